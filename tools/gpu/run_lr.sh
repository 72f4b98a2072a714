mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "left_and_right or coarse_euler or test_pool_local or test_pool_global or 128px" 2>&1 | tail -15 > gpurun_out/run_lr.txt
cat gpurun_out/run_lr.txt
