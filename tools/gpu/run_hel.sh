mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "symmetrise" 2>&1 | tail -6 > gpurun_out/run_hel.txt
cat gpurun_out/run_hel.txt
