mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_adapter_cpp.py -q -x -m gpu 2>&1 | tail -40; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "left_and_right or noise_filled or prepare" 2>&1 | tail -8) > gpurun_out/run_ad.txt 2>&1
cat gpurun_out/run_ad.txt
