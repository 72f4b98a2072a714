mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_adapter_cpp.py -q -x -m gpu 2>&1 | tail -30; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "only_classify or prepare" 2>&1 | tail -5) > gpurun_out/run_skip.txt 2>&1
cat gpurun_out/run_skip.txt
