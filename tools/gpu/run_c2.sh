mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "corrections_stage or euler" 2>&1 | tail -8
