mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prepare or two_device or star_and_mrc" 2>&1 | tail -4
