mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reconstruct or refinement_iterations" 2>&1 | tail -8
