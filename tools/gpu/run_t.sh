# fused local coarse kernel: more than 32 translations (3 passes), CC criterion; racecheck of a small pool
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "many_translations or always_cc or firstiter_cc or local_search" 2>&1 | tail -4
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_pool_local_search and 1" > gpurun_out/racecheck_r02.txt 2>&1
tail -6 gpurun_out/racecheck_r02.txt
