mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_pool_local_search or gradient_refinement or backproject_posed or prepare_noise or beam_tilt or many_translations or always_cc or band_major_path" > gpurun_out/memcheck_r02.txt 2>&1
tail -8 gpurun_out/memcheck_r02.txt
