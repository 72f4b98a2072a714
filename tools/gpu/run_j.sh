python -m pytest tests/test_adapter_cpp.py tests/test_particle_io.py -m gpu -x -q -s 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err; tail -c 6000 gpurun_out/bench_r02j.json; tail -5 gpurun_out/bench_r02j.err
