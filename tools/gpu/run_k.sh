# two GPUs: adapter C++ multi-rank program over real NCCL, the Python 2-GPU reduction test, the per-class prior test, and the 2-rank bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_adapter_cpp.py tests/test_gpu_multi.py tests/test_particle_io.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/multi_r02k_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "2d or class2d or prior" 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --other-workloads 0 > gpurun_out/bench_r02k_2gpu.json 2> gpurun_out/bench_r02k_2gpu.err
wc -l gpurun_out/bench_r02k_2gpu.json; tail -c 700 gpurun_out/bench_r02k_2gpu.json; tail -3 gpurun_out/bench_r02k_2gpu.err
