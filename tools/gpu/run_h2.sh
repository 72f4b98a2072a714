mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
wc -l gpurun_out/bench_r02_2gpu.json; python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['n_gpus'], d['ms_per_step'], d['e2e'], d['allreduce_ms'], d['allreduce_parity'], d['scaling'], d['parity'])
"; tail -3 gpurun_out/bench_r02_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -c 400
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_adapter_cpp.py -m gpu -q 2>&1 | tail -3
