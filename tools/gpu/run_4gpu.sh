mkdir -p gpurun_out
(
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('weak', d['n_gpus'], d['value'], d['e2e']['value'], d.get('allreduce_ms'), d.get('allreduce_parity'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong --total-pools 64 --kernels-only 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('strong', d['n_gpus'], d['value'], d.get('strong_scaling'))"
) > gpurun_out/run_4gpu_fin.txt 2>&1
cat gpurun_out/run_4gpu_fin.txt
