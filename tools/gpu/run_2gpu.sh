mkdir -p gpurun_out
(
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_adapter_cpp.py -q -x -m gpu 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['e2e'], d.get('allreduce_ms'), d.get('allreduce_parity'))"
) > gpurun_out/run_2gpu_fin.txt 2>&1
cat gpurun_out/run_2gpu_fin.txt
