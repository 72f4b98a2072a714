# multi-GPU: strong scaling (fixed 64 pools from a shared queue) and the K = 8 classification workload; N from $1 (default: all visible)
mkdir -p gpurun_out
NG=${1:-$(nvidia-smi -L | wc -l)}
tr() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" 2>>gpurun_out/run_w.err; }
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  echo "== strong scaling refine3d_256_local, 64 pools, N=$n"
  tr $n --kernels-only --scaling strong --total-pools 64 --steps 5 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['strong_scaling'], d['allreduce_ms'])"
done
for n in 1 $NG; do
  echo "== weak scaling class3d_256_global_k8, N=$n"
  tr $n --kernels-only --workload class3d_256_global_k8 --steps 5 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages'], d['allreduce_ms'], d['allreduce_parity'])"
  [ $NG = 1 ] && break
done
tail -5 gpurun_out/run_w.err
