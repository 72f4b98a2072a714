# padded block accumulator: parity (all GPU tests), timing of the E-step and of the posed back-projection
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"; }
run RB_X=0
run RB_BP_BLOCKS=0
for b in 1 0; do
  echo "== posed RB_BP_BLOCKS=$b"
  RB_BP_BLOCKS=$b python bench.py --workload reconstruct_256 --steps 5 --warmup 3 --other-workloads 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
done
