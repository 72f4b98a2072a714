mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages']['coarse'], d['stages']['total'])"; }
run RB_VARIANT=stages3
