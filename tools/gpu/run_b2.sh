mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages']['coarse'], d['stages']['total'])"; }
run RB_PIX_TILE=4x4
run RB_PIX_TILE=8x2
run RB_PIX_TILE=2x8
run RB_PIX_TILE=4x2
run RB_PIX_TILE=2x4
run RB_PIX_TILE=8x4
run RB_PIX_TILE=4x8
run RB_PIX_TILE=16x1
