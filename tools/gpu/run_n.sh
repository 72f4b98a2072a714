# TLB-aware layouts: coarse-window core of the x-pair copy, radius-sorted 4^3 blocks of the expanded reference
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"; }
run RB_X=0
run RB_BAND_PROJ_CTAS=2
run RB_BLOCK_SORT=0
run RB_COARSE_CORE=0
run RB_FUSED_G256=1
