# band-major local coarse pass with cell gathers: parity, timing, band counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x 2>&1 | tail -6
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"; }
run RB_X=0
run RB_COARSE_BANDS=2
run RB_COARSE_BANDS=6
run RB_COARSE_BANDS=8
run RB_COARSE_CELLS=0
