mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"; }
run RB_BAND_CHUNK_MIN=48
run RB_BAND_CHUNK_MIN=64
run RB_BAND_CHUNK_MIN=96
run RB_BAND_CHUNK_MIN=128
run RB_BAND_CHUNK_MIN=64 RB_BAND_STORE_CHUNK_MIN=32
run RB_BAND_CHUNK_MIN=64 RB_BAND_STORE_CHUNK_MIN=8
