# tile-major band arrays: parity, timing, ncu of the band kernels
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"; }
run RB_X=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x 2>&1 | tail -4
export RB_BAND_ROUNDS=1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band|k_diff2_slices|k_store_band" -s 9 -c 3 -o gpurun_out/prof_r02o python bench.py --kernels-only --steps 2 --warmup 3 > gpurun_out/ncu_o.log 2>&1
tail -2 gpurun_out/ncu_o.log | cut -c1-200
