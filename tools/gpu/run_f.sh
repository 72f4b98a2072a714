python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band or slice_cache or local_search or capacity or 2d_class or many_trans or global" 2>&1 | tail -3
K="python bench.py --kernels-only --steps 5 --warmup 3"
sel() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']; print('$1', d['value'], {k:s[k] for k in ('coarse','fine','store','fine_prep','fine_project','fine_diff2','store_list','store_band','total')})
"; }
$K 2>&1 | sel base
RB_BAND_PROJ_CTAS=2 $K 2>&1 | sel projctas2
RB_BAND_PROJ_CTAS=4 $K 2>&1 | sel projctas4
RB_BAND_CHUNK_MIN=32 $K 2>&1 | sel chunk32
export RB_BAND_ROUNDS=1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band" -s 3 -c 1 -o gpurun_out/prof_r02f $K > gpurun_out/ncu_f.log 2>&1
tail -1 gpurun_out/ncu_f.log | cut -c1-300
