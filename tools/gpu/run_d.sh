python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band or slice_cache or local_search or capacity or 2d_class or many_trans" 2>&1 | tail -3
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "256_local or class3d" 2>&1 | tail -3
K="python bench.py --kernels-only --steps 5 --warmup 3"
sel() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']; print('$1', d['value'], {k:s[k] for k in ('coarse','fine','store','fine_prep','fine_project','fine_diff2','store_list','store_band','total')})
"; }
$K 2>&1 | sel base
RB_BAND_CHUNK_MIN=8 $K 2>&1 | sel chunk8
RB_BAND_CHUNK_MIN=32 $K 2>&1 | sel chunk32
RB_BAND_STORE_CTAS=2 $K 2>&1 | sel storectas2
RB_BAND_STORE_CHUNK_MIN=16 $K 2>&1 | sel storechunk16
RB_BAND_TABLES=0 $K 2>&1 | sel notables
export RB_BAND_ROUNDS=1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band|k_diff2_slices|k_store_band" -s 9 -c 3 -o gpurun_out/prof_r02d $K > gpurun_out/ncu_d.log 2>&1
tail -1 gpurun_out/ncu_d.log | cut -c1-300
