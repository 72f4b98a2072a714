python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band or slice_cache or local_search or capacity or 2d_class or many_trans or global" 2>&1 | tail -3
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "256_local or class3d" 2>&1 | tail -3
K="python bench.py --kernels-only --steps 5 --warmup 3"
sel() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']; print('$1', d['value'], {k:s[k] for k in ('coarse','fine','store','fine_prep','fine_project','fine_diff2','store_list','store_band','total')})
"; }
$K 2>&1 | sel base
RB_BAND_PREFETCH=0 $K 2>&1 | sel noprefetch_proj
RB_BAND_STORE_PREFETCH=0 $K 2>&1 | sel noprefetch_store
RB_BAND_PREFETCH=4 RB_BAND_STORE_PREFETCH=4 $K 2>&1 | sel prefetch4
RB_BAND_PREFETCH=1 RB_BAND_STORE_PREFETCH=1 $K 2>&1 | sel prefetch1
RB_BAND_STORE_CHUNK_MIN=8 $K 2>&1 | sel storechunk8
RB_BAND_STORE_CHUNK_MIN=32 $K 2>&1 | sel storechunk32
export RB_BAND_ROUNDS=1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band|k_diff2_slices|k_store_band" -s 9 -c 3 -o gpurun_out/prof_r02e $K > gpurun_out/ncu_e.log 2>&1
tail -1 gpurun_out/ncu_e.log | cut -c1-300
