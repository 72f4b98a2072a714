K="python bench.py --kernels-only --steps 5 --warmup 3"
sel() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']; print('$1', d['value'], {k:s[k] for k in ('coarse','fine','store','fine_project','fine_diff2','store_band','total')})
"; }
$K 2>&1 | sel off
RB_BAND_PREFETCH=2 $K 2>&1 | sel pf2
RB_BAND_PREFETCH=4 $K 2>&1 | sel pf4
RB_BAND_PREFETCH=8 $K 2>&1 | sel pf8
RB_BAND_PREFETCH=4 RB_BAND_STORE_PREFETCH=4 $K 2>&1 | sel pf4_store4
RB_BAND_PREFETCH=4 RB_BAND_L2_128=0 $K 2>&1 | sel pf4_nol2hint
export RB_BAND_ROUNDS=1 RB_BAND_PREFETCH=4
ncu --set full --clock-control none -k regex:"k_project_band" -s 3 -c 1 -o gpurun_out/prof_r02h $K > gpurun_out/ncu_h.log 2>&1
