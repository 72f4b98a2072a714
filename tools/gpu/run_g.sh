K="python bench.py --kernels-only --steps 5 --warmup 3"
sel() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']; print('$1', d['value'], {k:s[k] for k in ('coarse','fine','store','fine_prep','fine_project','fine_diff2','store_list','store_band','total')})
"; }
$K 2>&1 | sel l2_128
RB_BAND_L2_128=0 $K 2>&1 | sel plain
RB_BAND_L2_128=1 RB_BAND_CHUNK_MIN=32 $K 2>&1 | sel l2_128_chunk32
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band or local_search" 2>&1 | tail -2
