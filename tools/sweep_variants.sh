#!/bin/bash
# A/B of the kernel variants selectable by environment (coarse x-pair layout, fine/store prefetch depth): stage times of bench.py
run() { env "$@" python bench.py --steps 5 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
s = d['stages']
print('$*', 'value', d['value'], 'e2e', d['e2e']['value'], 'coarse', s['coarse']['ms'], 'fine', s['fine']['ms'], 'store', s['store']['ms'], 'total', s['total']['ms'])
"; }
run RB_COARSE_XP=0
run RB_COARSE_XP=1
for v in 1 2 3 4 5; do run RB_FINE_VARIANT=$v; done
for v in 0 1 2 3 4; do run RB_STORE_VARIANT=$v; done
