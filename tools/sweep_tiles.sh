#!/bin/bash
# pixel-list tiling A/B on the headline workload: RB_PIX_TILE (coarse list), RB_PIX_TILE_F (store-stage list)
for t in 16x1 4x4 4x2 8x2 2x8; do
  RB_PIX_TILE_F=$t python bench.py --steps 5 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); s = d['stages']
print('fine-list $t', 'value', d['value'], 'coarse', s['coarse']['ms'], 'store', s['store']['ms'], 'total', s['total']['ms'])"
done
