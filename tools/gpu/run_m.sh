# hypothesis check: with a coarse window small enough for the expanded cells to stay in L2, are two 32-byte loads faster than four 16-byte ones?
for cs in 64 84 106; do for g in 0 1; do
  echo "== coarse_size=$cs RB_FUSED_G256=$g"
  RB_BENCH_COARSE_SIZE=$cs RB_FUSED_G256=$g python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages']['coarse'])"
done; done
