# per-class prior test, then the fused coarse kernel's gather variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prior_offsets" 2>&1 | tail -5
for g in 0 1; do for w in 16 8; do
  echo "== RB_FUSED_G256=$g RB_FUSED_WARPS=$w"
  RB_FUSED_G256=$g RB_FUSED_WARPS=$w python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages']['coarse'])"
done; done
RB_FUSED_G256=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "local" 2>&1 | tail -4
