mkdir -p gpurun_out
(
for w in 16 8; do
echo "== RB_FUSED_WARPS=$w"
RB_FUSED_WARPS=$w python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"
done
) > gpurun_out/run_npw.txt 2>&1
cat gpurun_out/run_npw.txt
