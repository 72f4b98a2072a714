mkdir -p gpurun_out
(
for o in 0 1; do for v in 0 1; do
echo "== RB_COARSE_ORDER=$o RB_COARSE_QUAD=$v"
RB_COARSE_ORDER=$o RB_COARSE_QUAD=$v python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"
done; done
echo "== tests with RB_COARSE_QUAD=1 (order on)"
RB_COARSE_QUAD=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "local or 128px or left_and_right or pipelined" 2>&1 | tail -4
) > gpurun_out/run_quad.txt 2>&1
cat gpurun_out/run_quad.txt
