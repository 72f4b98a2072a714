mkdir -p gpurun_out
(
for v in 4 3 2; do
echo "== RB_FUSED_G256=$v (quad; rows in flight: 4 -> 1, 3 -> 2, 2 -> 4)"
RB_FUSED_G256=$v python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"
done
echo "== tests (default = 2)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -x -k "local or 128px or left_and_right or pipelined or benchmark or scale" 2>&1 | tail -4
) > gpurun_out/run_quad_mlp.txt 2>&1
cat gpurun_out/run_quad_mlp.txt
