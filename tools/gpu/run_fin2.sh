# coarse kernel change: parity subset, stage timing, then the measurement set (launch list, full capture, bench line)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -q -x -k "local or left_and_right or scale or benchmark" 2>&1 | tail -3 | tee gpurun_out/parity_fin2.txt
python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])" | tee gpurun_out/stages_fin.txt
export RB_BAND_ROUNDS=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_fin.csv python bench.py --kernels-only --steps 2 --warmup 3 > gpurun_out/ncu_fin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band|k_diff2_slices|k_store_band|k_coarse_fused" -s 12 -c 4 -o gpurun_out/prof_r02_fin python bench.py --kernels-only --steps 2 --warmup 3 >> gpurun_out/ncu_fin.log 2>&1
unset RB_BAND_ROUNDS
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_fin.json 2> gpurun_out/bench_r02_fin.err; tail -c 300 gpurun_out/bench_r02_fin.json
