# round-2 measurement set: GPU tests, stage timing, ncu launch list + full capture of the four headline kernels, full bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --kernels-only --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['stages'])"
export RB_BAND_ROUNDS=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02x.csv python bench.py --kernels-only --steps 2 --warmup 3 > gpurun_out/ncu_x.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band|k_diff2_slices|k_store_band|k_coarse_fused" -s 12 -c 4 -o gpurun_out/prof_r02x python bench.py --kernels-only --steps 2 --warmup 3 >> gpurun_out/ncu_x.log 2>&1
tail -2 gpurun_out/ncu_x.log | cut -c1-200
unset RB_BAND_ROUNDS
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02x.json 2> gpurun_out/bench_r02x.err; tail -c 400 gpurun_out/bench_r02x.json; tail -3 gpurun_out/bench_r02x.err
