python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err; tail -c 2500 gpurun_out/bench_r02i.json; tail -3 gpurun_out/bench_r02i.err
export RB_BAND_ROUNDS=1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02i.csv python bench.py --kernels-only --steps 2 --warmup 3 > gpurun_out/ncu_i.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_project_band|k_diff2_slices|k_store_band|k_coarse_fused" -s 12 -c 4 -o gpurun_out/prof_r02i python bench.py --kernels-only --steps 2 --warmup 3 >> gpurun_out/ncu_i.log 2>&1
tail -2 gpurun_out/ncu_i.log | cut -c1-200
