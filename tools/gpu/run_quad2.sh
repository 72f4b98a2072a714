mkdir -p gpurun_out
export RB_COARSE_QUAD=1
ncu --set full --clock-control none --import-source on -k regex:"k_coarse_fused" -s 3 -c 1 -o gpurun_out/prof_r02_quad python bench.py --kernels-only --steps 1 --warmup 1 > gpurun_out/ncu_quad.log 2>&1
tail -2 gpurun_out/ncu_quad.log | cut -c1-200
