# band-major posed back-projection: parity + timing against the image-major kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "posed or reconstruct_256" 2>&1 | tail -4
for b in 1; do for cm in 8; do
  echo "== RB_POSED_BAND=$b RB_POSED_CHUNK_MIN=$cm"
  RB_POSED_BAND=$b RB_POSED_CHUNK_MIN=$cm python bench.py --workload reconstruct_256 --steps 5 --warmup 3 --other-workloads 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
done; done
ncu --set full --clock-control none --import-source on -k regex:"k_posed" -s 6 -c 2 -o gpurun_out/prof_r02p python bench.py --workload reconstruct_256 --steps 2 --warmup 3 --other-workloads 0 --kernels-only > gpurun_out/ncu_p.log 2>&1
tail -2 gpurun_out/ncu_p.log | cut -c1-200
