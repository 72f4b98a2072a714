mkdir -p gpurun_out
python bench.py --workload reconstruct_256 --steps 5 --warmup 3 --other-workloads 0 2>gpurun_out/run_a2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['e2e_from_raw_images'], d['roofline']['frac'])"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "posed" 2>&1 | tail -2
