mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_particle_io.py tests/test_adapter_cpp.py -m gpu -q -x -k "prepare or star or refinement_iterations or adapter or cpp or feed" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --other-workloads 0 --ref-cuda-sample 0 --parity-sample 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_from_raw_images']['value'], d['e2e_from_mrc_stacks'].get('value'))"
