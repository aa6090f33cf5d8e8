timeout 1200 python -m pytest tests/test_gpu_bestbasis.py -x -q 2>&1 | tail -4
timeout 300 python benchmarks/bench_paths.py --only jbb,lsdb 2>&1 | tee gpurun_out/bb_v3.jsonl
