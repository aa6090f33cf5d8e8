timeout 1500 python -m pytest tests/test_gpu_rwt.py tests/test_golden_fixtures.py -x -q 2>&1 | tail -4
timeout 600 python benchmarks/bench_paths.py --only iswt 2>&1 | tee gpurun_out/general_v2.jsonl | cut -c1-220
