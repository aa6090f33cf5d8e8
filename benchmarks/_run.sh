timeout 1200 python -m pytest tests/test_gpu_dwt.py -x -q -k "dwtall" 2>&1 | tail -8
timeout 300 python benchmarks/bench_paths.py --only bb,jbb 2>&1 | cut -c1-250 | tee gpurun_out/bb_bench.jsonl
timeout 300 python benchmarks/bench_paths.py --cpu --only cpu 2>&1 | cut -c1-250 | tee gpurun_out/cpu_port.jsonl
