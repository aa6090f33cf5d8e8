timeout 1200 python -m pytest tests/test_gpu_bestbasis.py tests/test_gpu_dwt.py -x -q -k "not 2d" 2>&1 | tail -4
P="wpdall_f64_db4,wpdall_f64_sym8,wpdall_f32_db4,jbb,lsdb"
timeout 300 python benchmarks/bench_paths.py --only $P 2>&1 | tee gpurun_out/bb_v2.jsonl
