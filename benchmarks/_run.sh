timeout 1200 python -m pytest tests/test_gpu_dwt.py -x -q -k "2d" 2>&1 | tail -8
P="wpd2d_f64_haar,wpd2d_f64_db4,wpd2d_f32_haar,wpd2d_f32_db4"
timeout 300 python benchmarks/bench_paths.py --only $P 2>&1 | tee gpurun_out/wpd2d_v2.jsonl
