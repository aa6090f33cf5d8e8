timeout 1200 python -m pytest tests/test_gpu_dwt.py -x -q -k "2d" 2>&1 | tail -4
P="wpd2d_f64_haar,wpd2d_f64_db4,wpd2d_f32_haar,wpd2d_f32_db4"
timeout 300 python benchmarks/bench_paths.py --only $P 2>&1 | tee gpurun_out/wpd2d_v3.jsonl
ncu --set full --clock-control none --import-source on -k regex:wpd1d_tma -s 3 -c 1 -o gpurun_out/prof_wpd1d_tma_full_r1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_wpd_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wpd2d_tile -s 3 -c 1 -o gpurun_out/prof_wpd2d_tile_r1c python benchmarks/bench_paths.py --only wpd2d_f64_db4 --scale 0.125 > gpurun_out/ncu_2dt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wpd2d_block -s 1 -c 1 -o gpurun_out/prof_wpd2d_block_r1c python benchmarks/bench_paths.py --only wpd2d_f64_db4 --scale 0.125 > gpurun_out/ncu_2db.log 2>&1
