cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 3000 python -m pytest tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_cfg_tests.log
tail -25 gpurun_out/r2_cfg_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -2 gpurun_out/r2_bench_c4.err; cat gpurun_out/r2_bench_c4.json
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; tail -2 gpurun_out/r2_bench_c5.err; cat gpurun_out/r2_bench_c5.json
