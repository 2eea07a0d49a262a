cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python profiles/bench_configs.py c1 c2 c3 c4 c5 > gpurun_out/configs_r02.jsonl 2> gpurun_out/configs_r02.err; tail -2 gpurun_out/configs_r02.err
cat gpurun_out/configs_r02.jsonl | cut -c1-600
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 600 python profiles/prof_r02.py > gpurun_out/plain_prof.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'eval_tile_kernel|stencil_sparse_kernel|stencil_dense_kernel' -c 8 -f -o gpurun_out/r02_kernels python profiles/prof_r02.py > gpurun_out/ncu_r02.log 2>&1
tail -3 gpurun_out/ncu_r02.log; cat gpurun_out/plain_prof.log
