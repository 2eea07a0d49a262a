cd $GRAFT_REPO_ROOT
timeout -s KILL 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout -s KILL 200 python profiles/prof_c1.py > gpurun_out/c1_fast.log 2>&1; head -5 gpurun_out/c1_fast.log
timeout -s KILL 300 python profiles/prof_stencil.py c4 1e8 f32 sparse,column,full > gpurun_out/col_c4_f32.log 2>&1; grep sorted gpurun_out/col_c4_f32.log
timeout -s KILL 300 python profiles/prof_stencil.py c4 1e8 f64 sparse,auto,column,full > gpurun_out/col_c4_auto.log 2>&1; grep sorted gpurun_out/col_c4_auto.log
timeout -s KILL 300 python profiles/prof_stencil.py c4 5e7 f64 sparse,column,full > gpurun_out/col_c4_5e7.log 2>&1; grep sorted gpurun_out/col_c4_5e7.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_col.json 2> gpurun_out/bench_col.err; tail -c 600 gpurun_out/bench_col.json
