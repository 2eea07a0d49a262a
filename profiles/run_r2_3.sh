set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_stencil.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_stencil_tests.log
tail -5 gpurun_out/r2_stencil_tests.log
REPS=4 timeout 600 python profiles/prof_stencil.py c4 1e8 f64 dense,sparse,records > gpurun_out/r2_prof_c4.jsonl 2> gpurun_out/r2_prof_c4.err
cat gpurun_out/r2_prof_c4.jsonl; tail -3 gpurun_out/r2_prof_c4.err
REPS=4 timeout 600 python profiles/prof_stencil.py c3 1e8 f64 dense,sparse > gpurun_out/r2_prof_c3.jsonl 2> gpurun_out/r2_prof_c3.err
cat gpurun_out/r2_prof_c3.jsonl; tail -3 gpurun_out/r2_prof_c3.err
REPS=2 timeout 600 python profiles/prof_stencil.py c4 1e8 f64 dense > gpurun_out/plain.log 2>&1 &&
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:stencil_dense -c 2 -f -o gpurun_out/r2_stencil_b python profiles/prof_stencil.py c4 1e8 f64 dense > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
