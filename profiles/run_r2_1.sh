set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests/test_gpu_stencil.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_stencil_tests.log
tail -5 gpurun_out/r2_stencil_tests.log
timeout 600 python profiles/prof_stencil.py c4 1e8 f64 > gpurun_out/r2_prof_c4.jsonl 2> gpurun_out/r2_prof_c4.err
cat gpurun_out/r2_prof_c4.jsonl; tail -3 gpurun_out/r2_prof_c4.err
timeout 600 python profiles/prof_stencil.py c3 1e8 f64 > gpurun_out/r2_prof_c3.jsonl 2> gpurun_out/r2_prof_c3.err
cat gpurun_out/r2_prof_c3.jsonl; tail -3 gpurun_out/r2_prof_c3.err
