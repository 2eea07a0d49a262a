cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.log
tail -6 gpurun_out/r2_gpu_tests.log
