cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 3000 python -m pytest tests/test_gpu_stencil.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -5
for c in "c4 1e8 f64" "c3 1e8 f64" "c3 2.5e8 f64"; do
REPS=4 timeout 900 python profiles/prof_stencil.py $c dense,sparse,full 2>&1 | grep cfg
done
