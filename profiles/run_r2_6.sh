cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_gpu_tests.log
tail -4 gpurun_out/r2_gpu_tests.log
for c in "c4 1e8 f64" "c4 1e8 f32" "c5 1.25e8 f64" "c5 1.25e8 f32" "c3 1e8 f64"; do
REPS=4 timeout 900 python profiles/prof_stencil.py $c dense,sparse,auto,full,records >> gpurun_out/r2_prof_all.jsonl 2>> gpurun_out/r2_prof_all.err
done
cat gpurun_out/r2_prof_all.jsonl; tail -3 gpurun_out/r2_prof_all.err
