cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
REPS=4 timeout 900 python profiles/prof_stencil.py c5 1.25e8 f64 sparse,records > gpurun_out/r2_prof_c5.jsonl 2> gpurun_out/r2_prof_c5.err
cat gpurun_out/r2_prof_c5.jsonl; tail -3 gpurun_out/r2_prof_c5.err
