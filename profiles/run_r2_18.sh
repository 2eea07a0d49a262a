cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
REPS=2 timeout 600 python profiles/prof_stencil.py c4 1e8 f64 dense,sparse > gpurun_out/plain3.log 2>&1 &&
REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:stencil_ -c 4 -f -o gpurun_out/r02_stencil_c4 python profiles/prof_stencil.py c4 1e8 f64 dense,sparse > gpurun_out/ncu_c4s.log 2>&1
tail -2 gpurun_out/ncu_c4s.log
