cd $GRAFT_REPO_ROOT
timeout -s KILL 120 python profiles/col_small.py > gpurun_out/col_small.log 2>&1; echo rc=$? >> gpurun_out/col_small.log; tail -3 gpurun_out/col_small.log
if grep -q "rc=0" gpurun_out/col_small.log && ! grep -q "equal False" gpurun_out/col_small.log; then
  timeout -s KILL 400 python profiles/prof_stencil.py c4 1e8 f64 column,full 2>&1 | grep "sorted\|Error\|error" | tail -3
  timeout -s KILL 900 python -m pytest tests/test_gpu_stencil.py -x -q 2>&1 | tail -3
  REPS=2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:stencil_column --launch-skip 1 --launch-count 1 -o gpurun_out/col_v10 -f python profiles/prof_stencil.py c4 1e8 f64 column > gpurun_out/col_v10_ncu.log 2>&1; tail -2 gpurun_out/col_v10_ncu.log
fi
