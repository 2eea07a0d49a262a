cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python profiles/prof_r02.py c5 c3 > gpurun_out/plain_prof2.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'eval_tile_kernel|stencil_sparse_kernel|stencil_dense_kernel' -c 12 -f -o gpurun_out/r02_kernels_b python profiles/prof_r02.py c5 c3 > gpurun_out/ncu_r02b.log 2>&1
tail -3 gpurun_out/ncu_r02b.log; cat gpurun_out/plain_prof2.log
