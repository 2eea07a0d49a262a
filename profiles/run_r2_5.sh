cd $GRAFT_REPO_ROOT
echo "== sparse only"; IPX_DEBUG=1 CUDA_LAUNCH_BLOCKING=1 REPS=2 timeout 600 python profiles/prof_stencil.py c5 2e7 f64 sparse 2>&1 | grep "ipx stencil\|cfg" | head
