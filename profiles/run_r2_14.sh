cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python profiles/prof_c1.py 2>&1 | head -5
timeout 3000 python -m pytest tests -q -m gpu 2>&1 | tail -5
