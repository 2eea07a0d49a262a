cd $GRAFT_REPO_ROOT
timeout 3000 python -m pytest tests/test_gpu_ad.py -x -q -m gpu 2>&1 | tail -40
