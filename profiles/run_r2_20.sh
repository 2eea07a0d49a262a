cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "many_periods or periodic or full_size" 2>&1 | tail -5
timeout 600 python profiles/prof_wrap_range.py 2>&1 | tail -5
