cd $GRAFT_REPO_ROOT
timeout 3000 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "promotes_with or (sweep and cubic2 and float32)" 2>&1 | grep -v "^ \|^$" | tail -30
