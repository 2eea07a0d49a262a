cd $GRAFT_REPO_ROOT
timeout 600 python profiles/prof_c1.py 2>&1 | head -60
