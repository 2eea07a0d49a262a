cd $GRAFT_REPO_ROOT
timeout 600 python profiles/prof_wrap_range.py 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['random_order']['value'], d['functional_form']['ms_per_step'], d['gpu_launches'])"
