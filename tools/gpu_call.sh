set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/ab.py C2 --reps 6 --out gpurun_out/c29_ab_c2.jsonl default 2>&1 | tail -2 | cut -c1-1500
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
