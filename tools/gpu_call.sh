set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/e2e_time.py 2>&1 | tail -14
