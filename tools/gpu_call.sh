set -x
cd $GRAFT_REPO_ROOT
timeout 800 python -m pytest tests/test_gpu_more.py -x -q -k "c5_shape" 2>&1 | tail -3
