set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_hop3.py tests/test_gpu_eval.py tests/test_gpu_build.py -x -q -s 2>&1 | tail -8
