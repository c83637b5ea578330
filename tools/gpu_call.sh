set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/ab.py C2 --reps 6 --out gpurun_out/c25_ab_c2.jsonl default 2>&1 | tail -2 | cut -c1-1200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py tests/test_gpu_paths.py -x -q 2>&1 | tail -3
