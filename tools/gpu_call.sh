set -x
cd $GRAFT_REPO_ROOT
BLP_HUB_MIN_DEG=915 timeout 300 python tools/ab.py C2 --reps 6 --out gpurun_out/c5_ab_c2.jsonl prev default 2>&1 | tail -3 | cut -c1-700
timeout 300 python tools/ab.py C2 --reps 6 --out gpurun_out/c5_ab_c2.jsonl default 2>&1 | tail -2 | cut -c1-700
timeout 300 python tools/ab.py C3:20000000 --reps 3 --out gpurun_out/c5_ab_c3.jsonl prev default 2>&1 | tail -3 | cut -c1-700
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -x -q 2>&1 | tail -3
