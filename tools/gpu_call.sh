set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/c23_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 --no-other-configs > gpurun_out/c23_bench_n2.json 2> gpurun_out/c23_bench_n2.err; echo "bench rc $?"; tail -c 400 gpurun_out/c23_bench_n2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c23_bench_n2.json').read().strip().splitlines()[-1])
    m=d['multi_gpu']; print('N=2 value', d['value'], 'ms', d['ms_per_step'], m['bytes_per_pair_over_nvlink']); print(m['fused_window']['ms_per_step'], m['scoring_only']['ms_per_step'], m['all_rows_match_unsharded_call'], m['nccl_gather_rows_match'], m['oracle_check']['ok']); print(m['fused_window']['what'])
except Exception as e: print('no json', e)
PY
