set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-configs > gpurun_out/c31_bench_n2.json 2> gpurun_out/c31_bench_n2.err; echo "bench rc $?"; tail -c 300 gpurun_out/c31_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c31_bench_n2.json').read().strip().splitlines()[-1])
m=d['multi_gpu']; print('N=2', d['value'], d['ms_per_step'], m['bytes_per_pair_over_nvlink'], m['all_rows_match_unsharded_call'], m['nccl_gather_rows_match'], m['oracle_check']['ok'], d['e2e']['ms_per_step'])
PY
