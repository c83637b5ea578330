set -x
cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/c19_bench_n$N.json 2> gpurun_out/c19_bench_n$N.err; echo "bench rc $?"; tail -c 600 gpurun_out/c19_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c19_bench_n$N.json').read().strip().splitlines()[-1])
    m=d['multi_gpu']; print('N=$N value', d['value'], 'ms', d['ms_per_step'], d['clocks']); print(json.dumps(m['per_rank_own_kernels_ms'])); print(m['fused_window']['ms_per_step'], m['scoring_only']['ms_per_step'], m['nccl_gather']['ms_per_step'], m['all_rows_match_unsharded_call'], m['nccl_gather_rows_match'], m['oracle_check']['ok'])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    for o in d['other_configs']: print(o['config'], o.get('value'), o.get('ms_per_step'), o.get('roofline',{}).get('frac'), o.get('parity_sample',{}).get('ok'), o.get('error'))
except Exception as e: print('no json', e)
PY
