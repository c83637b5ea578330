set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err; echo "bench2 rc $?"; tail -c 1500 gpurun_out/c7_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c7_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 value', d['value'], 'ms', d['ms_per_step']); print(json.dumps(d['multi_gpu'], indent=1)[:2500]); print(d['e2e']['value'], d['clocks'])
    print(json.dumps([{k:v for k,v in o.items() if k!='graph'} for o in d['other_configs']], indent=1)[:1500])
except Exception as e: print('no json', e)
PY
