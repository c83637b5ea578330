set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 2 --config C5 --steps 3 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c14_bench_c5_n2.json 2> gpurun_out/c14_bench_c5_n2.err; echo "bench rc $?"; tail -c 1500 gpurun_out/c14_bench_c5_n2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c14_bench_c5_n2.json').read().strip().splitlines()[-1])
    print('C5 N=2 value', d['value'], 'ms', d['ms_per_step']); print(json.dumps(d['multi_gpu'])[:1500]); r=d['roofline']; print(r['frac'], r['kernel_ms'], r['business_kernel']['kernel_ms'], r['launch'])
except Exception as e: print('no json', e)
PY
