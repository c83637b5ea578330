set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py > gpurun_out/c24_bench_n1.json 2> gpurun_out/c24_bench_n1.err; echo "bench rc $?"; tail -c 300 gpurun_out/c24_bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c24_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print('roof', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['traffic'], d['roofline'].get('intersection_phase',{}).get('frac'))
print('parity', d['parity_full_workload']['ok']); print('cpu', d['cpu_baseline']['value'])
for o in d['other_configs']: print(o['config'], o.get('value'), o.get('roofline',{}).get('frac'), o.get('parity_sample',{}).get('ok'), o.get('error'))
PY
