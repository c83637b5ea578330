set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/c15_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/c15_bench_n1.json 2> gpurun_out/c15_bench_n1.err; echo "bench rc $?"; tail -c 300 gpurun_out/c15_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/c15_bench_ref.json 2> gpurun_out/c15_bench_ref.err; echo "ref rc $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c15_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
print('roof', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['traffic'], d['roofline'].get('intersection_phase',{}).get('frac'))
print('parity', d['parity_full_workload']); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['target_100x'])
for o in d['other_configs']: print(o['config'], o.get('value'), o.get('roofline',{}).get('frac'), o.get('parity_sample',{}).get('ok'), o.get('error'))
r=json.loads(open('gpurun_out/c15_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['cores'], r['ms_per_step'])
PY
