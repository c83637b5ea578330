set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc $?" 
tail -5 gpurun_out/c1_pytest.log
timeout 400 python tools/ab.py C2 --reps 6 --out gpurun_out/c1_ab_c2.jsonl r1style tma_only hints_only default > gpurun_out/c1_ab_c2.log 2>&1; tail -6 gpurun_out/c1_ab_c2.log | cut -c1-600
timeout 200 python tools/phase_time.py C2 > gpurun_out/c1_phase_c2.log 2>&1; tail -24 gpurun_out/c1_phase_c2.log
timeout 500 python bench.py --steps 10 > gpurun_out/c1_bench_n1.json 2> gpurun_out/c1_bench_n1.err; echo "bench rc $?"; tail -c 600 gpurun_out/c1_bench_n1.err
timeout 300 python tools/check_config.py C3 20000000 0 > gpurun_out/c1_check_c3.log 2>&1; tail -4 gpurun_out/c1_check_c3.log
timeout 300 python tools/phase_time.py C3 20000000 > gpurun_out/c1_phase_c3.log 2>&1; tail -24 gpurun_out/c1_phase_c3.log
timeout 300 python tools/check_config.py C4 10000000 0 > gpurun_out/c1_check_c4.log 2>&1; tail -4 gpurun_out/c1_check_c4.log
timeout 300 python tools/phase_time.py C4 10000000 > gpurun_out/c1_phase_c4.log 2>&1; tail -24 gpurun_out/c1_phase_c4.log
