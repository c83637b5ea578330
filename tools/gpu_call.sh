set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 > gpurun_out/c4_bench_n1.json 2> gpurun_out/c4_bench_n1.err; echo "bench rc $?"; tail -c 600 gpurun_out/c4_bench_n1.err
for h in 700 850 1000 1200; do BLP_HUB_MIN_DEG=$h timeout 300 python tools/ab.py C2 --reps 5 --out gpurun_out/c4_ab_c2_hub.jsonl default 2>&1 | tail -2 | cut -c1-700; done
for h in 3000 4500; do BLP_HUB_MIN_DEG=$h timeout 300 python tools/ab.py C3:20000000 --reps 3 --out gpurun_out/c4_ab_c3_hub.jsonl default 2>&1 | tail -2 | cut -c1-700; done
timeout 300 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/c4_launches_c2.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c4_ncu1.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c4_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_score_(side|light)" -s 12 -c 4 -o gpurun_out/prof_r02_a python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c4_ncu2.log 2>&1
tail -3 gpurun_out/c4_ncu2.log
