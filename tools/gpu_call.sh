set -x
cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --config C3 --pairs 20000000 --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_score_(side|light)" -s 12 -c 2 -o gpurun_out/prof_r02_b_c3 python bench.py --config C3 --pairs 20000000 --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/c16_ncu.log 2>&1
tail -3 gpurun_out/c16_ncu.log; tail -c 600 gpurun_out/c16_plain.log
