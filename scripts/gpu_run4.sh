#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash scripts/gpu_check.sh > /dev/null 2>&1
grep -E "^===|passed|failed|overall" gpurun_out/tests.log | paste - - | grep -v "1 passed" | head -40
timeout 300 python scripts/pdl_ab.py > gpurun_out/pdl_ab.log 2>&1; tail -8 gpurun_out/pdl_ab.log
timeout 300 python scripts/mlp_split_emulation.py > gpurun_out/mlp_emul.log 2>&1; tail -3 gpurun_out/mlp_emul.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 3500 gpurun_out/bench_r2b.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sim_search_pair --launch-skip 0 -c 1 -o gpurun_out/r02_sim_search_pair -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
