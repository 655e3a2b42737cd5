#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash scripts/gpu_check.sh > /dev/null 2>&1
grep -E "^===|passed|failed|overall|skipped" gpurun_out/tests.log | paste - - | grep -v "1 passed" | head -40
grep -B5 -A30 "^E  " gpurun_out/tests.log | head -150
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -c 3800 gpurun_out/bench_r2f.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2_v2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sim_search_pair --launch-skip 0 -c 1 -o gpurun_out/r02_sim_search_pair_v2 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:vit_gemm_kernel --launch-skip 30 -c 4 -o gpurun_out/r02_vit_gemm_pair -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full2.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:attention_tc --launch-skip 3 -c 1 -o gpurun_out/r02_attention -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full3.log 2>&1
timeout 600 python scripts/stress_sim.py --out gpurun_out/stress_sim_r2b.json > gpurun_out/stress.log 2>&1; tail -c 1500 gpurun_out/stress.log
ls -la gpurun_out | tail -12
