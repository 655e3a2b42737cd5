#!/bin/bash
# One 1-GPU session: every -m gpu test in its own process, smoke(), the bench line (+ CPU and eager-GPU baselines), the ncu
# launch list, one `--set full` capture of the similarity kernel, and the reference arm on the box's host cores.
# (gpurun --timeout 3000 -- 'bash scripts/gpu_session_1gpu.sh'; outputs land in gpurun_out/)
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash scripts/gpu_check.sh > /dev/null 2>&1
grep -E "^===|passed|failed|overall|skipped" gpurun_out/tests.log | paste - - | grep -v "1 passed" | head -40
grep -B5 -A30 "^E  " gpurun_out/tests.log | head -100
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --gpu-eager-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 4500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sim_search_pair --launch-skip 0 -c 1 -o gpurun_out/r02_sim_search_pair_final -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full.log 2>&1
(time timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err) 2>&1 | tail -3; tail -c 1500 gpurun_out/bench_reference_arm.json
