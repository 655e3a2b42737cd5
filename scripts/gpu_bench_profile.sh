#!/bin/bash
# bench line + ncu launch list of ONE resident step + full capture of the similarity kernel (1 GPU)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:sim_search -c 1 -o gpurun_out/prof_sim -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out
