#!/bin/bash
# bench line + ncu launch list of ONE resident step (+ optional full capture of one kernel: KERNEL=regex) (1 GPU)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1
if [ -n "$KERNEL" ]; then
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$KERNEL --launch-skip ${KSKIP:-0} -c ${KCOUNT:-1} -o gpurun_out/prof_$KERNEL -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu_full.log 2>&1
fi
ls -la gpurun_out
