#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run12.log
: > $L
run() { echo "=== $*" | tee -a $L; timeout 900 "$@" 2>&1 | tail -n 25 | tee -a $L; echo "rc=${PIPESTATUS[0]}" | tee -a $L; }
run python -m pytest tests/test_gpu_vit.py tests/test_gpu_ist_trunk.py tests/test_gpu_retrieval.py -q
run python -m pytest tests/test_gpu_chain_parity.py tests/test_gpu_surface.py -q
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; tail -c 600 gpurun_out/bench_r2h.json | tee -a $L
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2_v4.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-range > gpurun_out/bench_ncu.log 2>&1
