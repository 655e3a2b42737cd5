#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run5.log
: > $L
run() { echo "=== $*" | tee -a $L; timeout 600 "$@" 2>&1 | tail -n 30 | tee -a $L; echo "rc=${PIPESTATUS[0]}" | tee -a $L; }
run python -m pytest tests/test_gpu_vit.py -x -q
run python -m pytest tests/test_gpu_chain_parity.py -x -q -k crop_level
run python -m pytest tests/test_gpu_surface.py tests/test_gpu_retrieval.py tests/test_gpu_fullsize.py -x -q
run python scripts/pdl_ab.py
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 1800 gpurun_out/bench_r2c.json | tee -a $L
