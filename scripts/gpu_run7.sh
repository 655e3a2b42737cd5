#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run7.log
: > $L
run() { echo "=== $*" | tee -a $L; timeout 900 "$@" 2>&1 | tail -n 45 | tee -a $L; echo "rc=${PIPESTATUS[0]}" | tee -a $L; }
run python -m pytest tests/test_gpu_retrieval.py -x -q
run python -m pytest tests/test_gpu_chain_parity.py -x -q
run python -m pytest tests/test_gpu_multi.py tests/test_gpu_surface.py tests/test_gpu_fullsize.py tests/test_gpu_edge.py -x -q
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 1200 gpurun_out/bench_r2d.json | tee -a $L
GIGAPOSE_MLP_SIMT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_simt.json 2> gpurun_out/bench_r2d_simt.err; tail -c 400 gpurun_out/bench_r2d_simt.json | tee -a $L
