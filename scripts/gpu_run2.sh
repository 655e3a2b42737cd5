#!/bin/bash
# round-2 GPU session: new kernels first (each under its own timeout), then A/B timings
mkdir -p gpurun_out
L=gpurun_out/run2.log
: > $L
run() { echo "=== $*" | tee -a $L; timeout 600 "$@" 2>&1 | tail -n 40 | tee -a $L; echo "rc=${PIPESTATUS[0]}" | tee -a $L; }
run python -m pytest tests/test_gpu_retrieval.py -x -q
run python -m pytest tests/test_gpu_vit.py tests/test_gpu_ist_trunk.py -x -q
run python -m pytest tests/test_gpu_multi.py tests/test_gpu_surface.py -x -q
run python -m pytest tests/test_gpu_chain_parity.py -x -q -k crop_level
run python scripts/e2e_probe.py
GIGAPOSE_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; tail -c 1500 gpurun_out/bench_pdl0.json | tee -a $L
GIGAPOSE_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; tail -c 1500 gpurun_out/bench_pdl1.json | tee -a $L
GIGAPOSE_PDL=1 GIGAPOSE_SIM_PAIR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl1_pair1.json 2> gpurun_out/bench_pdl1_pair1.err; tail -c 1500 gpurun_out/bench_pdl1_pair1.json | tee -a $L
run python scripts/stress_sim.py --out gpurun_out/stress_sim_r2.json
