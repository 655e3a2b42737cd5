#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run9.log
: > $L
run() { echo "=== $*" | tee -a $L; timeout 900 "$@" 2>&1 | tail -n 30 | tee -a $L; echo "rc=${PIPESTATUS[0]}" | tee -a $L; }
run python -m pytest tests/test_gpu_retrieval.py -q
run python -m pytest tests/test_gpu_chain_parity.py -q -k crop_level
run python -m pytest tests/test_gpu_multi.py tests/test_gpu_surface.py -q
run python scripts/mlp_split_emulation.py
