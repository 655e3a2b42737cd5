#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tests.log
bash scripts/gpu_check.sh > /dev/null 2>&1
grep -E "^===|passed|failed|overall|skipped" gpurun_out/tests.log | paste - - | grep -v "1 passed" | head -40
grep -A40 "FAILED\|Error" gpurun_out/tests.log | head -120
timeout 300 python scripts/mlp_split_emulation.py > gpurun_out/mlp_emul.log 2>&1; tail -2 gpurun_out/mlp_emul.log | cut -c1-1500
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; tail -c 1500 gpurun_out/bench_r2e.json
