#!/bin/bash
# One 8-GPU box: c4 at N = 8, c3 at N = 4, c2 at N = 2 and N = 1, back to back (gpurun --gpus 8 -- 'bash scripts/gpu_session_8gpu.sh').
# Add `--workload c5` to the first line for BASELINE configs[4].
mkdir -p gpurun_out
L=gpurun_out/run_n8b.log
: > $L
tr() { n=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@"; }
tr 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_c4_v2.json 2> gpurun_out/bench_n8_c4_v2.err; echo "c4 rc=$?" | tee -a $L; tail -c 1300 gpurun_out/bench_n8_c4_v2.json | tee -a $L
tr 4 --steps 10 --warmup 3 > gpurun_out/bench_n4_c3_v2.json 2> gpurun_out/bench_n4_c3_v2.err; echo "c3 rc=$?" | tee -a $L; tail -c 700 gpurun_out/bench_n4_c3_v2.json | tee -a $L
tr 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_c2_v3.json 2> gpurun_out/bench_n2_c2_v3.err; echo "c2 rc=$?" | tee -a $L; tail -c 700 gpurun_out/bench_n2_c2_v3.json | tee -a $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_samebox.json 2> gpurun_out/bench_n1_samebox.err; echo "n1 rc=$?" | tee -a $L; tail -c 500 gpurun_out/bench_n1_samebox.json | tee -a $L
