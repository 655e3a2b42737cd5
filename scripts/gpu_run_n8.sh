#!/bin/bash
# 8-GPU box: c4 (N=8), c5 stress (N=8), c3 (N=4), c2 (N=2), then the CUDA-graph variant of c4 under a tight timeout
mkdir -p gpurun_out
L=gpurun_out/run_n8.log
: > $L
nvidia-smi -L | tee -a $L
tr() { n=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@"; }
tr 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_c4.json 2> gpurun_out/bench_n8_c4.err; echo "c4 rc=$?" | tee -a $L; tail -c 2200 gpurun_out/bench_n8_c4.json | tee -a $L
tr 8 --steps 5 --warmup 3 --workload c5 > gpurun_out/bench_n8_c5.json 2> gpurun_out/bench_n8_c5.err; echo "c5 rc=$?" | tee -a $L; tail -c 2200 gpurun_out/bench_n8_c5.json | tee -a $L; tail -5 gpurun_out/bench_n8_c5.err | tee -a $L
tr 4 --steps 10 --warmup 3 > gpurun_out/bench_n4_c3.json 2> gpurun_out/bench_n4_c3.err; echo "c3 rc=$?" | tee -a $L; tail -c 900 gpurun_out/bench_n4_c3.json | tee -a $L
tr 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_c2.json 2> gpurun_out/bench_n2_c2.err; echo "c2 rc=$?" | tee -a $L; tail -c 900 gpurun_out/bench_n2_c2.json | tee -a $L
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29877 bench.py --gpus 8 --steps 10 --warmup 3 --multi-gpu-graph > gpurun_out/bench_n8_c4_graph.json 2> gpurun_out/bench_n8_c4_graph.err; echo "c4 graph rc=$?" | tee -a $L; tail -c 700 gpurun_out/bench_n8_c4_graph.json | tee -a $L; tail -3 gpurun_out/bench_n8_c4_graph.err | tee -a $L
