#!/bin/bash
# 2-GPU call: NCCL data-parallel test against the oracle, and the 2-GPU bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -x -q -s > gpurun_out/r02_dp_nccl_2gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dp_nccl_2gpu_final.log; tail -4 gpurun_out/r02_dp_nccl_2gpu_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_c3_n2_final.json 2> gpurun_out/bench_c3_n2_final.err; tail -1 gpurun_out/bench_c3_n2_final.json | cut -c1-300
