#!/bin/bash
# last GPU seconds of the round: smoke() and a short c3 bench line on the final HEAD
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py smoke > gpurun_out/r02l_smoke.log 2>&1; tail -1 gpurun_out/r02l_smoke.log | cut -c1-250
timeout 70 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_head.json 2> gpurun_out/bench_c3_head.err; cut -c1-260 gpurun_out/bench_c3_head.json
