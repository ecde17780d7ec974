#!/bin/bash
# ncu --set full of the north star's HBM-bound pooling kernels (blur-pool, GAP, SK pooled descriptor, 3x3/2 avg-pool)
mkdir -p gpurun_out
timeout 100 python tools/profile_pool.py > gpurun_out/r02j_pool_events.txt 2>&1; cat gpurun_out/r02j_pool_events.txt
ACNN_PROFILE_ONCE=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:"blurpool|avgpool|gap_bwd|image_reduce" -f -o /tmp/r02j_pool python tools/profile_pool.py > gpurun_out/ncu19.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02j_pool.ncu-rep --page raw --csv > gpurun_out/r02j_pool.raw.csv 2>/dev/null; wc -l gpurun_out/r02j_pool.raw.csv; tail -3 gpurun_out/ncu19.log
