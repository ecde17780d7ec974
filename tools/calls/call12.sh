#!/bin/bash
# GPU call: dual batch-norm backward (bn_bwd_reduce2 / apply2): lockstep + e2e tests, then step timing
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_plan_gpu.py tests/test_e2e_gpu.py tests/test_parity_fp32_gpu.py -x -q > gpurun_out/tests12.log 2>&1; echo "rc=$?" >> gpurun_out/tests12.log
tail -3 gpurun_out/tests12.log
for i in 1 2; do for o in 0 1; do echo "== fuse_bn_pairs=$o run $i"; ACNN_FUSE_BN_PAIRS=$o timeout 120 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'])"; done; done > gpurun_out/fuse_bn_pairs.txt 2>&1
cat gpurun_out/fuse_bn_pairs.txt
