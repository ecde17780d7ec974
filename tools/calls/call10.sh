#!/bin/bash
# GPU call: wgrad GEMMs on a second stream inside the captured step (env switch), two runs each
mkdir -p gpurun_out
for i in 1 2; do for o in 0 1; do echo "== overlap_wgrad=$o run $i"; ACNN_OVERLAP_WGRAD=$o timeout 120 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; done; done > gpurun_out/overlap_wgrad.txt 2>&1
cat gpurun_out/overlap_wgrad.txt
