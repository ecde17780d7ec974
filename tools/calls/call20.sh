#!/bin/bash
# GAP forward with the channels of an image split over 4 CTAs: op-by-op lockstep + e2e tests, event timings
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_plan_gpu.py tests/test_e2e_gpu.py -q -x > gpurun_out/r02k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_tests.log; tail -4 gpurun_out/r02k_tests.log
timeout 60 python tools/profile_pool.py > gpurun_out/r02k_pool_events.txt 2>&1; grep "gap" gpurun_out/r02k_pool_events.txt
