#!/bin/bash
# ncu evidence of round 2 (run on the GPU box from the repo root; writes gpurun_out/).  See /opt/skills/guides/B200_PROFILING.md.
set -x
O=gpurun_out
mkdir -p $O
tools/microbench/sfu_bench > $O/r02_sfu_microbench.json
# 1. launch list of two bench steps (per-launch device time; serialised, cold cache: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_ncu.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-overlap --no-streaming --no-torch-baseline --no-forward --no-diffuse > $O/r02_bench_under_ncu.json 2> /dev/null
# 2. full capture of the level-3 kernel as shipped (bulk staging, fixed-point exit) and with every iteration executed
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_w65x2 -s 2 -c 1 -f -o $O/r02_w65x2_bulk python tools/run_l3_once.py > /dev/null 2>&1
PATS_L3_FP_EXIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_w65x2 -s 2 -c 1 -f -o $O/r02_w65x2_bulk_noexit python tools/run_l3_once.py > /dev/null 2>&1
PATS_L3_FP_EXIT=0 PATS_L3_BULK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_w65x2 -s 2 -c 1 -f -o $O/r02_w65x2_direct_noexit python tools/run_l3_once.py > /dev/null 2>&1
# 3. full capture of the streaming (grid-cooperative) kernel on BASELINE.json configs[2]
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_grid_kernel -s 1 -c 1 -f -o $O/r02_grid_streaming python tools/run_streaming_once.py > /dev/null 2>&1
ls -la $O/*.ncu-rep
