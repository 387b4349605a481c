#!/bin/bash
# usage: scripts/profile.sh <tag>   -- run under gpurun: plain bench, then ncu --set full on the fused kernel
set -e
TAG=${1:-prof}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -o gpurun_out/$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
