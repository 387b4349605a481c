#!/bin/bash
# usage: scripts/profile.sh <tag>   -- run under gpurun: plain kernel bench, then ncu --set full on one fused launch
set -e
TAG=${1:-prof}
python scripts/kbench.py --cfg 2 --steps 5 > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -o gpurun_out/$TAG \
    python scripts/kbench.py --cfg 2 --steps 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
