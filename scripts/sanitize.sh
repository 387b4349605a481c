#!/bin/bash
# usage (under gpurun): scripts/sanitize.sh   -- compute-sanitizer memcheck / racecheck / synccheck / initcheck over
# scripts/sanitize_workload.py; logs in gpurun_out/sanitize_<tool>.log, one summary line per tool on stdout
python scripts/sanitize_workload.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "workload fails without the tool"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check full"
  [ $tool = initcheck ] && extra="--track-unused-memory no"
  timeout 600 compute-sanitizer --tool $tool $extra --print-limit 50 python scripts/sanitize_workload.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' gpurun_out/sanitize_$tool.log | tr '\n' ' ') $(grep -c 'mismatches' gpurun_out/sanitize_$tool.log) checks, ok=$(grep -c 'sanitize workload ok' gpurun_out/sanitize_$tool.log)"
done
