export MAPRAST_LIB=maprasterization.jl_b200/variants/libmaprast_sched.so
run() { python scripts/kbench.py --cfg $1 --lines $2 --steps 12 | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg $1 lines $2 F0=$MAPRAST_F0 TAILS=$MAPRAST_TAILS', round(d['ms_min'],4), round(d['ms_med'],4), d['digest'])"; }
for cfg in "2 0" "3 5000"; do
  unset MAPRAST_TAILS; export MAPRAST_F0=0.92; run $cfg
  for t in "0.04:16,0:8" "0.05:16,0.02:8,0:4" "0.03:24,0.03:16,0:8" "0.04:16,0.02:8,0:2"; do export MAPRAST_TAILS=$t; run $cfg; done
  export MAPRAST_F0=0.94; for t in "0.03:16,0:8" "0.03:16,0.02:8,0:4"; do export MAPRAST_TAILS=$t; run $cfg; done
  export MAPRAST_F0=0.90; for t in "0.05:24,0.03:16,0:8"; do export MAPRAST_TAILS=$t; run $cfg; done
done
