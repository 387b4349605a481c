export MAPRAST_LIB=maprasterization.jl_b200/variants/libmaprast_sched.so
for cfg in "2 0" "3 5000"; do set -- $cfg
for ft in "0.92 2" "0.92 1" "0.88 1" "0.84 1" "0.78 1" "0.88 2" "0.96 1"; do set -- $1 $2 $ft
  MAPRAST_F0=$3 MAPRAST_TAIL=$4 python scripts/kbench.py --cfg $1 --lines $2 --steps 12 | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg $1 lines $2 f0 $3 tail $4', round(d['ms_min'],4), round(d['ms_med'],4), d['digest'])"
done; done
