import sys
sys.path.insert(0, '/root/repo')
import maprast_b200 as mr
from maprast_b200 import synth
ctx = mr.Context(0)
for cfg in (2, 4):
    seed, palrgb, st, spec = synth.config_workload(cfg)
    for _ in range(3):
        ctx.set_palette(st.mean, st.lo, st.hi, mr._lib.LAB if st.colourspace == "lab" else mr._lib.RGB)
    print(cfg, ctx.palette_build_ms())
