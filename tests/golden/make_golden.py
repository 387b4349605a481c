"""Regenerate tests/golden/runtests_vectors.json from the reference's own test file.

Run in the build container only (/root/reference is not on the GPU box):
    python tests/golden/make_golden.py
Parses /root/reference/test/runtests.jl:6-20 (the only known-answer vectors the
reference holds for the categorise path).  The typo `max=0a.3` at :14 is read as 0.3
(SURVEY.md Appendix B).  tol = 1.0 is the tolerance under which the reference
arithmetic (src/categories.jl:76-133) reproduces the expected labels.
"""
import json
import os
import re

SRC = "/root/reference/test/runtests.jl"
text = open(SRC).read().replace("0a.3", "0.3")
chan = re.compile(r"(r|g|b)\s*=\s*\(mean=([\d.]+),\s*min=([\d.]+),\s*max=([\d.]+),\s*sd=([\d.]+)\)")
found = chan.findall(text)
assert len(found) == 6, found
cats = []
for k in range(0, 6, 3):
    cats.append({c: {"mean": float(m), "min": float(a), "max": float(b), "sd": float(s)}
                 for c, m, a, b, s in found[k:k + 3]})
vec = re.compile(r"_categorisecolor\(RGB\(([\d.]+),\s*([\d.]+),\s*([\d.]+)\),\s*cats\)\s*==\s*(\d+)")
vectors = [{"rgb": [float(r), float(g), float(b)], "expected": int(e)} for r, g, b, e in vec.findall(text)]
assert len(vectors) == 4
out = {"source": "test/runtests.jl:6-20", "tol": 1.0, "categories": cats, "vectors": vectors}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "runtests_vectors.json")
json.dump(out, open(dst, "w"), indent=1)
print(open(dst).read())
