#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["platform_h2d_probe"])
for c,v in d["configs"].items(): print(c, {m:round(x["ms_per_step"],2) for m,x in v["variants"].items()})
PY
