#!/bin/bash
# final check of the committed state: full GPU suite, smoke, the default bench line
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 1 gpurun_out/r2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2r_smoke.log
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
for c,v in d["configs"].items(): print(c, {m:(round(x["ms_per_step"],2), round(x["k0_frac_of_hbm_peak_on_valid_bytes"],3)) for m,x in v["variants"].items()}, v["parity_check"])
PY
