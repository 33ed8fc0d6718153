# full default bench line on N GPUs (weak cfg2/cfg3/cfg4, strong cfg5): builder-side copy of what the driver's SCALE run records
set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}
TAG=${2:-r2q}
S=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N > gpurun_out/${TAG}_bench_full_n$N.json 2> gpurun_out/${TAG}_bench_full_n$N.err; echo "bench rc=$? wall=$(( $(date +%s) - S )) s"
tail -c 300 gpurun_out/${TAG}_bench_full_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_full_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_aggregate_gbps"], d["e2e"]["platform_h2d_probe"])
for c,v in d["configs"].items(): print(c, v["scaling"], v["global_batch"], {m:(round(x["ms_per_step"],2), round(x["patterns_per_s"])) for m,x in v["variants"].items()})
PY
