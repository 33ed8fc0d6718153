#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --no-configs --no-extras --no-sampler > gpurun_out/probe_n1.json 2> gpurun_out/probe_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --no-configs --no-extras --no-sampler > gpurun_out/probe_n2.json 2> gpurun_out/probe_n2.err; echo "n2 rc=$?"
python - <<PY
import json
for n in (1,2):
    d=json.loads(open(f"gpurun_out/probe_n{n}.json").read().strip().splitlines()[-1])
    e=d["e2e"]; print(n, d["value"], e["value"], e["h2d_gbps_per_gpu"], e["h2d_aggregate_gbps"], e["platform_h2d_probe"])
PY
tail -c 300 gpurun_out/probe_n2.err
