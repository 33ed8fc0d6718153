# 8-GPU evidence (gpurun --gpus 8): NCCL test + the bench line (headline, e2e, policy-gradient all-reduce; side configs skipped to
# keep the 8x-charged call short). The driver's own SCALE run is the record; this is the builder-side check.
set -x
cd $GRAFT_REPO_ROOT
N=${1:-8}
TAG=${2:-r2p}
nvidia-smi --query-gpu=index,name --format=csv | head -12
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${TAG}_multi_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_multi_pytest_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-configs --no-extras --no-sampler > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("dp_training"))
PY
