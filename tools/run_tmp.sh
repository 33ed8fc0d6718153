cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_mma.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --no-e2e --no-extras --no-sampler --cpu-sample 2 --configs cfg3,cfg4,cfg5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
for n,c in d['configs'].items():
    for vn,v in c['variants'].items(): print(n, vn, round(v['ms_per_step'],2), round(v['patterns_per_s']), {a:round(b,3) for a,b in v['kernel_share'].items()}, c['parity_check'].get(vn))
PY
