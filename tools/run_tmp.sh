cd $GRAFT_REPO_ROOT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --no-e2e --no-extras --cpu-sample 4 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
for n,c in d['configs'].items():
    print(n, {k:(round(v,3) if isinstance(v,float) else v) for k,v in c.get('sample_to_reward(K4g taken-bitmask, no id lists)',{}).items()})
    for vn,v in c['variants'].items(): print('   ',vn, round(v['ms_per_step'],2), round(v['patterns_per_s']))
PY
