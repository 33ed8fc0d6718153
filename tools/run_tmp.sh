cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_reward.py tests/test_gpu_lut.py tests/test_gpu_fuzz.py tests/test_gpu_mma.py tests/test_gpu_configs.py -x -q 2>&1 | tail -2
for t in 32 64; do echo "== transpose $t"; SPAI_K0_TRANSPOSE=$t timeout 600 python bench.py --steps 10 --no-e2e --no-extras --no-configs --no-sampler --cpu-sample 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], {k.split('(')[0]:round(v['ms'],4) for k,v in d['kernels'].items()})"; done
