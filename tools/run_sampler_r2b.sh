set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_race.py tests/test_gpu_sampler.py -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --no-configs --no-e2e --steps 5 --cpu-sample 4 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(json.dumps(d['sampler'], indent=1))"
