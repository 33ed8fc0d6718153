set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_mma.py -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -30 gpurun_out/r2f_pytest.log
timeout 300 python tools/ab_k3m.py cfg3 1024 > gpurun_out/r2f_ab_cfg3.log 2>&1; tail -8 gpurun_out/r2f_ab_cfg3.log
timeout 300 python tools/ab_k3m.py cfg4 1024 > gpurun_out/r2f_ab_cfg4.log 2>&1; tail -8 gpurun_out/r2f_ab_cfg4.log
timeout 300 python tools/dbg_k3m.py cfg3 1024 3 2>&1 | tail -16
