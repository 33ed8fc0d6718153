#!/bin/bash
# scratch runner for gpurun: K0c tests + A/B
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_k0c.py -x -q > gpurun_out/k0c_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/k0c_pytest.log
timeout -s KILL 600 python tools/ab_k0.py cfg3 1024 bucket,cluster,cluster:4:2,cluster:4:4,cluster:8:4 > gpurun_out/ab_k0c_cfg3.log 2>&1; echo "ab rc=$?"
grep -v "^$" gpurun_out/ab_k0c_cfg3.log | cut -c1-400 | tail -30
