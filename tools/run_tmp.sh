#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_k0b.py tests/test_gpu_k0c.py tests/test_gpu_fullsize.py -x -q > gpurun_out/k0b8_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 1 gpurun_out/k0b8_pytest.log
for cfg in cfg3 cfg5; do
  B=1024; [ $cfg != cfg3 ] && B=512
  SPAI_K0B_TIMING=1 timeout -s KILL 600 python tools/ab_k0.py $cfg $B bucket > gpurun_out/ab_k0b8_${cfg}.log 2>&1; echo "ab $cfg rc=$?"
  grep "k0b\]\|\"input\"" gpurun_out/ab_k0b8_${cfg}.log | sed 's/"GBps_on_read.*"step_ms"/ step_ms/' | cut -c1-140 | sed -n '4,5p;14,15p'
done
