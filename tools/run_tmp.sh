#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_k0b.py tests/test_gpu_fullsize.py -x -q > gpurun_out/k0b6_pytest.log 2>&1; echo "pytest rc=$?"
for l in 1 2 4; do SPAI_K0B_LANES=$l timeout -s KILL 900 python -m pytest tests/test_gpu_k0b.py -x -q > gpurun_out/k0b6_pytest_l$l.log 2>&1; echo "pytest lanes $l rc=$?"; done
tail -n 1 gpurun_out/k0b6_pytest*.log
for cfg in cfg3 cfg4 cfg5; do
  B=1024; [ $cfg != cfg3 ] && B=512
  for l in 0 1 2 4; do
    [ $cfg = cfg3 ] && [ $l = 1 ] && continue
    SPAI_K0B_LANES=$l SPAI_K0B_TIMING=1 timeout -s KILL 600 python tools/ab_k0.py $cfg $B bucket > gpurun_out/ab_k0b6_${cfg}_l$l.log 2>&1; echo "ab $cfg lanes $l rc=$?"
    grep "k0b\]\|\"input\"" gpurun_out/ab_k0b6_${cfg}_l$l.log | sed 's/"GBps_on_read.*"step_ms"/ step_ms/' | cut -c1-140 | sed -n '4,5p'
  done
done
