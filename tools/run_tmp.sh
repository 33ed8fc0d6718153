cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_k0b.py tests/test_gpu_fullsize.py tests/test_gpu_mma.py -x -q 2>&1 | tail -2
for o in 0 1; do
  echo "== overlap=$o cfg3"; SPAI_K0B_OVERLAP=$o timeout 300 python tools/ab_k0.py cfg3 1024 bucket 2>&1 | grep -E "int64\+len|int32" | cut -c1-200
done
for g in 64 128 256; do echo "== overlap group=$g cfg3"; SPAI_K0B_GROUP=$g timeout 300 python tools/ab_k0.py cfg3 1024 bucket 2>&1 | grep -E "int64\+len" | cut -c1-200; done
for o in 0 1; do
  echo "== overlap=$o cfg4"; SPAI_K0B_OVERLAP=$o timeout 300 python tools/ab_k0.py cfg4 1024 bucket 2>&1 | grep -E "int64\+len" | cut -c1-200
done
