cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_mma.py -x -q 2>&1 | tail -2
SPAI_K3M_VERBOSE=1 timeout 300 python tools/ab_k3m.py cfg3 1024 2>&1 | grep -E "k3m\]|kernel" | uniq | cut -c1-230
SPAI_K3M_VERBOSE=1 timeout 300 python tools/ab_k3m.py cfg4 1024 2>&1 | grep -E "k3m\]|kernel" | uniq | cut -c1-230
