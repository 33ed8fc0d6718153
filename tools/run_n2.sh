set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2k_multi_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_multi_pytest.log; tail -4 gpurun_out/r2k_multi_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2k_bench_n2.err
