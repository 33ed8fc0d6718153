# Round-2 evidence for profiles/ (GPU box, through gpurun). Each ncu pass runs only after the same
# command exited 0 without ncu; numbers printed under ncu are never bench values.
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
C1="python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --no-configs --cpu-sample 2"
K='regex:k[0-5][a-z]*_'
$C1 > gpurun_out/${TAG}_c1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $C1 > gpurun_out/${TAG}_c1_ncu.log 2>&1
python tools/ncu_targets.py > gpurun_out/${TAG}_targets.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k0b_|k1_fill|k4_|k4g_|k4p_|k3s_sparse|k3t_lookup|k0_mask_build|k0_transpose|k3_copy|k3m_kernel' -c 60 -f -o gpurun_out/${TAG}_targets python tools/ncu_targets.py > gpurun_out/${TAG}_targets_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_targets.ncu-rep --page raw --csv > gpurun_out/${TAG}_targets.raw.csv 2>/dev/null
# gpurun_out/ travels back only below 64 MiB: keep the CSV pages, drop the report itself when it is large
if [ $(stat -c %s gpurun_out/${TAG}_targets.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/${TAG}_targets.ncu-rep; fi
ls -la gpurun_out | tail -20
