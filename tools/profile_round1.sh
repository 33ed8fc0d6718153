set -x
cd $GRAFT_REPO_ROOT
C1="python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --cpu-sample 2"
C2="python bench.py --mode ls_gram --dtype f64 --steps 2 --warmup 3 --no-e2e --no-extras --cpu-sample 1"
C3="python bench.py --mode ls_gram --dtype f32 --steps 2 --warmup 3 --no-e2e --no-extras --cpu-sample 1"
K='regex:k[0-4][a-z]*_'
$C1 > gpurun_out/p_c1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r1e_launches.csv $C1 > gpurun_out/p_c1_ncu.log 2>&1
$C1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3_copy_kernel -c 1 -o gpurun_out/r1e_k3 -f $C1 > gpurun_out/p_k3_ncu.log 2>&1
$C2 > gpurun_out/p_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2g_solve -c 1 -o gpurun_out/r1e_k2g_f64 -f $C2 > gpurun_out/p_c2_ncu.log 2>&1
$C3 > gpurun_out/p_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2g_solve -c 1 -o gpurun_out/r1e_k2g_f32 -f $C3 > gpurun_out/p_c3_ncu.log 2>&1
for f in r1e_k3 r1e_k2g_f64 r1e_k2g_f32; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -20
