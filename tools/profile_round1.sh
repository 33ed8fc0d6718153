# ncu evidence for profiles/ (run on the GPU box through gpurun; each ncu pass only after the
# same command exited 0 without ncu)
set -x
cd $GRAFT_REPO_ROOT
C1="python bench.py --steps 3 --warmup 3 --no-e2e --no-extras --cpu-sample 2"
C2="python bench.py --mode ls_gram --dtype f64 --steps 2 --warmup 3 --no-e2e --no-extras --cpu-sample 1"
K='regex:k[0-4][a-z]*_'
$C1 > gpurun_out/p_c1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r1f_launches.csv $C1 > gpurun_out/p_c1_ncu.log 2>&1
$C1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k3t_lookup_kernel|k0_mask_build_smem_kernel|k0_transpose_kernel" --launch-skip 30 -c 3 -o gpurun_out/r1f_copy -f $C1 > gpurun_out/p_copy_ncu.log 2>&1
$C2 > gpurun_out/p_c2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k3t_lookup_kernel -c 1 -o gpurun_out/r1f_lsgram -f $C2 > gpurun_out/p_c2_ncu.log 2>&1
for f in r1f_copy r1f_lsgram; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -12
