set -x
cd $GRAFT_REPO_ROOT
export SPAI_K3M_NTM=4
ncu --set full --clock-control none --import-source on -k regex:'k3m_kernel' -c 1 -f -o gpurun_out/r2g_k3m python tools/ncu_k3m.py cfg3 1024 > gpurun_out/r2g_ncu.log 2>&1
ncu -i gpurun_out/r2g_k3m.ncu-rep --page raw --csv > gpurun_out/r2g_k3m.raw.csv 2>/dev/null
ncu -i gpurun_out/r2g_k3m.ncu-rep --page source --csv > gpurun_out/r2g_k3m_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
