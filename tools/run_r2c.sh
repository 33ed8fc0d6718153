set -x
cd $GRAFT_REPO_ROOT
python tools/ncu_sampler.py > gpurun_out/r2c_sampler.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k4g_|k4p_|k4_sample' -c 12 -f -o gpurun_out/r2c_sampler python tools/ncu_sampler.py > gpurun_out/r2c_sampler_ncu.log 2>&1
ncu -i gpurun_out/r2c_sampler.ncu-rep --page raw --csv > gpurun_out/r2c_sampler.raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_sampler.ncu-rep --page source --csv -k regex:k4g_order > gpurun_out/r2c_order_source.csv 2>/dev/null
ls -la gpurun_out | tail
