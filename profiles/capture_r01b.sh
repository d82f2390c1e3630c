set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'raster_kernel|backward_tile' -s 6 -c 2 -o gpurun_out/prof_r01b -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/prof_r01b.log 2>&1
ncu -i gpurun_out/prof_r01b.ncu-rep --page raw --csv > gpurun_out/prof_r01b_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01b.ncu-rep --page source --csv --print-source cuda,sass -k regex:raster > gpurun_out/src_raster_r01b.csv 2>/dev/null
ncu -i gpurun_out/prof_r01b.ncu-rep --page source --csv --print-source cuda,sass -k regex:backward > gpurun_out/src_bwd_r01b.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 24 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > /dev/null 2>&1
rm -f gpurun_out/prof_r01b.ncu-rep
ls -la gpurun_out | tail -8
