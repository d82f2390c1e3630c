#!/bin/bash
# round-1 closing evidence: ncu capture of the hot kernels, launch list, then the bench lines (outside any profiler)
cd $GRAFT_REPO_ROOT
T=r01c
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'raster_kernel|backward_tile' -s 6 -c 2 -o gpurun_out/prof_$T -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/prof_$T.log 2>&1
ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv --print-source cuda,sass -k regex:raster > gpurun_out/src_raster_$T.csv 2>/dev/null
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv --print-source cuda,sass -k regex:backward > gpurun_out/src_bwd_$T.csv 2>/dev/null
rm -f gpurun_out/prof_$T.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 24 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > /dev/null 2>&1
python profiles/make_traffic.py gpurun_out/prof_${T}_raw.csv cfg3 profiles/r01_final_ncu_summary.txt > /dev/null
cp profiles/traffic.json gpurun_out/traffic.json
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference > gpurun_out/bench_${T}_reference.json 2>> gpurun_out/bench_$T.err
python bench.py --background uniform --no-e2e --no-cpu-baseline > gpurun_out/bench_${T}_uniform_bg.json 2>> gpurun_out/bench_$T.err
for wl in cfg5 cfg4 cfg2; do python bench.py --workload $wl --no-e2e --no-cpu-baseline > gpurun_out/bench_${T}_$wl.json 2>> gpurun_out/bench_$T.err; done
python -c "
import json
for n in ['', '_reference', '_uniform_bg', '_cfg5', '_cfg4', '_cfg2']:
    d = json.load(open('gpurun_out/bench_r01c%s.json' % n))
    print(n or 'cfg3', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('transfers_only_ms'), (d.get('roofline') or {}).get('frac'), (d.get('cpu_baseline') or {}).get('value'))
"
