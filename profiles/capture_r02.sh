#!/bin/bash
# round-2 evidence, run on the GPU box from the repo root:  bash profiles/capture_r02.sh [TAG]
#  1. ncu --set full of the two hot kernels at cfg3 and at cfg5 (one capture each; never a bench value)
#  2. launch list of two bench steps (kernel shares)
#  3. summaries -> gpurun_out/ (copied into profiles/ by hand afterwards)
T=${1:-r02}
O=gpurun_out
mkdir -p $O
for WL in cfg3 cfg5; do
  B=""; N=""; [ $WL = cfg5 ] && B="--batch 16" && N=16
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'setup_kernel|raster_kernel|backward_tile' -s 9 -c 3 \
      -o $O/prof_${T}_$WL -f python bench.py --workload $WL $B --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $O/prof_${T}_$WL.log 2>&1
  ncu -i $O/prof_${T}_$WL.ncu-rep --page raw --csv > $O/prof_${T}_${WL}_raw.csv 2>/dev/null
  ncu -i $O/prof_${T}_$WL.ncu-rep --page source --csv --print-source cuda,sass -k regex:raster > $O/src_raster_${T}_$WL.csv 2>/dev/null
  ncu -i $O/prof_${T}_$WL.ncu-rep --page source --csv --print-source cuda,sass -k regex:backward > $O/src_bwd_${T}_$WL.csv 2>/dev/null
  python profiles/ncu_summary.py $O/prof_${T}_${WL}_raw.csv > $O/${T}_ncu_summary_$WL.txt
  python profiles/ncu_lines.py $O/src_bwd_${T}_$WL.csv > $O/${T}_backward_lines_$WL.txt 2>/dev/null
  python profiles/ncu_lines.py $O/src_raster_${T}_$WL.csv > $O/${T}_raster_lines_$WL.txt 2>/dev/null
  python profiles/ncu_opcodes.py $O/src_bwd_${T}_$WL.csv > $O/${T}_backward_opcodes_$WL.txt 2>/dev/null
  python profiles/ncu_opcodes.py $O/src_raster_${T}_$WL.csv > $O/${T}_raster_opcodes_$WL.txt 2>/dev/null
  python profiles/make_traffic.py $O/prof_${T}_${WL}_raw.csv $WL profiles/${T}_ncu_summary_$WL.txt $N > /dev/null
  rm -f $O/prof_${T}_$WL.ncu-rep $O/src_*_${T}_$WL.csv
done
cp profiles/traffic.json $O/traffic.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 16 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > /dev/null 2>&1
