#!/bin/bash
# Round-2 profile captures on the GPU box.  Every .ncu-rep is exported to CSV and deleted on the box (gpurun copies back <= 64 MiB).
#   bash tools/capture_profiles.sh [launches|conv|hbm ...]
set -u
mkdir -p gpurun_out
what="${*:-launches conv hbm}"
for w in $what; do
  case $w in
    launches)
      timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_O_r02.csv python tools/ncu_step.py --batch 128 > gpurun_out/cap_ncuO.log 2>&1
      timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_V_r02.csv python tools/ncu_step.py --batch 32 --workload V > gpurun_out/cap_ncuV.log 2>&1
      ;;
    conv)
      for kind in fwd wgrad; do
        if [ $kind = fwd ]; then k=regex:conv_fwd_tc2; name=prof_conv_fwd_tc2_shapes_r02; else k=regex:conv_wgrad_tc; name=prof_conv_wgrad_tc_shapes_r02; fi
        timeout 300 ncu --set full --clock-control none --import-source on -k $k -f -o gpurun_out/$name python tools/conv_bench.py --only $kind --shapes 0,1,2 --no-check --reps 1 > gpurun_out/cap_$kind.log 2>&1
        ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
        ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.source.csv 2>/dev/null
        python tools/ncu_source_top.py gpurun_out/$name.source.csv 60 > gpurun_out/$name.source_top.txt 2>&1
        rm -f gpurun_out/$name.ncu-rep gpurun_out/$name.source.csv
      done
      ;;
    hbm)
      timeout 120 python tools/hbm_bench.py > gpurun_out/cap_hbm.log 2>&1
      timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -f -o gpurun_out/hbm_r02 python tools/hbm_bench.py --ncu > gpurun_out/cap_hbmncu.log 2>&1
      ncu -i gpurun_out/hbm_r02.ncu-rep --page raw --csv > gpurun_out/hbm_raw.csv 2>/dev/null
      rm -f gpurun_out/hbm_r02.ncu-rep
      ;;
  esac
done
du -sh gpurun_out
