#!/bin/bash
# Per-kernel time, DRAM bytes and tensor-pipe activity of one configs[4] log-linear step (ncu; serialised, cold caches)
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct
timeout 900 ncu --metrics $M --clock-control none --csv -s ${1:-40} -c ${2:-24} --log-file $O/ll5_metrics_r2b.csv \
  python tools/loglinear_bench.py 500000 200000 300 1024 3 1 > $O/ll5_metrics.log 2>&1
echo "ncu rc=$?"; tail -n 2 $O/ll5_metrics.log; wc -l $O/ll5_metrics_r2b.csv
