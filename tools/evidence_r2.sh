#!/bin/bash
# Round-2 profile evidence on one B200 (run under gpurun; outputs land in gpurun_out/).
set -x
O=gpurun_out
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
# launch lists (cold-cache, serialised: compare shares)
$L -s 60 -c 48 --log-file $O/launches_train_r2.csv python tools/vs_step_bench.py 40 1 > $O/ev_train.log 2>&1
$L -s 14 -c 10 --log-file $O/launches_score3_r2.csv python tools/score_bench.py 10000 50000 128 100 3 > $O/ev_s3.log 2>&1
$L -s 14 -c 10 --log-file $O/launches_score4_r2.csv python tools/score_bench.py 10000 1000000 256 100 3 > $O/ev_s4.log 2>&1
$L -s 44 -c 22 --log-file $O/launches_ll5_r2.csv python tools/loglinear_bench.py 500000 200000 300 1024 3 1 > $O/ev_ll5.log 2>&1
# full captures of the dominant kernels
F="ncu --set full --clock-control none --import-source on"
$F -k regex:dense_update -s 20 -c 1 -o $O/dense_update_r2 python tools/vs_step_bench.py 20 1 > $O/ev_du.log 2>&1
$F -k regex:vs_tile -s 10 -c 1 -o $O/vs_tile_r2 python tools/vs_step_bench.py 20 1 > $O/ev_vt.log 2>&1
$F -k regex:gemm_tc -s 5 -c 1 -o $O/gemm_tc_score4_r2 python tools/score_bench.py 10000 1000000 256 100 1 > $O/ev_g4.log 2>&1
$F -k regex:finalize_kernel -s 2 -c 1 -o $O/finalize_score4_r2 python tools/score_bench.py 10000 1000000 256 100 1 > $O/ev_f4.log 2>&1
$F -k regex:ll_dz_split -s 1 -c 1 -o $O/ll_dz_split_r2 python tools/loglinear_bench.py 500000 200000 300 1024 2 1 > $O/ev_dz.log 2>&1
tail -n 2 $O/ev_train.log $O/ev_s3.log $O/ev_s4.log $O/ev_ll5.log
ls -la $O/*_r2.ncu-rep
