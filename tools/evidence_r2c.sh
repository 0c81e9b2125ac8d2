#!/bin/bash
# Profile evidence of the second half of round 2 on one B200 (run under gpurun; outputs land in gpurun_out/).
O=gpurun_out
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
# launch lists (cold-cache, serialised: compare shares)
timeout 300 $L -s 60 -c 48 --log-file $O/launches_train_r2c.csv python tools/vs_step_bench.py 40 1 > $O/ev_train.log 2>&1
timeout 300 $L -s 14 -c 10 --log-file $O/launches_score3_r2c.csv python tools/score_bench.py 10000 50000 128 100 3 > $O/ev_s3.log 2>&1
timeout 300 $L -s 14 -c 10 --log-file $O/launches_score4_r2c.csv python tools/score_bench.py 10000 1000000 256 100 3 > $O/ev_s4.log 2>&1
# every kernel of one configs[4] log-linear step: time, DRAM bytes, tensor-pipe activity, L2 hit rate
bash tools/ll_profile.sh 40 24
# full captures: the projection GEMM (store epilogue, pair operands, clusters) and the N-major gWd GEMM of that step
F="ncu --set full --clock-control none --import-source on"
timeout 600 $F -k regex:gemm_tc -s 5 -c 3 -o $O/gemm_tc_ll5_r2c python tools/loglinear_bench.py 500000 200000 300 1024 2 1 > $O/ev_g5.log 2>&1
timeout 300 $F -k regex:dense_update -s 20 -c 1 -o $O/dense_update_r2c python tools/vs_step_bench.py 20 1 > $O/ev_du.log 2>&1
timeout 300 $F -k regex:vs_tile -s 10 -c 1 -o $O/vs_tile_r2c python tools/vs_step_bench.py 20 1 > $O/ev_vt.log 2>&1
tail -n 2 $O/ev_train.log $O/ev_s3.log $O/ev_s4.log $O/ev_g5.log
ls -la $O/*_r2c.ncu-rep
