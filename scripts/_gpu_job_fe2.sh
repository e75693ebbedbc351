set -x
mkdir -p gpurun_out/r2
timeout -k 10 150 ncu --set full --import-source on --clock-control none -k regex:"fe_intersect_cluster|fe_vote_peaks" --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2/fe_full python scripts/fe_time.py 1024 quick > gpurun_out/r2/fe_ncu_full.log 2>&1; tail -3 gpurun_out/r2/fe_ncu_full.log
timeout -k 10 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/fe_launches_b1.csv python scripts/fe_time.py 1 quick > gpurun_out/r2/fe_ncu_b1.log 2>&1; tail -2 gpurun_out/r2/fe_ncu_b1.log
