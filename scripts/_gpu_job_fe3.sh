set -x
mkdir -p gpurun_out/r2
timeout -k 10 150 ncu --set full --import-source on --clock-control none -k regex:"fe_intersect_cluster|fe_vote_peaks|fe_raster_list" --launch-skip 3 --launch-count 3 -f -o gpurun_out/r2/fe_full_final python scripts/fe_time.py 1024 quick > gpurun_out/r2/fe_ncu_full_final.log 2>&1; tail -3 gpurun_out/r2/fe_ncu_full_final.log
