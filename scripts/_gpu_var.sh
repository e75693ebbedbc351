mkdir -p gpurun_out/r2
M="sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gpu__time_duration.sum,smsp__inst_executed.sum"
for v in $VARS; do
export FS2_LIB=$PWD/fast_slam_b200/variants/libfs2_$v.so
timeout -k 10 100 python scripts/bench_update.py --steps 12 --tag "$v" > gpurun_out/r2/var_$v.log 2>gpurun_out/r2/var_$v.err; echo "rc=$?"; cut -c1-330 gpurun_out/r2/var_$v.log
timeout -k 10 200 ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:fs2_update_ws_kernelILb0 -s 5 -c 1 --csv python scripts/bench_update.py --steps 4 2>/dev/null | grep -E "^\"[0-9]" | cut -d, -f13- | tr -d '"'
done
