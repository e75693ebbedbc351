mkdir -p gpurun_out/r2
M="sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__inst_executed.sum"
timeout -k 10 200 ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:fs2_update_ws_kernelILb1 -s 1 -c 2 --csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-frontend --no-known 2>/dev/null | grep -E "^\"[0-9]" | cut -d, -f5,13- | tr -d '"' > gpurun_out/r2/icc_${TAG}_defer.txt
timeout -k 10 200 ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:fs2_update_ws_kernelILb0 -s 5 -c 1 --csv python scripts/bench_update.py --steps 4 2>/dev/null | grep -E "^\"[0-9]" | cut -d, -f5,13- | tr -d '"' > gpurun_out/r2/icc_${TAG}_plain.txt
timeout -k 10 200 ncu --metrics $M --clock-control none --kernel-name-base mangled -k regex:fs2_update_ws_kernelILb0 -s 25 -c 1 --csv python scripts/bench_update.py --steps 24 2>/dev/null | grep -E "^\"[0-9]" | cut -d, -f5,13- | tr -d '"' > gpurun_out/r2/icc_${TAG}_late.txt
for f in defer plain late; do echo "== $f"; cat gpurun_out/r2/icc_${TAG}_$f.txt; done
