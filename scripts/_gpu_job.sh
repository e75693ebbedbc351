set -x
mkdir -p gpurun_out/r2
T=${TAG:-j}
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -m gpu -q -x --timeout 120 -k "deferred or skewed or multi_step or motion_update or stress or golden" > gpurun_out/r2/pytest_${T}_sub.log 2>&1; echo "rc=$?" >> gpurun_out/r2/pytest_${T}_sub.log
tail -4 gpurun_out/r2/pytest_${T}_sub.log
timeout -k 10 120 python scripts/bench_update.py --steps 12 --tag "$T" > gpurun_out/r2/var_$T.log 2>gpurun_out/r2/var_$T.err; cut -c1-330 gpurun_out/r2/var_$T.log
FS2_BENCH_VERBOSE=1 timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-frontend --no-known > gpurun_out/r2/bench_$T.json 2> gpurun_out/r2/bench_$T.err; tail -3 gpurun_out/r2/bench_$T.err; cut -c1-700 gpurun_out/r2/bench_$T.json
timeout -k 10 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:fs2_update_ws_kernelILb1 -s 1 -c 1 -o gpurun_out/r2/upd_defer_$T -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-frontend --no-known > gpurun_out/r2/ncu_$T.log 2>&1; tail -3 gpurun_out/r2/ncu_$T.log | cut -c1-300
timeout -k 10 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:fs2_update_ws_kernelILb0 -s 5 -c 1 -o gpurun_out/r2/upd_plain_$T -f python scripts/bench_update.py --steps 4 > gpurun_out/r2/ncu_${T}_plain.log 2>&1; tail -3 gpurun_out/r2/ncu_${T}_plain.log | cut -c1-300
