set -x
mkdir -p gpurun_out/r2
T=${TAG:-k}
timeout -k 10 600 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/r2/pytest_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_$T.log
tail -4 gpurun_out/r2/pytest_$T.log
timeout -k 10 120 python scripts/bench_update.py --steps 12 --tag "$T" > gpurun_out/r2/var_$T.log 2>gpurun_out/r2/var_$T.err; cut -c1-330 gpurun_out/r2/var_$T.log
FS2_BENCH_VERBOSE=1 timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-frontend --no-known > gpurun_out/r2/bench_$T.json 2> gpurun_out/r2/bench_$T.err; tail -3 gpurun_out/r2/bench_$T.err; cut -c1-700 gpurun_out/r2/bench_$T.json
