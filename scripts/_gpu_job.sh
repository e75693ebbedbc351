set -x
mkdir -p gpurun_out/r2
T=${TAG:-z}
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py tests/test_gpu_api.py -m gpu -q -x --timeout 200 > gpurun_out/r2/pytest_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_$T.log
tail -3 gpurun_out/r2/pytest_$T.log
timeout -k 10 120 python scripts/bench_update.py --steps 24 --tag "$T" > gpurun_out/r2/var_$T.log 2>gpurun_out/r2/var_$T.err; cut -c1-330 gpurun_out/r2/var_$T.log
FS2_BENCH_VERBOSE=1 timeout -k 10 500 python bench.py --no-cpu-baseline --no-frontend --no-known > gpurun_out/r2/bench_$T.json 2> gpurun_out/r2/bench_$T.err; tail -3 gpurun_out/r2/bench_$T.err | cut -c1-400; cut -c1-300 gpurun_out/r2/bench_$T.json
