set -x
mkdir -p gpurun_out/r2
timeout -k 10 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_final.log
tail -6 gpurun_out/r2/pytest_final.log
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2/smoke_final.log; tail -2 gpurun_out/r2/smoke_final.log
timeout -k 10 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_driver_flags.json 2> gpurun_out/r2/bench_driver_flags.err; tail -2 gpurun_out/r2/bench_driver_flags.err | cut -c1-300; grep '^{' gpurun_out/r2/bench_driver_flags.json | cut -c1-600
