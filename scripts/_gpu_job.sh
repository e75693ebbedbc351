set -x
mkdir -p gpurun_out/r2
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/r2/pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_g.log
tail -6 gpurun_out/r2/pytest_g.log
rm -f gpurun_out/r2/var_g.log
for v in "" notail; do
  if [ -z "$v" ]; then lib=fast_slam_b200/libfs2.so; else lib=fast_slam_b200/variants/libfs2_$v.so; fi
  FS2_LIB=$PWD/$lib timeout -k 10 120 python scripts/bench_update.py --steps 12 --tag "${v:-new}" >> gpurun_out/r2/var_g.log 2>gpurun_out/r2/var_g_${v:-new}.err
done
cut -c1-330 gpurun_out/r2/var_g.log
timeout -k 10 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/bench_g.json 2> gpurun_out/r2/bench_g.err; tail -3 gpurun_out/r2/bench_g.err; cut -c1-1500 gpurun_out/r2/bench_g.json
