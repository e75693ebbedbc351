mkdir -p gpurun_out/r2
for v in a136; do
FS2_LIB=$PWD/fast_slam_b200/variants/libfs2_$v.so timeout -k 10 100 python scripts/bench_update.py --steps 12 --tag "$v" > gpurun_out/r2/var_$v.log 2>gpurun_out/r2/var_$v.err; echo "rc=$?"; cut -c1-330 gpurun_out/r2/var_$v.log
FS2_LIB=$PWD/fast_slam_b200/variants/libfs2_$v.so FS2_BENCH_VERBOSE=1 timeout -k 10 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-frontend --no-known > gpurun_out/r2/bench_$v.json 2> gpurun_out/r2/bench_$v.err; tail -3 gpurun_out/r2/bench_$v.err | head -2; cut -c1-330 gpurun_out/r2/bench_$v.json
done
