set -x
mkdir -p gpurun_out/r2
export FS2_DIST_PROFILE=
for mode in placed p2p; do
FS2_DIST=$mode timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/sharded_check.py > gpurun_out/r2/sharded2_$mode.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sharded2_$mode.log
tail -4 gpurun_out/r2/sharded2_$mode.log | cut -c1-600
done
for mode in placed p2p; do
FS2_DIST=$mode timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n2_$mode.json 2> gpurun_out/r2/bench_n2_$mode.err; tail -3 gpurun_out/r2/bench_n2_$mode.err | cut -c1-400; cut -c1-1200 gpurun_out/r2/bench_n2_$mode.json
done
