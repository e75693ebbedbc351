set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
FS2_DIST=placed timeout -k 10 300 $TR --nproc-per-node 8 --master-port 29517 tests/sharded_check.py > gpurun_out/r2/sharded8_placed.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sharded8_placed.log
tail -3 gpurun_out/r2/sharded8_placed.log | cut -c1-600
FS2_DIST_PROFILE=1 timeout -k 10 400 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n8_placed.json 2> gpurun_out/r2/bench_n8_placed.err; tail -2 gpurun_out/r2/bench_n8_placed.err | cut -c1-300; grep '^{' gpurun_out/r2/bench_n8_placed.json | cut -c1-900
timeout -k 10 400 $TR --nproc-per-node 8 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2/bench_n8_placed_b.json 2> gpurun_out/r2/bench_n8_placed_b.err; grep '^{' gpurun_out/r2/bench_n8_placed_b.json | cut -c1-700
timeout -k 10 600 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 12 --warmup 3 --workload cfg4 --no-cpu-baseline --no-parity > gpurun_out/r2/bench_n8_cfg4.json 2> gpurun_out/r2/bench_n8_cfg4.err; tail -2 gpurun_out/r2/bench_n8_cfg4.err | cut -c1-300; grep '^{' gpurun_out/r2/bench_n8_cfg4.json | cut -c1-900
timeout -k 10 400 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2/bench_n4_placed.json 2> gpurun_out/r2/bench_n4_placed.err; grep '^{' gpurun_out/r2/bench_n4_placed.json | cut -c1-700
