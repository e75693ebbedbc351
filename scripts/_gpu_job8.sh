set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -k 10 240 $TR --nproc-per-node 8 --master-port 29517 tests/sharded_check.py > gpurun_out/r2/sharded8_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2/sharded8_final.log
tail -3 gpurun_out/r2/sharded8_final.log | cut -c1-300
timeout -k 10 300 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n8_final.json 2> gpurun_out/r2/bench_n8_final.err; tail -2 gpurun_out/r2/bench_n8_final.err | cut -c1-300; grep '^{' gpurun_out/r2/bench_n8_final.json | cut -c1-900
timeout -k 10 300 $TR --nproc-per-node 4 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2/bench_n4_final.json 2> gpurun_out/r2/bench_n4_final.err; grep '^{' gpurun_out/r2/bench_n4_final.json | cut -c1-500
timeout -k 10 400 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --steps 12 --warmup 3 --workload cfg4 --no-cpu-baseline --no-parity > gpurun_out/r2/bench_n8_cfg4_final.json 2> gpurun_out/r2/bench_n8_cfg4_final.err; tail -2 gpurun_out/r2/bench_n8_cfg4_final.err | cut -c1-300; grep '^{' gpurun_out/r2/bench_n8_cfg4_final.json | cut -c1-700
