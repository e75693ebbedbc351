set -x
mkdir -p gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for d in 1 0; do
FS2_DEFER=$d timeout -k 10 400 $TR --nproc-per-node 2 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_n2_defer$d.json 2> gpurun_out/r2/bench_n2_defer$d.err; tail -2 gpurun_out/r2/bench_n2_defer$d.err | cut -c1-300; grep '^{' gpurun_out/r2/bench_n2_defer$d.json | cut -c1-1000
done
