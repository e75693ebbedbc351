set -x
mkdir -p gpurun_out/r2
for L in 256 258 272 288; do
timeout -k 10 100 python scripts/bench_update.py --landmarks $L --steps 5 --warmup 2 --tag L$L > gpurun_out/r2/tail_L$L.log 2> gpurun_out/r2/tail_L$L.err; cat gpurun_out/r2/tail_L$L.log | cut -c1-1500
done
timeout -k 10 100 python scripts/bench_update.py --landmarks 256 --steps 24 --warmup 2 --tag growth > gpurun_out/r2/tail_growth.log 2> gpurun_out/r2/tail_growth.err; cat gpurun_out/r2/tail_growth.log | cut -c1-2500
