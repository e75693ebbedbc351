set -x
mkdir -p gpurun_out/r2
T=${TAG:-aa}
for rep in 1 2; do
timeout -k 10 600 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/r2/pytest_${T}_$rep.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_${T}_$rep.log
tail -3 gpurun_out/r2/pytest_${T}_$rep.log
done
