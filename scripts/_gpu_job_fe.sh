set -x
mkdir -p gpurun_out/r2
timeout -k 10 240 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_stress.py -x -q -m gpu -k "frontend or hough or line_filter or landmark_utils" > gpurun_out/r2/pytest_fe6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_fe6.log
tail -15 gpurun_out/r2/pytest_fe6.log
timeout -k 10 120 python scripts/fe_time.py > gpurun_out/r2/fe_time6.json 2> gpurun_out/r2/fe_time6.err; cat gpurun_out/r2/fe_time6.json; tail -3 gpurun_out/r2/fe_time6.err
timeout -k 10 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/fe_launches6.csv python scripts/fe_time.py 1024 quick > gpurun_out/r2/fe_ncu6.log 2>&1; tail -2 gpurun_out/r2/fe_ncu6.log
