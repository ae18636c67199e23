set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r10_c1_launches.csv python tools/config1_launches.py > gpurun_out/r10_c1.log 2>&1; echo c1=$?
tail -2 gpurun_out/r10_c1.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest_gpu.log 2>&1; echo pytest=$?
tail -5 gpurun_out/r10_pytest_gpu.log
