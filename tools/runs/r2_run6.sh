set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r6_pytest_gpu.log 2>&1; echo pytest=$?
tail -5 gpurun_out/r6_pytest_gpu.log
timeout 300 python tools/config3_phases.py > gpurun_out/r6_config3_phases.json 2> gpurun_out/r6_config3_phases.err; echo c3=$?
