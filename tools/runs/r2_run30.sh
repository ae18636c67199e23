set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wide_formats.py -m gpu -q > gpurun_out/r30_pytest_wide.log 2>&1; echo wide=$?
tail -3 gpurun_out/r30_pytest_wide.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r30_pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/r30_pytest_gpu.log
