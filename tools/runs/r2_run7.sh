set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "streaming" > gpurun_out/r7_pytest_stream.log 2>&1; echo pytest=$?
tail -30 gpurun_out/r7_pytest_stream.log
timeout 600 python tools/stream_probe.py > gpurun_out/r7_stream_probe.json 2> gpurun_out/r7_stream_probe.err; echo probe=$?
cat gpurun_out/r7_stream_probe.json; tail -5 gpurun_out/r7_stream_probe.err
