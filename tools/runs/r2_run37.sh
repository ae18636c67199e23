set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_p1 tools/ubench_p1.cu && timeout 120 /tmp/ubench_p1 > gpurun_out/r37_ubench_p1.log 2>&1; echo ub=$?
cat gpurun_out/r37_ubench_p1.log
