set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_imma tools/ubench_imma.cu && timeout 120 /tmp/ubench_imma > gpurun_out/r31_ubench_imma.log 2>&1; echo ub=$?
cat gpurun_out/r31_ubench_imma.log
