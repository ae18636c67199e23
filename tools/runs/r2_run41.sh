set -x
mkdir -p gpurun_out
OUT=gpurun_out/r41_ab_seg_target.txt bash tools/ab_bench.sh "seg_target=8192" "seg_target=10240" "seg_target=6144" "seg_target=12288" "seg_target=8192"
