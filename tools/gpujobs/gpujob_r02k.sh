set -x
cd tools/experiments && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o packed_probe packed_fp32_probe.cu && ./packed_probe
