#!/bin/bash
# Build the standalone dev harnesses into build/ (git-ignored; travels to the GPU box with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p build
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
$NVCC $FLAGS $EXTRA -o build/dev_chol_flow${SUFFIX} scripts/dev_chol_flow.cu
