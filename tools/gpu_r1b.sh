#!/bin/bash
# VIMNMX3 micro-benchmark + tile-height variants of k_chain (device-resident bench per build)
mkdir -p gpurun_out
timeout 120 tools/ubench_minmax3 | tee gpurun_out/ubench_minmax3.txt
tools/gpu_variants.sh "$@"
