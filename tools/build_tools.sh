#!/bin/bash
# builds the stand-alone sm_100a microbenchmarks next to their sources (they travel to the GPU box with the snapshot; git ignores them):
#   ubench (instruction throughputs), latency (dependent-issue latencies), issue_probe (issue-slot model: what co-issues with FFMA2),
#   ffma2_forms (FFMA2 by operand form), role_bench (role A of the strip kernel in isolation), tma_probe, f32x2_test
set -e
cd "$(dirname "$0")/.."
N="/usr/local/cuda/bin/nvcc -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
for t in ubench latency issue_probe ffma2_forms; do $N -o tools/$t tools/$t.cu; done
$N -fmad=false -Iinclude -Ivideorenderer_b200/csrc -o tools/role_bench tools/role_bench.cu
$N -o tools/f32x2_test tools/f32x2_test.cu 2>/dev/null || true
$N -lcuda -o tools/tma_probe tools/tma_probe.cu 2>/dev/null || true
ls -la tools | grep -E "ubench|latency|issue_probe|ffma2_forms|role_bench" | grep -v "\.cu"
