#!/bin/bash
# usage: tools/sweep_threads.sh 512 640 1024  -- rebuilds the library with kThreads=N and prints the C2 bench line (on the GPU box)
for T in "$@"; do
  sed -i "s/constexpr int kThreads = [0-9]*;/constexpr int kThreads = $T;/" videorenderer_b200/csrc/kernel_fused.cu
  python videorenderer_b200/build.py --force > /dev/null 2>&1
  python bench.py --steps 100 --warmup 10 --no-cpu 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('threads $T', d['config']['path'], round(d['value'],1), 'fps', round(d['ms_per_step']*1000,1), 'us')"
done
