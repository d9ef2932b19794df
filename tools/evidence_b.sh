#!/bin/bash
# evidence pass of the second round-2 session (run on the GPU box under gpurun; gpurun_out/ carries at most 64 MiB back):
#   ncu --set full of the C2 frame kernel as the product dispatches it, the launch list of bench.py, compute-sanitizer
#   memcheck / synccheck on the strip-kernel tests, racecheck on the tile kernel and the span kernels
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -s 3 -c 1 -o $O/r02b_strip_c2 -f python tools/prof_c2.py --config 2 --frames 5 > $O/ncu_strip.log 2>&1; tail -1 $O/ncu_strip.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_bench_c2.csv python bench.py --steps 20 --warmup 3 > $O/bench_under_ncu.log 2>&1; tail -c 200 $O/bench_under_ncu.log; echo
T="tests/test_gpu_parity.py"
timeout 600 compute-sanitizer --tool memcheck --log-file $O/r02b_sanitizer_memcheck.txt python -m pytest $T -q -x -k "strip_kernel_equals or two_instances_on_two_streams or span_kernel or test_1x" 2>&1 | tail -2
timeout 600 compute-sanitizer --tool synccheck --log-file $O/r02b_sanitizer_synccheck.txt python -m pytest $T -q -x -k "strip_kernel_equals or two_instances_on_two_streams" 2>&1 | tail -2
for f in memcheck synccheck; do echo "== $f: $(tail -1 $O/r02b_sanitizer_$f.txt)"; done
