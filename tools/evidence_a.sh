#!/bin/bash
# evidence pass (run on the GPU box under gpurun, one part per call -- gpurun_out/ carries at most 64 MiB back):
#   part 1: ncu --set full of the C2 frame kernels (strip + tile) and the launch list of bench.py
#   part 2: ncu --set full of the C1 / C3 / C4 frames
#   part 3: compute-sanitizer logs, strip-kernel experiment table
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
O=gpurun_out
case "$1" in
1)
timeout 300 ncu --set full --import-source on --clock-control none -s 3 -c 1 -o $O/r02_strip_c2 -f python tools/prof_c2.py --config 2 --frames 5 > $O/ncu_strip.log 2>&1; tail -1 $O/ncu_strip.log
timeout 300 ncu --set full --clock-control none -s 3 -c 1 -o $O/r02_tile_c2 -f python tools/prof_c2.py --config 2 --frames 5 --tile > $O/ncu_tile.log 2>&1; tail -1 $O/ncu_tile.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_c2.csv python bench.py --steps 20 --warmup 3 > $O/bench_under_ncu.log 2>&1; tail -c 200 $O/bench_under_ncu.log; echo
;;
2)
timeout 300 ncu --set full --clock-control none -s 2 -c 1 -o $O/r02_c1 -f python tools/prof_c2.py --config 1 --frames 4 > $O/ncu_c1.log 2>&1; tail -1 $O/ncu_c1.log
timeout 300 ncu --set full --clock-control none -s 4 -c 4 -o $O/r02_c3 -f python tools/prof_c2.py --config 3 --frames 3 > $O/ncu_c3.log 2>&1; tail -1 $O/ncu_c3.log
timeout 300 ncu --set full --clock-control none -s 3 -c 3 -o $O/r02_c4 -f python tools/prof_c2.py --config 4 --frames 3 > $O/ncu_c4.log 2>&1; tail -1 $O/ncu_c4.log
;;
3)
T="tests/test_gpu_parity.py"
timeout 600 compute-sanitizer --tool memcheck --log-file $O/r02_sanitizer_memcheck.txt python -m pytest $T -q -x -k "strip_kernel_equals or two_instances_on_two_streams or span_kernel or resize_four or test_1x" 2>&1 | tail -2
timeout 600 compute-sanitizer --tool synccheck --log-file $O/r02_sanitizer_synccheck.txt python -m pytest $T -q -x -k "strip_kernel_equals or two_instances_on_two_streams or span_kernel" 2>&1 | tail -2
MPCVR_DEBUG_STRIP_KERNEL=0 timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --log-file $O/r02_sanitizer_racecheck_tile.txt python -m pytest $T -q -x -k "two_instances_on_two_streams or span_kernel" 2>&1 | tail -2
for f in memcheck synccheck racecheck_tile; do echo "== $f: $(tail -1 $O/r02_sanitizer_$f.txt)"; done
timeout 400 python tools/strip_check.py --cfgs 0,1,2,3,4,5 --waits 1,3,5 --masks 1,2,4,8,d,b,7,e,5,f 2>&1 | grep "TIME\|correctness" > $O/r02_strip_experiments.txt; cat $O/r02_strip_experiments.txt
;;
esac
