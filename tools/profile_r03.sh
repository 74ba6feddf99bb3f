#!/bin/bash
# round 2, second half: ncu evidence for the c2 path after the recurrence-compare / carried-state changes (B200_PROFILING.md recipe), one GPU.
# c3's and c4's kernels did not change: profiles/r02_fast_frame_c3.txt, r02_main_c4.txt, r02_launches_c3.csv / _c4.csv stand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu-baseline --no-full-trips"
# (1) launch list of the same command (cold-cache, serialised: the SHARES matter, not the absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_c2.csv $B --workload c2 --steps 2 --warmup 3 > /dev/null 2>&1
# (2) one strand, per-kernel table of a c2 frame (frame 4: 13 launches per frame)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum
CHAOS_STRANDS=1 timeout 600 ncu --metrics $M --clock-control none --launch-skip 39 --launch-count 13 --csv --log-file gpurun_out/r03_passes_c2.csv $B --workload c2 --steps 1 --warmup 3 > /dev/null 2>&1
# (3) full capture with source of the dominant kernels: c2's long + probe (pass A of frame 4)
CHAOS_STRANDS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"chaosLongDouble|chaosProbeDouble" --launch-skip 12 --launch-count 2 -o gpurun_out/r03_long_probe_c2 -f $B --workload c2 --steps 1 --warmup 3 > /dev/null 2>&1
echo done
