#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): the bench command first WITHOUT ncu, then (a) the launch list of the
# same command (device time of every launch, cold-cache and serialised: compare shares), (b) one `--set full` capture of
# every kernel of one step.  Outputs land in gpurun_out/; the summaries judged are copied into profiles/ from here.
# 21 kernel launches per step (place sort bp_count bp_scan bp_emit narrowphase slot_counts islands island_sort equations
# tile_build team<256> team<32> lanes<1> x3 lanes<2> x3 lanes<0> integrate) + 1 k_place at upload; 120 pre-roll + 3 warm-up
# steps, then the four k_out_* launches of the snapshot download, then the timed steps.
set -u
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pile --e2e-steps 1"
SKIP=$((1 + 123 * 21 + 4))
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 105 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -s $SKIP -c 21 -o gpurun_out/prof_step_final $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
