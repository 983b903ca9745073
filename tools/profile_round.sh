#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): the bench command first WITHOUT ncu, then (a) the launch list of the
# same command (device time of every launch, cold-cache and serialised: compare shares), (b) one `--set full` capture of
# every kernel of one step.  Outputs land in gpurun_out/; the summaries judged are copied into profiles/ from here.
#   tools/profile_round.sh [worlds] [launches_per_step]
# 8192 worlds (lane tiles): 23 kernel launches per step (place sort bp_count bp_scan bp_emit narrowphase_group x4 slot_counts
# islands island_sort equations team<256> team<32> lanes<1> x3 lanes<2> x3 lanes<0> integrate).  1024 worlds (quad family):
# 20 (... narrowphase ... equations qteam x4 quads x5 integrate).  + 1 k_place at upload; 120 pre-roll + 3 warm-up steps, then the four k_out_*
# launches of the snapshot download, then the timed steps.
set -u
W=${1:-8192}
LPS=${2:-23}
CMD="python bench.py --worlds $W --steps 5 --warmup 3 --no-cpu-baseline --no-pile --no-cars --no-weak --e2e-steps 1"
SKIP=$((1 + 123 * LPS + 4))
$CMD > gpurun_out/prof_plain_$W.json 2> gpurun_out/prof_plain_$W.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_$W.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $((5 * LPS)) --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_launches_$W.log 2>&1
ncu --set full --clock-control none --import-source on -s $SKIP -c $LPS -f -o gpurun_out/prof_step_$W $CMD > gpurun_out/ncu_full_$W.log 2>&1
tail -2 gpurun_out/ncu_full_$W.log
