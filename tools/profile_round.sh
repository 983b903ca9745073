#!/bin/bash
# Round profiling recipe (run under gpurun, one GPU): the bench command first WITHOUT ncu, then (a) the launch list of the
# same command (device time of every launch, cold-cache and serialised: compare shares), (b) one `--set full` capture of
# every kernel of one step.  Outputs land in gpurun_out/; the summaries judged are copied into profiles/ from here.
set -u
CMD="python bench.py --steps 5 --warmup 3 --preroll 20 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.err; exit 1; }
# preroll 20 + warmup 3 steps precede the timed region; ~22 launches per step
ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -s 520 -c 24 -o gpurun_out/prof_step $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
