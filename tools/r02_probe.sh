#!/bin/bash
# Round-2 first GPU call: toolchain probe, baseline timings, island class distribution, compute-sanitizer.
set -u
O=gpurun_out
{ echo "== which"; which node nodejs elm elm-test deno bun npm 2>&1; echo "rc=$?"; ls /usr/lib/node_modules /usr/local/lib/node_modules 2>&1 | head; nproc; nvidia-smi -L; } > $O/r02_toolchain_probe.txt 2>&1
python tools/time_scene.py mixed:1024 --steps 50 --preroll 120 2>/dev/null | tail -1 > $O/r02_base_mixed1024.json
python tools/time_scene.py mixed:8192 --steps 50 --preroll 120 2>/dev/null | tail -1 > $O/r02_base_mixed8192.json
EP_SOLVE_SERIAL=1 python tools/solve_classes.py 8192 2>/dev/null | tail -1 > $O/r02_classes_8192.json
EP_SOLVE_SERIAL=1 python tools/solve_classes.py 1024 2>/dev/null | tail -1 > $O/r02_classes_1024.json
python tools/time_scene.py cars:8192 --steps 30 --preroll 60 2>/dev/null | tail -1 > $O/r02_base_cars8192.json
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_scenes.py 4 > $O/r02_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> $O/r02_sanitizer_$tool.txt
  tail -3 $O/r02_sanitizer_$tool.txt
done
cat $O/r02_toolchain_probe.txt
cat $O/r02_base_mixed1024.json $O/r02_classes_1024.json $O/r02_classes_8192.json
