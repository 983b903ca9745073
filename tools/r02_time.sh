#!/bin/bash
# timing of the standard scenes on one box: tools/r02_time.sh [tag]
tag=${1:-run}
for sc in mixed:1024 mixed:8192 cars:8192; do
  python tools/time_scene.py $sc --steps 50 --preroll 120 2>/dev/null | tail -1 > gpurun_out/time_${tag}_${sc/:/}.json
  python - <<PY
import json
d=json.load(open("gpurun_out/time_${tag}_${sc/:/}.json")); k=d["kernels_ms"]
print("$sc step %.3f | " % d["ms_per_step"] + " ".join("%s %.3f" % (n.replace("k_","").replace("solve_",""), v) for n,v in k.items()))
PY
done
