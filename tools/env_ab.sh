#!/bin/bash
# A/B of an environment switch on one box: tools/env_ab.sh <scene> <VAR> <value>...   (alternated twice)
scene=$1; var=$2; shift 2
for round in 1 2; do for v in "$@"; do
  env $var=$v python tools/time_scene.py $scene --steps 40 --preroll 120 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms']
print('$var=$v', '$scene', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n.replace('k_','').replace('solve_',''), k.get(n, 0.0)) for n in ['k_narrowphase','k_equations','k_islands','k_solve','k_solve_lanes_b5','k_solve_team_warp','k_solve_team_cta']))"
done; done
