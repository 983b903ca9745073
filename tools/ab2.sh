#!/bin/bash
# A/B of library variants over several scenes: tools/ab2.sh "<scene>:<preroll> ..." <variant>...   (variants = elm_physics_b200/csrc/variants/lib<NAME>.so)
scenes=$1; shift
cd "$(dirname "$0")/../elm_physics_b200/csrc"
cp libelm_physics_b200.so /tmp/keep.so
for round in 1 2; do for sp in $scenes; do sc=${sp%@*}; pre=${sp#*@}; for v in "$@"; do
  cp variants/lib$v.so libelm_physics_b200.so
  python ../../tools/time_scene.py $sc --preroll $pre 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms']
print('$sc', '$v', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n.replace('k_',''), k.get(n, 0.0)) for n in ['k_broadphase','k_narrowphase','k_equations','k_islands','k_solve']))"
done; done; done
cp /tmp/keep.so libelm_physics_b200.so
