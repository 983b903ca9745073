#!/bin/bash
# A/B timing of library variants on one box: tools/ab.sh <scene> <variant>...   (variants = elm_physics_b200/csrc/variants/lib<NAME>.so)
scene=$1; shift
cd "$(dirname "$0")/../elm_physics_b200/csrc"
cp libelm_physics_b200.so /tmp/keep.so
for round in 1 2; do for v in "$@"; do
  cp variants/lib$v.so libelm_physics_b200.so
  python ../../tools/time_scene.py $scene 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms']
print('$v', 'step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n.replace('k_',''), k.get(n, 0.0)) for n in ['k_broadphase','k_narrowphase','k_equations','k_islands','k_tile_build','k_solve']))"
done; done
cp /tmp/keep.so libelm_physics_b200.so
