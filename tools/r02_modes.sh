#!/bin/bash
# lane tiles against the quad family at several batch sizes (same box): picks the EP_QUADS_MAX_BODIES threshold
for w in 512 1024 2048 4096 8192; do for m in lanes quads; do
  EP_SOLVER=$m python tools/time_scene.py mixed:$w --steps 40 --preroll 120 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels_ms']
print('mixed:$w $m step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (n.replace('k_','').replace('solve_',''), k[n]) for n in k if n in ('k_solve','k_tile_build','k_equations','k_islands','k_narrowphase')))"
done; done
for m in lanes quads; do EP_SOLVER=$m python tools/time_scene.py cars:4096 --steps 40 --preroll 60 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cars:4096 $m step %.3f solve %.3f' % (d['ms_per_step'], d['kernels_ms']['k_solve']))"; done
for s in boxes pile:2000; do for m in lanes quads; do EP_SOLVER=$m python tools/time_scene.py $s --steps 30 --preroll 300 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$s $m step %.3f solve %.3f' % (d['ms_per_step'], d['kernels_ms']['k_solve']))"; done; done
