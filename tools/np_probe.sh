for cfg in "32 32" "16 32" "8 32" "16 16" "8 16" "32 16"; do set -- $cfg
  EP_NP_LANES_G0=$1 EP_NP_LANES_G1=$2 python tools/time_scene.py mixed:8192 --steps 40 --preroll 120 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('G0=$1 G1=$2 step %.3f np %.3f' % (d['ms_per_step'], k['k_narrowphase']))"
done
