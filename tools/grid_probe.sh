for hg in "0,0,0,0,0,0" "0,0,2,0,0,0" "0,0,1,0,0,0" "4,3,2,0,0,0" "3,2,1,0,0,0" "2,2,2,4,2,1" "0,0,0,6,3,1" "6,4,2,6,4,2"; do
  EP_HEAD_GRID=$hg python tools/time_scene.py mixed:8192 --steps 40 --preroll 120 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels_ms']; print('HG=$hg step %.3f solve %.3f b0 %.2f b1 %.2f b2 %.2f b3 %.2f b4 %.2f b5 %.2f tw %.2f' % (d['ms_per_step'], k['k_solve'], k.get('k_solve_lanes_b0',0), k['k_solve_lanes_b1'], k['k_solve_lanes_b2'], k['k_solve_lanes_b3'], k['k_solve_lanes_b4'], k['k_solve_lanes_b5'], k['k_solve_team_warp']))"
done
