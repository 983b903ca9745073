for K in 2 3 4; do
  for S in auto lanes; do
    if [ $S = lanes ]; then export EP_SOLVER=lanes; else unset EP_SOLVER; fi
    python bench.py --steps 20 --warmup 3 --no-cars --no-pile --no-cpu-baseline --e2e-steps 20 --e2e-contexts $K 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('K=$K solver=$S value %.1f M  e2e %.1f M  every %.1f M' % (d['value']/1e6, e['value']/1e6, e['every_array_value']/1e6))"
  done
done
