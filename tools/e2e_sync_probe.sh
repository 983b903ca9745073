#!/bin/bash
# e2e of one shard with blocking vs spinning host waits, and the frame trace: tools/e2e_sync_probe.sh  (run on the GPU box)
run() {  # worlds contexts label env...
  W=$1; K=$2; L=$3; shift 3
  env "$@" python bench.py --worlds $W --e2e-contexts $K --no-cars --no-pile --no-cpu-baseline --steps 50 --e2e-steps 40 2>/tmp/err.txt | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$W', '$K', '$L', 'value %.1fM ms %.3f e2e %.1fM every %.1fM single %.1fM' % (d['value']/1e6, d['ms_per_step'], e['value']/1e6, e['every_array_value']/1e6, e['single_context_every_array_value']/1e6))"
  grep -i "frame\|trace" /tmp/err.txt | tail -6
}
run 1024 1 block X=1
run 1024 1 spin EP_BLOCKING_SYNC=0
run 1024 1 trace EP_TRACE_FRAME=1
run 8192 2 block X=1
run 8192 2 spin EP_BLOCKING_SYNC=0
run 8192 1 spin EP_BLOCKING_SYNC=0
