#!/bin/bash
# usage: bench_quick.sh <slots> [env...]  -> one line summary
s=$1; shift
env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --batch 96 --slots $s --cpu-sample 0 --e2e-queries 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; q=d['per_query']
print('slots=$s $*  value %.1f q/s e2e %.1f | push phase %.2f ms (kernel %.2f) walk %.2f ms | push %.1f G edges/s walk %.1f G hops/s' % (d['value'], d['e2e']['value'], q['push_phase_ms'], r['push']['kernel_ms']/(96*2), q['walk_phase_ms'], r['push']['edges_per_s']/1e9, r['walk']['steps_per_s']/1e9))"
