#!/bin/bash
# A/B of library variants built by tools/build_variant.sh: quick_bench on each scratch/lib_<name>.so, best kernel ms
for n in "$@"; do
  printf "%-12s " $n
  ORPCD_B200_LIB=/root/repo/scratch/lib_$n.so REPS=${REPS:-5} python tools/quick_bench.py 2>/dev/null | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin if l.startswith('{')]
print('kernel ms min %.3f  wall ms min %.3f  pairs %.4g  rmse_min %.12g iters %.4f' % (min(x['sweep_ms'] for x in r[1:]), 1e3*min(x['wall_s'] for x in r[1:]), r[-1]['pairs_visited'], r[-1]['rmse_min'], r[-1]['mean_iters']))"
done
