#!/bin/bash
# A/B of frozen builds: usage r2_ab.sh "<tag>[:opts] ..." [cases...]   -> gpurun_out/ab_<stamp>.txt
set -u
variants=$1; shift
out=gpurun_out/ab_$(date +%H%M%S).txt
for v in $variants; do
  tag=${v%%:*}; opts=""; [ "$tag" != "$v" ] && opts=${v#*:}
  echo "== $v" | tee -a $out
  RTG_LIB_DIR=$PWD/build_variants/$tag RTG_OPTS=$opts timeout 300 python scripts/quick_perf.py "$@" 2>&1 | grep -E "case|rror" | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['case'].ljust(28), 'ms', d['ms'], 'Mrays/s', d['Mrays/s'], 'frac', d['frac_74.4'], 'fill', d['fill'], 'exact/q', d['exact_per_query'], 'grid', d['grid'], 'smem', d['smem'])
" | tee -a $out
done
