#!/bin/bash
# A/B of library builds on the stand-alone multigrid benchmark: tools/ab_mg.sh OUT SIZES lib1 lib2 ...  (two interleaved passes)
out=$1; sizes=$2; shift 2
: > $out
for pass in 1 2; do
  for lib in "$@"; do
    echo "## $lib pass $pass" >> $out
    PINC_B200_LIB=$PWD/$lib timeout 300 python tools/mg_bench.py $sizes noise ${MGMODE:-2} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l); print(r['N'], r['vcycles'], round(r['us_per_vcycle'], 1))
    except Exception: print(l.rstrip()[:200])" >> $out
  done
done
