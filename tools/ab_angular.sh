#!/bin/bash
# A/B of the block-angular K1 tuning variants (tools/build_variant.py): config 3 at 1M points, one JSON summary line per variant
for v in default $(ls tools/variants/*.so 2>/dev/null); do
  if [ "$v" = default ]; then unset QRKIT_B200_LIB; else export QRKIT_B200_LIB=$PWD/$v; fi
  python bench_extra.py --workload angular --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$v', round(d['ms_per_step']*1e3, 2), 'us  colpiv-left', round(d['colpiv_left']['ms_per_step']*1e3, 2), 'us  frac', round(d['roofline']['frac'], 3))"
done
