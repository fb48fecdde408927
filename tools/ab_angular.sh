#!/bin/bash
# A/B of the block-angular kernels for 2x1 blocks: the default library (direct K1 + direct K3), its environment switches
# (QRK_ANG_STAGED=1: staged K1; QRK_ANG_K3_TILED=1: one-point-per-thread K3) and the variant libraries in tools/variants
run() { python bench_extra.py --workload angular --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$1', round(d['ms_per_step']*1e3, 2), 'us  colpiv-left', round(d['colpiv_left']['ms_per_step']*1e3, 2), 'us  frac', round(d['roofline']['frac'], 3))"; }
for v in default $(ls tools/variants/*.so 2>/dev/null); do
  if [ "$v" = default ]; then unset QRKIT_B200_LIB; else export QRKIT_B200_LIB=$PWD/$v; fi
  unset QRK_ANG_STAGED QRK_ANG_K3_TILED; run "$v"
  if [ "$v" = default ] && [ -n "$AB_SWITCHES" ]; then
    export QRK_ANG_K3_TILED=1; run "$v K3tiled"; unset QRK_ANG_K3_TILED
    export QRK_ANG_STAGED=1; run "$v K1staged"; unset QRK_ANG_STAGED
  fi
done
