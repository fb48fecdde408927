#!/bin/bash
# A/B of blocked-WY variants (tools/build_variant.py): parity tests of the WY shapes, then config 5 per size class
for v in default $(ls tools/variants/*.so 2>/dev/null); do
  if [ "$v" = default ]; then unset QRKIT_B200_LIB; else export QRKIT_B200_LIB=$PWD/$v; fi
  echo "== $v"
  python -m pytest tests/test_block_diagonal_gpu.py -m gpu -q -x -k "wy or mixed or config5" 2>&1 | tail -1
  python bench_extra.py --workload classes --no-cpu --steps 10 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(' '.join('%dx%d:%.3f' % (c['block'][0], c['block'][1], c['ms']) for c in d['classes']), 'total %.3f ms' % d['total_ms'])"
done
