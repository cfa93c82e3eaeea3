#!/bin/bash
# developer probe: constrained-assign time vs selection slack
for s in ${@:-16 24 32}; do
  RC_SINKHORN_SLACK=$s python tools/quick_bench.py assign 2>&1 | grep "M=48" | sed "s/^/slack $s: /"
  RC_SINKHORN_SLACK=$s python tools/prof_assign.py 50 2>/dev/null | grep lists
done
