#!/bin/bash
# dev tool: bench kernel_ms for each prebuilt library variant under build/var/
for f in build/var/*.so; do
  echo "== $f"
  TAB200_LIB=$PWD/$f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
