#!/bin/bash
# A/B of two builds under SUSTAINED load (power cap / clocks matter): bench.py value over 150 steps, alternating
cd "$(dirname "$0")/../.."
cp vstnet_b200/libvstb200.so /tmp/lib_keep.so
for rep in 1 2; do
  for v in A B; do
    cp tools/ab/lib$v.so vstnet_b200/libvstb200.so
    python bench.py --steps 150 --warmup 3 --no-cpu --no-images 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['value'],2), 'fps  e2e', round(d['e2e']['value'],2), ' sm_mhz', d['clocks']['sm_mhz'])"
  done
done
cp /tmp/lib_keep.so vstnet_b200/libvstb200.so
