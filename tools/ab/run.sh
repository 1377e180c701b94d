#!/bin/bash
# A/B of two builds of libvstb200.so (developer tool): alternates runs to average out clock / power drift
cd "$(dirname "$0")/../.."
cp vstnet_b200/libvstb200.so /tmp/lib_keep.so
for rep in 1 2 3; do
  for v in A B; do
    cp tools/ab/lib$v.so vstnet_b200/libvstb200.so
    python tools/quick_time.py photo 1080 1920 f16x2 2>&1 | head -1 | sed "s/^/$v: /"
  done
done
cp /tmp/lib_keep.so vstnet_b200/libvstb200.so
