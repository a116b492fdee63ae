#!/bin/bash
for so in tools/var/libwva_rev*.so; do
cp $so workload_variant_autoscaler_b200/libwva_b200.so
echo "== $so"
python tools/bench_configs.py 2>&1 | grep -E '"config": (4|5)' | cut -c1-160
done
