#!/bin/bash
for so in tools/var/libwva_d2.so tools/var/libwva_d4.so; do
cp $so workload_variant_autoscaler_b200/libwva_b200.so
echo "== $so"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "size or analyze or streaming or golden" 2>&1 | tail -1
python tools/bench_configs.py 2>&1 | grep -E '"config": (4|5)' | cut -c1-170
done
