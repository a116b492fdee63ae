#!/bin/bash
for so in tools/var/libwva_w16_t1.so tools/var/libwva_w20_t1.so tools/var/libwva_w24_t1.so; do
cp $so workload_variant_autoscaler_b200/libwva_b200.so
echo "== $so"
python tools/one_chain.py | tail -1
python bench.py --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
done
