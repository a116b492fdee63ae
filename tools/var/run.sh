#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
python tools/one_chain.py | tail -3
python tools/timeline.py | grep -v "^t=" | egrep "kernel ms|finish|warps      0|warps    200|warps   1000|warps   5000"
python bench.py --steps 30 --warmup 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
