#!/bin/bash
# run bench.py against every libsmc variant in profiles/micro/variants (tuning sweeps)
for f in profiles/micro/variants/libsmc_*.so; do
  n=$(basename $f .so)
  SMC_LIB_PATH=$PWD/$f python bench.py --steps 50 --warmup 3 --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print('$n', round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernel_ms'].items()})
    elif line.strip(): print('$n', line.strip()[:200])
"
done
