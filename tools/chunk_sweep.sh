#!/bin/bash
# throughput vs head chunk size (wave quantisation of the persistent conv kernels): tools/chunk_sweep.sh 256 259 ...
for c in "$@"; do
  python bench.py --no-cpu-baseline --chunk $c --e2e-chunk $c 2>/dev/null > /tmp/chunk_$c.json
  python -c "
import json,sys; d=json.load(open('/tmp/chunk_$c.json')); print('chunk', $c, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'], 3))"
done
