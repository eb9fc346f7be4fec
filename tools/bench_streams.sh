#!/bin/bash
# dev: headline value with one stream vs one stream per eye, both math modes
for s in 2 4 8; do for m in strict fast; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --streams $s --math $m 2>/dev/null | tail -1 > /tmp/b.json
  python -c "import json; d=json.load(open('/tmp/b.json')); print('streams', $s, '$m', round(d['value'],1), d.get('masked_r0.5'), d['clocks']['sm_mhz'])"
done; done
