#!/bin/bash
# prints value + breakdown of a short bench run (developer A/B helper)
python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-ext-microbench --no-sam-stage "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); b=d['breakdown_ms']
print('pairs/s %.3gM e2e %.3gM | smem %.1f seedpos %.1f chain %.1f c2a %.1f step %.1f e2e_step %.1f | ext GCUPS %.0f' % (d['value']/1e6, d['e2e']['value']/1e6, b['smem_stage'], b['seedpos_stage'], b['device_chaining'], b['c2a_stage'], b['step'], b['e2e_step'], d['roofline_ext']['achieved']))"
