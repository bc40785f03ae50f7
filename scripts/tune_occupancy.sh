#!/bin/bash
# Sweep the resident-CTAs-per-SM knob of the step kernel (run on the GPU box; writes gpurun_out/tune_*.json).
mkdir -p gpurun_out
for c in ${@:-0 5 6 7 8}; do
  DBSGYM_CTAS_PER_SM=$c python bench.py --steps 30 --warmup 3 > gpurun_out/tune_$c.json 2>gpurun_out/tune_$c.err
  python -c "
import json; d=json.load(open('gpurun_out/tune_$c.json')); print('ctas/sm=$c', 'step_ms', round(d['roofline']['kernel_ms'],4), 'obs_ms', round(d['roofline_obs']['kernel_ms'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
