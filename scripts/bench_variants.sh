#!/bin/bash
# A/B the tuning variants of libdbsgym (built by hand with -D flags) on the GPU box.
mkdir -p gpurun_out
for n in "$@"; do
  export DBSGYM_LIB=$PWD/dbsgym_b200/csrc/libdbsgym_$n.so
  python bench.py --steps 60 --warmup 3 > gpurun_out/var_$n.json 2>gpurun_out/var_$n.err
  python -c "
import json; d=json.load(open('gpurun_out/var_$n.json')); print('$n', 'step_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
export DBSGYM_LIB=$PWD/dbsgym_b200/csrc/libdbsgym_fast.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
