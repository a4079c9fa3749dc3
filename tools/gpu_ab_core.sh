#!/bin/bash
# "lib name" runs of the short bench, twice each, interleaved (libs under tools/ab/)
mkdir -p gpurun_out; : > gpurun_out/ab.txt
run() {
  name=$1; envs=$2; shift 2
  ( env $envs timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu-baseline "$@" 2> gpurun_out/ab_$name.err ) | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['config']['kernel']
    print('$name: dev ms %.4f  e2e ms %.4f  match %s  nt %d x %d regs %d smem %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['matches_device_leg'], k['threads'], k['blocks_per_sm'], k['regs'], k['smem_bytes']))
except Exception as e:
    print('$name failed', e)
" | tee -a gpurun_out/ab.txt
}
for rep in 1 2; do
for lib in $AB_LIBS; do
run $lib "TSC_B200_LIB=$PWD/tools/ab/lib_$lib.so"
done
done
