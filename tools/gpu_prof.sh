#!/bin/bash
# one full ncu capture of the step kernel in the busy part of the episode
mkdir -p gpurun_out
CAP=${CAP:-1024}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsc_step -s 450 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 460 --warmup 5 --no-cpu-baseline --vehicle-capacity $CAP > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
