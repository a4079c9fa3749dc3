#!/bin/bash
# last visit of a round with little budget left: bench, one full ncu capture, launch list, reference arm (in that order)
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tsc_step -s 450 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 452 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 100 python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 300 gpurun_out/bench_ref.json
