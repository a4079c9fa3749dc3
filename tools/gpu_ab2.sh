#!/bin/bash
mkdir -p gpurun_out
TSC_B200_LIB=$PWD/tools/ab/lib_xp.so timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "lockstep or oracle" > gpurun_out/pytest_engine.log 2>&1; tail -3 gpurun_out/pytest_engine.log
AB_LIBS="${AB_LIBS:-cur xp}" bash tools/gpu_ab_core.sh
