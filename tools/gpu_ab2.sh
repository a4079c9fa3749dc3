#!/bin/bash
mkdir -p gpurun_out
TSC_B200_LIB=$PWD/tools/ab/lib_nfrc.so timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_episode_gpu.py tests/test_edge_gpu.py -m gpu -x -q > gpurun_out/pytest_engine.log 2>&1; tail -5 gpurun_out/pytest_engine.log
AB_LIBS="${AB_LIBS:-sb nf nfrc rc}" bash tools/gpu_ab_core.sh
