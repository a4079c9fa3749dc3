#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_episode_gpu.py -m gpu -x -q > gpurun_out/pytest_episode.log 2>&1; tail -15 gpurun_out/pytest_episode.log
timeout 1500 bash tools/gpu_sanitize.sh > gpurun_out/sanitize.log 2>&1; cat gpurun_out/sanitizer/summary.txt
python -c "import importlib.util; print('cityflow importable:', importlib.util.find_spec('cityflow') is not None)"
nproc; lscpu | head -20
