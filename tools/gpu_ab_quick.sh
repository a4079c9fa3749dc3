#!/bin/bash
# like gpu_ab.sh but with a chosen subset of the parity tests: TESTS="tests/test_golden_gpu.py ..." tools/gpu_ab_quick.sh "VAR=1" ...
mkdir -p gpurun_out
timeout 900 python -m pytest ${TESTS:-tests/test_golden_gpu.py tests/test_engine_gpu.py::test_kernel_variants_agree_with_oracle} -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
SKIP_TESTS=1 bash tools/gpu_ab.sh "$@"
