#!/bin/bash
# round 2, session ak: the fit-loop tests after the worker-count rule (pinned-memory budget instead of fixed caps)
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_leiden.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r2ak_tests.log
