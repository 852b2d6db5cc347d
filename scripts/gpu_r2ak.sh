#!/bin/bash
# round 2, session ak: the fit loop after the worker-count rule (pinned-memory budget instead of fixed caps)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pipeline or default_constructor or end_to_end" 2>&1 | tail -3 | tee gpurun_out/r2ak_tests.log
