#!/bin/bash
# One gpurun call that evaluates the dense-build variants (DD_DENSE_V): parity first (the dense / end-to-end tests of the
# suite under each variant), then the stand-alone timing at c3.  Variant 4 (one shared-memory row buffer per warp, bulk
# store, 16-18 warps per SM) was written after round 1's GPU budget was spent and has never run: start here in round 2.
#   gpurun --timeout 900 -- 'bash scripts/gpu_dense_variants.sh'
mkdir -p gpurun_out
for v in 4 1 3; do
    for w in ${DENSE_WARPS:-16 18 12}; do
        [ "$v" != 4 ] && [ "$w" != 16 ] && continue
        tag="v${v}_w${w}"
        wenv=""; [ "$v" = 4 ] && wenv="DD_DENSE_WARPS=$w"
        echo "=== DD_DENSE_V=$v $wenv: parity" | tee -a gpurun_out/dense_variants.log
        env DD_DENSE_V=$v $wenv timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
            -k "dense or doublets or end_to_end or wide_matrix or pipeline_matches" 2>&1 | tail -3 | tee -a gpurun_out/dense_variants.log
        echo "=== DD_DENSE_V=$v $wenv: c3 timing" | tee -a gpurun_out/dense_variants.log
        env DD_DENSE_V=$v $wenv timeout 300 python scripts/dense_bench.py c3 6 2>&1 | tail -1 | tee -a gpurun_out/dense_variants.log
    done
done
