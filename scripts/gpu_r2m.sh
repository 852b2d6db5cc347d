#!/bin/bash
# round 2, session m: full GPU suite, full bench line (cpu_baseline, parity_c2, c2), ncu launch list + --set full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2m_tests.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2m_clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err
kill $SMI
tail -c 1500 gpurun_out/r2m_bench_c3.json
# launch list of the pipelined loop (2 iterations per call, 3 calls; per-launch times are cold-cache and serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2m_launches.csv python scripts/one_iter.py c3 2 > gpurun_out/r2m_ncu_launches.log 2>&1
# --set full of the hot kernels, stage by stage (second repetition)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_tc_gemm|k_dense_rows|k_km_assign|k_lists_other|k_jacobi" -s 12 -c 12 -o gpurun_out/r2m_full python scripts/stage_bench.py c3 2 > gpurun_out/r2m_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
