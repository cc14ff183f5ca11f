#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?" >> gpurun_out/bench_n1.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "== ref exit $?" >> gpurun_out/bench_ref.err
# launch list of one warm step (cold-cache, serialised: shares only)
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --only-value --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $?" >> gpurun_out/ncu_launch.log
# full capture of this repo's kernels in one warm step
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
    -k regex:"nms_tile|nms_select|nms_fixup|detector_head_kernel|normalize_desc|sample_descriptors|match_top2_tc|match_prep|match_recheck|match_decide|match_compact|match_flag" \
    -s 42 -c 14 -o gpurun_out/prof_r1 python bench.py --only-value --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $?" >> gpurun_out/ncu_full.log
tail -3 gpurun_out/tests_gpu.log gpurun_out/bench_n1.err gpurun_out/bench_ref.err gpurun_out/ncu_launch.log gpurun_out/ncu_full.log
cat gpurun_out/bench_n1.json gpurun_out/bench_ref.json
