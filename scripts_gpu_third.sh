#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -60 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?" >> gpurun_out/bench_n1.err
# launch list of the hot path at full batch size (cold-cache, serialised: shares only)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
echo "== ncu hot launches exit $?" >> gpurun_out/ncu_launch_hot.log
# launch list of the whole step (backbone included) at a reduced batch so ncu's save/restore stays cheap
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_pairs4.csv \
    python bench.py --only-value --pairs 4 --steps 1 --warmup 3 > gpurun_out/ncu_launch_step.log 2>&1
echo "== ncu step launches exit $?" >> gpurun_out/ncu_launch_step.log
# full capture of this repo's kernels in one warm hot-path step
timeout -k 10 600 ncu --set full --clock-control none --import-source on \
    -k regex:"nms_|detector_head_kernel|normalize_desc|sample_descriptors|match_" \
    -s 45 -c 15 -o gpurun_out/prof_hot python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $?" >> gpurun_out/ncu_full.log
for f in gpurun_out/tests_gpu.log gpurun_out/bench_n1.err gpurun_out/ncu_launch_hot.log gpurun_out/ncu_launch_step.log gpurun_out/ncu_full.log; do echo "--- $f"; tail -n 4 $f; done
cat gpurun_out/bench_n1.json
