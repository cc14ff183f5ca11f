#!/bin/bash
# Round-2 evidence run (one GPU-box visit):  gpurun --timeout 2400 -- 'bash tools/gpu_scripts/round2_evidence.sh'
# Outputs land in gpurun_out/; profiles/summarize.py turns the ncu reports into the tracked CSVs.
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_tests_gpu.log
tail -2 gpurun_out/r2_tests_gpu.log
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "== bench exit $?"; tail -c 300 gpurun_out/r2_bench_n1.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout -k 10 600 python tools/bench_match.py --sweep --out gpurun_out/r2_match.json > gpurun_out/r2_match.log 2>&1
timeout -k 10 600 python tools/bench_adapt.py --out gpurun_out/r2_adapt.json > gpurun_out/r2_adapt.log 2>&1
timeout -k 10 600 python tools/bench_nms.py --out gpurun_out/r2_nms.json > gpurun_out/r2_nms.log 2>&1
timeout -k 10 900 python tools/bench_kernels.py --out gpurun_out/r2_kernels.json > gpurun_out/r2_kernels.log 2>&1
# launch list of the hot path at the full batch size (cold-cache, serialised: compare shares)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_hot.csv \
    python bench.py --only-hot --steps 2 --warmup 3 > gpurun_out/r2_ncu_launch_hot.log 2>&1
# full capture of one warm hot-path step (every kernel of this library): 15 launches per step, 3 warm-up steps
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
    -k regex:"nms_|detector_head_kernel|normalize_desc|sample_descriptors|match_" \
    -s 45 -c 15 -o gpurun_out/r2_prof_hot python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/r2_ncu_full.log 2>&1
# adaptation kernels on the reference's homography distribution
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"warp_kernel|ha_aggregate|valid_mask" -s 12 -c 4 \
    -o gpurun_out/r2_prof_adapt python tools/bench_adapt.py --iters 2 > gpurun_out/r2_ncu_adapt.log 2>&1
# the dense NMS tile kernel (keep_top_k = 0) at 12.9 % candidates
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"nms_tile_fast_kernel" -s 3 -c 1 \
    -o gpurun_out/r2_prof_nms_dense python tools/bench_nms.py --only "dense 12" --iters 1 > gpurun_out/r2_ncu_nms.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d['hot_path']['ms_per_step'])
print(json.dumps(d['roofline']))
PY
