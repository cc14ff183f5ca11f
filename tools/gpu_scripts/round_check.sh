#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the hot path and a full
# ncu capture of this repo's kernels.  Run through:  gpurun --timeout 2400 -- 'bash tools/gpu_scripts/round_check.sh'
# Outputs land in gpurun_out/; summarise them into profiles/ with profiles/summarize.py.
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?"; tail -3 gpurun_out/bench_n1.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout -k 10 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
# launch list of the hot path at the full batch size (cold-cache, serialised: compare shares)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
# full capture of one warm hot-path step (every kernel of this library)
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
    -k regex:"nms_|detector_head_kernel|normalize_desc|sample_descriptors|match_" \
    -s 57 -c 19 -o gpurun_out/prof_hot python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
# dram traffic of the captured kernels
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['hot_path']['ms_per_step']); print(json.dumps(d['roofline']))
for r in d['hot_path']['kernels']: print(r)"
# the fused backbone glue kernel at the two full-resolution layer shapes (ncu --set full)
cat > gpurun_out/glue_ncu.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from multipoint_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn((64, 64, 512, 640), generator=g, device='cuda')
sc = torch.rand(64, device='cuda') + 0.5
sh = torch.randn(64, device='cuda')
for pool in (False, True):
    for _ in range(3):
        ops.relu_bn_pad(x, sc, sh, pool=pool, pad=1, reflect=True, conv_bias=sh)
img = torch.rand((64, 1, 512, 640), generator=g, device='cuda')
w = torch.randn((64, 1, 3, 3), generator=g, device='cuda')
for _ in range(2):
    ops.conv1_relu_bn_pad(img, w, sh, sc, sh)
torch.cuda.synchronize()
PY
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"relu_bn_pad" -s 2 -c 6 -o gpurun_out/prof_glue python gpurun_out/glue_ncu.py > gpurun_out/ncu_glue.log 2>&1
tail -2 gpurun_out/ncu_glue.log
