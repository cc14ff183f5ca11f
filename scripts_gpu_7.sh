#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
cat > /tmp/nms_ab.py <<'PY'
import sys, torch, json
sys.path.insert(0, '.')
from multipoint_b200 import ops
dev = torch.device('cuda')
g = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn((128, 65, 64, 80), generator=g, device=dev) * 2.0
logits[:, 64] += 5.0
prob = ops.detector_head(logits).reshape(128, 512, 640)
def timed(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
print(json.dumps({"nms_dense_ms": timed(lambda: ops.box_nms(prob, 4, 0.015)),
                  "nms_topk_kp_ms": timed(lambda: ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048))}))
PY
MP_NMS_TILE=0 python /tmp/nms_ab.py > gpurun_out/nms_tile0.json 2>&1
MP_NMS_TILE=1 python /tmp/nms_ab.py > gpurun_out/nms_tile1.json 2>&1
MP_NMS_TILE=1 timeout -k 10 600 python -m pytest tests -q -m gpu -k "box_nms" 2>&1 | tail -3 > gpurun_out/tests_tile1.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?" >> gpurun_out/bench_n1.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
for f in gpurun_out/tests_gpu.log gpurun_out/tests_tile1.log gpurun_out/bench_n1.err; do echo "--- $f"; tail -n 4 $f; done
cat gpurun_out/nms_tile0.json gpurun_out/nms_tile1.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['hot_path']['ms_per_step']); print({k:round(v['ms'],4) for k,v in d['hot_path']['stages'].items()})"
