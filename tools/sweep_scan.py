"""Sweep the (class slices, unroll) variants of yolo_v8_scan_kernel at b32 and print us/launch + GB/s.
Run on the GPU box:  python tools/sweep_scan.py"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L  # noqa: E402
from tensorrtx_b200 import plugins as P  # noqa: E402
from tensorrtx_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
lib = L.load()
B, R, K = 32, 4, 200
sets = [[torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=i)] for i in range(R)]
sets16 = [[h.half() for h in s] for s in sets]
nbytes = sum(h.numel() * 4 for h in sets[0])
res = []
for dt, ss, nb in ((L.F32, sets, nbytes), (L.F16, sets16, nbytes // 2)):
    plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32), in_dtype=dt)
    fused = P.FusedYoloDecodeNms(plug, B, device=dev)
    for slices, unroll in [(-1, 15), (1, 8), (1, 16), (2, 5), (2, 10), (2, 20), (4, 5), (4, 10), (4, 20), (8, 5)]:
        if slices < 0:   # TMA pipeline kernel, `unroll` = cap on stages (= consumer warps)
            plug.tune(tma=1, stages=unroll)
        else:
            plug.tune(slices=slices, rows=unroll)
        for i in range(10):
            fused.enqueue_scan(B, ss[i % R])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fused.enqueue_scan(B, ss[i % R])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / K * 1e3
        r = {"dtype": "f32" if dt == L.F32 else "f16", "slices": slices, "unroll": unroll, "us": round(us, 2),
             "GBps": round(nb / us / 1e3, 1)}
        res.append(r)
        print(json.dumps(r), flush=True)
# NMS alone and preprocess alone
plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32))
fused = P.FusedYoloDecodeNms(plug, B, device=dev)
fused.enqueue_scan(B, sets[0])
for name, fn in (("nms", lambda i: fused.enqueue_nms(B, sets[0])), ("fused", lambda i: fused.enqueue(B, sets[i % R]))):
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"kernel": name, "us": round(e0.elapsed_time(e1) / K * 1e3, 2)}), flush=True)
for odt in (torch.float32, torch.float16):
    frames = [torch.from_numpy(synth.frames(B, seed=i)).to(dev) for i in range(R)]
    dst = torch.empty((B, 3, 640, 640), dtype=odt, device=dev)
    plans = [P.PreprocessPlan(list(f.unbind(0)), dst, 640, 640) for f in frames]
    for i in range(10):
        plans[i % R].enqueue()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        plans[i % R].enqueue()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / K * 1e3
    nb = B * (640 * 640 * 3 + 3 * 640 * 640 * dst.element_size())
    print(json.dumps({"kernel": "letterbox", "out": str(odt), "us": round(us, 2), "GBps": round(nb / us / 1e3, 1)}), flush=True)
