"""Probe the TMA pipeline scan: stages sweep, background-only vs dense data."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L, plugins as P, synth
dev = torch.device("cuda", 0)
lib = L.load()
B, R, K = 32, 4, 200
def timeit(fused, ss):
    for i in range(10): fused.enqueue_scan(B, ss[i % R])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): fused.enqueue_scan(B, ss[i % R])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
for name, nobj in (("dense(64 obj)", 64), ("background only", 0)):
    sets = [[torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=i, n_obj=nobj)] for i in range(R)]
    plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32))
    fused = P.FusedYoloDecodeNms(plug, B, device=dev)
    for stages in (5, 8, 10, 12, 15):   # (the copy-only mode of round 1 lives on in tools/tma_bench.cu)
        plug.tune(tma=1, stages=stages)
        print(json.dumps({"data": name, "stages": stages, "us": round(timeit(fused, sets), 2)}), flush=True)
    plug.tune(slices=4, rows=5)
    print(json.dumps({"data": name, "register_kernel_4x5_us": round(timeit(fused, sets), 2)}), flush=True)
