"""Where does the scan kernel's time go?  us/launch vs batch (fixed cost + slope), stream launches vs one CUDA graph of
K launches, dense vs background heads, fp32 vs fp16."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L, plugins as P, synth
dev = torch.device("cuda", 0)
K = 100
def timeit(fn, n):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
base = {}
for nobj in (64, 0):
    base[nobj] = [synth.yolov8_heads(32, seed=i, n_obj=nobj) for i in range(4)]
for B in (8, 16, 32, 64, 128):
    for nobj in (64, 0):
        R = max(2, 128 // B)
        # build B-image sets by tiling the 32-image seeds
        sets = []
        for r in range(min(R, 4) if B <= 32 else 2):
            hs = base[nobj][r % 4]
            if B <= 32:
                sets.append([torch.from_numpy(h[:B]).to(dev) for h in hs])
            else:
                sets.append([torch.from_numpy(h).to(dev).repeat(B // 32, 1, 1).contiguous() for h in hs])
        for dt in (L.F32, L.F16):
            ss = sets if dt == L.F32 else [[h.half() for h in s] for s in sets]
            plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32), in_dtype=dt)
            fused = P.FusedYoloDecodeNms(plug, B, device=dev)
            n = len(ss)
            us = timeit(lambda i: fused.enqueue_scan(B, ss[i % n]), K)
            # one graph holding 20 back-to-back launches
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                g = torch.cuda.CUDAGraph()
                fused.enqueue_scan(B, ss[0]); torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    for i in range(20): fused.enqueue_scan(B, ss[i % n])
                usg = timeit(lambda i: g.replay(), 10) / 20
            nb = sum(h.numel() * h.element_size() for h in ss[0])
            print(json.dumps({"B": B, "n_obj": nobj, "dtype": "f32" if dt == L.F32 else "f16", "us_stream": round(us, 2),
                              "us_graph": round(usg, 2), "GBps_graph": round(nb / usg / 1e3, 1)}), flush=True)
        del sets
        torch.cuda.empty_cache()
