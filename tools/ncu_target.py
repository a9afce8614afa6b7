"""A short, deterministic launch sequence for ncu: rotating input sets (> L2), the three step kernels launched back to back.
usage: python tools/ncu_target.py [reps]   (12 launches of each kernel by default)"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L, plugins as P, synth
dev = torch.device("cuda", 0)
B, R = 32, 4
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dt = L.F16 if "--f16" in sys.argv else L.F32
sets = [[torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=i)] for i in range(R)]
if dt == L.F16:
    sets = [[h.half() for h in s] for s in sets]
frames = [torch.from_numpy(synth.frames(B, seed=i)).to(dev) for i in range(R)]
dst = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
plans = [P.PreprocessPlan(list(f.unbind(0)), dst, 640, 640) for f in frames]
plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32), in_dtype=dt)
fused = P.FusedYoloDecodeNms(plug, B, device=dev)
for i in range(reps):
    plans[i % R].enqueue()
    fused.enqueue(B, sets[i % R])
torch.cuda.synchronize()
print("done")
