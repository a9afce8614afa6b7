"""Phase timing of nms_kernel from clock64 stamps (thread 0 of every CTA)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L, plugins as P, synth
dev = torch.device("cuda", 0)
lib = L.load()
B = 32
heads = [torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=0)]
plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32))
fused = P.FusedYoloDecodeNms(plug, B, device=dev)
dbg = torch.zeros(B * 8, dtype=torch.int64, device=dev)
fused.enqueue(B, heads); torch.cuda.synchronize()
lib.trtx_tune_set_ptr(dbg.data_ptr())
acc = torch.zeros(7)
N = 20
for _ in range(N):
    fused.enqueue(B, heads); torch.cuda.synchronize()
    d = dbg.view(B, 8).cpu().double()
    acc[:6] += (d[:, 1:7] - d[:, 0:6]).mean(0).float()
    acc[6] += (d[:, 6] - d[:, 0]).max().float()
lib.trtx_tune_set_ptr(None)
names = ["A collect+stash", "C sort", "D permute", "E1 segment detect", "E2 long segs", "E3 short/medium segs", "max total (cycles)"]
for n, v in zip(names, (acc / N).tolist()):
    print(f"{n:24s} {v:10.0f} cycles  {v / 1965.0:7.2f} us")
