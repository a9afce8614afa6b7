"""Phase timing of nms_kernel from clock64 stamps (thread 0 of every CTA).
Needs the probe build of the library: python -c "from tensorrtx_b200 import build; build.build(probe=True)" writes
tensorrtx_b200/lib/libtrtx_hot_probe.so (release builds carry no stamps); run with TRTX_LIB=<that path>."""
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
dbg = torch.zeros(B * 16, dtype=torch.int64, device=dev)
fused.enqueue(B, heads); torch.cuda.synchronize()
import ctypes as C
lib.trtx_probe_set_nms_stamps.argtypes = [C.c_void_p]
lib.trtx_probe_set_nms_stamps(dbg.data_ptr())
# stamp ids in program order (nms.cu TRTX_STAMP): label = phase that ends at the stamp
order = [(0, "start"), (1, "A collect+stash"), (2, "C sort"), (8, "D permute"), (3, "E segment table"),
         (4, "E long segments"), (9, "E unit list"), (7, "E IoU bitmaps"), (5, "E resolve"), (6, "F output")]
N = 20
acc = torch.zeros(len(order))
for _ in range(N):
    fused.enqueue(B, heads); torch.cuda.synchronize()
    d = dbg.view(B, 16).cpu().double()
    for k in range(1, len(order)):
        acc[k] += (d[:, order[k][0]] - d[:, order[k - 1][0]]).mean().item()
    acc[0] += (d[:, 6] - d[:, 0]).max().item()
lib.trtx_probe_set_nms_stamps(None)
for k in range(1, len(order)):
    v = acc[k].item() / N
    print(f"{order[k][1]:24s} {v:10.0f} cycles  {v / 1965.0:7.2f} us")
v = acc[0].item() / N
print(f"{'max total':24s} {v:10.0f} cycles  {v / 1965.0:7.2f} us")
