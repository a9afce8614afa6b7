"""Three launches of the default scan kernel per data set (background / 8 objects / 64 objects per image), for
`ncu -k regex:yolo_v8_scan --set full`: which stall reason pays for the candidates?"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L, plugins as P, synth
dev = torch.device("cuda", 0)
lib = L.load()
B = 32
for name, nobj in (("background", 0), ("typical", 8), ("dense", 64)):
    heads = [torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=1, n_obj=nobj)]
    plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32))
    fused = P.FusedYoloDecodeNms(plug, B, device=dev)
    for _ in range(3):
        fused.enqueue_scan(B, heads)
    torch.cuda.synchronize()
    print(name, "done", flush=True)
