"""ncu target: a few launches of the RoIAlign kernels at the Faster R-CNN C4 shapes (same inputs as tools/roi_probe.py).
usage: ncu --set full --import-source on -k regex:roi_align -s 2 -c 1 -o gpurun_out/roi python tools/roi_ncu_target.py [direct]"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L
from tensorrtx_b200 import plugins as P
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
B, N, C, H, W, Pp = 1, 1000, 1024, 50, 67, 14
feat = torch.from_numpy(rng.standard_normal((B, C, H, W)).astype(np.float32)).to(dev)
x1 = rng.uniform(0, 900, (B, N)); y1 = rng.uniform(0, 700, (B, N))
w = np.exp(rng.uniform(np.log(16), np.log(600), (B, N))); h = np.exp(rng.uniform(np.log(16), np.log(500), (B, N)))
rois = torch.from_numpy(np.stack([x1, y1, x1 + w, y1 + h], -1).astype(np.float32)).to(dev)
out = torch.empty((B, N, C, Pp, Pp), dtype=torch.float32, device=dev)
plug = P.RoiAlignPlugin(Pp, 1 / 16, 0, N, C, H, W)
mode = L.ROI_DIRECT if "direct" in sys.argv else L.ROI_WINDOW
for _ in range(4):
    assert plug.enqueue(B, [rois, feat], [out], mode=mode) == 0
torch.cuda.synchronize()
