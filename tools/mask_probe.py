"""process_mask kernel timing: B images x M detections, 640x640 masks from 32x160x160 prototypes."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import plugins as P
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
B, K, M, R = 8, 100, 50, 39
proto = torch.from_numpy(rng.standard_normal((B, 32, 160, 160)).astype(np.float32)).to(dev)
dets = np.zeros((B, 1 + K * R), np.float32)
for b in range(B):
    dets[b, 0] = M
    rows = dets[b, 1:].reshape(K, R)
    for i in range(M):
        s = float(np.exp(rng.uniform(np.log(32), np.log(256))))
        rows[i, :4] = [rng.uniform(0, 640 - s), rng.uniform(0, 640 - s), s, s]
        rows[i, 7:39] = rng.standard_normal(32) * 0.5
dets = torch.from_numpy(dets).to(dev)
out = torch.empty((B, M, 640, 640), dtype=torch.float32, device=dev)
for _ in range(3):
    P.process_mask(proto, dets, K, R, 7, M, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 20
e0.record()
for _ in range(N):
    P.process_mask(proto, dets, K, R, 7, M, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / N * 1e3
nb = out.numel() * 4
print(json.dumps({"kernel": "process_mask", "masks": B * M, "us": round(us, 1), "written_GBps": round(nb / us / 1e3, 1),
                  "us_per_mask": round(us / (B * M), 3)}))
