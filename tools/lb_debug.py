"""Debug aid: run the letterbox on single images / mixed batches and compare with the oracle."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import plugins as P, synth
from oracle import oracle as O
dev = torch.device("cuda", 0)
for sizes in ([(640, 640)], [(480, 640)], [(640, 480)], [(639, 640)], [(640, 624)], [(640, 640), (480, 640), (640, 500), (1080, 1920), (640, 640)]):
    frs = [synth.frames(1, seed=7 + i, h=h, w=w)[0] for i, (h, w) in enumerate(sizes)]
    dst = torch.zeros((len(sizes), 3, 640, 640), dtype=torch.float32, device=dev)
    try:
        P.cuda_batch_preprocess([torch.from_numpy(f).to(dev) for f in frs], dst, 640, 640)
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        ok = [bool(np.array_equal(got[b], O.warpaffine(f, 640, 640))) or float(np.abs(got[b] - O.warpaffine(f, 640, 640)).max()) for b, f in enumerate(frs)]
        print(sizes, ok, flush=True)
    except Exception as e:
        print(sizes, "FAILED", str(e)[:200], flush=True)
        break
