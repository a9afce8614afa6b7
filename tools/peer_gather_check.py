"""Multi-GPU check of the gather fused into nms_kernel (trtx_gather) with REAL peer memory (CUDA IPC between the ranks):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/peer_gather_check.py
Every rank decodes its own seeded batch; after each step every rank's gathered buffer must hold, for every rank r, the
count and the kept rows of r's local output (compared with an NCCL all-gather of the local outputs)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import plugins as P, synth  # noqa: E402
from tensorrtx_b200.pipeline import PeerGather  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
B, K, SLOTS, STEPS = 8, 1000, 4, 11
plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, K, False, False, False, (8, 16, 32))
fused = P.FusedYoloDecodeNms(plug, B, 0.5, 0.45, device=dev)
pg = PeerGather(world, rank, B, 1 + K * 7, dev, slots=SLOTS)
bad = 0
for step in range(STEPS):
    heads = [torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=100 * step + rank, n_obj=20)]
    slot = step % SLOTS
    if step % 2 == 0:                     # alternate the two forms of the publish: from inside nms_kernel / push kernel
        out, _ = fused.enqueue(B, heads, gather=pg.desc(slot))
    else:
        out, _ = fused.enqueue(B, heads)
        pg.push(out, slot, K)
    pg.wait(slot)
    torch.cuda.synchronize()
    ref = torch.empty((world * B, 1 + K * 7), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(ref, out.contiguous())
    got = pg.result(step % SLOTS).cpu().numpy()
    ref = ref.cpu().numpy()
    for i in range(world * B):
        n = int(ref[i, 0])
        if got[i, 0] != n or not np.array_equal(got[i, 1:1 + n * 7], ref[i, 1:1 + n * 7]) or n < 3:
            bad += 1
    assert pg.error() == 0 and pg.published()[:, step % SLOTS].cpu().tolist() == [step // SLOTS + 1] * world, pg.published().cpu().tolist()
    dist.barrier()
t = torch.tensor([bad], device=dev)
dist.all_reduce(t)
if rank == 0:
    print(f"peer gather check: world {world}, {STEPS} steps, mismatching images: {int(t.item())}", flush=True)
pg.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
