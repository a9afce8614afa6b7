"""The multi-GPU gather (trtx_gather): protocol test on ONE GPU.  Two logical ranks share the device -- each owns a gathered
buffer, a flag array and a control block (plain device memory here; CUDA IPC mappings in production) and the two ranks' steps
are enqueued alternately on two streams, publishing from inside nms_kernel on even steps and with the push kernel on odd ones.
After every step both gathered buffers must hold both ranks' detections (count + kept rows) in the slot of that step, and the
publish counters must advance.  tools/peer_gather_check.py is the same check with real peers (torchrun, CUDA IPC)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tensorrtx_b200 import _lib as L
from tensorrtx_b200 import plugins as P
from tensorrtx_b200 import synth

pytestmark = pytest.mark.gpu


def test_peer_gather_protocol_two_logical_ranks(oracle, dev):
    lib = L.load()
    W, B, K, SLOTS = 2, 3, 1000, 4
    cols = 1 + K * 7
    outs = [torch.full((SLOTS, W * B, cols), -9.0, dtype=torch.float32, device=dev) for _ in range(W)]
    flags = [torch.zeros((W, SLOTS), dtype=torch.int32, device=dev) for _ in range(W)]
    ctrls = [torch.zeros(4, dtype=torch.int32, device=dev) for _ in range(W)]

    def desc(r, slot):
        g = L.Gather()
        g.world, g.rank, g.slots, g.slot = W, r, SLOTS, slot
        for p in range(W):
            g.out_dev[p], g.flags_dev[p] = outs[p].data_ptr(), flags[p].data_ptr()
        g.ctrl_dev = ctrls[r].data_ptr()
        return g
    plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, K, False, False, False, (8, 16, 32))
    fused = [P.FusedYoloDecodeNms(plug, B, 0.5, 0.45, device=dev) for _ in range(W)]
    streams = [torch.cuda.Stream(dev) for _ in range(W)]
    n_steps = 10                                            # every slot is reused: counters 1, 2, 3
    for step in range(n_steps):
        slot = step % SLOTS
        heads = [synth.yolov8_heads(B, seed=900 + 10 * step + r, n_obj=12) for r in range(W)]
        local = []
        for r in range(W):
            g = desc(r, slot)
            with torch.cuda.stream(streams[r]):
                hd = [torch.from_numpy(h).to(dev) for h in heads[r]]
                if step % 2 == 0:   # publish from inside nms_kernel
                    out, _ = fused[r].enqueue(B, hd, gather=g)
                else:               # plain decode + NMS, then the publish kernel
                    out, _ = fused[r].enqueue(B, hd)
                    L.check(lib.trtx_gather_push_enqueue(C.byref(g), out.data_ptr(), B, K, 0, streams[r].cuda_stream), "push")
                L.check(lib.trtx_gather_wait_enqueue(C.byref(g), streams[r].cuda_stream), "wait")
                local.append(out)
        torch.cuda.synchronize()
        for r in range(W):
            assert ctrls[r].cpu().tolist()[1:3] == [0, 0]                 # CTA counter reset, no timeout
            assert flags[r][:, slot].cpu().tolist() == [step // SLOTS + 1] * W
        for r in range(W):                                                # rank r's images, as seen by every rank p
            loc = local[r].cpu().numpy()
            ref, _ = oracle.yolov8_decode(heads[r])
            for p in range(W):
                got = outs[p][slot, r * B:(r + 1) * B].cpu().numpy()
                for b in range(B):
                    n = int(loc[b, 0])
                    res, _ = oracle.nms(0, ref[b], K, 90, 0.5, 0.45)
                    assert n == len(res) and n > 3 and got[b, 0] == n
                    assert np.array_equal(got[b, 1:1 + n * 7], loc[b, 1:1 + n * 7])   # same bytes as the local output
    # a round of two steps published by ONE kernel into slots 2, 3 and awaited by ONE kernel (trtx_gather_push_many / wait_many)
    outs2 = []
    for r in range(W):
        with torch.cuda.stream(streams[r]):
            pair = []
            for k in range(2):
                hd = [torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=990 + 2 * r + k, n_obj=10)]
                f = P.FusedYoloDecodeNms(plug, B, 0.5, 0.45, device=dev)
                pair.append(f.enqueue(B, hd)[0])
            g = desc(r, 2)
            ptrs = L.ptr_array([t.data_ptr() for t in pair])
            L.check(lib.trtx_gather_push_many_enqueue(C.byref(g), ptrs, 2, B, K, 0, streams[r].cuda_stream), "push_many")
            L.check(lib.trtx_gather_wait_many_enqueue(C.byref(g), 2, streams[r].cuda_stream), "wait_many")
            outs2.append(pair)
    torch.cuda.synchronize()
    for r in range(W):
        for k in range(2):
            loc = outs2[r][k].cpu().numpy()
            for p in range(W):
                got = outs[p][2 + k, r * B:(r + 1) * B].cpu().numpy()
                for b in range(B):
                    n = int(loc[b, 0])
                    assert got[b, 0] == n and n > 2 and np.array_equal(got[b, 1:1 + n * 7], loc[b, 1:1 + n * 7])
    # a rank whose peer never publishes gives up instead of hanging the GPU: error flag set
    f2 = torch.zeros((2, SLOTS), dtype=torch.int32, device=dev)
    f2[0, 1] = 5                                            # own counter of slot 1 is ahead of the silent peer's
    c2 = torch.zeros(4, dtype=torch.int32, device=dev)
    lone = L.Gather()
    lone.world, lone.rank, lone.slots, lone.slot = 2, 0, SLOTS, 1
    for p in range(2):
        lone.out_dev[p], lone.flags_dev[p] = outs[0].data_ptr(), f2.data_ptr()
    lone.ctrl_dev = c2.data_ptr()
    L.check(lib.trtx_gather_wait_enqueue(C.byref(lone), None), "wait")
    torch.cuda.synchronize()
    assert c2.cpu().tolist()[2] == 1
