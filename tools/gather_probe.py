"""What does the gather cost on ONE GPU (a single logical rank gathering to itself: same stores, fences, flags and wait, no
NVLink)?  The step of bench.py (4 steps per graph, letterbox chain || scan -> NMS chain) without gather, with the gather fused
into nms_kernel, and with the push kernel."""
import ctypes as C, json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tensorrtx_b200 import _lib as L, synth
from tensorrtx_b200.pipeline import DetectionPipeline
dev = torch.device("cuda", 0)
lib = L.load()
B, R, K = 32, 4, 1000
heads = [[torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=i)] for i in range(R)]
pipes = []
for i in range(R):
    p = DetectionPipeline(B, device=dev)
    p.frames_dev.copy_(torch.from_numpy(synth.frames(B, seed=i)).to(dev))
    pipes.append(p)
cols = 1 + K * 7
out = torch.zeros((4, B, cols), device=dev)
flags = torch.zeros((1, 4), dtype=torch.int32, device=dev)
ctrl = torch.zeros(4, dtype=torch.int32, device=dev)
gs = []
for sl in range(4):
    g = L.Gather()
    g.world, g.rank, g.slots, g.slot = 1, 0, 4, sl
    g.out_dev[0], g.flags_dev[0], g.ctrl_dev = out.data_ptr(), flags.data_ptr(), ctrl.data_ptr()
    gs.append(g)
chain_b = torch.cuda.Stream(dev, priority=-1)
chain_c = torch.cuda.Stream(dev)


def group(mode):
    def f():
        cur = torch.cuda.current_stream(dev)
        chain_b.wait_stream(cur)
        if mode == "push_c":
            chain_c.wait_stream(cur)
        with torch.cuda.stream(chain_b):
            for j in range(R):
                g = gs[j]
                if mode == "fused":
                    pipes[j].fused.enqueue(B, heads[j], gather=g)
                    L.check(lib.trtx_gather_wait_enqueue(C.byref(g), chain_b.cuda_stream), "wait")
                else:
                    o, _ = pipes[j].fused.enqueue(B, heads[j])
                    if mode == "push":
                        L.check(lib.trtx_gather_push_enqueue(C.byref(g), o.data_ptr(), B, K, 0, chain_b.cuda_stream), "push")
                        L.check(lib.trtx_gather_wait_enqueue(C.byref(g), chain_b.cuda_stream), "wait")
                    elif mode == "push_c":      # the push + wait kernel on a third chain: off the scan -> NMS critical path
                        ev = torch.cuda.Event()
                        ev.record(chain_b)
                        chain_c.wait_event(ev)
                        L.check(lib.trtx_gather_push_enqueue(C.byref(g), o.data_ptr(), B, K, 0, chain_c.cuda_stream), "push")
                        L.check(lib.trtx_gather_wait_enqueue(C.byref(g), chain_c.cuda_stream), "wait")
        for j in range(R):
            pipes[j].pre.enqueue()
        cur.wait_stream(chain_b)
        if mode == "push_c":
            cur.wait_stream(chain_c)
    return f


st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    for mode in ("none", "fused", "push", "push_c", "none"):
        gr = pipes[0].capture(group(mode))
        for _ in range(5):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"gather": mode, "us_per_step": round(e0.elapsed_time(e1) / 400 * 1e3, 2), "ctrl": ctrl.cpu().tolist()}), flush=True)
