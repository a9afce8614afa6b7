"""Where do the ~2.7 us that candidates add to the scan kernel go?  Experiment builds of the library (never the product):
  probe1: candidates are counted but not decoded / stored (no epilogue tail)
  probe2: no per-anchor replay in the class loop (running maxima only)
Graph timing of the scan at b32 on the dense heads, next to the release library.  Usage (GPU box): python tools/scan_probe.py"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import torch
    from tensorrtx_b200 import _lib as L, plugins as P, synth
    dev = torch.device("cuda", 0)
    B, R = 32, 4
    for data, nobj in (("dense", 64), ("background", 0)):
        sets = [[torch.from_numpy(h).to(dev) for h in synth.yolov8_heads(B, seed=i, n_obj=nobj)] for i in range(R)]
        plug = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32))
        fused = P.FusedYoloDecodeNms(plug, B, device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fused.enqueue_scan(B, sets[0]); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(20): fused.enqueue_scan(B, sets[i % R])
            for _ in range(3): g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): g.replay()
            e1.record(); torch.cuda.synchronize()
        print(json.dumps({"lib": os.environ.get("TRTX_LIB", "release"), "data": data, "us": round(e0.elapsed_time(e1) / 200 * 1e3, 2)}), flush=True)
    sys.exit(0)
from tensorrtx_b200 import build as B
libs = [None, B.build(defines=("TRTX_SCAN_PROBE=1",), suffix="1"), B.build(defines=("TRTX_SCAN_PROBE=2",), suffix="2")]
for lib in libs:
    env = dict(os.environ)
    if lib: env["TRTX_LIB"] = str(lib)
    else: env.pop("TRTX_LIB", None)
    subprocess.run([sys.executable, __file__, "--child"], env=env)
