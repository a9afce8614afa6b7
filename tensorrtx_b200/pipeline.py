"""Detection hot path as one object: pinned frames -> H2D -> batched letterbox pre-process ->
[backbone: TensorRT in the reference, injected here] -> fused YoloLayer decode + NMS -> D2H of the
compact detections.  This is the call a user of the reference's `yolov8_det -d` loop makes per batch
(yolov8/yolov8_det.cpp:223-244), minus disk IO and drawing.

Everything is asynchronous on one CUDA stream; the only host synchronisation is the caller's.
Multi-GPU: one pipeline per process/GPU (batch-sharded, images are independent -- SURVEY 8e);
`gather()` is the single NCCL collective (fixed-size all-gather of the compact detections).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import _lib as L
from .plugins import FusedYoloDecodeNms, PreprocessPlan, YoloLayerPlugin


class DetectionPipeline:
    def __init__(self, batch: int, src_h: int = 640, src_w: int = 640, net_h: int = 640, net_w: int = 640,
                 num_classes: int = 80, strides: Sequence[int] = (8, 16, 32), max_out: int = 1000,
                 conf_thresh: float = 0.5, nms_thresh: float = 0.45, device: str | torch.device = "cuda",
                 input_dtype: torch.dtype = torch.float32, head_dtype: int = L.F32,
                 backbone: Callable[[torch.Tensor], Sequence[torch.Tensor]] | None = None):
        self.batch, self.src_h, self.src_w, self.net_h, self.net_w = batch, src_h, src_w, net_h, net_w
        self.device = torch.device(device)
        self.backbone = backbone
        self.plugin = YoloLayerPlugin(num_classes, 17, 0.0, net_w, net_h, max_out, False, False, False, strides,
                                      in_dtype=head_dtype)
        self.fused = FusedYoloDecodeNms(self.plugin, batch, conf_thresh, nms_thresh, max_det=max_out,
                                        device=self.device, return_index=False)
        self.frames_dev = torch.empty((batch, src_h, src_w, 3), dtype=torch.uint8, device=self.device)
        self.net_input = torch.empty((batch, 3, net_h, net_w), dtype=input_dtype, device=self.device)
        self.out_host = torch.empty((batch, 1 + max_out * 7), dtype=torch.float32).pin_memory()
        self.h2d_bytes = self.frames_dev.numel()
        self.d2h_bytes = self.out_host.numel() * 4
        # descriptors of the (persistent) device frame buffers are built once, not per step
        self.pre = PreprocessPlan(list(self.frames_dev.unbind(0)), self.net_input, net_w, net_h)
        # decode + NMS of the heads run on a high-priority side stream next to the letterbox (see run_device)
        self._side = torch.cuda.Stream(self.device, priority=-1) if self.device.type == "cuda" else None

    def run(self, frames_host: torch.Tensor, heads: Sequence[torch.Tensor] | None = None, stream=None) -> torch.Tensor:
        """frames_host: pinned uint8 [B, H, W, 3] BGR.  heads: the backbone's per-stride outputs when no
        backbone callable was given.  Returns the pinned host buffer [B, 1+K*7] (valid after a stream sync)."""
        self.frames_dev.copy_(frames_host, non_blocking=True)
        self.pre.enqueue(stream)
        if self.backbone is not None:
            heads = self.backbone(self.net_input)
        out, _ = self.fused.enqueue(self.batch, heads, stream)
        self.out_host.copy_(out, non_blocking=True)
        return self.out_host

    def run_device(self, heads: Sequence[torch.Tensor], stream=None, overlap: bool = True) -> torch.Tensor:
        """Device-resident step: pre-process the frames already in HBM + decode + NMS (no host copies).

        The letterbox of this batch does not depend on the decode/NMS of the head tensors (with a real backbone the heads
        of batch i are decoded while the frames of batch i+1 are pre-processed), so the two halves are issued as
        parallel branches: scan -> NMS on a high-priority side stream forked from the current stream, the letterbox on
        the current stream, joined at the end.  Captured into a CUDA graph this becomes two concurrent branches; the
        latency-bound NMS (one CTA per image, 32 of 148 SMs) then runs underneath the HBM-bound letterbox instead of
        after it."""
        if not overlap or self._side is None or stream is not None:
            self.pre.enqueue(stream)
            out, _ = self.fused.enqueue(self.batch, heads, stream)
            return out
        cur = torch.cuda.current_stream(self.device)
        self._side.wait_stream(cur)                  # fork
        with torch.cuda.stream(self._side):
            out, _ = self.fused.enqueue(self.batch, heads)
        self.pre.enqueue()
        cur.wait_stream(self._side)                  # join
        return out

    # ---- overlapped host pipeline: H2D of batch i+1 runs on a copy stream while batch i is decoded ----
    def _make_slots(self, n: int) -> None:
        self._slots = []
        for _ in range(n):
            fd = torch.empty_like(self.frames_dev)
            self._slots.append({
                "frames": fd,
                "pre": PreprocessPlan(list(fd.unbind(0)), self.net_input, self.net_w, self.net_h),
                "out_host": torch.empty_like(self.out_host).pin_memory(),
                "h2d": torch.cuda.Event(), "done": torch.cuda.Event(),
            })
            self._slots[-1]["done"].record(torch.cuda.current_stream(self.device))
        self._copy_stream = torch.cuda.Stream(self.device)
        self._next = 0

    def submit(self, frames_host: torch.Tensor, heads: Sequence[torch.Tensor] | None = None, slots: int = 2):
        """Asynchronous step with double buffering: returns (pinned host result, event); the result is valid
        once the event has completed.  The H2D copy is issued on a dedicated copy stream so that it overlaps
        the kernels of the previous submit() (the reference serialises memcpy -> H2D -> kernel -> sync per
        image, yolov8/src/preprocess.cu:89-127)."""
        if not hasattr(self, "_slots"):
            self._make_slots(slots)
        sl = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        compute = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_event(sl["done"])           # slot buffers are free again
        with torch.cuda.stream(self._copy_stream):
            sl["frames"].copy_(frames_host, non_blocking=True)
            sl["h2d"].record(self._copy_stream)
        compute.wait_event(sl["h2d"])
        sl["pre"].enqueue()
        if self.backbone is not None:
            heads = self.backbone(self.net_input)
        out, _ = self.fused.enqueue(self.batch, heads)
        sl["out_host"].copy_(out, non_blocking=True)
        sl["done"].record(compute)
        return sl["out_host"], sl["done"]

    # ---- CUDA graphs: the step is launch-bound (4 kernels of 10-30 us), so replaying a captured graph
    # removes the per-launch host cost (guide: "capture launch-bound inner loops in CUDA graphs") ----
    def capture(self, fn: Callable[[], object], capture_error_mode: str = "global") -> "torch.cuda.CUDAGraph":
        """Capture `fn` (which must only enqueue work on the current stream, or on streams forked from and joined
        back into it) into a CUDA graph.  Use capture_error_mode="thread_local" when other threads of the process
        (e.g. the NCCL watchdog) may touch CUDA during the capture."""
        fn()  # warm-up outside capture (lazy module loads, cudaFuncSetAttribute)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=capture_error_mode):
            fn()
        return g


def gather(local: torch.Tensor, world_size: int) -> torch.Tensor:
    """The one collective of the path: fixed-size all-gather of [B_local, 1+K*7]."""
    import torch.distributed as dist

    if world_size == 1:
        return local
    out = torch.empty((world_size * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


class GatherRing:
    """The same all-gather, pipelined: step i's compact detections travel on a side stream while step i+1 computes.

    `slots` preallocated destination buffers ([world*B_local, 1+K*7]); `launch(local, slot)` orders the collective
    after everything queued so far on the current stream and returns immediately; `reuse(slot)` makes the current stream
    wait until that slot's previous collective has read its source (call it before overwriting the source buffer);
    `join()` makes the current stream wait for all collectives (call before the closing timing event / before reading).
    On a CPU process group (gloo, tests) there are no streams and every call is synchronous."""

    def __init__(self, world_size: int, rows: int, cols: int, device, slots: int = 4, dtype=torch.float32):
        self.world = world_size
        self.cuda = torch.device(device).type == "cuda"
        self.bufs = [torch.empty((world_size * rows, cols), dtype=dtype, device=device) for _ in range(slots)]
        if self.cuda:
            self.comm = torch.cuda.Stream(device=device)
            self.done = [torch.cuda.Event() for _ in range(slots)]
            self.ready = [torch.cuda.Event() for _ in range(slots)]
            self.used = [False] * slots

    def launch(self, local: torch.Tensor, slot: int) -> torch.Tensor:
        import torch.distributed as dist

        out = self.bufs[slot]
        if self.world == 1:
            return local
        if not self.cuda:
            dist.all_gather_into_tensor(out, local.contiguous())
            return out
        cur = torch.cuda.current_stream(local.device)
        self.ready[slot].record(cur)
        self.comm.wait_event(self.ready[slot])
        with torch.cuda.stream(self.comm):
            dist.all_gather_into_tensor(out, local)
            self.done[slot].record(self.comm)
        self.used[slot] = True
        return out

    def reuse(self, slot: int) -> None:
        if self.cuda and self.world > 1 and self.used[slot]:
            torch.cuda.current_stream(self.bufs[slot].device).wait_event(self.done[slot])

    def join(self) -> None:
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.bufs[0].device).wait_stream(self.comm)
