"""include/trtx_plugins.h (header-only TensorRT adapters) compiled against the mock NvInfer.h and driven the way an
engine builder / the TensorRT runtime drives plugins.  Host-only part runs without a GPU; --gpu enqueues on the device."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _build(tmp_path):
    from tensorrtx_b200 import _lib as L

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda = Path(nvcc).resolve().parents[1]
    exe = tmp_path / "trt_adapter_check"
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", str(ROOT / "include"), "-I", str(ROOT / "tests" / "mock_trt"),
           "-I", str(cuda / "include"), str(ROOT / "tests" / "trt_adapter_check.cpp"), "-o", str(exe),
           str(L.LIB_PATH), f"-Wl,-rpath,{L.LIB_PATH.parent}", "-L", str(cuda / "lib64"), "-lcudart", f"-Wl,-rpath,{cuda / 'lib64'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_trt_adapters_host_surface(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "trt adapter check ok" in r.stdout


@pytest.mark.gpu
def test_trt_adapters_enqueue_on_gpu(tmp_path, dev):   # `dev` skips on a box without CUDA
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu enqueue ok" in r.stdout


def _build_preprocess_compat(tmp_path):
    from tensorrtx_b200 import _lib as L

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuda = Path(nvcc).resolve().parents[1]
    exe = tmp_path / "preprocess_compat_check"
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-I", str(ROOT / "include"), "-I", str(ROOT / "oracle" / "shim"),
           "-I", str(cuda / "include"), str(ROOT / "tests" / "preprocess_compat_check.cpp"), "-o", str(exe),
           str(L.LIB_PATH), f"-Wl,-rpath,{L.LIB_PATH.parent}", "-L", str(cuda / "lib64"), "-lcudart", f"-Wl,-rpath,{cuda / 'lib64'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_preprocess_compat_header_compiles_and_links(tmp_path):
    """include/trtx_preprocess_compat.h = the reference's preprocess.h API (cuda_preprocess_init / _destroy / cuda_preprocess /
    cuda_batch_preprocess) on top of the C ABI; compiled here against the OpenCV type shim (no OpenCV in this image)."""
    exe = _build_preprocess_compat(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "preprocess compat check ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_preprocess_compat_header_on_gpu(tmp_path, dev):
    """The same header on the device: cuda_batch_preprocess (pinned ring, H2D, one launch) and cuda_preprocess give the
    bytes of a direct trtx_preprocess_batch_enqueue call."""
    exe = _build_preprocess_compat(tmp_path)
    r = subprocess.run([str(exe), "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "gpu preprocess compat ok" in r.stdout, r.stdout + r.stderr



def test_cpp_adapter_blobs_equal_python_mirror_and_reference_layouts(tmp_path):
    """The serialization layouts exist twice in this repository (include/trtx_plugins.h for TensorRT, tensorrtx_b200/plugins.py
    for the tests and the bench): the C++ adapters' serialized bytes must equal the Python mirror's for the same constructor
    arguments, and -- for the rcnn plugins, which the mirror does not serialize -- the reference's own write() sequences
    (rcnn/RpnDecodePlugin.h:64-75, RpnNmsPlugin.h:48-52, PredictorDecodePlugin.h:62-70, BatchedNmsPlugin.h:51-56,
    RoiAlignPlugin.h:53-61, MaskRcnnInferencePlugin.h:41-45) restated with struct.pack."""
    import struct

    from tensorrtx_b200 import plugins as P

    exe = _build(tmp_path)
    r = subprocess.run([str(exe), "--dump"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    blobs = {ln.split()[1]: bytes.fromhex(ln.split()[2]) for ln in r.stdout.splitlines() if ln.startswith("blob ")}
    # ---- C++ adapter == Python mirror ----
    v8 = P.YoloLayerPlugin(80, 17, 0.0, 640, 640, 1000, False, False, False, (8, 16, 32))
    assert blobs["yolov8"] == v8.serialize()
    kern = [P.YoloKernel(80, 80, [10, 13, 16, 30, 33, 23]), P.YoloKernel(40, 40, [30, 61, 62, 45, 59, 119]),
            P.YoloKernel(20, 20, [116, 90, 156, 198, 373, 326])]
    v5 = P.YoloLayerPluginV5(80, 640, 640, 1000, False, kern)
    assert blobs["yolov5"] == v5.serialize()
    assert blobs["decode"] == P.DecodePlugin().serialize()
    # ---- C++ adapter == the reference's write() order and types (size_t = 8 bytes, unsigned / int / float = 4) ----
    assert blobs["rpn_decode"] == struct.pack("<iQ60ffQQQQ", 6000, 60, *([1.0] * 60), 16.0, 50, 67, 800, 1067)
    assert blobs["rpn_nms"] == struct.pack("<fiQ", 0.7, 1000, 6000)
    assert blobs["predictor_decode"] == struct.pack("<IIIIQ4f", 1000, 80, 800, 1067, 4, 10.0, 10.0, 5.0, 5.0)
    assert blobs["batched_nms"] == struct.pack("<ifiQ", 1, 0.5, 100, 1000)
    assert blobs["roi_align"] == struct.pack("<ifiiiii", 14, 1.0 / 16, 0, 1000, 1024, 50, 67)
    assert blobs["mask_rcnn_inference"] == struct.pack("<iii", 100, 14, 1)
