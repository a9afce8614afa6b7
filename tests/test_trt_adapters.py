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

