"""The C-ABI library loads without a GPU and exports every symbol that
include/fpv_b200.h declares; argument validation that needs no device works."""
import ctypes as C
import os
import re

import fusion_power_video_b200 as fpv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fpv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = fpv.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fpv_b200.h but not exported"


def test_version_and_error_paths_without_device():
    assert "sm_100a" in fpv.version()
    L = fpv.lib()
    h = C.c_void_p()
    # invalid geometry is rejected before any CUDA call
    assert L.fpv_create(C.byref(h), 0, 0, 16, 0, 0, 1) == 1
    assert b"dimensions" in L.fpv_last_error(None)
    # a shift beyond 16 bits is refused outright (big-endian shift 9..16 gives a decode-only context, see the GPU tests)
    assert L.fpv_create(C.byref(h), 0, 16, 16, 17, 1, 1) == 3
    assert L.fpv_create(C.byref(h), 0, 16, 16, 3, 0, 70000) == 1   # max_batch is a grid dimension: <= 65535
    # NULL context is handled
    assert L.fpv_wait(None, 0) == 1
    assert L.fpv_plane_bytes(None) == 0


def test_no_cpu_fallback_without_device():
    """On a box without a GPU, creating a context must FAIL (no silent CPU path)."""
    if fpv.device_count() > 0:
        return
    try:
        fpv.Context(64, 64)
    except fpv.FpvError as e:
        assert e.code == 5
    else:
        raise AssertionError("Context creation succeeded without a CUDA device")


def test_product_does_not_reference_oracle():
    """Nothing under fusion_power_video_b200/ may include/link/call oracle/."""
    bad = []
    pkg = os.path.join(ROOT, "fusion_power_video_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
                t = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"fpv_oracle|libfpv_ref|oracle_binding|fpvo_|/oracle/", t):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_host_library_loads_and_fails_loudly_without_device():
    """libfusion_power_video_b200.so (fpvc:: classes + fpvh_* C entry points) loads on a CPU-only box;
    without a CUDA device encoding must fail with a message, not fall back to a CPU path."""
    import numpy as np

    from fusion_power_video_b200 import host

    L = host.lib()
    for n in ("fpvh_encode_stream", "fpvh_time_encode", "fpvh_decode_stream", "fpvh_random_access", "fpvh_unextract",
              "fpvh_last_error"):
        assert hasattr(L, n), n
    img = (np.arange(64 * 32, dtype=np.uint32) * 37 % 65536).astype(np.uint16)
    out = host.unextract(img, 64, 32, 4, True)
    assert out[0] == ((int(img[0]) >> 4) >> 8) and out[1] == ((int(img[0]) >> 4) & 0xff)
    if fpv.device_count() > 0:
        return
    try:
        host.encode_stream(img.reshape(1, -1), 64, 32, threads=0)
    except host.HostError as e:
        assert "fpv_create" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("encode_stream succeeded without a CUDA device")
