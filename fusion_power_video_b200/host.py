"""ctypes binding of the host layer's C entry points (csrc/host/capi.cc):
the fpvc::Encoder / StreamingDecoder / RandomAccessDecoder classes of
``lib/libfusion_power_video_b200.so`` driven end to end (GPU transform + host
brotli + stream framing).  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libfusion_power_video_b200.so")
_lib = None


def lib_path() -> str:
    return _LIB_PATH


def _preload_cabi():
    from . import binding

    binding.lib()  # the C ABI first, so the dependency resolves from lib/


def lib() -> C.CDLL:
    """Loads the host library (which links libfpv_b200.so).  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} not found: run `make -C fusion_power_video_b200/csrc all`")
    _preload_cabi()
    L = C.CDLL(_LIB_PATH)
    vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32
    L.fpvh_last_error.restype = C.c_char_p
    L.fpvh_encode_stream.argtypes = [sz, sz, i32, i32, sz, u32, i32, vp, vp, sz, vp, sz]
    L.fpvh_encode_stream.restype = sz
    L.fpvh_columnar_roundtrip.argtypes = [sz, sz, i32, i32, i32, i32, i32, vp, vp, sz, vp, vp, sz, C.POINTER(sz),
                                          C.POINTER(C.c_long), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(sz)]
    L.fpvh_columnar_roundtrip.restype = C.c_long
    L.fpvh_columnar_planes.argtypes = [sz, sz, i32, i32, i32, vp, vp, sz, vp, vp, vp, vp]
    L.fpvh_columnar_planes.restype = C.c_long
    L.fpvh_scan_coded_plane.argtypes = [vp, sz, sz, vp, sz, C.POINTER(sz), C.POINTER(sz)]
    L.fpvh_scan_coded_plane.restype = i32
    L.fpvh_encode_stream_multi.argtypes = [sz, sz, i32, i32, sz, u32, vp, i32, i32, vp, vp, sz, vp, sz, C.POINTER(C.c_double)]
    L.fpvh_encode_stream_multi.restype = sz
    L.fpvh_ingest.argtypes = [sz, sz, i32, i32, sz, u32, i32, i32, vp, sz, C.c_double, C.c_double, sz, C.POINTER(C.c_double)]
    L.fpvh_ingest.restype = i32
    L.fpvh_time_encode.argtypes = [sz, sz, i32, i32, sz, u32, i32, vp, vp, sz, C.POINTER(sz)]
    L.fpvh_time_encode.restype = C.c_double
    L.fpvh_decode_stream.argtypes = [vp, sz, sz, u32, i32, i32, i32, vp, sz, C.POINTER(sz), C.POINTER(sz),
                                     C.POINTER(C.c_double)]
    L.fpvh_decode_stream.restype = C.c_long
    L.fpvh_random_access.argtypes = [vp, sz, u32, i32, sz, sz, vp, vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
    L.fpvh_random_access.restype = i32
    L.fpvh_unextract.argtypes = [vp, sz, sz, i32, i32, vp]
    L.fpvh_unextract.restype = None
    _lib = L
    return L


class HostError(RuntimeError):
    pass


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def last_error() -> str:
    return lib().fpvh_last_error().decode()


class _entropy_mode:
    """Selects the Encoder's entropy stage for the calls inside the block (GpuOptions::gpu_entropy = -1 reads the
    environment variable FPV_GPU_ENTROPY in Encoder::Init)."""

    def __init__(self, gpu_entropy):
        self.value = "1" if gpu_entropy else "0"

    def __enter__(self):
        self.old = os.environ.get("FPV_GPU_ENTROPY")
        os.environ["FPV_GPU_ENTROPY"] = self.value

    def __exit__(self, *a):
        if self.old is None:
            del os.environ["FPV_GPU_ENTROPY"]
        else:
            os.environ["FPV_GPU_ENTROPY"] = self.old


def encode_stream(frames, xsize, ysize, shift=0, big_endian=False, threads=4, batch=8, delta=None, device=0,
                  gpu_entropy=False):
    """frames: uint16 [n, ysize*xsize]; delta defaults to frames[0].  Returns the stream as bytes.
    gpu_entropy: planes are entropy-coded on the GPU (valid brotli, not libbrotli's bytes) instead of host brotli."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, xsize * ysize)
    delta = frames[0] if delta is None else np.ascontiguousarray(delta, dtype=np.uint16).reshape(-1)
    n = frames.shape[0]
    cap = 64 + (n + 1) * (xsize * ysize * 5 // 2 + 4096)
    out = np.empty(cap, np.uint8)
    with _entropy_mode(gpu_entropy):
        size = L.fpvh_encode_stream(xsize, ysize, shift, int(big_endian), threads, batch, device, _p(delta), _p(frames),
                                    n, _p(out), cap)
    if size == 0:
        raise HostError(f"encode failed: {last_error()}")
    if size > cap:
        raise HostError("stream larger than the output buffer")
    return out[:size].tobytes()


def encode_stream_multi(frames, xsize, ysize, shift=0, big_endian=False, threads=4, batch=8, delta=None, devices=(0,),
                        gpu_entropy=False, return_time=False):
    """One fpvc::Encoder over several GPUs (GpuOptions::devices): batches round-robin over `devices`, the delta frame
    reaches devices[1:] by peer copy, the stream is the single-GPU stream byte for byte."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, xsize * ysize)
    delta = frames[0] if delta is None else np.ascontiguousarray(delta, dtype=np.uint16).reshape(-1)
    n = frames.shape[0]
    cap = 64 + (n + 1) * (xsize * ysize * 5 // 2 + 4096)
    out = np.empty(cap, np.uint8)
    devs = np.ascontiguousarray(list(devices), dtype=np.int32)
    sec = C.c_double(0)
    size = L.fpvh_encode_stream_multi(xsize, ysize, shift, int(big_endian), threads, batch, _p(devs), devs.size,
                                      int(bool(gpu_entropy)), _p(delta), _p(frames), n, _p(out), cap, C.byref(sec))
    if size == 0:
        raise HostError(f"encode failed: {last_error()}")
    if size > cap:
        raise HostError("stream larger than the output buffer")
    stream = out[:size].tobytes()
    return (stream, sec.value) if return_time else stream


def ingest(frames, xsize, ysize, fps, seconds, shift=0, big_endian=False, threads=4, batch=8, device=0, gpu_entropy=False,
           ring_frames=64):
    """Paced real-time ingest into fpvc::Encoder (capi.cc fpvh_ingest): frames arrive at `fps` for `seconds`; returns a
    dict with offered / encoded / dropped counts and the arrival-to-callback latency distribution (ms)."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, xsize * ysize)
    out = (C.c_double * 10)()
    rc = L.fpvh_ingest(xsize, ysize, shift, int(big_endian), threads, batch, device, int(bool(gpu_entropy)), _p(frames),
                       frames.shape[0], float(fps), float(seconds), ring_frames, out)
    if rc != 0:
        raise HostError(f"ingest failed: {last_error()}")
    keys = ["offered", "encoded", "dropped", "p50_ms", "p99_ms", "max_ms", "mean_ms", "achieved_fps", "stream_bytes", "wall_s"]
    d = dict(zip(keys, [float(v) for v in out]))
    for k in ("offered", "encoded", "dropped", "stream_bytes"):
        d[k] = int(d[k])
    d.update(offered_fps=float(fps), batch=batch, threads=threads, gpu_entropy=bool(gpu_entropy), ring_frames=ring_frames)
    return d


def time_encode(frames, xsize, ysize, shift=0, big_endian=False, threads=4, batch=8, delta=None, device=0,
                gpu_entropy=False):
    """(seconds, stream bytes) of Encoder Init + CompressFrame x n + Finish (benchmark.cc's timing window)."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, xsize * ysize)
    delta = frames[0] if delta is None else np.ascontiguousarray(delta, dtype=np.uint16).reshape(-1)
    size = C.c_size_t(0)
    with _entropy_mode(gpu_entropy):
        t = L.fpvh_time_encode(xsize, ysize, shift, int(big_endian), threads, batch, device, _p(delta), _p(frames),
                               frames.shape[0], C.byref(size))
    if t < 0:
        raise HostError(f"encode failed: {last_error()}")
    return t, size.value


def decode_stream(stream, max_frames, xsize, ysize, block=0, batch=32, raw_shift=-1, big_endian=False, device=0,
                  return_time=False, keep=True):
    """Decodes with StreamingDecoder in `block`-byte pieces.  Returns uint16 [n, ysize*xsize] (images, or the raw
    file words when raw_shift >= 0).  keep=False: the callback only counts the frames (timing the decoder without a
    consumer's copy of every frame); the returned array then has the right length and no contents."""
    L = lib()
    buf = np.frombuffer(stream, np.uint8)
    out = np.empty((max_frames if keep else 0, xsize * ysize), np.uint16)
    out[...] = 0            # touched before the timed call: the callback's memcpy must not pay the page faults
    W, H, sec = C.c_size_t(0), C.c_size_t(0), (C.c_double * 2)()
    n = L.fpvh_decode_stream(_p(buf), buf.size, block, batch, device, raw_shift, int(big_endian), _p(out) if keep else None,
                             max_frames, C.byref(W), C.byref(H), sec)
    if n < 0:
        raise HostError(f"decode failed: {last_error()}")
    if n and (W.value, H.value) != (xsize, ysize):
        raise HostError(f"stream is {W.value}x{H.value}, expected {xsize}x{ysize}")
    if not keep:
        out = np.empty((n, 0), np.uint16)
    if return_time == "both":
        return out[:n], sec[0], sec[1]      # whole call, time until the first frame came out
    return (out[:n], sec[0]) if return_time else out[:n]


def random_access(stream, first, count, xsize, ysize, batch=32, device=0, want_preview=True):
    """(numframes, frames uint16 [count, P], preview uint8 [(ysize//4)*(xsize//4)] of frame `first`)."""
    L = lib()
    buf = np.frombuffer(stream, np.uint8)
    frames = np.zeros((count, xsize * ysize), np.uint16)
    preview = np.zeros((ysize // 4) * (xsize // 4), np.uint8) if want_preview else None
    nf, W, H = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    ok = L.fpvh_random_access(_p(buf), buf.size, batch, device, first, count, _p(frames), _p(preview), C.byref(nf),
                              C.byref(W), C.byref(H))
    if not ok:
        raise HostError(f"random access decode failed: {last_error()}")
    return nf.value, frames, preview


def unextract(img, xsize, ysize, shift, big_endian):
    img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
    out = np.empty(xsize * ysize * 2, np.uint8)
    lib().fpvh_unextract(_p(img), xsize, ysize, shift, int(big_endian), _p(out))
    return out


IMAGE_PREVIEW, IMAGE_MSB8, IMAGE_FULL = 0, 1, 2


def columnar_roundtrip(frames, timestamps, xsize, ysize, shift=0, big_endian=False, frames_per_batch=10, image_type=IMAGE_FULL,
                       unshift=False):
    """frames -> ColumnarBatchEncoder -> Batches -> ColumnarBatchDecoder -> images (the reference's
    columnar_batch_decoder_test.cc wiring).  Returns (images, timestamps, info)."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, xsize * ysize)
    ts = np.ascontiguousarray(timestamps, dtype=np.int64)
    n = frames.shape[0]
    per = {IMAGE_PREVIEW: (xsize // 4) * (ysize // 4), IMAGE_MSB8: xsize * ysize, IMAGE_FULL: 2 * xsize * ysize}[image_type]
    out = np.zeros(max(n * per, 1), np.uint8)
    out_ts = np.zeros(max(n, 1), np.int64)
    bpi, batches, comp = C.c_size_t(0), C.c_long(0), C.c_size_t(0)
    ec, dc = C.c_int64(0), C.c_int64(0)
    cnt = L.fpvh_columnar_roundtrip(xsize, ysize, shift, int(big_endian), frames_per_batch, image_type, int(unshift), _p(frames),
                                    _p(ts), n, _p(out), _p(out_ts), out.size, C.byref(bpi), C.byref(batches), C.byref(ec),
                                    C.byref(dc), C.byref(comp))
    if cnt < 0:
        raise HostError(f"columnar batch round trip failed: {last_error()}")
    images = out[:cnt * per].reshape(cnt, per)
    if image_type == IMAGE_FULL:
        images = images.view(np.uint16)
    return images, out_ts[:cnt], {"batches": batches.value, "encoder_close": ec.value, "decoder_close": dc.value,
                                  "compressed_bytes": comp.value, "bytes_per_image": bpi.value}


def columnar_planes(frames, timestamps, xsize, ysize, shift=0, big_endian=False, frames_per_batch=10):
    """frames -> ColumnarBatchEncoder -> Batches; the plane columns of every batch brotli-decoded again.
    Returns (flags[n], high[n][P], low[n][P] (zeros where a frame has no low plane), preview[n][P/16])."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, xsize * ysize)
    ts = np.ascontiguousarray(timestamps, dtype=np.int64)
    n, P, PP = frames.shape[0], xsize * ysize, (xsize // 4) * (ysize // 4)
    flags = np.zeros(n, np.uint8)
    high, low, preview = np.zeros((n, P), np.uint8), np.zeros((n, P), np.uint8), np.zeros((n, PP), np.uint8)
    cnt = L.fpvh_columnar_planes(xsize, ysize, shift, int(big_endian), frames_per_batch, _p(frames), _p(ts), n, _p(flags),
                                 _p(high), _p(low), _p(preview))
    if cnt != n:
        raise HostError(f"columnar planes probe failed ({cnt} of {n} frames): {last_error()}")
    return flags, high, low, preview


def scan_coded_plane(stream, plane_bytes):
    """The host decoders' walk over the chunk directories of one plane stream (GPU entropy coder): (chunk offsets,
    stream length), or None if the stream carries none."""
    L = lib()
    buf = np.frombuffer(bytes(stream), np.uint8)
    cap = (plane_bytes + 65535) // 65536 + 1
    offs = np.zeros(cap, np.uint64)
    n, length = C.c_size_t(0), C.c_size_t(0)
    if not L.fpvh_scan_coded_plane(_p(buf), buf.size, plane_bytes, _p(offs), cap, C.byref(n), C.byref(length)):
        return None
    return [int(x) for x in offs[:n.value]], length.value
