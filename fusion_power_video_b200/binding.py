"""ctypes binding of the C ABI in include/fpv_b200.h.

Every function here forwards to ``libfpv_b200.so``; host arrays are numpy,
device buffers are raw pointers (ints), e.g. ``torch.Tensor.data_ptr()``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# FPV_B200_LIB: another build of the same library (tuning experiments: `make LIBDIR=../lib_x EXTRA=-D...`)
_LIB_PATH = os.environ.get("FPV_B200_LIB") or os.path.join(_HERE, "lib", "libfpv_b200.so")

ENC_DEFAULT, ENC_NO_DELTA, ENC_GENERIC = 0, 1, 2
DEC_DEFAULT, DEC_UNEXTRACT = 0, 1
FLAG_USE_DELTA, FLAG_USE_CG, FLAG_NO_LOW_BYTES = 1, 2, 4
IPC_HANDLE_BYTES = 64

_lib = None


class FpvError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"fpv_b200 error {code}: {message}")
        self.code = code


def lib_path() -> str:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Loads libfpv_b200.so.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C fusion_power_video_b200/csrc`). There is no CPU fallback."
        )
    L = C.CDLL(_LIB_PATH)
    vp, u32, i32, u64, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64, C.c_size_t
    L.fpv_create.argtypes = [C.POINTER(vp), i32, u32, u32, i32, i32, u32]
    L.fpv_create.restype = i32
    L.fpv_destroy.argtypes = [vp]
    L.fpv_destroy.restype = None
    L.fpv_last_error.argtypes = [vp]
    L.fpv_last_error.restype = C.c_char_p
    L.fpv_device_count.argtypes = []
    L.fpv_device_count.restype = i32
    L.fpv_version.argtypes = []
    L.fpv_version.restype = C.c_char_p
    L.fpv_plane_bytes.argtypes = [vp]
    L.fpv_plane_bytes.restype = sz
    L.fpv_preview_bytes.argtypes = [vp]
    L.fpv_preview_bytes.restype = sz
    L.fpv_kernel_launches.argtypes = [vp]
    L.fpv_kernel_launches.restype = u64
    L.fpv_enable_kernel_timing.argtypes = [vp, i32]
    L.fpv_enable_kernel_timing.restype = i32
    L.fpv_read_kernel_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u32)]
    L.fpv_read_kernel_timing.restype = i32
    L.fpv_host_alloc.argtypes = [sz]
    L.fpv_host_alloc.restype = vp
    L.fpv_host_free.argtypes = [vp]
    L.fpv_host_free.restype = None
    L.fpv_set_delta_raw.argtypes = [vp, vp]
    L.fpv_set_delta_raw.restype = i32
    L.fpv_set_delta_raw_device.argtypes = [vp, vp, vp]
    L.fpv_set_delta_raw_device.restype = i32
    L.fpv_set_delta_image.argtypes = [vp, vp]
    L.fpv_set_delta_image.restype = i32
    L.fpv_set_delta_image_device.argtypes = [vp, vp, vp]
    L.fpv_set_delta_image_device.restype = i32
    L.fpv_copy_delta_peer.argtypes = [vp, vp]
    L.fpv_copy_delta_peer.restype = i32
    L.fpv_delta_ipc_export.argtypes = [vp, vp]
    L.fpv_delta_ipc_export.restype = i32
    L.fpv_delta_ipc_import.argtypes = [vp, vp]
    L.fpv_delta_ipc_import.restype = i32
    L.fpv_bind_thread.argtypes = [vp]
    L.fpv_bind_thread.restype = i32
    L.fpv_device_of.argtypes = [vp]
    L.fpv_device_of.restype = i32
    L.fpv_encode.argtypes = [vp, vp, u32, u32, vp, vp, vp, vp]
    L.fpv_encode.restype = i32
    L.fpv_encode_device.argtypes = [vp, vp, u32, u32, vp, vp, vp, vp, vp]
    L.fpv_encode_device.restype = i32
    L.fpv_encode_submit.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp, vp]
    L.fpv_encode_submit.restype = i32
    L.fpv_wait.argtypes = [vp, u32]
    L.fpv_wait.restype = i32
    L.fpv_stream_bound.argtypes = [vp, u32]
    L.fpv_stream_bound.restype = sz
    L.fpv_entropy_device.argtypes = [vp, vp, vp, vp, vp, u32, vp, sz, vp, vp]
    L.fpv_entropy_device.restype = i32
    L.fpv_encode_stream_submit.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp, sz]
    L.fpv_encode_stream_submit.restype = i32
    L.fpv_decode.argtypes = [vp, vp, vp, vp, u32, u32, vp]
    L.fpv_decode.restype = i32
    L.fpv_decode_device.argtypes = [vp, vp, vp, vp, u32, u32, vp, vp]
    L.fpv_decode_device.restype = i32
    L.fpv_decode_submit.argtypes = [vp, u32, vp, vp, vp, u32, u32, vp]
    L.fpv_decode_submit.restype = i32
    L.fpv_encode_submit_v.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp, vp]
    L.fpv_encode_submit_v.restype = i32
    L.fpv_encode_stream_submit_v.argtypes = [vp, u32, vp, u32, u32, vp, vp, vp, sz]
    L.fpv_encode_stream_submit_v.restype = i32
    L.fpv_host_is_pinned.argtypes = [vp]
    L.fpv_host_is_pinned.restype = i32
    L.fpv_decode_coded.argtypes = [vp, vp, sz, vp, u32, vp, u32, u32, vp]
    L.fpv_decode_coded.restype = i32
    L.fpv_unpredict_planes.argtypes = [vp, vp, vp, vp, vp, u32]
    L.fpv_unpredict_planes.restype = i32
    _lib = L
    return L


def version() -> str:
    return lib().fpv_version().decode()


def device_count() -> int:
    return int(lib().fpv_device_count())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return C.c_void_p(a.ctypes.data)


class PinnedArray:
    """numpy view over pinned host memory from fpv_host_alloc."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().fpv_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("fpv_host_alloc failed")
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().fpv_host_free(self.ptr)
            self.ptr = None


class Context:
    """One fpv_ctx: a device, a frame geometry, a resident delta frame."""

    def __init__(self, xsize, ysize, shift=0, big_endian=False, max_batch=64, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        rc = self._L.fpv_create(C.byref(self._h), device, xsize, ysize, shift, int(bool(big_endian)), max_batch)
        if rc != 0:
            raise FpvError(rc, self._L.fpv_last_error(None).decode())
        self.xsize, self.ysize, self.shift, self.big_endian = xsize, ysize, shift, bool(big_endian)
        self.max_batch, self.device = max_batch, device
        self.P = xsize * ysize
        self.PP = (xsize // 4) * (ysize // 4)
        self.has_low = shift != 8

    # -- plumbing -----------------------------------------------------------
    def close(self):
        if self._h:
            self._L.fpv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise FpvError(rc, self._L.fpv_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    @property
    def kernel_launches(self) -> int:
        return int(self._L.fpv_kernel_launches(self._h))

    def enable_kernel_timing(self, on=True):
        self._check(self._L.fpv_enable_kernel_timing(self._h, int(on)))

    def read_kernel_timing(self):
        """(total device ms of the dominant kernel, number of launches) since the last read."""
        ms, cnt = C.c_double(0), C.c_uint32(0)
        self._check(self._L.fpv_read_kernel_timing(self._h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # -- delta frame --------------------------------------------------------
    def set_delta_raw(self, raw):
        if raw is not None:
            raw = np.ascontiguousarray(raw, dtype=np.uint16).reshape(-1)
            assert raw.size == self.P
        self._check(self._L.fpv_set_delta_raw(self._h, _ptr(raw)))

    def set_delta_raw_device(self, ptr, stream=0):
        self._check(self._L.fpv_set_delta_raw_device(self._h, _ptr(ptr), C.c_void_p(stream)))

    def set_delta_image(self, img):
        if img is not None:
            img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
            assert img.size == self.P
        self._check(self._L.fpv_set_delta_image(self._h, _ptr(img)))

    def set_delta_image_device(self, ptr, stream=0):
        self._check(self._L.fpv_set_delta_image_device(self._h, _ptr(ptr), C.c_void_p(stream)))

    def copy_delta_from(self, other: "Context"):
        self._check(self._L.fpv_copy_delta_peer(self._h, other._h))

    def delta_ipc_export(self) -> bytes:
        """CUDA IPC handle (64 bytes) of the resident delta image, for fpv_delta_ipc_import in ANOTHER process."""
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._L.fpv_delta_ipc_export(self._h, buf))
        return buf.raw

    def delta_ipc_import(self, handle: bytes):
        """Maps the exporting process's delta image and copies it device to device (peer copy over NVLink / PCIe)."""
        assert len(handle) == IPC_HANDLE_BYTES
        buf = C.create_string_buffer(bytes(handle), IPC_HANDLE_BYTES)
        self._check(self._L.fpv_delta_ipc_import(self._h, buf))

    def bind_thread(self):
        self._check(self._L.fpv_bind_thread(self._h))

    # -- encode -------------------------------------------------------------
    def encode(self, frames, options=ENC_DEFAULT):
        """frames: uint16 [n, ysize*xsize] (host).  Returns flags, high, low, preview."""
        frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, self.P)
        n = frames.shape[0]
        flags = np.zeros(n, np.uint8)
        high = np.zeros((n, self.P), np.uint8)
        low = np.zeros((n, self.P), np.uint8) if self.has_low else None
        preview = np.zeros((n, self.PP), np.uint8)
        self._check(self._L.fpv_encode(self._h, _ptr(frames), n, options, _ptr(flags), _ptr(high), _ptr(low), _ptr(preview)))
        return flags, high, low, preview

    def encode_device(self, frames_ptr, n, flags_ptr, high_ptr, low_ptr, preview_ptr, options=ENC_DEFAULT, stream=0):
        self._check(
            self._L.fpv_encode_device(
                self._h, _ptr(frames_ptr), n, options, _ptr(flags_ptr), _ptr(high_ptr),
                _ptr(low_ptr) if low_ptr else None, _ptr(preview_ptr), C.c_void_p(stream),
            )
        )

    def encode_submit(self, slot, frames, n, flags, high, low, preview, options=ENC_DEFAULT):
        self._check(self._L.fpv_encode_submit(self._h, slot, _ptr(frames), n, options, _ptr(flags), _ptr(high), _ptr(low), _ptr(preview)))

    def wait(self, slot):
        self._check(self._L.fpv_wait(self._h, slot))

    # -- GPU entropy coding ---------------------------------------------------
    def stream_bound(self, n) -> int:
        return int(self._L.fpv_stream_bound(self._h, n))

    def entropy_device(self, flags_ptr, high_ptr, low_ptr, preview_ptr, n, out_ptr, capacity, frame_off_ptr, stream=0):
        self._check(
            self._L.fpv_entropy_device(
                self._h, _ptr(flags_ptr), _ptr(high_ptr), _ptr(low_ptr) if low_ptr else None, _ptr(preview_ptr), n,
                _ptr(out_ptr), capacity, _ptr(frame_off_ptr), C.c_void_p(stream),
            )
        )

    def encode_stream_submit(self, slot, frames, n, flags, frame_off, out, capacity, options=ENC_DEFAULT):
        self._check(self._L.fpv_encode_stream_submit(self._h, slot, _ptr(frames), n, options, _ptr(flags), _ptr(frame_off),
                                                     _ptr(out), capacity))

    def encode_stream(self, frames, options=ENC_DEFAULT):
        """frames: uint16 [n, P] (host), n <= max_batch.  Returns flags and the list of the frames' container chunks
        (bytes), entropy-coded on the GPU."""
        frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, self.P)
        n = frames.shape[0]
        cap = self.stream_bound(n)
        flags = np.zeros(n, np.uint8)
        off = np.zeros(n + 1, np.uint64)
        out = np.zeros(cap, np.uint8)
        self.encode_stream_submit(0, frames, n, flags, off, out, cap, options)
        self.wait(0)
        return flags, [out[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)]

    # -- decode -------------------------------------------------------------
    def decode(self, high, low, flags, options=DEC_DEFAULT):
        high = np.ascontiguousarray(high, dtype=np.uint8).reshape(-1, self.P)
        n = high.shape[0]
        if low is not None:
            low = np.ascontiguousarray(low, dtype=np.uint8).reshape(-1, self.P)
        flags = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
        assert flags.size == n
        out = np.zeros((n, self.P), np.uint16)
        self._check(self._L.fpv_decode(self._h, _ptr(high), _ptr(low), _ptr(flags), n, options, _ptr(out)))
        return out

    def decode_coded(self, blob, chunks, flags, options=DEC_DEFAULT):
        """fpv_decode_coded: `blob` = coded bytes, `chunks` = list of (offset, frame, plane, index)."""
        blob = np.ascontiguousarray(np.frombuffer(bytes(blob), dtype=np.uint8))
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        n = flags.size
        tab = np.zeros((len(chunks), 6), np.uint32)
        for i, (off, f, pl, ix) in enumerate(chunks):
            tab[i] = (off & 0xFFFFFFFF, off >> 32, f, pl, ix, 0)
        out = np.zeros((n, self.P), np.uint16)
        self._check(self._L.fpv_decode_coded(self._h, _ptr(blob), blob.size, _ptr(tab), len(chunks), _ptr(flags), n, options,
                                             _ptr(out)))
        return out

    def decode_device(self, high_ptr, low_ptr, flags_ptr, n, out_ptr, options=DEC_DEFAULT, stream=0):
        self._check(
            self._L.fpv_decode_device(
                self._h, _ptr(high_ptr), _ptr(low_ptr) if low_ptr else None, _ptr(flags_ptr), n, options,
                _ptr(out_ptr), C.c_void_p(stream),
            )
        )

    def decode_submit(self, slot, high, low, flags, n, out, options=DEC_DEFAULT):
        self._check(self._L.fpv_decode_submit(self._h, slot, _ptr(high), _ptr(low), _ptr(flags), n, options, _ptr(out)))

    def unpredict_planes(self, high, low, preview, flags):
        high = np.ascontiguousarray(high, dtype=np.uint8).reshape(-1, self.P).copy()
        n = high.shape[0]
        low = None if low is None else np.ascontiguousarray(low, dtype=np.uint8).reshape(-1, self.P).copy()
        preview = None if preview is None else np.ascontiguousarray(preview, dtype=np.uint8).reshape(-1, self.PP).copy()
        flags = np.ascontiguousarray(flags, dtype=np.uint8).reshape(-1)
        self._check(self._L.fpv_unpredict_planes(self._h, _ptr(high), _ptr(low), _ptr(preview), _ptr(flags), n))
        return high, low, preview
