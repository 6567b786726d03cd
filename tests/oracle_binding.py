"""ctypes bindings for the TEST-ONLY CPU checkers under oracle/.

 * Oracle : oracle/libfpv_oracle.so  -- the plain-C restatement (fpv_oracle.c)
 * Ref    : oracle/_ref/libfpv_ref.so -- the unmodified reference compiled in
            place (only where it has been built; never required on the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use
this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libfpv_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libfpv_ref.so")

vp, sz, i32, u8, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint8, C.c_uint64


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "fpv_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle"], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self):
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.fpvo_estimate_entropy.argtypes = [vp]
        L.fpvo_estimate_entropy.restype = u64
        L.fpvo_clamped_gradient.argtypes = [u8, u8, u8]
        L.fpvo_clamped_gradient.restype = u8
        L.fpvo_split.argtypes = [vp, sz, i32, i32, vp, vp]
        L.fpvo_split.restype = u8
        L.fpvo_preview.argtypes = [vp, sz, sz, vp]
        L.fpvo_preview.restype = None
        L.fpvo_delta_decide.argtypes = [vp, sz]
        L.fpvo_delta_decide.restype = i32
        L.fpvo_cg_decide.argtypes = [vp, sz, sz]
        L.fpvo_cg_decide.restype = i32
        L.fpvo_cg_forward.argtypes = [vp, sz, sz, vp]
        L.fpvo_cg_forward.restype = None
        L.fpvo_cg_inverse.argtypes = [vp, sz, sz]
        L.fpvo_cg_inverse.restype = None
        L.fpvo_predict.argtypes = [vp, sz, sz, i32, i32, vp, vp, vp, vp, vp, vp]
        L.fpvo_predict.restype = u8
        L.fpvo_inverse.argtypes = [vp, vp, vp, sz, sz, u8, vp]
        L.fpvo_inverse.restype = None
        L.fpvo_unextract.argtypes = [vp, sz, i32, i32, vp]
        L.fpvo_unextract.restype = None
        L.fpvo_unpredict_planes.argtypes = [vp, vp, vp, vp, vp, sz, sz, u8]
        L.fpvo_unpredict_planes.restype = None
        self.L = L

    def estimate_entropy(self, counts):
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        assert c.size == 256
        return int(self.L.fpvo_estimate_entropy(_p(c)))

    def clamped_gradient(self, n, w, nw):
        return int(self.L.fpvo_clamped_gradient(n, w, nw))

    def split(self, img, shift, big_endian):
        img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
        high = np.zeros(img.size, np.uint8)
        low = np.zeros(img.size, np.uint8) if shift != 8 else None
        flags = self.L.fpvo_split(_p(img), img.size, shift, int(big_endian), _p(high), _p(low))
        return int(flags), high, low

    def delta_planes(self, delta_raw, shift, big_endian):
        """Split planes of a raw delta frame (what Encoder::Init keeps, .cc:1097)."""
        _, dh, dl = self.split(delta_raw, shift, big_endian)
        return dh, dl

    def delta_image(self, delta_raw, shift, big_endian):
        """The delta frame as the decoder sees it: (high << 8) | low."""
        dh, dl = self.delta_planes(delta_raw, shift, big_endian)
        return (dh.astype(np.uint16) << 8) | (dl.astype(np.uint16) if dl is not None else 0)

    def predict(self, img, W, H, shift, big_endian, delta_raw=None):
        """Frame ctor + Predict.  Returns flags, high, low (None for shift 8), preview."""
        img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
        assert img.size == W * H
        dh = dl = None
        if delta_raw is not None:
            dh, dl = self.delta_planes(delta_raw, shift, big_endian)
        high = np.zeros(W * H, np.uint8)
        low = np.zeros(W * H, np.uint8) if shift != 8 else None
        preview = np.zeros((W // 4) * (H // 4), np.uint8)
        scratch = np.zeros(W * H, np.uint8)
        flags = self.L.fpvo_predict(_p(img), W, H, shift, int(big_endian), _p(dh), _p(dl), _p(high), _p(low), _p(preview), _p(scratch))
        return int(flags), high, low, preview

    def inverse(self, high, low, delta_image, W, H, flags):
        high = np.ascontiguousarray(high, dtype=np.uint8).reshape(-1).copy()
        low = None if low is None else np.ascontiguousarray(low, dtype=np.uint8).reshape(-1)
        d = None if delta_image is None else np.ascontiguousarray(delta_image, dtype=np.uint16).reshape(-1)
        img = np.zeros(W * H, np.uint16)
        self.L.fpvo_inverse(_p(high), _p(low), _p(d), W, H, flags, _p(img))
        return img

    def unextract(self, img, shift, big_endian):
        img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
        out = np.zeros(img.size * 2, np.uint8)
        self.L.fpvo_unextract(_p(img), img.size, shift, int(big_endian), _p(out))
        return out

    def cg_forward(self, plane, W):
        plane = np.ascontiguousarray(plane, dtype=np.uint8).reshape(-1)
        out = np.zeros_like(plane)
        self.L.fpvo_cg_forward(_p(plane), W, plane.size, _p(out))
        return out

    def cg_inverse(self, plane, W):
        plane = np.ascontiguousarray(plane, dtype=np.uint8).reshape(-1).copy()
        self.L.fpvo_cg_inverse(_p(plane), W, plane.size)
        return plane

    def unpredict_planes(self, high, low, preview, dh, dl, W, H, flags):
        high = np.ascontiguousarray(high, dtype=np.uint8).reshape(-1).copy()
        low = None if low is None else np.ascontiguousarray(low, dtype=np.uint8).reshape(-1).copy()
        preview = None if preview is None else np.ascontiguousarray(preview, dtype=np.uint8).reshape(-1).copy()
        self.L.fpvo_unpredict_planes(_p(high), _p(low), _p(preview), _p(dh), _p(dl), W, H, flags)
        return high, low, preview


def ref_available():
    return os.path.exists(REF_SO)


class Ref:
    """The unmodified reference through oracle/ref_harness.cc."""

    def __init__(self):
        if not ref_available():
            raise FileNotFoundError(REF_SO)
        L = C.CDLL(REF_SO)
        L.ref_estimate_entropy.argtypes = [vp]
        L.ref_estimate_entropy.restype = u64
        L.ref_clamped_gradient.argtypes = [u8, u8, u8]
        L.ref_clamped_gradient.restype = u8
        L.ref_predict.argtypes = [sz, sz, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        L.ref_predict.restype = i32
        L.ref_split.argtypes = [sz, sz, vp, i32, i32, vp, vp, vp]
        L.ref_split.restype = i32
        L.ref_unpredict_planes.argtypes = [sz, sz, u8, vp, vp, sz, vp, vp, i32, i32]
        L.ref_unpredict_planes.restype = None
        L.ref_decompress_image.argtypes = [vp, vp, sz, sz, sz, vp]
        L.ref_decompress_image.restype = i32
        L.ref_unextract.argtypes = [vp, sz, sz, i32, i32, vp]
        L.ref_unextract.restype = None
        L.ref_encode_stream.argtypes = [sz, sz, i32, i32, sz, vp, vp, sz, vp, sz]
        L.ref_encode_stream.restype = sz
        L.ref_decode_stream.argtypes = [vp, sz, sz, vp, sz, vp, vp]
        L.ref_decode_stream.restype = C.c_long
        L.ref_random_access_decode.argtypes = [vp, sz, sz, vp, vp, vp]
        L.ref_random_access_decode.restype = i32
        for name in ("ref_time_transform", "ref_time_transform_shared"):
            fn = getattr(L, name)
            fn.argtypes = [sz, sz, i32, i32, vp, vp, sz, sz, vp]
            fn.restype = C.c_double
        L.ref_time_encode.argtypes = [sz, sz, i32, i32, sz, vp, vp, sz, vp]
        L.ref_time_encode.restype = C.c_double
        L.ref_time_unpredict.argtypes = [sz, sz, u8, vp, vp, vp, i32, i32, sz, sz]
        L.ref_time_unpredict.restype = C.c_double
        L.ref_hardware_threads.restype = C.c_uint
        self.L = L

    def estimate_entropy(self, counts):
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        return int(self.L.ref_estimate_entropy(_p(c)))

    def clamped_gradient(self, n, w, nw):
        return int(self.L.ref_clamped_gradient(n, w, nw))

    def split(self, img, W, H, shift, big_endian):
        img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
        high = np.zeros(W * H, np.uint8)
        low = np.zeros(W * H, np.uint8)
        ls = sz(0)
        flags = self.L.ref_split(W, H, _p(img), shift, int(big_endian), _p(high), _p(low), C.byref(ls))
        return int(flags), high, (low if ls.value else None)

    def predict(self, img, W, H, shift, big_endian, delta_raw=None):
        img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
        d = None if delta_raw is None else np.ascontiguousarray(delta_raw, dtype=np.uint16).reshape(-1)
        high = np.zeros(W * H, np.uint8)
        low = np.zeros(W * H, np.uint8)
        preview = np.zeros(max((W // 4) * (H // 4), 1), np.uint8)
        ls, ps = sz(0), sz(0)
        flags = self.L.ref_predict(W, H, _p(img), shift, int(big_endian), _p(d), _p(high), _p(low), _p(preview), C.byref(ls), C.byref(ps))
        return int(flags), high, (low if ls.value else None), preview[: ps.value]

    def unpredict_planes(self, high, low, preview, delta_raw, W, H, flags, shift, big_endian):
        high = np.ascontiguousarray(high, dtype=np.uint8).reshape(-1).copy()
        low = None if low is None else np.ascontiguousarray(low, dtype=np.uint8).reshape(-1).copy()
        preview = None if preview is None else np.ascontiguousarray(preview, dtype=np.uint8).reshape(-1).copy()
        d = None if delta_raw is None else np.ascontiguousarray(delta_raw, dtype=np.uint16).reshape(-1)
        self.L.ref_unpredict_planes(W, H, flags, _p(high), _p(low), 0 if low is None else low.size, _p(preview), _p(d), shift, int(big_endian))
        return high, low, preview

    def decompress_image(self, delta_image, data, W, H):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        d = None if delta_image is None else np.ascontiguousarray(delta_image, dtype=np.uint16).reshape(-1)
        img = np.zeros(W * H, np.uint16)
        ok = self.L.ref_decompress_image(_p(d), _p(data), data.size, W, H, _p(img))
        return bool(ok), img

    def unextract(self, img, W, H, shift, big_endian):
        img = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1)
        out = np.zeros(W * H * 2, np.uint8)
        self.L.ref_unextract(_p(img), W, H, shift, int(big_endian), _p(out))
        return out

    def encode_stream(self, frames, W, H, shift, big_endian, delta_raw, threads=2):
        frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, W * H)
        d = np.ascontiguousarray(delta_raw, dtype=np.uint16).reshape(-1)
        cap = frames.size * 3 + 65536
        out = np.zeros(cap, np.uint8)
        n = self.L.ref_encode_stream(W, H, shift, int(big_endian), threads, _p(d), _p(frames), frames.shape[0], _p(out), cap)
        assert n <= cap
        return out[:n].copy()

    def decode_stream(self, data, max_frames, W, H, block=65536):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        frames = np.zeros((max_frames, W * H), np.uint16)
        wo, ho = sz(0), sz(0)
        n = self.L.ref_decode_stream(_p(data), data.size, block, _p(frames), max_frames, C.byref(wo), C.byref(ho))
        return int(n), frames, wo.value, ho.value

    def random_access_decode(self, data, index, W, H):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        frame = np.zeros(W * H, np.uint16)
        preview = np.zeros((W // 4) * (H // 4), np.uint8)
        nf = sz(0)
        ok = self.L.ref_random_access_decode(_p(data), data.size, index, _p(frame), _p(preview), C.byref(nf))
        return bool(ok), frame, preview, nf.value

    def time_transform(self, frames, W, H, shift, big_endian, delta_raw, threads, shared=True):
        frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, W * H)
        d = np.ascontiguousarray(delta_raw, dtype=np.uint16).reshape(-1)
        chk = u64(0)
        fn = self.L.ref_time_transform_shared if shared else self.L.ref_time_transform
        return float(fn(W, H, shift, int(big_endian), _p(d), _p(frames), frames.shape[0], threads, C.byref(chk)))

    def time_encode(self, frames, W, H, shift, big_endian, delta_raw, threads):
        frames = np.ascontiguousarray(frames, dtype=np.uint16).reshape(-1, W * H)
        d = np.ascontiguousarray(delta_raw, dtype=np.uint16).reshape(-1)
        ss = sz(0)
        t = float(self.L.ref_time_encode(W, H, shift, int(big_endian), threads, _p(d), _p(frames), frames.shape[0], C.byref(ss)))
        return t, ss.value

    def time_unpredict(self, high, low, W, H, flags, delta_raw, shift, big_endian, threads):
        high = np.ascontiguousarray(high, dtype=np.uint8).reshape(-1, W * H)
        low = np.ascontiguousarray(low, dtype=np.uint8).reshape(-1, W * H)
        d = np.ascontiguousarray(delta_raw, dtype=np.uint16).reshape(-1)
        return float(self.L.ref_time_unpredict(W, H, flags, _p(high), _p(low), _p(d), shift, int(big_endian), high.shape[0], threads))

    def hardware_threads(self):
        return int(self.L.ref_hardware_threads())
