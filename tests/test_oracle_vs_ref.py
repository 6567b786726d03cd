"""Pins the C restatement (oracle/fpv_oracle.c) against the unmodified
reference compiled in place (oracle/_ref).  Runs only where the reference has
been built (this container); the same pin is carried to the GPU box by the
committed tests/golden vectors (test_golden.py)."""
import hashlib

import numpy as np
import pytest

from cases import CASE_NAMES, make_case
from oracle_binding import Oracle, Ref, _p, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def ref():
    return Ref()


def test_clamped_gradient_exhaustive(oracle, ref):
    import ctypes as C

    a = np.zeros(1 << 24, np.uint8)
    b = np.zeros(1 << 24, np.uint8)
    oracle.L.fpvo_cg_table.argtypes = [C.c_void_p]
    ref.L.ref_cg_table.argtypes = [C.c_void_p]
    oracle.L.fpvo_cg_table(_p(a))
    ref.L.ref_cg_table(_p(b))
    assert np.array_equal(a, b)
    # the two closed forms the kernels use (fpv_common.cuh cg1/cg2, fpv_decode.cu chain)
    n, w, nw = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    form1 = n + w - np.clip(nw, np.minimum(n, w), np.maximum(n, w))
    form2 = n + w - np.clip(w, np.minimum(n, nw), np.maximum(n, nw))
    assert np.array_equal(form1.reshape(-1).astype(np.uint8), b)
    assert np.array_equal(form2.reshape(-1).astype(np.uint8), b)
    assert form1.min() >= 0 and form1.max() <= 255


def test_estimate_entropy_random(oracle, ref):
    rng = np.random.default_rng(5)
    for k in range(3000):
        h = np.zeros(256, np.uint64)
        m = rng.integers(1, 257)
        idx = rng.choice(256, m, replace=False)
        h[idx] = rng.integers(0, 10 ** rng.integers(1, 8), m)
        assert oracle.estimate_entropy(h) == ref.estimate_entropy(h)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_predict_matches_reference(oracle, ref, name):
    c = make_case(name)
    W, H, shift, be = c["W"], c["H"], c["shift"], c["be"]
    frames = c["frames"].reshape(-1, W * H)
    for i in range(frames.shape[0]):
        fo = oracle.predict(frames[i], W, H, shift, be, c["delta"])
        fr = ref.predict(frames[i], W, H, shift, be, c["delta"])
        assert fo[0] == fr[0], f"flags frame {i}"
        assert np.array_equal(fo[1], fr[1]), "high"
        assert (fo[2] is None) == (fr[2] is None)
        if fo[2] is not None:
            assert np.array_equal(fo[2], fr[2]), "low"
        assert np.array_equal(fo[3], fr[3]), "preview"


@pytest.mark.parametrize("shift,be", [(0, 0), (8, 0), (3, 0), (7, 0), (12, 0), (16, 0), (0, 1), (8, 1), (1, 1), (4, 1), (7, 1)])
def test_split_all_values(oracle, ref, shift, be):
    img = np.arange(65536, dtype=np.uint16)
    fo = oracle.split(img, shift, be)
    fr = ref.split(img, 256, 256, shift, be)
    assert fo[0] == fr[0]
    assert np.array_equal(fo[1], fr[1])
    assert (fo[2] is None) == (fr[2] is None)
    if fo[2] is not None:
        assert np.array_equal(fo[2], fr[2])


@pytest.mark.parametrize("name", [n for n in CASE_NAMES if make_case(n)["delta"] is not None])
def test_inverse_matches_reference_decoder(oracle, ref, name):
    """oracle.inverse on the reference's own predicted planes == reference decode of its stream."""
    c = make_case(name)
    W, H, shift, be = c["W"], c["H"], c["shift"], c["be"]
    frames = c["frames"].reshape(-1, W * H)
    n = frames.shape[0]
    stream = ref.encode_stream(frames, W, H, shift, be, c["delta"], threads=1)
    nd, dec, _, _ = ref.decode_stream(stream, n, W, H, block=1000)
    assert nd == n
    dimg = oracle.delta_image(c["delta"], shift, be)
    for i in range(n):
        fl, h, l, p = ref.predict(frames[i], W, H, shift, be, c["delta"])
        low = None if (fl & 4) else l
        img = oracle.inverse(h, low, dimg, W, H, fl)
        assert np.array_equal(img, dec[i]), f"frame {i}"
        assert np.array_equal(oracle.unextract(img, shift, be), ref.unextract(dec[i], W, H, shift, be))


@pytest.mark.parametrize("name", ["plasma16_le0", "noise16_le0", "no_delta_frame", "be3_fullrange", "ramp"])
def test_unpredict_planes_matches_reference(oracle, ref, name):
    c = make_case(name)
    W, H, shift, be = c["W"], c["H"], c["shift"], c["be"]
    frames = c["frames"].reshape(-1, W * H)
    for i in range(frames.shape[0]):
        fl, h, l, p = ref.predict(frames[i], W, H, shift, be, c["delta"])
        dh = dl = None
        if c["delta"] is not None:
            dh, dl = oracle.delta_planes(c["delta"], shift, be)
        a = oracle.unpredict_planes(h, l, p, dh, dl, W, H, fl)
        b = ref.unpredict_planes(h, l, p, c["delta"], W, H, fl, shift, be)
        for x, y in zip(a, b):
            assert (x is None) == (y is None)
            if x is not None:
                assert np.array_equal(x, y)
