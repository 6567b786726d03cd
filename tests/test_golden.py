"""The oracle against the committed golden vectors (made by
tests/golden/make_golden.py from the unmodified reference).  Runs anywhere."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle_binding import Oracle, _p

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz")))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def test_golden_present():
    assert len(CASES) >= 15
    assert os.path.exists(os.path.join(GOLDEN, "scalars.npz"))


def test_scalars(oracle):
    import ctypes as C

    g = np.load(os.path.join(GOLDEN, "scalars.npz"))
    for h, e in zip(g["hists"], g["entropy"]):
        assert oracle.estimate_entropy(h) == int(e)
    table = np.zeros(1 << 24, np.uint8)
    oracle.L.fpvo_cg_table.argtypes = [C.c_void_p]
    oracle.L.fpvo_cg_table(_p(table))
    assert np.array_equal(np.frombuffer(hashlib.sha256(table.tobytes()).digest(), np.uint8), g["cg_table_sha256"])
    assert np.array_equal(table[::4099], g["cg_table_sample"])


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[5:-4] for p in CASES])
def test_case(oracle, path):
    g = np.load(path)
    W, H, shift, be = int(g["W"]), int(g["H"]), int(g["shift"]), int(g["be"])
    delta = g["delta"] if int(g["has_delta"]) else None
    frames = g["frames"]
    for i in range(frames.shape[0]):
        fl, h, l, p = oracle.predict(frames[i], W, H, shift, be, delta)
        assert fl == int(g["flags"][i])
        assert np.array_equal(h, g["high"][i])
        if shift != 8:
            assert np.array_equal(l, g["low"][i])
        else:
            assert l is None
        assert np.array_equal(p, g["preview"][i])
        if delta is not None:
            dimg = oracle.delta_image(delta, shift, be)
            assert np.array_equal(dimg, g["delta_image"])
            low = None if (fl & 4) else l
            img = oracle.inverse(h, low, dimg, W, H, fl)
            assert np.array_equal(img, g["decoded"][i])
            assert np.array_equal(oracle.unextract(img, shift, be), g["unextracted"][i])
