"""bin/fpv_encode | bin/fpv_decode, the counterparts of the reference's encode.cc / decode.cc (same argv order:
xsize ysize big_endian shift), must round-trip raw frame files -- with host brotli and with the GPU entropy coder."""
import os
import subprocess

import numpy as np
import pytest

from fusion_power_video_b200 import host, synth

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fusion_power_video_b200", "bin")


@pytest.mark.parametrize("gpu_entropy", [0, 1], ids=["brotli", "gpu_entropy"])
@pytest.mark.parametrize("W,H,bits,shift,be", [(1280, 160, 12, 4, 0), (256, 128, 16, 0, 1)])
def test_encode_decode_cli_round_trip(tmp_path, W, H, bits, shift, be, gpu_entropy):
    if not os.path.exists(os.path.join(BIN, "fpv_encode")):
        pytest.skip("CLI tools not built")
    n = 11
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=5).reshape(n, -1)
    raw = frames.byteswap().tobytes() if be else frames.tobytes()
    enc = subprocess.run([os.path.join(BIN, "fpv_encode"), str(W), str(H), str(be), str(shift), "3", "4", str(gpu_entropy)],
                         input=raw, capture_output=True, timeout=120)
    assert enc.returncode == 0, enc.stderr.decode()[-500:]
    stream = enc.stdout
    # like encode.cc:87-92 the first frame doubles as the delta frame
    lib_stream = host.encode_stream(frames.byteswap() if be else frames, W, H, shift, bool(be), threads=3, batch=4,
                                    gpu_entropy=bool(gpu_entropy))
    assert stream == lib_stream
    dec = subprocess.run([os.path.join(BIN, "fpv_decode"), str(W), str(H), str(be), str(shift)], input=stream,
                         capture_output=True, timeout=120)
    assert dec.returncode == 0, dec.stderr.decode()[-500:]
    assert dec.stdout == raw, "decoded raw file differs from the input"
