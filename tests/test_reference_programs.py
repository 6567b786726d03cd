"""Whole-program parity: the reference's UNMODIFIED programs (encode.cc, decode.cc, benchmark.cc,
columnar_batch/*_test.cc) compiled against the mirror header + this repository's libraries must behave like the
same programs compiled against the reference itself.

Three builds of every program exist (made by `__graft_entry__.build()` / the Makefiles, in this container where
/root/reference exists; the GPU box runs the prebuilt files):

    oracle/_ref/bin/ref_X    reference main + reference library, CPU              -> the expected behaviour
    oracle/_ref/bin/gpu_X    reference main + mirror header + the B200 libraries  -> `-m gpu` tests
    tests/_build/cpu_X       reference main + mirror header + the host layer on the oracle-backed stand-in of the
                             C ABI (tests/cpu_cabi, test infrastructure)          -> CPU tests of the HOST layer

plus tests/frame_parity.cc (a walk through the public fpvc::Frame surface) in the same three builds.
"""
import os
import subprocess

import numpy as np
import pytest

from fusion_power_video_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "bin")
CPUBIN = os.path.join(ROOT, "tests", "_build")

CONFIGS = [
    # W, H, bits, shift, big_endian, frames, threads
    (256, 128, 12, 4, 0, 9, 4),
    (128, 96, 16, 0, 1, 7, 3),
    (64, 64, 8, 8, 0, 5, 0),       # shift 8: no low plane; 0 threads: the synchronous path
    (1280, 160, 12, 4, 0, 21, 8),
]


@pytest.fixture(scope="session")
def cpu_programs():
    """Builds the stand-in C ABI, the host layer on it and the reference's programs on both (needs gcc; the
    programs need /root/reference)."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpu_cabi"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    if not os.path.exists(os.path.join(CPUBIN, "cpu_encode")) or not os.path.exists(os.path.join(REFBIN, "ref_encode")):
        pytest.skip("the reference's programs are not built (no /root/reference here and no prebuilt oracle/_ref/bin)")
    return CPUBIN + "/cpu_"


@pytest.fixture(scope="session")
def gpu_programs():
    if not os.path.exists(os.path.join(REFBIN, "gpu_encode")) or not os.path.exists(os.path.join(REFBIN, "ref_encode")):
        pytest.skip("oracle/_ref/bin is not built (run __graft_entry__.build() where /root/reference exists)")
    return REFBIN + "/gpu_"


def run(exe, args, stdin=None, timeout=300, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe] + [str(a) for a in args], input=stdin, capture_output=True, timeout=timeout, env=e)


def raw_file(W, H, bits, be, n, seed=11):
    frames = synth.plasma_frames(n, W, H, bits=bits, seed=seed).reshape(n, -1)
    return (frames.byteswap() if be else frames).tobytes()


def check_encode_decode(prefix, tmp_path, W, H, bits, shift, be, n, threads, env=None):
    raw = raw_file(W, H, bits, be, n)
    args = [W, H, be, shift, threads]      # the order encode.cc actually parses (encode.cc:41-48)
    ref = run(REFBIN + "/ref_encode", args, raw)
    ours = run(prefix + "encode", args, raw, env=env)
    assert ref.returncode == 0 and ours.returncode == 0, ours.stderr.decode()[-800:]
    assert ours.stdout == ref.stdout, "stream of the reference's encode.cc on this library differs from the reference's"
    # decode.cc both ways
    dargs = [W, H, be, shift]
    back = run(prefix + "decode", dargs, ref.stdout, env=env)
    assert back.returncode == 0, back.stderr.decode()[-800:]
    assert back.stdout == raw, "the reference's decode.cc on this library does not reproduce the raw file"
    back_ref = run(REFBIN + "/ref_decode", dargs, ours.stdout)
    assert back_ref.returncode == 0 and back_ref.stdout == raw


def check_benchmark(prefix, tmp_path, env=None):
    W, H, bits, shift, be, n = 256, 128, 12, 4, 0, 12
    path = tmp_path / "frames.raw"
    path.write_bytes(raw_file(W, H, bits, be, n))
    outs = []
    for exe in (REFBIN + "/ref_benchmark", prefix + "benchmark"):
        r = run(exe, [path, W, H, be, shift], env=env if exe.startswith(prefix) else None)
        assert r.returncode == 0, r.stderr.decode()[-800:]
        text = r.stdout.decode() + r.stderr.decode()
        # benchmark.cc:189-285: sizes, then "ok" for the streaming and the random access decoder
        assert text.count("\nok") == 2, text[-600:]
        outs.append([ln.split(", time")[0] for ln in text.splitlines() if ln.startswith("total:")])
    assert outs[0] == outs[1] and outs[0], "benchmark.cc reports different stream sizes: %r" % (outs,)


def check_frame_parity(prefix):
    ref = run(REFBIN + "/ref_frame_parity", [])
    ours = run(prefix + "frame_parity", [])
    assert ref.returncode == 0 and ours.returncode == 0, ours.stderr.decode()[-800:]
    a, b = ref.stdout.decode().splitlines(), ours.stdout.decode().splitlines()
    assert len(a) > 150
    for x, y in zip(a, b):
        assert x == y, f"fpvc::Frame differs from the reference:\n  reference: {x}\n  this repo: {y}"
    assert len(a) == len(b)


def check_columnar(prefix):
    # (1) the reference's columnar test mains on OUR columnar_batch mirror (batched GPU calls): the decoder test's
    #     three images come back without a single "Bad Pixel" (columnar_batch_decoder_test.cc:19-27)
    r = run(prefix + "mirror_columnar_batch_decoder_test", [], timeout=120)
    text = r.stdout.decode()
    assert r.returncode == 0, r.stderr.decode()[-800:]
    assert "Bad Pixel" not in text
    for k, ts in ((1, 123456), (2, 234567), (3, 345678)):
        assert f"Got the Image {k}! {ts}" in text
    assert "Closed Encoder - 345678." in text and "Closed Decoder - 345678." in text
    r = run(prefix + "mirror_columnar_batch_encoder_test", [], timeout=300)
    text = r.stdout.decode()
    assert r.returncode == 0, r.stderr.decode()[-800:]
    assert "Closed - 234567." in text and "Closed - 499." in text
    assert text.count("Got the Batch!") == 1 + (500 + 12) // 13
    # (2) the reference's OWN columnar library sources on the fpvc::Frame facade.  In the reference's own build both
    #     test programs die of the defects SURVEY.md section 2 (#17) lists (compressing into a zero-sized vector,
    #     columnar_batch.cc:10-22 -- undefined behaviour, so nothing is asserted about a crashing reference); where
    #     the reference build survives, the facade build must print the same images.
    for t in ("columnar_batch_decoder_test", "columnar_batch_encoder_test"):
        ref = run(REFBIN + "/ref_" + t, [], timeout=300)
        if ref.returncode != 0:
            continue
        # the undefined behaviour sits in the reference's columnar_batch.cc, which is compiled into BOTH programs: a
        # run may die of it by chance (seen: SIGSEGV in about one run in ten), so a crashed run is repeated
        for _ in range(4):
            ours = run(prefix + t, [], timeout=300)
            if ours.returncode >= 0:
                break
        assert ours.returncode == 0, (t, ours.returncode)
        if t == "columnar_batch_decoder_test":
            pick = lambda s: [ln for ln in s.decode().splitlines() if ln.startswith(("Got the Image", "Bad Pixel"))]
            assert not any(ln.startswith("Bad Pixel") for ln in pick(ours.stdout)), "facade build decoded a wrong pixel"
            # the reference build's undefined behaviour sometimes survives with corrupted images ("Bad Pixel" lines,
            # seen on one run in ~10): only a clean reference run is a meaningful expectation
            if not any(ln.startswith("Bad Pixel") for ln in pick(ref.stdout)):
                assert pick(ref.stdout) == pick(ours.stdout)


# ---- CPU: the host layer on the stand-in C ABI ----------------------------------------------------------------------

@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "x".join(map(str, c)))
def test_cpu_reference_encode_decode_on_mirror(cpu_programs, tmp_path, cfg):
    check_encode_decode(cpu_programs, tmp_path, *cfg)


@pytest.mark.parametrize("gpus", [2, 3])
def test_cpu_one_encoder_over_several_devices_writes_the_same_stream(cpu_programs, tmp_path, gpus):
    """GpuOptions::devices (here through FPV_GPUS, which the unmodified encode.cc cannot set any other way): batches
    go round-robin to per-device contexts and GPU threads, emission stays in submission order."""
    check_encode_decode(cpu_programs, tmp_path, 256, 128, 12, 4, 0, 37, 6, env={"FPV_GPUS": str(gpus)})


def test_cpu_reference_benchmark_on_mirror(cpu_programs, tmp_path):
    check_benchmark(cpu_programs, tmp_path)


def test_cpu_frame_facade_matches_reference(cpu_programs):
    check_frame_parity(cpu_programs)


def test_cpu_columnar_programs(cpu_programs):
    check_columnar(cpu_programs)


# ---- GPU: the same programs on the B200 libraries -----------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "x".join(map(str, c)))
def test_gpu_reference_encode_decode_on_mirror(gpu_programs, tmp_path, cfg):
    check_encode_decode(gpu_programs, tmp_path, *cfg)


@pytest.mark.gpu
def test_gpu_reference_encode_on_two_gpus_of_one_encoder(gpu_programs, tmp_path):
    """The unmodified encode.cc driving two GPUs through ONE fpvc::Encoder (FPV_GPUS=2): same bytes."""
    import fusion_power_video_b200 as fpv

    if fpv.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    check_encode_decode(gpu_programs, tmp_path, 1280, 160, 12, 4, 0, 40, 8, env={"FPV_GPUS": "2"})


@pytest.mark.gpu
def test_gpu_reference_benchmark_on_mirror(gpu_programs, tmp_path):
    check_benchmark(gpu_programs, tmp_path)


@pytest.mark.gpu
def test_gpu_frame_facade_matches_reference(gpu_programs):
    check_frame_parity(gpu_programs)


@pytest.mark.gpu
def test_gpu_columnar_programs(gpu_programs):
    check_columnar(gpu_programs)
