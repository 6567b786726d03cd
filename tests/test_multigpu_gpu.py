"""Two-GPU tests (skipped on a single-GPU box): frame-range sharding over devices gives exactly the
single-GPU result; the delta frame is the only thing that crosses GPUs (fpv_copy_delta_peer)."""
import numpy as np
import pytest

import fusion_power_video_b200 as fpv
from fusion_power_video_b200 import host, sharding, synth

pytestmark = pytest.mark.gpu


def _need_two():
    if fpv.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def test_planes_sharded_over_two_devices_match_one_device():
    _need_two()
    W, H, shift, n = 1280, 160, 4, 10
    frames = synth.plasma_frames(n, W, H, bits=12, seed=5).reshape(n, -1)
    with fpv.Context(W, H, shift, False, max_batch=n, device=0) as c0:
        c0.set_delta_raw(frames[0])
        want = c0.encode(frames)
        with fpv.Context(W, H, shift, False, max_batch=n, device=1) as c1:
            c1.copy_delta_from(c0)          # peer copy of the resident delta planes
            a, b = sharding.frame_range(n, 2, 1)
            got1 = c1.encode(frames[a:b])
            got0 = c0.encode(frames[:a])
            for k in range(4):
                assert np.array_equal(np.concatenate([got0[k], got1[k]]), want[k])
            # decode on the other device
            c1.set_delta_image(None)
        with fpv.Context(W, H, shift, False, max_batch=n, device=1) as c1:
            c1.copy_delta_from(c0)
            raw = c1.decode(want[1], want[2], want[0], fpv.DEC_UNEXTRACT)
            assert np.array_equal(raw, frames)


def test_streams_sharded_over_two_devices_merge_to_the_single_stream():
    _need_two()
    W, H, shift, n = 640, 96, 0, 11
    frames = synth.plasma_frames(n, W, H, bits=16, seed=6).reshape(n, -1)
    single = host.encode_stream(frames, W, H, shift, threads=2, batch=4, device=0)
    parts = []
    for r in range(2):
        a, b = sharding.frame_range(n, 2, r)
        parts.append(host.encode_stream(frames[a:b], W, H, shift, threads=2, batch=4, delta=frames[0], device=r))
    header, chunks0 = sharding.split_stream(parts[0])
    header1, chunks1 = sharding.split_stream(parts[1])
    assert header == header1
    assert sharding.merge_shards(header, [chunks0, chunks1]) == single


def test_one_encoder_over_two_devices_writes_the_single_gpu_stream():
    """GpuOptions::devices: batches alternate between two GPUs (own context and GPU thread each), the delta frame is
    uploaded to the first and peer-copied to the second, emission order is the submission order."""
    _need_two()
    W, H, shift, n = 640, 96, 0, 45
    frames = synth.plasma_frames(n, W, H, bits=16, seed=7).reshape(n, -1)
    single = host.encode_stream(frames, W, H, shift, threads=4, batch=4, device=0)
    for ge in (False, True):
        one = host.encode_stream_multi(frames, W, H, shift, threads=4, batch=4, devices=(0,), gpu_entropy=ge)
        two = host.encode_stream_multi(frames, W, H, shift, threads=4, batch=4, devices=(0, 1), gpu_entropy=ge)
        assert two == one
        if not ge:
            assert two == single
        back = host.decode_stream(two, n, W, H, block=0, batch=8, raw_shift=shift, device=1)
        assert np.array_equal(back, frames)


_IPC_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["FPV_ROOT"])
import fusion_power_video_b200 as fpv
from fusion_power_video_b200 import synth, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
W, H, shift, n = 1280, 160, 4, 12
frames = synth.plasma_frames(n, W, H, bits=12, seed=9).reshape(n, -1)
ctx = fpv.Context(W, H, shift, False, max_batch=n, device=rank)
box = [None]
if rank == 0:
    ctx.set_delta_raw(frames[0])
    box = [ctx.delta_ipc_export()]
dist.broadcast_object_list(box, src=0)
if rank != 0:
    ctx.delta_ipc_import(box[0])          # device-to-device copy out of rank 0's memory
dist.barrier()
a, b = sharding.frame_range(n, world, rank)
got = ctx.encode(frames[a:b])
parts = [None] * world
dist.all_gather_object(parts, [np.ascontiguousarray(x) for x in got])
if rank == 0:
    want = ctx.encode(frames)
    for k in range(4):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), want[k]), k
    print("IPC_SHARDS_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_delta_frame_crosses_processes_by_cuda_ipc(tmp_path):
    """One process per GPU (torchrun): rank 0 exports its resident delta image, rank 1 imports it
    (fpv_delta_ipc_export / _import) and encodes its frame range; the shards equal the single-GPU planes."""
    import os
    import subprocess
    import sys

    _need_two()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ipc_worker.py"
    script.write_text(_IPC_WORKER)
    env = dict(os.environ, FPV_ROOT=root)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "IPC_SHARDS_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
