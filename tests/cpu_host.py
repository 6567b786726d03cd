"""TEST INFRASTRUCTURE: loads a second copy of fusion_power_video_b200/host.py bound to the host layer built on the
oracle-backed stand-in of the C ABI (tests/cpu_cabi -> tests/_build/libfpv_host_cpustub.so), so that host-side logic
(Encoder pipeline, ordering, sharding, the ingest harness, the columnar classes) can be exercised without a GPU.
Nothing here is reachable from the product package."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
_mod = None


def build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpu_cabi"), os.path.join("..", "_build", "libfpv_host_cpustub.so")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the CPU stand-in failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])


def host():
    """fusion_power_video_b200.host, re-imported under another name and pointed at the stand-in library."""
    global _mod
    if _mod is not None:
        return _mod
    build()
    path = os.path.join(ROOT, "fusion_power_video_b200", "host.py")
    spec = importlib.util.spec_from_file_location("fusion_power_video_b200._host_cpustub", path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "fusion_power_video_b200"
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    mod._LIB_PATH = os.path.join(BUILD, "libfpv_host_cpustub.so")
    mod._preload_cabi = lambda: None      # its DT_NEEDED is libfpv_cpustub.so next to it ($ORIGIN rpath)
    _mod = mod
    return mod
