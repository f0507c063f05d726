"""The C++ host-side mirror of the reference interface (coxgraph_b200/host/coxgraph_b200.hpp):
it compiles with plain g++ against the C ABI, fails loudly without a GPU, and — on the GPU box —
the reference's call sequence (tsdf_recover.h:59-99 then map_server.cpp:59-73) written in C++
gives the oracle's layers."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "host_api_check")


def _binary():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", ROOT, "-s", "build/host_api_check"])
    return BIN


def test_host_program_builds_and_fails_loudly_without_a_gpu():
    r = subprocess.run([_binary(), "nogpu"], capture_output=True, text=True)
    # 0: no device, context creation raised CG_ERR_CUDA; 3: a device is present (GPU box)
    assert r.returncode in (0, 3), r.stdout + r.stderr
    if r.returncode == 0:
        assert "no CPU fallback" in r.stdout


def _read_layer(buf, off):
    (B,) = struct.unpack_from("<I", buf, off)
    off += 4
    idx = np.frombuffer(buf, np.int32, 3 * B, off).reshape(B, 3)
    off += 12 * B
    from coxgraph_b200 import VOXEL_DTYPE
    vox = np.frombuffer(buf, VOXEL_DTYPE, 4096 * B, off).reshape(B, 4096)
    off += 4096 * 12 * B
    return (idx, vox, np.zeros(B, np.uint8)), off


@pytest.mark.gpu
def test_cpp_call_sequence_matches_oracle(tmp_path):
    from coxgraph_b200 import synth
    from oracle import oracle_py as orc
    frames = util.small_frames(3, stride=8)
    T_M_S = np.asarray(synth.robot_map_offset(1), np.float32)
    voxel, trunc = 0.05, 0.15
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(struct.pack("<Iff", len(frames), voxel, trunc))
        f.write(T_M_S.tobytes())
        for (T, p, c) in frames:
            f.write(np.asarray(T, np.float32).tobytes())
            f.write(struct.pack("<I", len(p)))
            f.write(np.ascontiguousarray(p, np.float32).tobytes())
            f.write(np.ascontiguousarray(c, np.uint8).tobytes())
    r = subprocess.run([_binary(), "run", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    submap, off = _read_layer(buf, 0)
    combined, off = _read_layer(buf, off)
    assert off == len(buf)
    ocfg, _ = util.make_cfgs(default_truncation_distance=trunc)   # "fast" -> merged, const weight
    ol, og = orc.Layer(voxel), orc.Layer(voxel)
    for (T, p, c) in frames:
        ol.integrate(ocfg, T, p, c)
    og.merge_from(ol, T_M_S)
    util.compare_layers(submap, ol.download(), "C++ host: submap")
    util.compare_layers(combined, og.download(), "C++ host: combined")
