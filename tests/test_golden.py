"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
the oracle must keep reproducing them bit for bit (CPU), and the CUDA path must reproduce their
block sets exactly and their voxels within tolerance (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "digests.json")) as f:
    META = json.load(f)
CASES = sorted(k for k in META if "voxel_size" in META[k])   # the layer fixtures (mesh_frames: test_mesh_recover.py)


def _load(name):
    return np.load(os.path.join(GOLD, f"{name}.npz"))


def _frames(g):
    offs = g["offsets"].astype(int)
    return [(g["poses"][f], g["points"][offs[f]:offs[f + 1]], g["colors"][offs[f]:offs[f + 1]])
            for f in range(len(g["poses"]))]


def _sha(vox):
    return hashlib.sha256(np.ascontiguousarray(vox).tobytes()).hexdigest()


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    from oracle import oracle_py as orc
    g, meta = _load(name), META[name]
    cfg = orc.default_config(**meta["cfg"])
    L = orc.Layer(meta["voxel_size"])
    for k, (T, p, c) in enumerate(_frames(g)):
        L.integrate(cfg, T, p, c)
        assert L.last_blocks_touched == g["touched"][k]
    idx, vox, flags = L.download()
    assert np.array_equal(idx, g["block_idx"]) and np.array_equal(flags, g["flags"])
    assert _sha(vox) == meta["sha256"]
    pos = g["sample_pos"]
    assert np.array_equal(vox[pos[:, 0], pos[:, 1]], g["sample_vox"])
    if "merge_pose" in g:
        G = orc.Layer(meta["voxel_size"])
        G.merge_from(L, g["merge_pose"])
        gi, gv, gf = G.download()
        assert np.array_equal(gi, g["merge_block_idx"]) and np.array_equal(gf, g["merge_flags"])
        assert _sha(gv) == meta["merge_sha256"] and G.last_blocks_out == meta["merge_blocks_out"]


def _check_sample(vox, pos, want, what):
    got = vox[pos[:, 0], pos[:, 1]]
    for field in ("distance", "weight"):
        bad = np.abs(got[field] - want[field]) > np.maximum(util.ATOL, util.RTOL * np.abs(want[field]))
        assert not bad.any(), f"{what}: {field} out of tolerance at {bad.sum()} sampled voxels"
    assert np.abs(got["rgba"].astype(int) - want["rgba"].astype(int)).max() <= 1, what


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("batch", [False, True])
def test_cuda_reproduces_golden(gpu_ctx, name, batch):
    from coxgraph_b200 import Layer, TsdfIntegrator, TsdfIntegratorConfig, mergeLayerAintoLayerB
    g, meta = _load(name), META[name]
    cfg = TsdfIntegratorConfig(**meta["cfg"])
    L = Layer(gpu_ctx, meta["voxel_size"], max_blocks=4096)
    integ = TsdfIntegrator(cfg, L)
    if batch:
        integ.integrateBatch(g["poses"], g["points"], g["colors"], g["offsets"])
    else:
        for k, (T, p, c) in enumerate(_frames(g)):
            st = integ.integratePointCloud(T, p, c)
            assert st.blocks_touched == g["touched"][k]
    idx, vox, flags = L.download()
    assert np.array_equal(idx, g["block_idx"]), "allocated block set differs from the golden set"
    assert np.array_equal(flags, g["flags"])
    _check_sample(vox, g["sample_pos"], g["sample_vox"], name)
    if "merge_pose" in g:
        G = Layer(gpu_ctx, meta["voxel_size"], max_blocks=4096)
        st = mergeLayerAintoLayerB(L, g["merge_pose"], G)
        gi, gv, gf = G.download()
        assert np.array_equal(gi, g["merge_block_idx"]) and np.array_equal(gf, g["merge_flags"])
        assert st.blocks_out == meta["merge_blocks_out"]
        _check_sample(gv, g["merge_sample_pos"], g["merge_sample_vox"], name + " merge")
        G.close()
    L.close()


def _mesh_digests(begin, v, n, c):
    return dict(num_vertices=int(len(v)), begin_sha256=_sha(begin), vertices_sha256=_sha(v),
                normals_sha256=_sha(n), colors_sha256=_sha(c))


def test_oracle_reproduces_golden_layer_mesh():
    """Marching cubes over the fused merged_5cm layer: frozen digests of the oracle's output."""
    from tests.golden.make_golden import fused_golden_layer
    assert _mesh_digests(*fused_golden_layer().mesh()) == META["layer_mesh"]
    assert META["layer_mesh"]["num_vertices"] > 1000


@pytest.mark.gpu
def test_cuda_reproduces_golden_layer_mesh(gpu_ctx):
    from coxgraph_b200 import Layer
    from tests.golden.make_golden import fused_golden_layer
    idx, vox, flags = fused_golden_layer().download()
    L = Layer(gpu_ctx, 0.05, max_blocks=4096)
    L.upload(idx, vox, flags)
    gi, begin, v, n, c = L.generateMesh()
    assert np.array_equal(gi, idx)
    assert _mesh_digests(begin, v, n, c) == META["layer_mesh"]
    L.close()
