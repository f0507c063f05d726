#!/usr/bin/env python
"""Generate the golden fixtures of the TSDF hot path from the CPU oracle.

The reference holds no golden vectors for this path ("parity unpinned"), so
these fixtures freeze the ORACLE's behaviour: seeded inputs (stored, so they do not depend on
the torch build that rendered them) and the oracle's outputs (block index set, SHA-256 of the
voxel bytes, a sample of voxel records).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from coxgraph_b200 import synth  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "merged_5cm": dict(cfg=dict(default_truncation_distance=0.15, max_ray_length_m=5.0,
                                min_ray_length_m=0.1, use_const_weight=1, method=1),
                       voxel_size=0.05, robot=0, frames=2, stride=16, cam="640"),
    "simple_5cm_invz": dict(cfg=dict(default_truncation_distance=0.16, max_ray_length_m=5.0,
                                     min_ray_length_m=0.1, use_const_weight=0, method=0),
                            voxel_size=0.05, robot=1, frames=1, stride=16, cam="640"),
    "merged_2cm_720p": dict(cfg=dict(default_truncation_distance=0.06, max_ray_length_m=3.0,
                                     min_ray_length_m=0.1, use_const_weight=1, method=1),
                            voxel_size=0.02, robot=2, frames=1, stride=16, cam="1280"),
}
MERGE_POSE = synth.robot_map_offset(1)


def digest(vox):
    return hashlib.sha256(np.ascontiguousarray(vox).tobytes()).hexdigest()


def sample_voxels(idx, vox, n=3000, seed=7):
    rng = np.random.default_rng(seed)
    obs = np.argwhere(vox["weight"] > 0)
    pick = obs[rng.choice(len(obs), size=min(n, len(obs)), replace=False)]
    return pick.astype(np.int32), vox[pick[:, 0], pick[:, 1]]


def build_case(name, spec):
    cam = synth.CAM_640x480 if spec["cam"] == "640" else synth.CAM_1280x720
    frames = synth.submap_frames(spec["robot"], 0, spec["frames"], cam=cam, stride=spec["stride"])
    poses = np.stack([T for (T, _, _) in frames]).astype(np.float32)
    pts = [p.numpy() for (_, p, _) in frames]
    cols = [c.numpy() for (_, _, c) in frames]
    cfg = orc.default_config(**spec["cfg"])
    L = orc.Layer(spec["voxel_size"])
    touched = []
    for T, p, c in zip(poses, pts, cols):
        L.integrate(cfg, T, p, c)
        touched.append(L.last_blocks_touched)
    idx, vox, flags = L.download()
    pick, vals = sample_voxels(idx, vox)
    out = dict(poses=poses, points=np.concatenate(pts), colors=np.concatenate(cols),
               offsets=np.cumsum([0] + [len(p) for p in pts]).astype(np.uint64),
               block_idx=idx, flags=flags, sample_pos=pick, sample_vox=vals,
               touched=np.array(touched, np.int64))
    meta = dict(cfg=spec["cfg"], voxel_size=spec["voxel_size"], sha256=digest(vox),
                num_blocks=int(len(idx)))
    if name == "merged_5cm":  # the fused layer is also the source of the merge fixture
        G = orc.Layer(spec["voxel_size"])
        G.merge_from(L, MERGE_POSE)
        gi, gv, gf = G.download()
        gp, gs = sample_voxels(gi, gv, seed=11)
        out.update(merge_pose=MERGE_POSE, merge_block_idx=gi, merge_flags=gf, merge_sample_pos=gp,
                   merge_sample_vox=gs)
        meta.update(merge_sha256=digest(gv), merge_num_blocks=int(len(gi)),
                    merge_blocks_out=int(G.last_blocks_out))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    return meta


def build_mesh_case():
    """MeshConverter fixture: the synthetic mesh / trajectory of tests/test_mesh_recover.py (seeded
    numpy generators) and the digest of the per-pose clouds the oracle recovers from it."""
    from tests.test_mesh_recover import make_mesh, make_trajectory
    mesh = make_mesh(seed=0, blocks=9, tris_per_block=60)
    poses, stamps = make_trajectory(12, 0)
    offs, pts, cols = orc.mesh_to_frames(mesh, 0.05, poses, stamps)
    return dict(num_points=int(len(pts)), offsets_sha256=digest(offs), points_sha256=digest(pts),
                colors_sha256=digest(cols))


def fused_golden_layer(name="merged_5cm"):
    """The oracle layer of a layer fixture, re-fused from the stored inputs."""
    g = np.load(os.path.join(HERE, f"{name}.npz"))
    spec = CASES[name]
    cfg = orc.default_config(**spec["cfg"])
    L = orc.Layer(spec["voxel_size"])
    offs = g["offsets"].astype(int)
    for f in range(len(g["poses"])):
        L.integrate(cfg, g["poses"][f], g["points"][offs[f]:offs[f + 1]], g["colors"][offs[f]:offs[f + 1]])
    return L


def build_layer_mesh_case():
    """MeshIntegrator fixture: marching cubes over the fused merged_5cm layer (oracle)."""
    begin, v, n, c = fused_golden_layer().mesh()
    return dict(num_vertices=int(len(v)), begin_sha256=digest(begin), vertices_sha256=digest(v),
                normals_sha256=digest(n), colors_sha256=digest(c))


if __name__ == "__main__":
    if "--only-layer-mesh" in sys.argv:  # add the newest fixture without touching the others
        with open(os.path.join(HERE, "digests.json")) as f:
            metas = json.load(f)
        metas["layer_mesh"] = build_layer_mesh_case()
        with open(os.path.join(HERE, "digests.json"), "w") as f:
            json.dump(metas, f, indent=1, sort_keys=True)
        print("layer_mesh", metas["layer_mesh"])
        sys.exit(0)
    metas = {name: build_case(name, spec) for name, spec in CASES.items()}
    metas["mesh_frames"] = build_mesh_case()
    metas["layer_mesh"] = build_layer_mesh_case()
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(metas, f, indent=1, sort_keys=True)
    for k, v in metas.items():
        print(k, v.get("num_blocks", v.get("num_points")), v.get("sha256", v.get("points_sha256"))[:16])
