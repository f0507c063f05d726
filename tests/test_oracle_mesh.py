"""CPU: the oracle's restatement of voxblox MeshIntegrator / MarchingCubes.

The reference tree holds no mesh tests or fixtures (the code is in the un-vendored voxblox fork),
so the restatement is pinned by what marching cubes must satisfy whatever the implementation:
the triangle table's consistency with the corner signs, watertightness and orientation on a
closed surface, exact answers on a plane, plus the reference's own validity rules (min_weight,
missing neighbour blocks)."""
import numpy as np

from oracle import oracle_py as orc

EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6),
         (3, 7)]
VS = 0.05


def dense_layer(sdf_fn, blocks, weight=1.0, color_fn=None):
    """Layer with every voxel of the given blocks observed; distance = sdf_fn(centre)."""
    L = orc.Layer(VS)
    idx = np.array(blocks, np.int32)
    vox = np.zeros((len(idx), 4096), orc.VOXEL_DTYPE)
    lin = np.arange(4096)
    loc = np.stack([lin & 15, (lin >> 4) & 15, lin >> 8], -1)
    for b, bi in enumerate(idx):
        c = (bi[None, :] * 16 + loc + 0.5) * VS
        vox[b]["distance"] = sdf_fn(c)
        vox[b]["weight"] = weight
        vox[b]["rgba"] = color_fn(c) if color_fn else np.array([10, 20, 30, 255], np.uint8)
    L.upload(idx, vox)
    return L


def test_triangle_table_matches_corner_signs():
    T = orc.triangle_table()
    for cfg in range(256):
        inside = [(cfg >> i) & 1 for i in range(8)]
        crossing = {e for e, (a, b) in enumerate(EDGES) if inside[a] != inside[b]}
        row = list(T[cfg])
        k = row.index(-1)
        assert k % 3 == 0 and k <= 15 and all(x == -1 for x in row[k:])
        assert set(row[:k]) == crossing, f"configuration {cfg}"
        for t in range(0, k, 3):
            assert len(set(row[t:t + 3])) == 3
    # complementary configurations cut the same edges
    for cfg in range(256):
        assert set(T[cfg][T[cfg] >= 0]) == set(T[255 - cfg][T[255 - cfg] >= 0])


def test_sphere_mesh_is_closed_and_oriented():
    blocks = [(x, y, z) for z in (-1, 0) for y in (-1, 0) for x in (-1, 0)]
    ctr = np.array([0.013, -0.021, 0.007])
    L = dense_layer(lambda c: np.linalg.norm(c - ctr, axis=-1) - 0.5, blocks)
    begin, v, n, col = L.mesh()
    assert len(v) > 3000 and len(v) % 3 == 0 and begin[-1] == len(v)
    # vertices lie on the sphere up to the linear interpolation error, normals point outwards
    # (towards positive distance)
    r = np.linalg.norm(v - ctr, axis=1)
    assert np.abs(r - 0.5).max() < 2e-3
    tri = v.reshape(-1, 3, 3)
    cen = tri.mean(1)
    nn = n.reshape(-1, 3, 3)
    assert np.array_equal(nn[:, 0], nn[:, 1]) and np.array_equal(nn[:, 0], nn[:, 2])
    area2 = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    good = area2 > 1e-9
    assert np.abs(np.linalg.norm(nn[good, 0], axis=1) - 1).max() < 1e-5
    outward = np.einsum("ij,ij->i", nn[good, 0], cen[good] - ctr)
    assert (outward > 0).all() or (outward < 0).all()
    # watertight: after welding, every directed edge has exactly one opposite partner
    key = np.round(v / 1e-5).astype(np.int64)
    _, ids = np.unique(key, axis=0, return_inverse=True)
    t = ids.reshape(-1, 3)
    t = t[(t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])]
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    fwd = set(map(tuple, e))
    assert len(fwd) == len(e), "a directed edge is used twice"
    assert all((b, a) in fwd for (a, b) in fwd), "open edge: the mesh is not watertight"
    # surface area of the sphere within 1 %
    assert abs(0.5 * area2.sum() - 4 * np.pi * 0.25) < 0.01 * 4 * np.pi * 0.25


def test_plane_vertices_normals_and_colours():
    z0 = 0.26
    colf = lambda c: np.stack([np.floor(c[:, 0] / VS) % 200, np.floor(c[:, 1] / VS) % 200,
                               np.floor(c[:, 2] / VS) % 200, np.full(len(c), 255)], -1).astype(np.uint8)
    L = dense_layer(lambda c: c[:, 2] - z0, [(0, 0, 0), (1, 0, 0)], color_fn=colf)
    begin, v, n, col = L.mesh()
    assert np.abs(v[:, 2] - z0).max() < 1e-6
    assert np.abs(np.abs(n[:, 2]) - 1).max() < 1e-6 and (np.sign(n[:, 2]) == np.sign(n[0, 2])).all()
    tri = v.reshape(-1, 3, 3)
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1).sum()
    # block 0 reaches into block 1 through its max-X border; block 1 has no +x neighbour
    assert abs(area - (31 * VS) * (15 * VS)) < 1e-4
    assert begin[0] == 0 and 0 < begin[1] < begin[2] == len(v)
    # colour = the voxel holding the vertex
    vi = np.floor(v / VS + 1e-6).astype(int)
    want = np.stack([vi[:, 0] % 200, vi[:, 1] % 200, vi[:, 2] % 200, np.full(len(v), 255)], -1)
    assert np.array_equal(col, want.astype(np.uint8))


def test_unobserved_corners_and_missing_neighbours_make_no_triangles():
    L = dense_layer(lambda c: c[:, 2] - 0.26, [(0, 0, 0)], weight=1e-4)  # weight <= min_weight
    assert len(L.mesh()[1]) == 0
    assert len(L.mesh(min_weight=1e-5)[1]) > 0
    # only_updated: blocks uploaded without the flag are skipped
    assert len(L.mesh(min_weight=1e-5, only_updated=True)[1]) == 0
    # a single block: only cubes entirely inside it are meshed
    tri = L.mesh(min_weight=1e-5)[1].reshape(-1, 3, 3)
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1).sum()
    assert abs(area - (15 * VS) ** 2) < 1e-4


def test_connected_mesh_merges_exactly_coincident_vertices():
    """createConnectedMesh (MeshLayer::getConnectedMesh, the PLY step of saveAndPubCombinedMesh,
    server_visualizer.cpp:118-126) welds the per-block triangle soup with a 1e-10 m cell: only
    vertices whose coordinates agree to the bit (or within 1e-10 m near the origin) merge — an
    edge vertex interpolated from its two corners in the other order may differ in the last bit and
    stays separate, upstream too.  Unique vertices come in order of first occurrence."""
    blocks = [(x, y, z) for z in (-1, 0) for y in (-1, 0) for x in (-1, 0)]
    ctr = np.array([0.013, -0.021, 0.007])
    L = dense_layer(lambda c: np.linalg.norm(c - ctr, axis=-1) - 0.5, blocks)
    begin, v, n, c = L.mesh()
    idx, first = orc.connect_mesh(v)
    assert len(idx) == len(v) and len(first) < len(v) / 2
    assert np.array_equal(v[first][idx], v)  # every corner maps to a vertex with its own position
    assert (np.diff(first.astype(np.int64)) > 0).all()  # order of first occurrence
    assert np.array_equal(idx[first], np.arange(len(first)))
    # distinct unique vertices have distinct positions
    assert len(np.unique(v[first], axis=0)) == len(first)
    # hand case: -0.0 joins +0.0, 1e-12 m falls into the same 1e-10 m cell, a real neighbour does not
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [-0.0, 0, 0], [0, 0, 1e-12], [0, 0, 1e-6]], np.float32)
    idx, first = orc.connect_mesh(pts)
    assert idx.tolist() == [0, 1, 2, 0, 0, 3] and first.tolist() == [0, 1, 2, 5]
