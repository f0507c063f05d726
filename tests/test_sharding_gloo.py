"""N>1 host logic of the global merge on CPU: 2 ranks over gloo.  The GPU kernels are stood in by
numpy record packing and the oracle's aligned merge; what is under test is the plan itself
(submap assignment, block ownership, the uneven all-to-all, the fold order) — the same
coxgraph_b200.sharding code the NCCL path runs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coxgraph_b200 import sharding, synth
from tests import util

WORLD = 2
ROBOTS, SUBMAPS_PER_ROBOT = 2, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _submap_layer(orc, ocfg, sid):
    robot, sm = divmod(sid, SUBMAPS_PER_ROBOT)
    L = orc.Layer(0.05)
    for (T, p, c) in util.small_frames(1, stride=16, robot=robot, submap=sm):
        L.integrate(ocfg, T, p, c)
    return L, synth.robot_map_offset(robot)


def pack_records(idx, vox, flags, world):
    """numpy twin of cg_layer_pack_by_owner: records grouped by owner, (z,y,x) order inside."""
    owners = sharding.block_owners(idx, world)
    order = np.argsort(owners, kind="stable")
    rec = np.zeros((len(idx), sharding.RECORD_BYTES), np.uint8)
    for row, b in enumerate(order):
        hdr = np.array([idx[b][0], idx[b][1], idx[b][2], int(flags[b])], np.int32)
        rec[row, :16] = hdr.view(np.uint8)
        planes = np.concatenate([vox[b]["distance"].view(np.uint32), vox[b]["weight"].view(np.uint32),
                                 vox[b]["rgba"].copy().view(np.uint32).reshape(-1)])
        rec[row, 16:] = planes.view(np.uint8)
    counts = [int((owners == r).sum()) for r in range(world)]
    return rec, counts


def unpack_records(rec):
    from oracle import oracle_py as orc
    n = len(rec)
    hdr = rec[:, :16].copy().view(np.int32).reshape(n, 4)
    planes = rec[:, 16:].copy().view(np.uint32).reshape(n, 3, 4096)
    vox = np.zeros((n, 4096), orc.VOXEL_DTYPE)
    vox["distance"] = planes[:, 0].view(np.float32)
    vox["weight"] = planes[:, 1].view(np.float32)
    vox["rgba"] = planes[:, 2].copy().view(np.uint8).reshape(n, 4096, 4)
    return hdr[:, :3].copy(), vox, hdr[:, 3].astype(np.uint8)


def _worker(rank, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    from oracle import oracle_py as orc
    ocfg, _ = util.make_cfgs()
    mine = sharding.assign_submaps(ROBOTS, SUBMAPS_PER_ROBOT, WORLD)[rank]
    partial = orc.Layer(0.05)
    for sid in mine:
        L, T = _submap_layer(orc, ocfg, sid)
        partial.merge_from(L, T)
    idx, vox, flags = partial.download()
    rec, counts = pack_records(idx, vox, flags, WORLD)
    recv, recv_counts = sharding.exchange_records(torch.from_numpy(rec), counts)
    assert recv.shape[0] == sum(recv_counts)
    owned = orc.Layer(0.05)
    off = 0
    for src in range(WORLD):                 # ascending source rank = the fold order
        ridx, rvox, rfl = unpack_records(recv[off:off + recv_counts[src]].numpy())
        off += recv_counts[src]
        tmp = orc.Layer(0.05)
        tmp.upload(ridx, rvox, rfl)
        owned.merge_aligned_from(tmp)
    oi, ov, of = owned.download()
    assert (sharding.block_owners(oi, WORLD) == rank).all()
    np.savez(os.path.join(out_dir, f"owned{rank}.npz"), idx=oi, vox=ov, flags=of,
             sent=np.array(counts), got=np.array(recv_counts))
    dist.barrier()
    dist.destroy_process_group()


def test_assignment_keeps_robots_together():
    a = sharding.assign_submaps(8, 64, 8)
    assert all(len(x) == 64 for x in a) and a[3] == list(range(192, 256))
    a = sharding.assign_submaps(2, 20, 1)
    assert a == [list(range(40))]
    a = sharding.assign_submaps(3, 2, 2)
    assert a == [[0, 1, 4, 5], [2, 3]]
    assert sharding.assign_robots(8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]


def test_two_rank_global_merge_matches_single_process_fold(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    from oracle import oracle_py as orc
    ocfg, _ = util.make_cfgs()
    # single-process statement of the sharded semantics: partial layer per rank, folded in
    # ascending rank order with the aligned merge
    expect = orc.Layer(0.05)
    for rank in range(WORLD):
        partial = orc.Layer(0.05)
        for sid in sharding.assign_submaps(ROBOTS, SUBMAPS_PER_ROBOT, WORLD)[rank]:
            L, T = _submap_layer(orc, ocfg, sid)
            partial.merge_from(L, T)
        expect.merge_aligned_from(partial)
    ei, ev, ef = expect.download()
    parts = [np.load(os.path.join(tmp_path, f"owned{r}.npz")) for r in range(WORLD)]
    assert sum(len(p["idx"]) for p in parts) == len(ei)          # a partition: no block twice
    gi = np.concatenate([p["idx"] for p in parts])
    gv = np.concatenate([p["vox"] for p in parts])
    gf = np.concatenate([p["flags"] for p in parts])
    order = np.lexsort((gi[:, 0], gi[:, 1], gi[:, 2]))
    util.compare_layers((gi[order], gv[order], gf[order]), (ei, ev, ef), "2-rank fold", exact=True,
                        check_flags=True)
    assert parts[0]["sent"].sum() + parts[1]["sent"].sum() == \
        parts[0]["got"].sum() + parts[1]["got"].sum()
    # and it agrees with the plain sequential getProjectedMap within the merge's rounding
    seq = orc.Layer(0.05)
    for sid in range(ROBOTS * SUBMAPS_PER_ROBOT):
        L, T = _submap_layer(orc, ocfg, sid)
        seq.merge_from(L, T)
    si, sv, _ = seq.download()
    assert np.array_equal(si, ei)
    assert np.allclose(sv["weight"], ev["weight"], rtol=1e-5, atol=1e-6)
    assert np.allclose(sv["distance"], ev["distance"], rtol=1e-4, atol=1e-5)
    assert np.abs(sv["rgba"].astype(int) - ev["rgba"].astype(int)).max() <= 2
