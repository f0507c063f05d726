#!/usr/bin/env python
"""Multi-GPU parity of the sharded global merge, one process per GPU (torchrun, NCCL):
every rank fuses the submaps assigned to it with the CUDA library, projects them into a partial
global layer, and the partial layers are exchanged over NCCL (coxgraph_b200.sharding).  The union
of the owned layers is compared with the CPU oracle's statement of the same plan.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 scripts/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from coxgraph_b200 import (Context, Layer, TsdfIntegrator, getProjectedMap, sharding,  # noqa: E402
                           synth)
from tests import util  # noqa: E402

ROBOTS, SUBMAPS_PER_ROBOT, FRAMES, STRIDE = 4, 2, 2, 8


def frames_of(sid):
    robot, sm = divmod(sid, SUBMAPS_PER_ROBOT)
    return util.small_frames(FRAMES, stride=STRIDE, robot=robot, submap=sm), \
        synth.robot_map_offset(robot)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ocfg, gcfg = util.make_cfgs()
    ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
    mine = sharding.assign_submaps(ROBOTS, SUBMAPS_PER_ROBOT, world)[rank]
    layers, poses = [], []
    for sid in mine:
        fr, T_M_S = frames_of(sid)
        L = Layer(ctx, 0.05, max_blocks=2048)
        integ = TsdfIntegrator(gcfg, L)
        for (T, p, c) in fr:
            integ.integratePointCloud(T, p, c)
        layers.append(L)
        poses.append(T_M_S)
    partial, owned = Layer(ctx, 0.05, max_blocks=8192), Layer(ctx, 0.05, max_blocks=8192)
    sent, got = sharding.project_sharded(layers, np.stack(poses) if poses else np.zeros((0, 7)),
                                         partial, owned)
    oi, ov, of = owned.download()
    assert (sharding.block_owners(oi, world) == rank).all(), "a rank holds a block it does not own"
    # the same merge through the C ABI alone (csrc/comm.cu: holder-aware ownership, owner-pull over
    # NVLink peer memory, NCCL as bootstrap and barrier), twice in a row (the second call reuses
    # the peer mappings and the cleared partial layer).  Ownership differs from the packed path
    # (a block stays with one of the ranks that hold it), so the UNION of the owned layers is
    # compared below: it must be bit-identical.
    native, native_err = None, ""
    try:
        sharding.init_native(ctx)
        owned2 = Layer(ctx, 0.05, max_blocks=8192)
        for rep in range(2):
            owned2.clear()
            sharding.project_sharded_native(layers, np.stack(poses) if poses else np.zeros((0, 7)),
                                            partial, owned2)
            native = owned2.download()
        owned2.close()
    except Exception as e:  # noqa: BLE001 - reported through the gathered results below
        native, native_err = None, repr(e)
    gathered = [None] * world
    dist.all_gather_object(gathered, (oi, ov, of, sent, got, native, native_err))
    ok = True
    if rank == 0:
        from oracle import oracle_py as orc
        expect = orc.Layer(0.05)
        for r in range(world):                      # ascending source rank = the fold order
            part = orc.Layer(0.05)
            for sid in sharding.assign_submaps(ROBOTS, SUBMAPS_PER_ROBOT, world)[r]:
                fr, T_M_S = frames_of(sid)
                L = orc.Layer(0.05)
                for (T, p, c) in fr:
                    L.integrate(ocfg, T, p, c)
                part.merge_from(L, T_M_S)
            expect.merge_aligned_from(part)
        gi = np.concatenate([g[0] for g in gathered])
        gv = np.concatenate([g[1] for g in gathered])
        gf = np.concatenate([g[2] for g in gathered])
        order = np.lexsort((gi[:, 0], gi[:, 1], gi[:, 2]))
        try:
            util.compare_layers((gi[order], gv[order], gf[order]), expect.download(),
                                f"{world}-GPU sharded merge")
            assert sum(sum(g[3]) for g in gathered) == sum(sum(g[4]) for g in gathered)
            assert all(g[5] is not None for g in gathered), \
                f"native (C ABI) exchange failed: {[g[6] for g in gathered]}"
            ni = np.concatenate([g[5][0] for g in gathered])
            nv = np.concatenate([g[5][1] for g in gathered])
            no = np.lexsort((ni[:, 0], ni[:, 1], ni[:, 2]))
            assert len(np.unique(ni, axis=0)) == len(ni), "a block is owned by two ranks"
            util.compare_layers((ni[no], nv[no], None), (gi[order], gv[order], None),
                                "native vs packed exchange (union of the owned layers)", exact=True)
            # against the reference's single left fold over all submaps (rank-major
            # order), where the sharded plan re-associates: margins, not only pass / fail
            single = orc.Layer(0.05)
            for r in range(world):
                for sid in sharding.assign_submaps(ROBOTS, SUBMAPS_PER_ROBOT, world)[r]:
                    fr, T_M_S = frames_of(sid)
                    L = orc.Layer(0.05)
                    for (T, p, c) in fr:
                        L.integrate(ocfg, T, p, c)
                    single.merge_from(L, T_M_S)
            m = util.margins((gi[order], gv[order], gf[order]), single.download())
            print(f"margins vs the single-process left fold: {m}", flush=True)
            util.compare_layers((gi[order], gv[order], gf[order]), single.download(),
                                f"{world}-GPU sharded merge vs single left fold")
            print(f"multigpu_check ok: world {world}, {len(gi)} global blocks, "
                  f"records exchanged {sum(sum(g[3]) for g in gathered)}, native exchange identical",
                  flush=True)
        except AssertionError as e:
            print("multigpu_check FAILED:", e, flush=True)
            ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
