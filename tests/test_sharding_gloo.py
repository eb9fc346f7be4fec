"""The N>1 host logic on CPU: two gloo ranks agree on the broadcast constant blocks and split frames/eyes
without overlap (independent units, no data-path collective)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 1 deliberately holds a different radius: after the broadcast both must use root's constants
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5 if rank == 0 else 0.1)
    got = sharding.broadcast_constants(cfg, 1683, 1869, 2244, 2492)
    dist.barrier()
    q.put((rank, np.concatenate(got["upscale"] + got["sharpen"]).tolist(), sharding.frames_for_rank(10, rank, world),
           sharding.eye_for_rank(rank, world)))
    dist.destroy_process_group()


def test_two_ranks_share_constants_and_partition_work():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, c0, f0, e0), (_, c1, f1, e1) = res
    assert c0 == c1
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import sharding
    root = sharding.pack_constants(ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5), 1683, 1869, 2244, 2492)
    assert c0 == root.tolist()
    assert sorted(f0 + f1) == list(range(10)) and not set(f0) & set(f1)
    assert e0 == [0] and e1 == [1]


def test_pack_roundtrip_single_process():
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import sharding
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=2.0)
    d = sharding.broadcast_constants(cfg, 1683, 1869, 2244, 2492)
    assert np.array_equal(d["upscale"][0], ovr.make_upscale_constants(cfg, 0, True, 1683, 1869, 2244, 2492))
    assert np.array_equal(d["sharpen"][1], ovr.make_sharpen_constants(cfg, 1, True, 2244, 2492))
    assert sharding.frames_for_rank(256, 3, 8) == list(range(3, 256, 8))
