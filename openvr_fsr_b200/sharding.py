"""Multi-GPU partitioning of the path: eyes and frames are independent units, so the
only communication is a broadcast of the <=256-byte constant blocks from rank 0 and barriers around timed
regions.  One process per GPU over torch.distributed (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

from . import api


def eye_for_rank(rank: int, world: int):
    """2-GPU stereo mode: one eye per device (eye index = EVREye).  Returns the eyes this rank owns."""
    if world == 1:
        return [0, 1]
    if world == 2:
        return [rank]
    raise ValueError("eye sharding is the 2-GPU mode; use frames_for_rank beyond 2 GPUs")


def frames_for_rank(n_frames: int, rank: int, world: int):
    """Batched offline mode: frame f (both eyes) -> GPU f mod world."""
    return list(range(rank, n_frames, world))


def pack_constants(cfg: api.Config, in_w: int, in_h: int, out_w: int, out_h: int, only_one_eye: bool = True):
    """All constant blocks of one configuration as a single uint32 vector:
    [upscale eye0 (24) | upscale eye1 (24) | sharpen eye0 (12) | sharpen eye1 (12)]."""
    words = []
    for eye in (0, 1):
        words.append(api.make_upscale_constants(cfg, eye, only_one_eye, in_w, in_h, out_w, out_h))
    for eye in (0, 1):
        words.append(api.make_sharpen_constants(cfg, eye, only_one_eye, out_w, out_h))
    return np.concatenate(words).astype(np.uint32)


def unpack_constants(vec: np.ndarray):
    vec = np.asarray(vec, dtype=np.uint32)
    return {"upscale": [vec[0:24].copy(), vec[24:48].copy()], "sharpen": [vec[48:60].copy(), vec[60:72].copy()]}


def broadcast_constants(cfg: api.Config, in_w: int, in_h: int, out_w: int, out_h: int, device=None,
                        only_one_eye: bool = True):
    """Rank 0 builds the constant blocks, every rank receives root's copy (ncclBroadcast of 288 bytes).
    Without an initialised process group this is the local computation."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return unpack_constants(pack_constants(cfg, in_w, in_h, out_w, out_h, only_one_eye))
    if dist.get_rank() == 0:
        t = torch.from_numpy(pack_constants(cfg, in_w, in_h, out_w, out_h, only_one_eye).view(np.int32).copy())
    else:
        t = torch.zeros(72, dtype=torch.int32)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return unpack_constants(t.cpu().numpy().view(np.uint32))
