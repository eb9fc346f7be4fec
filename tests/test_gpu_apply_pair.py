"""ovrfsr_apply_pair: both eyes of a frame in one call, the right eye's passes forked onto a ctx-owned stream and joined
back.  It must be indistinguishable from apply(left) ; apply(right) on the caller's stream -- same bits, same per-eye
state, same ordering guarantees for whatever the caller queues next -- on the legacy stream, a side stream, across
size changes and when captured into a CUDA graph."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [
    dict(renderScale=0.75, sharpness=0.9, radius=2.0),
    dict(renderScale=0.75, sharpness=0.9, radius=0.5, projCentre=(.45, .5, .55, .5)),
    dict(renderScale=0.59, sharpness=0.6, radius=0.4, useNis=True, projCentre=(.45, .5, .55, .5)),
    dict(renderScale=1.0, sharpness=0.7, radius=0.5, useNis=True),
    dict(renderScale=0.77, sharpness=0.8, radius=0.3, debugMode=True),
]


def _frames(torch, synth, dev, w, h, n, seed):
    return [(torch.from_numpy(synth.natural_rgba8(w, h, seed + 2 * i)).to(dev), torch.from_numpy(synth.natural_rgba8(w, h, seed + 2 * i + 1)).to(dev))
            for i in range(n)]


@pytest.mark.parametrize("kw", CONFIGS, ids=lambda k: ("nis" if k.get("useNis") else "fsr") + f"-s{k['renderScale']}-r{k['radius']}")
@pytest.mark.parametrize("side_stream", [False, True])
def test_pair_matches_two_applies(cuda, kw, side_stream):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    cfg = ovr.Config(fsrEnabled=True, **kw)
    pair, plain = ovr.PostProcessor(cfg), ovr.PostProcessor(cfg)
    stream = torch.cuda.Stream(device=cuda) if side_stream else None
    frames = _frames(torch, synth, cuda, 403, 287, 4, 11)
    torch.cuda.synchronize()
    got, want = [], []
    with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream(cuda)):
        for (l, r) in frames:  # no synchronisation between frames: the clones are ordered by the caller's stream alone
            a, b = pair.apply_pair(l, r, stream=stream)
            got.append((a.clone(), b.clone()))
        for (l, r) in frames:
            want.append((plain.apply(0, l, stream=stream).clone(), plain.apply(1, r, stream=stream).clone()))
    torch.cuda.synchronize()
    for (ga, gb), (wa, wb) in zip(got, want):
        assert torch.equal(ga, wa) and torch.equal(gb, wb)
    assert not torch.equal(got[0][0], got[0][1])  # the eyes really differ (content and projection centre)
    for e in (0, 1):
        assert np.array_equal(pair.upscale_constants(e), plain.upscale_constants(e))
    pair.close(), plain.close()


def test_pair_survives_size_change_and_reset(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.6)
    pair, plain = ovr.PostProcessor(cfg), ovr.PostProcessor(cfg)
    for (w, h) in ((320, 200), (200, 260), (320, 200)):
        for (l, r) in _frames(torch, synth, cuda, w, h, 2, w):
            a, b = pair.apply_pair(l, r)
            assert torch.equal(a, plain.apply(0, l)) and torch.equal(b, plain.apply(1, r))
    pair.reset()
    l, r = _frames(torch, synth, cuda, 320, 200, 1, 5)[0]
    a, b = pair.apply_pair(l, r)
    assert torch.equal(a, plain.apply(0, l)) and torch.equal(b, plain.apply(1, r))
    pair.close(), plain.close()


def test_pair_with_shared_texture_is_the_two_calls(cuda):
    """One texture holding both eyes (uMax - uMin == 0.5): processed once, both outputs are the same image."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    tex = torch.from_numpy(synth.natural_rgba8(320, 120, 21)).to(cuda)
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.6, projCentre=(.45, .5, .55, .5))
    pair, plain = ovr.PostProcessor(cfg), ovr.PostProcessor(cfg)
    n0 = ovr.kernel_launches()
    a, b = pair.apply_pair(tex, tex, ovr.TextureBounds(0.0, 0.0, 0.5, 1.0))
    assert ovr.kernel_launches() - n0 == 2 and a.data_ptr() == b.data_ptr()
    want = plain.apply(0, tex, ovr.TextureBounds(0.0, 0.0, 0.5, 1.0))
    assert torch.equal(a, want)
    pair.close(), plain.close()


def test_pair_disabled_passes_through(cuda):
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    l, r = _frames(torch, synth, cuda, 64, 48, 1, 1)[0]
    pp = ovr.PostProcessor(ovr.Config(fsrEnabled=False))
    a, b = pp.apply_pair(l, r)
    assert a is l and b is r
    pp.close()


def test_pair_captures_into_a_cuda_graph(cuda):
    """The fork/join is event-based, so a stream capture follows it onto the ctx-owned stream: one graph replays both
    eyes' launches."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    cfg = ovr.Config(fsrEnabled=True, renderScale=0.75, sharpness=0.9, radius=0.5)
    pair, plain = ovr.PostProcessor(cfg), ovr.PostProcessor(cfg)
    l, r = _frames(torch, synth, cuda, 403, 287, 1, 31)[0]
    a, b = pair.apply_pair(l, r)  # lazy init outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream(cuda)
        a, b = pair.apply_pair(l, r, stream=s)
    # new content in the same input buffers, then replay
    l2, r2 = _frames(torch, synth, cuda, 403, 287, 1, 77)[0]
    l.copy_(l2), r.copy_(r2)
    a.zero_(), b.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(a, plain.apply(0, l2)) and torch.equal(b, plain.apply(1, r2))
    del g
    pair.close(), plain.close()
