"""GPU parity, proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

strict math mode -> bit-identical to the oracle (which is itself bit-identical to the reference's own lines
compiled on the host, tests/test_oracle_vs_ref.py); fast mode -> <= 1 LSB per RGBA8 channel (north_star).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [  # (inW, inH, renderScale)
    (17, 13, 0.75), (16, 16, 0.5), (33, 47, 0.75), (200, 150, 0.77), (129, 65, 0.59), (64, 64, 1.3), (301, 97, 0.67),
]


def _inputs(w, h):
    from openvr_fsr_b200 import synth
    yield "uniform", synth.uniform_rgba8(w, h, 0)
    yield "natural", synth.natural_rgba8(w, h, 1)


def _run_gpu_easu(cuda, src_np, ow, oh, consts, mode, out_dtype=np.uint8):
    import torch
    import openvr_fsr_b200 as ovr
    src = torch.from_numpy(src_np).to(cuda)
    dst = torch.zeros((oh, ow, 4), dtype=torch.uint8 if out_dtype == np.uint8 else torch.float16, device=cuda)
    ovr.fsr_easu(src, dst, consts, mode)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


def _run_gpu_rcas(cuda, src_np, consts, mode, out_dtype=np.uint8):
    import torch
    import openvr_fsr_b200 as ovr
    src = torch.from_numpy(src_np).to(cuda)
    dst = torch.zeros(src_np.shape, dtype=torch.uint8 if out_dtype == np.uint8 else torch.float16, device=cuda)
    ovr.fsr_rcas(src, dst, consts, mode)
    torch.cuda.synchronize()
    return dst.cpu().numpy()


@pytest.mark.parametrize("iw,ih,scale", SIZES)
@pytest.mark.parametrize("radius", [2.0, 0.5])
def test_easu_rcas_vs_oracle(cuda, iw, ih, scale, radius):
    import openvr_fsr_b200 as ovr
    from oracle import pyoracle as po
    ow, oh = po.output_size(iw, ih, scale)
    assert (ow, oh) == ovr.output_size(iw, ih, scale)
    uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
    sc = po.sharpen_constants(0, True, ow, oh, radius=radius, sharpness=0.9)
    for name, src in _inputs(iw, ih):
        ref_e = po.easu(src, ow, oh, uc)
        ref_r = po.rcas(ref_e, sc)
        # strict: bit-identical
        got_e = _run_gpu_easu(cuda, src, ow, oh, uc.words(), ovr.MATH_STRICT)
        assert np.array_equal(got_e, ref_e), f"EASU strict mismatch on {name}: {(got_e != ref_e).sum()} bytes"
        got_r = _run_gpu_rcas(cuda, ref_e, sc.words(), ovr.MATH_STRICT)
        assert np.array_equal(got_r, ref_r), f"RCAS strict mismatch on {name}: {(got_r != ref_r).sum()} bytes"
        # fast: <= 1 LSB per channel
        fast_e = _run_gpu_easu(cuda, src, ow, oh, uc.words(), ovr.MATH_FAST)
        assert np.abs(fast_e.astype(np.int16) - ref_e.astype(np.int16)).max() <= 1
        fast_r = _run_gpu_rcas(cuda, ref_e, sc.words(), ovr.MATH_FAST)
        assert np.abs(fast_r.astype(np.int16) - ref_r.astype(np.int16)).max() <= 1


def test_smoke_entry(cuda):
    import __graft_entry__ as g
    g.smoke()


def test_full_frame_c2_strict_bit_exact_and_fast_statistics(cuda):
    """BASELINE.json configs[1] at full size (1683x1869 -> 2244x2492), one eye, reference default radius 0.5 and
    mask-off radius 2.0: strict is bit-identical end to end; fast is <= 1 LSB per pass on identical inputs, and its
    composed EASU->RCAS deviation is reported (RCAS amplifies 1-LSB intermediates in dark regions)."""
    import os
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    iw, ih, scale = 1683, 1869, 0.75
    ow, oh = po.output_size(iw, ih, scale)
    src = synth.natural_rgba8(iw, ih, 1)
    t = torch.from_numpy(src).to(cuda)
    nt = os.cpu_count() or 1
    for radius in (0.5, 2.0):
        uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
        sc = po.sharpen_constants(0, True, ow, oh, radius=radius, sharpness=0.9)
        easu = po.easu(src, ow, oh, uc, nthreads=nt)
        want = po.rcas(easu, sc, nthreads=nt)
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=radius,
                                          mathMode=ovr.MATH_STRICT))
        got = pp.apply(0, t).cpu().numpy()
        pp.close()
        assert np.array_equal(got, want)
        f_easu = torch.empty((oh, ow, 4), dtype=torch.uint8, device=cuda)
        f_rcas = torch.empty_like(f_easu)
        ovr.fsr_easu(t, f_easu, uc.words(), ovr.MATH_FAST)
        ovr.fsr_rcas(torch.from_numpy(easu).to(cuda), f_rcas, sc.words(), ovr.MATH_FAST)
        torch.cuda.synchronize()
        de = np.abs(f_easu.cpu().numpy().astype(np.int16) - easu.astype(np.int16))
        dr = np.abs(f_rcas.cpu().numpy().astype(np.int16) - want.astype(np.int16))
        assert de.max() <= 1 and dr.max() <= 1
        assert (de > 0).mean() < 2e-3 and (dr > 0).mean() < 2e-3  # and rare
        pp = ovr.PostProcessor(ovr.Config(fsrEnabled=True, renderScale=scale, sharpness=0.9, radius=radius,
                                          mathMode=ovr.MATH_FAST))
        comp = np.abs(pp.apply(0, t).cpu().numpy().astype(np.int16) - want.astype(np.int16))
        pp.close()
        print(f"radius {radius}: fast per-pass mismatch EASU {(de > 0).mean():.2e} RCAS {(dr > 0).mean():.2e}; "
              f"composed: {(comp > 1).mean():.2e} of channel values beyond 1 LSB, max {comp.max()}")
        # measured over the C2-sized natural / noise images (tools/fastmath_error.py, profiles/r2_fastmath_composed_error.txt):
        # max 3 LSB, 6e-6 of the channel values beyond 1 LSB -- a 1-LSB difference in the RGBA8 intermediate is
        # amplified where RCAS divides by a dark ring maximum
        assert comp.max() <= 4 and (comp > 1).mean() < 1e-4


@pytest.mark.parametrize("iw,ih,scale", [(33, 47, 0.75), (211, 157, 0.75), (129, 65, 0.59), (100, 40, 1.0 / 1.07), (77, 50, 0.5)])
def test_tma_and_plain_loaders_agree(cuda, iw, ih, scale):
    """The same eye through the TMA tile loader (pitch-aligned image) and the plain-load fallback (tight pitch,
    odd width): both bit-identical to the oracle in strict mode, identical to each other in fast mode."""
    import torch
    import openvr_fsr_b200 as ovr
    from openvr_fsr_b200 import synth
    from oracle import pyoracle as po
    ow, oh = po.output_size(iw, ih, scale)
    src = synth.natural_rgba8(iw, ih, 7)
    for radius in (2.0, 0.45):
        uc = po.upscale_constants(0, True, iw, ih, ow, oh, radius=radius)
        sc = po.sharpen_constants(0, True, ow, oh, radius=radius, sharpness=0.8)
        easu = po.easu(src, ow, oh, uc)
        rcas = po.rcas(easu, sc)
        tight_src, tight_mid = torch.from_numpy(src).to(cuda), torch.from_numpy(easu).to(cuda)
        pitched_src, pitched_mid = ovr.to_image(src, cuda), ovr.to_image(easu, cuda)
        assert pitched_src.stride(0) % 16 == 0
        outs = {}
        for mode in (ovr.MATH_STRICT, ovr.MATH_FAST):
            for name, s_, m_ in (("tight", tight_src, tight_mid), ("pitched", pitched_src, pitched_mid)):
                e = ovr.alloc_image(ow, oh, torch.uint8, cuda) if name == "pitched" else torch.empty((oh, ow, 4), dtype=torch.uint8, device=cuda)
                r = torch.empty((oh, ow, 4), dtype=torch.uint8, device=cuda)
                ovr.fsr_easu(s_, e, uc.words(), mode)
                ovr.fsr_rcas(m_, r, sc.words(), mode)
                torch.cuda.synchronize()
                outs[(mode, name)] = (e.cpu().numpy(), r.cpu().numpy())
        for name in ("tight", "pitched"):
            assert np.array_equal(outs[(ovr.MATH_STRICT, name)][0], easu) and np.array_equal(outs[(ovr.MATH_STRICT, name)][1], rcas)
        assert np.array_equal(outs[(ovr.MATH_FAST, "tight")][0], outs[(ovr.MATH_FAST, "pitched")][0])
        assert np.array_equal(outs[(ovr.MATH_FAST, "tight")][1], outs[(ovr.MATH_FAST, "pitched")][1])


def test_strict_rcas_reciprocal_is_exact_on_every_unorm8_operand(cuda):
    """Strict RCAS uses MUFU.RCP + one Newton step instead of rcp.rn for UNORM8 sources; exhaustive device check over
    all 512 operands (4*k/255 and 4*k/255-4) that it equals rcp.rn bit for bit."""
    import ctypes as C
    from openvr_fsr_b200 import _lib as L
    bad, n = C.c_uint32(99), C.c_uint32(0)
    L.check(L.lib().ovrfsr_selftest_rcp(C.byref(bad), C.byref(n)))
    assert n.value == 512 and bad.value == 0
