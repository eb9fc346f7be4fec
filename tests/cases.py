"""Shared small parity cases (SURVEY.md section 8d 'parity corner set')."""
import numpy as np

from openvr_fsr_b200 import synth


def corner_images(w, h):
    """name -> (h, w, 4) uint8"""
    yy, xx = np.mgrid[0:h, 0:w]
    out = {}
    for v in (0, 128, 255):
        out[f"const{v}"] = np.full((h, w, 4), v, np.uint8)
    imp = np.zeros((h, w, 4), np.uint8); imp[h // 2, w // 2] = 255; out["impulse"] = imp
    chk = np.zeros((h, w, 4), np.uint8); chk[(xx + yy) % 2 == 0] = 255; out["checker"] = chk
    for name, mask in (("hstep", xx >= w // 2), ("vstep", yy >= h // 2), ("diag", xx * h >= yy * w)):
        im = np.full((h, w, 4), 30, np.uint8); im[mask] = (220, 180, 90, 255); out[name] = im
    out["uniform"] = synth.uniform_rgba8(w, h, 0)
    out["natural"] = synth.natural_rgba8(w, h, 1)
    for im in out.values():
        im[..., 3] = np.where(im[..., 3] == 0, 255, im[..., 3])
    return out


SMALL = [(17, 13, 0.75), (16, 16, 0.5), (33, 47, 0.75), (40, 24, 1.3), (21, 35, 0.59)]
