/*
 * selfcheck.c -- runs every pass of the restated oracle on small, ragged images; meant to be built with
 * -fsanitize=address,undefined (make -C oracle selfcheck) so that an out-of-bounds read or undefined operation in
 * the CHECKER itself cannot hide behind a lucky comparison.  TEST INFRASTRUCTURE ONLY (tests/test_oracle_sanitized.py).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ovr_oracle.h"

static uint32_t rng_state = 12345u;
static uint32_t rnd(void) { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static ovo_image make(int w, int h, int fmt, int fill) {
  ovo_image im;
  const int bpp = fmt == OVO_FMT_RGBA32F ? 16 : (fmt == OVO_FMT_RGBA16F ? 8 : 4);
  im.width = w; im.height = h; im.pitch = w * bpp; im.format = fmt;
  im.data = malloc((size_t)im.pitch * (size_t)h); /* exact size: any overrun is caught by ASan */
  if (fill) {
    uint8_t *p = (uint8_t *)im.data;
    for (size_t i = 0; i < (size_t)im.pitch * (size_t)h; ++i) p[i] = (uint8_t)rnd();
    if (fmt == OVO_FMT_RGBA16F) { /* keep the halves finite: clear the top exponent bit */
      uint16_t *q = (uint16_t *)im.data;
      for (size_t i = 0; i < (size_t)w * (size_t)h * 4; ++i) q[i] &= 0xbbffu;
    }
  }
  return im;
}

int main(void) {
  static const int sizes[][2] = {{1, 1}, {2, 3}, {5, 4}, {16, 16}, {17, 13}, {33, 47}, {64, 24}};
  static const float scales[] = {0.5f, 0.59f, 0.75f, 1.0f, 1.3f};
  static const int fmts[] = {OVO_FMT_RGBA8, OVO_FMT_BGRA8, OVO_FMT_RGBA16F, OVO_FMT_RGB10A2};
  const float proj[4] = {0.45f, 0.52f, 0.55f, 0.48f};
  int runs = 0;
  for (unsigned si = 0; si < sizeof(sizes) / sizeof(sizes[0]); ++si)
    for (unsigned ci = 0; ci < sizeof(scales) / sizeof(scales[0]); ++ci)
      for (unsigned fi = 0; fi < sizeof(fmts) / sizeof(fmts[0]); ++fi) {
        const int iw = sizes[si][0], ih = sizes[si][1], fmt = fmts[fi];
        const float scale = scales[ci], radius = (runs & 1) ? 0.4f : 2.0f;
        uint32_t ow, oh;
        ovo_output_size((uint32_t)iw, (uint32_t)ih, scale, &ow, &oh);
        if (ow == 0 || oh == 0) continue;
        const int ofmt = fmt == OVO_FMT_RGB10A2 ? OVO_FMT_RGB10A2 : ((runs & 2) ? OVO_FMT_RGBA32F : OVO_FMT_RGBA8);
        ovo_image src = make(iw, ih, fmt, 1), mid = make((int)ow, (int)oh, ofmt, 0), dst = make((int)ow, (int)oh, ofmt, 0);
        ovo_upscale_constants uc;
        ovo_sharpen_constants sc;
        ovo_make_upscale_constants(&uc, runs & 1, 1, (uint32_t)iw, (uint32_t)ih, ow, oh, proj, radius);
        ovo_make_sharpen_constants(&sc, runs & 1, 1, ow, oh, proj, radius, 0.9f, runs & 4);
        if (ovo_fsr_easu(&src, &mid, &uc, 2) != 0 || ovo_fsr_rcas(&mid, &dst, &sc, 2) != 0) { puts("FSR pass failed"); return 1; }
        if (scale <= 1.0f && fmt != OVO_FMT_RGB10A2) {
          ovo_nis_config nc;
          ovo_cas_constants cc;
          ovo_make_nis_config(&nc, scale == 1.0f, runs & 1, 1, (uint32_t)iw, (uint32_t)ih, ow, oh, proj, radius, 0.7f, runs & 4);
          if ((scale == 1.0f ? ovo_nis_sharpen(&src, &dst, &nc, 2) : ovo_nis_scaler(&src, &dst, &nc, 2)) != 0) { puts("NIS pass failed"); return 1; }
          ovo_cas_setup(&cc, 0.8f, (runs & 8) ? 0.1f : 1.0f, (float)iw, (float)ih, (float)ow, (float)oh);
          if (ovo_cas(&src, &dst, &cc, scale == 1.0f, 2) != 0) { puts("CAS pass failed"); return 1; }
          if (ovo_cas(&src, &dst, &cc, 0, 1) != 0) { puts("CAS upscale pass failed"); return 1; }
        }
        free(src.data); free(mid.data); free(dst.data);
        ++runs;
      }
  printf("oracle selfcheck: %d configurations ok\n", runs);
  return 0;
}
