/* abi_smoke.c -- include/ovrfsr.h used from plain C (C11, -pedantic): the host-only entry points of the C ABI, no GPU.
 * Built and run by tests/test_c_abi.py.  Prints one line per check; exit code 0 = all good. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ovrfsr.h"

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

int main(int argc, char **argv) {
  ovrfsr_config cfg;
  uint32_t w = 0, h = 0, con[16], rc4[4], up[24], sh[12], cas[8];
  char name[96];
  ovrfsr_config_default(&cfg);
  CHECK(cfg.struct_size == sizeof(cfg) && cfg.fsr_enabled == 0 && cfg.render_scale == 1.0f && cfg.sharpness == 0.75f &&
        cfg.radius == 0.5f && cfg.math_mode == OVRFSR_MATH_STRICT && cfg.output_format == OVRFSR_FORMAT_AUTO);
  ovrfsr_output_size(1683, 1869, 0.75f, &w, &h);
  CHECK(w == 2244 && h == 2492);
  ovrfsr_output_size(2244, 2492, 1.3f, &w, &h);
  CHECK(w == 2917 && h == 3239); /* the code's formula, not the README's 2915x3240 */
  /* known-answer words */
  ovrfsr_fsr_easu_con(con, 1683.f, 1869.f, 1683.f, 1869.f, 2244.f, 2492.f);
  CHECK(con[0] == 0x3f400000u && con[1] == 0x3f400000u && con[2] == 0xbe000000u && con[3] == 0xbe000000u);
  CHECK(con[4] == 0x3a1bc28cu && con[5] == 0x3a0c424bu && con[6] == 0x3a1bc28cu && con[7] == 0xba0c424bu);
  ovrfsr_fsr_rcas_con(rc4, 2.0f - 2.0f * 0.9f);
  CHECK(rc4[0] == 0x3f5edc66u && rc4[1] == 0x3af63af6u);
  cfg.fsr_enabled = 1; cfg.render_scale = 0.75f; cfg.sharpness = 0.9f;
  ovrfsr_make_upscale_constants(up, &cfg, 0, 1, 1683, 1869, 2244, 2492);
  ovrfsr_make_sharpen_constants(sh, &cfg, 0, 1, 2244, 2492);
  CHECK(up[16] == 1122 && up[17] == 1246 && up[20] == 623 && up[21] == 388129u && up[22] == 2244 && up[23] == 2492);
  CHECK(sh[0] == 0x3f5edc66u && sh[4] == 1122 && sh[9] == 388129u);
  /* render-size / MIP-bias policy */
  w = 2244; h = 2492;
  ovrfsr_recommended_render_size(&cfg, &w, &h);
  CHECK(w == 1683 && h == 1869);
  CHECK(ovrfsr_mip_lod_bias(960, 1920) == -1.0f);
  CHECK(ovrfsr_sampler_lod_bias(0.0f, 16, -0.5f) == -0.5f && ovrfsr_sampler_lod_bias(0.0f, 1, -0.5f) == 0.0f);
  /* CasSetup: scale terms of the C2 shape */
  ovrfsr_cas_setup(cas, 0.0f, 1.0f, 1683.f, 1869.f, 2244.f, 2492.f);
  CHECK(cas[0] == 0x3f400000u && cas[2] == 0xbe000000u && cas[4] == 0xbe000000u /* -1/8 */ && cas[7] == 0x3f800000u);
  CHECK(ovrfsr_capture_filename(&cfg, 0, name, sizeof(name)) == OVRFSR_OK && strstr(name, "_fsr_s90_r50.dds") != NULL);
  CHECK(ovrfsr_nis_coef_scale() != NULL && ovrfsr_nis_coef_usm() != NULL && ovrfsr_nis_coef_scale()[2] > 0.9f);
  CHECK(strcmp(ovrfsr_status_string(OVRFSR_OK), ovrfsr_status_string(OVRFSR_ERR_CUDA)) != 0);
  CHECK(ovrfsr_version() != 0);
  /* DDS round trip in a scratch directory */
  if (argc > 1) {
    unsigned char px[4 * 3 * 2];
    char path[512];
    ovrfsr_image im, back;
    int i;
    for (i = 0; i < (int)sizeof(px); ++i) px[i] = (unsigned char)(i * 7);
    memset(&im, 0, sizeof(im));
    im.data = px; im.width = 3; im.height = 2; im.pitch = 12; im.format = OVRFSR_FORMAT_RGBA8; im.array_slices = 1;
    snprintf(path, sizeof(path), "%s/abi_smoke.dds", argv[1]);
    CHECK(ovrfsr_dds_write(path, &im) == OVRFSR_OK);
    memset(&back, 0, sizeof(back));
    CHECK(ovrfsr_dds_read(path, &back) == OVRFSR_OK && back.width == 3 && back.height == 2 && back.format == OVRFSR_FORMAT_RGBA8 &&
          memcmp(back.data, px, sizeof(px)) == 0);
    ovrfsr_host_free(back.data);
  }
  /* a context can be created and configured without a device; GPU work then reports pass-through / CUDA errors */
  {
    ovrfsr_ctx *ctx = NULL;
    ovrfsr_config got;
    CHECK(ovrfsr_create(&ctx, &cfg) == OVRFSR_OK && ctx != NULL);
    CHECK(ovrfsr_get_config(ctx, &got) == OVRFSR_OK && got.render_scale == 0.75f);
    CHECK(ovrfsr_reset(ctx) == OVRFSR_OK);
    ovrfsr_destroy(ctx);
  }
  printf(fails ? "abi_smoke: %d check(s) failed\n" : "abi_smoke: ok\n", fails);
  return fails ? 1 : 0;
}
