// fsr_kernels.cuh -- FSR1 EASU (upscale) and RCAS (sharpen) for sm_100a.
//
// Replaces the two D3D11 compute dispatches of the reference:
//   g_FSRUpscaleShader = src/fsr/fsr_easu.hlsl:38-64 -> FsrEasuF, src/fsr/ffx_fsr1.h:315-437
//   g_FSRSharpenShader = src/fsr/fsr_rcas.hlsl:29-55 -> FsrRcasF, src/fsr/ffx_fsr1.h:684-769
// (citations relative to /root/reference/).  Not a translation of the HLSL: the work is
// re-tiled for a CTA of 8 warps over a 64x32 output tile,
//   * the source tile (+halo) is decoded ONCE per texel into shared memory (float4 rgb + luma);
//     the reference decodes every texel ~21 times through 36 Gather4s per pixel,
//   * EASU's direction/length analysis (FsrEasuSetF) depends only on the '+' neighbourhood of a
//     SOURCE texel, so it is evaluated once per source texel into a shared feature tile and
//     each output pixel bilinearly blends four float4 features instead of redoing 4 x ~25 ops,
//   * one warp owns one 16x16 output group, so the reference's per-workgroup radius test is a
//     warp-uniform branch (no divergence) with the bilinear / copy fast path fused in.
// Arithmetic order inside each reference function is kept where kStrict is set, so the strict
// build is bit-identical to the reference lines; the fast build regroups the 12 tap offsets
// (v = A*o - A*pp for integer o) and lets ptxas contract FMAs.
#pragma once

#include "device_common.cuh"

namespace ovrfsr {
inline namespace OVRFSR_MODE_NS {

constexpr int kTileW = 64;   // output tile of one CTA
constexpr int kTileH = 32;
constexpr int kThreads = 256; // 8 warps, one per 16x16 group

struct EasuArgs {
  ImageRO src;
  ImageRW dst;
  float c0x, c0y, c0z, c0w; // const0 of FsrEasuCon: out->in scale and offset
  uint32_t centre[4];
  uint32_t radiusSq;
  float radW, radH;         // (float)Radius.z, (float)Radius.w for Bilinear()
  int tileW, tileH;         // shared source tile extent in texels
};

struct RcasArgs {
  ImageRO src;
  ImageRW dst;
  float sharp;              // const0[0] as float
  float tintGB;             // 1 - debug*0.3 (fsr_rcas.hlsl:46)
  uint32_t centre[4];
  uint32_t radiusSq;
};

// out pixel -> source position; the same instruction sequence is used for the tile origin and
// for every pixel so that both agree to the bit.
__device__ __forceinline__ float easu_pos(int p, float scale, float offs) { return mul_add_unfused((float)p, scale, offs); }

// FsrEasuSetF (ffx_fsr1.h:275-313) minus the bilinear weight: the per-source-texel part.
//   a
// b c d      returns (dirX, dirY, lenX, lenY)
//   e
__device__ __forceinline__ float4 easu_feature(float lA, float lB, float lC, float lD, float lE) {
  const float dc = lD - lC, cb = lC - lB;
  float lenX = prx_lo_rcp(fmaxf(fabsf(dc), fabsf(cb)));
  const float dirX = lD - lB;
  lenX = __saturatef(fabsf(dirX) * lenX);
  lenX *= lenX;
  const float ec = lE - lC, ca = lC - lA;
  float lenY = prx_lo_rcp(fmaxf(fabsf(ec), fabsf(ca)));
  const float dirY = lE - lA;
  lenY = __saturatef(fabsf(dirY) * lenY);
  lenY *= lenY;
  return make_float4(dirX, dirY, lenX, lenY);
}

// FsrEasuTapF, ffx_fsr1.h:239-272 (reference operation order)
__device__ __forceinline__ void easu_tap_ref(float &aR, float &aG, float &aB, float &aW, float offX, float offY,
                                             float dirX, float dirY, float len0, float len1, float lob, float clp,
                                             const float4 c) {
  float vx = (offX * dirX) + (offY * dirY);
  float vy = (offX * (-dirY)) + (offY * dirX);
  vx *= len0;
  vy *= len1;
  float d2 = vx * vx + vy * vy;
  d2 = fminf(d2, clp);
  float wB = (float)(2.0 / 5.0) * d2 + (float)(-1.0);
  float wA = lob * d2 + (float)(-1.0);
  wB *= wB;
  wA *= wA;
  wB = (float)(25.0 / 16.0) * wB + (float)(-(25.0 / 16.0 - 1.0));
  const float w = wB * wA;
  aR += c.x * w; aG += c.y * w; aB += c.z * w;
  aW += w;
}

// tap with the rotated/scaled offset already formed (fast mode)
__device__ __forceinline__ void easu_tap_v(float &aR, float &aG, float &aB, float &aW, float vx, float vy, float lob,
                                           float clp, const float4 c) {
  float d2 = fminf(fmaf(vx, vx, vy * vy), clp);
  float wB = fmaf((float)(2.0 / 5.0), d2, -1.0f);
  float wA = fmaf(lob, d2, -1.0f);
  wB *= wB;
  wA *= wA;
  wB = fmaf((float)(25.0 / 16.0), wB, (float)(-(25.0 / 16.0 - 1.0)));
  const float w = wB * wA;
  aR = fmaf(c.x, w, aR); aG = fmaf(c.y, w, aG); aB = fmaf(c.z, w, aB);
  aW += w;
}

// FsrEasuF, ffx_fsr1.h:315-437, for the output pixel whose 'f' texel sits at tile coords (ix,iy)
// with sub-texel phase (ppx,ppy).  sC = decoded colours (xyz) , sF = per-texel features.
__device__ __forceinline__ float3 easu_filter(const float4 *__restrict__ sC, const float4 *__restrict__ sF, int tw,
                                              int ix, int iy, float ppx, float ppy) {
  // direction + length: bilinear blend of the four corner features (:380-386)
  const float4 *fr = sF + iy * tw + ix;
  const float4 Ff = fr[0], Fg = fr[1], Fj = fr[tw], Fk = fr[tw + 1];
  const float qx = 1.0f - ppx, qy = 1.0f - ppy;
  const float wf = qx * qy, wg = ppx * qy, wj = qx * ppy, wk = ppx * ppy;
  float dirX = 0.0f, dirY = 0.0f, len = 0.0f;
  dirX += Ff.x * wf; len += Ff.z * wf; dirY += Ff.y * wf; len += Ff.w * wf;
  dirX += Fg.x * wg; len += Fg.z * wg; dirY += Fg.y * wg; len += Fg.w * wg;
  dirX += Fj.x * wj; len += Fj.z * wj; dirY += Fj.y * wj; len += Fj.w * wj;
  dirX += Fk.x * wk; len += Fk.z * wk; dirY += Fk.y * wk; len += Fk.w * wk;

  // normalise (:389-395)
  const float dir2x = dirX * dirX, dir2y = dirY * dirY;
  float dirR = dir2x + dir2y;
  const bool zro = dirR < (float)(1.0 / 32768.0);
  dirR = prx_lo_rsq(dirR);
  dirR = zro ? 1.0f : dirR;
  dirX = zro ? 1.0f : dirX;
  dirX *= dirR;
  dirY *= dirR;
  // shape (:397-409)
  len = len * 0.5f;
  len *= len;
  const float stretch = (dirX * dirX + dirY * dirY) * prx_lo_rcp(fmaxf(fabsf(dirX), fabsf(dirY)));
  const float len0 = 1.0f + (stretch - 1.0f) * len;
  const float len1 = 1.0f + (-0.5f) * len;
  const float lob = 0.5f + (float)((1.0 / 4.0 - 0.04) - 0.5) * len;
  const float clp = prx_lo_rcp(lob);

  //    b c
  //  e f g h
  //  i j k l
  //    n o
  const float4 *r0 = sC + (iy - 1) * tw + ix;
  const float4 *r1 = r0 + tw, *r2 = r1 + tw, *r3 = r2 + tw;
  const float4 b = r0[0], c = r0[1];
  const float4 e = r1[-1], f = r1[0], g = r1[1], h = r1[2];
  const float4 i = r2[-1], j = r2[0], k = r2[1], l = r2[2];
  const float4 n = r3[0], o = r3[1];

  // min/max of the four nearest (:416-419)
  const float mnR = fminf(fminf(f.x, fminf(g.x, j.x)), k.x), mxR = fmaxf(fmaxf(f.x, fmaxf(g.x, j.x)), k.x);
  const float mnG = fminf(fminf(f.y, fminf(g.y, j.y)), k.y), mxG = fmaxf(fmaxf(f.y, fmaxf(g.y, j.y)), k.y);
  const float mnB = fminf(fminf(f.z, fminf(g.z, j.z)), k.z), mxB = fmaxf(fmaxf(f.z, fmaxf(g.z, j.z)), k.z);

  float aR = 0.0f, aG = 0.0f, aB = 0.0f, aW = 0.0f;
  if constexpr (kStrict) {
    // reference accumulation order (:423-434): b c i j f e k l h g o n
    easu_tap_ref(aR, aG, aB, aW, 0.0f - ppx, -1.0f - ppy, dirX, dirY, len0, len1, lob, clp, b);
    easu_tap_ref(aR, aG, aB, aW, 1.0f - ppx, -1.0f - ppy, dirX, dirY, len0, len1, lob, clp, c);
    easu_tap_ref(aR, aG, aB, aW, -1.0f - ppx, 1.0f - ppy, dirX, dirY, len0, len1, lob, clp, i);
    easu_tap_ref(aR, aG, aB, aW, 0.0f - ppx, 1.0f - ppy, dirX, dirY, len0, len1, lob, clp, j);
    easu_tap_ref(aR, aG, aB, aW, 0.0f - ppx, 0.0f - ppy, dirX, dirY, len0, len1, lob, clp, f);
    easu_tap_ref(aR, aG, aB, aW, -1.0f - ppx, 0.0f - ppy, dirX, dirY, len0, len1, lob, clp, e);
    easu_tap_ref(aR, aG, aB, aW, 1.0f - ppx, 1.0f - ppy, dirX, dirY, len0, len1, lob, clp, k);
    easu_tap_ref(aR, aG, aB, aW, 2.0f - ppx, 1.0f - ppy, dirX, dirY, len0, len1, lob, clp, l);
    easu_tap_ref(aR, aG, aB, aW, 2.0f - ppx, 0.0f - ppy, dirX, dirY, len0, len1, lob, clp, h);
    easu_tap_ref(aR, aG, aB, aW, 1.0f - ppx, 0.0f - ppy, dirX, dirY, len0, len1, lob, clp, g);
    easu_tap_ref(aR, aG, aB, aW, 1.0f - ppx, 2.0f - ppy, dirX, dirY, len0, len1, lob, clp, o);
    easu_tap_ref(aR, aG, aB, aW, 0.0f - ppx, 2.0f - ppy, dirX, dirY, len0, len1, lob, clp, n);
  } else {
    // v(o) = M*(o - pp) with M = diag(len0,len1)*Rot(dir): for the integer offsets o=(ox,oy),
    // ox,oy in {-1,0,1,2}, v = ox*colX + oy*colY - M*pp -> two adds per tap instead of 4 mul + 2 fma.
    const float m00 = dirX * len0, m01 = dirY * len0;  // vx = m00*ox + m01*oy
    const float m10 = -dirY * len1, m11 = dirX * len1; // vy = m10*ox + m11*oy
    const float cx = -fmaf(m00, ppx, m01 * ppy), cy = -fmaf(m10, ppx, m11 * ppy);
    const float x_m1 = cx - m00, x_0 = cx, x_1 = cx + m00, x_2 = fmaf(2.0f, m00, cx);
    const float y_m1 = cy - m10, y_0 = cy, y_1 = cy + m10, y_2 = fmaf(2.0f, m10, cy);
    const float m01_2 = m01 + m01, m11_2 = m11 + m11;
    easu_tap_v(aR, aG, aB, aW, x_0 - m01, y_0 - m11, lob, clp, b);
    easu_tap_v(aR, aG, aB, aW, x_1 - m01, y_1 - m11, lob, clp, c);
    easu_tap_v(aR, aG, aB, aW, x_m1 + m01, y_m1 + m11, lob, clp, i);
    easu_tap_v(aR, aG, aB, aW, x_0 + m01, y_0 + m11, lob, clp, j);
    easu_tap_v(aR, aG, aB, aW, x_0, y_0, lob, clp, f);
    easu_tap_v(aR, aG, aB, aW, x_m1, y_m1, lob, clp, e);
    easu_tap_v(aR, aG, aB, aW, x_1 + m01, y_1 + m11, lob, clp, k);
    easu_tap_v(aR, aG, aB, aW, x_2 + m01, y_2 + m11, lob, clp, l);
    easu_tap_v(aR, aG, aB, aW, x_2, y_2, lob, clp, h);
    easu_tap_v(aR, aG, aB, aW, x_1, y_1, lob, clp, g);
    easu_tap_v(aR, aG, aB, aW, x_1 + m01_2, y_1 + m11_2, lob, clp, o);
    easu_tap_v(aR, aG, aB, aW, x_0 + m01_2, y_0 + m11_2, lob, clp, n);
  }
  // normalise and de-ring (:437)
  const float r = rcp_mode(aW);
  return make_float3(fminf(mxR, fmaxf(mnR, aR * r)), fminf(mxG, fmaxf(mnG, aG * r)), fminf(mxB, fmaxf(mnB, aB * r)));
}

// Bilinear(), fsr_easu.hlsl:33-36: SampleLevel(linearClamp, float2(pos)/Radius.zw) -- no half-texel
// offset; coordinates snapped to 1/256 texel like D3D11's fixed-point sampler.  Reads the same clamped colour tile.
__device__ __forceinline__ float3 easu_bilinear(const float4 *__restrict__ sC, int tw, int th, int sx0, int sy0, int x,
                                                int y, const EasuArgs &a) {
  const float u = (float)x / a.radW, v = (float)y / a.radH;
  const float sx = snap_subtexel(mul_add_unfused(u, (float)a.src.w, -0.5f));
  const float sy = snap_subtexel(mul_add_unfused(v, (float)a.src.h, -0.5f));
  const float fx0 = floorf(sx), fy0 = floorf(sy);
  const float fx = sx - fx0, fy = sy - fy0;
  const int tx0 = clampi((int)fx0 - sx0, 0, tw - 1), tx1 = clampi((int)fx0 + 1 - sx0, 0, tw - 1);
  const int ty0 = clampi((int)fy0 - sy0, 0, th - 1), ty1 = clampi((int)fy0 + 1 - sy0, 0, th - 1);
  const float4 c00 = sC[ty0 * tw + tx0], c10 = sC[ty0 * tw + tx1];
  const float4 c01 = sC[ty1 * tw + tx0], c11 = sC[ty1 * tw + tx1];
  const float wx0 = 1.0f - fx, wy0 = 1.0f - fy;
  const float tR = c00.x * wx0 + c10.x * fx, bR = c01.x * wx0 + c11.x * fx;
  const float tG = c00.y * wx0 + c10.y * fx, bG = c01.y * wx0 + c11.y * fx;
  const float tB = c00.z * wx0 + c10.z * fx, bB = c01.z * wx0 + c11.z * fx;
  return make_float3(tR * wy0 + bR * fy, tG * wy0 + bG * fy, tB * wy0 + bB * fy);
}

template <int FIN, int FOUT>
__global__ void __launch_bounds__(kThreads, 2) easu_kernel(const EasuArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tw = a.tileW, th = a.tileH, tn = tw * th;
  float4 *sC = reinterpret_cast<float4 *>(smem_raw); // decoded colour, w = luma*2
  float4 *sF = sC + tn;                              // (dirX, dirY, lenX, lenY) per texel
  float *sL = reinterpret_cast<float *>(sF + tn);    // luma*2 plane (conflict-free stencil reads)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH;
  // source tile origin: one texel left/above the 'f' texel of the tile's first pixel
  const int sx0 = (int)floorf(easu_pos(ox0, a.c0x, a.c0z)) - 1;
  const int sy0 = (int)floorf(easu_pos(oy0, a.c0y, a.c0w)) - 1;

  // this warp's 16x16 group and its radius test (warp-uniform)
  const uint32_t ggx = blockIdx.x * (kTileW / 16) + (warp & 3), ggy = blockIdx.y * (kTileH / 16) + (warp >> 2);
  const bool inside = group_inside(ggx * 16u + 8u, ggy * 16u + 8u, a.centre, a.radiusSq);

  // ---- stage 1: decode the clamped source tile once ------------------------------------------
  for (int ty = warp; ty < th; ty += kThreads / 32) {
    const int gy = clampi(sy0 + ty, 0, a.src.h - 1);
    const uint8_t *row = a.src.ptr + (size_t)gy * a.src.pitch;
    for (int tx = lane; tx < tw; tx += 32) {
      const int gx = clampi(sx0 + tx, 0, a.src.w - 1);
      float4 c = fetch_texel<FIN>(row, gx);
      const float l2 = c.z * 0.5f + (c.x * 0.5f + c.y); // luma*2, ffx_fsr1.h:363 (x0.5 is exact, so FMA-safe)
      c.w = l2;
      sC[ty * tw + tx] = c;
      sL[ty * tw + tx] = l2;
    }
  }
  const int anyInside = __syncthreads_or(inside);

  // ---- stage 2: per-source-texel direction/length features (only where EASU will run) ---------
  if (anyInside) {
    for (int ty = 1 + warp; ty < th - 1; ty += kThreads / 32) {
      const float *l = sL + ty * tw;
      for (int tx = 1 + lane; tx < tw - 1; tx += 32)
        sF[ty * tw + tx] = easu_feature(l[tx - tw], l[tx - 1], l[tx], l[tx + 1], l[tx + tw]);
    }
    __syncthreads();
  }

  // ---- stage 3: one warp per 16x16 group, 8 pixels per lane ------------------------------------
  const int x = (int)ggx * 16 + (lane & 15);
  if (x >= a.dst.w) return;
  const float ppx_full = easu_pos(x, a.c0x, a.c0z);
  const float fpx = floorf(ppx_full);
  const float ppx = ppx_full - fpx;
  const int ix = (int)fpx - sx0;
#pragma unroll 2
  for (int k = 0; k < 8; ++k) {
    const int y = (int)ggy * 16 + (lane >> 4) + 2 * k;
    if (y >= a.dst.h) break;
    float3 c;
    if (inside) {
      const float ppy_full = easu_pos(y, a.c0y, a.c0w);
      const float fpy = floorf(ppy_full);
      c = easu_filter(sC, sF, tw, ix, (int)fpy - sy0, ppx, ppy_full - fpy);
    } else {
      c = easu_bilinear(sC, tw, th, sx0, sy0, x, y, a);
    }
    store_texel<FOUT>(a.dst.ptr + (size_t)y * a.dst.pitch, x, c.x, c.y, c.z, 1.0f);
  }
}

// ------------------------------------------------------------------------------------------------
// RCAS
// ------------------------------------------------------------------------------------------------
constexpr int kRcasTW = kTileW + 4; // 1-texel halo each side, padded to a multiple of 4 texels
constexpr int kRcasTH = kTileH + 2;

// FsrRcasF, ffx_fsr1.h:684-769 (FSR_RCAS_DENOISE / PASSTHROUGH_ALPHA undefined: fsr_rcas.hlsl:1-4)
__device__ __forceinline__ float3 rcas_filter(const float4 b, const float4 d, const float4 e, const float4 f,
                                              const float4 h, float sharp) {
  const float mn4R = fminf(fminf(b.x, fminf(d.x, f.x)), h.x), mx4R = fmaxf(fmaxf(b.x, fmaxf(d.x, f.x)), h.x);
  const float mn4G = fminf(fminf(b.y, fminf(d.y, f.y)), h.y), mx4G = fmaxf(fmaxf(b.y, fmaxf(d.y, f.y)), h.y);
  const float mn4B = fminf(fminf(b.z, fminf(d.z, f.z)), h.z), mx4B = fmaxf(fmaxf(b.z, fmaxf(d.z, f.z)), h.z);
  // limiters need full-precision reciprocals (:748-755)
  const float hitMinR = mn4R * rcp_mode(4.0f * mx4R);
  const float hitMinG = mn4G * rcp_mode(4.0f * mx4G);
  const float hitMinB = mn4B * rcp_mode(4.0f * mx4B);
  const float hitMaxR = (1.0f - mx4R) * rcp_mode(4.0f * mn4R + (-4.0f));
  const float hitMaxG = (1.0f - mx4G) * rcp_mode(4.0f * mn4G + (-4.0f));
  const float hitMaxB = (1.0f - mx4B) * rcp_mode(4.0f * mn4B + (-4.0f));
  const float lobeR = fmaxf(-hitMinR, hitMaxR), lobeG = fmaxf(-hitMinG, hitMaxG), lobeB = fmaxf(-hitMinB, hitMaxB);
  const float lobe =
      fmaxf((float)(-(0.25 - (1.0 / 16.0))), fminf(fmaxf(lobeR, fmaxf(lobeG, lobeB)), 0.0f)) * sharp;
  const float rcpL = prx_med_rcp(4.0f * lobe + 1.0f);
  return make_float3((lobe * b.x + lobe * d.x + lobe * h.x + lobe * f.x + e.x) * rcpL,
                     (lobe * b.y + lobe * d.y + lobe * h.y + lobe * f.y + e.y) * rcpL,
                     (lobe * b.z + lobe * d.z + lobe * h.z + lobe * f.z + e.z) * rcpL);
}

template <int FIN, int FOUT>
__global__ void __launch_bounds__(kThreads, 2) rcas_kernel(const RcasArgs a) {
  __shared__ __align__(16) float4 sC[kRcasTH * kRcasTW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH;
  const int sx0 = ox0 - 1, sy0 = oy0 - 1;

  // decode tile + 1-texel ring; Texture2D.Load semantics: out of bounds reads 0 (fsr_rcas.hlsl:18)
  for (int ty = warp; ty < kRcasTH; ty += kThreads / 32) {
    const int gy = sy0 + ty;
    const bool rowOk = gy >= 0 && gy < a.src.h;
    const uint8_t *row = a.src.ptr + (size_t)(rowOk ? gy : 0) * a.src.pitch;
    for (int tx = lane; tx < kTileW + 2; tx += 32) {
      const int gx = sx0 + tx;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowOk && gx >= 0 && gx < a.src.w) c = fetch_texel<FIN>(row, gx);
      sC[ty * kRcasTW + tx] = c;
    }
  }
  __syncthreads();

  const uint32_t ggx = blockIdx.x * (kTileW / 16) + (warp & 3), ggy = blockIdx.y * (kTileH / 16) + (warp >> 2);
  const bool inside = group_inside(ggx * 16u + 8u, ggy * 16u + 8u, a.centre, a.radiusSq);
  const int x = (int)ggx * 16 + (lane & 15);
  if (x >= a.dst.w) return;
  const int tx = x - sx0;
#pragma unroll 2
  for (int k = 0; k < 8; ++k) {
    const int y = (int)ggy * 16 + (lane >> 4) + 2 * k;
    if (y >= a.dst.h) break;
    const float4 *p = sC + (y - sy0) * kRcasTW + tx;
    const float4 e = p[0];
    uint8_t *drow = a.dst.ptr + (size_t)y * a.dst.pitch;
    if (inside) {
      const float3 c = rcas_filter(p[-kRcasTW], p[-1], e, p[1], p[kRcasTW], a.sharp);
      store_texel<FOUT>(drow, x, c.x, c.y, c.z, 1.0f);
    } else {
      // OutputTexture[p] = mul * InputTexture[p], alpha included (fsr_rcas.hlsl:45-53)
      store_texel<FOUT>(drow, x, 1.0f * e.x, a.tintGB * e.y, a.tintGB * e.z, 1.0f * e.w);
    }
  }
}

} // inline namespace OVRFSR_MODE_NS
} // namespace ovrfsr
