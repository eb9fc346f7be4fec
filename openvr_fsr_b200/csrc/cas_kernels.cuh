// cas_kernels.cuh -- the legacy FidelityFX CAS path for sm_100a.
//
// Replaces the two shaders the reference keeps under src/cas but never dispatches:
//   cas.sharpen.hlsl (CAS_SHARPEN_ONLY + CAS_BETTER_DIAGONALS) and cas.upscale.hlsl, both = cas.compute.h:25-47 calling
//   CasFilter, src/cas/ffx_cas.h:409-893 (non-packed float path; bit-trick approximations of src/cas/ffx_a.h:1455-1457).
// With CAS_SLOW undefined only the GREEN amplitude chain feeds the filter weights (ffx_cas.h:514-523,869-878), so the
// red / blue chains of the source are dead code there and are not evaluated here.
//   sharpen: one CTA = 64x32 pixels; the 66x34 source box arrives by TMA (zero fill outside = CasLoad's Texture2D.Load)
//            and is decoded once; a lane walks 8 rows of one column and keeps the 3x3 window in registers.
//   upscale: one CTA = 64x32 output pixels; the source tile is decoded once and the per-SOURCE-TEXEL quantities of
//            the four neighbourhoods (weight w = APrxLoSqrt(amp)*peak and the "thin edge" reciprocal) are computed once
//            per texel into a second tile instead of four times per pixel; a pixel then reads 12 texels + 4 features.
// kStrict keeps the reference's operation order, so it is bit-identical to the header compiled on the host.
#pragma once

#include "fsr_kernels.cuh"

namespace ovrfsr {
inline namespace OVRFSR_MODE_NS {

struct CasArgs {
  ImageRO src;
  ImageRW dst;
  float c0x, c0y, c0z, c0w;  // const0: in/out scale and offset (CasSetup, ffx_cas.h:385-388)
  float peak;                // const1.x
  float maxColorDelta;       // const1.w (this fork's clamp, ffx_cas.h:546-551)
};

constexpr int kCasTW = 72, kCasTH = 36;  // upscale source tile (64x32 outputs, out->in step <= 1: 63+1+4, 31+1+4)
constexpr int kCasScaleSmem = kCasTW * kCasTH * (16 + 8);

__device__ __forceinline__ float cas_min3(float x, float y, float z) { return fminf(x, fminf(y, z)); } // AMin3F1
__device__ __forceinline__ float cas_max3(float x, float y, float z) { return fmaxf(x, fmaxf(y, z)); } // AMax3F1
__device__ __forceinline__ float prx_lo_sqrt(float a) { return u2f((f2u(a) >> 1) + 0x1fbc4639u); }    // APrxLoSqrtF1

// ffx_cas.h:424-551 on a 3x3 window (rows a b c / d e f / g h i)
__device__ __forceinline__ float3 cas_sharpen_filter(const float4 a, const float4 b, const float4 c, const float4 d, const float4 e,
                                                     const float4 f, const float4 g, const float4 h, const float4 i, float peak,
                                                     float mcd) {
  float mn = cas_min3(cas_min3(d.y, e.y, f.y), b.y, h.y);
  const float mnD = cas_min3(cas_min3(mn, a.y, c.y), g.y, i.y);
  mn = mn + mnD;
  float mx = cas_max3(cas_max3(d.y, e.y, f.y), b.y, h.y);
  const float mxD = cas_max3(cas_max3(mx, a.y, c.y), g.y, i.y);
  mx = mx + mxD;
  float amp = __saturatef(fminf(mn, 2.0f - mx) * prx_lo_rcp(mx));
  amp = prx_lo_sqrt(amp);
  const float w = amp * peak;
  const float rcpW = prx_med_rcp(1.0f + 4.0f * w);
  const float pR = __saturatef((b.x * w + d.x * w + f.x * w + h.x * w + e.x) * rcpW);
  const float pG = __saturatef((b.y * w + d.y * w + f.y * w + h.y * w + e.y) * rcpW);
  const float pB = __saturatef((b.z * w + d.z * w + f.z * w + h.z * w + e.z) * rcpW);
  return make_float3(fminf(fmaxf(pR, e.x - mcd), e.x + mcd), fminf(fmaxf(pG, e.y - mcd), e.y + mcd),
                     fminf(fmaxf(pB, e.z - mcd), e.z + mcd));
}

template <int FIN, int FOUT, bool TMA>
__global__ void __launch_bounds__(kThreads, 2) cas_sharpen_kernel(const __grid_constant__ CasArgs a,
                                                                  const __grid_constant__ CUtensorMap srcMap) {
  __shared__ __align__(128) float4 sC[kRcasTH * kRcasTW];
  __shared__ __align__(128) uint32_t sRaw[TMA ? kRcasTH * kRcasRawW : 1];
  __shared__ uint64_t tileBar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH;
  const int sx0 = ox0 - 1, sy0 = oy0 - 1;
  if constexpr (TMA) {
    if (tid == 0) {
      mbar_init(&tileBar, 1);
      fence_barrier_init();
      mbar_arrive_expect_tx(&tileBar, (uint32_t)(kRcasTH * kRcasRawW * 4));
      tma_load_2d(sRaw, &srcMap, ox0 - 4, sy0, &tileBar); // x origin 16-byte aligned: 3 unused texels on the left
    }
    __syncthreads(); // barrier init visible before anybody polls it
    mbar_wait(&tileBar, 0);
    for (int ty = warp; ty < kRcasTH; ty += kThreads / 32)
      for (int tx = lane; tx < kTileW + 2; tx += 32) sC[ty * kRcasTW + tx] = decode_rgba<FIN>(sRaw[ty * kRcasRawW + tx + 3]);
  } else {
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
  }
  __syncthreads();

  const int x = ox0 + (warp & 3) * 16 + (lane & 15);
  if (x >= a.dst.w) return;
  const int y0 = oy0 + (warp >> 2) * 16 + (lane >> 4) * 8; // this lane's 8 consecutive rows
  const float4 *p = sC + (y0 - sy0) * kRcasTW + (x - sx0);
  float4 ta = p[-kRcasTW - 1], tb = p[-kRcasTW], tc = p[-kRcasTW + 1], td = p[-1], te = p[0], tf = p[1];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int y = y0 + k;
    if (y >= a.dst.h) break;
    const float4 tg = p[kRcasTW - 1], th = p[kRcasTW], ti = p[kRcasTW + 1];
    const float3 c = cas_sharpen_filter(ta, tb, tc, td, te, tf, tg, th, ti, a.peak, a.maxColorDelta);
    store_opaque<FOUT>(a.dst.ptr + (size_t)y * a.dst.pitch, x, c.x, c.y, c.z);
    ta = td; tb = te; tc = tf; td = tg; te = th; tf = ti;
    p += kRcasTW;
  }
}

// per-source-texel part of one neighbourhood of the scaling branch (ffx_cas.h:610-706,708-826 for green, no
// CAS_BETTER_DIAGONALS): the '+' of up, left, centre, right, down
__device__ __forceinline__ float2 cas_texel_feature(float up, float left, float centre, float right, float down, float peak) {
  const float mn = cas_min3(cas_min3(up, left, centre), right, down);
  const float mx = cas_max3(cas_max3(up, left, centre), right, down);
  float amp = __saturatef(fminf(mn, 1.0f - mx) * prx_lo_rcp(mx));
  amp = prx_lo_sqrt(amp);
  const float thinB = 1.0f / 32.0f;
  return make_float2(amp * peak, prx_lo_rcp(thinB + (mx - mn)));
}

template <int FIN, int FOUT>
__global__ void __launch_bounds__(kThreads, 3) cas_scale_kernel(const __grid_constant__ CasArgs a) {
  extern __shared__ __align__(16) uint8_t cas_smem[];
  float4 *sC = reinterpret_cast<float4 *>(cas_smem);           // decoded colour
  float2 *sW = reinterpret_cast<float2 *>(sC + kCasTW * kCasTH); // (w, thin reciprocal) of the texel's neighbourhood
  const int tid = threadIdx.x;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH;
  // tile origin: one texel left / above the 'f' texel of the tile's first pixel (ffx_cas.h:569-572)
  const int sx0 = (int)floorf(mul_add_unfused((float)(uint32_t)ox0, a.c0x, a.c0z)) - 1;
  const int sy0 = (int)floorf(mul_add_unfused((float)(uint32_t)oy0, a.c0y, a.c0w)) - 1;

  for (int q = tid; q < kCasTW * kCasTH; q += kThreads) {
    const int ty = q / kCasTW, tx = q - ty * kCasTW;
    const int gx = sx0 + tx, gy = sy0 + ty;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f); // CasLoad = Texture2D.Load: zero outside the image
    if (gx >= 0 && gy >= 0 && gx < a.src.w && gy < a.src.h) c = fetch_texel<FIN>(a.src.ptr + (size_t)gy * a.src.pitch, gx);
    sC[q] = c;
  }
  __syncthreads();
  for (int q = tid; q < (kCasTW - 2) * (kCasTH - 2); q += kThreads) {
    const int ty = 1 + q / (kCasTW - 2), tx = 1 + q % (kCasTW - 2);
    const float4 *c = sC + ty * kCasTW + tx;
    sW[ty * kCasTW + tx] = cas_texel_feature(c[-kCasTW].y, c[-1].y, c[0].y, c[1].y, c[kCasTW].y, a.peak);
  }
  __syncthreads();

  const int x = ox0 + (tid & 63);
  if (x >= a.dst.w) return;
  float ppx = mul_add_unfused((float)(uint32_t)x, a.c0x, a.c0z);
  const float fx = floorf(ppx);
  ppx -= fx;
  const int ix = (int)fx - sx0;
#pragma unroll 2
  for (int k = 0; k < 8; ++k) {
    const int y = oy0 + (tid >> 6) + 4 * k;
    if (y >= a.dst.h) break;
    float ppy = mul_add_unfused((float)(uint32_t)y, a.c0y, a.c0w);
    const float fy = floorf(ppy);
    ppy -= fy;
    const float4 *r = sC + ((int)fy - sy0) * kCasTW + ix; // 'f'
    const float2 *wr = sW + ((int)fy - sy0) * kCasTW + ix;
    //    b c
    //  e f g h
    //  i j k l
    //    n o
    const float4 b = r[-kCasTW], c = r[-kCasTW + 1];
    const float4 e = r[-1], f = r[0], g = r[1], h = r[2];
    const float4 i = r[kCasTW - 1], j = r[kCasTW], kk = r[kCasTW + 1], l = r[kCasTW + 2];
    const float4 n = r[2 * kCasTW], o = r[2 * kCasTW + 1];
    const float2 Ff = wr[0], Fg = wr[1], Fj = wr[kCasTW], Fk = wr[kCasTW + 1];
    float s = (1.0f - ppx) * (1.0f - ppy), t = ppx * (1.0f - ppy), u = (1.0f - ppx) * ppy, v = ppx * ppy;
    s *= Ff.y; t *= Fg.y; u *= Fj.y; v *= Fk.y;
    const float wf = Ff.x, wg = Fg.x, wj = Fj.x, wk = Fk.x;
    const float qbe = wf * s, qch = wg * t;
    const float qf = wg * t + wj * u + s, qg = wf * s + wk * v + t, qj = wf * s + wk * v + u, qk = wg * t + wj * u + v;
    const float qin = wj * u, qlo = wk * v;
    const float rcpW = prx_med_rcp(2.0f * qbe + 2.0f * qch + 2.0f * qin + 2.0f * qlo + qf + qg + qj + qk);
    const float pR = __saturatef((b.x * qbe + e.x * qbe + c.x * qch + h.x * qch + i.x * qin + n.x * qin + l.x * qlo + o.x * qlo +
                                  f.x * qf + g.x * qg + j.x * qj + kk.x * qk) * rcpW);
    const float pG = __saturatef((b.y * qbe + e.y * qbe + c.y * qch + h.y * qch + i.y * qin + n.y * qin + l.y * qlo + o.y * qlo +
                                  f.y * qf + g.y * qg + j.y * qj + kk.y * qk) * rcpW);
    const float pB = __saturatef((b.z * qbe + e.z * qbe + c.z * qch + h.z * qch + i.z * qin + n.z * qin + l.z * qlo + o.z * qlo +
                                  f.z * qf + g.z * qg + j.z * qj + kk.z * qk) * rcpW);
    store_opaque<FOUT>(a.dst.ptr + (size_t)y * a.dst.pitch, x, pR, pG, pB);
  }
}

} // inline namespace OVRFSR_MODE_NS
} // namespace ovrfsr
