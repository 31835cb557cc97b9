// spline_reg.cuh -- monotonic rational-quadratic spline of ONE scalar with all 3K-1 logits in registers
// (K <= 8 * NCH <= 32).  fp32, SFU transcendentals (ex2 / rcp, ~2 ulp): every logit is soft-clipped and
// exponentiated once, the horizontal and vertical knot scans run in one pass.  Shared by the tcgen05 flow
// kernel (logits come out of TMEM) and the standalone spline kernel (logits come out of shared memory).
// Semantics: zuko MonotonicRQSTransform.call_and_ladj / _inverse (+ the circular / PS pre-maps of the flows).
#pragma once
#include "flow_common.cuh"

namespace mf {

__device__ __forceinline__ float softclipf(float v, float inv) { return __fdividef(v, 1.f + fabsf(v * inv)); }

// exp(x) on the SFU (ex2.approx, 2 ulp) with the rounding of x * log2(e) compensated: ~2^-22 relative for the
// bounded arguments of the soft-clipped softmax, at 5 instructions instead of expf's ~15
__device__ __forceinline__ float exp_sfu(float x) {
  const float kL2E = 1.4426950216293335f, kL2E_lo = 1.925963033500107e-08f;
  const float t = x * kL2E;
  float r = fmaf(x, kL2E, -t);
  r = fmaf(x, kL2E_lo, r);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return fmaf(e, r * 0.6931471805599453f, e);
}

// lw, lh: K width / height logits; ld: K-1 derivative logits (entries beyond are ignored).
template <int NCH>
__device__ __forceinline__ void rqs_eval_reg(const float* lw, const float* lh, const float* ld,
                                             const SplineCfg<float>& cfg, bool inverse, float v, float& out,
                                             float& ladj, int& bin, bool& inside) {
  constexpr int KP = 8 * NCH;
  const int K = cfg.K;
  float ladj0 = 0.f;
  if (!inverse) {
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) v = circular_wrap(v, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) { v = (v + 1.f) * 0.5f; ladj0 = -0.6931471805599453f; }
  }
  float ew[KP], eh[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    ew[j] = softclipf(lw[j], cfg.inv_ls2);
    eh[j] = softclipf(lh[j], cfg.inv_ls2);
  }
  float mw = 0.f, mh = 0.f;
  if (cfg.inv_ls2 == 0.f) {  // no soft clip: logits are unbounded, subtract the maximum
    mw = -INFINITY; mh = -INFINITY;
#pragma unroll
    for (int j = 0; j < KP; ++j)
      if (j < K) { mw = fmaxf(mw, ew[j]); mh = fmaxf(mh, eh[j]); }
  }
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    ew[j] = (j < K) ? exp_sfu(ew[j] - mw) : 0.f;
    eh[j] = (j < K) ? exp_sfu(eh[j] - mh) : 0.f;
    sw += ew[j];
    sh += eh[j];
  }
  const float isw = __fdividef(1.f, sw), ish = __fdividef(1.f, sh);
  int k = -1;
  float x0 = 0.f, x1 = 1.f, y0 = 0.f, y1 = 1.f, cw = 0.f, ch = 0.f, prevw = -cfg.bound, prevh = -cfg.bound;
  float t0 = 0.f, t1 = 0.f;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    if (j < K) {
      cw += ew[j] * isw;
      ch += eh[j] * ish;
      const float knotw = cfg.bound * (2.f * cw - 1.f), knoth = cfg.bound * (2.f * ch - 1.f);
      if ((inverse ? prevh : prevw) < v) {  // k = #(knots < v) - 1 on the searched axis
        k = j; x0 = prevw; x1 = knotw; y0 = prevh; y1 = knoth;
        t1 = ld[j];                      // derivative logit k   (knot k+1)
        t0 = ld[j > 0 ? j - 1 : 0];      // derivative logit k-1 (knot k)
      }
      prevw = knotw; prevh = knoth;
    }
  }
  if ((inverse ? prevh : prevw) < v) k = K;
  inside = (k >= 0) && (k < K);
  k = k < 0 ? k + K : (k >= K ? k - K : k);
  bin = k;
  const float d0 = (k > 0) ? exp_sfu(softclipf(t0, cfg.inv_ls1)) : 1.f;
  const float d1 = (k < K - 1) ? exp_sfu(softclipf(t1, cfg.inv_ls1)) : 1.f;
  if (!inside) { x0 = 0.f; x1 = 1.f; y0 = 0.f; y1 = 1.f; }
  const float dx = x1 - x0, dy = y1 - y0;
  const float s = dy * __fdividef(1.f, dx);
  if (!inverse) {
    const float z = inside ? (v - x0) * __fdividef(1.f, dx) : 0.f;
    const float omz = 1.f - z;
    const float den = s + (d0 + d1 - 2.f * s) * z * omz;
    const float iden = __fdividef(1.f, den);
    const float y = y0 + dy * (s * z * z + d0 * z * omz) * iden;
    const float jac = s * s * (2.f * s * z * omz + d0 * omz * omz + d1 * z * z) * iden * iden;
    out = inside ? y : v;
    ladj = (inside ? logf(jac) : 0.f) + ladj0;
  } else {
    const float y_ = inside ? (v - y0) : 0.f;
    const float a = dy * (s - d0) + y_ * (d0 + d1 - 2.f * s);
    const float b = dy * d0 - y_ * (d0 + d1 - 2.f * s);
    const float c = -s * y_;
    const float z = 2.f * c / (-b - sqrtf(b * b - 4.f * a * c));
    float x = x0 + z * dx;
    x = inside ? x : v;
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) x = circular_wrap(x, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) x = 2.f * x - 1.f;
    out = x;
    ladj = 0.f;
  }
}

}  // namespace mf
