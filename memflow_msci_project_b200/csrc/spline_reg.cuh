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

// Phase 1: knots and bin.  lw / lh hold the K width / height logits on entry and are overwritten (soft clip ->
// softmax numerators).  v is pre-mapped (circular shift / PS map) in place for the forward direction.
template <int NCH>
__device__ __forceinline__ void rqs_scan_reg(float* lw, float* lh, const SplineCfg<float>& cfg, bool inverse, float& v,
                                             int& k, bool& inside, float& x0, float& x1, float& y0, float& y1,
                                             float& ladj0) {
  constexpr int KP = 8 * NCH;
  const int K = cfg.K;
  ladj0 = 0.f;
  if (!inverse) {
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) v = circular_wrap(v, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) { v = (v + 1.f) * 0.5f; ladj0 = -0.6931471805599453f; }
  }
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    lw[j] = softclipf(lw[j], cfg.inv_ls2);
    lh[j] = softclipf(lh[j], cfg.inv_ls2);
  }
  float mw = 0.f, mh = 0.f;
  if (cfg.inv_ls2 == 0.f) {  // no soft clip: logits are unbounded, subtract the maximum
    mw = -INFINITY; mh = -INFINITY;
#pragma unroll
    for (int j = 0; j < KP; ++j)
      if (j < K) { mw = fmaxf(mw, lw[j]); mh = fmaxf(mh, lh[j]); }
  }
  float sw = 0.f, sh = 0.f;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    // padding entries (j >= K) go through the same straight-line code with an argument that underflows to 0:
    // a conditional around the inline-asm exp would be compiled as one branch per element
    lw[j] = exp_sfu((j < K) ? lw[j] - mw : -200.f);
    lh[j] = exp_sfu((j < K) ? lh[j] - mh : -200.f);
    sw += lw[j];
    sh += lh[j];
  }
  const float isw = __fdividef(1.f, sw), ish = __fdividef(1.f, sh);
  k = -1;
  x0 = 0.f; x1 = 1.f; y0 = 0.f; y1 = 1.f;
  // branch-free scan (selects only): with a runtime K the padded entries add 0 to the running sums and are
  // masked out of the bin search, so the loop stays one basic block the scheduler can interleave
  float cw = 0.f, ch = 0.f, prevw = -cfg.bound, prevh = -cfg.bound;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    cw += lw[j] * isw;
    ch += lh[j] * ish;
    const float knotw = cfg.bound * (2.f * cw - 1.f), knoth = cfg.bound * (2.f * ch - 1.f);
    const bool take = (j < K) & ((inverse ? prevh : prevw) < v);  // k = #(knots < v) - 1 on the searched axis
    k = take ? j : k;
    x0 = take ? prevw : x0; x1 = take ? knotw : x1;
    y0 = take ? prevh : y0; y1 = take ? knoth : y1;
    prevw = (j < K) ? knotw : prevw;
    prevh = (j < K) ? knoth : prevh;
  }
  if ((inverse ? prevh : prevw) < v) k = K;
  inside = (k >= 0) && (k < K);
  k = k < 0 ? k + K : (k >= K ? k - K : k);
  if (!inside) { x0 = 0.f; x1 = 1.f; y0 = 0.f; y1 = 1.f; }
}

// Phase 2: the rational-quadratic map of bin k given the derivative logits t0 (knot k) and t1 (knot k+1).
__device__ __forceinline__ void rqs_finish(const SplineCfg<float>& cfg, bool inverse, float v, int k, bool inside,
                                           float x0, float x1, float y0, float y1, float t0, float t1, float ladj0,
                                           float& out, float& ladj) {
  const int K = cfg.K;
  // boundary derivatives are 1 = exp(0): select the argument, not the result (keeps the inline-asm exp unconditional)
  const float d0 = exp_sfu((k > 0) ? softclipf(t0, cfg.inv_ls1) : 0.f);
  const float d1 = exp_sfu((k < K - 1) ? softclipf(t1, cfg.inv_ls1) : 0.f);
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

// pick logits k-1 and k out of a register array without dynamic indexing
template <int KP>
__device__ __forceinline__ void select_pair(const float* ld, int k, float& t0, float& t1) {
  t0 = 0.f; t1 = 0.f;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    t0 = (j == k - 1) ? ld[j] : t0;
    t1 = (j == k) ? ld[j] : t1;
  }
}

// lw, lh: K width / height logits (overwritten); ld: K-1 derivative logits (entries beyond are ignored).
template <int NCH>
__device__ __forceinline__ void rqs_eval_reg(float* lw, float* lh, const float* ld, const SplineCfg<float>& cfg,
                                             bool inverse, float v, float& out, float& ladj, int& bin, bool& inside) {
  int k;
  float x0, x1, y0, y1, ladj0, t0, t1;
  rqs_scan_reg<NCH>(lw, lh, cfg, inverse, v, k, inside, x0, x1, y0, y1, ladj0);
  select_pair<8 * NCH>(ld, k, t0, t1);
  bin = k;
  rqs_finish(cfg, inverse, v, k, inside, x0, x1, y0, y1, t0, t1, ladj0, out, ladj);
}

}  // namespace mf
