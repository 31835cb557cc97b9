// spline_reg.cuh -- monotonic rational-quadratic spline of ONE scalar with its 3K-1 logits in registers
// (K <= 8 * NCH <= 40), fp32, written for the issue-slot budget of an SM sub-partition.
// Semantics: zuko MonotonicRQSTransform.call_and_ladj / _inverse (+ the circular / PS pre-maps of the flows);
// restated in oracle/flow_oracle.py (rqs_knots, rqs_forward, rqs_inverse).
//
// Formulation ("masked sums", no per-bin selects, no normalised knots):
//   e_j  = exp(softclip(w_j)),  f_j = exp(softclip(h_j))          unnormalised softmax numerators
//   S    = sum e_j, cw_j = e_0 + .. + e_j                          (same summation order as a cumsum)
//   the searched knot comparison  bound (2 cw_j / S - 1) < v  becomes  cw_j < vt,  vt = (v / (2 bound) + 1/2) S
//   m_j  = [cw_j < vt] in {0, 1},  m_{-1} = [-bound < v],  o_j = m_{j-1} - m_j  (one-hot of the bin k)
//   X0 = sum e_j m_j (= cw_{k-1}),  E = sum e_j o_j (= e_k),  Y0 = sum f_j m_j,  F = sum f_j o_j,
//   derivative logits of the two knots  t0 = sum d_{j-1} o_j,  t1 = sum d_j o_j  (0 at the boundaries -> exp = 1)
//   z = (vt - X0) / E,  s = (F / Sh) / (E / S),  y0 = bound (2 Y0 / Sh - 1),  dy = 2 bound F / Sh.
// Nothing is normalised before it is needed, the bin width E is the numerator itself (no x1 - x0 cancellation), and
// every step is a multiply-add that packs two lanes into one FFMA2 / FADD2 / FMUL2 (sm_100 packed fp32).
// Measured against the fp64 oracle the error distribution is the one of the numpy-fp32 oracle or better
// (tools/proto_spline_masked_sums.py).
//
// Logits arrive PRE-SCALED: lam = logit * inv_ls (inv_ls2 = 2 / ln(slope) for widths / heights, inv_ls1 = 1 / ln(slope)
// for derivatives), which the callers get for free from one multiply-add (accumulator * inv + bias * inv).  Then
//   exp(softclip(logit)) = ex2( lam / (1 + |lam|) * (log2(e) / inv_ls) ).
// The soft clip bounds the exponent by |ln(slope)| / 2, so no maximum has to be subtracted.  Flows built with
// slope = None (no soft clip) do not take this path (callers fall back to rqs_univariate in flow_common.cuh).
#pragma once
#include "flow_common.cuh"

namespace mf {

// ---- packed fp32 pairs (FFMA2 / FADD2 / FMUL2) ---------------------------------------------------
__device__ __forceinline__ unsigned long long f2_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long r) {
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
  return d;
}
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ float2 f2_bcast(float s) { return make_float2(s, s); }

__device__ __forceinline__ float ex2_sfu(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}
__device__ __forceinline__ float rcp_sfu(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// 1.0f / 0.0f without a predicate round trip (FSET.BF)
__device__ __forceinline__ float lt_as_float(float a, float b) {
  float r;
  asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

__device__ __forceinline__ float softclipf(float v, float inv) { return __fdividef(v, 1.f + fabsf(v * inv)); }

// exp(x) on the SFU for the derivative knots
__device__ __forceinline__ float exp_sfu(float x) { return ex2_sfu(x * 1.4426950408889634f); }

// per-kernel constants of the scaled-logit spline
struct RqsScaled {
  int K;
  int univariate;
  float bound;
  float inv2, inv1;  // 2 / ln(slope), 1 / ln(slope): the scale factors of the incoming logits
  float cl2, cl1;    // log2(e) / inv2, log2(e) / inv1
  float half_ib;     // 1 / (2 bound)
};

__device__ __forceinline__ RqsScaled make_rqs_scaled(int K, int univariate, double bound, double slope) {
  RqsScaled c;
  c.K = K; c.univariate = univariate; c.bound = (float)bound;
  const double ls = log(slope);
  c.inv2 = (float)(2.0 / ls); c.inv1 = (float)(1.0 / ls);
  c.cl2 = (float)(1.4426950408889634 * ls / 2.0); c.cl1 = (float)(1.4426950408889634 * ls);
  c.half_ib = (float)(0.5 / bound);
  return c;
}

// ---- step 1: softmax numerators -----------------------------------------------------------------
// lw, lh: KP = 8 NCH scaled width / height logits (entries >= K are ignored), overwritten with e_j / f_j (0 beyond K).
template <int NCH>
__device__ __forceinline__ void rqs_ms_exp(float* lw, float* lh, const RqsScaled& cf, float& sw, float& sh) {
  constexpr int KP = 8 * NCH;
  const int K = cf.K;
  const float2 cl = f2_bcast(cf.cl2);
  float2 accw = make_float2(0.f, 0.f), acch = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < KP; j += 2) {
    const float2 pw = make_float2(lw[j], lw[j + 1]), ph = make_float2(lh[j], lh[j + 1]);
    const float2 rw = make_float2(rcp_sfu(1.f + fabsf(pw.x)), rcp_sfu(1.f + fabsf(pw.y)));
    const float2 rh = make_float2(rcp_sfu(1.f + fabsf(ph.x)), rcp_sfu(1.f + fabsf(ph.y)));
    float2 aw = f2_mul(f2_mul(pw, cl), rw), ah = f2_mul(f2_mul(ph, cl), rh);
    if (j >= KP - 8) {  // only the last chunk of eight can hold padding (KP = roundup(K, 8)): exponent -200 -> 0
      aw.x = (j < K) ? aw.x : -200.f; aw.y = (j + 1 < K) ? aw.y : -200.f;
      ah.x = (j < K) ? ah.x : -200.f; ah.y = (j + 1 < K) ? ah.y : -200.f;
    }
    const float2 e = make_float2(ex2_sfu(aw.x), ex2_sfu(aw.y)), f = make_float2(ex2_sfu(ah.x), ex2_sfu(ah.y));
    lw[j] = e.x; lw[j + 1] = e.y; lh[j] = f.x; lh[j + 1] = f.y;
    accw = f2_add(accw, e); acch = f2_add(acch, f);
  }
  sw = accw.x + accw.y;
  sh = acch.x + acch.y;
}

// ---- step 2: masked sums, eight entries at a time -----------------------------------------------
struct RqsMsState {
  float vt, m_first, m_prev, cs, t0, kf, d_prev;
  float2 x0, y0, ek, fk, t1;
};

template <bool INV>
__device__ __forceinline__ void rqs_ms_begin(RqsMsState& st, const RqsScaled& cf, float v, float sw, float sh) {
  const float u = fmaf(v, cf.half_ib, 0.5f);
  st.vt = u * (INV ? sh : sw);
  st.m_first = lt_as_float(-cf.bound, v);
  st.m_prev = st.m_first;
  st.cs = 0.f; st.t0 = 0.f; st.kf = 0.f; st.d_prev = 0.f;
  st.x0 = st.y0 = st.ek = st.fk = st.t1 = make_float2(0.f, 0.f);
}

// e8, f8: numerators j0 .. j0+7;  d8: scaled derivative logits j0 .. j0+7 (entries >= K-1 MUST be zero: they stand
// for the boundary derivative exp(0) = 1)
template <bool INV>
__device__ __forceinline__ void rqs_ms_chunk(RqsMsState& st, const float* e8, const float* f8, const float* d8, int j0) {
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    const float2 e = make_float2(e8[i], e8[i + 1]), f = make_float2(f8[i], f8[i + 1]);
    st.cs += INV ? f.x : e.x;
    const float ma = lt_as_float(st.cs, st.vt);
    st.cs += INV ? f.y : e.y;
    const float mb = lt_as_float(st.cs, st.vt);
    const float2 m = make_float2(ma, mb), o = make_float2(st.m_prev - ma, ma - mb);
    st.x0 = f2_fma(e, m, st.x0); st.y0 = f2_fma(f, m, st.y0);
    st.ek = f2_fma(e, o, st.ek); st.fk = f2_fma(f, o, st.fk);
    st.t1 = f2_fma(make_float2(d8[i], d8[i + 1]), o, st.t1);
    st.t0 = fmaf(st.d_prev, o.x, st.t0);
    st.t0 = fmaf(d8[i], o.y, st.t0);
    st.kf = fmaf((float)(j0 + i), o.x, st.kf);
    st.kf = fmaf((float)(j0 + i + 1), o.y, st.kf);
    st.d_prev = d8[i + 1];
    st.m_prev = mb;
  }
}

// ---- step 3: the rational-quadratic map of the bin ------------------------------------------------
// INV = false: out = f(v), ladj = log f'(v);  INV = true: out = f^-1(v), ladj = 0.
template <bool INV>
__device__ __forceinline__ void rqs_ms_finish(const RqsMsState& st, const RqsScaled& cf, float v, float sw, float sh,
                                              float ladj0, float& out, float& ladj, int& bin, bool& inside) {
  const float t1 = st.t1.x + st.t1.y, t0 = st.t0;
  const float X0 = st.x0.x + st.x0.y, Y0 = st.y0.x + st.y0.y, E = st.ek.x + st.ek.y, F = st.fk.x + st.fk.y;
  inside = (st.m_first != 0.f) && (st.m_prev == 0.f);
  // zuko: k = searchsorted - 1, then k mod K: K -> 0 above the last knot (kf = 0 there), -1 -> K - 1 below the first
  bin = (st.m_first != 0.f) ? (int)st.kf : cf.K - 1;
  const float d0 = ex2_sfu(t0 * cf.cl1 * rcp_sfu(1.f + fabsf(t0)));
  const float d1 = ex2_sfu(t1 * cf.cl1 * rcp_sfu(1.f + fabsf(t1)));
  const float ish = rcp_sfu(sh), iE = rcp_sfu(E);
  const float s = (F * ish) * (sw * iE);
  const float e2 = d0 + d1 - 2.f * s;
  if (!INV) {
    const float z = inside ? (st.vt - X0) * iE : 0.f;
    const float omz = 1.f - z;
    const float den = fmaf(e2, z * omz, s);
    const float iden = rcp_sfu(den);
    const float yb = fmaf(2.f * Y0, ish, -1.f) * cf.bound;  // y0
    const float dy = 2.f * cf.bound * (F * ish);
    const float y = fmaf(dy * (s * z * z + d0 * z * omz), iden, yb);
    const float jac = s * s * (2.f * s * z * omz + d0 * omz * omz + d1 * z * z) * iden * iden;
    out = inside ? y : v;
    ladj = (inside ? logf(jac) : 0.f) + ladj0;
  } else {
    const float zeta = inside ? (st.vt - Y0) * rcp_sfu(F) : 0.f;  // (y - y0) / (y1 - y0)
    const float a = (s - d0) + zeta * e2;
    const float b = d0 - zeta * e2;
    const float c = -s * zeta;
    const float z = 2.f * c / (-b - sqrtf(b * b - 4.f * a * c));
    float x = fmaf(2.f * fmaf(z, E, X0), rcp_sfu(sw), -1.f) * cf.bound;
    x = inside ? x : v;
    if (cf.univariate == MF_UNIV_CIRCULAR_RQS) x = circular_wrap(x, cf.bound);
    if (cf.univariate == MF_UNIV_PS_RQS) x = 2.f * x - 1.f;
    out = x;
    ladj = 0.f;
  }
}

// forward pre-map of the flows' composed univariates (circular shift / PS map)
__device__ __forceinline__ float rqs_premap(const RqsScaled& cf, float v, float& ladj0) {
  ladj0 = 0.f;
  if (cf.univariate == MF_UNIV_CIRCULAR_RQS) v = circular_wrap(v, cf.bound);
  if (cf.univariate == MF_UNIV_PS_RQS) { v = (v + 1.f) * 0.5f; ladj0 = -0.6931471805599453f; }
  return v;
}

// all-in-registers form: lw, lh, ld hold the KP scaled logits (ld entries >= K-1 zero); lw / lh are overwritten
template <int NCH, bool INV>
__device__ __forceinline__ void rqs_masked_sums(float* lw, float* lh, const float* ld, const RqsScaled& cf, float v,
                                                float& out, float& ladj, int& bin, bool& inside) {
  float ladj0 = 0.f;
  if (!INV) v = rqs_premap(cf, v, ladj0);
  float sw, sh;
  rqs_ms_exp<NCH>(lw, lh, cf, sw, sh);
  RqsMsState st;
  rqs_ms_begin<INV>(st, cf, v, sw, sh);
#pragma unroll
  for (int c = 0; c < NCH; ++c) rqs_ms_chunk<INV>(st, lw + 8 * c, lh + 8 * c, ld + 8 * c, 8 * c);
  rqs_ms_finish<INV>(st, cf, v, sw, sh, ladj0, out, ladj, bin, inside);
}

}  // namespace mf
