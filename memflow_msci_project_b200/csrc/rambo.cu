// rambo.cu -- fused RAMBO-on-diet kernels for sm_100a.
//
// One launch replaces the thousands of ATen element-wise ops behind
//   FlatInvertiblePhasespace.generateKinematics_batch  (memflow/phasespace/rambo_generator.py:168-398)
//   FlatInvertiblePhasespace.getPSpoint_batch          (memflow/phasespace/rambo_generator.py:615-736)
//   Compute_ParticlesTensor.get_PS                     (memflow/unfolding_flow/utils.py:1056-1152)
//
// Layout: the API keeps the reference's array-of-structures tensors (r [N,3n-2], momenta
// [N,n+2,4]).  A CTA owns a tile of EPB consecutive events, i.e. ONE contiguous byte range of
// every tensor: the tile is staged global<->shared with a single 1-D bulk async copy (TMA engine,
// SASS UBLKCP) or, for ragged/unaligned tiles, with coalesced 16-byte loads; each thread then
// works on one event out of shared memory with everything else in registers (one warp = 32
// consecutive events).  Per-event scalars (weight, x1, x2, ...) are written straight from
// registers, which is coalesced by construction.
//
// Arithmetic: TC=double follows the reference (which computes in fp64) and is the parity mode.
// TC=float is a cancellation-free reformulation for throughput (see DESIGN.md for its error
// budget).  The 120-step bisection of the reference (:400-446) is replaced by a bracketed Newton
// iteration on the same equation, which converges to the same fp64 root.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace mf {
namespace {

constexpr int kEPB = 128;  // events per CTA (one per thread)
__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
struct Mth;
// fp64 sqrt / reciprocal for the parity mode: fp32 SFU seed + Newton steps in fp64 (faithfully rounded,
// ~1 ulp) instead of the IEEE sequences with their slow-path calls.  Arguments outside the fp32 exponent
// range (never produced by physical kinematics) take the library routine.
__device__ __forceinline__ double fast_sqrt(double x) {
  const float xf = (float)x;
  if (!(xf > 1e-30f && xf < 1e30f)) return sqrt(x);  // 0, negative (NaN like the reference), NaN, huge
  double y = (double)rsqrtf(xf);
  y = y * fma(-0.5 * x, y * y, 1.5);
  double sq = x * y;
  return fma(fma(-sq, sq, x), 0.5 * y, sq);
}
__device__ __forceinline__ double fast_rcp(double x) {
  const float xf = (float)x;
  if (!(fabsf(xf) > 1e-30f && fabsf(xf) < 1e30f)) return 1.0 / x;
  double y = (double)__fdividef(1.f, xf);
  y = y * fma(-x, y, 2.0);
  y = y * fma(-x, y, 2.0);
  return y;
}
__device__ __forceinline__ double fast_div(double a, double b) {
  const double y = fast_rcp(b);
  const double q = a * y;
  return fma(fma(-q, b, a), y, q);
}
__device__ __forceinline__ float fast_div(float a, float b) { return a / b; }

template <>
struct Mth<double> {
  static __device__ __forceinline__ double sqrt_(double x) { return fast_sqrt(x); }
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double pow_(double x, double y) { return pow(x, y); }
  static __device__ __forceinline__ void sincospi_(double x, double* s, double* c) { sincospi(x, s, c); }
  static __device__ __forceinline__ double atan_(double x) { return atan(x); }
};
template <>
struct Mth<float> {
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __device__ __forceinline__ float log_(float x) { return logf(x); }
  static __device__ __forceinline__ float pow_(float x, float y) { return powf(x, y); }
  static __device__ __forceinline__ void sincospi_(float x, float* s, float* c) { sincospif(x, s, c); }
  static __device__ __forceinline__ float atan_(float x) { return atanf(x); }
};

template <typename T>
__device__ __forceinline__ T ipow(T x, int e) {
  T r = T(1);
  for (int i = 0; i < e; ++i) r *= x;
  return r;
}

// Root u in [0,1] of v = (e+1) u^e - e u^(e+1)   (massless_map, rambo_generator.py:142-143;
// solved there by bisect_vec_batch, :400-446).  Bracketed Newton from the asymptotic seeds
//   u ~ (v/(e+1))^(1/e)  (v -> 0),   1-u ~ sqrt(2(1-v)/(e(e+1)))  (v -> 1).
__device__ __forceinline__ float solve_massless_u_f32(float v, int e) {
  const float fe = (float)e;
  const float vc = 2.f * ipow((fe - 1.f) / fe, e);  // value at the inflection point u = (e-1)/e
  float u;
  if (v < vc)
    u = (e == 2) ? sqrtf(v * (1.f / 3.f)) : powf(v / (fe + 1.f), 1.f / fe);
  else
    u = 1.f - sqrtf(2.f * (1.f - v) / (fe * (fe + 1.f)));
  float lo = 0.f, hi = 1.f;
  const int iters = e <= 8 ? 4 : 12;
  for (int it = 0; it < iters; ++it) {
    const float ue1 = ipow(u, e - 1);
    const float g = ue1 * u * ((fe + 1.f) - fe * u) - v;
    const float gp = fe * (fe + 1.f) * ue1 * (1.f - u);
    if (g < 0.f) lo = u;
    if (g > 0.f) hi = u;
    const float un = u - g / gp;
    u = (un >= lo && un <= hi) ? un : 0.5f * (lo + hi);
  }
  return u;
}

template <typename T>
__device__ __forceinline__ T solve_massless_u(T v, int e) {
  if (!(v > T(0))) return T(0);  // the bisection walks to the left/right end for v outside (0,1)
  if (!(v < T(1))) return T(1);
  if (e == 1) {                  // v = 2u - u^2
    T s = Mth<T>::sqrt_(T(1) - v);
    return fast_div(v, T(1) + s);
  }
  float uf = solve_massless_u_f32((float)v, e);
  if (sizeof(T) == 4) return (T)uf;
  // fp64: polish the fp32 root with bracketed Newton steps (quadratic: 1e-4 -> 1e-8 -> 1e-16)
  const T fe = T(e);
  uf = fminf(fmaxf(uf, 1e-30f), 1.f);
  T u = (T)uf, lo = T(0), hi = T(1);
  const int iters = e <= 8 ? 3 : 6;
  for (int it = 0; it < iters; ++it) {
    T ue1 = ipow(u, e - 1);
    T g = ue1 * u * ((fe + T(1)) - fe * u) - v;
    T gp = fe * (fe + T(1)) * ue1 * (T(1) - u);
    if (g < T(0)) lo = u;
    if (g > T(0)) hi = u;
    T un = u - fast_div(g, gp);
    u = (un >= lo && un <= hi) ? un : T(0.5) * (lo + hi);
  }
  return u;
}

// host-derived extras that are not part of the public descriptor
struct RamboAux {
  double e_thr;      // sqrt_s * exp(-q/2): E_cm at u_b = 1 as the (fp32-valued) q defines it
  double thr_gap;    // e_thr - sum_mass
  double dS[MF_MAX_FINAL];  // tail_sum_fwd[i] - tail_sum_fwd[i+1] - mass[i]  (fp32 cumsum residue)
  double inv_A, inv_c2;     // 1/x_A, 1/x_c2
  double vol_scale;         // vol_coef / vol_fact
  int cuts_on;
  int perm[MF_MAX_FINAL];
  // lab-frame get_PS (mf_rambo_get_ps_lab, inverse kernel mode 2)
  double m_min_tot;      // boost fix-up threshold (<= 0: no fix-up)
  void* lab_mom_cm;      // optional output: momenta boosted to the CM frame [N,n,4]
  void* lab_boost_out;   // optional output: the boost four-vectors after the fix-up [N,4]
};

__device__ __forceinline__ double pseudo_rap_d(double px, double py, double pz) {
  const double eps = 1.4901161193847656e-08, huge = 1.7976931348623157e308;
  double pt = sqrt(px * px + py * py);
  if (pt < eps && fabs(pz) < eps) return huge;
  return -log(tan(atan2(pt, pz) * 0.5));
}

// lab-frame cuts of the reference (rambo_generator.py:350-393, phasespace/utils.py:180-257).
// row = CM momenta of one event in shared memory.  Returns the 0/1 factor.
template <typename TIO>
__device__ __noinline__ double rambo_cut_factor(const mf_rambo_desc& d, int n, const TIO* row,
                                                double x1, double x2) {
  double bx = 0, by = 0, bz = 0;
  bool do_boost = false;
  {
    double r0 = (double)row[0] * x1 + (double)row[4] * x2;
    double r1 = (double)row[1] * x1 + (double)row[5] * x2;
    double r2 = (double)row[2] * x1 + (double)row[6] * x2;
    double r3 = (double)row[3] * x1 + (double)row[7] * x2;
    if ((r1 * r1 + r2 * r2 + r3 * r3) != 0.0 && (x1 != 1.0 || x2 != 1.0)) {
      do_boost = true;
      bx = r1 / r0; by = r2 / r0; bz = r3 / r0;
    }
  }
  const double b2 = bx * bx + by * by + bz * bz;
  const double gamma = 1.0 / sqrt(1.0 - b2);
  const double gamma2 = b2 > 0.0 ? (gamma - 1.0) / b2 : 0.0;
  const double huge = 1.7976931348623157e308;
  double factor = 1.0, ptmin = INFINITY, etamax = -INFINITY;
  double lx[MF_MAX_FINAL], ly[MF_MAX_FINAL], lz[MF_MAX_FINAL];
  for (int i = 0; i < n; ++i) {
    double E = row[4 * (i + 2)], px = row[4 * (i + 2) + 1], py = row[4 * (i + 2) + 2],
           pz = row[4 * (i + 2) + 3];
    if (do_boost) {
      double bp = px * bx + py * by + pz * bz;
      double f = gamma2 * bp + gamma * E;
      px += f * bx; py += f * by; pz += f * bz;
    }
    lx[i] = px; ly[i] = py; lz[i] = pz;
    double pt = sqrt(px * px + py * py);
    ptmin = fmin(ptmin, pt);
    double eta = pseudo_rap_d(px, py, pz);
    etamax = fmax(etamax, eta);
  }
  if (ptmin < d.pt_min_cut) factor = 0.0;
  if (d.delr_min_cut > 0.0) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) {
        double pt1 = sqrt(lx[i] * lx[i] + ly[i] * ly[i]);
        double pt2 = sqrt(lx[j] * lx[j] + ly[j] * ly[j]);
        double t = (lx[i] * lx[j] + ly[i] * ly[j]) / (pt1 * pt2);
        double dphi = fabs(t) > 1.0 ? acos(t / fabs(t)) : acos(t);
        if (pt1 == 0.0 || pt2 == 0.0) dphi = huge;
        double deta = pseudo_rap_d(lx[i], ly[i], lz[i]) - pseudo_rap_d(lx[j], ly[j], lz[j]);
        if (fabs(sqrt(deta * deta + dphi * dphi)) < d.delr_min_cut) factor = 0.0;
      }
  }
  if (d.rap_max_cut > 0.0 && d.rap_max_cut < fabs(etamax)) factor = 0.0;
  return factor;
}

// -------------------------------------------------------------------------------------------
// forward map of one event.  rin: this event's uniforms in shared memory; row: this event's
// output momenta in shared memory.
template <typename TIO, typename TC, int NF>
__device__ __forceinline__ void rambo_forward_event(const mf_rambo_desc& d, const RamboAux& aux,
                                                    const TIO* rin, TIO* row, TC& weight_out,
                                                    TC& logw_out, TC& x1_out, TC& x2_out,
                                                    int& flags, bool want_logw) {
  constexpr int NMAX = NF ? NF : MF_MAX_FINAL;
  constexpr int UNR = NF ? NF : 1;  // full unroll only when n is a compile-time constant
  constexpr bool kF32 = sizeof(TC) == 4;
  const int n = NF ? NF : d.n_final;
  using M = Mth<TC>;

  // ---- x1, x2 from the last two uniforms (get_x1x2_from_uniform, phasespace/utils.py:260-284)
  const TC ua = (TC)rin[3 * n - 4], ub = (TC)rin[3 * n - 3];
  if (ua != ua || ub != ub) flags |= MF_FLAG_NAN_INPUT;  // rambo_generator.py:191-199
  TC x1, x2, E_cm, K0, wgt_x;
  if constexpr (!kF32) {
    TC mlogx1 = ((TC)d.x_c1 + M::sqrt_((TC)d.x_c1sq + (TC(-2) * ua) * (TC)aux.inv_A)) * (TC)aux.inv_c2;
    TC mlogx2 = (-mlogx1 + (TC)d.x_q) * ub;
    x1 = M::exp_(-mlogx1);
    x2 = M::exp_(-mlogx2);
    wgt_x = (TC)d.x_A * x1 * x2;
    E_cm = M::sqrt_(x1 * x2) * (TC)d.sqrt_s;  // rambo_generator.py:226
    K0 = E_cm - (TC)d.sum_mass;               // :478
  } else {
    // same map written without cancellations: -ln x1 = q (1 - sqrt(1-ua)) = q ua / (1 + s)
    const TC q = (TC)d.x_q;
    TC s = M::sqrt_(TC(1) - ua);
    TC l1 = q * ua / (TC(1) + s);
    TC rest = q * s;  // q - l1
    TC l2 = rest * ub;
    x1 = M::exp_(-l1);
    x2 = M::exp_(-l2);
    wgt_x = (TC)d.x_A * x1 * x2;
    TC half_delta = TC(0.5) * rest * (TC(1) - ub);  // (q - l1 - l2)/2 >= 0
    TC em1 = expm1f(half_delta);
    E_cm = (TC)aux.e_thr * (em1 + TC(1));
    K0 = (TC)aux.e_thr * em1 + (TC)aux.thr_gap;  // E_cm - sum_mass
  }
  if (!(K0 > TC(0))) flags |= MF_FLAG_BELOW_THRESHOLD;

  // ---- intermediate masses (:448-509)
  TC K[NMAX], Mi[NMAX], U[NMAX];
  K[0] = K0;
  {
    TC prod_u = TC(1);
#pragma unroll UNR
    for (int c = 0; c < NMAX - 2; ++c) {
      if (c < n - 2) {
        TC v = (TC)rin[c];
        if (v != v) flags |= MF_FLAG_NAN_INPUT;
        TC u = solve_massless_u<TC>(v, n - 2 - c);
        U[c] = u;
        prod_u *= u;
        K[c + 1] = K0 * prod_u;  // cumprod, :455-458
      }
    }
  }
#pragma unroll UNR
  for (int i = 0; i < NMAX - 1; ++i)
    if (i < n - 1) Mi[i] = K[i] + (TC)d.tail_sum_fwd[i];  // :487
  Mi[n - 1] = (TC)d.mass[n - 1];                           // :303-309

  // ---- weight (:489-507) and rest-frame momenta q_i = 4 M_i rho(M_i, M_{i+1}, m_i) (:310-314)
  double w = (double)wgt_x;  // product spans > fp32 range for n = 8 ((E^2)^6 ~ 1e49)
  TC qi[NMAX];
#pragma unroll UNR
  for (int i = 0; i < NMAX - 1; ++i) {
    if (i < n - 1) {
      const TC m = (TC)d.mass[i];
      const TC Mn = Mi[i + 1];
      TC a, b;
      if constexpr (!kF32) {
        const TC Msq = Mi[i] * Mi[i];
        a = Msq - (Mn + m) * (Mn + m);
        b = Msq - (Mn - m) * (Mn - m);
      } else {
        // M_i - M_{i+1} - m_i = K_i - K_{i+1} + (S_i - S_{i+1} - m_i): no cancellation
        TC dK = (i < n - 2) ? K[i] * (TC(1) - U[i]) : K[i];
        TC gap = dK + (TC)aux.dS[i];
        a = gap * (Mi[i] + Mn + m);
        b = (Mi[i] - Mn + m) * (Mi[i] + Mn - m);
      }
      const TC sab = M::sqrt_(a * b);
      const TC invM = fast_div(TC(1), Mi[i]);
      qi[i] = TC(0.5) * sab * invM;
      if (i < n - 2) {
        // rho(M_i,M_{i+1},m_i) / rho(K_i,K_{i+1},0) * M_{i+1}/K_{i+1}
        TC omu2 = (TC(1) - U[i]) * (TC(1) + U[i]);  // (K_i^2 - K_{i+1}^2)/K_i^2
        w *= (double)fast_div(sab * invM * invM * Mn, omu2 * K[i + 1]);
      }
    }
  }
  {
    // 8 rho(M_{n-2}, m_{n-1}, m_{n-2}) with the fp32-rounded squares of the reference (:489-493)
    const TC Msq = Mi[n - 2] * Mi[n - 2];
    TC t = fast_div(M::sqrt_((Msq - (TC)d.last_plus_sq) * (Msq - (TC)d.last_minus_sq)), Msq);
    w *= (double)t;
    TC km = fast_div(K[0], Mi[0]);
    w *= (double)ipow(km, 2 * n - 4);
    // Vol(E_cm, n), :111-140
    double E2 = (double)E_cm * (double)E_cm;
    w *= aux.vol_scale * ipow(E2, n - 2);
  }
  logw_out = want_logw ? (TC)log(w) : TC(0);
  if (d.quirks & MF_QUIRK_F32_VOLUME) {
    // the reference evaluates (E_cm^2)^(n-2) in fp32: +inf above ~1.6 TeV for n = 8
    float Ef = (float)E_cm;
    double p = 1.0, e2 = (double)(Ef * Ef);
    for (int i = 0; i < n - 2; ++i) p *= e2;
    if (p > (double)FLT_MAX) w = (double)INFINITY;
  }

  // ---- decay chain (:294-345), boost written with gamma = E_Q / M_i (Q^2 = M_i^2 by construction)
  TC Qx = TC(0), Qy = TC(0), Qz = TC(0), QE = Mi[0];
#pragma unroll UNR
  for (int i = 0; i < NMAX - 1; ++i) {
    if (i < n - 1) {
      TC rc = (TC)rin[n - 2 + 2 * i], rp = (TC)rin[n - 1 + 2 * i];
      if (rc != rc || rp != rp) flags |= MF_FLAG_NAN_INPUT;
      TC cos_t = TC(2) * rc - TC(1);
      TC sin_t = M::sqrt_(kF32 ? TC(4) * rc * (TC(1) - rc) : TC(1) - cos_t * cos_t);
      TC sin_p, cos_p;
      M::sincospi_(TC(2) * rp, &sin_p, &cos_p);
      TC px = qi[i] * sin_t * cos_p, py = qi[i] * sin_t * sin_p, pz = qi[i] * cos_t;
      const TC msq = (TC)d.mass_sq[i];
      TC Ep = M::sqrt_(qi[i] * qi[i] + msq);
      // boost_t(p, Q_vec/Q_E) (phasespace/utils.py:88-118)
      TC Mq = Mi[i];
      TC bp = Qx * px + Qy * py + Qz * pz;
      TC inv = fast_div(TC(1), Mq * (QE + Mq));
      TC f = bp * inv + Ep * (QE + Mq) * inv;  // gamma2*bp*/... + gamma*E, per unit of Q_vec
      px += f * Qx; py += f * Qy; pz += f * Qz;
      TC E = M::sqrt_(px * px + py * py + pz * pz + msq);  // set_square_t, :339
      TIO* o = row + 4 * (i + 2);
      o[0] = (TIO)E; o[1] = (TIO)px; o[2] = (TIO)py; o[3] = (TIO)pz;
      Qx -= px; Qy -= py; Qz -= pz;
      QE = M::sqrt_(Qx * Qx + Qy * Qy + Qz * Qz + Mi[i + 1] * Mi[i + 1]);  // :341-343
    }
  }
  {
    TIO* o = row + 4 * (n + 1);  // last particle = remaining Q, :345
    o[0] = (TIO)QE; o[1] = (TIO)Qx; o[2] = (TIO)Qy; o[3] = (TIO)Qz;
    // incoming massless partons from float(E_cm), :533-551
    TIO h = (TIO)((double)(float)E_cm * 0.5);
    row[0] = h; row[1] = (TIO)0; row[2] = (TIO)0; row[3] = h;
    row[4] = h; row[5] = (TIO)0; row[6] = (TIO)0; row[7] = -h;
  }
  if (aux.cuts_on) w *= rambo_cut_factor<TIO>(d, n, row, (double)x1, (double)x2);
  weight_out = (TC)w;
  x1_out = x1;
  x2_out = x2;
}

// staging helpers -------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void tile_load(T* s_dst, const T* g_src, int count, bool vec_ok) {
  if (vec_ok && (count * sizeof(T)) % 16 == 0) {
    const int nv = count * (int)sizeof(T) / 16;
    const int4* g4 = reinterpret_cast<const int4*>(g_src);
    int4* s4 = reinterpret_cast<int4*>(s_dst);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) s4[i] = __ldg(g4 + i);
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) s_dst[i] = g_src[i];
  }
}
template <typename T>
__device__ __forceinline__ void tile_store(T* g_dst, const T* s_src, int count, bool vec_ok) {
  if (vec_ok && (count * sizeof(T)) % 16 == 0) {
    const int nv = count * (int)sizeof(T) / 16;
    int4* g4 = reinterpret_cast<int4*>(g_dst);
    const int4* s4 = reinterpret_cast<const int4*>(s_src);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) g4[i] = s4[i];
  } else {
    for (int i = threadIdx.x; i < count; i += blockDim.x) g_dst[i] = s_src[i];
  }
}

template <typename TIO, typename TC, int NF>
__global__ void __launch_bounds__(kEPB)
rambo_forward_kernel(const __grid_constant__ mf_rambo_desc d, const __grid_constant__ RamboAux aux,
                     const TIO* __restrict__ g_r, long long n_events, TIO* __restrict__ g_mom,
                     TIO* __restrict__ g_w, TIO* __restrict__ g_logw, TIO* __restrict__ g_x1,
                     TIO* __restrict__ g_x2, int* __restrict__ g_flags, int use_tma, int vec_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = NF ? NF : d.n_final;
  const int ndim = 3 * n - 2, mdim = 4 * (n + 2);
  TIO* s_in = reinterpret_cast<TIO*>(smem_raw);
  TIO* s_out = s_in + kEPB * ndim;  // kEPB*ndim*sizeof(TIO) is a multiple of 16
  __shared__ __align__(8) uint64_t bar;

  const long long first = (long long)blockIdx.x * kEPB;
  const int tile = (int)min((long long)kEPB, n_events - first);
  const bool full = (tile == kEPB);
  const bool tma = use_tma && full;
  const int tid = threadIdx.x;

  if (tma) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(kEPB * ndim * sizeof(TIO));
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(s_in, g_r + first * ndim, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    tile_load(s_in, g_r + first * ndim, tile * ndim, vec_ok != 0);
    __syncthreads();
  }

  if (tid < tile) {
    TC w, lw, x1, x2;
    int flags = 0;
    rambo_forward_event<TIO, TC, NF>(d, aux, s_in + tid * ndim, s_out + tid * mdim, w, lw, x1, x2,
                                     flags, g_logw != nullptr);
    const long long e = first + tid;
    if (g_w) g_w[e] = (TIO)w;
    if (g_logw) g_logw[e] = (TIO)lw;
    if (g_x1) g_x1[e] = (TIO)x1;
    if (g_x2) g_x2[e] = (TIO)x2;
    if (g_flags) g_flags[e] = flags;
  }

  if (g_mom) {
    if (tma) {
      fence_proxy_async_smem();  // make the generic-proxy smem writes visible to the bulk engine
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(g_mom + first * mdim, s_out, (uint32_t)(kEPB * mdim * sizeof(TIO)));
        bulk_commit();
        bulk_wait_read0();  // smem must stay alive until the engine has read it
      }
    } else {
      __syncthreads();
      tile_store(g_mom + first * mdim, s_out, tile * mdim, vec_ok != 0);
    }
  }
}

// -------------------------------------------------------------------------------------------
// inverse map of one event (getPSpoint_batch :615-736, get_PS unfolding_flow/utils.py:1056-1152).
// pin: this event's final-state momenta [n][4] in shared memory (un-permuted).  rout: uniforms.
// swz: particle slot XOR of this event's row (0, or row & 7 when the tile was staged with the 16-byte-chunk swizzle
// that makes 128-byte rows conflict-free: all lanes read the same particle index at the same time).
template <typename TIO, typename TC, int NF>
__device__ __forceinline__ void rambo_inverse_event(const mf_rambo_desc& d, const RamboAux& aux,
                                                    const TIO* pin, TIO* rout, TC x1, TC x2,
                                                    TC& prob_out, int& flags, bool want_prob,
                                                    bool& r_out_of_range, int swz) {
  constexpr int NMAX = NF ? NF : MF_MAX_FINAL;
  constexpr int UNR = NF ? NF : 1;
  const int n = NF ? NF : d.n_final;
  using M = Mth<TC>;
  TC Mi[NMAX], K[NMAX];
  // suffix sums Q_j = sum_{k>=j} P_k and M_j = sqrt(Q_j^2), K_j = M_j - sum_{k>=j} m_k (:656-660)
  TC QE = TC(0), Qx = TC(0), Qy = TC(0), Qz = TC(0);
  // first pass from the back: masses
#pragma unroll UNR
  for (int jj = 0; jj < NMAX; ++jj) {
    const int j = n - 1 - jj;
    if (jj < n) {
      const TIO* p = pin + 4 * (aux.perm[j] ^ swz);
      QE += (TC)p[0]; Qx += (TC)p[1]; Qy += (TC)p[2]; Qz += (TC)p[3];
      TC m2 = QE * QE - Qx * Qx - Qy * Qy - Qz * Qz;
      Mi[j] = M::sqrt_(m2);
      K[j] = Mi[j] - (TC)d.tail_sum_inv[j];
    }
  }
  // total momentum now in Q: CM check (:630-635)
  if (!(Qx * Qx + Qy * Qy + Qz * Qz < TC(1e-3))) flags |= MF_FLAG_NOT_CM;
  r_out_of_range = false;
  // second pass: j = n-1 .. 1, rebuild Q_{j-1} = Q_j + P_{j-1} from the back
  TC SE = TC(0), Sx = TC(0), Sy = TC(0), Sz = TC(0);
  {
    const TIO* p = pin + 4 * (aux.perm[n - 1] ^ swz);
    SE = (TC)p[0]; Sx = (TC)p[1]; Sy = (TC)p[2]; Sz = (TC)p[3];
  }
#pragma unroll UNR
  for (int jj = 0; jj < NMAX - 1; ++jj) {
    const int j = n - 1 - jj;  // j = n-1 .. 1
    if (jj < n - 1) {
      const int i = j + 1;     // reference's 1-based loop index
      TC u = fast_div(K[j], K[j - 1]);
      const int e = n - i;
      TC rm = TC(e + 1) * ipow(u, e) - TC(e) * ipow(u, e + 1);  // :672-674
      const TIO* p = pin + 4 * (aux.perm[j - 1] ^ swz);
      TC pE = (TC)p[0], px = (TC)p[1], py = (TC)p[2], pz = (TC)p[3];
      SE += pE; Sx += px; Sy += py; Sz += pz;  // Q_{j-1}
      // boost_t(P_{j-1}, -Q_vec/Q_E) (:678, phasespace/utils.py:88-118)
      const TC iSE = fast_div(TC(1), SE);
      TC bx = -(Sx * iSE), by = -(Sy * iSE), bz = -(Sz * iSE);
      TC b2 = bx * bx + by * by + bz * bz;
      TC gamma = fast_div(TC(1), M::sqrt_(TC(1) - b2));
      TC bp = px * bx + py * by + pz * bz;
      TC gamma2 = b2 > TC(0) ? fast_div(gamma - TC(1), b2) : TC(0);
      TC f = gamma2 * bp + gamma * pE;
      px += f * bx; py += f * by; pz += f * bz;
      TC rcos = (fast_div(pz, M::sqrt_(px * px + py * py + pz * pz)) + TC(1)) * TC(0.5);  // :681-683
      // fp32 outputs: the angle is stored as an fp32 uniform, an fp32-accurate arctangent of the fp64 ratio is enough
      // (|d phi| <= 1.2e-7 -> 2e-8 of the unit interval, below the fp32 rounding of r); fp64 outputs keep the fp64 atan
      const TC ratio = fast_div(py, px);
      TC phi = (sizeof(TIO) == 4) ? (TC)atanf((float)ratio) : M::atan_(ratio);                // :687
      const TC pi = TC(3.14159265358979323846);
      TC dphi = (py < TC(0) && px > TC(0)) ? TC(2) * pi : TC(0);
      dphi += (px < TC(0)) ? pi : TC(0);
      phi += dphi;
      TC rphi = phi * TC(0.15915494309189533577);  // / (2 pi)
      // the reference keeps r in an fp32 tensor (:663): round there for the range test
      float f0 = (float)rm, f1 = (float)rcos, f2 = (float)rphi;
      r_out_of_range |= (f0 < 0.f) | (f0 > 1.f) | (f1 < 0.f) | (f1 > 1.f) | (f2 < 0.f) | (f2 > 1.f);
      rout[j - 1] = (TIO)f0;
      rout[n - 5 + 2 * i - 1] = (TIO)f1;
      rout[n - 4 + 2 * i - 1] = (TIO)f2;
    }
  }
  // x1, x2 -> uniforms (get_uniform_from_x1x2, phasespace/utils.py:287-309)
  TC l1 = -M::log_(x1), l2 = -M::log_(x2);
  TC u1 = fast_div(TC(-0.5) * (l1 * l1) + (TC)d.x_q * l1, (TC)d.x_A);
  TC u2 = fast_div(l2, -l1 + (TC)d.x_q);
  rout[3 * n - 4] = (TIO)u1;
  rout[3 * n - 3] = (TIO)u2;
  prob_out = TC(0);
  if (want_prob) {
    TC jac = fast_div(TC(1), (TC)d.x_A * x1 * x2);
    TC E_cm = M::sqrt_(x1 * x2) * (TC)d.sqrt_s;  // :700
    TC w = fast_div((TC)d.vol_coef * ipow(E_cm * E_cm, n - 2), (TC)d.vol_fact);
    const TC Msq = Mi[n - 2] * Mi[n - 2];
    w *= fast_div(M::sqrt_((Msq - (TC)d.last_plus_sq) * (Msq - (TC)d.last_minus_sq)), Msq);  // 8 rho, :709-713
#pragma unroll UNR
    for (int i = 0; i < NMAX - 2; ++i) {
      if (i < n - 2) {
        const TC m = (TC)d.mass[i];
        const TC Ms = Mi[i] * Mi[i];
        TC a = Ms - (Mi[i + 1] + m) * (Mi[i + 1] + m);
        TC b = Ms - (Mi[i + 1] - m) * (Mi[i + 1] - m);
        // (rhoM / rhoK) * (M_{i+1} / K_{i+1}) with rhoM = sqrt(ab) / 8 Ms, rhoK = (Ks - K1^2) / 8 Ks: one division
        TC Ks = K[i] * K[i];
        TC num = M::sqrt_(a * b) * Ks * Mi[i + 1];
        TC den = Ms * (Ks - K[i + 1] * K[i + 1]) * K[i + 1];
        w *= fast_div(num, den);  // :714-725
      }
    }
    w *= ipow(fast_div(K[0], Mi[0]), 2 * n - 4);  // :727
    TC prob = fast_div(jac, w);
    if (d.quirks & MF_QUIRK_F32_VOLUME) {
      float Ef = (float)E_cm;
      double p = 1.0, e2 = (double)(Ef * Ef);
      for (int i = 0; i < n - 2; ++i) p *= e2;
      if (p > (double)FLT_MAX) prob = TC(0);  // 1/inf
    }
    prob_out = prob;
  }
}

// mode 0: getPSpoint_batch (x1,x2 given);  mode 1: get_PS (boost four-vector given, bool mask out)
template <typename TIO, typename TC, int NF>
__global__ void __launch_bounds__(kEPB)
rambo_inverse_kernel(const __grid_constant__ mf_rambo_desc d, const __grid_constant__ RamboAux aux,
                     const TIO* __restrict__ g_mom, const TIO* __restrict__ g_x1,
                     const TIO* __restrict__ g_x2, const TIO* __restrict__ g_boost,
                     long long n_events, TIO* __restrict__ g_r, TIO* __restrict__ g_prob,
                     int* __restrict__ g_flags, unsigned char* __restrict__ g_mask, int mode,
                     int use_tma, int vec_ok) {
  using M = Mth<TC>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = NF ? NF : d.n_final;
  const int ndim = 3 * n - 2, mdim = 4 * n;
  TIO* s_in = reinterpret_cast<TIO*>(smem_raw);
  TIO* s_out = s_in + kEPB * mdim;
  __shared__ __align__(8) uint64_t bar;

  const long long first = (long long)blockIdx.x * kEPB;
  const int tile = (int)min((long long)kEPB, n_events - first);
  const bool full = (tile == kEPB);
  const bool tma = use_tma && full;
  const int tid = threadIdx.x;
  // 128-byte rows (n = 8, fp32): every lane reads the same particle of its own row, i.e. the same bank group -- a
  // 32-way conflict on a flat tile.  Stage those tiles with coalesced 16-byte loads and store chunk c of row r at
  // c ^ (r & 7); the reads below XOR the particle index the same way.
  const bool swizzled = (sizeof(TIO) == 4) && (n == 8) && (vec_ok != 0);
  const int swz = swizzled ? (tid & 7) : 0;

  if (swizzled) {
    const float4* src = reinterpret_cast<const float4*>(g_mom + first * mdim);
    float4* dst = reinterpret_cast<float4*>(s_in);
    for (int i = tid; i < tile * 8; i += kEPB) {
      const int row = i >> 3, c = i & 7;
      dst[row * 8 + (c ^ (row & 7))] = __ldg(src + i);
    }
    __syncthreads();
  } else if (tma) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(kEPB * mdim * sizeof(TIO));
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(s_in, g_mom + first * mdim, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    tile_load(s_in, g_mom + first * mdim, tile * mdim, vec_ok != 0);
    __syncthreads();
  }

  if (tid < tile) {
    const long long e = first + tid;
    TC x1, x2, prob;
    int flags = 0;
    bool problem = false;
    if (mode == 0) {
      x1 = (TC)g_x1[e];
      x2 = (TC)g_x2[e];
    } else {
      // unfolding_flow/utils.py:1066-1071
      const TIO* b = g_boost + 4 * e;
      TC bE = (TC)b[0], bz = (TC)b[3];
      if (mode == 2) {
        // memflow/unfolding_flow/unfolding_flow.py:137-148 in front of get_PS: fix an unphysical regressed boost, then
        // move the lab-frame momenta to the CM frame with boost_tt(p, -boostVector_t(boost)) (phasespace/utils.py:36-42,
        // :121-147).  The CM momenta go back into the tile in the I/O dtype (the reference's tensors keep the model dtype).
        TC bx = (TC)b[1], by = (TC)b[2];
        if (aux.m_min_tot > 0.0) {
          const TC mm = M::sqrt_(bE * bE - bz * bz);   // NaN for a space-like boost: comparison false, like torch
          if (mm < (TC)aux.m_min_tot) bE = M::sqrt_(bz * bz + (TC)(aux.m_min_tot * aux.m_min_tot) + TC(1e-3));
        }
        if (aux.lab_boost_out) {
          TIO* bo = static_cast<TIO*>(aux.lab_boost_out) + 4 * e;
          bo[0] = (TIO)bE; bo[1] = (TIO)bx; bo[2] = (TIO)by; bo[3] = (TIO)bz;
        }
        bE = (TC)(TIO)bE;   // the reference reads the fixed value back from the tensor
        const TC ibE = fast_div(TC(1), bE);
        const TC vx = -(bx * ibE), vy = -(by * ibE), vz = -(bz * ibE);   // -beta
        const TC b2 = vx * vx + vy * vy + vz * vz;
        const TC gamma = fast_div(TC(1), M::sqrt_(TC(1) - b2));
        const TC gamma2 = b2 > TC(0) ? fast_div(gamma - TC(1), b2) : TC(0);
        for (int k = 0; k < n; ++k) {
          TIO* p = s_in + tid * mdim + 4 * (k ^ swz);
          const TC pE = (TC)p[0], px = (TC)p[1], py = (TC)p[2], pz = (TC)p[3];
          const TC bp = px * vx + py * vy + pz * vz;
          const TC f = gamma2 * bp + gamma * pE;
          p[0] = (TIO)(gamma * (pE + bp));
          p[1] = (TIO)(px + f * vx); p[2] = (TIO)(py + f * vy); p[3] = (TIO)(pz + f * vz);
        }
      }
      x1 = (bE + bz) / (TC)d.sqrt_s;
      x2 = (bE - bz) / (TC)d.sqrt_s;
      problem = (x1 > TC(1)) | (x1 < TC(0)) | (x2 < TC(0)) | (x1 > TC(1));  // sic, :1069
      x1 = fmin(fmax(x1, TC(0)), TC(1));
      x2 = fmin(fmax(x2, TC(0)), TC(1));
    }
    bool oor = false;
    rambo_inverse_event<TIO, TC, NF>(d, aux, s_in + tid * mdim, s_out + tid * ndim, x1, x2, prob,
                                     flags, (mode == 0) && g_prob != nullptr, oor, swz);
    if (mode == 0) {
      if (g_prob) g_prob[e] = (TIO)prob;
      if (g_flags) g_flags[e] = flags;
    } else {
      // CM test of get_PS is 1e-5 (:1079-1081); recompute from the tile
      TC sx = TC(0), sy = TC(0), sz = TC(0);
      for (int k = 0; k < n; ++k) {
        const TIO* p = s_in + tid * mdim + 4 * (k ^ swz);
        sx += (TC)p[1]; sy += (TC)p[2]; sz += (TC)p[3];
      }
      problem |= (sx * sx + sy * sy + sz * sz > TC(1e-5));
      problem |= oor;
      if (g_mask) g_mask[e] = problem ? 1 : 0;
    }
  }

  if (tma) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(g_r + first * ndim, s_out, (uint32_t)(kEPB * ndim * sizeof(TIO)));
      if (mode == 2 && aux.lab_mom_cm)
        bulk_s2g(static_cast<TIO*>(aux.lab_mom_cm) + first * mdim, s_in, (uint32_t)(kEPB * mdim * sizeof(TIO)));
      bulk_commit();
      bulk_wait_read0();
    }
  } else {
    __syncthreads();
    tile_store(g_r + first * ndim, s_out, tile * ndim, vec_ok != 0);
    if (mode == 2 && aux.lab_mom_cm) {
      TIO* dst = static_cast<TIO*>(aux.lab_mom_cm) + first * mdim;
      if (swizzled) {   // undo the 16-byte-chunk swizzle of the staged tile
        const float4* src = reinterpret_cast<const float4*>(s_in);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < tile * 8; i += kEPB) { const int row = i >> 3, c = i & 7; d4[i] = src[row * 8 + (c ^ (row & 7))]; }
      } else {
        tile_store(dst, s_in, tile * mdim, vec_ok != 0 && aligned16_dev(dst));
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// host launchers
int fill_aux(const mf_rambo_desc* d, const int32_t* perm, RamboAux* aux) {
  const int n = d->n_final;
  MF_REQUIRE(n >= 3 && n <= MF_MAX_FINAL, "n_final=%d outside 3..%d", n, MF_MAX_FINAL);
  aux->e_thr = d->sqrt_s * exp(-0.5 * d->x_q);
  aux->thr_gap = aux->e_thr - d->sum_mass;
  aux->inv_A = 1.0 / d->x_A;
  aux->inv_c2 = 1.0 / d->x_c2;
  aux->vol_scale = d->vol_coef / d->vol_fact;
  for (int i = 0; i < MF_MAX_FINAL; ++i) {
    aux->dS[i] = 0.0;
    aux->perm[i] = i;
  }
  for (int i = 0; i + 1 < n; ++i)
    aux->dS[i] = d->tail_sum_fwd[i] - d->tail_sum_fwd[i + 1] - d->mass[i];
  aux->cuts_on = (d->pt_min_cut > 0.0 || d->delr_min_cut > 0.0 || d->rap_max_cut > 0.0) ? 1 : 0;
  aux->m_min_tot = 0.0; aux->lab_mom_cm = nullptr; aux->lab_boost_out = nullptr;
  if (perm) {
    unsigned seen = 0;
    for (int i = 0; i < n; ++i) {
      MF_REQUIRE(perm[i] >= 0 && perm[i] < n, "perm[%d]=%d outside 0..%d", i, perm[i], n - 1);
      seen |= 1u << perm[i];
      aux->perm[i] = perm[i];
    }
    MF_REQUIRE(seen == (1u << n) - 1u, "perm is not a permutation of 0..%d", n - 1);
  }
  return MF_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename TIO, typename TC, int NF>
int launch_forward_n(const mf_rambo_desc& d, const RamboAux& aux, const TIO* r, long long N,
                     TIO* mom, TIO* w, TIO* lw, TIO* x1, TIO* x2, int* flags, cudaStream_t st) {
  const int n = d.n_final;
  const size_t smem = (size_t)kEPB * ((3 * n - 2) + 4 * (n + 2)) * sizeof(TIO);
  auto kern = rambo_forward_kernel<TIO, TC, NF>;
  MF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vec_ok = aligned16(r) && (mom == nullptr || aligned16(mom));
  const int use_tma = vec_ok && tma_enabled();
  const unsigned grid = (unsigned)((N + kEPB - 1) / kEPB);
  kern<<<grid, kEPB, smem, st>>>(d, aux, r, N, mom, w, lw, x1, x2, flags, use_tma, vec_ok);
  return post_launch("rambo_forward_kernel");
}

template <typename TIO, typename TC>
int launch_forward(const mf_rambo_desc& d, const RamboAux& aux, const TIO* r, long long N, TIO* mom,
                   TIO* w, TIO* lw, TIO* x1, TIO* x2, int* flags, cudaStream_t st) {
  switch (d.n_final) {
    case 3: return launch_forward_n<TIO, TC, 3>(d, aux, r, N, mom, w, lw, x1, x2, flags, st);
    case 4: return launch_forward_n<TIO, TC, 4>(d, aux, r, N, mom, w, lw, x1, x2, flags, st);
    case 5: return launch_forward_n<TIO, TC, 5>(d, aux, r, N, mom, w, lw, x1, x2, flags, st);
    case 6: return launch_forward_n<TIO, TC, 6>(d, aux, r, N, mom, w, lw, x1, x2, flags, st);
    case 8: return launch_forward_n<TIO, TC, 8>(d, aux, r, N, mom, w, lw, x1, x2, flags, st);
    default: return launch_forward_n<TIO, TC, 0>(d, aux, r, N, mom, w, lw, x1, x2, flags, st);
  }
}

template <typename TIO, typename TC, int NF>
int launch_inverse_n(const mf_rambo_desc& d, const RamboAux& aux, const TIO* mom, const TIO* x1,
                     const TIO* x2, const TIO* boost, long long N, TIO* r, TIO* prob, int* flags,
                     unsigned char* mask, int mode, cudaStream_t st) {
  const int n = d.n_final;
  const size_t smem = (size_t)kEPB * ((3 * n - 2) + 4 * n) * sizeof(TIO);
  auto kern = rambo_inverse_kernel<TIO, TC, NF>;
  MF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vec_ok = aligned16(mom) && aligned16(r);
  const int use_tma = vec_ok && tma_enabled();
  const unsigned grid = (unsigned)((N + kEPB - 1) / kEPB);
  kern<<<grid, kEPB, smem, st>>>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, use_tma,
                                 vec_ok);
  return post_launch("rambo_inverse_kernel");
}

template <typename TIO, typename TC>
int launch_inverse(const mf_rambo_desc& d, const RamboAux& aux, const TIO* mom, const TIO* x1,
                   const TIO* x2, const TIO* boost, long long N, TIO* r, TIO* prob, int* flags,
                   unsigned char* mask, int mode, cudaStream_t st) {
  switch (d.n_final) {
    case 3: return launch_inverse_n<TIO, TC, 3>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, st);
    case 4: return launch_inverse_n<TIO, TC, 4>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, st);
    case 5: return launch_inverse_n<TIO, TC, 5>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, st);
    case 6: return launch_inverse_n<TIO, TC, 6>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, st);
    case 8: return launch_inverse_n<TIO, TC, 8>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, st);
    default: return launch_inverse_n<TIO, TC, 0>(d, aux, mom, x1, x2, boost, N, r, prob, flags, mask, mode, st);
  }
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" {

int mf_rambo_forward(const mf_rambo_desc* desc, const void* d_r, int64_t n_events, void* d_momenta,
                     void* d_weight, void* d_logw, void* d_x1, void* d_x2, int32_t* d_flags,
                     int io_dtype, int compute, void* stream) {
  MF_REQUIRE(n_events >= 0, "mf_rambo_forward: negative event count");
  MF_REQUIRE(desc && (d_r || n_events == 0), "mf_rambo_forward: null desc or uniforms");
  RamboAux aux;
  int rc = fill_aux(desc, nullptr, &aux);
  if (rc != MF_OK) return rc;
  if (n_events == 0) return MF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (io_dtype == MF_F32 && compute == MF_COMPUTE_F64)
    return launch_forward<float, double>(*desc, aux, (const float*)d_r, n_events, (float*)d_momenta,
                                         (float*)d_weight, (float*)d_logw, (float*)d_x1,
                                         (float*)d_x2, d_flags, st);
  if (io_dtype == MF_F32 && compute == MF_COMPUTE_F32)
    return launch_forward<float, float>(*desc, aux, (const float*)d_r, n_events, (float*)d_momenta,
                                        (float*)d_weight, (float*)d_logw, (float*)d_x1,
                                        (float*)d_x2, d_flags, st);
  if (io_dtype == MF_F64 && compute == MF_COMPUTE_F64)
    return launch_forward<double, double>(*desc, aux, (const double*)d_r, n_events,
                                          (double*)d_momenta, (double*)d_weight, (double*)d_logw,
                                          (double*)d_x1, (double*)d_x2, d_flags, st);
  set_error("mf_rambo_forward: unsupported io_dtype=%d / compute=%d", io_dtype, compute);
  return MF_EUNSUPPORTED;
}

int mf_rambo_forward_f32(const mf_rambo_desc* desc, const float* d_r, int64_t n_events,
                         float* d_momenta, float* d_weight, float* d_logw, float* d_x1, float* d_x2,
                         int32_t* d_flags, void* stream) {
  return mf_rambo_forward(desc, d_r, n_events, d_momenta, d_weight, d_logw, d_x1, d_x2, d_flags,
                          MF_F32, MF_COMPUTE_F64, stream);
}

int mf_rambo_forward_f64(const mf_rambo_desc* desc, const double* d_r, int64_t n_events,
                         double* d_momenta, double* d_weight, double* d_logw, double* d_x1,
                         double* d_x2, int32_t* d_flags, void* stream) {
  return mf_rambo_forward(desc, d_r, n_events, d_momenta, d_weight, d_logw, d_x1, d_x2, d_flags,
                          MF_F64, MF_COMPUTE_F64, stream);
}

static int inverse_dispatch(const mf_rambo_desc* desc, const void* mom, const void* x1,
                            const void* x2, const void* boost, const int32_t* perm, int64_t N,
                            void* r, void* prob, int32_t* flags, uint8_t* mask, int mode,
                            int io_dtype, int compute, void* stream, double m_min_tot = 0.0,
                            void* mom_cm = nullptr, void* boost_out = nullptr) {
  RamboAux aux;
  int rc = fill_aux(desc, perm, &aux);
  if (rc != MF_OK) return rc;
  aux.m_min_tot = m_min_tot; aux.lab_mom_cm = mom_cm; aux.lab_boost_out = boost_out;
  if (N == 0) return MF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (io_dtype == MF_F32 && compute == MF_COMPUTE_F64)
    return launch_inverse<float, double>(*desc, aux, (const float*)mom, (const float*)x1,
                                         (const float*)x2, (const float*)boost, N, (float*)r,
                                         (float*)prob, flags, mask, mode, st);
  if (io_dtype == MF_F32 && compute == MF_COMPUTE_F32)
    return launch_inverse<float, float>(*desc, aux, (const float*)mom, (const float*)x1,
                                        (const float*)x2, (const float*)boost, N, (float*)r,
                                        (float*)prob, flags, mask, mode, st);
  if (io_dtype == MF_F64 && compute == MF_COMPUTE_F64)
    return launch_inverse<double, double>(*desc, aux, (const double*)mom, (const double*)x1,
                                          (const double*)x2, (const double*)boost, N, (double*)r,
                                          (double*)prob, flags, mask, mode, st);
  set_error("rambo inverse: unsupported io_dtype=%d / compute=%d", io_dtype, compute);
  return MF_EUNSUPPORTED;
}

int mf_rambo_inverse(const mf_rambo_desc* desc, const void* d_momenta, const void* d_x1,
                     const void* d_x2, const int32_t* perm, int64_t n_events, void* d_r,
                     void* d_detjacinv, int32_t* d_flags, int io_dtype, int compute, void* stream) {
  MF_REQUIRE(n_events >= 0, "mf_rambo_inverse: negative event count");
  MF_REQUIRE(desc && ((d_momenta && d_x1 && d_x2 && d_r) || n_events == 0), "mf_rambo_inverse: null argument");
  return inverse_dispatch(desc, d_momenta, d_x1, d_x2, nullptr, perm, n_events, d_r, d_detjacinv,
                          d_flags, nullptr, 0, io_dtype, compute, stream);
}

int mf_rambo_inverse_f32(const mf_rambo_desc* desc, const float* d_momenta, const float* d_x1,
                         const float* d_x2, const int32_t* perm, int64_t n_events, float* d_r,
                         float* d_detjacinv, int32_t* d_flags, void* stream) {
  return mf_rambo_inverse(desc, d_momenta, d_x1, d_x2, perm, n_events, d_r, d_detjacinv, d_flags,
                          MF_F32, MF_COMPUTE_F64, stream);
}

int mf_rambo_inverse_f64(const mf_rambo_desc* desc, const double* d_momenta, const double* d_x1,
                         const double* d_x2, const int32_t* perm, int64_t n_events, double* d_r,
                         double* d_detjacinv, int32_t* d_flags, void* stream) {
  return mf_rambo_inverse(desc, d_momenta, d_x1, d_x2, perm, n_events, d_r, d_detjacinv, d_flags,
                          MF_F64, MF_COMPUTE_F64, stream);
}

int mf_rambo_get_ps(const mf_rambo_desc* desc, const void* d_momenta, const void* d_boost,
                    const int32_t* perm, int64_t n_events, void* d_ps, uint8_t* d_mask, int io_dtype,
                    int compute, void* stream) {
  MF_REQUIRE(desc && ((d_momenta && d_boost && d_ps) || n_events == 0), "mf_rambo_get_ps: null argument");
  MF_REQUIRE(desc->n_final == 4, "mf_rambo_get_ps: the reference hard-codes n = 4 (got %d)",
             desc->n_final);
  MF_REQUIRE(n_events >= 0, "mf_rambo_get_ps: negative event count");
  return inverse_dispatch(desc, d_momenta, nullptr, nullptr, d_boost, perm, n_events, d_ps, nullptr,
                          nullptr, d_mask, 1, io_dtype, compute, stream);
}

int mf_rambo_get_ps_lab(const mf_rambo_desc* desc, const void* d_momenta_lab, const void* d_boost,
                        const int32_t* perm, int64_t n_events, double m_min_tot, void* d_momenta_cm,
                        void* d_boost_fixed, void* d_ps, uint8_t* d_mask, int io_dtype, int compute, void* stream) {
  MF_REQUIRE(desc && ((d_momenta_lab && d_boost && d_ps) || n_events == 0), "mf_rambo_get_ps_lab: null argument");
  MF_REQUIRE(desc->n_final == 4, "mf_rambo_get_ps_lab: the reference hard-codes n = 4 (got %d)", desc->n_final);
  MF_REQUIRE(n_events >= 0, "mf_rambo_get_ps_lab: negative event count");
  return inverse_dispatch(desc, d_momenta_lab, nullptr, nullptr, d_boost, perm, n_events, d_ps, nullptr,
                          nullptr, d_mask, 2, io_dtype, compute, stream, m_min_tot, d_momenta_cm, d_boost_fixed);
}

}  // extern "C"
