// flow_common.cuh -- pieces shared by the flow kernels: blob layout, the rational-quadratic spline,
// base densities.  Semantics follow zuko (MonotonicRQSTransform, DiagNormal, BoxUniform); see
// SURVEY.md section 8(c) and oracle/flow_oracle.py for the restatement they are tested against.
#pragma once
#include <math.h>

#include "common.cuh"

namespace mf {

constexpr int kKC = 16;   // K-chunk of the staged weight tiles (CUDA-core GEMM)
constexpr int kNC = 128;  // output columns per GEMM pass
constexpr int kMaxLayers = MF_MAX_HIDDEN_LAYERS + 1;

inline __host__ __device__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Layout of the packed-weight blob of the CUDA-core ("exact") path, in ELEMENTS of the dtype.
// Per transform: for every layer l < L a transposed, zero-padded matrix Wt[kpad_l][ld_l]
// (Wt[k][j] = mask[j][k] * W[j][k]) followed by bias[ld_l]; the last layer is split into feature
// chunks of `fpc` features, each its own Wt[kpad_L][pw_ld] whose column f*PF + j holds spline
// parameter j of feature f0 + f.
struct FlowLayout {
  int n_layers;  // L + 1
  int D, C, K, T, P;  // P = 3K-1
  int PF;        // per-feature parameter stride inside a chunk (P rounded up to 8)
  int fpc;       // features per last-layer chunk
  int n_chunks;  // ceil(D / fpc)
  int pw_ld;     // leading dimension of a last-layer chunk (multiple of kNC)
  int in_dim[kMaxLayers], out_dim[kMaxLayers], kpad[kMaxLayers], ld[kMaxLayers];
  long long w_off[kMaxLayers], b_off[kMaxLayers];  // last layer: offset of chunk 0
  long long chunk_stride;                           // elements per last-layer chunk (W + bias)
  long long t_stride;                               // elements per transform
  int buf_w;     // width of the activation ping-pong buffers (elements)
  int in0_ld;    // leading dimension of the (x | context) input tile
};

inline int make_flow_layout(const mf_flow_desc& d, FlowLayout* L) {
  MF_REQUIRE(d.features >= 1 && d.features <= MF_MAX_FEATURES, "features=%d outside 1..%d",
             d.features, MF_MAX_FEATURES);
  MF_REQUIRE(d.context >= 0, "negative context");
  MF_REQUIRE(d.bins >= 2 && d.bins <= 127, "bins=%d outside 2..127", d.bins);
  MF_REQUIRE(d.transforms >= 1 && d.transforms <= MF_MAX_TRANSFORMS, "transforms=%d outside 1..%d",
             d.transforms, MF_MAX_TRANSFORMS);
  MF_REQUIRE(d.n_hidden >= 1 && d.n_hidden <= MF_MAX_HIDDEN_LAYERS, "n_hidden=%d outside 1..%d",
             d.n_hidden, MF_MAX_HIDDEN_LAYERS);
  L->n_layers = d.n_hidden + 1;
  L->D = d.features; L->C = d.context; L->K = d.bins; L->T = d.transforms;
  L->P = 3 * d.bins - 1;
  L->PF = round_up(L->P, 8);
  const int cap = round_up(L->PF, kNC);  // parameter buffer holds at least one feature
  L->fpc = cap / L->PF;
  if (L->fpc > d.features) L->fpc = d.features;
  L->n_chunks = (d.features + L->fpc - 1) / L->fpc;
  L->pw_ld = round_up(L->fpc * L->PF, kNC);
  long long off = 0;
  int hmax = 0;
  for (int l = 0; l < L->n_layers; ++l) {
    L->in_dim[l] = (l == 0) ? d.features + d.context : d.hidden[l - 1];
    L->out_dim[l] = (l < d.n_hidden) ? d.hidden[l] : d.features * L->P;
    MF_REQUIRE(L->in_dim[l] >= 1 && L->out_dim[l] >= 1, "layer %d has an empty dimension", l);
    L->kpad[l] = round_up(L->in_dim[l], kKC);
    if (l < d.n_hidden) {
      L->ld[l] = round_up(L->out_dim[l], kNC);
      if (L->ld[l] > hmax) hmax = L->ld[l];
      L->w_off[l] = off; off += (long long)L->kpad[l] * L->ld[l];
      L->b_off[l] = off; off += L->ld[l];
    } else {
      L->ld[l] = L->pw_ld;
      L->chunk_stride = (long long)L->kpad[l] * L->pw_ld + L->pw_ld;
      L->w_off[l] = off;
      L->b_off[l] = off + (long long)L->kpad[l] * L->pw_ld;
      off += L->chunk_stride * L->n_chunks;
    }
  }
  L->t_stride = off;
  L->buf_w = (hmax > L->pw_ld ? hmax : L->pw_ld) + 4;
  L->in0_ld = L->kpad[0] + 4;
  return MF_OK;
}

struct FlowKernelArgs {
  const void* blob;
  const void* x;      // mode 0: data x; mode 1: noise
  const void* ctx;
  long long n;
  void* logprob;
  void* z;            // mode 0: z; mode 1: sampled x
  void* ladj;
  int8_t* bins;
  uint8_t* inside;
  const unsigned short* klim;        // mma mode: k-step limits per 128-column block (NULL: full K)
  const unsigned long long* sched;  // sampling: [T][MF_MAX_FEATURES] chunk bitmasks per sweep (NULL: every chunk every sweep)
  const unsigned long long* skip;   // tcgen05 path: zero-block table [T][jobs] (flow_tc_skip_kernel)
  void* work;         // caller-provided scratch (mf_flow_workspace_bytes)
  long long work_bytes;
  int ctx_group;      // x rows per context row (>= 1)
  int mode;           // 0 = log_prob / transform, 1 = sample
  int passes;
  int base;
  int univariate;
  double bound, slope;
};

struct BaseParams {
  double a[MF_MAX_FEATURES];
  double b[MF_MAX_FEATURES];
  // element-wise map of the data side (mf_flow_xmap): applied to x when it is loaded (log_prob) / stored (sample)
  int xmap_kind;
  double xmap_eps;
  double xmap_scale[MF_MAX_FEATURES];
  double xmap_shift[MF_MAX_FEATURES];
};


// Sampling schedule kept at the tail of every blob: word [t][s] has bit g set when feature chunk g of transform t holds a
// feature that becomes final in sweep s (mf_flow_set_sample_schedule); all ones after packing = zuko's literal sweeps.
constexpr long long kSchedBytes = (long long)MF_MAX_TRANSFORMS * MF_MAX_FEATURES * 8;

// warp-level MMA path for wide hyper-networks (flow_mma.cu); uses the blob of the exact path
struct FlowLayout;
struct FlowKernelArgs;
struct BaseParams;
int flow_mma_run(const FlowLayout& L, const FlowKernelArgs& a, const BaseParams& bp, cudaStream_t st);
// in place: fp32 k-major rows -> fp16 hi/lo k-pairs; then the k-step limits of every 128-column block (klim region)
int flow_mma_presplit(const FlowLayout& L, float* blob, unsigned short* klim, cudaStream_t st);
// mma mode keeps, after the sampling schedule, [T][kKlimEntries] uint16: k16-steps up to the last nonzero weight row of
// every 128-column block (hidden layer l, block nb -> 4 l + nb; last-layer chunk g -> 32 + g)
constexpr int kKlimEntries = 128;
constexpr long long kKlimBytes = (long long)MF_MAX_TRANSFORMS * kKlimEntries * 2;

// tcgen05 path (flow_tc.cu)
int flow_tc_blob_bytes(const mf_flow_desc& d, int gemm, long long* bytes);
int flow_tc_pack(const mf_flow_desc& d, const void* const* w, const void* const* b, const uint8_t* const* m,
                 int gemm, void* blob, long long blob_bytes, cudaStream_t st);
int flow_tc_run(const mf_flow_desc& d, const FlowKernelArgs& a, const BaseParams& bp, int gemm, cudaStream_t st);
long long flow_tc_workspace_bytes(const mf_flow_desc& d, long long n, int gemm, int op);
int flow_tc_sched_layout(const mf_flow_desc& d, int gemm, long long* offset, int* features_per_chunk);

#ifdef __CUDACC__
template <typename T> struct FM;
template <> struct FM<float> {
  static __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __device__ __forceinline__ float log_(float x) { return logf(x); }
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float floor_(float x) { return floorf(x); }
};
template <> struct FM<double> {
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double floor_(double x) { return floor(x); }
};

// spline constants in the compute dtype
template <typename T>
struct SplineCfg {
  int K;
  int univariate;
  T bound;
  T inv_ls2;  // 2 / ln(slope)  (0: soft clip disabled)
  T inv_ls1;  // 1 / ln(slope)
};

template <typename T>
__device__ __forceinline__ T softclip(T v, T inv) {
  return v / (T(1) + FM<T>::abs_(v * inv));
}

// torch.remainder(x + b, 2b) - b   (zuko CircularShiftTransform)
template <typename T>
__device__ __forceinline__ T circular_wrap(T x, T b) {
  T t = x + b, m = T(2) * b;
  T r = t - FM<T>::floor_(t / m) * m;
  return r - b;
}

// One univariate transform of one feature.  p = [K widths | K heights | K-1 derivatives] (unconstrained
// hyper-network outputs) with element stride `ps`; they are overwritten with scratch values.
// inverse=false: MonotonicRQSTransform.call_and_ladj;  inverse=true: MonotonicRQSTransform._inverse.
template <typename T>
__device__ __forceinline__ void rqs_univariate(T* p, int ps, const SplineCfg<T>& cfg, bool inverse,
                                               T v, T& out, T& ladj, int& bin, bool& inside) {
  const int K = cfg.K;
  using M = FM<T>;
  T ladj0 = T(0);
  if (!inverse) {
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) v = circular_wrap(v, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) { v = (v + T(1)) / T(2); ladj0 = T(-0.6931471805599453); }
  }
  T* pw = p;
  T* ph = p + (size_t)K * ps;
  T* pd = p + (size_t)2 * K * ps;
  // softmax numerators (soft-clipped logits), kept in place
  T mw = -INFINITY, mh = -INFINITY;
  for (int j = 0; j < K; ++j) {
    T a = softclip(pw[j * ps], cfg.inv_ls2), b = softclip(ph[j * ps], cfg.inv_ls2);
    pw[j * ps] = a; ph[j * ps] = b;
    mw = a > mw ? a : mw; mh = b > mh ? b : mh;
  }
  T sw = T(0), sh = T(0);
  for (int j = 0; j < K; ++j) {
    T a = M::exp_(pw[j * ps] - mw), b = M::exp_(ph[j * ps] - mh);
    pw[j * ps] = a; ph[j * ps] = b;
    sw += a; sh += b;
  }
  // knots: bound * (2 cumsum(softmax) - 1); k = #(knots < v) - 1 on the searched axis
  T* ps_search = inverse ? ph : pw;
  T* ps_other = inverse ? pw : ph;
  const T ss = inverse ? sh : sw, so = inverse ? sw : sh;
  int count = 0;
  {
    T cum = T(0);
    T knot = cfg.bound * (T(2) * cum - T(1));
    count += (knot < v) ? 1 : 0;
    for (int j = 0; j < K; ++j) {
      cum += ps_search[j * ps] / ss;
      knot = cfg.bound * (T(2) * cum - T(1));
      count += (knot < v) ? 1 : 0;
    }
  }
  int k = count - 1;
  inside = (k >= 0) && (k < K);
  k = k < 0 ? k + K : (k >= K ? k - K : k);  // k mod K
  bin = k;
  // gather the two knots of bin k on both axes (same summation order as the scan above)
  T s0 = T(0), s1, o0 = T(0), o1;
  for (int j = 0; j < k; ++j) { s0 += ps_search[j * ps] / ss; o0 += ps_other[j * ps] / so; }
  s1 = s0 + ps_search[k * ps] / ss;
  o1 = o0 + ps_other[k * ps] / so;
  const T a0 = cfg.bound * (T(2) * s0 - T(1)), a1 = cfg.bound * (T(2) * s1 - T(1));  // searched axis
  const T b0 = cfg.bound * (T(2) * o0 - T(1)), b1 = cfg.bound * (T(2) * o1 - T(1));  // other axis
  const T x0 = inverse ? b0 : a0, x1 = inverse ? b1 : a1;
  const T y0 = inverse ? a0 : b0, y1 = inverse ? a1 : b1;
  const T d0 = (k == 0) ? T(1) : M::exp_(softclip(pd[(k - 1) * ps], cfg.inv_ls1));
  const T d1 = (k == K - 1) ? T(1) : M::exp_(softclip(pd[k * ps], cfg.inv_ls1));
  const T s = (y1 - y0) / (x1 - x0);
  const T msk = inside ? T(1) : T(0);
  if (!inverse) {
    const T z = msk * (v - x0) / (x1 - x0);
    const T omz = T(1) - z;
    const T den = s + (d0 + d1 - T(2) * s) * z * omz;
    const T y = y0 + (y1 - y0) * (s * z * z + d0 * z * omz) / den;
    const T jac = s * s * (T(2) * s * z * omz + d0 * omz * omz + d1 * z * z) / (den * den);
    out = inside ? y : v;
    ladj = msk * M::log_(jac) + ladj0;
  } else {
    const T y_ = msk * (v - y0);
    const T a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - T(2) * s);
    const T b = (y1 - y0) * d0 - y_ * (d0 + d1 - T(2) * s);
    const T c = -s * y_;
    const T z = T(2) * c / (-b - M::sqrt_(b * b - T(4) * a * c));
    T x = x0 + z * (x1 - x0);
    x = inside ? x : v;
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) x = circular_wrap(x, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) x = T(2) * x - T(1);
    out = x;
    ladj = T(0);
  }
}

// mf_flow_xmap on one value of feature f (see include/memflow_b200.h)
template <typename T>
__device__ __forceinline__ T xmap_apply(const BaseParams& bp, T v, int f) {
  if (bp.xmap_kind == MF_XMAP_SIGMOID_AFFINE) {
    const T t = v * (T)bp.xmap_scale[f] + (T)bp.xmap_shift[f];
    return T(1) / (T(1) + FM<T>::exp_(-t));
  }
  if (bp.xmap_kind == MF_XMAP_LOGIT_AFFINE) {
    const T e = (T)bp.xmap_eps;
    T c = v < e ? e : (v > T(1) - e ? T(1) - e : v);   // torch.logit(x, eps): clamp to [eps, 1 - eps]
    const T l = FM<T>::log_(c / (T(1) - c));
    return (l - (T)bp.xmap_shift[f]) / (T)bp.xmap_scale[f];
  }
  return v;
}

// base.log_prob contribution of one feature
template <typename T>
__device__ __forceinline__ T base_log_prob_1(int base, T z, T a, T b) {
  if (base == MF_BASE_DIAG_NORMAL) {
    const T d = z - a;
    return -(d * d) / (T(2) * b * b) - FM<T>::log_(b) - T(0.9189385332046727);  // log sqrt(2 pi)
  }
  const bool in = (a <= z) && (b > z);  // torch.distributions.Uniform.log_prob
  return (in ? T(0) : T(-INFINITY)) - FM<T>::log_(b - a);
}
#endif  // __CUDACC__

}  // namespace mf
