// flow_tc.cu -- fused conditional RQ-spline MAF log_prob with the conditioner GEMMs on tcgen05.
//
// One persistent CTA per SM works on PAIRS of 128-object tiles.  Warp roles (352 threads):
//   warp 0        weight loader: streams the packed, pre-swizzled weight K-blocks from L2 into a 3-stage
//                 shared-memory ring with 1-D bulk async copies (TMA engine) completing on mbarriers
//   warps 1, 2    MMA issuers of tile slot 0 / 1: one elected lane issues tcgen05.mma (M=128, N<=128, K=16,
//                 fp32 accumulate in TMEM) and tcgen05.commit; both consume every weight stage (measured: one
//                 issuer sustains ~1 MMA per 105 cycles, the tensor core needs 64 -> two issuers)
//   warps 3-6     epilogue group of tile slot 0 (thread = object = TMEM lane)
//   warps 7-10    epilogue group of tile slot 1
// The two tile slots ping-pong: while the tensor core works on one tile's layer, the other tile's epilogue
// (TMEM accumulator -> registers -> bias + ReLU -> fp16 hi/lo operand planes -> back into TMEM with tcgen05.st)
// runs on the CUDA cores.  Hidden activations therefore never leave TMEM: layer l+1 takes its A operand straight
// from TMEM (tcgen05.mma with a TMEM A operand), only the layer-0 operand (x | context) sits in shared memory,
// where the context part is staged ONCE per tile and only the D columns of x are refreshed per transform.
// The spline parameters of the last layer are consumed straight out of TMEM by the thread that owns the
// object, so they never touch shared memory or HBM either.
// TMEM map of tile slot t (256 columns): [0,128) fp32 accumulator, [128,192) hi plane, [192,256) lo plane
// (64 columns = 128 packed 16-bit K elements).
//
// Numerics.  MF_GEMM_TC_3XF16: every fp32 operand value v is split into fp16 hi = rn(v), lo = rn(v - hi) and
// each product is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation: 22 significant bits, i.e. fp32-class
// accuracy (measured 4e-7 of sum|terms|, tests/test_tc_probe_gpu.py).  MF_GEMM_TC_BF16: single bf16 MMA.
//
// Supported shapes (others are served by the CUDA-core kernel in flow.cu): D + C <= 128, every hidden width
// <= 128, 3 * roundup(K, 8) <= 128.
#include <algorithm>

#include "flow_common.cuh"
#include "spline_reg.cuh"
#include "tc_common.cuh"

namespace mf {
namespace {

// Optional cycle-level trace of CTA 0 / first tile pair (build with -DMF_TC_TRACE): events are written as
// (clock64, id) pairs into the `inside` debug buffer.  ids: 100+ji MMA sees a_ready, 200+ji MMA issued,
// 300+ji epilogue sees acc_full, 400+ji epilogue done (tile slot 0 only).
#ifdef MF_TC_TRACE
#define MF_TRACE(cond, id)                                                          \
  do {                                                                              \
    if ((cond) && blockIdx.x == 0 && trace_n < 4000) {                              \
      long long* tb_ = reinterpret_cast<long long*>(args.inside);                   \
      tb_[2 * trace_n] = clock64(); tb_[2 * trace_n + 1] = (id); ++trace_n;         \
    }                                                                               \
  } while (0)
#else
#define MF_TRACE(cond, id) do { } while (0)
#endif

constexpr int kMaxStages = 8;                   // weight ring depth: as many as fit next to the layer-0 operands
constexpr int kMaxJobs = 40;
constexpr uint32_t kWPartBytes = 128 * 128;     // one [<=128 rows x 64] 16-bit weight K-block
constexpr uint32_t kKBlockBytes = 128 * 128;    // one [128 x 64] 16-bit SWIZZLE_128B K-block of the layer-0 operand
constexpr uint32_t kTailBytes = 128 * 32;       // one [128 x 16] 16-bit SWIZZLE_32B k-step of the layer-0 operand
constexpr size_t kStaticSmem = 4096;            // barriers, bias slots (static __shared__), alignment slack

struct TcJob {
  int N;        // MMA N (multiple of 16, <= 128)
  int ksteps;   // K / 16
  int nkb;      // K-blocks of 64
  int needs_a;  // a new A operand is produced right before this job
  int last;     // 1: spline chunk of the last layer
  int feat0, nfeat;
  int layer;    // source layer in zuko layout
  int row0;     // last layer: first zuko output row of the chunk is feat0 * (3K-1)
  int in_dim, out_dim;
  long long w_off;  // bytes from the transform base to K-block 0 ([hi block][lo block] per K-block)
  int b_off;        // floats from the transform bias base
};

struct TcLayout {
  int D, C, K, T, L, nsplit;
  int Kp;      // roundup(K, 8): widths at [0,Kp), heights at [Kp,2Kp), derivatives at [2Kp, 3Kp) of a feature
  int PFt;     // 3 * Kp
  int njobs;
  int in0_ksteps;
  int xcol0;             // layer-0 operand columns: context at [0, C), zeros, x at [xcol0, xcol0 + D), zeros
  int in0_full;          // layer-0 operand: full 64-wide K-blocks (SWIZZLE_128B) ...
  int in0_tail;          // ... followed by this many single k-steps in SWIZZLE_32B blocks
  int in0_plane_bytes;   // bytes of one operand plane of one tile
  int stages;            // weight ring depth
  long long t_stride;    // bytes per transform (weights)
  long long bias_off;    // byte offset of the fp32 bias region in the blob
  int bias_stride;       // floats per transform
  TcJob job[kMaxJobs];
};

int make_tc_layout(const mf_flow_desc& d, int gemm, TcLayout* L) {
  MF_REQUIRE(gemm == MF_GEMM_TC_3XF16 || gemm == MF_GEMM_TC_BF16, "not a tensor-core gemm mode: %d", gemm);
  MF_REQUIRE(d.features >= 1 && d.features <= MF_MAX_FEATURES && d.context >= 0, "bad features/context");
  MF_REQUIRE(d.transforms >= 1 && d.transforms <= MF_MAX_TRANSFORMS, "bad transforms");
  MF_REQUIRE(d.n_hidden >= 1 && d.n_hidden <= MF_MAX_HIDDEN_LAYERS, "bad n_hidden");
  L->D = d.features; L->C = d.context; L->K = d.bins; L->T = d.transforms; L->L = d.n_hidden;
  L->nsplit = (gemm == MF_GEMM_TC_3XF16) ? 3 : 1;
  L->Kp = round_up(d.bins, 8);
  L->PFt = 3 * L->Kp;
  L->xcol0 = round_up(d.context, 8);
  if (L->xcol0 + d.features > 128 || L->PFt > 128 || d.bins < 2) {
    set_error("tensor-core flow path needs roundup(C,8)+D <= 128 and 3*roundup(bins,8) <= 128 (got D=%d, C=%d, bins=%d)",
              d.features, d.context, d.bins);
    return MF_EUNSUPPORTED;
  }
  for (int l = 0; l < d.n_hidden; ++l)
    if (d.hidden[l] > 128 || d.hidden[l] < 1) {
      set_error("tensor-core flow path needs hidden widths <= 128 (layer %d has %d)", l, d.hidden[l]);
      return MF_EUNSUPPORTED;
    }
  const int parts = (L->nsplit == 3) ? 2 : 1;
  int nj = 0;
  long long woff = 0;
  int boff = 0;
  for (int l = 0; l < d.n_hidden; ++l) {
    TcJob& j = L->job[nj++];
    j.in_dim = (l == 0) ? d.features + d.context : d.hidden[l - 1];
    j.out_dim = d.hidden[l];
    j.N = round_up(j.out_dim, 16);
    j.ksteps = round_up(l == 0 ? L->xcol0 + d.features : j.in_dim, 16) / 16;
    j.nkb = (j.ksteps + 3) / 4;
    j.needs_a = 1; j.last = 0; j.feat0 = 0; j.nfeat = 0; j.layer = l; j.row0 = 0;
    j.w_off = woff; woff += (long long)j.nkb * parts * j.N * 128;
    j.b_off = boff; boff += j.N;
  }
  L->in0_ksteps = L->job[0].ksteps;
  L->in0_full = L->in0_ksteps / 4;
  L->in0_tail = L->in0_ksteps % 4;
  L->in0_plane_bytes = L->in0_full * (int)kKBlockBytes + L->in0_tail * (int)kTailBytes;
  {
    const size_t avail = 232448 - kStaticSmem - (size_t)2 * parts * L->in0_plane_bytes;
    L->stages = (int)std::min<size_t>(kMaxStages, avail / ((size_t)parts * kWPartBytes));
  }
  const int fpc = std::max(1, 128 / L->PFt);
  for (int f0 = 0; f0 < d.features; f0 += fpc) {
    MF_REQUIRE(nj < kMaxJobs, "too many last-layer chunks for the tensor-core path");
    TcJob& j = L->job[nj++];
    j.in_dim = d.hidden[d.n_hidden - 1];
    j.nfeat = std::min(fpc, d.features - f0);
    j.feat0 = f0;
    j.out_dim = j.nfeat * L->PFt;
    j.N = round_up(j.out_dim, 16);
    j.ksteps = round_up(j.in_dim, 16) / 16;
    j.nkb = (j.ksteps + 3) / 4;
    j.needs_a = (f0 == 0) ? 1 : 0; j.last = 1; j.layer = d.n_hidden; j.row0 = f0 * (3 * d.bins - 1);
    j.w_off = woff; woff += (long long)j.nkb * parts * j.N * 128;
    j.b_off = boff; boff += j.N;
  }
  L->njobs = nj;
  L->t_stride = woff;
  L->bias_off = woff * d.transforms;
  L->bias_stride = boff;
  return MF_OK;
}

// ------------------------------------------------------------------------------------------------
// pack: zuko-layout parameters -> swizzled fp16 hi/lo (or bf16) K-blocks + fp32 biases
template <typename T>
struct TcPackArgs {
  const T* w[kMaxLayers];
  const T* b[kMaxLayers];
  const uint8_t* m[kMaxLayers];
};

template <typename T>
__global__ void flow_tc_pack_kernel(TcLayout L, TcPackArgs<T> a, unsigned char* blob_t, float* bias_t) {
  const int parts = (L.nsplit == 3) ? 2 : 1;
  const int P = 3 * L.K - 1;
  for (int ji = 0; ji < L.njobs; ++ji) {
    const TcJob j = L.job[ji];
    const int kdim = j.nkb * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.N * kdim; i += gridDim.x * blockDim.x) {
      const int n = i / kdim, k = i % kdim;
      // source row in zuko layout
      int src_row = -1;
      if (!j.last) {
        if (n < j.out_dim) src_row = n;
      } else {
        const int fl = n / L.PFt, c = n % L.PFt;
        if (fl < j.nfeat) {
          const int seg = c / L.Kp, jj = c % L.Kp;  // 0 widths, 1 heights, 2 derivatives
          const int len = (seg == 2) ? L.K - 1 : L.K;
          if (jj < len) src_row = (j.feat0 + fl) * P + seg * L.K + jj;
        }
      }
      float v = 0.f;
      int src_col = k;  // layer 0: operand column -> zuko input column (x | context)
      if (ji == 0) src_col = (k < L.C) ? L.D + k : ((k >= L.xcol0 && k < L.xcol0 + L.D) ? k - L.xcol0 : -1);
      if (src_row >= 0 && src_col >= 0 && src_col < j.in_dim) {
        const long long s = (long long)src_row * j.in_dim + src_col;
        v = a.m[j.layer][s] ? (float)a.w[j.layer][s] : 0.f;
      }
      const int kb = k >> 6;
      unsigned char* blk = blob_t + j.w_off + (long long)kb * parts * j.N * 128;
      const uint32_t off = tc::sw128_offset(n, k & 63, j.N);
      if (L.nsplit == 3) {
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(blk + off) = h;
        *reinterpret_cast<__half*>(blk + (size_t)j.N * 128 + off) = __float2half_rn(v - __half2float(h));
      } else {
        *reinterpret_cast<__nv_bfloat16*>(blk + off) = __float2bfloat16_rn(v);
      }
      if (k == 0) {
        float bv = 0.f;
        if (src_row >= 0) bv = (float)a.b[j.layer][src_row];
        bias_t[j.b_off + n] = bv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  tc::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Spline of one feature with its 3K-1 unconstrained parameters in TMEM (this thread's lane):
// columns [0,K) widths, [Kp,Kp+K) heights, [2Kp, 2Kp+K-1) derivatives.  Forward transform + log-det.
// sb: the parameters' biases in shared memory (same column layout).
__device__ __forceinline__ void ld8b(uint32_t taddr, const float* sb, float* v) {
  tmem_ld8(taddr, v);
  const float4 b0 = *reinterpret_cast<const float4*>(sb), b1 = *reinterpret_cast<const float4*>(sb + 4);
  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
}

__device__ __forceinline__ void rqs_forward_tmem(uint32_t taddr, const float* sb, const SplineCfg<float>& cfg, int Kp,
                                                 float v, float& out, float& ladj, int& bin, bool& inside) {
  const int K = cfg.K;
  const int nch = Kp >> 3;
  float ladj0 = 0.f;
  if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) v = circular_wrap(v, cfg.bound);
  if (cfg.univariate == MF_UNIV_PS_RQS) { v = (v + 1.f) * 0.5f; ladj0 = -0.6931471805599453f; }
  const bool clipped = cfg.inv_ls2 != 0.f;  // soft-clipped logits are bounded by |ln slope| / 2: no max pass
  float mw = 0.f, mh = 0.f;
  if (!clipped) {
    mw = -INFINITY; mh = -INFINITY;
    for (int c = 0; c < nch; ++c) {
      float a[8], b[8];
      ld8b(taddr + c * 8, sb + c * 8, a);
      ld8b(taddr + Kp + c * 8, sb + Kp + c * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c * 8 + j < K) {
          mw = fmaxf(mw, softclipf(a[j], cfg.inv_ls2));
          mh = fmaxf(mh, softclipf(b[j], cfg.inv_ls2));
        }
    }
  }
  float sw = 0.f, sh = 0.f;
  for (int c = 0; c < nch; ++c) {
    float a[8], b[8];
    ld8b(taddr + c * 8, sb + c * 8, a);
    ld8b(taddr + Kp + c * 8, sb + Kp + c * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c * 8 + j < K) {
        sw += __expf(softclipf(a[j], cfg.inv_ls2) - mw);
        sh += __expf(softclipf(b[j], cfg.inv_ls2) - mh);
      }
  }
  // scan the horizontal knots: bound * (2 cumsum(softmax(w)) - 1); k = #(knots < v) - 1
  int k = -1;
  float x0 = 0.f, x1 = 1.f, cum = 0.f, prev = -cfg.bound;
  for (int c = 0; c < nch; ++c) {
    float a[8];
    ld8b(taddr + c * 8, sb + c * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c * 8 + j < K) {
        cum += __expf(softclipf(a[j], cfg.inv_ls2) - mw) / sw;
        const float knot = cfg.bound * (2.f * cum - 1.f);
        if (prev < v) { k = c * 8 + j; x0 = prev; x1 = knot; }
        prev = knot;
      }
  }
  if (prev < v) k = K;
  inside = (k >= 0) && (k < K);
  k = k < 0 ? k + K : (k >= K ? k - K : k);
  bin = k;
  // vertical knots of bin k
  float part = 0.f, ek = 0.f;
  for (int c = 0; c < nch; ++c) {
    float b[8];
    ld8b(taddr + Kp + c * 8, sb + Kp + c * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int jj = c * 8 + j;
      if (jj <= k) {
        const float e = __expf(softclipf(b[j], cfg.inv_ls2) - mh) / sh;
        if (jj < k) part += e; else ek = e;
      }
    }
  }
  const float y0 = cfg.bound * (2.f * part - 1.f), y1 = cfg.bound * (2.f * (part + ek) - 1.f);
  // derivatives at the two knots (boundary derivatives are 1)
  float d0 = 1.f, d1 = 1.f;
  {
    // tcgen05.ld is a warp-collective (.sync.aligned) with a warp-uniform address: all derivative chunks
    // are loaded by every lane and each lane picks the two columns of its own bin in registers
    const int i0 = k - 1, i1 = k;  // indices into the K-1 derivative logits
    float t0 = 0.f, t1 = 0.f;
    for (int c = 0; c < nch; ++c) {
      float dv[8];
      ld8b(taddr + 2 * Kp + c * 8, sb + 2 * Kp + c * 8, dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        t0 = (c * 8 + j == i0) ? dv[j] : t0;
        t1 = (c * 8 + j == i1) ? dv[j] : t1;
      }
    }
    if (i0 >= 0) d0 = __expf(softclipf(t0, cfg.inv_ls1));
    if (i1 < K - 1) d1 = __expf(softclipf(t1, cfg.inv_ls1));
  }
  if (!inside) { x0 = 0.f; x1 = 1.f; }
  const float s = (y1 - y0) / (x1 - x0);
  const float msk = inside ? 1.f : 0.f;
  const float z = msk * (v - x0) / (x1 - x0);
  const float omz = 1.f - z;
  const float den = s + (d0 + d1 - 2.f * s) * z * omz;
  const float y = y0 + (y1 - y0) * (s * z * z + d0 * z * omz) / den;
  const float jac = s * s * (2.f * s * z * omz + d0 * omz * omz + d1 * z * z) / (den * den);
  out = inside ? y : v;
  ladj = (inside ? __logf(jac) : 0.f) + ladj0;
}

// Same transform with all logits of the feature held in registers (Kp = 8 * NCH <= 32): every TMEM column is
// loaded once, every exponential is evaluated once, and the two knot scans run in one pass.
__device__ __forceinline__ void tmem_ld8_raw(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

template <int NCH>
__device__ __forceinline__ void rqs_forward_tmem_reg(uint32_t taddr, const float* sb, const SplineCfg<float>& cfg,
                                                     bool inverse, float v, float& out, float& ladj, int& bin,
                                                     bool& inside, long long* ts = nullptr) {
  constexpr int KP = 8 * NCH;
  float lw[KP], lh[KP];
  if (ts) ts[0] = clock64();
  {
    uint32_t rw[KP], rh[KP];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tmem_ld8_raw(taddr + c * 8, rw + c * 8);
      tmem_ld8_raw(taddr + KP + c * 8, rh + c * 8);
    }
    tc::tmem_ld_wait();
    if (ts) ts[1] = clock64();
#pragma unroll
    for (int j = 0; j < KP; ++j) {  // accumulator + bias
      lw[j] = __uint_as_float(rw[j]) + sb[j];
      lh[j] = __uint_as_float(rh[j]) + sb[KP + j];
    }
  }
  int k;
  float x0, x1, y0, y1, ladj0, t0, t1;
  rqs_scan_reg<NCH>(lw, lh, cfg, inverse, v, k, inside, x0, x1, y0, y1, ladj0);
  if (ts) ts[2] = clock64() + (long long)(x0 == 12345.f);
  {  // derivative logits are fetched only now (short register live ranges); the load is warp-uniform
    uint32_t rd[KP];
#pragma unroll
    for (int c = 0; c < NCH; ++c) tmem_ld8_raw(taddr + 2 * KP + c * 8, rd + c * 8);
    tc::tmem_ld_wait();
    t0 = 0.f; t1 = 0.f;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      const float dj = __uint_as_float(rd[j]) + sb[2 * KP + j];
      t0 = (j == k - 1) ? dj : t0;
      t1 = (j == k) ? dj : t1;
    }
  }
  bin = k;
  if (ts) ts[3] = clock64() + (long long)(t0 == 12345.f);
  rqs_finish(cfg, inverse, v, k, inside, x0, x1, y0, y1, t0, t1, ladj0, out, ladj);
  if (ts) ts[4] = clock64() + (long long)(out == 12345.f);
}

// ------------------------------------------------------------------------------------------------
// EW = epilogue warp sets per tile slot: with EW = 2 two warps share every TMEM lane quarter and split the
// columns of a hidden layer / the features of a spline chunk between them (16 epilogue warps per CTA).
// INV: sampling (transforms in reverse, `passes` sweeps each) -- a compile-time flag so that the log_prob kernel carries
// none of the sweep / schedule logic
template <int NSPLIT, int EW, bool INV>
__global__ void __launch_bounds__((3 + 8 * EW) * 32, 1)
flow_tc_kernel(const __grid_constant__ TcLayout L, const __grid_constant__ FlowKernelArgs args,
               const __grid_constant__ BaseParams bp) {
  constexpr int kParts = (NSPLIT == 3) ? 2 : 1;           // operand planes (hi, lo)
  constexpr uint32_t kStageBytes = kParts * kWPartBytes;  // one weight K-block, all planes
  extern __shared__ __align__(1024) unsigned char smem[];
  // [tile][part] activation operands, then the weight ring
  unsigned char* a_buf = smem;
  const uint32_t kPlane = (uint32_t)L.in0_plane_bytes;
  const uint32_t kFullBytes = (uint32_t)L.in0_full * kKBlockBytes;  // tail k-steps start here inside a plane
  unsigned char* w_ring = smem + 2 * kParts * kPlane;
  const uint32_t kStages = (uint32_t)L.stages;
  __shared__ __align__(8) uint64_t w_full[kMaxStages], w_empty[kMaxStages], a_ready[2], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sbias[2][2][128];  // [tile slot][job parity][column]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n = args.n;
  const long long n_tiles = (n + 127) / 128;
  const long long n_pairs = (n_tiles + 1) / 2;
  constexpr bool inverse = INV;
  const int sweeps = inverse ? args.passes : 1;
  const int n_units = L.T * sweeps;                    // one unit = one hyper-network evaluation of a tile
  const unsigned char* blob = static_cast<const unsigned char*>(args.blob);
  const float* bias_all = reinterpret_cast<const float*>(blob + L.bias_off);
  // sampling schedule (mf_flow_set_sample_schedule): in sweep s only the last-layer chunks holding a feature that becomes
  // final there are evaluated; every role walks the same (unit, job) sequence and skips the same jobs
  const unsigned long long* sched = inverse ? args.sched : nullptr;
  auto job_live = [&](int unit, int ji) -> bool {
    if (sched == nullptr || !L.job[ji].last) return true;
    const int tr = L.T - 1 - unit / sweeps;
    return ((sched[(size_t)tr * MF_MAX_FEATURES + (unit % sweeps)] >> L.job[ji].feat0) & ((1ull << L.job[ji].nfeat) - 1ull)) != 0ull;
  };

  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&a_ready[t], 128 * EW);
      mbar_init(&acc_full[t], 1);
      mbar_init(&acc_free[t], 128 * EW);
    }
    fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===================== weight loader =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (long long p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int unit = 0; unit < n_units; ++unit) {
          const int tr = inverse ? (L.T - 1 - unit / sweeps) : unit;
          const unsigned char* tb = blob + (long long)tr * L.t_stride;
          for (int ji = 0; ji < L.njobs; ++ji) {
            if (!job_live(unit, ji)) continue;
            const TcJob& j = L.job[ji];
            const uint32_t part_bytes = (uint32_t)j.N * 128u;
            for (int kb = 0; kb < j.nkb; ++kb, ++it) {
              const uint32_t s = it % kStages, ph = (it / kStages) & 1;
              mbar_wait_relaxed(&w_empty[s], ph ^ 1, 2000);
              mbar_expect_tx(&w_full[s], kParts * part_bytes);
              const unsigned char* src = tb + j.w_off + (long long)kb * kParts * part_bytes;
              unsigned char* dst = w_ring + s * kStageBytes;
#pragma unroll
              for (int q = 0; q < kParts; ++q) bulk_g2s(dst + q * kWPartBytes, src + (size_t)q * part_bytes, part_bytes, &w_full[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One warp walks the (warp-uniform) schedule and waits on the barriers; per K-block one elected lane issues
    // the MMAs back to back.  The two tile slots take turns job by job (slot 0's job j, slot 1's job j, slot 0's
    // job j+1, ...): while one slot's MMAs occupy the tensor pipe the other slot's epilogue drains its
    // accumulator, so in steady state the pipe never waits for an epilogue that is shorter than an MMA phase.
    uint32_t it = 0;                 // weight stages consumed so far
    int trace_n = 0; (void)trace_n;
    uint32_t jobctr[2] = {0, 0};     // accumulator uses per tile slot
    uint32_t aph[2] = {0, 0};        // a_ready phases
    for (long long p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      const int nact = (2 * p + 1 < n_tiles) ? 2 : 1;  // odd tile count: slot 1 idles in the last pair
      for (int unit = 0; unit < n_units; ++unit) {
        bool first_last = true;  // the first EXECUTED last-layer chunk of a unit waits for the last hidden activations
        for (int ji = 0; ji < L.njobs; ++ji) {
          if (!job_live(unit, ji)) continue;
          const TcJob& j = L.job[ji];
          const uint32_t idesc = tc::make_idesc_f16(NSPLIT == 3 ? 0 : 1, 128, j.N);
          const bool wait_a = j.last ? first_last : (j.needs_a != 0);
          if (j.last) first_last = false;
          for (int t = 0; t < nact; ++t) {
            const uint32_t d_tmem = tmem + (uint32_t)(t * 256);
            const uint32_t a_tmem = d_tmem + 128;  // hi plane; lo plane 64 columns further
            const uint32_t a_base = smem_u32(a_buf + (size_t)t * kParts * kPlane);
            if (wait_a) { mbar_wait(&a_ready[t], aph[t]); aph[t] ^= 1; }
            if (jobctr[t] >= 1) mbar_wait(&acc_free[t], (jobctr[t] - 1) & 1);  // accumulator drained
            MF_TRACE(lane == 0, 100 + 10 * t + ji);
            tc::fence_after_thread_sync();
            for (int kb = 0; kb < j.nkb; ++kb) {
              const uint32_t s = (it + kb) % kStages, ph = ((it + kb) / kStages) & 1;
              if (t == 0) mbar_wait(&w_full[s], ph);
              tc::fence_after_thread_sync();
              const uint32_t w_base = smem_u32(w_ring + s * kStageBytes);
              const int ks_n = min(4, j.ksteps - kb * 4);
              // measured: 75 clk per 128x128x16 MMA issued like this against ~120 when every MMA sits in its own
              // elect / branch / reconverge region
              if (tc::elect_one()) {
                const uint64_t dbh = tc::make_smem_desc(w_base);
                const uint64_t dbl = tc::make_smem_desc(w_base + kWPartBytes);
                if (ji == 0) {  // layer 0: (x | ctx) operand from shared memory
                  const bool full = kb < L.in0_full;
                  const uint64_t dah = full ? tc::make_smem_desc(a_base + kb * kKBlockBytes) : tc::make_smem_desc_sw32(a_base + kFullBytes);
                  const uint64_t dal = full ? tc::make_smem_desc(a_base + kPlane + kb * kKBlockBytes)
                                            : tc::make_smem_desc_sw32(a_base + kPlane + kFullBytes);
                  const uint64_t astep = full ? 2u : (kTailBytes >> 4);  // descriptor address units of 16 bytes
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    if (ks < ks_n) {
                      tc::mma_f16_ss(d_tmem, dah + ks * astep, dbh + ks * 2u, idesc, (kb | ks) ? 1u : 0u);
                      if (NSPLIT == 3) {
                        tc::mma_f16_ss(d_tmem, dah + ks * astep, dbl + ks * 2u, idesc, 1u);
                        tc::mma_f16_ss(d_tmem, dal + ks * astep, dbh + ks * 2u, idesc, 1u);
                      }
                    }
                  }
                } else {        // hidden activations: operand planes in TMEM
                  const uint32_t ah = a_tmem + (uint32_t)(kb * 4) * 8u;
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) {
                    if (ks < ks_n) {
                      tc::mma_f16_ts(d_tmem, ah + ks * 8u, dbh + ks * 2u, idesc, (kb | ks) ? 1u : 0u);
                      if (NSPLIT == 3) {
                        tc::mma_f16_ts(d_tmem, ah + ks * 8u, dbl + ks * 2u, idesc, 1u);
                        tc::mma_f16_ts(d_tmem, ah + 64u + ks * 8u, dbh + ks * 2u, idesc, 1u);
                      }
                    }
                  }
                }
                if (t == nact - 1) tc::mma_commit(&w_empty[s]);  // both slots are done with the stage
                if (kb == j.nkb - 1) tc::mma_commit(&acc_full[t]);
              }
              __syncwarp();
            }
            MF_TRACE(lane == 0, 200 + 10 * t + ji);
            jobctr[t]++;
          }
          it += j.nkb;
        }
      }
    }
  } else if (warp == 2) {
    // spare warp (keeps the epilogue warps aligned to the TMEM lane quarters: warp % 4)
  } else {
    // ===================== epilogue groups =====================
    const int eset = (warp - 3) >> 3;         // warp set inside the tile's epilogue group (0 .. EW-1)
    const int t = ((warp - 3) & 7) >> 2;      // tile slot
    const int q = warp & 3;                   // TMEM lane quarter this warp may access (4 consecutive warps cover all)
    const int row = q * 32 + lane;            // object inside the tile
    const int gt = eset * 128 + ((warp - 3) & 3) * 32 + lane;  // thread index inside the group (cooperative loops)
    constexpr int GT = 128 * EW;              // threads per epilogue group
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    unsigned char* a_hi = a_buf + (size_t)t * kParts * kPlane;
    unsigned char* a_lo = a_hi + kPlane;
    const int kfull_cols = L.in0_full * 64;
    auto in0_offset = [&](int r, int k) -> uint32_t {
      return k < kfull_cols ? tc::sw128_offset(r, k, 128)
                            : kFullBytes + (uint32_t)((k - kfull_cols) >> 4) * kTailBytes + tc::sw32_offset(r, k & 15);
    };
    const float* gx = static_cast<const float*>(args.x);
    const float* gc = static_cast<const float*>(args.ctx);
    float* gz = static_cast<float*>(args.z);
    const int D = L.D, C = L.C, xcol0 = L.xcol0;
    SplineCfg<float> cfg;
    cfg.K = L.K; cfg.univariate = args.univariate; cfg.bound = (float)args.bound;
    cfg.inv_ls2 = args.slope > 0.0 ? (float)(2.0 / log(args.slope)) : 0.f;
    cfg.inv_ls1 = args.slope > 0.0 ? (float)(1.0 / log(args.slope)) : 0.f;
    uint32_t jobctr = 0;
    int trace_n = 2000 + 1000 * t; (void)trace_n;

    auto put1 = [&](int r, int k, float v) {
      const uint32_t off = in0_offset(r, k);
      if (NSPLIT == 3) {
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(a_hi + off) = h;
        *reinterpret_cast<__half*>(a_lo + off) = __float2half_rn(v - __half2float(h));
      } else {
        *reinterpret_cast<__nv_bfloat16*>(a_hi + off) = __float2bfloat16_rn(v);
      }
    };

    for (long long p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      const long long tile = 2 * p + t;
      if (tile >= n_tiles) break;  // odd tile count: slot 1 idles in the last pair (the MMA warp skips it too)
      const long long tile0 = tile * 128;
      const int rows = (int)min((long long)128, n - tile0);
      const bool valid = row < rows;
      const long long g = tile0 + row;
      // working copy of x lives in the output buffer (each thread touches only its own row); when sampling,
      // gx holds the noise and the base sample a + b * noise is the starting point
      MF_TRACE(gt == 0, 950 + 10 * t);
      if (eset == 0)
        for (int f0 = 0; f0 < D; f0 += 8) {  // eight loads in flight
          float xv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) xv[u] = (valid && f0 + u < D) ? gx[g * D + f0 + u] : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int f = f0 + u;
            if (f < D) {
              float v0 = xv[u];
              if (inverse && valid)
                v0 = (args.base == MF_BASE_DIAG_NORMAL) ? (float)bp.a[f] + (float)bp.b[f] * v0
                                                       : (float)bp.a[f] + ((float)bp.b[f] - (float)bp.a[f]) * v0;
              if (valid) gz[g * D + f] = v0;
              put1(row, xcol0 + f, v0);  // x columns of the layer-0 operand
            }
          }
        }
      float* gy = static_cast<float*>(args.work);  // sampling: target y of the transform being inverted
      float ladj = 0.f;
      MF_TRACE(gt == 0, 951 + 10 * t);
      // ---- stage (ctx | 0) of the layer-0 operand once per tile; every MMA that read the previous tile's
      //      operand has retired (this group waited on the last accumulator of that tile)
      {
        const int kfill = L.in0_ksteps * 16;
        if (eset == 0)
          for (int k = xcol0 + D; k < kfill; ++k) put1(row, k, 0.f);
        // context: one 16-byte operand chunk (8 columns of one row) per thread and step, four chunks in flight;
        // consecutive threads take consecutive chunks, so a warp reads a contiguous stretch of global memory
        const int nchunk = xcol0 >> 3;  // chunks per row (the last one zero-padded when C % 8 != 0)
        const int total = 128 * nchunk;
        const float* gct = gc + tile0 * C;
        for (int q0 = gt; q0 < total; q0 += 4 * GT) {
          float v[4][8];
          int rr[4], kc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int q = q0 + u * GT;
            rr[u] = q / nchunk;
            kc[u] = q - rr[u] * nchunk;
            const bool live = (q < total) && (rr[u] < rows);
            const float* src = gct + (long long)rr[u] * C + kc[u] * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[u][e] = (live && kc[u] * 8 + e < C) ? __ldg(src + e) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (q0 + u * GT < total) {
              const int r = rr[u], c8 = kc[u];
              const uint32_t off = (c8 * 8 < kfull_cols)
                                       ? (uint32_t)(c8 >> 3) * kKBlockBytes + (uint32_t)r * 128u + (uint32_t)(((c8 & 7) ^ (r & 7)) << 4)
                                       : kFullBytes + (uint32_t)((c8 * 8 - kfull_cols) >> 4) * kTailBytes + (uint32_t)r * 32u +
                                             (uint32_t)(((c8 & 1) ^ ((r >> 2) & 1)) << 4);
              uint4 hi, lo;
              if (NSPLIT == 3) {
                tc::split_f16(v[u][0], v[u][1], hi.x, lo.x); tc::split_f16(v[u][2], v[u][3], hi.y, lo.y);
                tc::split_f16(v[u][4], v[u][5], hi.z, lo.z); tc::split_f16(v[u][6], v[u][7], hi.w, lo.w);
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
                *reinterpret_cast<uint4*>(a_lo + off) = lo;
              } else {
                hi.x = tc::pack_bf16(v[u][0], v[u][1]); hi.y = tc::pack_bf16(v[u][2], v[u][3]);
                hi.z = tc::pack_bf16(v[u][4], v[u][5]); hi.w = tc::pack_bf16(v[u][6], v[u][7]);
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
              }
            }
          }
        }
      }
      MF_TRACE(gt == 0, 952 + 10 * t);
      {  // pull this slot's next tile (x rows and context rows) into L2 while this one is being processed
        const long long nt0 = (tile + 2 * (long long)gridDim.x) * 128;
        if (nt0 < n) {
          const long long nrows = min((long long)128, n - nt0);
          const char* px = reinterpret_cast<const char*>(gx + nt0 * D);
          const char* pc = reinterpret_cast<const char*>(gc + nt0 * C);
          const long long xb = nrows * D * 4, cb = nrows * C * 4;
          for (long long o = (long long)gt * 128; o < xb; o += (long long)GT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
          for (long long o = (long long)gt * 128; o < cb; o += (long long)GT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
        }
      }
      if (EW > 1) group_bar(1 + t, GT);  // x written by warp set 0 is read by the other set below
      float lp_base = 0.f;               // base log-density, accumulated as the last transform's outputs appear
      float bias_next;
      {
        const int tr0 = inverse ? (L.T - 1) : 0;
        bias_next = (gt < L.job[0].N) ? __ldg(bias_all + (size_t)tr0 * L.bias_stride + L.job[0].b_off + gt) : 0.f;
      }
      MF_TRACE(gt == 0, 953 + 10 * t);
      for (int unit = 0; unit < n_units; ++unit) {
        const int tr = inverse ? (L.T - 1 - unit / sweeps) : unit;
        if (inverse && (unit % sweeps) == 0 && eset == 0) {  // y <- x, x <- 0  (AutoregressiveTransform._inverse)
          for (int f = 0; f < D; ++f) {
            if (valid) { gy[g * D + f] = gz[g * D + f]; gz[g * D + f] = 0.f; }
            put1(row, xcol0 + f, 0.f);
          }
        }
        // ---- layer-0 operand: the context part was staged at the tile start, the D columns of x were written
        //      by this thread when it produced them (tile start / spline output of the previous unit)
        fence_proxy_async_smem();
        mbar_arrive(&a_ready[t]);
        for (int ji = 0; ji < L.njobs; ++ji) {
          if (!job_live(unit, ji)) continue;
          const TcJob& j = L.job[ji];
          const uint32_t jpar = jobctr & 1;
          ++jobctr;
          // this job's biases -> shared memory (double-buffered by job parity), overlapping the MMA
          float* sb = sbias[t][jpar];
          if (gt < j.N) sb[gt] = bias_next;  // fetched one job ahead (global latency off the critical path)
          {
            int ju = unit, jn = ji + 1;
            if (jn == L.njobs) { jn = 0; ++ju; }
            while (ju < n_units && !job_live(ju, jn)) {  // next job that is actually executed
              if (++jn == L.njobs) { jn = 0; ++ju; }
            }
            if (ju < n_units) {
              const int trn = inverse ? (L.T - 1 - ju / sweeps) : ju;
              const TcJob& nj = L.job[jn];
              bias_next = (gt < nj.N) ? __ldg(bias_all + (size_t)trn * L.bias_stride + nj.b_off + gt) : 0.f;
            }
          }
          group_bar(1 + t, GT);
          // inputs of this chunk's splines, fetched before the accumulator wait (at most 5 features per chunk)
          const float* xsrc = (inverse ? gy : gz) + g * D + j.feat0;
          float xnext = (j.last && eset < j.nfeat && valid) ? xsrc[eset] : 0.f;
          MF_TRACE(gt == 0, 500 + 300 * t + ji);
          mbar_wait(&acc_full[t], jpar);
          MF_TRACE(gt == 0, 300 + 300 * t + ji);
          tc::fence_after_thread_sync();
          const uint32_t taddr = tmem + lane_off + (uint32_t)(t * 256);
          if (!j.last) {
            // hidden layer: TMEM -> +bias -> ReLU -> operand planes of the next layer (in place)
            const int csplit = (EW == 1) ? j.N : min(j.N, ((j.N / 2 + 31) / 32) * 32);
            const int cbeg = eset == 0 ? 0 : csplit, cend = eset == 0 ? csplit : j.N;
            for (int c00 = cbeg; c00 < cend; c00 += 32) {
              float vv[32];
              tc::tmem_ld16(taddr + c00, vv);
              if (c00 + 16 < cend) tc::tmem_ld16(taddr + c00 + 16, vv + 16);
              tc::tmem_ld_wait();
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const int c0 = c00 + half * 16;
                if (c0 < cend) {
                  float* v = vv + half * 16;
#pragma unroll
                  for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + i);
                    v[i] = fmaxf(v[i] + b4.x, 0.f); v[i + 1] = fmaxf(v[i + 1] + b4.y, 0.f);
                    v[i + 2] = fmaxf(v[i + 2] + b4.z, 0.f); v[i + 3] = fmaxf(v[i + 3] + b4.w, 0.f);
                  }
                  // 16 values -> 8 packed 16-bit pairs per plane -> operand planes in TMEM
                  if (NSPLIT == 3) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) tc::split_f16(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                    tc::tmem_st8(taddr + 128 + c0 / 2, hi);
                    tc::tmem_st8(taddr + 192 + c0 / 2, lo);
                  } else {
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = tc::pack_bf16(v[2 * i], v[2 * i + 1]);
                    tc::tmem_st8(taddr + 128 + c0 / 2, pk);
                  }
                }
              }
            }
            tc::tmem_st_wait();
            tc::fence_before_thread_sync();
            mbar_arrive(&acc_free[t]);
            mbar_arrive(&a_ready[t]);
            MF_TRACE(gt == 0, 400 + 300 * t + ji);
          } else {
            // last-layer chunk: splines of the chunk's features straight out of TMEM
            const unsigned long long fmask = sched ? sched[(size_t)tr * MF_MAX_FEATURES + (unit % sweeps)] : ~0ull;
#pragma unroll 1
            for (int fl = eset; fl < j.nfeat; fl += EW) {  // one copy of the spline code (instruction cache)
              const int f = j.feat0 + fl;
              const float xv = xnext;
              if (fl + EW < j.nfeat && valid) xnext = xsrc[fl + EW];  // next feature's input, in flight during this spline
              if (!((fmask >> f) & 1ull)) continue;  // sampling: provisional feature of this sweep, keeps its value
              float out, la;
              int bin;
              bool ins;
              const uint32_t ta = taddr + fl * L.PFt;
              const float* sbf = sb + fl * L.PFt;
              switch (L.Kp) {
                case 8: rqs_forward_tmem_reg<1>(ta, sbf, cfg, inverse, xv, out, la, bin, ins); break;
#ifdef MF_TC_TRACE
                case 16: {
                  long long ts5[5];
                  rqs_forward_tmem_reg<2>(ta, sbf, cfg, inverse, xv, out, la, bin, ins, ts5);
                  if (gt == 0 && t == 0 && blockIdx.x == 0 && trace_n < 2900) {
                    long long* tb_ = reinterpret_cast<long long*>(args.inside);
                    for (int q5 = 0; q5 < 5; ++q5) { tb_[2 * trace_n] = ts5[q5]; tb_[2 * trace_n + 1] = 900 + q5; ++trace_n; }
                  }
                } break;
#else
                case 16: rqs_forward_tmem_reg<2>(ta, sbf, cfg, inverse, xv, out, la, bin, ins); break;
#endif
                case 24: rqs_forward_tmem_reg<3>(ta, sbf, cfg, inverse, xv, out, la, bin, ins); break;
                case 32: rqs_forward_tmem_reg<4>(ta, sbf, cfg, inverse, xv, out, la, bin, ins); break;
                default: rqs_forward_tmem(ta, sbf, cfg, L.Kp, xv, out, la, bin, ins); break;
              }
              put1(row, xcol0 + f, valid ? out : 0.f);  // next unit's layer-0 operand (its last reader, job 0, has retired)
              if (valid) {
                gz[g * D + f] = out;
                ladj += la;
                if (EW == 1 && !inverse && unit == n_units - 1)
                  lp_base += base_log_prob_1<float>(args.base, out, (float)bp.a[f], (float)bp.b[f]);
                if (!inverse && args.bins) args.bins[((size_t)tr * n + g) * D + f] = (int8_t)bin;
                if (!inverse && args.inside) args.inside[((size_t)tr * n + g) * D + f] = ins ? 1 : 0;
              }
            }
            tc::fence_before_thread_sync();
            mbar_arrive(&acc_free[t]);
            MF_TRACE(gt == 0, 400 + 300 * t + ji);
          }
        }
      }
      if (EW > 1 && !inverse) {  // the other warp set's share of ladj travels through the workspace
        float* wl = static_cast<float*>(args.work);
        if (eset == 1 && valid) wl[g] = ladj;
        group_bar(1 + t, GT);
        if (eset == 0 && valid) ladj += wl[g];
      }
      if (valid && !inverse && eset == 0) {
        if (args.ladj) static_cast<float*>(args.ladj)[g] = ladj;
        if (args.logprob) {
          float lp = lp_base;
          if (EW > 1)
            for (int f = 0; f < D; ++f) lp += base_log_prob_1<float>(args.base, gz[g * D + f], (float)bp.a[f], (float)bp.b[f]);
          static_cast<float*>(args.logprob)[g] = lp + ladj;
        }
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
static long long tc_sched_offset(const TcLayout& L) {  // sampling schedule after the biases, 8-byte aligned
  return (L.bias_off + (long long)L.bias_stride * L.T * 4 + 7) / 8 * 8;
}

int flow_tc_blob_bytes(const mf_flow_desc& d, int gemm, long long* bytes) {
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  *bytes = tc_sched_offset(L) + kSchedBytes;
  return MF_OK;
}

int flow_tc_sched_layout(const mf_flow_desc& d, int gemm, long long* offset, int* features_per_chunk) {
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  *offset = tc_sched_offset(L);
  *features_per_chunk = std::max(1, 128 / L.PFt);
  return MF_OK;
}

int flow_tc_pack(const mf_flow_desc& d, const void* const* w, const void* const* b, const uint8_t* const* m,
                 int gemm, void* blob, long long blob_bytes, cudaStream_t st) {
  // parameters are fp32 on this path
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  const long long need = tc_sched_offset(L) + kSchedBytes;
  MF_REQUIRE(blob_bytes >= need, "blob too small: %lld < %lld bytes", blob_bytes, need);
  const int nl = d.n_hidden + 1;
  for (int t = 0; t < L.T; ++t) {
    TcPackArgs<float> a;
    for (int l = 0; l < nl; ++l) {
      a.w[l] = (const float*)w[t * nl + l];
      a.b[l] = (const float*)b[t * nl + l];
      a.m[l] = m[t * nl + l];
      MF_REQUIRE(a.w[l] && a.b[l] && a.m[l], "null parameter pointer (transform %d layer %d)", t, l);
    }
    flow_tc_pack_kernel<float><<<148, 256, 0, st>>>(L, a, (unsigned char*)blob + (long long)t * L.t_stride,
                                                   reinterpret_cast<float*>((unsigned char*)blob + L.bias_off) +
                                                       (size_t)t * L.bias_stride);
    rc = post_launch("flow_tc_pack_kernel");
    if (rc != MF_OK) return rc;
  }
  MF_CUDA_CHECK(cudaMemsetAsync((char*)blob + tc_sched_offset(L), 0xFF, kSchedBytes, st));
  return MF_OK;
}

int tc_epilogue_sets() {  // MF_TC_EW=2 selects the 16-epilogue-warp variant (measured: no gain, kept for A/B runs)
  static const int ew = [] {
    const char* e = getenv("MF_TC_EW");
    return (e && e[0] == '2') ? 2 : 1;
  }();
  return ew;
}

long long flow_tc_workspace_bytes(const mf_flow_desc& d, long long n, int gemm, int op) {
  (void)gemm;
  if (op == MF_FLOW_OP_SAMPLE) return n * d.features * 4;  // y targets of the transform being inverted
  return n * 4;                                            // ladj share of the second epilogue warp set
}

int flow_tc_run(const mf_flow_desc& d, const FlowKernelArgs& a, const BaseParams& bp, int gemm, cudaStream_t st) {
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  MF_REQUIRE(a.z != nullptr, "tensor-core flow path needs the z / x output buffer (it is the working copy of x)");
  FlowKernelArgs a_run = a;
  if (a.mode == 1) a_run.sched = reinterpret_cast<const unsigned long long*>(static_cast<const char*>(a.blob) + tc_sched_offset(L));
  if (a.ctx_group > 1) {
    set_error("context grouping (ctx_group=%d) is implemented on the MF_GEMM_EXACT path only", a.ctx_group);
    return MF_EUNSUPPORTED;
  }
  if (a.mode == 1) {
    if (L.Kp > 32) {
      set_error("tensor-core sampling needs bins <= 32 (got %d); use MF_GEMM_EXACT", d.bins);
      return MF_EUNSUPPORTED;
    }
    MF_REQUIRE(a.work != nullptr && a.work_bytes >= flow_tc_workspace_bytes(d, a.n, gemm, MF_FLOW_OP_SAMPLE),
               "tensor-core sampling needs a workspace of mf_flow_workspace_bytes() bytes");
  }
  const int parts = (L.nsplit == 3) ? 2 : 1;
  MF_REQUIRE(L.stages >= 2, "layer-0 operand too large for the weight ring");
  const size_t smem = (size_t)2 * parts * L.in0_plane_bytes + (size_t)L.stages * parts * kWPartBytes;
  const long long n_pairs = ((a.n + 127) / 128 + 1) / 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<long long>(n_pairs, sms);
  if (a.mode == 0)
    MF_REQUIRE(a.work != nullptr && a.work_bytes >= a.n * 4, "tensor-core log_prob needs a workspace of mf_flow_workspace_bytes() bytes");
  const int ew = tc_epilogue_sets();
#define MF_TC_LAUNCH2(NS, EWV, INVV)                                                                                  \
  do {                                                                                                                \
    MF_CUDA_CHECK(cudaFuncSetAttribute(flow_tc_kernel<NS, EWV, INVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    flow_tc_kernel<NS, EWV, INVV><<<grid, (3 + 8 * EWV) * 32, smem, st>>>(L, a_run, bp);                              \
  } while (0)
#define MF_TC_LAUNCH(NS, EWV) do { if (a.mode == 1) MF_TC_LAUNCH2(NS, EWV, true); else MF_TC_LAUNCH2(NS, EWV, false); } while (0)
  if (L.nsplit == 3) { if (ew == 2) MF_TC_LAUNCH(3, 2); else MF_TC_LAUNCH(3, 1); }
  else               { if (ew == 2) MF_TC_LAUNCH(1, 2); else MF_TC_LAUNCH(1, 1); }
#undef MF_TC_LAUNCH2
#undef MF_TC_LAUNCH
  return post_launch("flow_tc_kernel");
}

}  // namespace mf
