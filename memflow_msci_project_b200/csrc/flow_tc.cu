// flow_tc.cu -- fused conditional RQ-spline MAF log_prob / sample with the conditioner GEMMs on tcgen05.
//
// One persistent CTA per SM works on PAIRS of 128-object tiles (two "tile slots").  Warp roles (320 threads):
//   warps 0-3     epilogue group of tile slot 0 (thread = object = TMEM lane; warp % 4 = TMEM lane quarter)
//   warps 4-7     epilogue group of tile slot 1
//   warp 8        weight loader: streams the packed, pre-swizzled weight K-blocks from L2 into a shared-memory ring
//                 (as many 32 KB stages as fit next to the layer-0 operands) with 1-D bulk async copies (TMA engine)
//                 completing on mbarriers
//   warp 9        MMA issuer: walks the job schedule, the two tile slots take turns job by job; per job ONE elected lane
//                 issues the tcgen05.mma (M=128, N<=128, K=16, kind::f16, fp32 accumulate in TMEM) of the whole job back
//                 to back and commits the stage / accumulator mbarriers
// The control warps carry the HIGHEST warp ids of their scheduler on purpose: the sub-partition arbiter prefers the
// highest warp id, and with the issuer as warp 1 it lost the issue slot to the epilogue warps of its scheduler for
// ~1000 cycles per job (per-warp trace profiles/trace_flow_tc_r2f.txt: dependencies met at t, batch issued at t + 1000).
// While the tensor core works on one slot's job, the other slot's epilogue (TMEM accumulator -> registers -> bias +
// ReLU -> fp16 hi/lo operand planes -> back into TMEM with tcgen05.st) runs on the CUDA cores.  Hidden activations never
// leave TMEM: layer l+1 takes its A operand straight from TMEM (tcgen05.mma, TS form); only the layer-0 operand
// (context | x) sits in shared memory, where the context part is staged ONCE per tile and only the D columns of x are
// refreshed per transform.  The spline parameters of the last layer are consumed straight out of TMEM by the thread
// that owns the object (spline_reg.cuh), so they never touch shared memory or HBM either.
// TMEM map of tile slot t (256 columns): [0,128) fp32 accumulator, [128,192) hi plane, [192,256) lo plane (64 columns =
// 128 packed 16-bit K elements).  Last-layer chunks whose parameters fit 64 columns (3 roundup(K,8) <= 64) alternate
// between the accumulator halves [0,64) / [64,128): the MMAs of chunk c+1 run while the splines of chunk c are evaluated.
//
// Numerics.  MF_GEMM_TC_3XF16: every fp32 operand value v is split into fp16 hi = rn(v), lo = rn(v - hi) and
// each product is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation: 22 significant bits, i.e. fp32-class
// accuracy (measured 4e-7 of sum|terms|, tests/test_tc_probe_gpu.py).  MF_GEMM_TC_BF16: single bf16 MMA.
//
// Supported shapes (others are served by flow_mma.cu / flow.cu): roundup(C,8) + D <= 128, every hidden width <= 128,
// 3 * roundup(K, 8) <= 128 (K <= 40), soft clip enabled (slope > 0).
#include <algorithm>

#include "flow_common.cuh"
#include "spline_reg.cuh"
#include "tc_common.cuh"

namespace mf {
namespace {

// Optional cycle-level trace of CTA 0 (build with -DMF_TC_TRACE): every traced warp writes (clock64, id) pairs into its
// own region of the `inside` debug buffer; id = warp * 1000 + code * 20 + job.  Codes: issuer 1 = weights landed,
// 2 = operand / accumulator dependencies met, 3 = batch issued; epilogue warps 5 = waiting for the accumulator,
// 6 = accumulator complete, 7 = done (arrived on the next barrier); 8/9 = tile start / tile staged.  tools/trace_tc.py.
#ifdef MF_TC_TRACE
#define MF_TRACE(cond, code, job)                                                    \
  do {                                                                               \
    if ((cond) && blockIdx.x == 0 && trace_n < trace_end) {                          \
      long long* tb_ = reinterpret_cast<long long*>(args.inside);                    \
      tb_[2 * trace_n] = clock64(); tb_[2 * trace_n + 1] = (threadIdx.x >> 5) * 1000 + (code) * 20 + (job); ++trace_n; \
    }                                                                                \
  } while (0)
#else
#define MF_TRACE(cond, code, job) do { } while (0)
#endif

constexpr int kTcThreads = 12 * 32;              // 8 epilogue warps + loader + one MMA issuer per tile slot + 1 idle
constexpr int kCtrlRegs = 104, kEpiRegs = 200;   // setmaxnreg: registers move from the control warps to the epilogue warps
constexpr int kLoaderWarp = 8, kIssuerWarp = 9;
constexpr int kMaxStages = 8;                   // weight ring depth: as many as fit next to the layer-0 operands
constexpr int kMaxJobs = 40;
constexpr uint32_t kWPartBytes = 128 * 128;     // one [<=128 rows x 64] 16-bit weight K-block
constexpr uint32_t kKBlockBytes = 128 * 128;    // one [128 x 64] 16-bit SWIZZLE_128B K-block of the layer-0 operand
constexpr uint32_t kTailBytes = 128 * 32;       // one [128 x 16] 16-bit SWIZZLE_32B k-step of the layer-0 operand
constexpr size_t kStaticSmem = 10240;
constexpr int kPktWords = 40;                   // 32-bit words of one job packet (struct JobPkt)           // barriers, per-warp bias slots (static __shared__), alignment slack

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
  int nstages;      // weight ring stages of the job: both K-blocks of a narrow job (N <= 64) share one stage
  int region;       // accumulator region (barrier pair) of a last-layer chunk; hidden layers use region 0 and all columns
  int acc_col;      // first accumulator column of the job (0 or 64)
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
  int pkt_table;         // 1: the job packets of all (transform, job) pairs are prepared once per CTA in shared memory
  int fpc;               // features per last-layer chunk
  int nregions;          // 2: last-layer chunks alternate between the accumulator halves
  long long t_stride;    // bytes per transform (weights)
  long long bias_off;    // byte offset of the fp32 bias region in the blob
  int bias_stride;       // floats per transform
  double slope;          // soft-clip slope: the last-layer biases are stored pre-scaled by the soft-clip factors
  TcJob job[kMaxJobs];
};

int make_tc_layout(const mf_flow_desc& d, int gemm, TcLayout* L) {
  MF_REQUIRE(gemm == MF_GEMM_TC_3XF16 || gemm == MF_GEMM_TC_BF16, "not a tensor-core gemm mode: %d", gemm);
  MF_REQUIRE(d.features >= 1 && d.features <= MF_MAX_FEATURES && d.context >= 0, "bad features/context");
  MF_REQUIRE(d.transforms >= 1 && d.transforms <= MF_MAX_TRANSFORMS, "bad transforms");
  MF_REQUIRE(d.n_hidden >= 1 && d.n_hidden <= MF_MAX_HIDDEN_LAYERS, "bad n_hidden");
  L->D = d.features; L->C = d.context; L->K = d.bins; L->T = d.transforms; L->L = d.n_hidden;
  L->slope = d.slope;
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
    j.needs_a = 1; j.last = 0; j.feat0 = 0; j.nfeat = 0; j.layer = l; j.row0 = 0; j.region = 0; j.acc_col = 0;
    j.nstages = (j.nkb == 1 || j.N <= 64) ? 1 : 2;
    j.w_off = woff; woff += (long long)j.nkb * parts * j.N * 128;
    j.b_off = boff; boff += j.N;
  }
  L->in0_ksteps = L->job[0].ksteps;
  L->in0_full = L->in0_ksteps / 4;
  L->in0_tail = L->in0_ksteps % 4;
  L->in0_plane_bytes = L->in0_full * (int)kKBlockBytes + L->in0_tail * (int)kTailBytes;
  L->stages = 0;  // fixed below, once the number of jobs (packet table size) is known
  // features per last-layer chunk: as many as fit the 128 accumulator columns; when that takes more than one chunk
  // and a feature fits 64 columns, chunks of <= 64 columns alternate between the two accumulator halves instead
  int fpc = std::max(1, 128 / L->PFt);
  L->nregions = 1;
  if (d.features > fpc && L->PFt <= 64) { fpc = 64 / L->PFt; L->nregions = 2; }
  L->fpc = fpc;
  for (int f0 = 0, ci = 0; f0 < d.features; f0 += fpc, ++ci) {
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
    j.region = (L->nregions == 2) ? (ci & 1) : 0; j.acc_col = j.region * 64;
    j.nstages = (j.nkb == 1 || j.N <= 64) ? 1 : 2;
    j.w_off = woff; woff += (long long)j.nkb * parts * j.N * 128;
    j.b_off = boff; boff += j.N;
  }
  L->njobs = nj;
  {
    // shared memory: layer-0 operands | weight ring | job-packet table [T][njobs] (when it fits next to >= 4 stages)
    const size_t avail = 232448 - kStaticSmem - (size_t)2 * parts * L->in0_plane_bytes;
    const size_t stage_bytes = (size_t)parts * kWPartBytes, table = (size_t)d.transforms * nj * kPktWords * 4;
    int st = (int)std::min<size_t>(kMaxStages, avail / stage_bytes);
    L->pkt_table = 0;
    if (st >= 2 && avail - (size_t)st * stage_bytes >= table) L->pkt_table = 1;
    else if (st > 4 && avail - (size_t)(st - 1) * stage_bytes >= table) { --st; L->pkt_table = 1; }
    L->stages = st;
  }
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

// masked weight of packed row n, operand column k of job j (0 for padding) and the zuko row it comes from
template <typename T>
__device__ __forceinline__ float tc_packed_weight(const TcLayout& L, const TcPackArgs<T>& a, const TcJob& j, int ji, int n, int k,
                                                  int* src_row_out) {
  const int P = 3 * L.K - 1;
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
  if (src_row_out) *src_row_out = src_row;
  return v;
}

template <typename T>
__global__ void flow_tc_pack_kernel(TcLayout L, TcPackArgs<T> a, unsigned char* blob_t, float* bias_t) {
  const int parts = (L.nsplit == 3) ? 2 : 1;
  for (int ji = 0; ji < L.njobs; ++ji) {
    const TcJob j = L.job[ji];
    const int kdim = j.nkb * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.N * kdim; i += gridDim.x * blockDim.x) {
      const int n = i / kdim, k = i % kdim;
      int src_row;
      const float v = tc_packed_weight(L, a, j, ji, n, k, &src_row);
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
        if (j.last) {
          // the spline works on logits pre-multiplied by 2 / ln(slope) (widths, heights) or 1 / ln(slope) (derivatives,
          // spline_reg.cuh): accumulator * inv + bias * inv, with the second product done here once
          const RqsScaled cf = make_rqs_scaled(L.K, 0, 1.0, L.slope);
          bv *= (((n % L.PFt) / L.Kp) == 2) ? cf.inv1 : cf.inv2;
        }
        bias_t[j.b_off + n] = bv;
      }
    }
  }
}

// Zero-block table: for job ji and k-step ks (16 operand columns) byte ks of word [ji] = (number of LEADING packed rows
// whose weights in these columns are all exactly zero) / 16, i.e. the MMA of that k-step only has to produce the
// accumulator columns from 16 * byte on (byte = N / 16: the whole k-step is skipped).  With the hidden units sorted by
// autoregressive degree (host mirror, _sort_hidden_units) the masked weight matrices are block-triangular and ~1/4 of
// the tensor work falls away; the skipped products are exact zeros, so the results do not change.  k-step 0 always
// covers every column (it initialises the accumulator).
template <typename T>
__global__ void flow_tc_skip_kernel(TcLayout L, TcPackArgs<T> a, unsigned long long* skip_t) {
  const int ji = blockIdx.x, ks = threadIdx.x;  // 8 threads per job
  if (ji >= L.njobs) return;
  const TcJob j = L.job[ji];
  int n0 = 0;
  if (ks > 0 && ks < j.ksteps) {
    for (; n0 < j.N; ++n0) {
      bool nz = false;
      for (int k = ks * 16; k < ks * 16 + 16; ++k) nz |= (tc_packed_weight(L, a, j, ji, n0, k, nullptr) != 0.f);
      if (nz) break;
    }
  }
  const unsigned long long byte = (unsigned long long)(n0 / 16) << (8 * ks);
  // combine the 8 bytes of the job (threads 0..7 of the block)
  __shared__ unsigned long long w[8];
  w[ks] = byte;
  __syncthreads();
  if (ks == 0) skip_t[ji] = w[0] | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7];
}

// ------------------------------------------------------------------------------------------------
// non-suspending wait (mbarrier.test_wait poll): for the MMA issuer, whose wake-up from a suspended try_wait on a
// barrier completed by plain thread arrivals took ~1000 cycles in the per-warp traces
__device__ __forceinline__ void mbar_poll(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// arrival of every thread of an epilogue warp (barrier count 128)
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar, int lane) {
  (void)lane;  // every thread arrives (count 128): measured 1 % faster than one elected arrival per warp behind a __syncwarp
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Spline of one feature with its 3K-1 unconstrained parameters in TMEM (this thread's lane): columns [0,K) widths,
// [KP,KP+K) heights, [2KP, 2KP+K-1) derivatives (KP = 8 NCH; padded columns hold exact zeros).
// sb: the parameters' biases in shared memory in the same column layout, PRE-SCALED at pack time by the soft-clip
// factors (widths / heights by inv2, derivatives by inv1), so that accumulator * inv + sb is the scaled logit of
// spline_reg.cuh.  (Reading them straight from the blob with ld.global.nc was measured slower, r2v: this kernel's
// shared-memory footprint leaves ~4 KB of L1, so every such load is an L2 round trip.)
__device__ __forceinline__ void tmem_ld8_raw(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// one aligned 32-byte sector of read-only global memory (SASS LDG.E.256.CONSTANT)
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
      : "l"(p));
}

// Cold paths of the tile start, kept out of line: the tile prologue runs once per ~150 k cycles, its code is never in
// the instruction cache when it starts, and all eight epilogue warps wait for the same cache lines (per-warp traces
// r2x-r3c: the context loop took 5.6 k cycles even with its global loads removed).
struct F8 { float v[8]; };
__device__ __noinline__ F8 ldg_sector_guarded(const float* sp, const float* end, bool live) {
  F8 r;
#pragma unroll
  for (int e = 0; e < 8; ++e) r.v[e] = (live && sp + e < end) ? __ldg(sp + e) : 0.f;
  return r;
}
__device__ __noinline__ float xmap_cold(const BaseParams& bp, float v, int f) { return xmap_apply<float>(bp, v, f); }

// 8 accumulator columns -> scaled logits: acc * inv + sb, two columns per FFMA2
__device__ __forceinline__ void scale8(const uint32_t* r, const float* sb, float inv, float* o) {
  const float4 b0 = *reinterpret_cast<const float4*>(sb), b1 = *reinterpret_cast<const float4*>(sb + 4);
  const float2 i2 = f2_bcast(inv);
  float2 t;
  t = f2_fma(make_float2(__uint_as_float(r[0]), __uint_as_float(r[1])), i2, make_float2(b0.x, b0.y)); o[0] = t.x; o[1] = t.y;
  t = f2_fma(make_float2(__uint_as_float(r[2]), __uint_as_float(r[3])), i2, make_float2(b0.z, b0.w)); o[2] = t.x; o[3] = t.y;
  t = f2_fma(make_float2(__uint_as_float(r[4]), __uint_as_float(r[5])), i2, make_float2(b1.x, b1.y)); o[4] = t.x; o[5] = t.y;
  t = f2_fma(make_float2(__uint_as_float(r[6]), __uint_as_float(r[7])), i2, make_float2(b1.z, b1.w)); o[6] = t.x; o[7] = t.y;
}

template <int NCH, bool INV>
__device__ __forceinline__ void spline_tmem(uint32_t taddr, const float* sb, const RqsScaled& cf, float v, float& out,
                                            float& ladj, int& bin, bool& inside) {
  constexpr int KP = 8 * NCH;
  float lw[KP], lh[KP];
  {
    uint32_t rw[KP], rh[KP];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      tmem_ld8_raw(taddr + c * 8, rw + c * 8);
      tmem_ld8_raw(taddr + KP + c * 8, rh + c * 8);
    }
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      scale8(rw + c * 8, sb + c * 8, cf.inv2, lw + c * 8);
      scale8(rh + c * 8, sb + KP + c * 8, cf.inv2, lh + c * 8);
    }
  }
  uint32_t rd[2][8];
  tmem_ld8_raw(taddr + 2 * KP, rd[0]);  // derivative logits: in flight during the exponentials, one chunk ahead after that
  float ladj0 = 0.f;
  if (!INV) v = rqs_premap(cf, v, ladj0);
  float sw, sh;
  rqs_ms_exp<NCH>(lw, lh, cf, sw, sh);
  RqsMsState st;
  rqs_ms_begin<INV>(st, cf, v, sw, sh);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    tc::tmem_ld_wait();
    if (c + 1 < NCH) tmem_ld8_raw(taddr + 2 * KP + (c + 1) * 8, rd[(c + 1) & 1]);
    float d8[8];
    scale8(rd[c & 1], sb + 2 * KP + c * 8, cf.inv1, d8);
    rqs_ms_chunk<INV>(st, lw + c * 8, lh + c * 8, d8, c * 8);
  }
  rqs_ms_finish<INV>(st, cf, v, sw, sh, ladj0, out, ladj, bin, inside);
}

// ------------------------------------------------------------------------------------------------
// fp32 pair -> (hi, lo) fp16 pairs with hi + lo = v to ~2^-22 relative; v - float(hi) is exact in fp32
// (the remainder is one mixed-precision FMA per value, `fma.rn.f32.f16` -> SASS FHFMA: 4 instructions per pair)
__device__ __forceinline__ void split_f16_pair(float2 v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.y), "f"(v.x));
  float l0, l1;
  asm("{\n"
      ".reg .b16 h0, h1, m1;\n"
      "mov.b32 {h0, h1}, %2;\n"
      "mov.b16 m1, 0xBC00;\n"  // -1.0 in fp16
      "fma.rn.f32.f16 %0, h0, m1, %3;\n"
      "fma.rn.f32.f16 %1, h1, m1, %4;\n"
      "}\n"
      : "=f"(l0), "=f"(l1)
      : "r"(hi), "f"(v.x), "f"(v.y));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}

// ReLU + split in two conversions: hi = rz(max(u, 0)), lo = rn(max(u - hi, 0)).  Rounding hi TOWARDS ZERO makes the
// remainder of a positive u non-negative, so the ReLU of both halves rides on the conversions (`cvt.*.relu`) and the
// two FMNMX per pair disappear (a negative u gives hi = 0 and a negative remainder, which the second ReLU clears).
// |u - hi - lo| <= 2^-21 |u| (one bit more than the round-to-nearest split): still fp32-class.
// The remainder u - hi is one mixed-precision FMA per value (`fma.rn.f32.f16`: fp16 x fp16 + fp32 -> SASS FHFMA with a
// half selector on the packed hi register), exact in fp32: four instructions per PAIR of values in all.
__device__ __forceinline__ void split_relu_f16_pair(float2 u, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(u.y), "f"(u.x));
  float l0, l1;
  asm("{\n"
      ".reg .b16 h0, h1, m1;\n"
      "mov.b32 {h0, h1}, %2;\n"
      "mov.b16 m1, 0xBC00;\n"  // -1.0 in fp16
      "fma.rn.f32.f16 %0, h0, m1, %3;\n"
      "fma.rn.f32.f16 %1, h1, m1, %4;\n"
      "}\n"
      : "=f"(l0), "=f"(l1)
      : "r"(hi), "f"(u.x), "f"(u.y));
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(l1), "f"(l0));
}
__device__ __forceinline__ uint32_t pack_relu_bf16(float2 u) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(u.y), "f"(u.x));
  return r;
}

// Hidden-layer epilogue of one warp, full-width layer (N = 128): accumulator columns -> + bias -> ReLU -> operand planes.
// Straight-line code (no loop control, immediate TMEM offsets) and software-pipelined: the TMEM loads of column block
// b + 1 are issued BEFORE the arithmetic of block b, so the TMEM round trip hides behind ~130 instructions.  The static
// schedule of the rolled version was 310 issue cycles + ~165 cycles of exposed TMEM / shared-memory latency per 32
// columns (190 instructions, 62 of them loop control, address moves and uniform-register transfers).
template <int NSPLIT>
__device__ __forceinline__ void hidden_epilogue_128(uint32_t taddr, const float* sb) {
  float v[2][32];
  tc::tmem_ld16(taddr, v[0]);
  tc::tmem_ld16(taddr + 16, v[0] + 16);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    float4 bb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bb[i] = *reinterpret_cast<const float4*>(sb + 32 * b + 4 * i);
    tc::tmem_ld_wait();  // block b has landed
    if (b + 1 < 4) {     // next block: in flight during this block's arithmetic
      tc::tmem_ld16(taddr + 32 * (b + 1), v[(b + 1) & 1]);
      tc::tmem_ld16(taddr + 32 * (b + 1) + 16, v[(b + 1) & 1] + 16);
    }
    const float* vb = v[b & 1];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float2 w2[8];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b4 = bb[hf * 4 + i / 4];
        w2[i / 2] = f2_add(make_float2(vb[hf * 16 + i], vb[hf * 16 + i + 1]), make_float2(b4.x, b4.y));
        w2[i / 2 + 1] = f2_add(make_float2(vb[hf * 16 + i + 2], vb[hf * 16 + i + 3]), make_float2(b4.z, b4.w));
      }
      const uint32_t c0 = (uint32_t)(32 * b + 16 * hf);
      if (NSPLIT == 3) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_relu_f16_pair(w2[i], hi[i], lo[i]);
        tc::tmem_st8(taddr + 128 + c0 / 2, hi);
        tc::tmem_st8(taddr + 192 + c0 / 2, lo);
      } else {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_relu_bf16(w2[i]);
        tc::tmem_st8(taddr + 128 + c0 / 2, pk);
      }
    }
  }
}

// tcgen05.mma under a (uniform) enable predicate: the tail k-steps of a job are switched off without branches.
// Shared-memory descriptors travel as (low word, high word): only the 14-bit start-address field of the LOW word
// changes from MMA to MMA and it cannot carry (shared-memory addresses are < 2^18 bytes = 2^14 sixteen-byte units), so
// all per-MMA address arithmetic is 32-bit -- with 64-bit descriptors the compiler emitted ~300 carry-propagating
// instructions per job in front of the first MMA (~900 cycles of an idle tensor pipe per job, per-warp trace r2h).
__device__ __forceinline__ void mma_ts_p(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate, uint32_t enable) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 db;\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "setp.ne.b32 q, %6, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(enable)
      : "memory");
}
__device__ __forceinline__ void mma_ss_p(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate, uint32_t enable) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.ne.b32 q, %7, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(enable)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t desc_hi(uint64_t d) { return (uint32_t)(d >> 32); }

// Everything the MMA batch of one job needs that does not depend on the tile slot, prepared ONE JOB AHEAD by the
// issuer (while the previous job's MMAs execute): the per-job scalar work (parameter-bank loads, zero-block word,
// per-k-step operand offsets) is ~400 dependent integer instructions = ~1500 cycles of a single warp, which used to sit
// between "dependencies met" and "first MMA issued" of every job (per-warp traces profiles/trace_flow_tc_r2h.txt, r2k).
struct JobPkt {
  uint32_t kd[8];   // accumulator column offset of k-step ks (its leading all-zero weight rows, flow_tc_skip_kernel)
  uint32_t kb[8];   // weight descriptor offset of k-step ks inside its K-block (16-byte units): rows skipped + k offset
  uint32_t ki[8];   // instruction descriptor of k-step ks (N reduced by the skipped columns)
  uint32_t ka[8];   // layer 0: A descriptor offset of k-step ks (16-byte units); bit 31: SWIZZLE_32B tail k-step
  uint32_t en;      // bit ks: k-step ks is issued (inside K and not an all-zero weight block)
  int N, last, reg, nst;
  uint32_t col;     // first accumulator column
  int ji, pad;      // 40 words = 160 bytes: one entry of the shared-memory packet table [transform][job]
};
static_assert(sizeof(JobPkt) == kPktWords * 4, "JobPkt layout");

// One job of one tile slot, issued by ONE thread as a single straight-line batch (up to 8 k-steps = 2 weight K-blocks,
// 3 MMAs per k-step in the 3xFP16 mode), so the tensor pipe is fed back to back (bookkeeping between the K-blocks of a
// job used to cost ~500 cycles of an idle pipe per K-block).
//   hidden layers / last-layer chunks: A operand planes in TMEM (hi at a_tmem, lo 64 columns further)
template <int NSPLIT>
__device__ __forceinline__ void issue_job_ts(const JobPkt& q, uint32_t d_tmem, uint32_t a_tmem, uint32_t w_base0, uint32_t w_base1,
                                             uint64_t* empty0, uint64_t* empty1, uint64_t* full) {
  const uint32_t part16 = (uint32_t)q.N * 8u;  // [hi][lo] planes of a K-block are dense in the stage (16-byte units)
  const uint64_t d0 = tc::make_smem_desc(w_base0), d1 = tc::make_smem_desc(w_base1);
  const uint32_t bhi = desc_hi(d0), b0 = desc_lo(d0), b1 = desc_lo(d1);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t en = (q.en >> ks) & 1u;
    const uint32_t d = d_tmem + q.kd[ks];
    const uint32_t bh = (ks < 4 ? b0 : b1) + q.kb[ks], bl = bh + part16;
    const uint32_t a = a_tmem + (uint32_t)ks * 8u;
    mma_ts_p(d, a, bh, bhi, q.ki[ks], ks ? 1u : 0u, en);
    if (NSPLIT == 3) {
      mma_ts_p(d, a, bl, bhi, q.ki[ks], 1u, en);
      mma_ts_p(d, a + 64u, bh, bhi, q.ki[ks], 1u, en);
    }
    if (ks == 3 && empty0) tc::mma_commit(empty0);  // first stage consumed by both slots
  }
  if (empty1) tc::mma_commit(empty1);
  tc::mma_commit(full);
}
//   layer 0: (ctx | x) operand planes in shared memory (a_hi / a_lo: plane bases): 64-wide SWIZZLE_128B K-blocks, then
//   single k-steps in SWIZZLE_32B blocks; the offsets come prepared in q.ka
template <int NSPLIT>
__device__ __forceinline__ void issue_job_ss(const JobPkt& q, uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t w_base0,
                                             uint32_t w_base1, uint64_t* empty0, uint64_t* empty1, uint64_t* full) {
  const uint32_t part16 = (uint32_t)q.N * 8u;
  const uint64_t d0 = tc::make_smem_desc(w_base0), d1 = tc::make_smem_desc(w_base1);
  const uint32_t bhi = desc_hi(d0), b0 = desc_lo(d0), b1 = desc_lo(d1);
  const uint64_t fh = tc::make_smem_desc(a_hi), fl = tc::make_smem_desc(a_lo);
  const uint64_t th = tc::make_smem_desc_sw32(a_hi), tl = tc::make_smem_desc_sw32(a_lo);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t en = (q.en >> ks) & 1u;
    const uint32_t bh = (ks < 4 ? b0 : b1) + q.kb[ks], bl = bh + part16;
    const bool tail = (q.ka[ks] >> 31) != 0u;
    const uint32_t off = q.ka[ks] & 0x7fffffffu;
    const uint32_t ah = (tail ? desc_lo(th) : desc_lo(fh)) + off, al = (tail ? desc_lo(tl) : desc_lo(fl)) + off;
    const uint32_t ahi = tail ? desc_hi(th) : desc_hi(fh);
    mma_ss_p(d_tmem, ah, ahi, bh, bhi, q.ki[ks], ks ? 1u : 0u, en);
    if (NSPLIT == 3) {
      mma_ss_p(d_tmem, ah, ahi, bl, bhi, q.ki[ks], 1u, en);
      mma_ss_p(d_tmem, al, ahi, bh, bhi, q.ki[ks], 1u, en);
    }
    if (ks == 3 && empty0) tc::mma_commit(empty0);
  }
  if (empty1) tc::mma_commit(empty1);
  tc::mma_commit(full);
}

// INV: sampling (transforms in reverse, `passes` sweeps each) -- a compile-time flag so that the log_prob kernel carries
// none of the sweep / schedule logic.  NCHK = roundup(K, 8) / 8 (spline logits held in registers).
template <int NSPLIT, int NCHK, bool INV>
__global__ void __launch_bounds__(kTcThreads, 1)
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
  __shared__ __align__(8) uint64_t w_full[kMaxStages], w_empty[kMaxStages], a_ready[2], acc_full[2][2], acc_free[2][2], acc_hid[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t issue_lock;  // one MMA batch at a time enters the tensor core's queue
  __shared__ __align__(16) float sbias[8][2][128];  // [epilogue warp][job parity][column]: every warp stages its own copy
  // what an epilogue warp needs to know about job ji, one 16-byte shared-memory load instead of a chain of indexed
  // parameter-bank loads per job (x = N | last << 8 | region << 9 | nfeat << 16 | feat0 << 24, y = acc_col, z = b_off)
  __shared__ uint4 epi_s[kMaxJobs];

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
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 2); }  // released by both issuers
    issue_lock = 0u;
    for (int t = 0; t < 2; ++t) {
      mbar_init(&a_ready[t], 128);
      for (int r = 0; r < 2; ++r) { mbar_init(&acc_full[t][r], 1); mbar_init(&acc_free[t][r], 128); }
      // hidden-layer accumulators complete on their own barrier: all eight epilogue warps wait for EVERY completion of
      // it (a warp that skipped completions of a barrier would be released early by parity aliasing)
      mbar_init(&acc_hid[t], 1);
    }
    fence_mbar_init();
  }
  if (tid < L.njobs) {
    const TcJob& j = L.job[tid];
    epi_s[tid] = make_uint4((uint32_t)j.N | ((uint32_t)j.last << 8) | ((uint32_t)j.region << 9) | ((uint32_t)j.nfeat << 16) |
                                ((uint32_t)j.feat0 << 24),
                            (uint32_t)j.acc_col, (uint32_t)j.b_off, 0u);
  }
  if (warp == kIssuerWarp) tc::tmem_alloc(&tmem_base_s, 512);
  // job packets of every (transform, job), built once per CTA by all threads (the tensor pipe idles for ~1000 cycles per
  // job when the issuer derives them on the fly)
  JobPkt* pkt_s = reinterpret_cast<JobPkt*>(w_ring + (size_t)kStages * kStageBytes);
  auto build_pkt = [&](JobPkt& q, int tr, int ji_) {
    const int ksteps = L.job[ji_].ksteps;
    q.ji = ji_; q.pad = 0;
    q.N = L.job[ji_].N; q.last = L.job[ji_].last; q.reg = L.job[ji_].region; q.nst = L.job[ji_].nstages;
    q.col = (uint32_t)L.job[ji_].acc_col;
    const uint32_t idesc = tc::make_idesc_f16(NSPLIT == 3 ? 0 : 1, 128, q.N);
    unsigned long long skipw = 0ull;
    if (ji_ != 0 && args.skip) skipw = __ldg(args.skip + (size_t)tr * kMaxJobs + ji_);
    const uint32_t slo = (uint32_t)skipw, shi = (uint32_t)(skipw >> 32);
    uint32_t en = 0u;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t byte = ((ks < 4 ? slo : shi) >> (8 * (ks & 3))) & 0xFFu;  // leading all-zero weight rows / 16
      q.kd[ks] = byte << 4;
      q.kb[ks] = (byte << 7) + (uint32_t)((ks & 3) * 2);   // 128-byte weight rows, 32-byte k-steps (16-byte units)
      q.ki[ks] = idesc - (byte << 18);                     // N - 16 * byte columns (N field = N >> 3 at bit 17)
      en |= ((ks < ksteps && (int)(byte << 4) < q.N) ? 1u : 0u) << ks;
      const bool isfull = (ks >> 2) < L.in0_full;           // layer-0 operand: 64-wide K-blocks, then 32-byte tail k-steps
      q.ka[ks] = isfull ? (uint32_t)(ks >> 2) * (kKBlockBytes >> 4) + (uint32_t)((ks & 3) * 2)
                        : (0x80000000u | ((kFullBytes >> 4) + (uint32_t)(ks - 4 * L.in0_full) * (kTailBytes >> 4)));
    }
    q.en = en;
  };
  if (L.pkt_table)
    for (int i = tid; i < L.T * L.njobs; i += kTcThreads) build_pkt(pkt_s[i], i / L.njobs, i % L.njobs);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;

  // Register reallocation: the kernel starts with 168 registers per thread (12 warps); the control warpgroup (warps
  // 8-11: loader, two issuers, one idle warp) gives registers back and the two epilogue warpgroups take them, so the
  // spline / hidden-layer code is not squeezed into 168 registers (it spilled, and every change of the cold tile-start
  // code shifted the allocation of the hot loops: r3f was 12 % slower for that reason alone).
  if (warp >= 8) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
  if (warp == kLoaderWarp) {
    // ===================== weight loader =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;  // ring position / phase of the next stage to fill
      for (long long p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        for (int unit = 0; unit < n_units; ++unit) {
          const int tr = inverse ? (L.T - 1 - unit / sweeps) : unit;
          const unsigned char* tb = blob + (long long)tr * L.t_stride;
          for (int ji = 0; ji < L.njobs; ++ji) {
            if (!job_live(unit, ji)) continue;
            const TcJob& j = L.job[ji];
            // one bulk copy per ring stage: the K-blocks of a job are contiguous in the blob ([hi][lo] per K-block) and keep
            // that dense layout in the stage; a narrow job (N <= 64) brings both of its K-blocks into ONE stage
            const uint32_t kb_bytes = kParts * (uint32_t)j.N * 128u;
            const uint32_t st_bytes = (j.nstages == 1) ? kb_bytes * (uint32_t)j.nkb : kb_bytes;
            for (int q = 0; q < j.nstages; ++q) {
              mbar_wait_relaxed(&w_empty[s], ph ^ 1, 2000);
              mbar_expect_tx(&w_full[s], st_bytes);
              bulk_g2s(w_ring + s * kStageBytes, tb + j.w_off + (long long)q * kb_bytes, st_bytes, &w_full[s]);
              if (++s == kStages) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kIssuerWarp || warp == kIssuerWarp + 1) {
    // ===================== MMA issuers: one warp per tile slot =====================
    // Each issuer walks the (warp-uniform) job schedule for ITS tile slot: wait for the job's weight stages, for the
    // slot's dependencies (a hidden layer and the first last-layer chunk of a unit wait for the A operand -- a_ready:
    // every thread of the slot has finished the previous epilogue, so the accumulator is drained as well; a later
    // last-layer chunk waits until the previous chunk in ITS accumulator region has been consumed -- acc_free), then ONE
    // elected lane takes the submission lock and issues the MMAs of the whole job back to back.  The lock keeps the
    // batches of the two slots from interleaving in the tensor core's queue (interleaved, both accumulators complete at
    // the same time, both epilogues run at once and the pipe idles meanwhile); everything else an issuer does per job --
    // ~200 dependent scalar instructions, 600-1400 cycles of a single warp (per-warp traces r2m, r2s) -- now overlaps the
    // OTHER slot's batch instead of sitting between two batches.  A weight stage is released by the commits of both
    // issuers (w_empty counts 2).
    const int t = warp - kIssuerWarp;
    uint32_t stage = 0, sphase = 0;  // ring position / phase of the next weight K-block to consume
    int trace_n = t * 750, trace_end = trace_n + 750; (void)trace_n; (void)trace_end;
    uint32_t aph = 0;                          // a_ready phase of the slot
    uint32_t nchunk0 = 0, nchunk1 = 0;         // last-layer chunks issued so far per accumulator region
    struct JobRef { long long p; int unit; };
    // advance (p, unit, ji) to the next job that is executed; false after the last one of this CTA
    auto next_job = [&](long long& p_, int& unit_, int& ji_) -> bool {
      for (;;) {
        if (++ji_ == L.njobs) {
          ji_ = 0;
          if (++unit_ == n_units) { unit_ = 0; p_ += gridDim.x; }
        }
        if (p_ >= n_pairs) return false;
        if (job_live(unit_, ji_)) return true;
      }
    };
    auto prepare = [&](JobPkt& q, JobRef& r, long long p_, int unit_, int ji_) {
      r.p = p_; r.unit = unit_;
      const int tr = inverse ? (L.T - 1 - unit_ / sweeps) : unit_;
      if (L.pkt_table) {
        const uint4* src = reinterpret_cast<const uint4*>(pkt_s + tr * L.njobs + ji_);
        uint4* dst = reinterpret_cast<uint4*>(&q);
#pragma unroll
        for (int i = 0; i < kPktWords / 4; ++i) dst[i] = src[i];
      } else {
        build_pkt(q, tr, ji_);
      }
    };
    // pins a prepared packet in front of the next volatile statement (the compiler must not sink the preparation into
    // the critical path behind the barrier waits)
    auto pin = [&](JobPkt& q) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) asm volatile("" : "+r"(q.kd[ks]), "+r"(q.kb[ks]), "+r"(q.ki[ks]), "+r"(q.ka[ks]));
      asm volatile("" : "+r"(q.en), "+r"(q.N), "+r"(q.col));
    };

    JobPkt cur, nxt;
    JobRef rcur, rnxt;
    long long jp = blockIdx.x;
    int junit = 0, jji = -1;
    bool have = (jp < n_pairs) && next_job(jp, junit, jji);
    if (have) { prepare(cur, rcur, jp, junit, jji); pin(cur); }
    bool first_last = true;          // the first EXECUTED last-layer chunk of a unit waits for the last hidden activations
    int in_unit0 = 0, in_unit1 = 0;  // chunks of the current unit issued per region
    long long key_p = -1;
    int key_unit = -1;
    while (have) {
      if (rcur.p != key_p || rcur.unit != key_unit) {  // a new unit starts
        key_p = rcur.p; key_unit = rcur.unit;
        first_last = true; in_unit0 = 0; in_unit1 = 0;
      }
      const int ji = cur.ji;
      const int jreg = cur.reg;
      const bool two = cur.nst > 1;
      const bool active = (t == 0) || (2 * rcur.p + 1 < n_tiles);  // odd tile count: slot 1 idles in the last pair
      const bool wait_a = cur.last ? first_last : true;
      const bool wait_free = cur.last && (jreg ? in_unit1 : in_unit0) > 0;
      const uint32_t free_par = ((jreg ? nchunk1 : nchunk0) - 1) & 1;  // phase of the previous chunk of this region
      if (cur.last) {
        first_last = false;
        if (jreg) { ++in_unit1; ++nchunk1; } else { ++in_unit0; ++nchunk0; }
      }
      // ring stages of the job's (at most two) weight K-blocks: wait until they have landed, advance the ring position
      const uint32_t cs0 = stage, ph0 = sphase;
      if (++stage == kStages) { stage = 0; sphase ^= 1; }
      uint32_t cs1 = cs0;
      mbar_wait(&w_full[cs0], ph0);
      const uint32_t wb0 = smem_u32(w_ring + cs0 * kStageBytes);
      uint32_t wb1;
      if (two) {
        cs1 = stage;
        mbar_wait(&w_full[cs1], sphase);
        if (++stage == kStages) { stage = 0; sphase ^= 1; }
        wb1 = smem_u32(w_ring + cs1 * kStageBytes);
      } else {
        wb1 = wb0 + kParts * (uint32_t)cur.N * 128u;  // both K-blocks of a narrow job share the stage
      }
      MF_TRACE(lane == 0, 1, ji);
      if (active) {
        const uint32_t d_tmem = tmem + (uint32_t)(t * 256) + cur.col;
        const uint32_t a_tmem = tmem + (uint32_t)(t * 256) + 128;  // hi plane; lo plane 64 columns further
        const uint32_t a_base = smem_u32(a_buf + (size_t)t * kParts * kPlane);
        if (wait_a) {
          mbar_wait(&a_ready[t], aph);
          aph ^= 1;
        }
        if (wait_free) mbar_wait(&acc_free[t][jreg], free_par);
        MF_TRACE(lane == 0, 2, ji);
        tc::fence_after_thread_sync();
        if (tc::elect_one()) {
          while (atomicCAS(&issue_lock, 0u, 1u) != 0u) __nanosleep(64);
          // both issuers commit to the stage barriers: the first stage right after its k-steps when the job spans two
          uint64_t* e0 = two ? &w_empty[cs0] : nullptr;
          uint64_t* e1 = &w_empty[cs1];
          if (ji == 0)
            issue_job_ss<NSPLIT>(cur, d_tmem, a_base, a_base + kPlane, wb0, wb1, e0, e1, &acc_hid[t]);
          else
            issue_job_ts<NSPLIT>(cur, d_tmem, a_tmem, wb0, wb1, e0, e1, cur.last ? &acc_full[t][jreg] : &acc_hid[t]);
          atomicExch(&issue_lock, 0u);
        }
        __syncwarp();
        MF_TRACE(lane == 0, 3, ji);
      } else if (lane == 0) {
        // idle slot of the last pair: keep the weight ring moving (the stage was seen full above, so this arrival belongs
        // to the same use of the stage as the other issuer's commit)
        if (two) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_empty[cs0])) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_empty[cs1])) : "memory");
      }
      // the next job's packet: prepared while the MMAs just issued (or the other slot's) execute
      have = next_job(jp, junit, jji);
      if (have) {
        prepare(nxt, rnxt, jp, junit, jji);
        pin(nxt);
        cur = nxt;
        rcur = rnxt;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    // ===================== epilogue groups =====================
    const int t = warp >> 2;                  // tile slot
    const int q = warp & 3;                   // TMEM lane quarter this warp may access (4 consecutive warps cover all)
    const int row = q * 32 + lane;            // object inside the tile
    const int gt = (warp & 3) * 32 + lane;    // thread index inside the group (cooperative loops)
    constexpr int GT = 128;                   // threads per epilogue group
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
    const RqsScaled cf = make_rqs_scaled(L.K, args.univariate, args.bound, args.slope);
    // Biases: every warp keeps its own copy in shared memory (lane l stages columns 4l .. 4l+3 with one 16-byte load,
    // fetched one job ahead), so the four warps of a slot never wait for each other outside the mbarriers.
    // Soft-clip scale of the last-layer columns (same column layout in every chunk): the spline works on logits
    // pre-multiplied by 2 / ln(slope) (widths, heights) or 1 / ln(slope) (derivatives); a 4-aligned group of columns
    // never straddles a segment (segments are multiples of 8 columns wide).
    float* const sb_warp = &sbias[warp][0][0];
    uint32_t jobctr = 0;                      // bias double-buffer parity
    uint32_t fh0 = 0, fh1 = 0;   // hidden-layer accumulator completions seen so far, per slot (every warp waits for all)
    uint32_t fo0 = 0, fo1 = 0;   // last-layer chunk completions of the warp's OWN slot, per accumulator region
    int trace_n = 1500 + warp * 300, trace_end = trace_n + 300; (void)trace_n; (void)trace_end;

    struct EpiJob { int N, last, region, nfeat, feat0, acc_col, b_off; };
    auto epi_job = [&](int ji_) -> EpiJob {
      const uint4 r = epi_s[ji_];
      EpiJob e;
      e.N = (int)(r.x & 0xFFu); e.last = (int)((r.x >> 8) & 1u); e.region = (int)((r.x >> 9) & 1u);
      e.nfeat = (int)((r.x >> 16) & 0xFFu); e.feat0 = (int)(r.x >> 24); e.acc_col = (int)r.y; e.b_off = (int)r.z;
      return e;
    };
    const int njobs = L.njobs, nhid = L.L, bias_stride = L.bias_stride, PFt = L.PFt;

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

    const int nchunk = xcol0 >> 3;  // 16-byte operand chunks per context row (the last one zero-padded when C % 8 != 0)
    const float* ctx_end = gc + ((n - 1) / args.ctx_group + 1) * (long long)C;  // one past the last context float
    for (long long p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      const long long tile = 2 * p + t;
      if (tile >= n_tiles) break;  // odd tile count: slot 1 idles in the last pair (the MMA warp skips it too)
      const bool has_tile = true;
      const long long tile0 = tile * 128;
      const int rows = has_tile ? (int)min((long long)128, n - tile0) : 0;
      const bool valid = row < rows;
      const long long g = tile0 + row;
      // working copy of x lives in the output buffer (each thread touches only its own row); when sampling,
      // gx holds the noise and the base sample a + b * noise is the starting point
      float* gy = static_cast<float*>(args.work);  // sampling: target y of the transform being inverted
      float ladj = 0.f;
      if (has_tile) {
        MF_TRACE(lane == 0, 8, 0);
        // x: the first eight features are requested BEFORE the context staging and consumed after it (their latency
        // hides behind the context loop); further groups of eight follow the old load-then-use pattern
        float xv0[8];
  #pragma unroll
        for (int u = 0; u < 8; ++u) xv0[u] = (valid && u < D) ? gx[g * D + u] : 0.f;
        auto x_in = [&](float v0, int f) {
          if (inverse && valid)
            v0 = (args.base == MF_BASE_DIAG_NORMAL) ? (float)bp.a[f] + (float)bp.b[f] * v0
                                                   : (float)bp.a[f] + ((float)bp.b[f] - (float)bp.a[f]) * v0;
          if (!inverse && valid && bp.xmap_kind != MF_XMAP_NONE) v0 = xmap_cold(bp, v0, f);  // logit bridge
          if (valid) gz[g * D + f] = v0;
          put1(row, xcol0 + f, v0);  // x columns of the layer-0 operand
        };
        MF_TRACE(lane == 0, 10, 0);
        // ---- stage (ctx | 0) of the layer-0 operand once per tile; every MMA that read the previous tile's
        //      operand has retired (this group waited on the last accumulator of that tile)
        {
          const int kfill = L.in0_ksteps * 16;
          for (int k = xcol0 + D; k < kfill; ++k) put1(row, k, 0.f);
          // context: every thread stages ITS OWN row (thread = object = operand row), 16-byte operand chunks of 8 columns.
          // A context row is only 4-byte aligned in global memory (row stride C floats): it is fetched as the 32-byte
          // ALIGNED sectors that hold it (ld.global.nc.v8, SASS LDG.256, four sectors in flight, every sector requested
          // once) and each chunk is shifted into place out of two neighbouring sectors with three select stages whose
          // predicates are per-thread constants of the tile.  Eight consecutive rows hit the eight swizzle positions of
          // a 128-byte line, so the 16-byte stores of a warp are conflict-free.
          // (History, per-warp traces r2t-r3b: one chunk per thread with consecutive threads on consecutive chunks cost
          // 7-9 k cycles per tile pair -- first 8 scalar loads per chunk, each an L2 round trip because this kernel
          // leaves ~4 KB of L1, then ~150 instructions of per-chunk bookkeeping; the loop without any load still took
          // 5.6 k cycles.)
          const int G = args.ctx_group;   // x rows per context row (context hoisted per event in the MEM sweep)
          long long crow = tile0 + row;
          if (G != 1) crow = crow / G;
          const float* rowp = gc + crow * C;
          const int mis = (int)((reinterpret_cast<uintptr_t>(rowp) >> 2) & 7);  // floats past the sector boundary
          const float* wp = rowp - mis;
          const bool s4 = (mis & 4) != 0, s2 = (mis & 2) != 0, s1 = (mis & 1) != 0;
          MF_TRACE(lane == 0, 11, 0);
          // groups of three chunks = four sectors
          auto load_group = [&](int c0, float (&sct)[4][8]) {
  #pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float* sp = wp + 8 * (c0 + j);
              if (valid && c0 + j <= nchunk && sp + 8 <= ctx_end) {
                ldg256(sp, sct[j]);
              } else {  // rows past the end of the batch (zeros) / the last sector of the tensor (element by element)
                const F8 g8 = ldg_sector_guarded(sp, ctx_end, valid && c0 + j <= nchunk);
  #pragma unroll
                for (int e = 0; e < 8; ++e) sct[j][e] = g8.v[e];
              }
            }
          };
          auto convert_group = [&](int c0, const float (&sct)[4][8]) {
  #pragma unroll
            for (int u = 0; u < 3; ++u) {
              const int c8 = c0 + u;
              if (c8 < nchunk) {
                float v[8];
                {
                  // v[e] = window[e + mis], window = sectors u, u + 1
                  float t4[12], t2[10];
  #pragma unroll
                  for (int e = 0; e < 12; ++e) t4[e] = s4 ? (e + 4 < 8 ? sct[u][e + 4] : sct[u + 1][e - 4]) : (e < 8 ? sct[u][e] : sct[u + 1][e - 8]);
  #pragma unroll
                  for (int e = 0; e < 10; ++e) t2[e] = s2 ? t4[e + 2] : t4[e];
  #pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = s1 ? t2[e + 1] : t2[e];
                  if (c8 == nchunk - 1) {  // columns past C belong to the next row
  #pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = (c8 * 8 + e < C) ? v[e] : 0.f;
                  }
                }
                const int r = row;
                const uint32_t off = (c8 * 8 < kfull_cols)
                                         ? (uint32_t)(c8 >> 3) * kKBlockBytes + (uint32_t)r * 128u + (uint32_t)(((c8 & 7) ^ (r & 7)) << 4)
                                         : kFullBytes + (uint32_t)((c8 * 8 - kfull_cols) >> 4) * kTailBytes + (uint32_t)r * 32u +
                                               (uint32_t)(((c8 & 1) ^ ((r >> 2) & 1)) << 4);
                uint4 hi, lo;
                if (NSPLIT == 3) {
                  split_f16_pair(make_float2(v[0], v[1]), hi.x, lo.x); split_f16_pair(make_float2(v[2], v[3]), hi.y, lo.y);
                  split_f16_pair(make_float2(v[4], v[5]), hi.z, lo.z); split_f16_pair(make_float2(v[6], v[7]), hi.w, lo.w);
                  *reinterpret_cast<uint4*>(a_hi + off) = hi;
                  *reinterpret_cast<uint4*>(a_lo + off) = lo;
                } else {
                  hi.x = tc::pack_bf16(v[0], v[1]); hi.y = tc::pack_bf16(v[2], v[3]);
                  hi.z = tc::pack_bf16(v[4], v[5]); hi.w = tc::pack_bf16(v[6], v[7]);
                  *reinterpret_cast<uint4*>(a_hi + off) = hi;
                }
              }
            }
          };
          // (Measured and not kept: staging the NEXT tile's context three chunks at a time in front of the last unit's
          // hidden-layer waits -- the tile start shrank from ~10 k to 3.8 k cycles but the last unit grew by as much,
          // profiles/trace_flow_tc_r3j_prestaged_context.txt; evaluating the two splines of a chunk pair side by side for
          // instruction-level parallelism -- 7.19 instead of 6.47 ms on TF-A, profiles/trace_flow_tc_r3h_paired_splines.txt.)
          // (requesting group k + 1 before converting group k hides ~2.5 k cycles of load latency per tile but made the
          // whole kernel 12 % slower, r3f: 64 more live registers at this point change the allocation of the hot loops)
          for (int c0 = 0; c0 < nchunk; c0 += 3) {
            float sct[4][8];
            load_group(c0, sct);
            convert_group(c0, sct);
          }
        }
        // x columns (requested before the context loop)
  #pragma unroll
        for (int u = 0; u < 8; ++u)
          if (u < D) x_in(xv0[u], u);
        for (int f0 = 8; f0 < D; f0 += 8) {  // eight loads in flight
          float xv[8];
  #pragma unroll
          for (int u = 0; u < 8; ++u) xv[u] = (valid && f0 + u < D) ? gx[g * D + f0 + u] : 0.f;
  #pragma unroll
          for (int u = 0; u < 8; ++u)
            if (f0 + u < D) x_in(xv[u], f0 + u);
        }
        MF_TRACE(lane == 0, 12, 0);
        {  // pull this slot's next tile (x rows and context rows) into L2 while this one is being processed
          const long long nt0 = (tile + 2 * (long long)gridDim.x) * 128;
          if (nt0 < n) {
            const long long nrows = min((long long)128, n - nt0);
            const char* px = reinterpret_cast<const char*>(gx + nt0 * D);
            const char* pc = reinterpret_cast<const char*>(gc + (nt0 / args.ctx_group) * C);
            const long long xb = nrows * D * 4, cb = ((nt0 + nrows - 1) / args.ctx_group - nt0 / args.ctx_group + 1) * C * 4;
            // one prefetch per 32-byte SECTOR (a prefetch of one address per 128-byte line left the tile start waiting
            // for DRAM: 7 k cycles for the three context rounds, per-warp traces r2x-r2z)
            for (long long o = (long long)gt * 32; o < xb; o += (long long)GT * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + o));
            for (long long o = (long long)gt * 32; o < cb; o += (long long)GT * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
          }
        }
      }
      float lp_base = 0.f;               // base log-density, accumulated as the last transform's outputs appear
      float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const int tr0 = inverse ? (L.T - 1) : 0;
        if (4 * lane < L.job[0].N)
          bias_next = __ldg(reinterpret_cast<const float4*>(bias_all + (size_t)tr0 * L.bias_stride + L.job[0].b_off) + lane);
      }
      // biases of the job after (unit_, ji_) -> bias_next (a register per lane; stored to the warp's slot when that job
      // starts)
      auto fetch_next_bias = [&](int unit_, int ji_) {
        int ju = unit_, jn = ji_ + 1;
        if (jn == njobs) { jn = 0; ++ju; }
        if (inverse) {
          while (ju < n_units && !job_live(ju, jn)) {  // next job that is actually executed
            if (++jn == njobs) { jn = 0; ++ju; }
          }
        }
        if (ju < n_units) {
          const int trn = inverse ? (L.T - 1 - ju / sweeps) : ju;
          const EpiJob nj = epi_job(jn);
          if (4 * lane < nj.N)
            bias_next = __ldg(reinterpret_cast<const float4*>(bias_all + (size_t)trn * bias_stride + nj.b_off) + lane);
        }
      };
      MF_TRACE(lane == 0, 9, 0);
      for (int unit = 0; unit < n_units; ++unit) {
        const int tr = inverse ? (L.T - 1 - unit / sweeps) : unit;
        if (inverse && (unit % sweeps) == 0 && has_tile) {  // y <- x, x <- 0  (AutoregressiveTransform._inverse)
          for (int f = 0; f < D; ++f) {
            if (valid) { gy[g * D + f] = gz[g * D + f]; gz[g * D + f] = 0.f; }
            put1(row, xcol0 + f, 0.f);
          }
        }
        // ---- layer-0 operand: the context part was staged at the tile start, the D columns of x were written
        //      by this thread when it produced them (tile start / spline output of the previous unit).
        fence_proxy_async_smem();
        mbar_arrive_warp(&a_ready[t], lane);
        // the first spline input of the unit (every input of a unit is known when the unit starts): in flight during the
        // hidden layers
        const float* xs = (inverse ? gy : gz) + g * D;
        float xfirst = valid ? xs[0] : 0.f;  // (chunks hold consecutive features, the first one feature 0)
        // ---- hidden layers: jobs [0, nhid)
        for (int ji = 0; ji < nhid; ++ji) {
          const EpiJob j = epi_job(ji);
          const uint32_t jpar = jobctr & 1;
          ++jobctr;
          // this job's biases -> the warp's shared-memory slot (double-buffered by job parity), overlapping the MMA
          float* sb = sb_warp + jpar * 128;
          if (4 * lane < j.N) reinterpret_cast<float4*>(sb)[lane] = bias_next;
          fetch_next_bias(unit, ji);  // the next executed job's biases, one job ahead
          __syncwarp();
          // TMEM -> +bias -> ReLU -> fp16 hi/lo operand planes of the next layer (in place).
          // (Sharing every hidden epilogue between the warps of BOTH slots -- same lane quarter, half the columns each --
          // was slower overall, 23.5 vs 20.8 ms, profiles/trace_flow_tc_r2p.txt: the epilogue is bound by the issue rate
          // of its scheduler's ALU / FMA pipes, not by latency.)
          MF_TRACE(lane == 0, 5, ji);
          mbar_wait(&acc_hid[t], (t ? fh1 : fh0) & 1);
          if (t) ++fh1; else ++fh0;
          MF_TRACE(lane == 0, 6, ji);
          tc::fence_after_thread_sync();
          const uint32_t taddr = tmem + lane_off + (uint32_t)(t * 256);
          if (j.N == 128) {
            hidden_epilogue_128<NSPLIT>(taddr, sb);
          } else {
            const int ce = j.N;
            for (int c00 = 0; c00 < ce; c00 += 32) {
              float vv[32];
              tc::tmem_ld16(taddr + c00, vv);
              if (c00 + 16 < ce) tc::tmem_ld16(taddr + c00 + 16, vv + 16);
              float4 bb[8];  // the biases travel from shared memory while the TMEM loads are in flight
#pragma unroll
              for (int i = 0; i < 8; ++i) bb[i] = *reinterpret_cast<const float4*>(sb + c00 + 4 * i);
              tc::tmem_ld_wait();
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const int c0 = c00 + hf * 16;
                if (c0 < ce) {
                  float* v = vv + hf * 16;
                  float2 w2[8];
#pragma unroll
                  for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = bb[hf * 4 + i / 4];
                    w2[i / 2] = f2_add(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y));
                    w2[i / 2 + 1] = f2_add(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w));
                  }
                  // 16 values -> 8 packed 16-bit pairs per plane -> operand planes in TMEM
                  if (NSPLIT == 3) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) split_relu_f16_pair(w2[i], hi[i], lo[i]);
                    tc::tmem_st8(taddr + 128 + c0 / 2, hi);
                    tc::tmem_st8(taddr + 192 + c0 / 2, lo);
                  } else {
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_relu_bf16(w2[i]);
                    tc::tmem_st8(taddr + 128 + c0 / 2, pk);
                  }
                }
              }
            }
          }
          MF_TRACE(lane == 0, 4, ji);
          tc::tmem_st_wait();
          tc::fence_before_thread_sync();
          mbar_arrive_warp(&a_ready[t], lane);
          MF_TRACE(lane == 0, 7, ji);
        }
        // ---- last layer: jobs [nhid, njobs), one chunk of features each.  A tight loop: the chunk's biases come straight
        //      from the blob (pre-scaled at pack time), its record is one shared-memory load, and the first input of the
        //      NEXT chunk is fetched before this chunk's splines (all inputs of a unit are known when the unit starts).
        const unsigned long long fmask = sched ? sched[(size_t)tr * MF_MAX_FEATURES + (unit % sweeps)] : ~0ull;
        const uint32_t taddr = tmem + lane_off + (uint32_t)(t * 256);
        for (int ji = nhid; ji < njobs; ++ji) {
          const EpiJob j = epi_job(ji);
          const float x0 = xfirst;
          if (ji + 1 < njobs && valid) xfirst = xs[j.feat0 + j.nfeat];
          if (!inverse || job_live(unit, ji)) {
            const uint32_t jpar = jobctr & 1;
            ++jobctr;
            float* sb = sb_warp + jpar * 128;
            if (4 * lane < j.N) reinterpret_cast<float4*>(sb)[lane] = bias_next;
            fetch_next_bias(unit, ji);
            __syncwarp();
            MF_TRACE(lane == 0, 5, ji);
            mbar_wait(&acc_full[t][j.region], (j.region ? fo1 : fo0) & 1);
            if (j.region) ++fo1; else ++fo0;
            MF_TRACE(lane == 0, 6, ji);
            tc::fence_after_thread_sync();
            float xnext = x0;
#pragma unroll 1
            for (int fl = 0; fl < j.nfeat; ++fl) {  // one copy of the spline code (instruction cache)
              const int f = j.feat0 + fl;
              const float xv = xnext;
              if (fl + 1 < j.nfeat && valid) xnext = xs[f + 1];  // next feature's input, in flight during this spline
              if (!((fmask >> f) & 1ull)) continue;  // sampling: provisional feature of this sweep, keeps its value
              float out, la;
              int bin;
              bool ins;
              spline_tmem<NCHK, INV>(taddr + j.acc_col + fl * PFt, sb + fl * PFt, cf, xv, out, la, bin, ins);
              put1(row, xcol0 + f, valid ? out : 0.f);  // next unit's layer-0 operand (its last reader, job 0, has retired)
              if (valid) {
                gz[g * D + f] = out;
                ladj += la;
                if (!inverse && unit == n_units - 1)
                  lp_base += base_log_prob_1<float>(args.base, out, (float)bp.a[f], (float)bp.b[f]);
                if (!inverse && args.bins) args.bins[((size_t)tr * n + g) * D + f] = (int8_t)bin;
                if (!inverse && args.inside) args.inside[((size_t)tr * n + g) * D + f] = ins ? 1 : 0;
              }
            }
            tc::fence_before_thread_sync();
            mbar_arrive_warp(&acc_free[t][j.region], lane);
            MF_TRACE(lane == 0, 7, ji);
          }
        }
      }
      if (valid && !inverse) {
        if (args.ladj) static_cast<float*>(args.ladj)[g] = ladj;
        if (args.logprob) static_cast<float*>(args.logprob)[g] = lp_base + ladj;
      }
      if (valid && inverse && bp.xmap_kind != MF_XMAP_NONE)  // output map of sample (sigmoid bridge) on the thread's own row
        for (int f = 0; f < D; ++f) gz[g * D + f] = xmap_cold(bp, gz[g * D + f], f);
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == kIssuerWarp) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
static long long tc_sched_offset(const TcLayout& L) {  // sampling schedule after the biases, 8-byte aligned
  return (L.bias_off + (long long)L.bias_stride * L.T * 4 + 7) / 8 * 8;
}

// zero-block table [T][kMaxJobs] uint64 after the schedule
static long long tc_skip_offset(const TcLayout& L) { return tc_sched_offset(L) + kSchedBytes; }
static long long tc_blob_total(const TcLayout& L) { return tc_skip_offset(L) + (long long)L.T * kMaxJobs * 8; }

int flow_tc_blob_bytes(const mf_flow_desc& d, int gemm, long long* bytes) {
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  *bytes = tc_blob_total(L);
  return MF_OK;
}

int flow_tc_sched_layout(const mf_flow_desc& d, int gemm, long long* offset, int* features_per_chunk) {
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  *offset = tc_sched_offset(L);
  *features_per_chunk = L.fpc;
  return MF_OK;
}

int flow_tc_pack(const mf_flow_desc& d, const void* const* w, const void* const* b, const uint8_t* const* m,
                 int gemm, void* blob, long long blob_bytes, cudaStream_t st) {
  // parameters are fp32 on this path
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  const long long need = tc_blob_total(L);
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
    flow_tc_skip_kernel<float><<<L.njobs, 8, 0, st>>>(
        L, a, reinterpret_cast<unsigned long long*>((unsigned char*)blob + tc_skip_offset(L)) + (size_t)t * kMaxJobs);
    rc = post_launch("flow_tc_skip_kernel");
    if (rc != MF_OK) return rc;
  }
  MF_CUDA_CHECK(cudaMemsetAsync((char*)blob + tc_sched_offset(L), 0xFF, kSchedBytes, st));
  return MF_OK;
}

long long flow_tc_workspace_bytes(const mf_flow_desc& d, long long n, int gemm, int op) {
  (void)gemm;
  if (op == MF_FLOW_OP_SAMPLE) return n * d.features * 4;  // y targets of the transform being inverted
  return 0;
}

template <int NS, int NCHK, bool INV>
static int tc_launch(const TcLayout& L, const FlowKernelArgs& a, const BaseParams& bp, unsigned grid, size_t smem, cudaStream_t st) {
  MF_CUDA_CHECK(cudaFuncSetAttribute(flow_tc_kernel<NS, NCHK, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  flow_tc_kernel<NS, NCHK, INV><<<grid, kTcThreads, smem, st>>>(L, a, bp);
  return MF_OK;
}
template <int NS, bool INV>
static int tc_launch_k(const TcLayout& L, const FlowKernelArgs& a, const BaseParams& bp, unsigned grid, size_t smem, cudaStream_t st) {
  switch (L.Kp) {
    case 8: return tc_launch<NS, 1, INV>(L, a, bp, grid, smem, st);
    case 16: return tc_launch<NS, 2, INV>(L, a, bp, grid, smem, st);
    case 24: return tc_launch<NS, 3, INV>(L, a, bp, grid, smem, st);
    case 32: return tc_launch<NS, 4, INV>(L, a, bp, grid, smem, st);
    case 40: return tc_launch<NS, 5, INV>(L, a, bp, grid, smem, st);
  }
  set_error("tensor-core flow path: unsupported bins (roundup(K,8) = %d)", L.Kp);
  return MF_EUNSUPPORTED;
}

int flow_tc_run(const mf_flow_desc& d, const FlowKernelArgs& a, const BaseParams& bp, int gemm, cudaStream_t st) {
  TcLayout L;
  int rc = make_tc_layout(d, gemm, &L);
  if (rc != MF_OK) return rc;
  MF_REQUIRE(a.z != nullptr, "tensor-core flow path needs the z / x output buffer (it is the working copy of x)");
  FlowKernelArgs a_run = a;
  if (a.mode == 1) a_run.sched = reinterpret_cast<const unsigned long long*>(static_cast<const char*>(a.blob) + tc_sched_offset(L));
  a_run.skip = reinterpret_cast<const unsigned long long*>(static_cast<const char*>(a.blob) + tc_skip_offset(L));
  if (!(a.slope > 0.0)) {
    set_error("the tcgen05 flow path needs the soft clip of the spline logits (slope > 0); use MF_GEMM_EXACT");
    return MF_EUNSUPPORTED;
  }
  if (a.mode == 1)
    MF_REQUIRE(a.work != nullptr && a.work_bytes >= flow_tc_workspace_bytes(d, a.n, gemm, MF_FLOW_OP_SAMPLE),
               "tensor-core sampling needs a workspace of mf_flow_workspace_bytes() bytes");
  const int parts = (L.nsplit == 3) ? 2 : 1;
  MF_REQUIRE(L.stages >= 2, "layer-0 operand too large for the weight ring");
  const size_t smem = (size_t)2 * parts * L.in0_plane_bytes + (size_t)L.stages * parts * kWPartBytes +
                      (L.pkt_table ? (size_t)L.T * L.njobs * kPktWords * 4 : 0);
  const long long n_pairs = ((a.n + 127) / 128 + 1) / 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<long long>(n_pairs, sms);
  if (L.nsplit == 3) rc = (a.mode == 1) ? tc_launch_k<3, true>(L, a_run, bp, grid, smem, st) : tc_launch_k<3, false>(L, a_run, bp, grid, smem, st);
  else               rc = (a.mode == 1) ? tc_launch_k<1, true>(L, a_run, bp, grid, smem, st) : tc_launch_k<1, false>(L, a_run, bp, grid, smem, st);
  if (rc != MF_OK) return rc;
  return post_launch("flow_tc_kernel");
}

}  // namespace mf
