// flow_mma.cu -- fused MAF log_prob / sample for hyper-networks WIDER than the tcgen05 kernel takes
// (hidden width 129..512: the unfolding flow [512]x2, the 1-D transfer flows [256]x3), fp32 tensors.
//
// Why a second tensor-core kernel: flow_tc.cu keeps a layer's activations as two fp16 planes in TMEM; at width 512
// those planes alone are 128 x 512 x 4 B = 256 KB = all of TMEM (and more than shared memory), so that scheme stops
// at width 128.  Here a tile of 64 objects keeps ONE fp32 activation buffer in shared memory, a whole
// layer's output is accumulated in REGISTERS (64 x 512 fp32 = 128 accumulators per thread) and written back in place
// once every warp is done reading the layer's input.  The products run on the warp-level tensor path
// (mma.sync m16n8k16, fp16 operands, fp32 accumulate) with the same 3-term hi/lo split as the tcgen05 kernel:
// a*b ~ hi(a)hi(b) + hi(a)lo(b) + lo(a)hi(b), 22 significant bits.  Measured legacy-path rate on B200:
// 8 cycles per m16n8k8 TF32 MMA per scheduler (tools/probe_mma_sync.py), i.e. ~1/8 of the tcgen05 peak -- still
// several times the CUDA-core FMA GEMM of flow.cu for these shapes.
//
// Weights: the pre-masked, transposed blob of the exact path (Wt[kpad][ld], flow.cu) converted in place at pack time
// to fp16 hi/lo K-PAIRS -- the 8-byte slot (k/2, n) holds {half2 hi(k, k+1), half2 lo(k, k+1)}, i.e. one B-fragment
// register pair of the MMA -- and streamed through a cp.async ring; four lanes evaluate one spline (rqs_quad).
// Semantics: zuko NormalizingFlow.log_prob / sample as in flow.cu (same reference lines).
#include <algorithm>

#include "flow_common.cuh"
#include "tc_common.cuh"

namespace mf {
namespace {

constexpr int kWM = 64;        // objects per tile
constexpr int kWThreads = 256; // 8 warps
constexpr int kWK = 16;        // k-step of the MMA = staged weight rows

__device__ __forceinline__ void cp16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kStageFloats = 5 * kWK * (128 + 4);  // weight ring: 5 stages of [16 k-rows][128 columns (+4 pad)]

// Activations live in shared memory PRE-SPLIT: the float2 slot of columns (2c, 2c+1) holds {half2 hi, half2 lo} with
// hi + lo = value to 22 bits, so an A fragment of the MMA is one 64-bit load and no conversion (every value is split
// once by the thread that produced it instead of once per consuming warp).
__device__ __forceinline__ void store_split(float* slot, float v0, float v1) {
  uint32_t hi, lo;
  tc::split_f16(v0, v1, hi, lo);
  *reinterpret_cast<uint2*>(slot) = make_uint2(hi, lo);
}

constexpr int kNS = 5;  // weight ring depth (stages of [16 k-rows][128 columns])

// Out[64][0..ncols) = act( A[64][K] * Wt[K][0..ncols) + bias ); A is pre-split, Out is written pre-split when
// SPLIT_OUT (next layer's operand, may alias A) or as fp32 (spline parameters).
// NB N-blocks of 128 columns are accumulated in registers; warp w owns columns 16w.. of every block.
// Weights: cp.async ring, kNS - 1 stages in flight, one block-wide barrier per k-step.
template <int NB, bool SPLIT_OUT>
__device__ __forceinline__ void mma_layer(const float* A, int lda, int K, const float* __restrict__ Wt, int ldw,
                                          int ncols, const float* __restrict__ bias, float* Out, int ldo, bool relu,
                                          float* wstage, const unsigned short* __restrict__ klim) {
  constexpr int NT = 2;
  constexpr int NBW = 64 * NT;                        // columns per N-block
  constexpr int LDS = NBW + 4;                        // staged row stride: 2t*LDS + g hits 32 distinct banks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n_w = warp * 8 * NT;
  float acc[NB][4][NT][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) { acc[nb][i][j][0] = acc[nb][i][j][1] = acc[nb][i][j][2] = acc[nb][i][j][3] = 0.f; }

  const int nk = K / kWK;
  const int nblocks = min(NB, (ncols + NBW - 1) / NBW);
  // autoregressive masks leave the weight rows beyond some k exactly zero for a whole 128-column block (hidden units
  // sorted by degree make the masked matrices block-triangular): only the k-steps up to that row are streamed and issued
  int nkb[NB];
  int total = 0;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    nkb[nb] = (nb < nblocks) ? (klim ? max(1, min(nk, (int)klim[nb])) : nk) : 0;
    total += nkb[nb];
  }
  auto steps_of = [&](int nb) {
    int v = nkb[0];
#pragma unroll
    for (int i = 1; i < NB; ++i) v = (nb == i) ? nkb[i] : v;
    return v;
  };
  // staging: stage q of the step sequence = [8 k-pair rows][128 columns] x 8 bytes of (block nb, k-chunk kc); ncols
  // is a multiple of 128, so every thread copies two 16-byte pieces per stage: pair rows tid/64 and tid/64 + 4,
  // columns 2 (tid % 64).  A pair row of the blob is 2 ldw floats long.
  int st_nb = 0, st_kc = 0, st_slot = 0, st_step = 0;
  const int st_row = threadIdx.x >> 6, st_col = (threadIdx.x & 63) << 2;  // float offsets inside a pair row
  auto stage = [&]() {
    if (st_step < total) {
      float* dst = wstage + st_slot * (kWK * LDS) + st_row * (2 * LDS) + st_col;
      const float* src = Wt + (size_t)(st_kc * (kWK / 2) + st_row) * (2 * ldw) + st_nb * (2 * NBW) + st_col;
      cp16(dst, src);
      cp16(dst + 4 * (2 * LDS), src + (size_t)4 * (2 * ldw));
      if (++st_kc == steps_of(st_nb)) { st_kc = 0; ++st_nb; }
    }
    ++st_step;
    st_slot = (st_slot + 1 == kNS) ? 0 : st_slot + 1;
    cp_commit();  // one group per step, empty past the end: wait_group<kNS-2> below always means "this step's stage landed"
  };
#pragma unroll
  for (int p = 0; p < kNS - 1; ++p) stage();
  int rd_slot = 0;
  int s = 0;
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    if (nb >= nblocks) break;
    const bool mine = nb * NBW + n_w < ncols;  // warp-uniform
    for (int kc = 0; kc < nkb[nb]; ++kc, ++s) {
      cp_wait<kNS - 2>();
      __syncthreads();     // stage s visible to all; everyone is done with step s-1, whose buffer is refilled next
      stage();
      const float* W = wstage + rd_slot * (kWK * LDS);
      rd_slot = (rd_slot + 1 == kNS) ? 0 : rd_slot + 1;
      if (mine) {
        const int k0 = kc * kWK;
        uint32_t ah[4][4], al[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* r0 = A + (size_t)(16 * i + g) * lda + k0 + 2 * t;
          const float* r1 = r0 + 8 * (size_t)lda;
          const uint2 v0 = *reinterpret_cast<const uint2*>(r0), v1 = *reinterpret_cast<const uint2*>(r1);
          const uint2 v2 = *reinterpret_cast<const uint2*>(r0 + 8), v3 = *reinterpret_cast<const uint2*>(r1 + 8);
          ah[i][0] = v0.x; al[i][0] = v0.y; ah[i][1] = v1.x; al[i][1] = v1.y;
          ah[i][2] = v2.x; al[i][2] = v2.y; ah[i][3] = v3.x; al[i][3] = v3.y;
        }
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j) {  // staged as [8 pair rows][132] 8-byte slots: t*132 + g covers 16 distinct 8-byte banks
          const uint2* wc = reinterpret_cast<const uint2*>(W) + n_w + 8 * j + g;
          const uint2 p0 = wc[t * LDS], p1 = wc[(t + 4) * LDS];
          bh[j][0] = p0.x; bl[j][0] = p0.y; bh[j][1] = p1.x; bl[j][1] = p1.y;
        }
        // the three terms of a product accumulate into the same registers: issue term by term so that dependent MMAs
        // are 4*NT instructions apart instead of back to back
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) mma16816(acc[nb][i][j], ah[i], bh[j][0], bh[j][1]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) mma16816(acc[nb][i][j], ah[i], bl[j][0], bl[j][1]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) mma16816(acc[nb][i][j], al[i], bh[j][0], bh[j][1]);
      }
    }
  }
  cp_wait<0>();
  __syncthreads();  // every warp has finished reading A and the ring: write the layer's output (possibly in place)
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int col = nb * NBW + n_w + 8 * j + 2 * t;
      if (col < ncols) {
        const float b0 = bias[col], b1 = bias[col + 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v0 = acc[nb][i][j][0] + b0, v1 = acc[nb][i][j][1] + b1;
          float v2 = acc[nb][i][j][2] + b0, v3 = acc[nb][i][j][3] + b1;
          if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
          float* o0 = Out + (size_t)(16 * i + g) * ldo + col;
          float* o1 = Out + (size_t)(16 * i + g + 8) * ldo + col;
          if (SPLIT_OUT) {
            store_split(o0, v0, v1);
            store_split(o1, v2, v3);
          } else {
            o0[0] = v0; o0[1] = v1;  // scalar stores: the parameter rows have an odd stride (conflict-free spline reads)
            o1[0] = v2; o1[1] = v3;
          }
        }
      }
    }
  }
  __syncthreads();
}


// One univariate spline evaluated by FOUR consecutive lanes (a quad): lane q owns bins [q*Kq, (q+1)*Kq).  The
// logits p = [K widths | K heights | K-1 derivatives] live in shared memory and are overwritten in place with the
// upper knots of every bin, so that any lane can read the two knots of the selected bin.  Same operations as
// rqs_univariate (flow_common.cuh: soft clip, max-subtracted softmax, cumulative knots, k = #(knots < v) - 1,
// rational-quadratic map); only the order of the softmax / prefix sums differs (quad tree instead of sequential).
// All 32 lanes of the warp must call it; every lane of the quad returns the result.
__device__ __forceinline__ void rqs_quad(float* p, const SplineCfg<float>& cfg, bool inverse, float v, float& out,
                                         float& ladj, int& bin, bool& inside) {
  const int K = cfg.K, lane = threadIdx.x & 31, q = lane & 3, qbase = lane & ~3;
  const int Kq = (K + 3) >> 2, j0 = min(K, q * Kq), j1 = min(K, j0 + Kq);
  float ladj0 = 0.f;
  if (!inverse) {
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) v = circular_wrap(v, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) { v = (v + 1.f) * 0.5f; ladj0 = -0.6931471805599453f; }
  }
  float* pw = p;
  float* ph = p + K;
  float* pd = p + 2 * K;
  float mw = -INFINITY, mh = -INFINITY;
  for (int j = j0; j < j1; ++j) {
    const float a = softclip(pw[j], cfg.inv_ls2), b = softclip(ph[j], cfg.inv_ls2);
    pw[j] = a; ph[j] = b;
    mw = fmaxf(mw, a); mh = fmaxf(mh, b);
  }
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, o));
  }
  float sw = 0.f, sh = 0.f;
  for (int j = j0; j < j1; ++j) {
    const float a = expf(pw[j] - mw), b = expf(ph[j] - mh);
    pw[j] = a; ph[j] = b;
    sw += a; sh += b;
  }
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
  }
  // normalised local prefix sums, then the quad-exclusive offsets
  float cw = 0.f, ch = 0.f;
  for (int j = j0; j < j1; ++j) {
    cw += pw[j] / sw; ch += ph[j] / sh;
    pw[j] = cw; ph[j] = ch;
  }
  float ow = 0.f, oh = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float tw = __shfl_sync(0xffffffffu, cw, qbase + i), th = __shfl_sync(0xffffffffu, ch, qbase + i);
    if (i < q) { ow += tw; oh += th; }
  }
  int count = (q == 0 && -cfg.bound < v) ? 1 : 0;  // left edge knot
  for (int j = j0; j < j1; ++j) {
    const float kw = cfg.bound * (2.f * (ow + pw[j]) - 1.f), kh = cfg.bound * (2.f * (oh + ph[j]) - 1.f);
    pw[j] = kw; ph[j] = kh;
    count += ((inverse ? kh : kw) < v) ? 1 : 0;
  }
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
  __syncwarp();
  int k = count - 1;
  inside = (k >= 0) && (k < K);
  k = k < 0 ? k + K : (k >= K ? k - K : k);
  bin = k;
  const float x0 = (k == 0) ? -cfg.bound : pw[k - 1], x1 = pw[k];
  const float y0 = (k == 0) ? -cfg.bound : ph[k - 1], y1 = ph[k];
  const float d0 = (k == 0) ? 1.f : expf(softclip(pd[k - 1], cfg.inv_ls1));
  const float d1 = (k == K - 1) ? 1.f : expf(softclip(pd[k], cfg.inv_ls1));
  const float s = (y1 - y0) / (x1 - x0);
  const float msk = inside ? 1.f : 0.f;
  if (!inverse) {
    const float z = msk * (v - x0) / (x1 - x0);
    const float omz = 1.f - z;
    const float den = s + (d0 + d1 - 2.f * s) * z * omz;
    const float y = y0 + (y1 - y0) * (s * z * z + d0 * z * omz) / den;
    const float jac = s * s * (2.f * s * z * omz + d0 * omz * omz + d1 * z * z) / (den * den);
    out = inside ? y : v;
    ladj = msk * logf(jac) + ladj0;
  } else {
    const float y_ = msk * (v - y0);
    const float a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2.f * s);
    const float b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2.f * s);
    const float c = -s * y_;
    const float z = 2.f * c / (-b - sqrtf(b * b - 4.f * a * c));
    float x = x0 + z * (x1 - x0);
    x = inside ? x : v;
    if (cfg.univariate == MF_UNIV_CIRCULAR_RQS) x = circular_wrap(x, cfg.bound);
    if (cfg.univariate == MF_UNIV_PS_RQS) x = 2.f * x - 1.f;
    out = x;
    ladj = 0.f;
  }
}

struct WideSmem {
  int ld0, ldA, ldP;
  size_t in0, act, par, wst, xcur, xnew, ytgt, ladj, total;  // offsets in floats
};

WideSmem wide_smem(const FlowLayout& L, int hmax) {
  WideSmem w;
  w.ld0 = L.kpad[0] + 8;  // row stride == 8 (mod 32): conflict-free 64-bit A-fragment loads
  w.ldA = hmax + 8;
  w.ldP = L.pw_ld + 5;    // odd: thread-per-row spline reads hit 32 distinct banks
  size_t o = 0;
  w.in0 = o; o += (size_t)kWM * w.ld0;
  w.act = o; o += (size_t)kWM * w.ldA;
  w.par = o; o += (size_t)kWM * w.ldP;
  w.wst = o; o += (size_t)kStageFloats;
  w.xcur = o; o += (size_t)kWM * L.D;
  w.xnew = o; o += (size_t)kWM * L.D;
  w.ytgt = o; o += (size_t)kWM * L.D;
  w.ladj = o; o += kWM;
  w.total = o;
  return w;
}

template <int NB>
__global__ void __launch_bounds__(kWThreads, 1)
flow_mma_kernel(const __grid_constant__ FlowLayout L, const __grid_constant__ FlowKernelArgs args,
                const __grid_constant__ BaseParams bp, const __grid_constant__ WideSmem S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sm = reinterpret_cast<float*>(smem_raw);
  float* in0 = sm + S.in0;    // [64][ld0]  (x | ctx | 0-pad), pre-split pairs
  float* act = sm + S.act;    // [64][ldA]  hidden activations (pre-split pairs), updated in place layer by layer
  float* par = sm + S.par;    // [64][ldP]  spline parameters of the current feature chunk (fp32)
  float* wstage = sm + S.wst; // weight ring
  float* xcur = sm + S.xcur;  // [64][D]    current x (fp32)
  float* xnew = sm + S.xnew;  // [64][D]
  float* ytgt = sm + S.ytgt;  // [64][D]
  float* ladj_s = sm + S.ladj;
  const float* blob = static_cast<const float*>(args.blob);
  const float* gx = static_cast<const float*>(args.x);
  const float* gc = static_cast<const float*>(args.ctx);
  const int D = L.D, C = L.C, Lh = L.n_layers - 1;
  const int tid = threadIdx.x;
  const int npair0 = L.kpad[0] >> 1;  // column pairs of the input tile

  SplineCfg<float> cfg;
  cfg.K = L.K;
  cfg.univariate = args.univariate;
  cfg.bound = (float)args.bound;
  cfg.inv_ls2 = args.slope > 0.0 ? (float)(2.0 / log(args.slope)) : 0.f;
  cfg.inv_ls1 = args.slope > 0.0 ? (float)(1.0 / log(args.slope)) : 0.f;

  for (long long tile0 = (long long)blockIdx.x * kWM; tile0 < args.n; tile0 += (long long)gridDim.x * kWM) {
    const int rows = (int)min((long long)kWM, args.n - tile0);
    // value of column c of the input row m: x (or the base sample) for c < D, context for c < D + C, else 0
    auto in_value = [&](int m, int c) -> float {
      if (m >= rows) return 0.f;
      if (c < D) return xcur[(size_t)m * D + c];
      if (c < D + C) return gc[((tile0 + m) / args.ctx_group) * C + (c - D)];
      return 0.f;
    };
    for (int i = tid; i < kWM * D; i += kWThreads) {
      const int m = i / D, c = i - m * D;
      float v = 0.f;
      if (m < rows) {
        v = gx[(tile0 + m) * D + c];
        if (args.mode == 1)  // base sample from caller-supplied noise
          v = (args.base == MF_BASE_DIAG_NORMAL) ? (float)bp.a[c] + (float)bp.b[c] * v
                                                 : (float)bp.a[c] + ((float)bp.b[c] - (float)bp.a[c]) * v;
        else if (bp.xmap_kind != MF_XMAP_NONE)
          v = xmap_apply<float>(bp, v, c);  // input map of log_prob (logit bridge)
      }
      xcur[i] = v;
    }
    for (int i = tid; i < kWM; i += kWThreads) ladj_s[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < kWM * npair0; i += kWThreads) {
      const int m = i / npair0, c = (i - m * npair0) << 1;
      store_split(in0 + (size_t)m * S.ld0 + c, in_value(m, c), in_value(m, c + 1));
    }
    __syncthreads();

    for (int ti = 0; ti < L.T; ++ti) {
      const int t = (args.mode == 0) ? ti : (L.T - 1 - ti);
      const float* tb = blob + (size_t)t * L.t_stride;
      int sweeps = 1;
      if (args.mode == 1) {
        sweeps = args.passes;
        for (int i = tid; i < kWM * D; i += kWThreads) {  // y <- current value, x <- 0 (AutoregressiveTransform._inverse)
          ytgt[i] = xcur[i];
          xcur[i] = 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kWM * ((D + 1) >> 1); i += kWThreads) {
          const int m = i / ((D + 1) >> 1), c = (i - m * ((D + 1) >> 1)) << 1;
          store_split(in0 + (size_t)m * S.ld0 + c, in_value(m, c), in_value(m, c + 1));
        }
        __syncthreads();
      }
      for (int sw = 0; sw < sweeps; ++sw) {
        // ---- hyper-network: layer 0 reads the input tile, the others update `act` in place
        const float* cur = in0;
        int lda = S.ld0;
        for (int l = 0; l < Lh; ++l) {
          mma_layer<NB, true>(cur, lda, L.kpad[l], tb + L.w_off[l], L.ld[l], L.ld[l], tb + L.b_off[l], act, S.ldA, true, wstage,
                              args.klim ? args.klim + (size_t)t * kKlimEntries + 4 * l : nullptr);
          cur = act;
          lda = S.ldA;
        }
        // sampling: only the chunks holding a feature that becomes final in this sweep (flow.cu, same schedule)
        const unsigned long long live = (args.mode == 1 && args.sched) ? args.sched[(size_t)t * MF_MAX_FEATURES + sw] : ~0ull;
        for (int gch = 0; gch < L.n_chunks; ++gch) {
          if (!((live >> (gch * L.fpc)) & ((1ull << min(L.fpc, D - gch * L.fpc)) - 1ull))) {
            const int f0s = gch * L.fpc, nfs = min(L.fpc, D - f0s);
            for (int p = tid; p < kWM * nfs; p += kWThreads) {
              const int m = p % kWM, f = f0s + p / kWM;
              xnew[(size_t)m * D + f] = xcur[(size_t)m * D + f];
            }
            continue;
          }
          const float* wl = tb + L.w_off[Lh] + (size_t)gch * L.chunk_stride;
          const float* bl = tb + L.b_off[Lh] + (size_t)gch * L.chunk_stride;
          mma_layer<1, false>(cur, lda, L.kpad[Lh], wl, L.pw_ld, L.pw_ld, bl, par, S.ldP, false, wstage,
                              args.klim ? args.klim + (size_t)t * kKlimEntries + 32 + gch : nullptr);
          // ---- splines of the features of this chunk: four lanes per (row, feature)
          const int f0 = gch * L.fpc;
          const int nf = min(L.fpc, D - f0);
          for (int p0 = 0; p0 < kWM * nf; p0 += kWThreads / 4) {  // warp-uniform trip count
            const int p = p0 + (tid >> 2);
            const int m = p % kWM, fl = min(p / kWM, nf - 1), f = f0 + fl;
            const bool live_q = (p < kWM * nf) && (m < rows);
            const bool final_f = (live >> f) & 1ull;  // provisional features of this sweep keep their current value
            float* pp = par + (size_t)m * S.ldP + fl * L.PF;
            const float v = (args.mode == 0) ? xcur[(size_t)m * D + f] : ytgt[(size_t)m * D + f];
            float out, la;
            int bin;
            bool ins;
            rqs_quad(pp, cfg, args.mode == 1, v, out, la, bin, ins);
            if (live_q && (tid & 3) == 0) {
              xnew[(size_t)m * D + f] = final_f ? out : xcur[(size_t)m * D + f];
              if (args.mode == 0) {
                ytgt[(size_t)m * D + f] = la;  // per-feature ladj, summed in feature order below
                if (args.bins) args.bins[((size_t)t * args.n + tile0 + m) * D + f] = (int8_t)bin;
                if (args.inside) args.inside[((size_t)t * args.n + tile0 + m) * D + f] = ins ? 1 : 0;
              }
            }
          }
          __syncthreads();
        }
        __syncthreads();
        for (int i = tid; i < kWM * D; i += kWThreads) {
          const int m = i / D;
          if (m < rows) xcur[i] = xnew[i];
        }
        if (args.mode == 0) {
          for (int m = tid; m < rows; m += kWThreads) {
            float sacc = 0.f;
            for (int f = 0; f < D; ++f) sacc += ytgt[(size_t)m * D + f];
            ladj_s[m] += sacc;
          }
        }
        __syncthreads();
        for (int i = tid; i < kWM * ((D + 1) >> 1); i += kWThreads) {  // refresh the x pairs of the MMA operand
          const int m = i / ((D + 1) >> 1), c = (i - m * ((D + 1) >> 1)) << 1;
          store_split(in0 + (size_t)m * S.ld0 + c, in_value(m, c), in_value(m, c + 1));
        }
        __syncthreads();
      }
    }

    // ---- outputs
    float* gz = static_cast<float*>(args.z);
    if (gz)
      for (int i = tid; i < rows * D; i += kWThreads) {
        float v = xcur[i];
        if (args.mode == 1 && bp.xmap_kind != MF_XMAP_NONE) v = xmap_apply<float>(bp, v, i % D);  // output map of sample
        gz[tile0 * D + i] = v;
      }
    if (args.mode == 0) {
      for (int m = tid; m < rows; m += kWThreads) {
        const float la = ladj_s[m];
        if (args.ladj) static_cast<float*>(args.ladj)[tile0 + m] = la;
        if (args.logprob) {
          float lp = 0.f;
          for (int f = 0; f < D; ++f)
            lp += base_log_prob_1<float>(args.base, xcur[(size_t)m * D + f], (float)bp.a[f], (float)bp.b[f]);
          static_cast<float*>(args.logprob)[tile0 + m] = lp + la;
        }
      }
    }
    __syncthreads();
  }
}

// in place: two consecutive k-rows [ld] of fp32 weights -> one pair row of [ld] {half2 hi, half2 lo} slots
__global__ void flow_mma_presplit_kernel(float* W, int kpad, int ld) {
  extern __shared__ float rowbuf[];  // [2][ld]
  for (int kp = blockIdx.x; kp < kpad / 2; kp += gridDim.x) {
    float* base = W + (size_t)kp * 2 * ld;
    for (int i = threadIdx.x; i < 2 * ld; i += blockDim.x) rowbuf[i] = base[i];
    __syncthreads();
    for (int n = threadIdx.x; n < ld; n += blockDim.x) store_split(base + 2 * n, rowbuf[n], rowbuf[ld + n]);
    __syncthreads();
  }
}

// one thread per 128-column block: k16-steps up to the last pair row holding a nonzero weight (presplit layout)
__global__ void flow_mma_klimit_kernel(const float* W, int kpad, int ld, unsigned short* out) {
  const int nb = blockIdx.x * blockDim.x + threadIdx.x;
  if (nb * 128 >= ld) return;
  int last = -1;
  for (int kp = kpad / 2 - 1; kp >= 0 && last < 0; --kp) {
    const uint2* row = reinterpret_cast<const uint2*>(W + (size_t)kp * 2 * ld) + nb * 128;
    for (int n = 0; n < 128; ++n)
      if ((row[n].x & 0x7fff7fffu) != 0u) { last = kp; break; }
  }
  out[nb] = (unsigned short)(last < 0 ? 1 : last / 8 + 1);
}

}  // namespace

int flow_mma_presplit(const FlowLayout& L, float* blob, unsigned short* klim, cudaStream_t st) {
  MF_CUDA_CHECK(cudaMemsetAsync(klim, 0xFF, kKlimBytes, st));  // 0xFFFF = no limit for entries nobody computes
  const int Lh = L.n_layers - 1;
  for (int t = 0; t < L.T; ++t) {
    float* tb = blob + (size_t)t * L.t_stride;
    for (int l = 0; l < Lh; ++l) {
      flow_mma_presplit_kernel<<<64, 256, 2 * L.ld[l] * sizeof(float), st>>>(tb + L.w_off[l], L.kpad[l], L.ld[l]);
      int rc = post_launch("flow_mma_presplit_kernel");
      if (rc != MF_OK) return rc;
      if (l > 0 && l < 8) {  // layer 0 reads (x | context): the context columns come last, nothing to trim
        flow_mma_klimit_kernel<<<1, 32, 0, st>>>(tb + L.w_off[l], L.kpad[l], L.ld[l], klim + (size_t)t * kKlimEntries + 4 * l);
        rc = post_launch("flow_mma_klimit_kernel");
        if (rc != MF_OK) return rc;
      }
    }
    for (int g = 0; g < L.n_chunks; ++g) {
      flow_mma_presplit_kernel<<<64, 256, 2 * L.pw_ld * sizeof(float), st>>>(tb + L.w_off[Lh] + (size_t)g * L.chunk_stride,
                                                                            L.kpad[Lh], L.pw_ld);
      int rc = post_launch("flow_mma_presplit_kernel");
      if (rc != MF_OK) return rc;
      if (g < kKlimEntries - 32) {
        flow_mma_klimit_kernel<<<1, 32, 0, st>>>(tb + L.w_off[Lh] + (size_t)g * L.chunk_stride, L.kpad[Lh], L.pw_ld,
                                                klim + (size_t)t * kKlimEntries + 32 + g);
        rc = post_launch("flow_mma_klimit_kernel");
        if (rc != MF_OK) return rc;
      }
    }
  }
  return MF_OK;
}

int flow_mma_run(const FlowLayout& L, const FlowKernelArgs& a, const BaseParams& bp, cudaStream_t st) {
  int hmax = 0;
  for (int l = 0; l < L.n_layers - 1; ++l) hmax = std::max(hmax, L.ld[l]);
  if (hmax > 512 || L.pw_ld != 128) {
    set_error("mma gemm mode needs hidden widths <= 512 and 3*bins-1 <= 128 (got width %d, parameter row %d)", hmax, L.pw_ld);
    return MF_EUNSUPPORTED;
  }
  const WideSmem S = wide_smem(L, hmax);
  const size_t smem = S.total * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("mma gemm mode: tile does not fit shared memory (%zu bytes)", smem);
    return MF_EUNSUPPORTED;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (a.n + kWM - 1) / kWM;
  const unsigned grid = (unsigned)std::min<long long>(tiles, sms);
  if (hmax <= 256) {  // NB = 128-column blocks of the widest hidden layer, all accumulated in registers
    MF_CUDA_CHECK(cudaFuncSetAttribute(flow_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flow_mma_kernel<2><<<grid, kWThreads, smem, st>>>(L, a, bp, S);
  } else {
    MF_CUDA_CHECK(cudaFuncSetAttribute(flow_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    flow_mma_kernel<4><<<grid, kWThreads, smem, st>>>(L, a, bp, S);
  }
  return post_launch("flow_mma_kernel");
}

}  // namespace mf
