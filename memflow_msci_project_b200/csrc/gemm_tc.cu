// gemm_tc.cu -- C[M,N] = A[M,K] * B[N,K]^T (+ bias, ReLU) in fp32-class arithmetic on tcgen05 (mf_gemm_nt_3xf16).
//
// The GEMMs of the hyper-network BACKWARD of the training step (BASELINE config 4; reference:
// scripts/run_model_labframe_sampling.py:499-501, scripts/run_transferFlow_paperVersion_AllPartons_btag.py:385, where
// torch autograd runs them as cuBLAS fp32 SIMT kernels -- 42 of the 68 ms of a likelihood step):
//   forward recompute   H_l  = relu(H_{l-1} W_l^T + b_l)        A = H_{l-1} [rows, in],   B = W_l   [out, in]
//   input gradient      dH   = dZ W_l                           A = dZ      [rows, out],  B = W_l^T [in, out]
//   weight gradient     dW_l = dZ^T H_{l-1}                     A = dZ^T    [out, rows],  B = H^T   [in, rows]   (split-K)
// Arithmetic as in flow_tc.cu: every fp32 operand value is split into fp16 hi = rn(v), lo = rn(v - hi) and each product
// is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM.  Gradients are far below the fp16 range, so each
// operand can be multiplied by a power of two read from device memory before the split (d_scale_a / d_scale_b; the
// result is divided by their product) -- the caller derives it from the operand's largest magnitude without a host sync.
//
// One CTA = one 128-row tile of A and one K-range; 256 threads: all eight warps convert the operand K-blocks (64
// columns) from global memory into the swizzled fp16 hi/lo layout of tc_common.cuh; one thread issues the 12 MMAs of a
// K-block and commits them to the buffer's mbarrier; warps 0-3 (thread = row = TMEM lane) run the epilogue, transposing
// through shared memory so that C is written in full row pieces.  Long K-ranges (the weight gradient) use two operand
// buffers -- K-block i + 1 is staged while the tensor core works on K-block i --, short ones a single buffer so that
// two or three CTAs share an SM and hide each other's load / MMA / store latencies.  With k_split > 1 the partial
// sums of the K-ranges are added into C with red.global.add.f32 (C must be zeroed by the caller).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "spline_reg.cuh"
#include "tc_common.cuh"

namespace mf {
namespace {

constexpr int kGemmThreads = 256;

struct GemmArgs {
  const float* A; long long lda;
  const float* B; long long ldb;
  float* C; long long ldc;
  long long M, K;
  int N, Np;            // Np = N rounded up to 16 (MMA N)
  const float* bias;
  int relu, k_split;
  long long kb_per_split;  // K-blocks of 64 per K-range
  const float* scale_a; const float* scale_b;
  int vec_a, vec_b, vec_m;  // rows are 16-byte aligned: float4 loads
  int trans;            // 0: C = A B^T with A [M,K], B [N,K] (K contiguous); 1: C = A^T B with A [K,M], B [K,N] (K = rows)
  const float* mask; long long ldm;  // optional, same shape as A: A is multiplied by (mask > 0) (ReLU derivative)
  const float* out_mask; long long ldo;  // optional, shaped like C: C is multiplied by (out_mask > 0) (trans = 0, no split)
  float* colsum;        // trans = 1: optional [M] output, += sum_k A[k, m] (the bias gradient), as column N of the product
  int nbuf;             // operand buffers: 2 when a CTA walks many K-blocks (staging overlaps the MMAs), else 1 (more CTAs per SM)
  int tmem_cols;        // power of two >= Np
};

__device__ __forceinline__ void store_half_row(const float* v, float scale, unsigned char* hi, unsigned char* lo, int r,
                                               int half);

// 32 consecutive columns [k0, k0 + 32) of one operand row -> hi / lo planes of the K-block (half = 0 / 1: which half of
// the 64 columns); src == nullptr: a row past the end of the matrix (zeros)
__device__ __forceinline__ void stage_half_row(const float* src, const float* msk, long long k_left, bool vec, bool vec_m,
                                               float scale, unsigned char* hi, unsigned char* lo, int r, int half) {
  float v[32];
  if (src == nullptr || k_left <= 0) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = 0.f;
  } else if (vec && k_left >= 32) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src) + q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = (e < k_left) ? __ldg(src + e) : 0.f;
  }
  if (msk != nullptr && src != nullptr) {
    if (vec_m && k_left >= 32) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(msk) + q);
        v[4 * q] = t.x > 0.f ? v[4 * q] : 0.f; v[4 * q + 1] = t.y > 0.f ? v[4 * q + 1] : 0.f;
        v[4 * q + 2] = t.z > 0.f ? v[4 * q + 2] : 0.f; v[4 * q + 3] = t.w > 0.f ? v[4 * q + 3] : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = (e < k_left && __ldg(msk + e) > 0.f) ? v[e] : 0.f;
    }
  }
  store_half_row(v, scale, hi, lo, r, half);
}

// 32 consecutive K values of operand row r (half = which half of the 64-wide K-block) -> fp16 hi / lo planes
__device__ __forceinline__ void store_half_row(const float* v, float scale, unsigned char* hi, unsigned char* lo, int r,
                                               int half) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {  // four 16-byte chunks of 8 columns
    uint4 h, l;
    tc::split_f16(v[8 * c] * scale, v[8 * c + 1] * scale, h.x, l.x);
    tc::split_f16(v[8 * c + 2] * scale, v[8 * c + 3] * scale, h.y, l.y);
    tc::split_f16(v[8 * c + 4] * scale, v[8 * c + 5] * scale, h.z, l.z);
    tc::split_f16(v[8 * c + 6] * scale, v[8 * c + 7] * scale, h.w, l.w);
    const int chunk = half * 4 + c;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// trans = 1: one K-block (64 rows k of a [K, C] row-major matrix) -> the K-major operand tile [R rows c][64 k].  A thread
// takes one k and 32 consecutive columns (128 contiguous bytes of global memory); consecutive lanes hold consecutive k, so
// the 2-byte shared-memory stores of a warp fall into one 128-byte row (no bank conflicts).
// (Measured alternatives for the 128 x 128 weight gradient of 655 360 rows, 0.59 ms with this mapping: eight k x one
// column per lane -- coalesced scalar loads, one 16-byte store per lane -- 0.76 ms; four k x eight columns per thread --
// 16-byte loads, 8-byte stores -- 0.73 ms.  The mapping whose threads read the longest contiguous piece of a row wins:
// the kernel is bound by its global loads, not by the shared-memory stores.)
__device__ __forceinline__ void stage_tn(const float* base, long long ld, const float* msk, long long ldm, long long k0,
                                         long long K, long long c0, long long C, int R, int ones_col, float scale, bool vec,
                                         bool vec_m, unsigned char* hi, unsigned char* lo, int tid) {
  const int quarters = (R + 31) / 32;
  for (int idx = tid; idx < 64 * quarters; idx += kGemmThreads) {
    const int k = idx & 63, q = idx >> 6;
    const long long kk = k0 + k, cc = c0 + 32 * q;
    float v[32];
    const bool live = kk < K;
    if (live && vec && cc + 32 <= C) {
      const float4* src = reinterpret_cast<const float4*>(base + kk * ld + cc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(src + j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (live && cc + j < C) ? __ldg(base + kk * ld + cc + j) : 0.f;
    }
    if (msk != nullptr) {
      if (live && vec_m && cc + 32 <= C) {
        const float4* mp = reinterpret_cast<const float4*>(msk + kk * ldm + cc);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(mp + j);
          v[4 * j] = t.x > 0.f ? v[4 * j] : 0.f; v[4 * j + 1] = t.y > 0.f ? v[4 * j + 1] : 0.f;
          v[4 * j + 2] = t.z > 0.f ? v[4 * j + 2] : 0.f; v[4 * j + 3] = t.w > 0.f ? v[4 * j + 3] : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (live && cc + j < C && __ldg(msk + kk * ldm + cc + j) > 0.f) ? v[j] : 0.f;
      }
    }
    if (ones_col >= 0 && ones_col >= 32 * q && ones_col < 32 * q + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (32 * q + j == ones_col) v[j] = live ? 1.f / scale : 0.f;  // (the operand scale is applied below)
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int r = 32 * q + j;
      if (r < R) {
        const float x = v[j] * scale;
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
        *reinterpret_cast<__half*>(hi + off) = h;
        *reinterpret_cast<__half*>(lo + off) = l;
      }
    }
  }
}

// Epilogue of warps 0-3 through shared memory (the operand buffers are free by then): a thread owns a ROW of the
// accumulator, the rows of C are contiguous in global memory -- each warp transposes 32 rows x 64 columns through a
// scratch tile (row stride 65 floats: conflict-free both ways) and writes / adds full 256-byte row pieces.
__device__ __forceinline__ void gemm_epilogue(const GemmArgs& g, float* scratch_base, uint32_t tmem, long long m0, float sa,
                                              float sb, int warp, int lane) {
  {
      // Epilogue through shared memory (the operand buffers are free now): a thread owns a ROW of the accumulator, the
      // rows of C are contiguous in global memory -- each warp transposes 32 rows x 64 columns through a scratch tile
      // (row stride 65 floats: conflict-free both ways) and writes full 256-byte row pieces.
      float* scratch = scratch_base + (size_t)warp * 32 * 65;
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      const float inv = 1.f / (sa * sb);
      const bool add_bias = g.bias != nullptr && blockIdx.y == 0;
      const long long row0 = m0 + warp * 32;
      const int rows_w = (int)max(0LL, min(32LL, g.M - row0));  // rows of this warp inside the matrix
      for (int cb = 0; cb < g.Np; cb += 64) {
        // this block's biases: lane l holds columns cb + l and cb + 32 + l (one coalesced load instead of an L2 round
        // trip per element: with this kernel's shared-memory footprint little L1 is left), spread by shuffles
        const float bias_lo = (add_bias && cb + lane < g.N) ? __ldg(g.bias + cb + lane) : 0.f;
        const float bias_hi = (add_bias && cb + 32 + lane < g.N) ? __ldg(g.bias + cb + 32 + lane) : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c0 = cb + 16 * q;
          if (c0 < g.Np) {   // warp-uniform
            float v[16];
            tc::tmem_ld16(taddr + c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float bv = __shfl_sync(0xffffffffu, (q < 2) ? bias_lo : bias_hi, (16 * q + e) & 31);
              float x = v[e] * inv + bv;
              if (g.relu) x = fmaxf(x, 0.f);
              scratch[lane * 65 + 16 * q + e] = x;
            }
          }
        }
        __syncwarp();
        // rows of C: eight rows per step, loads first, then stores (a rolled load -> store loop is a chain of
        // shared-memory latencies: 4.5 k cycles per warp and tile, ncu)
        for (int r0 = 0; r0 < rows_w; r0 += 8) {
          float x0[8], x1[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            x0[u] = scratch[(r0 + u) * 65 + lane];
            x1[u] = scratch[(r0 + u) * 65 + lane + 32];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (r0 + u < rows_w) {
              const long long row = row0 + r0 + u;
              float* dst = g.C + row * g.ldc + cb;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                const float x = h ? x1[u] : x0[u];
                if (cb + c < g.N) {
                  if (g.k_split > 1) atomicAdd(dst + c, x);
                  else if (g.out_mask != nullptr) dst[c] = (__ldg(g.out_mask + row * g.ldo + cb + c) > 0.f) ? x : 0.f;
                  else dst[c] = x;
                } else if (g.colsum != nullptr && cb + c == g.N) {
                  atomicAdd(g.colsum + row, x);
                }
              }
            }
          }
        }
        __syncwarp();
      }
      }
}

__global__ void __launch_bounds__(kGemmThreads, 3) gemm_nt_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_plane = 128u * 128u;                 // [128 rows x 64] 16-bit
  const uint32_t b_plane = (uint32_t)g.Np * 128u;       // [Np rows x 64] 16-bit
  const uint32_t buf_bytes = 2u * a_plane + 2u * b_plane;
  const long long m0 = (long long)blockIdx.x * 128;
  const long long n_kb = (g.K + 63) / 64;
  const long long kb0 = (long long)blockIdx.y * g.kb_per_split;
  const long long kb1 = min(n_kb, kb0 + g.kb_per_split);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)g.tmem_cols);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;
  const float sa = g.scale_a ? __ldg(g.scale_a) : 1.f, sb = g.scale_b ? __ldg(g.scale_b) : 1.f;
  const uint32_t idesc = tc::make_idesc_f16(0, 128, g.Np);

  long long it = 0;
  for (long long kb = kb0; kb < kb1; ++kb, ++it) {
    const int b = (g.nbuf == 2) ? (int)(it & 1) : 0;
    unsigned char* a_hi = smem + (size_t)b * buf_bytes;
    unsigned char* a_lo = a_hi + a_plane;
    unsigned char* b_hi = a_lo + a_plane;
    unsigned char* b_lo = b_hi + b_plane;
    if (it >= g.nbuf) mbar_wait(&bar[b], (uint32_t)(((it / g.nbuf) - 1) & 1));  // the MMAs that read this buffer have retired
    const long long k0 = kb * 64;
    if (g.trans) {
      stage_tn(g.A, g.lda, g.mask, g.ldm, k0, g.K, m0, g.M, 128, -1, sa, g.vec_a != 0, g.vec_m != 0, a_hi, a_lo, tid);
      stage_tn(g.B, g.ldb, nullptr, 0, k0, g.K, 0, g.N, g.Np, g.colsum ? g.N : -1, sb, g.vec_b != 0, false, b_hi, b_lo, tid);
    } else {
      // A: 128 rows x 2 halves = 256 half rows, one per thread
      {
        const int r = tid >> 1, half = tid & 1;
        const long long row = m0 + r, kk = k0 + half * 32;
        stage_half_row(row < g.M ? g.A + row * g.lda + kk : nullptr, (row < g.M && g.mask) ? g.mask + row * g.ldm + kk : nullptr,
                       g.K - kk, g.vec_a != 0, g.vec_m != 0, sa, a_hi, a_lo, r, half);
      }
      // B: Np rows x 2 halves
      for (int i = tid; i < 2 * g.Np; i += kGemmThreads) {
        const int r = i >> 1, half = i & 1;
        const long long kk = k0 + half * 32;
        stage_half_row(r < g.N ? g.B + (long long)r * g.ldb + kk : nullptr, nullptr, g.K - kk, g.vec_b != 0, false, sb, b_hi, b_lo, r, half);
      }
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_thread_sync();
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = tc::make_smem_desc(ah + ks * 32u), dal = tc::make_smem_desc(al + ks * 32u);
        const uint64_t dbh = tc::make_smem_desc(bh + ks * 32u), dbl = tc::make_smem_desc(bl + ks * 32u);
        tc::mma_f16_ss(tmem, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        tc::mma_f16_ss(tmem, dah, dbl, idesc, 1u);
        tc::mma_f16_ss(tmem, dal, dbh, idesc, 1u);
      }
      tc::mma_commit(&bar[b]);
    }
  }
  if (it > 0) {
    // the last commit completes when ALL MMAs of the issuing thread have retired
    const long long last = it - 1;
    mbar_wait(&bar[(g.nbuf == 2) ? (last & 1) : 0], (uint32_t)((last / g.nbuf) & 1));
    tc::fence_after_thread_sync();
    if (warp < 4) gemm_epilogue(g, reinterpret_cast<float*>(smem), tmem, m0, sa, sb, warp, lane);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)g.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Persistent variant of the non-transposed product for K = 64 or 128 and a CONTIGUOUS A (lda = K: every hidden-layer
// GEMM of the training step): one CTA per SM walks 128-row tiles of A.  B (the weights) is converted ONCE per CTA and
// stays in shared memory; a whole raw fp32 tile of A (128 rows x K, one contiguous block of global memory) arrives by ONE
// bulk async copy completing on an mbarrier, requested as soon as the previous tile has been converted, so the
// global-memory latency -- what bounds the one-tile-per-CTA kernel above (25 % issue-active, ncu) -- hides behind the
// previous tile's MMAs and epilogue.  All eight warps convert the landed tile: consecutive threads take consecutive
// 16-byte pieces (conflict-free reads), each piece = four k-values of one row = one 8-byte store per operand plane.
// (A first persistent version fetched the tile as 128 row pieces of 256 bytes, one bulk copy per thread: 0.54 ms per
// GEMM instead of 0.40 -- small bulk copies cost ~100 cycles each in the copy engine.)
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_nt_persistent_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t raw_full, mma_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_kb = (int)(g.K / 64);
  const int q4 = (int)(g.K / 4), q4_shift = (g.K == 128) ? 5 : 4;  // float4 pieces per row
  const uint32_t b_plane = (uint32_t)g.Np * 128u, a_plane = 128u * 128u;
  unsigned char* b_conv = smem;                                   // [n_kb][hi | lo]
  unsigned char* conv = b_conv + (size_t)n_kb * 2 * b_plane;      // [n_kb][hi | lo], 1024-byte aligned (b_plane % 2048 == 0)
  unsigned char* raw = conv + (size_t)n_kb * 2 * a_plane;         // [128 rows x K] fp32
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)g.tmem_cols);
  if (tid == 0) {
    mbar_init(&raw_full, 1);
    mbar_init(&mma_done, 1);
    fence_mbar_init();
  }
  const float sa = g.scale_a ? __ldg(g.scale_a) : 1.f, sb = g.scale_b ? __ldg(g.scale_b) : 1.f;
  // B: all K-blocks, converted once
  for (int kb = 0; kb < n_kb; ++kb)
    for (int i = tid; i < 2 * g.Np; i += kGemmThreads) {
      const int r = i >> 1, half = i & 1;
      const long long kk = (long long)kb * 64 + half * 32;
      stage_half_row(r < g.N ? g.B + (long long)r * g.ldb + kk : nullptr, nullptr, g.K - kk, g.vec_b != 0, false, sb,
                     b_conv + (size_t)kb * 2 * b_plane, b_conv + (size_t)kb * 2 * b_plane + b_plane, r, half);
    }
  fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = tc::make_idesc_f16(0, 128, g.Np);
  const long long m_tiles = (g.M + 127) / 128;
  auto request = [&](long long tile) {  // the raw tile: one contiguous block
    if (tid == 0 && tile < m_tiles) {
      const long long rows = min((long long)128, g.M - tile * 128);
      const uint32_t bytes = (uint32_t)(rows * g.K * 4);
      mbar_expect_tx(&raw_full, bytes);
      bulk_g2s(raw, g.A + tile * 128 * g.K, bytes, &raw_full);
    }
  };
  request(blockIdx.x);
  long long it = 0;
  for (long long tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
    const long long m0 = tile * 128;
    const int rows = (int)min((long long)128, g.M - m0);
    mbar_wait(&raw_full, (uint32_t)(it & 1));  // the raw tile has landed
    // (the operand buffers are free: this tile's predecessor waited for its MMAs before its epilogue)
    for (int f = tid; f < 128 * q4; f += kGemmThreads) {
      const int r = f >> q4_shift, c4 = f & (q4 - 1);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) {
        t = reinterpret_cast<const float4*>(raw)[f];
        if (g.mask != nullptr) {
          const float4 mk = __ldg(reinterpret_cast<const float4*>(g.mask + (m0 + r) * g.ldm) + c4);
          t.x = mk.x > 0.f ? t.x : 0.f; t.y = mk.y > 0.f ? t.y : 0.f; t.z = mk.z > 0.f ? t.z : 0.f; t.w = mk.w > 0.f ? t.w : 0.f;
        }
      }
      uint2 h, l;
      tc::split_f16(t.x * sa, t.y * sa, h.x, l.x);
      tc::split_f16(t.z * sa, t.w * sa, h.y, l.y);
      const int kb = c4 >> 4, k = (c4 & 15) * 4;
      unsigned char* hi = conv + (size_t)kb * 2 * a_plane;
      const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2);
      *reinterpret_cast<uint2*>(hi + off) = h;
      *reinterpret_cast<uint2*>(hi + a_plane + off) = l;
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
    __syncthreads();           // (also: every thread is done with the raw tile)
    if (tid == 0) {
      tc::fence_after_thread_sync();
      for (int kb = 0; kb < n_kb; ++kb) {
        const uint32_t ah = smem_u32(conv + (size_t)kb * 2 * a_plane), al = ah + a_plane;
        const uint32_t bh = smem_u32(b_conv + (size_t)kb * 2 * b_plane), bl = bh + b_plane;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dah = tc::make_smem_desc(ah + ks * 32u), dal = tc::make_smem_desc(al + ks * 32u);
          const uint64_t dbh = tc::make_smem_desc(bh + ks * 32u), dbl = tc::make_smem_desc(bl + ks * 32u);
          tc::mma_f16_ss(tmem, dah, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          tc::mma_f16_ss(tmem, dah, dbl, idesc, 1u);
          tc::mma_f16_ss(tmem, dal, dbh, idesc, 1u);
        }
      }
      tc::mma_commit(&mma_done);
    }
    request(tile + gridDim.x);  // the raw buffer is free again: the next tile travels during the MMAs and the epilogue
    // output mask (input gradient of a layer fed by a ReLU): this thread's mask values of the whole tile are requested
    // NOW, into registers, so that their L2 latency passes under the MMAs (fetched inside the epilogue they doubled it)
    float mk[5][2][8];
    if (g.out_mask != nullptr) {
      const int wq_ = warp & 3, wh_ = warp >> 2, c_ = lane & 15, rsub_ = lane >> 4;
      const long long row0_ = m0 + wq_ * 32;
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int col = 16 * wh_ + 32 * j + c_;
            const long long row = row0_ + 16 * h + 2 * u + rsub_;
            mk[j][h][u] = (col < g.N && row < g.M) ? __ldg(g.out_mask + row * g.ldo + col) : 0.f;
          }
    }
    mbar_wait(&mma_done, (uint32_t)(it & 1));
    tc::fence_after_thread_sync();
    {
      // Epilogue by ALL eight warps: warp w reads TMEM lanes 32 (w % 4) .. + 31 (its rows); warps 0-3 take the 16-column
      // blocks with an even index, warps 4-7 the odd ones (with four warps the epilogue was ~70 % of a tile, ncu: the
      // other four waited at the next barrier).  Each block goes through the warp's scratch tile (row stride 17 floats)
      // so that C is written in 64-byte row pieces, two rows per store.
      const int wq = warp & 3, wh = warp >> 2;
      const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16);
      const float inv = 1.f / (sa * sb);
      const long long row0 = m0 + wq * 32;
      const int rows_w = (int)max(0LL, min(32LL, g.M - row0));
      float* sc = reinterpret_cast<float*>(raw + (size_t)128 * g.K * 4) + (size_t)warp * 32 * 17;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int c0 = 16 * wh + 32 * j;
        if (c0 >= g.Np) break;
        const float bias_l = (g.bias != nullptr && lane < 16 && c0 + lane < g.N) ? __ldg(g.bias + c0 + lane) : 0.f;
        float v[16];
        tc::tmem_ld16(taddr + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float x = v[e] * inv + __shfl_sync(0xffffffffu, bias_l, e);
          if (g.relu) x = fmaxf(x, 0.f);
          sc[lane * 17 + e] = x;
        }
        __syncwarp();
        const int c = lane & 15, rsub = lane >> 4;
        if (c0 + c < g.N) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r0 = 16 * h;
            if (r0 >= rows_w) break;
            float x0[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x0[u] = sc[(r0 + 2 * u + rsub) * 17 + c];
            if (g.out_mask != nullptr) {
#pragma unroll
              for (int u = 0; u < 8; ++u) x0[u] = mk[j][h][u] > 0.f ? x0[u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (r0 + 2 * u + rsub < rows_w) g.C[(row0 + r0 + 2 * u + rsub) * g.ldc + c0 + c] = x0[u];
          }
        }
        __syncwarp();
      }
      tc::fence_before_thread_sync();
    }
    // (the next tile's conversion only touches shared memory; its MMAs are issued behind the __syncthreads above, i.e.
    // after every warp has left this epilogue)
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)g.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Transposed product (weight gradient) with bulk-copied operands: for contiguous A [K, M <= 128] and B [K, N] a K-block
// of 64 rows is ONE contiguous block of each matrix.  Thread 0 requests the raw fp32 blocks two K-blocks ahead (two
// stages, completing on mbarriers); all warps transpose a landed block out of shared memory: a warp takes 32 consecutive
// columns and a group of eight k -- eight conflict-free row reads, then every lane owns the eight k-values of its
// column = one 16-byte chunk of the K-major operand row.  (The same mapping fed by global loads was the slowest of the
// three tried above; out of shared memory it is ~300 instead of ~1 350 instructions per thread and K-block and no load
// sits on the critical path.)  No ReLU mask here: the caller passes the masked gradient.
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tn_tma_kernel(const __grid_constant__ GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t raw_full[2], mma_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = (int)g.M, N = g.N;
  const uint32_t a_plane = 128u * 128u, b_plane = (uint32_t)g.Np * 128u;
  const uint32_t raw_a = 64u * (uint32_t)M * 4u, raw_b = 64u * (uint32_t)N * 4u;
  const uint32_t raw_stage = (raw_a + raw_b + 127u) / 128u * 128u;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_plane;
  unsigned char* b_hi = a_lo + a_plane;
  unsigned char* b_lo = b_hi + b_plane;
  unsigned char* raw = b_lo + b_plane;
  const long long n_kb = (g.K + 63) / 64;
  const long long kb0 = (long long)blockIdx.y * g.kb_per_split;
  const long long kb1 = min(n_kb, kb0 + g.kb_per_split);
  const long long n_it = kb1 - kb0;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, (uint32_t)g.tmem_cols);
  if (tid == 0) {
    mbar_init(&raw_full[0], 1); mbar_init(&raw_full[1], 1); mbar_init(&mma_done, 1);
    fence_mbar_init();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;
  const float sa = g.scale_a ? __ldg(g.scale_a) : 1.f, sb = g.scale_b ? __ldg(g.scale_b) : 1.f;
  const uint32_t idesc = tc::make_idesc_f16(0, 128, g.Np);
  auto request = [&](long long it) {
    if (tid == 0 && it < n_it) {
      const long long k0 = (kb0 + it) * 64;
      const long long rows = min((long long)64, g.K - k0);
      unsigned char* dst = raw + (size_t)(it & 1) * raw_stage;
      uint64_t* bar = &raw_full[it & 1];
      mbar_expect_tx(bar, (uint32_t)(rows * (M + N) * 4));
      bulk_g2s(dst, g.A + k0 * M, (uint32_t)(rows * M * 4), bar);
      bulk_g2s(dst + raw_a, g.B + k0 * N, (uint32_t)(rows * N * 4), bar);
    }
  };
  // one operand: raw [64 k][C] fp32 -> K-major [R rows c][64 k] hi / lo planes
  auto transpose = [&](const float* src, int C, int R, int rows, int ones_col, float scale, unsigned char* hi, unsigned char* lo) {
    const int blocks = (R + 31) / 32;
    for (int item = warp; item < 8 * blocks; item += kGemmThreads / 32) {
      const int kg = item & 7, cb = item >> 3;
      const int r = 32 * cb + lane;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = 8 * kg + e;
        float x = (k < rows && r < C) ? src[k * C + r] : 0.f;
        if (r == ones_col) x = (k < rows) ? 1.f / scale : 0.f;  // (the operand scale is applied below)
        v[e] = x * scale;
      }
      if (r < R) {
        uint4 h, l;
        tc::split_f16(v[0], v[1], h.x, l.x); tc::split_f16(v[2], v[3], h.y, l.y);
        tc::split_f16(v[4], v[5], h.z, l.z); tc::split_f16(v[6], v[7], h.w, l.w);
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((kg ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
      }
    }
  };
  request(0);
  request(1);
  for (long long it = 0; it < n_it; ++it) {
    const int s = (int)(it & 1);
    const long long k0 = (kb0 + it) * 64;
    const int rows = (int)min((long long)64, g.K - k0);
    mbar_wait(&raw_full[s], (uint32_t)((it >> 1) & 1));          // the raw blocks have landed
    if (it >= 1) mbar_wait(&mma_done, (uint32_t)((it - 1) & 1));  // the previous K-block's MMAs have read the operands
    const float* ra = reinterpret_cast<const float*>(raw + (size_t)s * raw_stage);
    transpose(ra, M, 128, rows, -1, sa, a_hi, a_lo);
    transpose(reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(ra) + raw_a), N, g.Np, rows,
              g.colsum ? N : -1, sb, b_hi, b_lo);
    fence_proxy_async_smem();
    __syncthreads();  // (also: every thread is done with raw stage s)
    if (tid == 0) {
      tc::fence_after_thread_sync();
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = tc::make_smem_desc(ah + ks * 32u), dal = tc::make_smem_desc(al + ks * 32u);
        const uint64_t dbh = tc::make_smem_desc(bh + ks * 32u), dbl = tc::make_smem_desc(bl + ks * 32u);
        tc::mma_f16_ss(tmem, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
        tc::mma_f16_ss(tmem, dah, dbl, idesc, 1u);
        tc::mma_f16_ss(tmem, dal, dbh, idesc, 1u);
      }
      tc::mma_commit(&mma_done);
    }
    request(it + 2);
  }
  if (n_it > 0) {
    mbar_wait(&mma_done, (uint32_t)((n_it - 1) & 1));
    tc::fence_after_thread_sync();
    __syncthreads();  // every warp is past its last operand access: the epilogue scratch overlays the operand planes
    if (warp < 4) gemm_epilogue(g, reinterpret_cast<float*>(smem), tmem, 0, sa, sb, warp, lane);
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)g.tmem_cols);
}

}  // namespace
}  // namespace mf

using namespace mf;

static int gemm_launch(int trans, const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                       int64_t m, int32_t n, int64_t k, const float* d_bias, int32_t relu, int32_t k_split,
                       const float* d_mask, int64_t ldm, const float* d_out_mask, int64_t ldo, float* d_colsum,
                       const float* d_scale_a, const float* d_scale_b, void* stream, const char* who) {
  MF_REQUIRE(m >= 0 && n >= 0 && k >= 0, "%s: negative dimension", who);
  if (m == 0 || n == 0) return MF_OK;
  MF_REQUIRE(d_a && d_b && d_c, "%s: null argument", who);
  MF_REQUIRE(n + (d_colsum ? 1 : 0) <= 256, "%s: n=%d > 256 (one accumulator tile)", who, n);
  MF_REQUIRE(k >= 1, "%s: k must be positive", who);
  if (trans) MF_REQUIRE(lda >= m && ldb >= n && ldc >= n, "%s: leading dimension smaller than the row length", who);
  else MF_REQUIRE(lda >= k && ldb >= k && ldc >= n, "%s: leading dimension smaller than the row length", who);
  MF_REQUIRE(k_split >= 1, "%s: k_split must be >= 1", who);
  MF_REQUIRE(!(relu && k_split > 1), "%s: ReLU cannot be applied to partial sums (k_split > 1)", who);
  GemmArgs g;
  g.A = d_a; g.lda = lda; g.B = d_b; g.ldb = ldb; g.C = d_c; g.ldc = ldc;
  g.M = m; g.K = k; g.N = n; g.Np = (n + (d_colsum ? 1 : 0) + 15) / 16 * 16;
  g.bias = d_bias; g.relu = relu;
  g.trans = trans; g.mask = d_mask; g.ldm = ldm; g.colsum = d_colsum; g.out_mask = d_out_mask; g.ldo = ldo;
  MF_REQUIRE(!(d_out_mask && (trans || k_split > 1)), "%s: an output mask needs the plain (unsplit, non-transposed) product", who);
  const long long n_kb = (k + 63) / 64;
  const long long ks = std::min<long long>(k_split, n_kb);
  g.kb_per_split = (n_kb + ks - 1) / ks;
  const long long splits = (n_kb + g.kb_per_split - 1) / g.kb_per_split;
  g.k_split = (int)splits;
  g.scale_a = d_scale_a; g.scale_b = d_scale_b;
  g.vec_a = ((reinterpret_cast<uintptr_t>(d_a) & 15u) == 0 && lda % 4 == 0) ? 1 : 0;
  g.vec_b = ((reinterpret_cast<uintptr_t>(d_b) & 15u) == 0 && ldb % 4 == 0) ? 1 : 0;
  g.vec_m = (d_mask && (reinterpret_cast<uintptr_t>(d_mask) & 15u) == 0 && ldm % 4 == 0) ? 1 : 0;
  g.nbuf = (g.kb_per_split > 3) ? 2 : 1;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.Np) g.tmem_cols *= 2;
  const size_t smem = std::max<size_t>((size_t)g.nbuf * (2 * 128 * 128 + 2 * (size_t)g.Np * 128), 4 * 32 * 65 * 4);
  const long long m_tiles = (m + 127) / 128;
  // persistent variant: resident weights, bulk-copied A K-blocks (see gemm_nt_persistent_kernel)
  // (MF_GEMM_PERSISTENT=0 selects the one-tile-per-CTA kernel for every shape: the comparison of tools/bench_gemm.py)
  static const bool use_persistent = !(getenv("MF_GEMM_PERSISTENT") != nullptr && getenv("MF_GEMM_PERSISTENT")[0] == '0');
  // transposed product with bulk-copied operands (see gemm_tn_tma_kernel): contiguous A [k, m <= 128] and B [k, n], no mask
  if (use_persistent && trans && d_mask == nullptr && lda == m && ldb == n && m <= 128 && m % 4 == 0 && (n % 4 == 0 || k % 64 == 0) &&  // (bulk copies move multiples of 16 bytes)
     
      (reinterpret_cast<uintptr_t>(d_a) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_b) & 15u) == 0) {
    const size_t raw_stage = ((size_t)64 * (m + n) * 4 + 127) / 128 * 128;
    const size_t tsmem = 2 * 128 * 128 + 2 * (size_t)g.Np * 128 + 2 * raw_stage;
    if (tsmem <= 225 * 1024 && tsmem >= 4 * 32 * 65 * 4) {
      MF_CUDA_CHECK(cudaFuncSetAttribute(gemm_tn_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
      dim3 tgrid(1, (unsigned)splits);
      gemm_tn_tma_kernel<<<tgrid, kGemmThreads, tsmem, static_cast<cudaStream_t>(stream)>>>(g);
      return post_launch("gemm_tn_tma_kernel");
    }
  }
  const bool mask_ok = d_mask == nullptr || (ldm == k && (reinterpret_cast<uintptr_t>(d_mask) & 15u) == 0);
  if (use_persistent && !trans && g.k_split == 1 && (k == 64 || k == 128) && lda == k && (reinterpret_cast<uintptr_t>(d_a) & 15u) == 0 &&
      mask_ok && m_tiles >= 2) {
    const size_t n_kb = (size_t)(k / 64);
    const size_t psmem = n_kb * 2 * (size_t)g.Np * 128 + n_kb * 2 * 128 * 128 + (size_t)128 * k * 4 + 8 * 32 * 17 * 4;
    if (psmem <= 225 * 1024) {
      int dev = 0, sms = 148;
      MF_CUDA_CHECK(cudaGetDevice(&dev));
      MF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      MF_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      const unsigned pgrid = (unsigned)std::min<long long>(m_tiles, sms);
      gemm_nt_persistent_kernel<<<pgrid, kGemmThreads, psmem, static_cast<cudaStream_t>(stream)>>>(g);
      return post_launch("gemm_nt_persistent_kernel");
    }
  }
  MF_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MF_REQUIRE(m_tiles <= 0x7fffffffLL && splits <= 65535, "%s: grid too large", who);
  dim3 grid((unsigned)m_tiles, (unsigned)splits);
  gemm_nt_kernel<<<grid, kGemmThreads, smem, static_cast<cudaStream_t>(stream)>>>(g);
  return post_launch("gemm_nt_kernel");
}

extern "C" int mf_gemm_nt_3xf16(const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                                int64_t m, int32_t n, int64_t k, const float* d_bias, int32_t relu, int32_t k_split,
                                const float* d_mask, int64_t ldm, const float* d_out_mask, int64_t ldo,
                                const float* d_scale_a, const float* d_scale_b, void* stream) {
  return gemm_launch(0, d_a, lda, d_b, ldb, d_c, ldc, m, n, k, d_bias, relu, k_split, d_mask, ldm, d_out_mask, ldo, nullptr,
                     d_scale_a, d_scale_b, stream, "mf_gemm_nt_3xf16");
}

extern "C" int mf_gemm_tn_3xf16(const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                                int64_t m, int32_t n, int64_t k, int32_t k_split, const float* d_mask, int64_t ldm,
                                float* d_colsum, const float* d_scale_a, const float* d_scale_b, void* stream) {
  return gemm_launch(1, d_a, lda, d_b, ldb, d_c, ldc, m, n, k, nullptr, 0, k_split, d_mask, ldm, nullptr, 0, d_colsum,
                     d_scale_a, d_scale_b, stream, "mf_gemm_tn_3xf16");
}
