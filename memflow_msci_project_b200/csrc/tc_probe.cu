// tc_probe.cu -- single-CTA tcgen05 GEMM used by the tests to validate the tensor-core building blocks
// (swizzled operand layout, shared-memory / instruction descriptors, TMEM allocation, MMA issue, commit,
// TMEM read-back) in isolation:  out[128, N] = A[128, K] * W[N, K]^T  with fp32 inputs.
//   mode 0: fp16 hi/lo split of both operands, 3 MMAs per k-step (hi*hi + hi*lo + lo*hi)
//   mode 1: bf16 single MMA        mode 2: fp16 single MMA
//   mode 3: fp16 hi/lo split with the A operand staged in TMEM (tcgen05.st) instead of shared memory
//   mode 4: like mode 0, but the k-steps of A beyond the last full 64-wide K-block use 32-byte-swizzle blocks
//   mode 100+: issue-rate timing (tc_issue_probe_kernel)
#include "spline_reg.cuh"
#include "tc_common.cuh"

namespace mf {
namespace {

__global__ void __launch_bounds__(128, 1)
tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ out, int K, int N,
                int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = (K + 63) / 64 * 64;                  // K-blocks are 64 wide
  const uint32_t a_bytes = 128u * (uint32_t)Kp * 2u;  // one [128 x Kp] 16-bit operand
  const uint32_t b_bytes = (uint32_t)N * (uint32_t)Kp * 2u;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;

  // zero-fill, then convert + scatter into the swizzled layout
  for (uint32_t i = tid; i < (2 * a_bytes + 2 * b_bytes) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  auto put = [&](unsigned char* hi, unsigned char* lo, int R, int r, int k, float v) {
    const uint32_t off = tc::sw128_offset(r, k, R);
    if (mode == 1) {
      *reinterpret_cast<__nv_bfloat16*>(hi + off) = __float2bfloat16_rn(v);
    } else {
      const __half h = __float2half_rn(v);
      *reinterpret_cast<__half*>(hi + off) = h;
      *reinterpret_cast<__half*>(lo + off) = __float2half_rn(v - __half2float(h));
    }
  };
  for (int i = tid; i < 128 * K; i += 128) put(a_hi, a_lo, 128, i / K, i % K, A[i]);
  for (int i = tid; i < N * K; i += 128) put(b_hi, b_lo, N, i / K, i % K, W[i]);

  unsigned char* a_tail = b_lo + b_bytes;  // mode 4: [tail k-step][plane][128 x 32 B]
  const int full_kb = K / 64, tail_steps = (K % 64) / 16;
  if (mode == 4) {
    for (int i = tid; i < 128 * tail_steps * 16; i += 128) {
      const int r = i / (tail_steps * 16), kk = i % (tail_steps * 16);
      const int ts = kk / 16, k = kk % 16;
      const float v = A[(size_t)r * K + full_kb * 64 + kk];
      const __half h = __float2half_rn(v);
      unsigned char* blk = a_tail + (size_t)ts * 2 * 4096;
      *reinterpret_cast<__half*>(blk + tc::sw32_offset(r, k)) = h;
      *reinterpret_cast<__half*>(blk + 4096 + tc::sw32_offset(r, k)) = __float2half_rn(v - __half2float(h));
    }
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;

  if (mode == 3) {
    // A planes into TMEM: hi at columns [256, 256 + K/2), lo at [384, 384 + K/2); thread = row
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t hi[8], lo[8];
      for (int j = 0; j < 8; ++j) tc::split_f16(A[(size_t)row * K + k0 + 2 * j], A[(size_t)row * K + k0 + 2 * j + 1], hi[j], lo[j]);
      tc::tmem_st8(tmem + lane_off + 256 + k0 / 2, hi);
      tc::tmem_st8(tmem + lane_off + 384 + k0 / 2, lo);
    }
    tc::tmem_st_wait();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
  }
  if (tid == 0 && mode == 3) {
    const uint32_t idesc = tc::make_idesc_f16(0, 128, N);
    const int ksteps = K / 16;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
      const uint64_t dbh = tc::make_smem_desc(smem_u32(b_hi) + kb * (uint32_t)N * 128u + kin);
      const uint64_t dbl = tc::make_smem_desc(smem_u32(b_lo) + kb * (uint32_t)N * 128u + kin);
      tc::mma_f16_ts(tmem, tmem + 256 + ks * 8, dbh, idesc, ks ? 1u : 0u);
      tc::mma_f16_ts(tmem, tmem + 256 + ks * 8, dbl, idesc, 1u);
      tc::mma_f16_ts(tmem, tmem + 384 + ks * 8, dbh, idesc, 1u);
    }
    tc::mma_commit(&bar);
  }
  if (tid == 0 && mode == 4) {
    const uint32_t idesc = tc::make_idesc_f16(0, 128, N);
    const int ksteps = K / 16;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
      uint64_t dah, dal;
      if ((int)kb < full_kb) {
        dah = tc::make_smem_desc(smem_u32(a_hi) + kb * 128u * 128u + kin);
        dal = tc::make_smem_desc(smem_u32(a_lo) + kb * 128u * 128u + kin);
      } else {
        const int ts = ks - full_kb * 4;
        dah = tc::make_smem_desc_sw32(smem_u32(a_tail) + ts * 8192u);
        dal = tc::make_smem_desc_sw32(smem_u32(a_tail) + ts * 8192u + 4096u);
      }
      const uint64_t dbh = tc::make_smem_desc(smem_u32(b_hi) + kb * (uint32_t)N * 128u + kin);
      const uint64_t dbl = tc::make_smem_desc(smem_u32(b_lo) + kb * (uint32_t)N * 128u + kin);
      tc::mma_f16_ss(tmem, dah, dbh, idesc, ks ? 1u : 0u);
      tc::mma_f16_ss(tmem, dah, dbl, idesc, 1u);
      tc::mma_f16_ss(tmem, dal, dbh, idesc, 1u);
    }
    tc::mma_commit(&bar);
  }
  if (tid == 0 && mode != 3 && mode != 4) {
    const uint32_t idesc = tc::make_idesc_f16(mode == 1 ? 1 : 0, 128, N);
    const int ksteps = (K + 15) / 16;
    uint32_t acc = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t kb = ks >> 2, kin = (ks & 3) * 32;
      const uint64_t dah = tc::make_smem_desc(smem_u32(a_hi) + kb * 128u * 128u + kin);
      const uint64_t dal = tc::make_smem_desc(smem_u32(a_lo) + kb * 128u * 128u + kin);
      const uint64_t dbh = tc::make_smem_desc(smem_u32(b_hi) + kb * (uint32_t)N * 128u + kin);
      const uint64_t dbl = tc::make_smem_desc(smem_u32(b_lo) + kb * (uint32_t)N * 128u + kin);
      tc::mma_f16_ss(tmem, dah, dbh, idesc, acc);
      acc = 1;
      if (mode == 0) {
        tc::mma_f16_ss(tmem, dah, dbl, idesc, 1);
        tc::mma_f16_ss(tmem, dal, dbh, idesc, 1);
      }
    }
    tc::mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc::fence_after_thread_sync();
  // read back: thread = row (TMEM lane)
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(size_t)row * N + c0 + j] = v[j];
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// Issue-rate probe: `issuers` warps each push `count` MMAs (M=128, N=n, K=16) at their own accumulator, operands
// either both in shared memory (ts = 0) or A in TMEM (ts = 1).  out[2w] = cycles until warp w has issued everything,
// out[2w+1] = cycles until its commit barrier fires.
__global__ void __launch_bounds__(128, 1) tc_issue_probe_kernel(float* __restrict__ out, int count, int n, int ts, int issuers, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // variant 3: random fp16 operands in (0.25, 2) with random signs instead of zeros (data-dependent MMA rate?)
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) {
    uint32_t h = (uint32_t)i * 2654435761u;
    h ^= h >> 15;
    reinterpret_cast<uint32_t*>(smem)[i] = (variant == 3) ? ((h & 0x83ff83ffu) | 0x34003400u | ((h >> 3) & 0x08000800u)) : 0u;
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&bar[i], 1);
    for (int i = 2; i < 4; ++i) mbar_init(&bar[i], 1000000);
    fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;
  if (variant == 3) {  // random A operand planes in TMEM (columns 128..191 of both accumulators' slots)
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    for (int slot = 0; slot < 2; ++slot)
      for (int c = 0; c < 64; c += 8) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t h = (uint32_t)(tid * 64 + c + i + slot * 7919) * 2246822519u;
          h ^= h >> 13;
          r[i] = (h & 0x83ff83ffu) | 0x34003400u | ((h >> 3) & 0x08000800u);
        }
        tc::tmem_st8(tmem + lane_off + slot * 256 + 128 + c, r);
      }
    tc::tmem_st_wait();
    tc::fence_before_thread_sync();
    __syncthreads();
    tc::fence_after_thread_sync();
  }
  if (warp < issuers) {
    const uint32_t idesc = tc::make_idesc_f16(0, 128, n);
    const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 16384);
    const uint32_t d_tmem = tmem + warp * 256, a_tmem = d_tmem + 128;
    const long long t0 = clock64();
    if (variant == 2) {  // like variant 1, with a tcgen05.commit to a spare barrier after every 12 MMAs
      if (tc::elect_one()) {
        uint64_t db[4], da[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { db[ks] = tc::make_smem_desc(b_s + ks * 32u); da[ks] = tc::make_smem_desc(a_s + ks * 32u); }
        for (int i = 0; i < count; i += 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ts) tc::mma_f16_ts(d_tmem, a_tmem + ks * 8u, db[ks], idesc, 1u);
            else tc::mma_f16_ss(d_tmem, da[ks], db[ks], idesc, 1u);
          }
          if ((i + 4) % 12 == 0) tc::mma_commit(&bar[2 + warp]);
        }
      }
      __syncwarp();
    } else
    if (variant == 1 || variant == 3) {  // the whole issue loop inside one elected thread, four k-steps unrolled
      if (tc::elect_one()) {
        uint64_t db[4], da[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { db[ks] = tc::make_smem_desc(b_s + ks * 32u); da[ks] = tc::make_smem_desc(a_s + ks * 32u); }
        for (int i = 0; i < count; i += 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ts) tc::mma_f16_ts(d_tmem, a_tmem + ks * 8u, db[ks], idesc, 1u);
            else tc::mma_f16_ss(d_tmem, da[ks], db[ks], idesc, 1u);
          }
        }
      }
      __syncwarp();
    } else
    for (int i = 0; i < count; ++i) {
      const uint32_t ks = i & 3;
      const uint64_t db = tc::make_smem_desc(b_s + ks * 32u);
      if (ts) {
        if (tc::elect_one()) tc::mma_f16_ts(d_tmem, a_tmem + ks * 8u, db, idesc, 1u);
      } else {
        const uint64_t da = tc::make_smem_desc(a_s + ks * 32u);
        if (tc::elect_one()) tc::mma_f16_ss(d_tmem, da, db, idesc, 1u);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::mma_commit(&bar[warp]);
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&bar[warp], 0);
    const long long t2 = clock64();
    if ((tid & 31) == 0 && blockIdx.x == 0) { out[2 * warp] = (float)(t1 - t0); out[2 * warp + 1] = (float)(t2 - t0); }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// Spline throughput probe: every thread runs `iters` register splines (K = 16, masked-sum form of spline_reg.cuh)
// back to back on one SM; out[0] = cycles per spline seen by warp 0.  variant 0: forward, 1: inverse.
__global__ void spline_probe_kernel(float* __restrict__ out, int iters, int variant, float seed) {
  const RqsScaled cf = make_rqs_scaled(16, MF_UNIV_RQS, 1.0, 1e-3);
  const float tf = seed * (float)(threadIdx.x % 29) * 0.01f;
#define MF_PB(j) (tf + (float)(((j) * 13) % 31 - 15) * 0.1f)
  float v = 0.01f * (threadIdx.x % 64) - 0.3f, acc = 0.f;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float lw[16], lh[16], ld[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      lw[j] = (MF_PB(j) + acc) * cf.inv2; lh[j] = (MF_PB(16 + j) - acc) * cf.inv2;
      ld[j] = (j < 15) ? (MF_PB(32 + j) + acc) * cf.inv1 : 0.f;
    }
    float o, la;
    int k;
    bool inside;
    if (variant == 1) rqs_masked_sums<2, true>(lw, lh, ld, cf, v, o, la, k, inside);
    else rqs_masked_sums<2, false>(lw, lh, ld, cf, v, o, la, k, inside);
    acc = 1e-6f * (o + la + (float)k);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (float)(t1 - t0) / (float)iters;
  if (acc == 12345.f) out[1] = acc;
}


// TMEM bandwidth probe: warps 0-3 (one per lane quarter) run `iters` passes over a 128-column fp32 accumulator like the
// hidden-layer epilogue of the flow kernel.  variant bits: 1 = tcgen05.ld of the 128 columns, 2 = tcgen05.st of the two
// 64-column operand planes, 4 = the bias / ReLU / fp16 hi-lo split arithmetic in between, 8 = warp 4 issues TS-form MMAs
// (M=128, N=128, K=16) back to back into the OTHER half of TMEM during the measurement.
// out[w] = cycles per pass seen by warp w, out[8] = cycles per MMA of warp 4.
__global__ void __launch_bounds__(160, 1) tmem_bw_probe_kernel(float* __restrict__ out, int iters, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32768 / 4; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0x34003400u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); stop_flag = 0; }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp < 4) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c00 = 0; c00 < 128; c00 += 32) {
        float vv[32];
        if (variant & 1) {
          tc::tmem_ld16(taddr + c00, vv);
          tc::tmem_ld16(taddr + c00 + 16, vv + 16);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) vv[i] = acc + (float)(i + c00);
        }
        uint32_t hi[16], lo[16];
        if (variant & 4) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = fmaxf(vv[2 * i] + 0.25f, 0.f), b = fmaxf(vv[2 * i + 1] - 0.125f, 0.f);
            tc::split_f16(a, b, hi[i], lo[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { hi[i] = __float_as_uint(vv[2 * i]); lo[i] = __float_as_uint(vv[2 * i + 1]); }
        }
        if (variant & 2) {
          tc::tmem_st8(taddr + 128 + c00 / 2, hi);
          tc::tmem_st8(taddr + 128 + c00 / 2 + 8, hi + 8);
          tc::tmem_st8(taddr + 192 + c00 / 2, lo);
          tc::tmem_st8(taddr + 192 + c00 / 2 + 8, lo + 8);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += __uint_as_float(hi[i] ^ lo[i]) * 1e-30f;
        }
      }
      if (variant & 2) tc::tmem_st_wait();
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = (float)(t1 - t0) / (float)iters;
    if (acc == 12345.f) out[15] = acc;
    __syncwarp();
    if (warp == 0 && lane == 0) stop_flag = 1;
  } else if (variant & 8) {
    const uint32_t idesc = tc::make_idesc_f16(0, 128, 128);
    const uint32_t b_s = smem_u32(smem);
    const uint32_t d_tmem = tmem + 256, a_tmem = d_tmem + 128;
    long long n_mma = 0;
    const long long t0 = clock64();
    while (!stop_flag) {
      if (tc::elect_one()) {
        const uint64_t db = tc::make_smem_desc(b_s);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          tc::mma_f16_ts(d_tmem, a_tmem + (ks & 7) * 8u, db + (uint64_t)((ks & 3) * 2), idesc, 1u);
          tc::mma_f16_ts(d_tmem, a_tmem + (ks & 7) * 8u, db + (uint64_t)((ks & 3) * 2), idesc, 1u);
          tc::mma_f16_ts(d_tmem, a_tmem + 64u + (ks & 7) * 8u, db + (uint64_t)((ks & 3) * 2), idesc, 1u);
        }
      }
      __syncwarp();
      n_mma += 24;
    }
    if (tc::elect_one()) tc::mma_commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (lane == 0) { out[8] = (float)(t1 - t0) / (float)n_mma; out[9] = (float)n_mma; }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// Legacy warp-level MMA rate probe: every warp issues `iters` x 8 independent mma.sync.m16n8k8 TF32 (fp32 accumulate).
// out[0] = cycles per MMA per scheduler.  f16 = 1: m16n8k16 fp16 instead.
__global__ void mma_sync_probe_kernel(float* __restrict__ out, int iters, int f16) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u}, b[2] = {threadIdx.x * 5u, 13u};
  const long long t0 = clock64();
  if (f16) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (threadIdx.x == 0) out[0] = (float)(t1 - t0) / (float)(iters * 8) / (float)(blockDim.x / 128);
  if (s == 12345.f) out[1] = s;
}

}  // namespace
}  // namespace mf

using namespace mf;

// test-only entry point of libmemflow_b200_probe.so (NOT part of the product ABI, not declared in include/memflow_b200.h)
extern "C" int mf_tc_probe_gemm(const float* d_a, const float* d_w, float* d_out, int32_t k, int32_t n, int32_t mode,
                                void* stream) {
  MF_REQUIRE(d_a && d_w && d_out, "mf_tc_probe_gemm: null argument");
  MF_REQUIRE(k >= 16 && k % 16 == 0 && k <= 256, "mf_tc_probe_gemm: K=%d must be a multiple of 16 in 16..256", k);
  MF_REQUIRE(n >= 16 && n % 16 == 0 && n <= 256, "mf_tc_probe_gemm: N=%d must be a multiple of 16 in 16..256", n);
  int all_sms = 0;
  if (mode >= 1000) { all_sms = 1; mode -= 1000; }
  if (mode >= 400) {  // TMEM bandwidth: mode = 400 + variant bits (1 ld, 2 st, 4 split arithmetic, 8 concurrent MMAs); k = passes
    MF_REQUIRE(d_out && mode - 400 < 16 && k >= 1, "mf_tc_probe_gemm: bad TMEM probe mode");
    MF_CUDA_CHECK(cudaFuncSetAttribute(tmem_bw_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    tmem_bw_probe_kernel<<<all_sms ? 148 : 1, 160, 32768, static_cast<cudaStream_t>(stream)>>>(d_out, k, mode - 400);
    return post_launch("tmem_bw_probe_kernel");
  }
  if (mode >= 300) {  // legacy mma.sync rate: mode = 300 + warps per scheduler; k = iterations
    const int wps = (mode - 300) % 10, f16 = (mode - 300) / 10;
    MF_REQUIRE(d_out && wps >= 1 && wps <= 8 && f16 <= 1, "mf_tc_probe_gemm: bad mma.sync probe mode");
    mma_sync_probe_kernel<<<1, 128 * wps, 0, static_cast<cudaStream_t>(stream)>>>(d_out, k, f16);
    return post_launch("mma_sync_probe_kernel");
  }
  if (mode >= 200) {  // spline throughput: mode = 200 + 10 * variant + warps per scheduler (1..4); k = iterations
    const int variant = (mode - 200) / 10, wps = (mode - 200) % 10;
    MF_REQUIRE(d_out && wps >= 1 && wps <= 8 && variant <= 1, "mf_tc_probe_gemm: bad spline probe mode");
    spline_probe_kernel<<<1, 128 * wps, 0, static_cast<cudaStream_t>(stream)>>>(d_out, k, variant, 1.0f);
    return post_launch("spline_probe_kernel");
  }
  if (mode >= 100) {  // timing: mode = 100 + 10 * (A in TMEM) + issuers; k = MMA count per issuer
    const int grid = all_sms ? 148 : 1;  // mode + 1000: the same probe on every SM at once (chip-level effects)
    const int variant = (mode - 100) / 20, ts = ((mode - 100) / 10) % 2, issuers = (mode - 100) % 10;
    MF_REQUIRE(d_out && issuers >= 1 && issuers <= 2 && ts <= 1 && variant <= 3 && n <= 128 && n % 16 == 0, "mf_tc_probe_gemm: bad timing mode");
    const size_t sm = 16384 + 32768;
    MF_CUDA_CHECK(cudaFuncSetAttribute(tc_issue_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    tc_issue_probe_kernel<<<grid, 128, sm, static_cast<cudaStream_t>(stream)>>>(d_out, k, n, ts, issuers, variant);
    return post_launch("tc_issue_probe_kernel");
  }
  MF_REQUIRE(mode >= 0 && mode <= 4, "mf_tc_probe_gemm: bad mode");
  const int kp = (k + 63) / 64 * 64;
  const size_t smem = (size_t)2 * 128 * kp * 2 + (size_t)2 * n * kp * 2 + 3 * 8192 + 1024;
  MF_REQUIRE(smem <= 227 * 1024, "mf_tc_probe_gemm: operands do not fit shared memory");
  MF_CUDA_CHECK(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_a, d_w, d_out, k, n, mode);
  return post_launch("tc_probe_kernel");
}
