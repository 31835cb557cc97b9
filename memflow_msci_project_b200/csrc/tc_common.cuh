// tc_common.cuh -- tcgen05 / TMEM building blocks (sm_100a) for the conditioner GEMMs.
//
// Operand layout in shared memory ("K-major, 128-byte swizzle", the canonical UMMA layout): a matrix of
// R rows x Kdim 16-bit elements is split into K-blocks of 64 elements; block kb lives at byte offset
// kb * R * 128; inside a block row r occupies the 128 bytes at r * 128 and its eight 16-byte chunks are
// permuted: chunk c is stored at position c ^ (r & 7).  The block base must be 1024-byte aligned.
// One tcgen05.mma (kind::f16) consumes K = 16 elements = 32 bytes of every row: successive k-steps advance
// the descriptor start address by 32 bytes inside the block, then jump to the next block.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace mf {
namespace tc {

constexpr int kBlockK = 64;          // elements per swizzled K-block (128 bytes of fp16/bf16)
constexpr uint32_t kRowBytes = 128;  // bytes per row inside a K-block

// byte offset of element (r, k) of an [R x Kdim] 16-bit operand
__host__ __device__ __forceinline__ uint32_t sw128_offset(int r, int k, int R) {
  const int kb = k >> 6, kk = k & 63;
  const int chunk = (kk >> 3) ^ (r & 7);
  return (uint32_t)kb * (uint32_t)R * kRowBytes + (uint32_t)r * kRowBytes + (uint32_t)chunk * 16u +
         (uint32_t)(kk & 7) * 2u;
}

// 32-byte-swizzle K-major layout of ONE k-step (16 elements = 32 bytes per row): row r at r * 32, its two 16-byte
// chunks swapped when bit 2 of r is set (Swizzle<1,4,3> on the byte address); 8-row groups 256 bytes apart.
// Used for the tail k-steps of the layer-0 operand, where a full 128-byte K-block would waste shared memory.
__host__ __device__ __forceinline__ uint32_t sw32_offset(int r, int k) {
  return (uint32_t)r * 32u + (uint32_t)(((k >> 3) ^ ((r >> 2) & 1)) * 16) + (uint32_t)(k & 7) * 2u;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;  // stride byte offset: 8 rows * 32 bytes
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;           // layout type SWIZZLE_32B
  return d;
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major), [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100), [46,48)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B, [61,64)
  return d;
}

// instruction descriptor, kind::f16: A/B format (0 = fp16, 1 = bf16), fp32 accumulate, K-major A and B
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// true in exactly one lane of a fully converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand from TMEM (lane = row, 32-bit column c holds K elements 2c, 2c+1 as a packed 16-bit pair),
// B from shared memory
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// registers -> TMEM: this thread's lane, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMEM -> registers: this thread's lane (32 * (warp % 4) + lane), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 -> (hi, lo) fp16 pair with hi + lo = v to ~2^-22 relative
__device__ __forceinline__ void split_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
#endif  // __CUDACC__

}  // namespace tc
}  // namespace mf
