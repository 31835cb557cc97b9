// rqs.cu -- standalone monotonic rational-quadratic spline transform on materialised parameters.
//
// zuko's MonotonicRQSTransform(widths, heights, derivatives, bound, slope) (a17 in SURVEY.md section 8):
// forward (call_and_ladj) or inverse of N independent scalars, each with its own 3K-1 unconstrained
// parameters.  This is the HBM-bound form of the spline (the fused flow kernels never materialise the
// parameters): 4(3K-1) + 4 bytes in, 8 bytes out per element.
//
// Layout / schedule: parameters are packed [N, 3K-1] = (K widths | K heights | K-1 derivatives), so a CTA's
// tile of 256 consecutive elements is one contiguous range that is brought into shared memory with a single
// 1-D bulk async copy (TMA engine).  One thread then evaluates one element out of shared memory: the row
// stride 3K-1 is odd, so the 32 lanes of a warp hit 32 different banks.  A thread-per-element schedule
// needs ~17 warp-instructions per element (K = 16); a warp-per-element schedule built on shuffles (log-step
// max / sum / scan, ballot bin search) needs ~150, which would make this streaming kernel issue-bound at
// ~25 % of the HBM roofline, so shuffles are not used here.
#include "flow_common.cuh"
#include "spline_reg.cuh"

namespace mf {
namespace {

constexpr int kRqsRows = 256;

// fp32 fast path: all logits of the element in registers (K <= 8 * NCH)
template <int NCH>
__device__ __forceinline__ void rqs_row_reg(const float* p, const SplineCfg<float>& cfg, bool inverse, float v,
                                            float& out, float& la, int& bin, bool& ins) {
  constexpr int KP = 8 * NCH;
  const int K = cfg.K;
  float lw[KP], lh[KP], ld[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    // clamped indices instead of guarded loads (a conditional load compiles to one branch per element); the
    // duplicated entries beyond K are masked out inside the spline
    const int jw = min(j, K - 1), jd = max(min(j, K - 2), 0);
    lw[j] = p[jw];
    lh[j] = p[K + jw];
    ld[j] = p[2 * K + jd];
  }
  rqs_eval_reg<NCH>(lw, lh, ld, cfg, inverse, v, out, la, bin, ins);
}

template <typename T>
__device__ __forceinline__ void rqs_row(T* p, const SplineCfg<T>& cfg, bool inverse, T v, T& out, T& la, int& bin,
                                        bool& ins) {
  rqs_univariate<T>(p, 1, cfg, inverse, v, out, la, bin, ins);
}
template <>
__device__ __forceinline__ void rqs_row<float>(float* p, const SplineCfg<float>& cfg, bool inverse, float v, float& out,
                                               float& la, int& bin, bool& ins) {
  const int K = cfg.K;
  if (K <= 8) rqs_row_reg<1>(p, cfg, inverse, v, out, la, bin, ins);
  else if (K <= 16) rqs_row_reg<2>(p, cfg, inverse, v, out, la, bin, ins);
  else if (K <= 24) rqs_row_reg<3>(p, cfg, inverse, v, out, la, bin, ins);
  else if (K <= 32) rqs_row_reg<4>(p, cfg, inverse, v, out, la, bin, ins);
  else rqs_univariate<float>(p, 1, cfg, inverse, v, out, la, bin, ins);
}

template <typename T>
__global__ void __launch_bounds__(kRqsRows, 2) rqs_kernel(const T* __restrict__ g_par, const T* __restrict__ g_x, long long n, int K, int univariate,
                           double bound, double slope, int inverse, T* __restrict__ g_y, T* __restrict__ g_ladj,
                           int8_t* __restrict__ g_bins, uint8_t* __restrict__ g_inside, int use_tma) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* s_par = reinterpret_cast<T*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  const int R = blockDim.x;
  const int P = 3 * K - 1;
  const int tid = threadIdx.x;
  SplineCfg<T> cfg;
  cfg.K = K; cfg.univariate = univariate; cfg.bound = (T)bound;
  cfg.inv_ls2 = slope > 0.0 ? (T)(2.0 / log(slope)) : T(0);
  cfg.inv_ls1 = slope > 0.0 ? (T)(1.0 / log(slope)) : T(0);
  const long long first = (long long)blockIdx.x * R;
  const int rows = (int)min((long long)R, n - first);
  const uint32_t tile_bytes = (uint32_t)((size_t)R * P * sizeof(T));
  const bool tma = use_tma && rows == R && (tile_bytes % 16 == 0);
  if (tma) {
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar, tile_bytes);
      bulk_g2s(s_par, g_par + first * P, tile_bytes, &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    for (int i = tid; i < rows * P; i += R) s_par[i] = g_par[first * P + i];
    __syncthreads();
  }
  if (tid < rows) {
    T out, la;
    int bin;
    bool ins;
    rqs_row<T>(s_par + (size_t)tid * P, cfg, inverse != 0, g_x[first + tid], out, la, bin, ins);
    g_y[first + tid] = out;
    if (g_ladj) g_ladj[first + tid] = la;
    if (g_bins) g_bins[first + tid] = (int8_t)bin;
    if (g_inside) g_inside[first + tid] = ins ? 1 : 0;
  }
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" int mf_rqs_transform(const void* d_params, const void* d_x, int64_t n, int32_t bins, int32_t univariate,
                                double bound, double slope, int32_t inverse, void* d_y, void* d_ladj, int8_t* d_bins,
                                uint8_t* d_inside, int dtype, void* stream) {
  MF_REQUIRE(n >= 0, "mf_rqs_transform: negative count");
  MF_REQUIRE((d_params && d_x && d_y) || n == 0, "mf_rqs_transform: null argument");
  MF_REQUIRE(bins >= 2 && bins <= 127, "mf_rqs_transform: bins=%d outside 2..127", bins);
  MF_REQUIRE(bound > 0.0, "mf_rqs_transform: bound must be positive");
  MF_REQUIRE(dtype == MF_F32 || dtype == MF_F64, "mf_rqs_transform: unsupported dtype %d", dtype);
  if (n == 0) return MF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = 3 * bins - 1;
  const size_t esz = dtype == MF_F32 ? 4 : 8;
  int rows = dtype == MF_F32 ? 128 : kRqsRows;  // fp32: 128 rows x 128 registers -> four CTAs per SM (measured best)
  while (rows > 32 && (size_t)rows * P * esz > 56 * 1024) rows >>= 1;  // <= 56 KB per CTA: four CTAs per SM
  const size_t smem = (size_t)rows * P * esz;
  MF_REQUIRE(smem <= 220 * 1024, "mf_rqs_transform: bins too large for shared memory");
  const unsigned grid = (unsigned)((n + rows - 1) / rows);
  const int use_tma = tma_enabled() && (reinterpret_cast<uintptr_t>(d_params) & 15u) == 0;
  if (dtype == MF_F32) {
    MF_CUDA_CHECK(cudaFuncSetAttribute(rqs_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rqs_kernel<float><<<grid, rows, smem, st>>>((const float*)d_params, (const float*)d_x, n, bins, univariate, bound,
                                               slope, inverse, (float*)d_y, (float*)d_ladj, d_bins, d_inside, use_tma);
  } else {
    MF_CUDA_CHECK(cudaFuncSetAttribute(rqs_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rqs_kernel<double><<<grid, rows, smem, st>>>((const double*)d_params, (const double*)d_x, n, bins, univariate, bound,
                                                slope, inverse, (double*)d_y, (double*)d_ladj, d_bins, d_inside, use_tma);
  }
  return post_launch("rqs_kernel");
}
