// mmd.cu -- empirical maximum mean discrepancy between two samples, value and gradient in one pass.
//
// Replaces memflow/unfolding_flow/mmd_loss.py:4-41 (`MMD(x, y, kernel, device, dtype)`), which the sampling-loss
// training steps call per particle and once on the concatenation (scripts/run_model_labframe_sampling.py:114-127).
// The reference materialises three B x B Gram matrices with torch.mm, three B x B distance matrices and one
// B x B accumulator per bandwidth; here the 2B points z = (x | y) are treated as one signed pair sum
//
//     MMD = 1/B^2 * sum_{i,j < 2B} s_ij * k(|z_i - z_j|^2),   s_ij = +1 inside a sample, -1 across the samples
//
// (the cross block appears twice, which is the reference's -2 * XY), nothing of size B^2 is ever stored, and the
// squared distance is the direct sum of squared differences instead of r_i + r_j - 2 <z_i, z_j>.
//
//     multiscale: k(d) = sum_a a^2 / (a^2 + d),  a in {0.1, 0.2, 0.5, 0.9, 1.3}          (mmd_loss.py:25-31)
//     rbf:        k(d) = sum_a exp(-d / (2 a)),  a in {10, 15, 20, 50}                    (mmd_loss.py:33-39)
//
// Gradient (what torch autograd derives in the reference's loss.backward()): dMMD/dz_i = 4/B^2 * sum_j s_ij k'(d_ij)(z_i - z_j).
//
// Work split: one thread owns one point z_i in registers; a CTA of 128 threads walks its chunk of the j range in tiles
// of 128 points staged in shared memory (every thread reads the same z_j: a broadcast, no bank conflicts).  The j range
// is cut into `splits` chunks so that the grid covers the 148 SMs at small B; each (chunk, i) writes its partial sums
// to the caller's workspace and a second one-CTA kernel adds them in a fixed order -- no atomics, results do not depend
// on scheduling.  Compute-bound on the CUDA cores (about 2 D + 20 flops per pair); HBM traffic is negligible.
#include "common.cuh"

namespace mf {
namespace {

constexpr int MMD_THREADS = 128;

// fp32: approximate division / exponential (MUFU, about 2 ulp); fp64: IEEE
__device__ __forceinline__ float mmd_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double mmd_div(double a, double b) { return a / b; }
__device__ __forceinline__ float mmd_exp(float v) { return __expf(v); }
__device__ __forceinline__ double mmd_exp(double v) { return exp(v); }

template <typename T, int KERNEL>
__device__ __forceinline__ void mmd_kernel_value(T d, T& k, T& dk) {
  if (KERNEL == MF_MMD_MULTISCALE) {
    const T a2[5] = {T(0.1) * T(0.1), T(0.2) * T(0.2), T(0.5) * T(0.5), T(0.9) * T(0.9), T(1.3) * T(1.3)};
    const T inv_a2[5] = {T(1) / a2[0], T(1) / a2[1], T(1) / a2[2], T(1) / a2[3], T(1) / a2[4]};  // folded by the compiler
    k = T(0);
    dk = T(0);
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const T q = mmd_div(a2[b], a2[b] + d);  // a^2 (a^2 + d)^-1
      k += q;
      dk -= q * q * inv_a2[b];          // d/dd = -a^2 / (a^2 + d)^2
    }
  } else {
    const T c[4] = {T(-0.5) / T(10), T(-0.5) / T(15), T(-0.5) / T(20), T(-0.5) / T(50)};  // -1 / (2 a)
    k = T(0);
    dk = T(0);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const T e = mmd_exp(c[b] * d);
      k += e;
      dk += c[b] * e;
    }
  }
}

// part_loss [splits, 2n], part_grad [splits, 2n, d] (doubles).
template <typename T, int DP, int KERNEL, bool GRAD>
__global__ void __launch_bounds__(MMD_THREADS)
mmd_pair_kernel(const T* __restrict__ x, const T* __restrict__ y, int n, int d, int chunk,
                double* __restrict__ part_loss, double* __restrict__ part_grad) {
  __shared__ __align__(16) T tile[MMD_THREADS * DP];
  const int n2 = 2 * n;
  const int i = blockIdx.x * MMD_THREADS + threadIdx.x;
  const bool live = i < n2;
  const bool i_in_x = i < n;
  T zi[DP];
  {
    const T* src = live ? (i_in_x ? x + (size_t)i * d : y + (size_t)(i - n) * d) : nullptr;
#pragma unroll
    for (int c = 0; c < DP; ++c) zi[c] = (live && c < d) ? src[c] : T(0);
  }
  double loss = 0.0;
  double grad[GRAD ? DP : 1];
#pragma unroll
  for (int c = 0; c < (GRAD ? DP : 1); ++c) grad[c] = 0.0;

  const int j_begin = blockIdx.y * chunk;
  const int j_end = min(n2, j_begin + chunk);
  for (int j0 = j_begin; j0 < j_end; j0 += MMD_THREADS) {
    const int cnt = min(MMD_THREADS, j_end - j0);
    __syncthreads();
    {  // stage z_j, zero-padded to DP columns
      const int j = j0 + threadIdx.x;
      if (threadIdx.x < cnt) {
        const T* src = j < n ? x + (size_t)j * d : y + (size_t)(j - n) * d;
#pragma unroll
        for (int c = 0; c < DP; ++c) tile[threadIdx.x * DP + c] = c < d ? src[c] : T(0);
      }
    }
    __syncthreads();
    if (!live) continue;
    T tl = T(0);
    T tg[GRAD ? DP : 1];
#pragma unroll
    for (int c = 0; c < (GRAD ? DP : 1); ++c) tg[c] = T(0);
    for (int t = 0; t < cnt; ++t) {
      T diff[DP];
      T dist = T(0);
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        diff[c] = zi[c] - tile[t * DP + c];
        dist += diff[c] * diff[c];
      }
      T k, dk;
      mmd_kernel_value<T, KERNEL>(dist, k, dk);
      const bool same = ((j0 + t) < n) == i_in_x;
      tl += same ? k : -k;
      if (GRAD) {
        const T w = same ? dk : -dk;
#pragma unroll
        for (int c = 0; c < DP; ++c) tg[c] += w * diff[c];
      }
    }
    loss += (double)tl;
    if (GRAD) {
#pragma unroll
      for (int c = 0; c < DP; ++c) grad[c] += (double)tg[c];
    }
  }
  if (!live) return;
  part_loss[(size_t)blockIdx.y * n2 + i] = loss;
  if (GRAD) {
    double* g = part_grad + ((size_t)blockIdx.y * n2 + i) * d;
#pragma unroll
    for (int c = 0; c < DP; ++c)
      if (c < d) g[c] = grad[c];
  }
}

// Fixed-order sums of the partials.  CTA 0: out_loss[0] = sum / n^2; CTAs 1..: grad_x / grad_y = 4 / n^2 * sum over the
// chunks, one element per thread.
template <typename T>
__global__ void __launch_bounds__(1024)
mmd_finish_kernel(const double* __restrict__ part_loss, const double* __restrict__ part_grad, int n, int d, int splits,
                  T* __restrict__ out_loss, T* __restrict__ grad_x, T* __restrict__ grad_y) {
  __shared__ double red[1024];
  const int n2 = 2 * n;
  const double inv = 1.0 / ((double)n * (double)n);
  if (blockIdx.x > 0) {
    const size_t total = (size_t)n2 * d;
    const size_t e = (size_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x;
    if (e >= total) return;
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += part_grad[(size_t)s * total + e];
    const T v = (T)(4.0 * inv * a);
    const size_t nx = (size_t)n * d;
    if (e < nx) {
      if (grad_x) grad_x[e] = v;
    } else if (grad_y) {
      grad_y[e - nx] = v;
    }
    return;
  }
  double acc = 0.0;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    double a = 0.0;
    for (int s = 0; s < splits; ++s) a += part_loss[(size_t)s * n2 + i];
    acc += a;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out_loss[0] = (T)(red[0] * inv);
}

inline int mmd_tiles(int64_t n) { return (int)((2 * n + MMD_THREADS - 1) / MMD_THREADS); }
// j-range chunks: enough CTAs for four per SM (16 warps) at small n, whole tiles per chunk; depends on n only (reproducible sums)
inline int mmd_splits(int64_t n) {
  const int tiles = mmd_tiles(n);
  int s = (4 * 148 + tiles - 1) / tiles;
  if (s > tiles) s = tiles;
  if (s < 1) s = 1;
  const int tiles_per = (tiles + s - 1) / s;
  return (tiles + tiles_per - 1) / tiles_per;
}

template <typename T, int DP, int KERNEL>
int mmd_launch(const T* x, const T* y, int n, int d, bool want_grad, double* ws, T* out, T* gx, T* gy, cudaStream_t st) {
  const int tiles = mmd_tiles(n), splits = mmd_splits(n);
  const int chunk = ((tiles + splits - 1) / splits) * MMD_THREADS;
  double* part_loss = ws;
  double* part_grad = ws + (size_t)splits * 2 * n;
  dim3 grid(tiles, splits);
  if (want_grad)
    mmd_pair_kernel<T, DP, KERNEL, true><<<grid, MMD_THREADS, 0, st>>>(x, y, n, d, chunk, part_loss, part_grad);
  else
    mmd_pair_kernel<T, DP, KERNEL, false><<<grid, MMD_THREADS, 0, st>>>(x, y, n, d, chunk, part_loss, part_grad);
  int rc = post_launch("mmd_pair_kernel");
  if (rc != MF_OK) return rc;
  const unsigned fin = 1u + (want_grad ? (unsigned)(((size_t)2 * n * d + 1023) / 1024) : 0u);
  mmd_finish_kernel<T><<<fin, 1024, 0, st>>>(part_loss, part_grad, n, d, splits, out, want_grad ? gx : nullptr,
                                          want_grad ? gy : nullptr);
  return post_launch("mmd_finish_kernel");
}

template <typename T, int KERNEL>
int mmd_dispatch_d(const T* x, const T* y, int n, int d, bool want_grad, double* ws, T* out, T* gx, T* gy,
                   cudaStream_t st) {
  if (d <= 4) return mmd_launch<T, 4, KERNEL>(x, y, n, d, want_grad, ws, out, gx, gy, st);
  if (d <= 8) return mmd_launch<T, 8, KERNEL>(x, y, n, d, want_grad, ws, out, gx, gy, st);
  if (d <= 12) return mmd_launch<T, 12, KERNEL>(x, y, n, d, want_grad, ws, out, gx, gy, st);
  if (d <= 16) return mmd_launch<T, 16, KERNEL>(x, y, n, d, want_grad, ws, out, gx, gy, st);
  if (d <= 24) return mmd_launch<T, 24, KERNEL>(x, y, n, d, want_grad, ws, out, gx, gy, st);
  return mmd_launch<T, 32, KERNEL>(x, y, n, d, want_grad, ws, out, gx, gy, st);
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" int64_t mf_mmd_workspace_bytes(int64_t n, int32_t d, int32_t want_grad) {
  if (n < 1 || d < 1) return 0;
  return (int64_t)mmd_splits(n) * 2 * n * (1 + (want_grad ? d : 0)) * (int64_t)sizeof(double);
}

extern "C" int mf_mmd(const void* d_x, const void* d_y, int64_t n, int32_t d, int32_t kernel, void* d_workspace,
                      void* d_out, void* d_grad_x, void* d_grad_y, int dtype, void* stream) {
  MF_REQUIRE(n >= 1 && n <= (1 << 24), "mf_mmd: n = %lld outside [1, 2^24]", (long long)n);
  MF_REQUIRE(d >= 1 && d <= MF_MMD_MAX_DIM, "mf_mmd: d = %d outside [1, %d]", d, MF_MMD_MAX_DIM);
  MF_REQUIRE(kernel == MF_MMD_MULTISCALE || kernel == MF_MMD_RBF, "mf_mmd: unknown kernel %d", kernel);
  MF_REQUIRE(dtype == MF_F32 || dtype == MF_F64, "mf_mmd: unsupported dtype %d", dtype);
  MF_REQUIRE(d_x && d_y && d_out && d_workspace, "mf_mmd: null argument");
  const bool want_grad = d_grad_x || d_grad_y;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* ws = static_cast<double*>(d_workspace);
#define MF_MMD_CASE(T, K)                                                                                     \
  return mmd_dispatch_d<T, K>((const T*)d_x, (const T*)d_y, (int)n, d, want_grad, ws, (T*)d_out, (T*)d_grad_x, \
                              (T*)d_grad_y, st)
  if (dtype == MF_F32) {
    if (kernel == MF_MMD_MULTISCALE) MF_MMD_CASE(float, MF_MMD_MULTISCALE);
    MF_MMD_CASE(float, MF_MMD_RBF);
  }
  if (kernel == MF_MMD_MULTISCALE) MF_MMD_CASE(double, MF_MMD_MULTISCALE);
  MF_MMD_CASE(double, MF_MMD_RBF);
#undef MF_MMD_CASE
}
