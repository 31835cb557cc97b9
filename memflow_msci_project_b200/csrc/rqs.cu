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
#include <algorithm>
#include "flow_common.cuh"
#include "spline_reg.cuh"

namespace mf {
namespace {

constexpr int kRqsRows = 256;

// fp32 fast path: all logits of the element in registers (K <= 8 * NCH), masked-sum spline of spline_reg.cuh
template <int NCH, bool INV>
__device__ __forceinline__ void rqs_row_reg(const float* p, const RqsScaled& cf, float v, float& out, float& la, int& bin,
                                            bool& ins) {
  constexpr int KP = 8 * NCH;
  const int K = cf.K;
  float lw[KP], lh[KP], ld[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    // clamped indices instead of guarded loads (a conditional load compiles to one branch per element); the
    // duplicated width / height entries beyond K are masked out inside the spline, derivative entries from K-1 on
    // must be zero (boundary derivative exp(0) = 1)
    const int jw = min(j, K - 1), jd = max(min(j, K - 2), 0);
    lw[j] = p[jw] * cf.inv2;
    lh[j] = p[K + jw] * cf.inv2;
    ld[j] = (j < K - 1) ? p[2 * K + jd] * cf.inv1 : 0.f;
  }
  rqs_masked_sums<NCH, INV>(lw, lh, ld, cf, v, out, la, bin, ins);
}

template <typename T>
__device__ __forceinline__ void rqs_row(T* p, const SplineCfg<T>& cfg, double slope, bool inverse, T v, T& out, T& la,
                                        int& bin, bool& ins) {
  rqs_univariate<T>(p, 1, cfg, inverse, v, out, la, bin, ins);
}
template <>
__device__ __forceinline__ void rqs_row<float>(float* p, const SplineCfg<float>& cfg, double slope, bool inverse, float v,
                                               float& out, float& la, int& bin, bool& ins) {
  const int K = cfg.K;
  if (K > 40 || !(slope > 0.0)) {  // no soft clip (unbounded logits) or more bins than the register path holds
    rqs_univariate<float>(p, 1, cfg, inverse, v, out, la, bin, ins);
    return;
  }
  const RqsScaled cf = make_rqs_scaled(K, cfg.univariate, (double)cfg.bound, slope);
#define MF_RQS_CASE(NCH) \
  { if (inverse) rqs_row_reg<NCH, true>(p, cf, v, out, la, bin, ins); else rqs_row_reg<NCH, false>(p, cf, v, out, la, bin, ins); }
  if (K <= 8) MF_RQS_CASE(1)
  else if (K <= 16) MF_RQS_CASE(2)
  else if (K <= 24) MF_RQS_CASE(3)
  else if (K <= 32) MF_RQS_CASE(4)
  else MF_RQS_CASE(5)
#undef MF_RQS_CASE
}

// Persistent: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  With `nbuf` = 2 the parameter tile of the NEXT
// tile travels (one bulk async copy completing on an mbarrier) while the splines of the current tile are evaluated; with
// one tile per CTA and load -> wait -> compute in sequence the kernel was latency-bound (67.8 % of the HBM roofline at
// K = 16 with four CTAs per SM: the copy engine idles while a CTA computes).
template <typename T>
__global__ void __launch_bounds__(kRqsRows, 2) rqs_kernel(const T* __restrict__ g_par, const T* __restrict__ g_x, long long n, int K, int univariate,
                           double bound, double slope, int inverse, T* __restrict__ g_y, T* __restrict__ g_ladj,
                           int8_t* __restrict__ g_bins, uint8_t* __restrict__ g_inside, int use_tma, int nbuf) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  const int R = blockDim.x;
  const int P = 3 * K - 1;
  const int tid = threadIdx.x;
  SplineCfg<T> cfg;
  cfg.K = K; cfg.univariate = univariate; cfg.bound = (T)bound;
  cfg.inv_ls2 = slope > 0.0 ? (T)(2.0 / log(slope)) : T(0);
  cfg.inv_ls1 = slope > 0.0 ? (T)(1.0 / log(slope)) : T(0);
  const uint32_t tile_bytes = (uint32_t)((size_t)R * P * sizeof(T));
  const long long n_tiles = (n + R - 1) / R;
  const bool tma_ok = use_tma && (tile_bytes % 16 == 0);
  auto buf_of = [&](int b) { return reinterpret_cast<T*>(smem_raw + (size_t)b * tile_bytes); };
  auto is_tma_tile = [&](long long tile) { return tma_ok && tile < n_tiles && (tile + 1) * R <= n; };
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  __syncthreads();
  uint32_t ph0 = 0, ph1 = 0;  // phases of the two barriers
  long long tile = blockIdx.x;
  if (tid == 0 && is_tma_tile(tile)) {
    mbar_expect_tx(&bar[0], tile_bytes);
    bulk_g2s(buf_of(0), g_par + tile * R * P, tile_bytes, &bar[0]);
  }
  int b = 0;
  for (; tile < n_tiles; tile += gridDim.x) {
    const long long first = tile * R;
    const int rows = (int)min((long long)R, n - first);
    const long long next = tile + gridDim.x;
    const int nb = (nbuf == 2) ? (b ^ 1) : 0;
    // the next tile's parameters: requested now when there is a second buffer (its last readers passed the barrier at
    // the end of the previous iteration), otherwise after this tile's splines
    if (nbuf == 2 && tid == 0 && is_tma_tile(next)) {
      mbar_expect_tx(&bar[nb], tile_bytes);
      bulk_g2s(buf_of(nb), g_par + next * R * P, tile_bytes, &bar[nb]);
    }
    const T xv = (tid < rows) ? g_x[first + tid] : T(0);
    T* s_par = buf_of(b);
    if (is_tma_tile(tile)) {
      mbar_wait(&bar[b], b ? ph1 : ph0);
      if (b) ph1 ^= 1; else ph0 ^= 1;
    } else {
      for (int i = tid; i < rows * P; i += R) s_par[i] = g_par[first * P + i];
      __syncthreads();
    }
    if (tid < rows) {
      T out, la;
      int bin;
      bool ins;
      rqs_row<T>(s_par + (size_t)tid * P, cfg, slope, inverse != 0, xv, out, la, bin, ins);
      g_y[first + tid] = out;
      if (g_ladj) g_ladj[first + tid] = la;
      if (g_bins) g_bins[first + tid] = (int8_t)bin;
      if (g_inside) g_inside[first + tid] = ins ? 1 : 0;
    }
    __syncthreads();  // every thread is done with this buffer
    if (nbuf != 2 && tid == 0 && is_tma_tile(next)) {
      mbar_expect_tx(&bar[0], tile_bytes);
      bulk_g2s(buf_of(0), g_par + next * R * P, tile_bytes, &bar[0]);
    }
    b = nb;
  }
}


// ------------------------------------------------------------------------------------------------
// Backward of the forward spline (x -> y, ladj): given dL/dy and dL/dladj per element, the gradient w.r.t. the 3K-1
// unconstrained parameters of the element and w.r.t. x.  One thread per element on its parameter row in shared
// memory; the row is overwritten IN PLACE with the gradient row and the tile is written back as one block.
//   knots X_j = B (2 cumsum(softmax(sc(a)))_j - 1), same for Y; D_j = exp(sc(c_j));  sc(v) = v / (1 + |f v|)
//   bin k: x0 = X_{k-1}, x1 = X_k, ..., d0 = D_{k-1} (1 at the boundary), d1 = D_k;
//   z = (x-x0)/dx, q = z(1-z), s = dy/dx, den = s + (d0+d1-2s) q, num = s z^2 + d0 q, A = 2 s q + d0 (1-z)^2 + d1 z^2
//   y = y0 + dy num/den, ladj = 2 ln s + ln A - 2 ln den.
// Outside the spline interval (mask false) y = x and ladj = 0: dx = dy, zero parameter gradient, like the reference's
// torch.where formulation (zuko MonotonicRQSTransform.call_and_ladj under autograd).
template <typename T>
__device__ void rqs_backward_row(T* p, const SplineCfg<T>& cfg, T x, T gy, T gl, T& gx) {
  using M = FM<T>;
  const int K = cfg.K;
  T* pw = p;
  T* ph = p + K;
  T* pd = p + 2 * K;
  const T B = cfg.bound;
  // forward pieces: soft clip (keeping its derivative in place of the logit is not possible: two passes)
  T mw = -INFINITY, mh = -INFINITY;
  for (int j = 0; j < K; ++j) {
    mw = fmax(mw, softclip(pw[j], cfg.inv_ls2));
    mh = fmax(mh, softclip(ph[j], cfg.inv_ls2));
  }
  T sw = T(0), sh = T(0);
  for (int j = 0; j < K; ++j) {
    sw += M::exp_(softclip(pw[j], cfg.inv_ls2) - mw);
    sh += M::exp_(softclip(ph[j], cfg.inv_ls2) - mh);
  }
  // scan: bin k and its knots (cumulative softmax on both axes)
  int k = -1;
  T cw = T(0), ch = T(0), cw0 = T(0), cw1 = T(0), ch0 = T(0), ch1 = T(0);
  T prevw = T(0), prevh = T(0);
  for (int j = 0; j < K; ++j) {
    const T knot_prev = B * (T(2) * cw - T(1));
    cw += M::exp_(softclip(pw[j], cfg.inv_ls2) - mw) / sw;
    ch += M::exp_(softclip(ph[j], cfg.inv_ls2) - mh) / sh;
    if (knot_prev < x) { k = j; cw0 = prevw; cw1 = cw; ch0 = prevh; ch1 = ch; }
    prevw = cw; prevh = ch;
  }
  const bool inside = (k >= 0) && !(B * (T(2) * cw - T(1)) < x);
  if (!inside) {
    for (int j = 0; j < 3 * K - 1; ++j) p[j] = T(0);
    gx = gy;
    return;
  }
  const T x0 = B * (T(2) * cw0 - T(1)), x1 = B * (T(2) * cw1 - T(1));
  const T y0 = B * (T(2) * ch0 - T(1)), y1 = B * (T(2) * ch1 - T(1));
  const T d0 = (k == 0) ? T(1) : M::exp_(softclip(pd[k - 1], cfg.inv_ls1));
  const T d1 = (k == K - 1) ? T(1) : M::exp_(softclip(pd[k], cfg.inv_ls1));
  const T dx = x1 - x0, dy = y1 - y0;
  const T s = dy / dx, z = (x - x0) / dx, omz = T(1) - z, q = z * omz;
  const T den = s + (d0 + d1 - T(2) * s) * q;
  const T num = s * z * z + d0 * q;
  const T A = T(2) * s * q + d0 * omz * omz + d1 * z * z;
  // reverse mode through (y, ladj)
  const T g_num = gy * dy / den;
  const T g_den = -gy * dy * num / (den * den) - T(2) * gl / den;
  const T g_A = gl / A;
  T g_dy = gy * num / den;
  T g_s = T(2) * gl / s + g_den * (T(1) - T(2) * q) + g_num * z * z + g_A * T(2) * q;
  const T g_d0 = (g_den + g_num) * q + g_A * omz * omz;
  const T g_d1 = g_den * q + g_A * z * z;
  const T g_q = g_den * (d0 + d1 - T(2) * s) + g_num * d0 + g_A * T(2) * s;
  const T g_z = g_num * T(2) * s * z + g_A * (T(2) * d1 * z - T(2) * d0 * omz) + g_q * (T(1) - T(2) * z);
  g_dy += g_s / dx;
  T g_dx = -g_s * s / dx - g_z * z / dx;
  gx = g_z / dx;
  const T g_x0 = -g_z / dx - g_dx, g_x1 = g_dx;
  const T g_y0 = gy - g_dy, g_y1 = g_dy;
  // knots -> cumulative softmax -> softmax -> soft-clipped logits -> logits
  //   dL/dW_i = 2B (g_x0 [i <= k-1] + g_x1 [i <= k]);  sum_j W_j dL/dW_j = 2B (g_x0 cumW_{k-1} + g_x1 cumW_k)
  const T tb = T(2) * B;
  const T Sw = tb * (g_x0 * cw0 + g_x1 * cw1), Sh = tb * (g_y0 * ch0 + g_y1 * ch1);
  const T gd0 = (k == 0) ? T(0) : g_d0 * d0, gd1 = (k == K - 1) ? T(0) : g_d1 * d1;
  const T c0 = (k > 0) ? pd[k - 1] : T(0), c1 = (k < K - 1) ? pd[k] : T(0);
  for (int j = 0; j < K; ++j) {
    const T a = pw[j], b = ph[j];
    const T da = T(1) + M::abs_(a * cfg.inv_ls2), db = T(1) + M::abs_(b * cfg.inv_ls2);
    const T Wj = M::exp_(a / da - mw) / sw, Hj = M::exp_(b / db - mh) / sh;
    const T gW = tb * ((j <= k - 1 ? g_x0 : T(0)) + (j <= k ? g_x1 : T(0)));
    const T gH = tb * ((j <= k - 1 ? g_y0 : T(0)) + (j <= k ? g_y1 : T(0)));
    pw[j] = Wj * (gW - Sw) / (da * da);
    ph[j] = Hj * (gH - Sh) / (db * db);
  }
  for (int j = 0; j < K - 1; ++j) pd[j] = T(0);
  if (k > 0) { const T dd = T(1) + M::abs_(c0 * cfg.inv_ls1); pd[k - 1] = gd0 / (dd * dd); }
  if (k < K - 1) { const T dd = T(1) + M::abs_(c1 * cfg.inv_ls1); pd[k] = gd1 / (dd * dd); }
}

template <typename T>
__global__ void rqs_backward_kernel(const T* __restrict__ g_par, const T* __restrict__ g_x, long long n, int K, double bound,
                                    double slope, const T* __restrict__ g_gy, const T* __restrict__ g_gl,
                                    T* __restrict__ g_gpar, T* __restrict__ g_gx) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* s_par = reinterpret_cast<T*>(smem_raw);
  const int R = blockDim.x, P = 3 * K - 1, tid = threadIdx.x;
  SplineCfg<T> cfg;
  cfg.K = K; cfg.univariate = MF_UNIV_RQS; cfg.bound = (T)bound;
  cfg.inv_ls2 = slope > 0.0 ? (T)(2.0 / log(slope)) : T(0);
  cfg.inv_ls1 = slope > 0.0 ? (T)(1.0 / log(slope)) : T(0);
  const long long first = (long long)blockIdx.x * R;
  const int rows = (int)min((long long)R, n - first);
  for (int i = tid; i < rows * P; i += R) s_par[i] = g_par[first * P + i];  // coalesced, contiguous tile
  __syncthreads();
  if (tid < rows) {
    T gx;
    rqs_backward_row<T>(s_par + (size_t)tid * P, cfg, g_x[first + tid], g_gy[first + tid], g_gl ? g_gl[first + tid] : T(0), gx);
    g_gx[first + tid] = gx;
  }
  __syncthreads();
  for (int i = tid; i < rows * P; i += R) g_gpar[first * P + i] = s_par[i];
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
  const size_t tile = (size_t)rows * P * esz;
  MF_REQUIRE(tile <= 220 * 1024, "mf_rqs_transform: bins too large for shared memory");
  // persistent CTAs, four per SM; two parameter buffers per CTA when four CTAs x two tiles fit the SM's shared memory
  const int nbuf = (tile * 2 * 4 <= 216 * 1024) ? 2 : 1;
  const size_t smem = tile * nbuf;
  int sms = 148;
  {
    int dev = 0;
    MF_CUDA_CHECK(cudaGetDevice(&dev));
    MF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long n_tiles = (n + rows - 1) / rows;
  const unsigned grid = (unsigned)std::min<long long>(n_tiles, (long long)sms * 4);
  const int use_tma = tma_enabled() && (reinterpret_cast<uintptr_t>(d_params) & 15u) == 0;
  if (dtype == MF_F32) {
    MF_CUDA_CHECK(cudaFuncSetAttribute(rqs_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rqs_kernel<float><<<grid, rows, smem, st>>>((const float*)d_params, (const float*)d_x, n, bins, univariate, bound,
                                               slope, inverse, (float*)d_y, (float*)d_ladj, d_bins, d_inside, use_tma, nbuf);
  } else {
    MF_CUDA_CHECK(cudaFuncSetAttribute(rqs_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rqs_kernel<double><<<grid, rows, smem, st>>>((const double*)d_params, (const double*)d_x, n, bins, univariate, bound,
                                                slope, inverse, (double*)d_y, (double*)d_ladj, d_bins, d_inside, use_tma, nbuf);
  }
  return post_launch("rqs_kernel");
}

extern "C" int mf_rqs_backward(const void* d_params, const void* d_x, int64_t n, int32_t bins, double bound, double slope,
                               const void* d_grad_y, const void* d_grad_ladj, void* d_grad_params, void* d_grad_x, int dtype,
                               void* stream) {
  MF_REQUIRE(n >= 0, "mf_rqs_backward: negative count");
  MF_REQUIRE((d_params && d_x && d_grad_y && d_grad_params && d_grad_x) || n == 0, "mf_rqs_backward: null argument");
  MF_REQUIRE(bins >= 2 && bins <= 127, "mf_rqs_backward: bins=%d outside 2..127", bins);
  MF_REQUIRE(bound > 0.0, "mf_rqs_backward: bound must be positive");
  MF_REQUIRE(dtype == MF_F32 || dtype == MF_F64, "mf_rqs_backward: unsupported dtype %d", dtype);
  if (n == 0) return MF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = 3 * bins - 1;
  const size_t esz = dtype == MF_F32 ? 4 : 8;
  int rows = 128;
  while (rows > 32 && (size_t)rows * P * esz > 56 * 1024) rows >>= 1;
  const size_t smem = (size_t)rows * P * esz;
  const unsigned grid = (unsigned)((n + rows - 1) / rows);
  if (dtype == MF_F32) {
    MF_CUDA_CHECK(cudaFuncSetAttribute(rqs_backward_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rqs_backward_kernel<float><<<grid, rows, smem, st>>>((const float*)d_params, (const float*)d_x, n, bins, bound, slope,
                                                        (const float*)d_grad_y, (const float*)d_grad_ladj, (float*)d_grad_params,
                                                        (float*)d_grad_x);
  } else {
    MF_CUDA_CHECK(cudaFuncSetAttribute(rqs_backward_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rqs_backward_kernel<double><<<grid, rows, smem, st>>>((const double*)d_params, (const double*)d_x, n, bins, bound, slope,
                                                         (const double*)d_grad_y, (const double*)d_grad_ladj,
                                                         (double*)d_grad_params, (double*)d_grad_x);
  }
  return post_launch("rqs_backward_kernel");
}
