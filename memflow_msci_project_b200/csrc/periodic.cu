// periodic.cu -- regression loss of a periodic variable (phi), SURVEY 8(f)-4.
//
// Replaces `loss_fn_periodic(inp, target, loss_fn, device)` of the training scripts
// (scripts/run_model_labframe_sampling.py:65-77, called four times per step from compute_regr_losses :100-111) with
// loss_fn = torch.nn.HuberLoss(delta) or torch.nn.MSELoss:
//
//     inp    <- inp - 2 pi where inp > pi,  inp + 2 pi where inp < -pi          (one wrap, :68-69)
//     dphi   =  target - inp;  dphi - 2 pi where dphi > pi;  dphi + 2 pi where dphi <= -pi   (:71-73)
//     loss_i =  huber_delta(dphi) | dphi^2                                    (loss_fn(dphi, 0), :74)
//
// The reference runs about ten elementwise torch kernels plus the loss; here one pass writes loss_i, d loss_i / d dphi
// (for the backward: d/d target = +, d/d inp = -) and per-CTA partial sums, and a one-CTA kernel adds the partials in a
// fixed order.  HBM-bound, 2 reads + up to 2 writes per element; at the reference's batch sizes it is launch-bound.
#include "common.cuh"

namespace mf {
namespace {

constexpr int PER_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(PER_THREADS)
periodic_loss_kernel(const T* __restrict__ inp, const T* __restrict__ target, long long n, int kind, T delta,
                     T* __restrict__ elem, T* __restrict__ deriv, double* __restrict__ partial) {
  __shared__ double red[PER_THREADS];
  const T PI = T(3.14159265358979323846);  // the scripts' PI = torch.pi, rounded to the tensors' dtype by torch
  const T TWO_PI = T(2) * PI;
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)PER_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * PER_THREADS) {
    T a = inp[i];
    if (a > PI) a = a - TWO_PI;
    else if (a < -PI) a = a + TWO_PI;
    T dphi = target[i] - a;
    if (dphi > PI) dphi = dphi - TWO_PI;
    else if (dphi <= -PI) dphi = dphi + TWO_PI;
    T l, g;
    if (kind == MF_LOSS_MSE) {
      l = dphi * dphi;
      g = T(2) * dphi;
    } else {  // torch.nn.HuberLoss: 0.5 d^2 if |d| < delta else delta (|d| - 0.5 delta)
      // written like torch's kernels so that a NaN input gives NaN loss AND NaN gradient as in the reference
      const T ad = fabs(dphi);
      l = ad < delta ? T(0.5) * ad * ad : delta * (ad - T(0.5) * delta);
      g = dphi < -delta ? -delta : (dphi > delta ? delta : dphi);
    }
    if (elem) elem[i] = l;
    if (deriv) deriv[i] = g;
    acc += (double)l;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = PER_THREADS / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

template <typename T>
__global__ void periodic_finish_kernel(const double* __restrict__ partial, int parts, double scale, T* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int p = 0; p < parts; ++p) a += partial[p];
    out[0] = (T)(a * scale);
  }
}

inline int periodic_grid(int64_t n) {
  const int64_t b = (n + PER_THREADS - 1) / PER_THREADS;
  return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" int64_t mf_periodic_loss_workspace_bytes(int64_t n) {
  return n < 1 ? 0 : (int64_t)periodic_grid(n) * (int64_t)sizeof(double);
}

extern "C" int mf_periodic_loss(const void* d_inp, const void* d_target, int64_t n, int32_t kind, double huber_delta,
                                int32_t mean, void* d_workspace, void* d_elem, void* d_deriv, void* d_total, int dtype,
                                void* stream) {
  MF_REQUIRE(n >= 1, "mf_periodic_loss: n = %lld < 1", (long long)n);
  MF_REQUIRE(kind == MF_LOSS_HUBER || kind == MF_LOSS_MSE, "mf_periodic_loss: unknown loss kind %d", kind);
  MF_REQUIRE(kind != MF_LOSS_HUBER || huber_delta > 0.0, "mf_periodic_loss: Huber delta must be positive");
  MF_REQUIRE(dtype == MF_F32 || dtype == MF_F64, "mf_periodic_loss: unsupported dtype %d", dtype);
  MF_REQUIRE(d_inp && d_target && d_workspace, "mf_periodic_loss: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = periodic_grid(n);
  double* ws = static_cast<double*>(d_workspace);
  const double scale = mean ? 1.0 / (double)n : 1.0;
  if (dtype == MF_F32)
    periodic_loss_kernel<float><<<grid, PER_THREADS, 0, st>>>((const float*)d_inp, (const float*)d_target, n, kind,
                                                              (float)huber_delta, (float*)d_elem, (float*)d_deriv, ws);
  else
    periodic_loss_kernel<double><<<grid, PER_THREADS, 0, st>>>((const double*)d_inp, (const double*)d_target, n, kind,
                                                               huber_delta, (double*)d_elem, (double*)d_deriv, ws);
  int rc = post_launch("periodic_loss_kernel");
  if (rc != MF_OK || !d_total) return rc;
  if (dtype == MF_F32)
    periodic_finish_kernel<float><<<1, 32, 0, st>>>(ws, grid, scale, (float*)d_total);
  else
    periodic_finish_kernel<double><<<1, 32, 0, st>>>(ws, grid, scale, (double*)d_total);
  return post_launch("periodic_finish_kernel");
}
