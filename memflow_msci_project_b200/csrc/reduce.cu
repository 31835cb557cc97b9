// reduce.cu -- per-event reduction of per-object log-probabilities.
//
// Fuses the tail of the reference's event likelihood:
//   flow_prob_batch = sum_objects(flow_prob_jet * maskExist)          (transfer_flow_paper_AllPartons_btag.py:216)
//   logp_g = flow().log_prob(ps) - log(rambo_jac)                     (notebooks/TestFlows_ImportanceSampling.ipynb cell 52)
// One thread per event; objects of an event are contiguous ([B, P] row-major), so a warp reads
// 32*P consecutive values.
#include "common.cuh"

namespace mf {
namespace {

template <typename T>
__global__ void event_reduce_kernel(const T* __restrict__ lp, const uint8_t* __restrict__ mask,
                                    const T* __restrict__ logw, long long B, int P, T* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < B;
       e += (long long)gridDim.x * blockDim.x) {
    T acc = T(0);
    for (int p = 0; p < P; ++p) {
      const T v = lp[e * P + p];
      if (!mask || mask[e * P + p]) acc += v;  // masked-out objects contribute exactly 0 (v * 0)
    }
    if (logw) acc -= logw[e];
    out[e] = acc;
  }
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" int mf_event_reduce(const void* d_logprob, const uint8_t* d_mask, const void* d_logw,
                               int64_t n_events, int32_t n_objects, void* d_out, int dtype, void* stream) {
  MF_REQUIRE(n_events >= 0 && n_objects >= 1, "mf_event_reduce: bad shape");
  MF_REQUIRE((d_logprob && d_out) || n_events == 0, "mf_event_reduce: null argument");
  MF_REQUIRE(dtype == MF_F32 || dtype == MF_F64, "mf_event_reduce: unsupported dtype %d", dtype);
  if (n_events == 0) return MF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((n_events + 255) / 256 < 148 * 16 ? (n_events + 255) / 256 : 148 * 16);
  if (dtype == MF_F32)
    event_reduce_kernel<float><<<grid, 256, 0, st>>>((const float*)d_logprob, d_mask, (const float*)d_logw,
                                                     n_events, n_objects, (float*)d_out);
  else
    event_reduce_kernel<double><<<grid, 256, 0, st>>>((const double*)d_logprob, d_mask, (const double*)d_logw,
                                                      n_events, n_objects, (double*)d_out);
  return post_launch("event_reduce_kernel");
}
