// flow.cu -- fused conditional RQ-spline MAF (zuko NSF family) on CUDA cores: the "exact" path.
//
// Replaces, per call of flow(c).log_prob(x) / .sample(), T x (L+1) masked F.linear launches plus
// ~40 element-wise launches per transform whose spline parameters [N, D, 3K-1] round-trip HBM
//  with ONE kernel: a CTA owns a tile of TM objects, keeps (x | context),
// the hidden activations and the spline parameters of the tile in shared memory, streams the
// pre-masked weights (L2 resident, a few MB) through a cp.async double buffer, and only reads
// x, context and writes log_prob / z / x from and to HBM.  Arithmetic is plain FMA in the I/O dtype
// (fp32 or fp64), i.e. the same class of arithmetic as the reference torch path; the tcgen05
// variant lives in flow_tc.cu.
//
// zuko semantics: MaskedAutoregressiveTransform.meta (hyper-network on cat(x, c)),
// AutoregressiveTransform.call_and_ladj / _inverse (`passes` fixed-point sweeps),
// MonotonicRQSTransform, NormalizingFlow.log_prob / rsample.
#include <algorithm>

#include "flow_common.cuh"

namespace mf {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// pack: zuko-layout (W [out,in], bias [out], mask [out,in]) -> blob
template <typename T>
struct PackArgs {
  const T* w[kMaxLayers];
  const T* b[kMaxLayers];
  const uint8_t* m[kMaxLayers];
};

template <typename T>
__global__ void flow_pack_kernel(FlowLayout L, PackArgs<T> a, T* blob /* this transform */) {
  const long long total = L.t_stride;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // find the layer this element belongs to
    int l = L.n_layers - 1;
    for (int q = 0; q + 1 < L.n_layers; ++q)
      if (i < L.w_off[q + 1]) { l = q; break; }
    T val = T(0);
    if (l < L.n_layers - 1) {
      const long long r = i - L.w_off[l];
      const long long nW = (long long)L.kpad[l] * L.ld[l];
      if (r < nW) {
        const int k = (int)(r / L.ld[l]), j = (int)(r % L.ld[l]);
        if (k < L.in_dim[l] && j < L.out_dim[l]) {
          const long long src = (long long)j * L.in_dim[l] + k;
          val = a.m[l][src] ? a.w[l][src] : T(0);
        }
      } else {
        const int j = (int)(r - nW);
        if (j < L.out_dim[l]) val = a.b[l][j];
      }
    } else {
      const long long r = i - L.w_off[l];
      const int g = (int)(r / L.chunk_stride);
      const long long rr = r % L.chunk_stride;
      const long long nW = (long long)L.kpad[l] * L.pw_ld;
      int k = -1, col;
      if (rr < nW) { k = (int)(rr / L.pw_ld); col = (int)(rr % L.pw_ld); } else { col = (int)(rr - nW); }
      const int fl = col / L.PF, j = col % L.PF, f = g * L.fpc + fl;
      if (fl < L.fpc && f < L.D && j < L.P) {
        const int o = f * L.P + j;  // zuko: phi.unflatten(-1, (D, 3K-1))
        if (k < 0) val = a.b[l][o];
        else if (k < L.in_dim[l]) {
          const long long src = (long long)o * L.in_dim[l] + k;
          val = a.m[l][src] ? a.w[l][src] : T(0);
        }
      }
    }
    blob[i] = val;
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage one [kKC][kNC] weight tile (row stride ldw in global) into shared memory
template <typename T>
__device__ __forceinline__ void stage_w(T* dst, const T* __restrict__ src, int ldw) {
  constexpr int kVec = 16 / sizeof(T);              // elements per 16-byte copy
  constexpr int kPerRow = kNC / kVec;               // copies per row
  constexpr int kTotal = kKC * kPerRow;
  for (int i = threadIdx.x; i < kTotal; i += kThreads) {
    const int k = i / kPerRow, c = (i % kPerRow) * kVec;
    cp_async16(dst + k * kNC + c, src + (size_t)k * ldw + c);
  }
}

// 4-element vector load/store on shared memory (16-byte aligned for float, 32 for double)
__device__ __forceinline__ void ld4(float* d, const float* s) {
  const float4 v = *reinterpret_cast<const float4*>(s);
  d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
}
__device__ __forceinline__ void ld4(double* d, const double* s) {
  const double2 a = *reinterpret_cast<const double2*>(s), b = *reinterpret_cast<const double2*>(s + 2);
  d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
}
__device__ __forceinline__ void st4(float* d, const float* s) {
  *reinterpret_cast<float4*>(d) = make_float4(s[0], s[1], s[2], s[3]);
}
__device__ __forceinline__ void st4(double* d, const double* s) {
  *reinterpret_cast<double2*>(d) = make_double2(s[0], s[1]);
  *reinterpret_cast<double2*>(d + 2) = make_double2(s[2], s[3]);
}

// Out[TM][0..127] (+col0) = act( A[TM][K] * Wt[K][col0..col0+127] + bias ), TM = 16*RM, 256 threads,
// thread (ty = tid/16, tx = tid%16) owns rows ty*RM.. and the 8 columns tx*4..+3 and 64+tx*4..+3.
template <typename T, int RM>
__device__ __forceinline__ void gemm_pass(const T* A, int lda, int K, const T* __restrict__ Wt, int ldw,
                                          const T* __restrict__ bias, T* Out, int ldo, bool relu,
                                          T* wstage) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T acc[RM][8];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = T(0);
  const int nk = K / kKC;
  stage_w(wstage, Wt, ldw);
  cp_async_commit();
  for (int kc = 0; kc < nk; ++kc) {
    T* cur = wstage + (kc & 1) * (kKC * kNC);
    if (kc + 1 < nk) {
      stage_w(wstage + ((kc + 1) & 1) * (kKC * kNC), Wt + (size_t)(kc + 1) * kKC * ldw, ldw);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const T* a_base = A + (size_t)(ty * RM) * lda + kc * kKC;
#pragma unroll
    for (int k4 = 0; k4 < kKC; k4 += 4) {
      T av[RM][4];
#pragma unroll
      for (int i = 0; i < RM; ++i) ld4(av[i], a_base + (size_t)i * lda + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        T w[8];
        ld4(w, cur + (k4 + kk) * kNC + tx * 4);
        ld4(w + 4, cur + (k4 + kk) * kNC + 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i][kk], w[j], acc[i][j]);
      }
    }
    __syncthreads();  // everyone done with `cur` before it is overwritten two iterations later
  }
  T bj[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bj[j] = bias[tx * 4 + j];
    bj[4 + j] = bias[64 + tx * 4 + j];
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    T v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = acc[i][j] + bj[j];
      if (relu) v[j] = v[j] > T(0) ? v[j] : T(0);
    }
    T* o = Out + (size_t)(ty * RM + i) * ldo;
    st4(o + tx * 4, v);
    st4(o + 64 + tx * 4, v + 4);
  }
}

template <typename T, int RM>
__global__ void __launch_bounds__(kThreads, 1)
flow_kernel(const __grid_constant__ FlowLayout L, const __grid_constant__ FlowKernelArgs args,
            const __grid_constant__ BaseParams bp) {
  constexpr int TM = 16 * RM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* in0 = reinterpret_cast<T*>(smem_raw);           // [TM][in0_ld]  (x | ctx | 0-pad)
  T* buf0 = in0 + (size_t)TM * L.in0_ld;             // [TM][buf_w]
  T* buf1 = buf0 + (size_t)TM * L.buf_w;             // [TM][buf_w]
  T* wstage = buf1 + (size_t)TM * L.buf_w;           // [2][kKC][kNC]
  T* xnew = wstage + 2 * kKC * kNC;                  // [TM][D]
  T* ytgt = xnew + (size_t)TM * L.D;                 // [TM][D]   (inverse: target of the transform)
  T* ladj_s = ytgt + (size_t)TM * L.D;               // [TM]

  const T* blob = static_cast<const T*>(args.blob);
  const T* gx = static_cast<const T*>(args.x);
  const T* gc = static_cast<const T*>(args.ctx);
  const int D = L.D, C = L.C, Lh = L.n_layers - 1;
  const int tid = threadIdx.x;

  SplineCfg<T> cfg;
  cfg.K = L.K;
  cfg.univariate = args.univariate;
  cfg.bound = (T)args.bound;
  cfg.inv_ls2 = args.slope > 0.0 ? (T)(2.0 / log(args.slope)) : T(0);
  cfg.inv_ls1 = args.slope > 0.0 ? (T)(1.0 / log(args.slope)) : T(0);

  for (long long tile0 = (long long)blockIdx.x * TM; tile0 < args.n; tile0 += (long long)gridDim.x * TM) {
    const int rows = (int)min((long long)TM, args.n - tile0);
    // ---- load the tile: x (or base sample) and context, zero padding
    for (int i = tid; i < TM * L.in0_ld; i += kThreads) {
      const int m = i / L.in0_ld, c = i % L.in0_ld;
      T v = T(0);
      if (m < rows) {
        if (c < D) {
          v = gx[(tile0 + m) * D + c];
          if (args.mode == 1) {  // base sample from caller-supplied noise
            v = (args.base == MF_BASE_DIAG_NORMAL) ? (T)bp.a[c] + (T)bp.b[c] * v
                                                  : (T)bp.a[c] + ((T)bp.b[c] - (T)bp.a[c]) * v;
          } else if (bp.xmap_kind != MF_XMAP_NONE) {
            v = xmap_apply<T>(bp, v, c);   // input map of log_prob (logit bridge)
          }
        } else if (c < D + C) {
          v = gc[((tile0 + m) / args.ctx_group) * C + (c - D)];
        }
      }
      in0[i] = v;
    }
    for (int i = tid; i < TM; i += kThreads) ladj_s[i] = T(0);
    __syncthreads();

    const int n_outer = L.T;
    for (int ti = 0; ti < n_outer; ++ti) {
      const int t = (args.mode == 0) ? ti : (L.T - 1 - ti);
      const T* tb = blob + (size_t)t * L.t_stride;
      int sweeps = 1;
      if (args.mode == 1) {
        sweeps = args.passes;
        // y <- current value, x <- 0   (AutoregressiveTransform._inverse)
        for (int i = tid; i < TM * D; i += kThreads) {
          const int m = i / D, f = i % D;
          ytgt[i] = in0[(size_t)m * L.in0_ld + f];
          in0[(size_t)m * L.in0_ld + f] = T(0);
        }
        __syncthreads();
      }
      for (int sw = 0; sw < sweeps; ++sw) {
        // ---- hyper-network
        const T* cur = in0;
        int lda = L.in0_ld;
        for (int l = 0; l < Lh; ++l) {
          T* out = (l & 1) ? buf1 : buf0;
          for (int c0 = 0; c0 < L.ld[l]; c0 += kNC)
            gemm_pass<T, RM>(cur, lda, L.kpad[l], tb + L.w_off[l] + c0, L.ld[l], tb + L.b_off[l] + c0,
                             out + c0, L.buf_w, true, wstage);
          __syncthreads();
          cur = out;
          lda = L.buf_w;
        }
        T* par = ((Lh - 1) & 1) ? buf0 : buf1;  // the buffer not holding the last hidden layer
        // sampling: a chunk is evaluated in sweep `sw` only if one of its features becomes final there (its inputs
        // are final, later sweeps would recompute the same value; the other features of the pass are provisional
        // and masked out of everything that is kept) -- same result as zuko's full sweeps, fewer last-layer GEMMs
        const unsigned long long live = (args.mode == 1 && args.sched) ? args.sched[(size_t)t * MF_MAX_FEATURES + sw] : ~0ull;
        for (int g = 0; g < L.n_chunks; ++g) {
          if (!((live >> (g * L.fpc)) & ((1ull << min(L.fpc, D - g * L.fpc)) - 1ull))) {
            const int f0s = g * L.fpc, nfs = min(L.fpc, D - f0s);
            for (int p = tid; p < TM * nfs; p += kThreads) {
              const int m = p % TM, f = f0s + p / TM;
              xnew[(size_t)m * D + f] = in0[(size_t)m * L.in0_ld + f];
            }
            continue;
          }
          const T* wl = tb + L.w_off[Lh] + (size_t)g * L.chunk_stride;
          const T* bl = tb + L.b_off[Lh] + (size_t)g * L.chunk_stride;
          for (int c0 = 0; c0 < L.pw_ld; c0 += kNC)
            gemm_pass<T, RM>(cur, lda, L.kpad[Lh], wl + c0, L.pw_ld, bl + c0, par + c0, L.buf_w, false,
                             wstage);
          __syncthreads();
          // ---- splines of the features of this chunk: thread per (row, feature)
          const int f0 = g * L.fpc;
          const int nf = min(L.fpc, D - f0);
          for (int p = tid; p < TM * nf; p += kThreads) {
            const int m = p % TM, fl = p / TM, f = f0 + fl;
            if (m < rows) {
              T* pp = par + (size_t)m * L.buf_w + fl * L.PF;
              const T v = (args.mode == 0) ? in0[(size_t)m * L.in0_ld + f] : ytgt[(size_t)m * D + f];
              T out, la;
              int bin;
              bool ins;
              if (!((live >> f) & 1ull)) {  // provisional feature of this sweep: keep the current value
                xnew[(size_t)m * D + f] = in0[(size_t)m * L.in0_ld + f];
                continue;
              }
              rqs_univariate<T>(pp, 1, cfg, args.mode == 1, v, out, la, bin, ins);
              xnew[(size_t)m * D + f] = out;
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
        // ---- x <- transformed values; ladj += sum over features (DependentTransform(.., 1))
        for (int i = tid; i < TM * D; i += kThreads) {
          const int m = i / D, f = i % D;
          in0[(size_t)m * L.in0_ld + f] = xnew[i];
        }
        if (args.mode == 0) {
          for (int m = tid; m < rows; m += kThreads) {
            T sacc = T(0);
            for (int f = 0; f < D; ++f) sacc += ytgt[(size_t)m * D + f];
            ladj_s[m] += sacc;
          }
        }
        __syncthreads();
      }
    }

    // ---- outputs
    if (args.mode == 0) {
      if (args.z) {
        T* gz = static_cast<T*>(args.z);
        for (int i = tid; i < rows * D; i += kThreads) gz[tile0 * D + i] = in0[(size_t)(i / D) * L.in0_ld + (i % D)];
      }
      for (int m = tid; m < rows; m += kThreads) {
        const T la = ladj_s[m];
        if (args.ladj) static_cast<T*>(args.ladj)[tile0 + m] = la;
        if (args.logprob) {
          T lp = T(0);
          for (int f = 0; f < D; ++f)
            lp += base_log_prob_1<T>(args.base, in0[(size_t)m * L.in0_ld + f], (T)bp.a[f], (T)bp.b[f]);
          static_cast<T*>(args.logprob)[tile0 + m] = lp + la;
        }
      }
    } else {
      T* gz = static_cast<T*>(args.z);
      for (int i = tid; i < rows * D; i += kThreads) {
        T v = in0[(size_t)(i / D) * L.in0_ld + (i % D)];
        if (bp.xmap_kind != MF_XMAP_NONE) v = xmap_apply<T>(bp, v, i % D);   // output map of sample (sigmoid bridge)
        gz[tile0 * D + i] = v;
      }
    }
    __syncthreads();
  }
}

template <typename T, int RM>
size_t flow_smem_bytes(const FlowLayout& L) {
  const size_t TM = 16 * RM;
  return sizeof(T) * (TM * L.in0_ld + 2 * TM * L.buf_w + 2 * kKC * kNC + 2 * TM * L.D + TM);
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <typename T, int RM>
int launch_flow_rm(const FlowLayout& L, const FlowKernelArgs& a, const BaseParams& bp, cudaStream_t st) {
  const size_t smem = flow_smem_bytes<T, RM>(L);
  auto kern = flow_kernel<T, RM>;
  MF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long tiles = (a.n + 16 * RM - 1) / (16 * RM);
  const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)num_sms());  // persistent, 1 CTA / SM
  kern<<<grid, kThreads, smem, st>>>(L, a, bp);
  return post_launch("flow_kernel");
}

template <typename T>
int launch_flow(const FlowLayout& L, const FlowKernelArgs& a, const BaseParams& bp, cudaStream_t st) {
  constexpr size_t kMaxSmem = 227 * 1024;
  if (flow_smem_bytes<T, 8>(L) <= kMaxSmem) return launch_flow_rm<T, 8>(L, a, bp, st);
  if (flow_smem_bytes<T, 4>(L) <= kMaxSmem) return launch_flow_rm<T, 4>(L, a, bp, st);
  if (flow_smem_bytes<T, 2>(L) <= kMaxSmem) return launch_flow_rm<T, 2>(L, a, bp, st);
  if (flow_smem_bytes<T, 1>(L) <= kMaxSmem) return launch_flow_rm<T, 1>(L, a, bp, st);
  set_error("flow too wide for shared memory (hidden width / bins too large)");
  return MF_EUNSUPPORTED;
}

inline long long exact_sched_offset(const FlowLayout& L, int dtype) {  // 8-byte aligned end of the weight region
  const long long w = (long long)L.t_stride * L.T * (dtype == MF_F32 ? 4 : 8);
  return (w + 7) / 8 * 8;
}

inline bool is_tcgen05(int gemm) { return gemm == MF_GEMM_TC_3XF16 || gemm == MF_GEMM_TC_BF16; }

int check_flow_common(const mf_flow_desc* desc, int dtype, int gemm) {
  MF_REQUIRE(desc != nullptr, "null flow descriptor");
  MF_REQUIRE(dtype == MF_F32 || dtype == MF_F64, "unsupported dtype %d", dtype);
  MF_REQUIRE(gemm == MF_GEMM_EXACT || gemm == MF_GEMM_TC_3XF16 || gemm == MF_GEMM_TC_BF16 || gemm == MF_GEMM_MMA_3XF16,
             "unknown gemm mode %d", gemm);
  if (gemm != MF_GEMM_EXACT && dtype != MF_F32) {
    set_error("tensor-core gemm modes take fp32 tensors (fp64 flows use MF_GEMM_EXACT)");
    return MF_EUNSUPPORTED;
  }
  return MF_OK;
}

}  // namespace
}  // namespace mf

using namespace mf;

extern "C" {

int64_t mf_flow_blob_bytes(const mf_flow_desc* desc, int dtype, int gemm) {
  if (check_flow_common(desc, dtype, gemm) != MF_OK) return -1;
  if (is_tcgen05(gemm)) {
    long long b = 0;
    return flow_tc_blob_bytes(*desc, gemm, &b) == MF_OK ? (int64_t)b : -1;
  }
  FlowLayout L;
  if (make_flow_layout(*desc, &L) != MF_OK) return -1;
  return exact_sched_offset(L, dtype) + kSchedBytes + (gemm == MF_GEMM_MMA_3XF16 ? kKlimBytes : 0);
}

int mf_flow_pack_weights(const mf_flow_desc* desc, const void* const* d_weights,
                         const void* const* d_biases, const uint8_t* const* d_masks, int dtype, int gemm,
                         void* d_blob, int64_t blob_bytes, void* stream) {
  int rc = check_flow_common(desc, dtype, gemm);
  if (rc != MF_OK) return rc;
  MF_REQUIRE(d_weights && d_biases && d_masks && d_blob, "mf_flow_pack_weights: null argument");
  if (is_tcgen05(gemm))
    return flow_tc_pack(*desc, d_weights, d_biases, d_masks, gemm, d_blob, blob_bytes, static_cast<cudaStream_t>(stream));
  FlowLayout L;
  rc = make_flow_layout(*desc, &L);
  if (rc != MF_OK) return rc;
  const int64_t need = exact_sched_offset(L, dtype) + kSchedBytes + (gemm == MF_GEMM_MMA_3XF16 ? kKlimBytes : 0);
  MF_REQUIRE(blob_bytes >= need, "blob too small: %lld < %lld bytes", (long long)blob_bytes, (long long)need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nl = L.n_layers;
  for (int t = 0; t < L.T; ++t) {
    const unsigned grid = (unsigned)std::min<long long>((L.t_stride + 255) / 256, 148LL * 8);
    if (dtype == MF_F32) {
      PackArgs<float> a;
      for (int l = 0; l < nl; ++l) {
        a.w[l] = (const float*)d_weights[t * nl + l];
        a.b[l] = (const float*)d_biases[t * nl + l];
        a.m[l] = d_masks[t * nl + l];
        MF_REQUIRE(a.w[l] && a.b[l] && a.m[l], "null parameter pointer (transform %d layer %d)", t, l);
      }
      flow_pack_kernel<float><<<grid, 256, 0, st>>>(L, a, (float*)d_blob + (size_t)t * L.t_stride);
    } else {
      PackArgs<double> a;
      for (int l = 0; l < nl; ++l) {
        a.w[l] = (const double*)d_weights[t * nl + l];
        a.b[l] = (const double*)d_biases[t * nl + l];
        a.m[l] = d_masks[t * nl + l];
        MF_REQUIRE(a.w[l] && a.b[l] && a.m[l], "null parameter pointer (transform %d layer %d)", t, l);
      }
      flow_pack_kernel<double><<<grid, 256, 0, st>>>(L, a, (double*)d_blob + (size_t)t * L.t_stride);
    }
    rc = post_launch("flow_pack_kernel");
    if (rc != MF_OK) return rc;
  }
  MF_CUDA_CHECK(cudaMemsetAsync((char*)d_blob + exact_sched_offset(L, dtype), 0xFF, kSchedBytes, st));
  if (gemm == MF_GEMM_MMA_3XF16)
    return flow_mma_presplit(L, (float*)d_blob,
                             reinterpret_cast<unsigned short*>((char*)d_blob + exact_sched_offset(L, dtype) + kSchedBytes), st);
  return MF_OK;
}

static int fill_xmap(BaseParams* bp, const mf_flow_xmap* map, int mode, int features) {
  bp->xmap_kind = MF_XMAP_NONE; bp->xmap_eps = 0.0;
  for (int f = 0; f < MF_MAX_FEATURES; ++f) { bp->xmap_scale[f] = 1.0; bp->xmap_shift[f] = 0.0; }
  if (map == nullptr || map->kind == MF_XMAP_NONE) return MF_OK;
  MF_REQUIRE((mode == 0 && map->kind == MF_XMAP_LOGIT_AFFINE) || (mode == 1 && map->kind == MF_XMAP_SIGMOID_AFFINE),
             "mf_flow_xmap kind %d does not apply to this call (log_prob takes MF_XMAP_LOGIT_AFFINE on its input, sample "
             "MF_XMAP_SIGMOID_AFFINE on its output)", map->kind);
  bp->xmap_kind = map->kind; bp->xmap_eps = map->eps;
  for (int f = 0; f < features; ++f) {
    MF_REQUIRE(map->kind != MF_XMAP_LOGIT_AFFINE || map->scale[f] != 0.0, "mf_flow_xmap: zero scale for feature %d", f);
    bp->xmap_scale[f] = map->scale[f]; bp->xmap_shift[f] = map->shift[f];
  }
  return MF_OK;
}

static int flow_run(const mf_flow_desc* desc, const FlowKernelArgs& a, int dtype, int gemm, const mf_flow_xmap* map, void* stream) {
  int rc = check_flow_common(desc, dtype, gemm);
  if (rc != MF_OK) return rc;
  FlowLayout L;
  rc = make_flow_layout(*desc, &L);
  if (rc != MF_OK) return rc;
  if (is_tcgen05(gemm)) {
    MF_REQUIRE(desc->context == 0 || a.ctx != nullptr || a.n == 0, "context pointer is null but context=%d", desc->context);
    MF_REQUIRE(desc->bound > 0.0, "bound must be positive");
    if (a.n == 0) return MF_OK;
    BaseParams bpt;
    for (int f = 0; f < MF_MAX_FEATURES; ++f) { bpt.a[f] = desc->base_a[f]; bpt.b[f] = desc->base_b[f]; }
    rc = fill_xmap(&bpt, map, a.mode, desc->features);
    if (rc != MF_OK) return rc;
    return flow_tc_run(*desc, a, bpt, gemm, static_cast<cudaStream_t>(stream));
  }
  MF_REQUIRE(desc->context == 0 || a.ctx != nullptr || a.n == 0, "context pointer is null but context=%d", desc->context);
  MF_REQUIRE(desc->bound > 0.0, "bound must be positive");
  if (a.n == 0) return MF_OK;
  BaseParams bp;
  for (int f = 0; f < MF_MAX_FEATURES; ++f) { bp.a[f] = desc->base_a[f]; bp.b[f] = desc->base_b[f]; }
  rc = fill_xmap(&bp, map, a.mode, desc->features);
  if (rc != MF_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FlowKernelArgs as = a;
  if (a.mode == 1)
    as.sched = reinterpret_cast<const unsigned long long*>(static_cast<const char*>(a.blob) + exact_sched_offset(L, dtype));
  if (gemm == MF_GEMM_MMA_3XF16) {
    as.klim = reinterpret_cast<const unsigned short*>(static_cast<const char*>(a.blob) + exact_sched_offset(L, dtype) + kSchedBytes);
    return flow_mma_run(L, as, bp, st);
  }
  return dtype == MF_F32 ? launch_flow<float>(L, as, bp, st) : launch_flow<double>(L, as, bp, st);
}

int mf_flow_log_prob_mapped(const mf_flow_desc* desc, const void* d_blob, const void* d_x, const void* d_ctx,
                            int64_t n, void* d_logprob, void* d_z, void* d_ladj, int8_t* d_bins,
                            uint8_t* d_inside, int dtype, int gemm, void* d_workspace, int64_t workspace_bytes,
                            const mf_flow_xmap* in_map, void* stream) {
  MF_REQUIRE(n >= 0, "mf_flow_log_prob: negative count");
  MF_REQUIRE(desc && d_blob && (d_x || n == 0), "mf_flow_log_prob: null argument");
  FlowKernelArgs a{};
  a.blob = d_blob; a.x = d_x; a.ctx = d_ctx; a.n = n;
  a.logprob = d_logprob; a.z = d_z; a.ladj = d_ladj; a.bins = d_bins; a.inside = d_inside;
  a.work = d_workspace; a.work_bytes = workspace_bytes;
  a.ctx_group = desc->ctx_group > 1 ? desc->ctx_group : 1;
  a.mode = 0; a.passes = desc->passes; a.base = desc->base; a.univariate = desc->univariate;
  a.bound = desc->bound; a.slope = desc->slope;
  return flow_run(desc, a, dtype, gemm, in_map, stream);
}

int mf_flow_log_prob(const mf_flow_desc* desc, const void* d_blob, const void* d_x, const void* d_ctx,
                     int64_t n, void* d_logprob, void* d_z, void* d_ladj, int8_t* d_bins,
                     uint8_t* d_inside, int dtype, int gemm, void* d_workspace, int64_t workspace_bytes,
                     void* stream) {
  return mf_flow_log_prob_mapped(desc, d_blob, d_x, d_ctx, n, d_logprob, d_z, d_ladj, d_bins, d_inside, dtype, gemm,
                                 d_workspace, workspace_bytes, nullptr, stream);
}

int64_t mf_flow_workspace_bytes(const mf_flow_desc* desc, int64_t n, int dtype, int gemm, int op) {
  if (check_flow_common(desc, dtype, gemm) != MF_OK || n < 0) return -1;
  if (!is_tcgen05(gemm)) return 0;
  return flow_tc_workspace_bytes(*desc, n, gemm, op);
}

int mf_flow_sample_mapped(const mf_flow_desc* desc, const void* d_blob, const void* d_noise, const void* d_ctx,
                          int64_t n, void* d_x, int dtype, int gemm, void* d_workspace, int64_t workspace_bytes,
                          const mf_flow_xmap* out_map, void* stream) {
  MF_REQUIRE(n >= 0, "mf_flow_sample: negative count");
  MF_REQUIRE(desc && d_blob && ((d_noise && d_x) || n == 0), "mf_flow_sample: null argument");
  MF_REQUIRE(desc->passes >= 1 && desc->passes <= desc->features, "passes=%d outside 1..features", desc->passes);
  FlowKernelArgs a{};
  a.blob = d_blob; a.x = d_noise; a.ctx = d_ctx; a.n = n;
  a.z = d_x;
  a.work = d_workspace; a.work_bytes = workspace_bytes;
  a.ctx_group = desc->ctx_group > 1 ? desc->ctx_group : 1;
  a.mode = 1; a.passes = desc->passes; a.base = desc->base; a.univariate = desc->univariate;
  a.bound = desc->bound; a.slope = desc->slope;
  return flow_run(desc, a, dtype, gemm, out_map, stream);
}

int mf_flow_sample(const mf_flow_desc* desc, const void* d_blob, const void* d_noise, const void* d_ctx,
                   int64_t n, void* d_x, int dtype, int gemm, void* d_workspace, int64_t workspace_bytes,
                   void* stream) {
  return mf_flow_sample_mapped(desc, d_blob, d_noise, d_ctx, n, d_x, dtype, gemm, d_workspace, workspace_bytes, nullptr,
                               stream);
}

int64_t mf_flow_sched_offset(const mf_flow_desc* desc, int dtype, int gemm) {
  if (check_flow_common(desc, dtype, gemm) != MF_OK) return -1;
  long long off = 0;
  int fpc = 1;
  if (is_tcgen05(gemm)) return flow_tc_sched_layout(*desc, gemm, &off, &fpc) == MF_OK ? (int64_t)off : -1;
  FlowLayout L;
  if (make_flow_layout(*desc, &L) != MF_OK) return -1;
  return exact_sched_offset(L, dtype);
}

int mf_flow_set_sample_schedule(const mf_flow_desc* desc, const int32_t* h_orders, int dtype, int gemm, void* d_blob,
                                int64_t blob_bytes, void* stream) {
  int rc = check_flow_common(desc, dtype, gemm);
  if (rc != MF_OK) return rc;
  MF_REQUIRE(h_orders && d_blob, "mf_flow_set_sample_schedule: null argument");
  MF_REQUIRE(desc->passes >= 1 && desc->passes <= desc->features, "passes=%d outside 1..features", desc->passes);
  long long off = 0;
  int fpc = 1;
  if (is_tcgen05(gemm)) {
    rc = flow_tc_sched_layout(*desc, gemm, &off, &fpc);
    if (rc != MF_OK) return rc;
  } else {
    FlowLayout L;
    rc = make_flow_layout(*desc, &L);
    if (rc != MF_OK) return rc;
    off = exact_sched_offset(L, dtype);
    fpc = L.fpc;
  }
  MF_REQUIRE(blob_bytes >= off + kSchedBytes, "blob too small for the sampling schedule");
  static thread_local unsigned long long words[MF_MAX_TRANSFORMS * MF_MAX_FEATURES];
  for (int t = 0; t < desc->transforms; ++t)
    for (int s = 0; s < MF_MAX_FEATURES; ++s) {
      unsigned long long m = 0;
      for (int f = 0; f < desc->features; ++f) {
        const int group = h_orders[t * desc->features + f];
        MF_REQUIRE(group >= 0 && group < desc->passes, "order[%d][%d]=%d outside 0..passes-1", t, f, group);
        if (group == s) m |= 1ull << f;
      }
      if (m == 0) m = 1;  // a sweep without a feature would leave the kernels' job sequence without its last-layer step
      words[t * MF_MAX_FEATURES + s] = m;
    }
  MF_CUDA_CHECK(cudaMemcpyAsync((char*)d_blob + off, words, (size_t)desc->transforms * MF_MAX_FEATURES * 8,
                                cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  MF_CUDA_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));  // `words` is reused by the next call
  return MF_OK;
}

}  // extern "C"
