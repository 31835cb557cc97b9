/*
 * memflow_b200.h -- C ABI of the B200-native MEMFlow hot path.
 *
 * Drop-in boundary for the batched per-event evaluation path of MEMFlow
 * (reference: themrluke/memflow-msci-project, citations are file:line relative
 * to the reference tree):
 *
 *   - RAMBO-on-diet phase-space map + Jacobian
 *       memflow/phasespace/rambo_generator.py:168-398  (generateKinematics_batch)
 *       memflow/phasespace/rambo_generator.py:615-736  (getPSpoint_batch)
 *       memflow/unfolding_flow/utils.py:1056-1152      (Compute_ParticlesTensor.get_PS)
 *   - conditional rational-quadratic-spline masked autoregressive flow
 *       zuko (third party, pinned zuko=1.2.0 in env.yml:40): NSF/MAF .log_prob / .sample,
 *       called from e.g. memflow/unfolding_flow/unfolding_flow.py:179,186 and
 *       memflow/transfer_flow/transfer_flow_paper_AllPartons_btag.py:202-213
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device,
 *     every other pointer is a HOST pointer that is only read during the call;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - all functions are asynchronous w.r.t. the host, re-entrant, keep no global
 *     mutable state apart from a thread-local last-error string;
 *   - return value: MF_OK (0) or a negative MF_E* code; mf_last_error() returns a
 *     human readable message for the calling thread.  No exception crosses the ABI;
 *   - four-vectors are (E, px, py, pz), array-of-structures, row-major contiguous,
 *     exactly as the reference torch tensors are laid out.
 */
#ifndef MEMFLOW_B200_H
#define MEMFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_ABI_VERSION 1

/* error codes */
#define MF_OK 0
#define MF_EINVAL -1   /* bad argument (null pointer, n out of range, unsupported shape) */
#define MF_ECUDA -2    /* CUDA runtime error (launch failure, ...) */
#define MF_EUNSUPPORTED -3

/* dtypes of device buffers */
#define MF_F32 0
#define MF_F64 1

/* arithmetic used inside the RAMBO kernels */
#define MF_COMPUTE_F64 0 /* reference arithmetic (the reference works in fp64), parity mode */
#define MF_COMPUTE_F32 1 /* cancellation-free fp32 formulation, throughput mode */

#define MF_MAX_FINAL 16 /* largest supported final-state multiplicity */

/* reference quirks that can be reproduced bit-for-bit on request */
#define MF_QUIRK_F32_VOLUME 1 /* get_flatWeights casts E_cm to fp32 (rambo_generator.py:132):
                                 (E_cm^2)^(n-2) overflows to +inf for n=8 above ~1.6 TeV */

/* per-event anomaly flags (optional int32 buffer, one word per event) */
#define MF_FLAG_NAN_INPUT 1      /* NaN in the uniforms (reference raises, rambo_generator.py:191-199) */
#define MF_FLAG_NOT_CM 2         /* |sum p|^2 >= 1e-3 in the inverse (reference raises, :630-635) */
#define MF_FLAG_BELOW_THRESHOLD 4 /* E_cm <= sum of masses */

/*
 * Process constants of one FlatInvertiblePhasespace instance.
 * The reference evaluates every mass-only quantity in fp32 because masses_t is an
 * fp32 tensor (rambo_generator.py:57-58); the host shim computes them with the same
 * torch fp32 ops and hands the (fp32-representable) results over as doubles so that
 * device code and oracle reproduce the reference exactly.
 */
typedef struct mf_rambo_desc {
  int32_t n_final;  /* n, 3..MF_MAX_FINAL (the reference fails below 3) */
  int32_t quirks;   /* MF_QUIRK_* */
  double sqrt_s;    /* collider_energy */
  double mass[MF_MAX_FINAL];         /* masses_t[i] */
  double mass_sq[MF_MAX_FINAL];      /* float32(masses_t[i]**2), set_square_t argument (:337,339) */
  double tail_sum_fwd[MF_MAX_FINAL]; /* flip(cumsum(flip(masses_t)))[i] (:484-486) */
  double tail_sum_inv[MF_MAX_FINAL]; /* torch.sum(masses_t[i:]) (:660) */
  double sum_mass;                   /* tot_final_state_masses (:58) */
  double x_q;    /* q = -log((sum_mass/sqrt_s)**2), fp32 (phasespace/utils.py:266-268) */
  double x_A;    /* A = (m/2) q^2 + q^2 with m=-1, fp32 (:269) */
  double x_c1;   /* float32(-q/A) (:272) */
  double x_c1sq; /* float32((-q/A)**2) */
  double x_c2;   /* float32(m/A) = float32(-1/A) */
  double last_plus_sq;  /* float32((m[n-1]+m[n-2])**2) (rambo_generator.py:489-493) */
  double last_minus_sq; /* float32((m[n-1]-m[n-2])**2) */
  double vol_coef; /* (2 pi)^(4-3n) (pi/2)^(n-1) (:134-135) */
  double vol_fact; /* (n-1)! (n-2)! (:138) */
  double pt_min_cut;   /* <=0: disabled.  (:352-365) */
  double delr_min_cut; /* <=0: disabled.  (:367-381) */
  double rap_max_cut;  /* <=0: disabled.  (:383-391) */
} mf_rambo_desc;

/* ------------------------------------------------------------------------- */
/* library                                                                    */
int mf_abi_version(void);
const char* mf_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long mf_launch_count(void);

/* ------------------------------------------------------------------------- */
/* RAMBO on diet                                                              */

/*
 * Forward map, replaces FlatInvertiblePhasespace.generateKinematics_batch
 * (rambo_generator.py:168-398) as called by PhaseSpace.get_momenta_from_ps
 * (phasespace.py:82-103).
 *   d_r        [N, 3n-2]   uniforms: [0:n-2] masses, [n-2:3n-4] (cos theta, phi) pairs, [-2:] x1x2
 *   d_momenta  [N, n+2, 4] CM-frame momenta, rows 0,1 incoming            (may be NULL)
 *   d_weight   [N]         Jacobian weight (times 0/1 for cuts)           (may be NULL)
 *   d_logw     [N]         log of the un-cut weight, finite where d_weight overflows (may be NULL)
 *   d_x1,d_x2  [N]                                                        (may be NULL)
 *   d_flags    [N] int32   MF_FLAG_*                                      (may be NULL)
 * io_dtype selects fp32 or fp64 for ALL floating buffers of the call.
 */
int mf_rambo_forward(const mf_rambo_desc* desc, const void* d_r, int64_t n_events,
                     void* d_momenta, void* d_weight, void* d_logw, void* d_x1, void* d_x2,
                     int32_t* d_flags, int io_dtype, int compute, void* stream);
int mf_rambo_forward_f32(const mf_rambo_desc* desc, const float* d_r, int64_t n_events,
                         float* d_momenta, float* d_weight, float* d_logw, float* d_x1,
                         float* d_x2, int32_t* d_flags, void* stream);
int mf_rambo_forward_f64(const mf_rambo_desc* desc, const double* d_r, int64_t n_events,
                         double* d_momenta, double* d_weight, double* d_logw, double* d_x1,
                         double* d_x2, int32_t* d_flags, void* stream);

/*
 * Inverse map, replaces getPSpoint_batch (rambo_generator.py:615-736) as called by
 * PhaseSpace.get_ps_from_momenta (phasespace.py:105-109).
 *   d_momenta  [N, n, 4]  final-state momenta in the CM frame
 *   d_x1,d_x2  [N]
 *   perm       [n] host   `order` argument (NULL = identity); desc must hold the PERMUTED masses
 *   d_r        [N, 3n-2]  uniforms
 *   d_detjacinv[N]        1/weight                                        (may be NULL)
 *   d_flags    [N] int32  MF_FLAG_NOT_CM, ...                             (may be NULL)
 */
int mf_rambo_inverse(const mf_rambo_desc* desc, const void* d_momenta, const void* d_x1,
                     const void* d_x2, const int32_t* perm, int64_t n_events, void* d_r,
                     void* d_detjacinv, int32_t* d_flags, int io_dtype, int compute,
                     void* stream);
int mf_rambo_inverse_f32(const mf_rambo_desc* desc, const float* d_momenta, const float* d_x1,
                         const float* d_x2, const int32_t* perm, int64_t n_events, float* d_r,
                         float* d_detjacinv, int32_t* d_flags, void* stream);
int mf_rambo_inverse_f64(const mf_rambo_desc* desc, const double* d_momenta, const double* d_x1,
                         const double* d_x2, const int32_t* perm, int64_t n_events, double* d_r,
                         double* d_detjacinv, int32_t* d_flags, void* stream);

/*
 * Gradient of the forward map w.r.t. its uniforms: what torch autograd does for
 * generateKinematics_batch(..., requires_grad=True) (rambo_generator.py:168-398; call sites
 * scripts/run_model_labframe_sampling.py:432,709).  fp64 (the reference runs this path in fp64).
 *   d_r             [N, 3n-2]   the uniforms of the forward call
 *   d_grad_momenta  [N, n+2, 4] cotangent of the momenta                       (may be NULL)
 *   d_grad_weight   [N]         cotangent of the weight                        (may be NULL)
 *   d_grad_x1/x2    [N]         cotangents of x1, x2                           (may be NULL)
 *   d_weight_fwd    [N]         the forward call's weight: where it is 0 (cut) the weight cotangent is dropped (may be NULL)
 *   d_grad_r        [N, 3n-2]   out
 * full_derivative = 0 reproduces the REFERENCE's autograd graph (no gradient through the bisection root, the
 * intermediate masses do not see x1/x2 -- rambo_generator.py:264-284, 400-446); 1 = the full derivative of the map.
 */
int mf_rambo_forward_backward(const mf_rambo_desc* desc, const double* d_r, int64_t n_events,
                              const double* d_grad_momenta, const double* d_grad_weight,
                              const double* d_grad_x1, const double* d_grad_x2, const double* d_weight_fwd,
                              int32_t full_derivative, double* d_grad_r, void* stream);

/*
 * Compute_ParticlesTensor.get_PS (unfolding_flow/utils.py:1056-1152): n=4 inverse
 * without Jacobian; x1,x2 from the boost four-vector; bool problem mask.
 *   d_momenta [N,4,4], d_boost [N,4] (E,px,py,pz of the total system in the lab)
 *   d_ps      [N,10]   cat(r, r_x1x2)
 *   d_mask    [N] uint8 (torch.bool) mask_problems
 * desc->sqrt_s is E_CM, desc masses are target_mass[perm].
 */
int mf_rambo_get_ps(const mf_rambo_desc* desc, const void* d_momenta, const void* d_boost,
                    const int32_t* perm, int64_t n_events, void* d_ps, uint8_t* d_mask,
                    int io_dtype, int compute, void* stream);

/*
 * Lab-frame get_PS: memflow/unfolding_flow/unfolding_flow.py:137-163 in ONE launch:
 *   1. fix an unphysical regressed boost (:140-143): where sqrt(E^2 - pz^2) < m_min_tot the energy becomes
 *      sqrt(pz^2 + m_min_tot^2 + 1e-3)                                        (m_min_tot <= 0: skipped)
 *   2. momenta_cm = boost_tt(momenta_lab, -boostVector_t(boost))               (phasespace/utils.py:36-42, 121-147)
 *   3. get_PS(momenta_cm, boost) as mf_rambo_get_ps
 *   d_momenta_lab [N,4,4], d_boost [N,4] (E,px,py,pz)
 *   d_momenta_cm [N,4,4] (may be NULL), d_boost_fixed [N,4] (may be NULL), d_ps [N,10], d_mask [N] (may be NULL)
 */
int mf_rambo_get_ps_lab(const mf_rambo_desc* desc, const void* d_momenta_lab, const void* d_boost,
                        const int32_t* perm, int64_t n_events, double m_min_tot, void* d_momenta_cm,
                        void* d_boost_fixed, void* d_ps, uint8_t* d_mask, int io_dtype, int compute, void* stream);

/* ------------------------------------------------------------------------- */
/* conditional RQ-spline masked autoregressive flow (zuko NSF / MAF family)  */

#define MF_MAX_HIDDEN_LAYERS 8
#define MF_MAX_FEATURES 64
#define MF_MAX_TRANSFORMS 64

/* univariate transformation of every feature (zuko.transforms) */
#define MF_UNIV_RQS 0          /* MonotonicRQSTransform(*phi, bound) */
#define MF_UNIV_CIRCULAR_RQS 1 /* CircularShiftTransform(bound) then RQS: NCSF, periodicNSF_gaussian.py:15-21 */
#define MF_UNIV_PS_RQS 2       /* PSTransform ((x+1)/2, ladj log 1/2) then RQS(bound=1): ttH/transfer_flow/custom_flows.py:10-48 */

/* base distribution (zuko.distributions) */
#define MF_BASE_DIAG_NORMAL 0 /* base_a = loc, base_b = scale */
#define MF_BASE_BOX_UNIFORM 1 /* base_a = lower, base_b = upper; log_prob = -inf outside [lower, upper) */

/* arithmetic of the conditioner (hyper-network) GEMMs */
#define MF_GEMM_EXACT 0  /* CUDA-core FMA in the I/O dtype (fp32 or fp64): bit-for-bit class of torch fp32 */
#define MF_GEMM_TC_3XF16 1 /* tcgen05, fp16 hi/lo split operands, 3 MMAs per product, fp32 accumulate */
#define MF_GEMM_TC_BF16 2  /* tcgen05, bf16 operands, fp32 accumulate: throughput mode */
#define MF_GEMM_MMA_3XF16 3 /* warp-level mma.sync, fp16 hi/lo split, 3 MMAs per product: hidden widths 129..512 (fp32) */

/*
 * Static description of one flow = zuko.flows.NSF(features, context, transforms, bins,
 * hidden_features, passes, ...) and its local subclasses (a13 in SURVEY.md section 8).
 * All transforms share one shape; masks are already folded into the packed weights.
 */
typedef struct mf_flow_desc {
  int32_t features;   /* D */
  int32_t context;    /* C */
  int32_t bins;       /* K: 3K-1 spline parameters per feature */
  int32_t transforms; /* T */
  int32_t n_hidden;   /* L hidden layers */
  int32_t hidden[MF_MAX_HIDDEN_LAYERS];
  int32_t passes;     /* AutoregressiveTransform.passes (fixed-point sweeps of the inverse) */
  int32_t univariate; /* MF_UNIV_* */
  int32_t base;       /* MF_BASE_* */
  int32_t ctx_group;  /* consecutive x rows sharing ONE context row (context hoisted per event in the MEM
                         sweep, BASELINE config 5); 0 or 1 = one context row per x row */
  double bound;       /* spline domain [-bound, bound] */
  double slope;       /* soft-clip slope of MonotonicRQSTransform (1e-3 in zuko >= 0.2); <= 0 disables */
  double base_a[MF_MAX_FEATURES];
  double base_b[MF_MAX_FEATURES];
} mf_flow_desc;

/* bytes of the packed-weight blob for (desc, dtype, gemm mode) */
int64_t mf_flow_blob_bytes(const mf_flow_desc* desc, int dtype, int gemm);

/*
 * Pack zuko-layout parameters into the kernel blob (mask folded in, transposed, padded).
 *   d_weights[t*(L+1)+l]  device pointer to MaskedLinear.weight [out_l, in_l] (dtype)
 *   d_biases [t*(L+1)+l]  device pointer to MaskedLinear.bias   [out_l]
 *   d_masks  [t*(L+1)+l]  device pointer to MaskedLinear.mask   [out_l, in_l] (uint8 / torch.bool)
 * The three pointer tables live on the HOST.  Replaces the per-call `mask * weight` of
 * zuko.nn.MaskedLinear.forward.
 */
int mf_flow_pack_weights(const mf_flow_desc* desc, const void* const* d_weights,
                         const void* const* d_biases, const uint8_t* const* d_masks, int dtype,
                         int gemm, void* d_blob, int64_t blob_bytes, void* stream);

/*
 * flow(c).log_prob(x) / flow(c).transform(x)  (a14, a15).
 *   d_x [N,D], d_ctx [N,C] (NULL when C = 0)
 *   d_logprob [N]  base.log_prob(z) + ladj           (may be NULL)
 *   d_z [N,D], d_ladj [N]                             (may be NULL)
 *   d_bins [T,N,D] int8 bin index of every spline evaluation, d_inside [T,N,D] uint8 in-range mask
 *                                                     (parity instrumentation, may be NULL)
 */
int mf_flow_log_prob(const mf_flow_desc* desc, const void* d_blob, const void* d_x,
                     const void* d_ctx, int64_t n, void* d_logprob, void* d_z, void* d_ladj,
                     int8_t* d_bins, uint8_t* d_inside, int dtype, int gemm, void* d_workspace,
                     int64_t workspace_bytes, void* stream);

/*
 * flow(c).sample() / .rsample()  (a16): x = T^-1(z), z = base sample built from caller-supplied
 * noise (standard normals for DiagNormal, uniforms in [0,1) for BoxUniform) so that reference and
 * kernel can be fed identical noise.
 *   d_noise [N,D] -> d_x [N,D];  d_ctx [N,C]
 */
int mf_flow_sample(const mf_flow_desc* desc, const void* d_blob, const void* d_noise,
                   const void* d_ctx, int64_t n, void* d_x, int dtype, int gemm, void* d_workspace,
                   int64_t workspace_bytes, void* stream);

/*
 * Element-wise maps of the flow's data side, fused into the kernels' load / store of x (a19 in SURVEY.md section 8:
 * the logit / sigmoid bridge between the unit hypercube of the phase space and the flow's unbounded space):
 *   MF_XMAP_SIGMOID_AFFINE  on the OUTPUT of mf_flow_sample:   x' = sigmoid(x * scale[f] + shift[f])
 *                           (scripts/run_model_labframe_sampling.py:427: torch.sigmoid(flow_samples*scale_ps + mean_ps))
 *   MF_XMAP_LOGIT_AFFINE    on the INPUT of mf_flow_log_prob:  x' = (logit(clamp(x, eps, 1-eps)) - shift[f]) / scale[f]
 *                           (memflow/unfolding_flow/unfolding_flow.py:170-171: torch.logit(ps, eps=5e-5), then the
 *                           (x - mean) / std scaling).  Like the reference, the map is a data transformation: it does not
 *                           contribute to ladj / log_prob.
 * A NULL map (or kind MF_XMAP_NONE) leaves the call identical to the unmapped entry points.
 */
#define MF_XMAP_NONE 0
#define MF_XMAP_SIGMOID_AFFINE 1
#define MF_XMAP_LOGIT_AFFINE 2
typedef struct mf_flow_xmap {
  int32_t kind; /* MF_XMAP_* */
  int32_t reserved;
  double eps;   /* clamp of the logit (torch.logit's eps); unused by the sigmoid */
  double scale[MF_MAX_FEATURES];
  double shift[MF_MAX_FEATURES];
} mf_flow_xmap;
int mf_flow_log_prob_mapped(const mf_flow_desc* desc, const void* d_blob, const void* d_x,
                            const void* d_ctx, int64_t n, void* d_logprob, void* d_z, void* d_ladj,
                            int8_t* d_bins, uint8_t* d_inside, int dtype, int gemm, void* d_workspace,
                            int64_t workspace_bytes, const mf_flow_xmap* in_map, void* stream);
int mf_flow_sample_mapped(const mf_flow_desc* desc, const void* d_blob, const void* d_noise,
                          const void* d_ctx, int64_t n, void* d_x, int dtype, int gemm, void* d_workspace,
                          int64_t workspace_bytes, const mf_flow_xmap* out_map, void* stream);

/*
 * Scratch the two calls above need for (desc, n, dtype, gemm); op 0 = log_prob, 1 = sample.  The caller
 * allocates it (the ABI never allocates) and may reuse it across calls on one stream.  0 = none needed.
 */
#define MF_FLOW_OP_LOG_PROB 0
#define MF_FLOW_OP_SAMPLE 1
int64_t mf_flow_workspace_bytes(const mf_flow_desc* desc, int64_t n, int dtype, int gemm, int op);

/*
 * Per-event likelihood tail: out[e] = sum_p logprob[e,p] * mask[e,p] - logw[e]
 * (transfer_flow_paper_AllPartons_btag.py:216; notebooks/TestFlows_ImportanceSampling.ipynb cell 52).
 *   d_logprob [B,P]; d_mask [B,P] uint8 (NULL = all objects exist); d_logw [B] (NULL = none)
 */
int mf_event_reduce(const void* d_logprob, const uint8_t* d_mask, const void* d_logw, int64_t n_events,
                    int32_t n_objects, void* d_out, int dtype, void* stream);

/*
 * Standalone zuko MonotonicRQSTransform on materialised parameters (a17): forward + log|dy/dx| or inverse of n
 * independent scalars.
 *   d_params [n, 3K-1]  (K widths | K heights | K-1 derivatives), unconstrained hyper-network outputs
 *   d_x [n] -> d_y [n], d_ladj [n] (NULL for the inverse), d_bins [n] int8 / d_inside [n] uint8 (may be NULL)
 *   univariate: MF_UNIV_* (circular shift / PS pre-map applied around the spline like in the flows)
 */
int mf_rqs_transform(const void* d_params, const void* d_x, int64_t n, int32_t bins, int32_t univariate, double bound,
                     double slope, int32_t inverse, void* d_y, void* d_ladj, int8_t* d_bins, uint8_t* d_inside,
                     int dtype, void* stream);

/* Sampling schedule.  zuko inverts an autoregressive transform with `passes` full sweeps of the hyper-network; the
 * features of order group s are final after sweep s and every later sweep recomputes them to the same value.  Given the
 * effective orders (zuko `MaskedAutoregressiveTransform.order`, the buffer `transforms.{t}.order` of a checkpoint,
 * values 0..passes-1) as a HOST array [transforms][features], this marks per sweep the features that become final
 * there; mf_flow_sample then evaluates only those splines and only the last-layer chunks that hold one of them.  Optional: after
 * mf_flow_pack_weights every chunk is evaluated in every sweep (the literal algorithm); results are the same.
 * Synchronises the stream.
 */
int mf_flow_set_sample_schedule(const mf_flow_desc* desc, const int32_t* h_orders, int dtype, int gemm, void* d_blob,
                                int64_t blob_bytes, void* stream);
/* Byte offset of the schedule inside the blob: MF_MAX_TRANSFORMS x MF_MAX_FEATURES uint64 words, word [t][s] = bitmask of
 * the features of transform t that become final in sweep s.  A caller that re-packs every training step can keep the
 * words on the device and copy them in place instead of calling mf_flow_set_sample_schedule (which synchronises). */
int64_t mf_flow_sched_offset(const mf_flow_desc* desc, int dtype, int gemm);

/*
 * C[m, n] = A[m, k] * B[n, k]^T (+ bias[n]) (ReLU) in fp32-class arithmetic on the tensor cores (fp16 hi/lo split
 * operands, three MMAs per product, fp32 accumulate): the GEMMs of the hyper-network backward of the training step,
 * which torch autograd runs as cuBLAS fp32 SIMT kernels in the reference (scripts/run_model_labframe_sampling.py:499-501,
 * scripts/run_transferFlow_paperVersion_AllPartons_btag.py:385).  Row-major fp32 operands with leading dimensions;
 * n <= 256.  k_split > 1: the K range is cut into that many pieces whose partial sums are ADDED into C (the caller
 * zeroes C; no ReLU).  d_mask (optional, shaped like A, leading dimension ldm): A is multiplied by (mask > 0), the
 * derivative of the ReLU whose output is `mask`.  d_out_mask (optional, shaped like C, leading dimension ldo; k_split = 1):
 * C is multiplied by (out_mask > 0) -- the input gradient of a layer fed by a ReLU leaves already masked for the
 * layer below.  d_scale_a / d_scale_b: optional device scalars (powers of two) the
 * operands are multiplied by before the fp16 split -- gradients are far below the fp16 range --; the result is
 * divided by their product.
 */
int mf_gemm_nt_3xf16(const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                     int64_t m, int32_t n, int64_t k, const float* d_bias, int32_t relu, int32_t k_split,
                     const float* d_mask, int64_t ldm, const float* d_out_mask, int64_t ldo,
                     const float* d_scale_a, const float* d_scale_b, void* stream);
/*
 * C[m, n] (+)= A[k, m]^T * B[k, n]: the weight gradient dW = dZ^T H of a linear layer straight from the row-major
 * activations (k = rows; A = dY [rows, out] with the optional ReLU mask, B = H [rows, in]); same arithmetic and
 * k_split / scale conventions.  d_colsum (optional, [m]): += sum_k A[k, m], the bias gradient (zeroed by the caller).
 */
int mf_gemm_tn_3xf16(const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                     int64_t m, int32_t n, int64_t k, int32_t k_split, const float* d_mask, int64_t ldm,
                     float* d_colsum, const float* d_scale_a, const float* d_scale_b, void* stream);

/* Backward of the forward spline of mf_rqs_transform (plain MF_UNIV_RQS): gradients of a scalar loss w.r.t. the packed
 * parameters [n, 3K-1] and w.r.t. x [n], given dL/dy [n] and dL/dladj [n] (d_grad_ladj may be NULL = 0).
 * Replaces what torch autograd derives from zuko MonotonicRQSTransform.call_and_ladj in the reference's training steps
 * (`loss.backward()` after `flow(c).log_prob(x)`, e.g. scripts/run_model_labframe_sampling.py:466-520).  The inverse
 * direction follows by the implicit function theorem on the same entry point (flows/_autograd.py).
 */
int mf_rqs_backward(const void* d_params, const void* d_x, int64_t n, int32_t bins, double bound, double slope,
                    const void* d_grad_y, const void* d_grad_ladj, void* d_grad_params, void* d_grad_x, int dtype,
                    void* stream);

/*
 * Empirical maximum mean discrepancy between two samples x, y [n, d] (row-major, `dtype`), SURVEY 8(f)-4:
 * replaces memflow/unfolding_flow/mmd_loss.py:4-41 `MMD(x, y, kernel, device, dtype)` (three torch.mm Gram matrices
 * + one n x n accumulator per bandwidth) and, when a gradient pointer is given, what torch autograd derives from it in
 * the reference's `loss.backward()` (scripts/run_model_labframe_sampling.py:114-127,466-520).
 *   kernel: MF_MMD_MULTISCALE (bandwidths 0.1, 0.2, 0.5, 0.9, 1.3) or MF_MMD_RBF (10, 15, 20, 50)
 *   d_out [1] = mean(XX + YY - 2 XY); d_grad_x / d_grad_y [n, d] = d out / d x, d out / d y (either may be NULL)
 *   d_workspace: mf_mmd_workspace_bytes(n, d, want_grad) bytes of scratch (the ABI never allocates)
 * Two launches (pair sums, fixed-order finish); no n x n intermediate; results do not depend on scheduling.
 */
#define MF_MMD_MULTISCALE 0
#define MF_MMD_RBF 1
#define MF_MMD_MAX_DIM 32
int64_t mf_mmd_workspace_bytes(int64_t n, int32_t d, int32_t want_grad);
int mf_mmd(const void* d_x, const void* d_y, int64_t n, int32_t d, int32_t kernel, void* d_workspace, void* d_out,
           void* d_grad_x, void* d_grad_y, int dtype, void* stream);

/*
 * Regression loss of a periodic variable (phi), SURVEY 8(f)-4: replaces `loss_fn_periodic(inp, target, loss_fn, device)`
 * (scripts/run_model_labframe_sampling.py:65-77) for loss_fn = torch.nn.HuberLoss(delta) / torch.nn.MSELoss.
 * inp is wrapped once into [-pi, pi], dphi = target - inp is wrapped into (-pi, pi], loss_i = loss_fn(dphi, 0).
 *   d_inp, d_target [n]; kind: MF_LOSS_HUBER (huber_delta > 0) or MF_LOSS_MSE
 *   d_elem [n] (optional): loss_i (reduction='none'); d_deriv [n] (optional): d loss_i / d dphi (= d/d target = -d/d inp)
 *   d_total [1] (optional): sum_i loss_i, divided by n when mean != 0
 *   d_workspace: mf_periodic_loss_workspace_bytes(n) bytes
 */
#define MF_LOSS_HUBER 0
#define MF_LOSS_MSE 1
int64_t mf_periodic_loss_workspace_bytes(int64_t n);
int mf_periodic_loss(const void* d_inp, const void* d_target, int64_t n, int32_t kind, double huber_delta, int32_t mean,
                     void* d_workspace, void* d_elem, void* d_deriv, void* d_total, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEMFLOW_B200_H */
