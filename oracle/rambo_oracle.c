/*
 * rambo_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU, fp64 restatement of the reference's RAMBO-on-diet phase-space map
 * (themrluke/memflow-msci-project).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call into this file; the
 * product path (memflow_msci_project_b200/) never does.
 *
 * Every function cites the reference lines it follows (paths relative to the
 * reference tree).  The arithmetic is kept in the reference's own order of
 * operations, including its quirks:
 *   - mass-only constants arrive pre-evaluated in fp32 inside mf_rambo_desc
 *     (the reference's masses_t is an fp32 tensor, rambo_generator.py:57-58);
 *   - get_flatWeights works in fp32 (rambo_generator.py:132) and the weight chain
 *     that follows is accumulated in place in an fp32 tensor (:489-507, :709-734);
 *   - the incoming momenta are built from float(E_cm) (:533);
 *   - the inverse stores its uniforms in an fp32 tensor (:663).
 * Pinning: checked element-wise against the imported reference (fed fp64 uniforms)
 * by tools/make_golden_rambo.py; the resulting vectors live in tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/memflow_b200.h"

#define PI 3.14159265358979323846

/* rambo_generator.py:145-149 */
static double rho(double M, double N, double m) {
  double Msqr = M * M;
  return sqrt((Msqr - (N + m) * (N + m)) * (Msqr - (N - m) * (N - m))) / (8.0 * Msqr);
}

/* rho with the squares of (N+m), (N-m) already rounded to fp32: rambo_generator.py:489-493
 * (both arguments are 0-dim fp32 tensors there) */
static double rho_presq(double M, double plus_sq, double minus_sq) {
  double Msqr = M * M;
  return sqrt((Msqr - plus_sq) * (Msqr - minus_sq)) / (8.0 * Msqr);
}

/* rambo_generator.py:142-143 */
static double massless_map(double x, double e) { return pow(x, e) * ((e + 1.0) - e * x); }

/* rambo_generator.py:400-446, one column; the reference runs rounds of 60 halvings
 * until the batch-wide error stops improving; past ~55 halvings nothing moves. */
static double bisect(double v, double e) {
  double left = 0.0, right = 1.0, u = -1.0;
  for (int level = 0; level < 120; ++level) {
    u = (left + right) * pow(0.5, level + 1);
    double check = massless_map(u, e);
    left *= 2.0;
    right *= 2.0;
    double adder = (v <= check) ? -0.5 : 0.5;
    left = left + (adder + 0.5);
    right = right + (adder - 0.5);
  }
  return u;
}

/* rambo_generator.py:111-140 (tensor branch: fp32) */
static float flat_weights_f32(double E_cm, int n, const mf_rambo_desc* d) {
  float E = (float)E_cm;
  float E2 = E * E;
  float p = powf(E2, (float)(n - 2));
  float frac = p / (float)d->vol_fact;
  return (float)d->vol_coef * frac;
}

/* phasespace/utils.py:6-22 */
static void set_square(double* p, double square) {
  p[0] = sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3] + square);
}

/* phasespace/utils.py:88-118, boost_vector = b[3] */
static void boost(double* p, const double* b) {
  double b2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
  double gamma = 1.0 / sqrt(1.0 - b2);
  double bp = p[1] * b[0] + p[2] * b[1] + p[3] * b[2];
  double gamma2 = (b2 > 0.0) ? (gamma - 1.0) / b2 : 0.0;
  double factor = gamma2 * bp + gamma * p[0];
  double E = p[0];
  p[1] += factor * b[0];
  p[2] += factor * b[1];
  p[3] += factor * b[2];
  p[0] = gamma * (E + bp);
}

/* phasespace/utils.py:205-217 (eps = sqrt(DBL_EPSILON), huge = DBL_MAX) */
static double pseudo_rap(const double* p) {
  const double eps = 1.4901161193847656e-08, huge = 1.7976931348623157e308;
  double pt = sqrt(p[1] * p[1] + p[2] * p[2]);
  if (pt < eps && fabs(p[3]) < eps) return huge;
  double th = atan2(pt, p[3]);
  return -log(tan(th / 2.0));
}

/* phasespace/utils.py:233-257 */
static double delta_r(const double* a, const double* b) {
  const double huge = 1.7976931348623157e308;
  double pt1 = sqrt(a[1] * a[1] + a[2] * a[2]);
  double pt2 = sqrt(b[1] * b[1] + b[2] * b[2]);
  double tmp = (a[1] * b[1] + a[2] * b[2]) / (pt1 * pt2);
  double dphi = (fabs(tmp) > 1.0) ? acos(tmp / fabs(tmp)) : acos(tmp);
  if (pt1 == 0.0 || pt2 == 0.0) dphi = huge;
  double deta = pseudo_rap(a) - pseudo_rap(b);
  return sqrt(deta * deta + dphi * dphi);
}

/*
 * generateKinematics_batch, rambo_generator.py:168-398 (pdf_active, uniform_x1x2).
 * r[3n-2] -> mom[(n+2)*4], weight, x1, x2.  Returns MF_FLAG_* bits.
 */
static int forward_one(const mf_rambo_desc* d, const double* r, double* mom, double* weight_out,
                       double* logw_out, double* x1_out, double* x2_out) {
  const int n = d->n_final;
  int flags = 0;
  for (int c = 0; c < 3 * n - 2; ++c)
    if (isnan(r[c])) flags |= MF_FLAG_NAN_INPUT; /* :191-199 */

  /* get_x1x2_from_uniform, phasespace/utils.py:260-284 */
  const double ua = r[3 * n - 4], ub = r[3 * n - 3];
  double mlogx1 = (d->x_c1 + sqrt(d->x_c1sq + (2.0 * -1.0 * ua) / d->x_A)) / d->x_c2;
  double maxv = -1.0 * mlogx1 + d->x_q; /* uniform_distr(unif[:,1], 0, m*minus_logx1+q) */
  double mlogx2 = 0.0 + (maxv - 0.0) * ub;
  double x1 = exp(-mlogx1), x2 = exp(-mlogx2);
  double detjacinv = 1.0 / (d->x_A * x1 * x2);
  double wgt_x = 1.0 / detjacinv;
  double E_cm = sqrt(x1 * x2) * d->sqrt_s; /* rambo_generator.py:226 */

  /* generateIntermediatesMassive/Massless_batch, :448-509 */
  double K[MF_MAX_FINAL], M[MF_MAX_FINAL];
  double K0 = E_cm - d->sum_mass; /* :478 */
  if (!(K0 > 0.0)) flags |= MF_FLAG_BELOW_THRESHOLD;
  double prod_u = 1.0;
  K[0] = K0 * prod_u;
  for (int c = 0; c < n - 2; ++c) { /* cumprod of [1, u...] (:455-458) */
    double u = bisect(r[c], (double)(n - 2 - c));
    prod_u *= u;
    K[c + 1] = K0 * prod_u;
  }
  for (int i = 0; i < n - 1; ++i) M[i] = K[i] + d->tail_sum_fwd[i]; /* :487 */

  float wf = flat_weights_f32(E_cm, n, d); /* fp32 tensor, updated in place (:473,:489-507) */
  double logw = log(d->vol_coef) + (double)(n - 2) * log(E_cm * E_cm) - log(d->vol_fact);
  {
    double t = 8.0 * rho_presq(M[n - 2], d->last_plus_sq, d->last_minus_sq);
    wf = (float)((double)wf * t);
    logw += log(t);
    double prod = 1.0;
    for (int i = 0; i < n - 2; ++i) {
      double term =
          (rho(M[i], M[i + 1], d->mass[i]) / rho(K[i], K[i + 1], 0.0)) * (M[i + 1] / K[i + 1]);
      prod *= term;
    }
    wf = (float)((double)wf * prod);
    logw += log(prod);
    double pw = pow(K[0] / M[0], (double)(2 * n - 4));
    wf = (float)((double)wf * pw);
    logw += log(pw);
  }
  double weight = wgt_x * (double)wf; /* :259,:291 */
  logw += log(wgt_x);
  if (!(d->quirks & MF_QUIRK_F32_VOLUME)) weight = exp(logw);

  /* decay chain, :294-345 */
  M[n - 1] = d->mass[n - 1]; /* :303-309 */
  double Q[4] = {M[0], 0.0, 0.0, 0.0};
  const double* rnd = r + (n - 2);
  for (int i = 0; i < n - 1; ++i) {
    double q = 4.0 * M[i] * rho(M[i], M[i + 1], d->mass[i]); /* :310-314 */
    double cos_t = 2.0 * rnd[2 * i] - 1.0;
    double sin_t = sqrt(1.0 - cos_t * cos_t);
    double phi = 2.0 * PI * rnd[2 * i + 1];
    double cos_p = cos(phi);
    double s = sqrt(1.0 - cos_p * cos_p);
    double sin_p = (phi > PI) ? -s : s;
    double p2[4] = {0.0, q * sin_t * cos_p, q * sin_t * sin_p, q * cos_t};
    set_square(p2, d->mass_sq[i]);
    double bv[3] = {Q[1] / Q[0], Q[2] / Q[0], Q[3] / Q[0]}; /* boostVector_t, utils.py:36-42 */
    boost(p2, bv);
    set_square(p2, d->mass_sq[i]);
    memcpy(mom + 4 * (i + 2), p2, sizeof p2);
    for (int k = 0; k < 4; ++k) Q[k] -= p2[k];
    set_square(Q, M[i + 1] * M[i + 1]);
  }
  memcpy(mom + 4 * (n + 1), Q, sizeof Q); /* :345 */

  /* setInitialStateMomenta_batch, massless beams, :533-551: E_cm is cast to fp32 */
  double Ef = (double)(float)E_cm;
  mom[0] = Ef / 2.0; mom[1] = 0.0; mom[2] = 0.0; mom[3] = Ef / 2.0;
  mom[4] = Ef / 2.0; mom[5] = 0.0; mom[6] = 0.0; mom[7] = -1.0 * Ef / 2.0;

  /* cuts in the lab frame, :350-393 (no-ops at the default -1) */
  double factor2 = 1.0;
  if (d->pt_min_cut > 0.0 || d->delr_min_cut > 0.0 || d->rap_max_cut > 0.0) {
    double lab[(MF_MAX_FINAL + 2) * 4];
    memcpy(lab, mom, sizeof(double) * 4 * (n + 2));
    /* boost_to_lab_frame, utils.py:180-202 */
    double ref[4];
    for (int k = 0; k < 4; ++k) ref[k] = mom[k] * x1 + mom[4 + k] * x2;
    double rr = ref[1] * ref[1] + ref[2] * ref[2] + ref[3] * ref[3];
    if (rr != 0.0 && (x1 != 1.0 || x2 != 1.0)) {
      double bv[3] = {ref[1] / ref[0], ref[2] / ref[0], ref[3] / ref[0]};
      for (int i = 0; i < n + 2; ++i) boost(lab + 4 * i, bv);
    }
    double ptmin = INFINITY, etamax = -INFINITY;
    for (int i = 2; i < n + 2; ++i) {
      double pt = fabs(sqrt(lab[4 * i + 1] * lab[4 * i + 1] + lab[4 * i + 2] * lab[4 * i + 2]));
      if (pt < ptmin) ptmin = pt;
      double eta = pseudo_rap(lab + 4 * i);
      if (eta > etamax) etamax = eta;
    }
    if (ptmin < d->pt_min_cut) factor2 = 0.0; /* :361-365 */
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j)
        if (fabs(delta_r(lab + 4 * (i + 2), lab + 4 * (j + 2))) < d->delr_min_cut)
          factor2 = 0.0; /* :367-381 */
    if (d->rap_max_cut > 0.0 && d->rap_max_cut < fabs(etamax)) factor2 = 0.0; /* :383-391 */
  }
  *weight_out = weight * factor2;
  *logw_out = logw;
  *x1_out = x1;
  *x2_out = x2;
  return flags;
}

/* Minkowski square, phasespace/utils.py:56-65 */
static double msquare(const double* p) {
  return p[0] * p[0] - p[1] * p[1] - p[2] * p[2] - p[3] * p[3];
}

/* shared by getPSpoint_batch (:656-696) and get_PS (unfolding_flow/utils.py:1083-1131):
 * P[n][4] (already permuted) -> r[3n-4] (rounded to fp32 as the reference's r tensor is),
 * M[n], K[n].  tail_sum = sums of the (permuted) masses from j on. */
static void inverse_core(int n, double P[][4], const double* tail_sum, double* r, double* M,
                         double* K) {
  double Q[MF_MAX_FINAL][4];
  for (int j = n - 1; j >= 0; --j) { /* :656-660 */
    double s[4] = {0, 0, 0, 0};
    for (int k = j; k < n; ++k)
      for (int c = 0; c < 4; ++c) s[c] += P[k][c];
    M[j] = sqrt(msquare(s));
    K[j] = M[j] - tail_sum[j];
  }
  memcpy(Q[n - 1], P[n - 1], sizeof Q[0]); /* :647 */
  for (int i = n; i > 1; --i) {            /* :665-696 */
    int j = i - 1;
    double u = K[j] / K[j - 1];
    r[j - 1] = (float)((n + 1 - i) * pow(u, (double)(n - i)) - (n - i) * pow(u, (double)(n + 1 - i)));
    for (int c = 0; c < 4; ++c) Q[j - 1][c] = Q[j][c] + P[j - 1][c];
    double bv[3] = {-1.0 * (Q[j - 1][1] / Q[j - 1][0]), -1.0 * (Q[j - 1][2] / Q[j - 1][0]),
                    -1.0 * (Q[j - 1][3] / Q[j - 1][0])};
    boost(P[j - 1], bv);
    const double* p = P[j - 1];
    r[n - 5 + 2 * i - 1] = (float)(((p[3] / sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3])) + 1.0) / 2.0);
    double phi = atan(p[2] / p[1]);
    double dphi = (p[2] < 0.0 && p[1] > 0.0) ? 2.0 * PI : 0.0;
    dphi += (p[1] < 0.0) ? PI : 0.0;
    phi += dphi;
    r[n - 4 + 2 * i - 1] = (float)(phi / (2.0 * PI));
  }
}

/* get_uniform_from_x1x2, phasespace/utils.py:287-309 */
static void uniform_from_x1x2(const mf_rambo_desc* d, double x1, double x2, double* u1, double* u2,
                              double* jac) {
  double mlogx1 = -log(x1), mlogx2 = -log(x2);
  *u1 = ((-1.0 / 2.0) * (mlogx1 * mlogx1) + d->x_q * mlogx1) / d->x_A;
  *u2 = mlogx2 / (-1.0 * mlogx1 + d->x_q);
  *jac = 1.0 / (d->x_A * x1 * x2);
}

/* getPSpoint_batch, rambo_generator.py:615-736.  mom[n*4] final-state momenta. */
static int inverse_one(const mf_rambo_desc* d, const int32_t* perm, const double* mom, double x1,
                       double x2, double* r_out, double* prob_out) {
  const int n = d->n_final;
  int flags = 0;
  double P[MF_MAX_FINAL][4], M[MF_MAX_FINAL], K[MF_MAX_FINAL];
  for (int i = 0; i < n; ++i) memcpy(P[i], mom + 4 * (perm ? perm[i] : i), sizeof P[0]); /* :624-625 */
  {
    double s[4] = {0, 0, 0, 0}; /* :630-635 */
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 4; ++c) s[c] += P[i][c];
    if (!(s[1] * s[1] + s[2] * s[2] + s[3] * s[3] < 1e-3)) flags |= MF_FLAG_NOT_CM;
  }
  inverse_core(n, P, d->tail_sum_inv, r_out, M, K);
  double u1, u2, jac;
  uniform_from_x1x2(d, x1, x2, &u1, &u2, &jac); /* :699 */
  r_out[3 * n - 4] = u1;
  r_out[3 * n - 3] = u2;
  double E_cm = sqrt(x1 * x2) * d->sqrt_s; /* :700 */
  float wf = flat_weights_f32(E_cm, n, d);  /* :707, fp32 in-place chain :709-727 */
  double logw = log(d->vol_coef) + (double)(n - 2) * log(E_cm * E_cm) - log(d->vol_fact);
  double t = 8.0 * rho_presq(M[n - 2], d->last_plus_sq, d->last_minus_sq);
  wf = (float)((double)wf * t);
  logw += log(t);
  double prod = 1.0;
  for (int i = 0; i < n - 2; ++i)
    prod *= (rho(M[i], M[i + 1], d->mass[i]) / rho(K[i], K[i + 1], 0.0)) * (M[i + 1] / K[i + 1]);
  wf = (float)((double)wf * prod);
  logw += log(prod);
  double pw = pow(K[0] / M[0], (double)(2 * n - 4));
  wf = (float)((double)wf * pw);
  logw += log(pw);
  if (d->quirks & MF_QUIRK_F32_VOLUME) {
    float prob = 1.0f / wf;              /* :732, fp32 */
    prob = (float)((double)prob * jac);  /* :734 */
    *prob_out = (double)prob;
  } else {
    *prob_out = exp(-logw) * jac;
  }
  return flags;
}

/* Compute_ParticlesTensor.get_PS, unfolding_flow/utils.py:1056-1152 (n = 4) */
static int get_ps_one(const mf_rambo_desc* d, const int32_t* perm, const double* mom,
                      const double* boost_reco, double* ps_out) {
  const int n = 4;
  double P[MF_MAX_FINAL][4], M[MF_MAX_FINAL], K[MF_MAX_FINAL];
  for (int i = 0; i < n; ++i) memcpy(P[i], mom + 4 * (perm ? perm[i] : i), sizeof P[0]);
  double x1 = (boost_reco[0] + boost_reco[3]) / d->sqrt_s; /* :1066-1067 */
  double x2 = (boost_reco[0] - boost_reco[3]) / d->sqrt_s;
  int mask = (x1 > 1.0) | (x1 < 0.0) | (x2 < 0.0) | (x1 > 1.0); /* :1069, sic */
  x1 = fmin(fmax(x1, 0.0), 1.0);
  x2 = fmin(fmax(x2, 0.0), 1.0);
  double s[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c) s[c] += P[i][c];
  mask |= (s[1] * s[1] + s[2] * s[2] + s[3] * s[3] > 1e-5); /* :1079-1081 */
  inverse_core(n, P, d->tail_sum_inv, ps_out, M, K);
  for (int c = 0; c < 8; ++c) mask |= (ps_out[c] < 0.0) | (ps_out[c] > 1.0); /* :1142 */
  double u1, u2, jac;
  uniform_from_x1x2(d, x1, x2, &u1, &u2, &jac); /* :1150 */
  ps_out[8] = u1;
  ps_out[9] = u2;
  return mask;
}

/* ----------------------------------------------------------------------------------- */
/* batch entry points (OpenMP over events; used single- or multi-threaded by the caller) */

void oracle_rambo_forward(const mf_rambo_desc* d, const double* r, int64_t N, double* mom,
                          double* weight, double* logw, double* x1, double* x2, int32_t* flags) {
  const int n = d->n_final;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < N; ++e) {
    int f = forward_one(d, r + e * (3 * n - 2), mom + e * 4 * (n + 2), weight + e, logw + e,
                        x1 + e, x2 + e);
    if (flags) flags[e] = f;
  }
}

void oracle_rambo_inverse(const mf_rambo_desc* d, const double* mom, const double* x1,
                          const double* x2, const int32_t* perm, int64_t N, double* r,
                          double* prob, int32_t* flags) {
  const int n = d->n_final;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < N; ++e) {
    int f = inverse_one(d, perm, mom + e * 4 * n, x1[e], x2[e], r + e * (3 * n - 2), prob + e);
    if (flags) flags[e] = f;
  }
}

void oracle_rambo_get_ps(const mf_rambo_desc* d, const double* mom, const double* boost_reco,
                         const int32_t* perm, int64_t N, double* ps, uint8_t* mask) {
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < N; ++e)
    mask[e] = (uint8_t)get_ps_one(d, perm, mom + e * 16, boost_reco + e * 4, ps + e * 10);
}

/* fp32-I/O convenience used by the CPU baseline leg (same arithmetic, fp32 buffers) */
void oracle_rambo_forward_f32(const mf_rambo_desc* d, const float* r, int64_t N, float* mom,
                              float* weight, float* x1, float* x2) {
  const int n = d->n_final;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < N; ++e) {
    double rd[3 * MF_MAX_FINAL], md[4 * (MF_MAX_FINAL + 2)], w, lw, a, b;
    for (int c = 0; c < 3 * n - 2; ++c) rd[c] = (double)r[e * (3 * n - 2) + c];
    forward_one(d, rd, md, &w, &lw, &a, &b);
    for (int c = 0; c < 4 * (n + 2); ++c) mom[e * 4 * (n + 2) + c] = (float)md[c];
    weight[e] = (float)w;
    x1[e] = (float)a;
    x2[e] = (float)b;
  }
}
