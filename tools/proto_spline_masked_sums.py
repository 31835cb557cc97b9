#!/usr/bin/env python
"""Numerical prototype of the "masked sums" spline formulation of csrc/spline_reg.cuh, in numpy fp32, against the
fp64 oracle (oracle/flow_oracle.py) and next to the oracle evaluated in numpy fp32 (the arithmetic class of the
reference's torch fp32 path).  Shows that the reformulation (unnormalised cumulative sums, bin found by comparing
against a rescaled input, bin quantities gathered with 0/1 masks, z = (vt - X0) / E) has the error distribution of
plain fp32 arithmetic or better, forward and inverse, for logit scales 0.3 / 1 / 3.

    python tools/proto_spline_masked_sums.py
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import flow_oracle as fo
f32 = np.float32
def softclip(x, inv): return x / (f32(1) + np.abs(x * inv))
def new_forward(w, h, d, v, bound, slope):
    # all fp32; w,h [N,K], d [N,K-1]
    K = w.shape[-1]
    ls = math.log(slope)
    inv2, inv1 = f32(2/ls), f32(1/ls)
    B = f32(bound)
    ew = np.exp(softclip(w, inv2)).astype(f32); eh = np.exp(softclip(h, inv2)).astype(f32)
    cw = np.zeros(w.shape[0], f32); ch = np.zeros_like(cw)
    cws = []; 
    for j in range(K):
        cw = cw + ew[:, j]; ch = ch + eh[:, j]; cws.append(cw.copy())
    sw, sh = cw, ch
    u = (v * f32(0.5 / bound) + f32(0.5))   # (v/B+1)/2
    vt = u * sw
    m_prev = (v > -B).astype(f32)
    x0u = np.zeros_like(cw); y0u = np.zeros_like(cw); ewk = np.zeros_like(cw); ehk = np.zeros_like(cw)
    t0 = np.zeros_like(cw); t1 = np.zeros_like(cw); kf = -1 + m_prev
    mlast = None
    for j in range(K):
        m = (cws[j] < vt).astype(f32)
        o = m_prev - m
        x0u = x0u + ew[:, j] * m; y0u = y0u + eh[:, j] * m
        ewk = ewk + ew[:, j] * o; ehk = ehk + eh[:, j] * o
        if j < K - 1: t1 = t1 + d[:, j] * o
        if j >= 1: t0 = t0 + d[:, j - 1] * o
        kf = kf + m
        m_prev = m
    inside = ((v > -B) & (m_prev == 0))
    isw = f32(1) / sw; ish = f32(1) / sh
    d0 = np.exp(softclip(t0, inv1)).astype(f32); d1 = np.exp(softclip(t1, inv1)).astype(f32)
    # z = (v - x0)/dx with x0 = B(2 x0u/sw - 1), dx = 2B ewk/sw  ->  z = (vt - x0u)/ewk   (all unnormalised!)
    with np.errstate(all='ignore'):
        z = np.where(inside, (vt - x0u) / ewk, f32(0)).astype(f32)
        s = (ehk * ish) / (ewk * isw)
        omz = f32(1) - z
        den = s + (d0 + d1 - f32(2) * s) * z * omz
        y0 = B * (f32(2) * (y0u * ish) - f32(1)); dy = f32(2) * B * (ehk * ish)
        y = y0 + dy * (s * z * z + d0 * z * omz) / den
        jac = s * s * (f32(2) * s * z * omz + d0 * omz * omz + d1 * z * z) / (den * den)
    out = np.where(inside, y, v).astype(f32)
    ladj = np.where(inside, np.log(jac), f32(0)).astype(f32)
    return out, ladj, kf.astype(np.int64), inside

rng = np.random.default_rng(0)
N, K = 200000, 16
for scale in (0.3, 1.0, 3.0):
    w = (rng.standard_normal((N, K)) * scale); h = rng.standard_normal((N, K)) * scale; d = rng.standard_normal((N, K - 1)) * scale
    bound = 5.0; slope = 1e-3
    v = rng.uniform(-5.5, 5.5, N)
    hor, ver, der = fo.rqs_knots(w, h, d, bound, slope)
    y64, l64, k64, m64 = fo.rqs_forward(v, hor, ver, der)
    hor32, ver32, der32 = fo.rqs_knots(w.astype(f32), h.astype(f32), d.astype(f32), bound, slope)
    y32, l32, k32, m32 = fo.rqs_forward(v.astype(f32), hor32, ver32, der32)
    yn, ln, kn, mn = new_forward(w.astype(f32), h.astype(f32), d.astype(f32), v.astype(f32), bound, slope)
    def stats(a, b): 
        e = np.abs(a.astype(np.float64) - b); return f"med {np.median(e):.2e} q999 {np.quantile(e,0.999):.2e} max {e.max():.2e}"
    print(f"scale {scale}: numpy-fp32 ref  y: {stats(y32,y64)}  ladj: {stats(l32,l64)}  binmis {np.sum((k32!=k64)&m64)} maskmis {np.sum(m32!=m64)}")
    print(f"           new form        y: {stats(yn,y64)}  ladj: {stats(ln,l64)}  binmis {np.sum((kn!=k64)&m64)} maskmis {np.sum(mn!=m64)}")

def new_inverse(w, h, d, y, bound, slope):
    K = w.shape[-1]
    ls = math.log(slope); inv2, inv1 = f32(2/ls), f32(1/ls); B = f32(bound)
    ew = np.exp(softclip(w, inv2)).astype(f32); eh = np.exp(softclip(h, inv2)).astype(f32)
    sw = np.zeros(w.shape[0], f32); sh = np.zeros_like(sw)
    for j in range(K): sw = sw + ew[:, j]; sh = sh + eh[:, j]
    u = (y * f32(0.5 / bound) + f32(0.5)); vt = u * sh
    m_prev = (y > -B).astype(f32)
    x0u = np.zeros_like(sw); y0u = np.zeros_like(sw); ewk = np.zeros_like(sw); ehk = np.zeros_like(sw)
    t0 = np.zeros_like(sw); t1 = np.zeros_like(sw); cw = np.zeros_like(sw); ch = np.zeros_like(sw); kf = np.zeros_like(sw)
    for j in range(K):
        cw = cw + ew[:, j]; ch = ch + eh[:, j]
        m = (ch < vt).astype(f32); o = m_prev - m
        x0u = x0u + ew[:, j] * m; y0u = y0u + eh[:, j] * m
        ewk = ewk + ew[:, j] * o; ehk = ehk + eh[:, j] * o
        if j < K - 1: t1 = t1 + d[:, j] * o
        if j >= 1: t0 = t0 + d[:, j - 1] * o
        kf = kf + f32(j) * o
        m_prev = m
    inside = ((y > -B) & (m_prev == 0))
    d0 = np.exp(softclip(t0, inv1)).astype(f32); d1 = np.exp(softclip(t1, inv1)).astype(f32)
    with np.errstate(all='ignore'):
        isw = f32(1) / sw; ish = f32(1) / sh
        zeta = np.where(inside, (vt - y0u) / ehk, f32(0)).astype(f32)
        s = (ehk * ish) / (ewk * isw)
        e = d0 + d1 - f32(2) * s
        a = (s - d0) + zeta * e; b = d0 - zeta * e; c = -s * zeta
        z = f32(2) * c / (-b - np.sqrt(b * b - f32(4) * a * c))
        x = B * (f32(2) * ((x0u + z * ewk) * isw) - f32(1))
    return np.where(inside, x, y).astype(f32), kf.astype(np.int64), inside

print("inverse")
for scale in (0.3, 1.0, 3.0):
    w = (rng.standard_normal((N, K)) * scale); h = rng.standard_normal((N, K)) * scale; d = rng.standard_normal((N, K - 1)) * scale
    y = rng.uniform(-5.5, 5.5, N)
    hor, ver, der = fo.rqs_knots(w, h, d, bound, slope)
    x64, k64, m64 = fo.rqs_inverse(y, hor, ver, der)
    hor32, ver32, der32 = fo.rqs_knots(w.astype(f32), h.astype(f32), d.astype(f32), bound, slope)
    x32, k32, m32 = fo.rqs_inverse(y.astype(f32), hor32, ver32, der32)
    xn, kn, mn = new_inverse(w.astype(f32), h.astype(f32), d.astype(f32), y.astype(f32), bound, slope)
    print(f"scale {scale}: numpy-fp32 ref x: {stats(x32,x64)} binmis {np.sum((k32!=k64)&m64)} maskmis {np.sum(m32!=m64)}")
    print(f"           new form       x: {stats(xn,x64)} binmis {np.sum((kn!=k64)&m64)} maskmis {np.sum(mn!=m64)}")
