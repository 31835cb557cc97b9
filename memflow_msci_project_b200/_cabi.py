"""ctypes binding of libmemflow_b200.so (the C ABI declared in include/memflow_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the
product path raises.  PyTorch only provides device memory and streams here.
"""
import ctypes
import os

import torch

from .phasespace._desc import RamboDesc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libmemflow_b200.so")

MF_F32, MF_F64 = 0, 1
MF_COMPUTE_F64, MF_COMPUTE_F32 = 0, 1
MF_FLAG_NAN_INPUT, MF_FLAG_NOT_CM, MF_FLAG_BELOW_THRESHOLD = 1, 2, 4

_lib = None


class MemflowB200Error(RuntimeError):
    pass


def lib():
    """Load the CUDA extension; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MemflowB200Error(
                f"{LIB_PATH} not found: build it with `python -m memflow_msci_project_b200.build` "
                "(there is no CPU or eager fallback on purpose)")
        L = ctypes.CDLL(LIB_PATH)
        L.mf_last_error.restype = ctypes.c_char_p
        L.mf_launch_count.restype = ctypes.c_ulonglong
        L.mf_abi_version.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mf_last_error().decode(errors="replace")
        raise MemflowB200Error(f"{what} failed (rc={rc}): {msg}")


def launch_count():
    return int(lib().mf_launch_count())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def io_dtype_of(t):
    if t.dtype == torch.float32:
        return MF_F32
    if t.dtype == torch.float64:
        return MF_F64
    raise TypeError(f"unsupported dtype {t.dtype} (float32/float64 only)")


def require_cuda(t, name):
    if not t.is_cuda:
        raise MemflowB200Error(
            f"{name} lives on {t.device}: the B200 path has no CPU fallback, move it to a CUDA device")


def rambo_forward(desc: RamboDesc, r, momenta, weight, logw, x1, x2, flags, compute):
    with torch.cuda.device(r.device):
        rc = lib().mf_rambo_forward(ctypes.byref(desc), _ptr(r), ctypes.c_int64(r.shape[0]),
                                    _ptr(momenta), _ptr(weight), _ptr(logw), _ptr(x1), _ptr(x2),
                                    _ptr(flags), ctypes.c_int(io_dtype_of(r)), ctypes.c_int(compute),
                                    _stream(r.device))
    check(rc, "mf_rambo_forward")


def rambo_inverse(desc: RamboDesc, momenta, x1, x2, perm, r, detjacinv, flags, compute):
    n = desc.n_final
    perm_arr = (ctypes.c_int32 * n)(*perm) if perm is not None else None
    with torch.cuda.device(momenta.device):
        rc = lib().mf_rambo_inverse(ctypes.byref(desc), _ptr(momenta), _ptr(x1), _ptr(x2), perm_arr,
                                    ctypes.c_int64(momenta.shape[0]), _ptr(r), _ptr(detjacinv),
                                    _ptr(flags), ctypes.c_int(io_dtype_of(momenta)),
                                    ctypes.c_int(compute), _stream(momenta.device))
    check(rc, "mf_rambo_inverse")


def rambo_get_ps(desc: RamboDesc, momenta, boost, perm, ps, mask, compute):
    perm_arr = (ctypes.c_int32 * 4)(*perm) if perm is not None else None
    with torch.cuda.device(momenta.device):
        rc = lib().mf_rambo_get_ps(ctypes.byref(desc), _ptr(momenta), _ptr(boost), perm_arr,
                                   ctypes.c_int64(momenta.shape[0]), _ptr(ps), _ptr(mask),
                                   ctypes.c_int(io_dtype_of(momenta)), ctypes.c_int(compute),
                                   _stream(momenta.device))
    check(rc, "mf_rambo_get_ps")


def rambo_get_ps_lab(desc: RamboDesc, momenta_lab, boost, perm, m_min_tot, momenta_cm, boost_fixed, ps, mask, compute):
    """Fused boost fix-up + lab -> CM boost + get_PS (mf_rambo_get_ps_lab)."""
    perm_arr = (ctypes.c_int32 * 4)(*perm) if perm is not None else None
    with torch.cuda.device(momenta_lab.device):
        rc = lib().mf_rambo_get_ps_lab(ctypes.byref(desc), _ptr(momenta_lab), _ptr(boost), perm_arr,
                                       ctypes.c_int64(momenta_lab.shape[0]), ctypes.c_double(m_min_tot), _ptr(momenta_cm),
                                       _ptr(boost_fixed), _ptr(ps), _ptr(mask), ctypes.c_int(io_dtype_of(momenta_lab)),
                                       ctypes.c_int(compute), _stream(momenta_lab.device))
    check(rc, "mf_rambo_get_ps_lab")


def rambo_forward_backward(desc: RamboDesc, r, g_mom, g_w, g_x1, g_x2, w_fwd, full_derivative, grad_r):
    """Cotangent of the uniforms of the RAMBO forward map (mf_rambo_forward_backward); fp64 tensors, None = absent."""
    with torch.cuda.device(r.device):
        rc = lib().mf_rambo_forward_backward(ctypes.byref(desc), _ptr(r), ctypes.c_int64(r.shape[0]), _ptr(g_mom), _ptr(g_w),
                                             _ptr(g_x1), _ptr(g_x2), _ptr(w_fwd), ctypes.c_int32(1 if full_derivative else 0),
                                             _ptr(grad_r), _stream(r.device))
    check(rc, "mf_rambo_forward_backward")


def gemm_nt(a, b, c, bias=None, relu=False, k_split=1, mask=None, out_mask=None, scale_a=None, scale_b=None):
    """c[m, n] (+)= (a * (mask > 0))[m, k] @ b[n, k].T (+ bias) (ReLU) on tcgen05, fp32-class (mf_gemm_nt_3xf16).  2-D fp32
    CUDA tensors with unit inner stride; any n (columns are processed 256 at a time)."""
    m, k = a.shape
    n = b.shape[0]
    assert b.shape[1] == k and c.shape == (m, n) and a.stride(1) == 1 and b.stride(1) == 1 and c.stride(1) == 1
    assert mask is None or (mask.shape == a.shape and mask.stride(1) == 1)
    assert out_mask is None or (out_mask.shape == c.shape and out_mask.stride(1) == 1)
    with torch.cuda.device(a.device):
        for n0 in range(0, n, 256):
            n1 = min(n, n0 + 256)
            bb, cc = b[n0:n1], c[:, n0:n1]
            om = out_mask[:, n0:n1] if out_mask is not None else None
            rc = lib().mf_gemm_nt_3xf16(_ptr(a), ctypes.c_int64(a.stride(0)), _ptr(bb), ctypes.c_int64(b.stride(0)), _ptr(cc),
                                        ctypes.c_int64(c.stride(0)), ctypes.c_int64(m), ctypes.c_int32(n1 - n0),
                                        ctypes.c_int64(k), _ptr(bias[n0:n1]) if bias is not None else ctypes.c_void_p(0),
                                        ctypes.c_int32(1 if relu else 0), ctypes.c_int32(k_split), _ptr(mask),
                                        ctypes.c_int64(mask.stride(0) if mask is not None else 0), _ptr(om),
                                        ctypes.c_int64(out_mask.stride(0) if out_mask is not None else 0), _ptr(scale_a),
                                        _ptr(scale_b), _stream(a.device))
            check(rc, "mf_gemm_nt_3xf16")


def gemm_tn(a, b, c, k_split=1, mask=None, colsum=None, scale_a=None, scale_b=None):
    """c[m, n] (+)= (a * (mask > 0))[k, m].T @ b[k, n]; colsum[m] += column sums of the masked a (mf_gemm_tn_3xf16).
    n (+1 with colsum) <= 256."""
    k, m = a.shape
    n = b.shape[1]
    assert b.shape[0] == k and c.shape == (m, n) and a.stride(1) == 1 and b.stride(1) == 1 and c.stride(1) == 1
    assert mask is None or (mask.shape == a.shape and mask.stride(1) == 1)
    with torch.cuda.device(a.device):
        rc = lib().mf_gemm_tn_3xf16(_ptr(a), ctypes.c_int64(a.stride(0)), _ptr(b), ctypes.c_int64(b.stride(0)), _ptr(c),
                                    ctypes.c_int64(c.stride(0)), ctypes.c_int64(m), ctypes.c_int32(n), ctypes.c_int64(k),
                                    ctypes.c_int32(k_split), _ptr(mask), ctypes.c_int64(mask.stride(0) if mask is not None else 0),
                                    _ptr(colsum), _ptr(scale_a), _ptr(scale_b), _stream(a.device))
    check(rc, "mf_gemm_tn_3xf16")


# ---- conditional spline flows ---------------------------------------------------------------------
def _dtype_code(dtype):
    return {torch.float32: MF_F32, torch.float64: MF_F64}[dtype]


def flow_blob_bytes(desc, dtype, gemm):
    L = lib()
    L.mf_flow_blob_bytes.restype = ctypes.c_int64
    n = L.mf_flow_blob_bytes(ctypes.byref(desc), ctypes.c_int(_dtype_code(dtype)), ctypes.c_int(gemm))
    if n < 0:
        check(-1, "mf_flow_blob_bytes")
    return int(n)


def flow_pack_weights(desc, weights, biases, masks, dtype, gemm, blob):
    n = len(weights)
    arr_t = ctypes.c_void_p * n
    w = arr_t(*[t.data_ptr() for t in weights])
    b = arr_t(*[t.data_ptr() for t in biases])
    m = arr_t(*[t.data_ptr() for t in masks])
    with torch.cuda.device(blob.device):
        rc = lib().mf_flow_pack_weights(ctypes.byref(desc), w, b, m, ctypes.c_int(_dtype_code(dtype)),
                                        ctypes.c_int(gemm), _ptr(blob), ctypes.c_int64(blob.numel()),
                                        _stream(blob.device))
    check(rc, "mf_flow_pack_weights")


def flow_sched_offset(desc, dtype, gemm):
    L = lib()
    L.mf_flow_sched_offset.restype = ctypes.c_int64
    off = L.mf_flow_sched_offset(ctypes.byref(desc), ctypes.c_int(_dtype_code(dtype)), ctypes.c_int(gemm))
    if off < 0:
        check(-1, "mf_flow_sched_offset")
    return int(off)


def flow_set_sample_schedule(desc, orders, dtype, gemm, blob):
    """Mark, per sweep, the last-layer chunks whose features become final there (mf_flow_set_sample_schedule).
    orders: host int32 array [transforms][features] of zuko's effective orders."""
    arr = (ctypes.c_int32 * len(orders))(*[int(v) for v in orders])
    with torch.cuda.device(blob.device):
        rc = lib().mf_flow_set_sample_schedule(ctypes.byref(desc), arr, ctypes.c_int(_dtype_code(dtype)), ctypes.c_int(gemm),
                                               _ptr(blob), ctypes.c_int64(blob.numel()), _stream(blob.device))
    check(rc, "mf_flow_set_sample_schedule")


def _flow_workspace(desc, n, dtype, gemm, op, device):
    L = lib()
    L.mf_flow_workspace_bytes.restype = ctypes.c_int64
    nb = L.mf_flow_workspace_bytes(ctypes.byref(desc), ctypes.c_int64(n), ctypes.c_int(_dtype_code(dtype)),
                                   ctypes.c_int(gemm), ctypes.c_int(op))
    if nb < 0:
        check(-1, "mf_flow_workspace_bytes")
    return torch.empty(int(nb), dtype=torch.uint8, device=device) if nb > 0 else None


MF_MAX_FEATURES = 64
MF_XMAP_NONE, MF_XMAP_SIGMOID_AFFINE, MF_XMAP_LOGIT_AFFINE = 0, 1, 2


class FlowXmap(ctypes.Structure):
    """ctypes mirror of ``struct mf_flow_xmap`` (element-wise map fused into the x load / store of the flow kernels)."""

    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("eps", ctypes.c_double),
                ("scale", ctypes.c_double * MF_MAX_FEATURES), ("shift", ctypes.c_double * MF_MAX_FEATURES)]


def make_xmap(kind, scale, shift, features, eps=0.0):
    m = FlowXmap()
    m.kind, m.eps = kind, float(eps)
    sc = [float(v) for v in torch.as_tensor(scale, dtype=torch.float64).reshape(-1).expand(features).tolist()] \
        if torch.as_tensor(scale).numel() in (1, features) else None
    sh = [float(v) for v in torch.as_tensor(shift, dtype=torch.float64).reshape(-1).expand(features).tolist()] \
        if torch.as_tensor(shift).numel() in (1, features) else None
    if sc is None or sh is None:
        raise ValueError(f"x-map scale / shift must be scalars or have {features} entries")
    for f in range(features):
        m.scale[f], m.shift[f] = sc[f], sh[f]
    return m


def flow_log_prob(desc, blob, x, ctx, logprob, z, ladj, bins, inside, gemm, xmap=None):
    ws = _flow_workspace(desc, x.shape[0], x.dtype, gemm, 0, x.device)
    with torch.cuda.device(x.device):
        rc = lib().mf_flow_log_prob_mapped(ctypes.byref(desc), _ptr(blob), _ptr(x), _ptr(ctx),
                                           ctypes.c_int64(x.shape[0]), _ptr(logprob), _ptr(z), _ptr(ladj),
                                           _ptr(bins), _ptr(inside), ctypes.c_int(_dtype_code(x.dtype)),
                                           ctypes.c_int(gemm), _ptr(ws), ctypes.c_int64(ws.numel() if ws is not None else 0),
                                           ctypes.byref(xmap) if xmap is not None else None, _stream(x.device))
    check(rc, "mf_flow_log_prob")


def flow_sample(desc, blob, noise, ctx, x, gemm, xmap=None):
    ws = _flow_workspace(desc, noise.shape[0], noise.dtype, gemm, 1, noise.device)
    with torch.cuda.device(noise.device):
        rc = lib().mf_flow_sample_mapped(ctypes.byref(desc), _ptr(blob), _ptr(noise), _ptr(ctx),
                                         ctypes.c_int64(noise.shape[0]), _ptr(x),
                                         ctypes.c_int(_dtype_code(noise.dtype)), ctypes.c_int(gemm), _ptr(ws),
                                         ctypes.c_int64(ws.numel() if ws is not None else 0),
                                         ctypes.byref(xmap) if xmap is not None else None, _stream(noise.device))
    check(rc, "mf_flow_sample")


def event_reduce(logprob, mask, logw, out):
    """out[e] = sum_p logprob[e,p]*mask[e,p] - logw[e]  (mf_event_reduce)."""
    B, P = logprob.shape
    with torch.cuda.device(logprob.device):
        rc = lib().mf_event_reduce(_ptr(logprob), _ptr(mask), _ptr(logw), ctypes.c_int64(B), ctypes.c_int32(P),
                                   _ptr(out), ctypes.c_int(_dtype_code(logprob.dtype)), _stream(logprob.device))
    check(rc, "mf_event_reduce")


MF_MMD_MULTISCALE, MF_MMD_RBF, MF_MMD_MAX_DIM = 0, 1, 32


def mmd(x, y, kernel, out, grad_x=None, grad_y=None):
    """out[0] = MMD(x, y) and, when given, its gradients w.r.t. x and y (mf_mmd); x, y [n, d] contiguous."""
    n, d = x.shape
    L = lib()
    L.mf_mmd_workspace_bytes.restype = ctypes.c_int64
    want_grad = grad_x is not None or grad_y is not None
    nbytes = int(L.mf_mmd_workspace_bytes(ctypes.c_int64(n), ctypes.c_int32(d), ctypes.c_int32(1 if want_grad else 0)))
    with torch.cuda.device(x.device):
        ws = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=x.device)
        rc = L.mf_mmd(_ptr(x), _ptr(y), ctypes.c_int64(n), ctypes.c_int32(d), ctypes.c_int32(kernel), _ptr(ws), _ptr(out),
                      _ptr(grad_x), _ptr(grad_y), ctypes.c_int(_dtype_code(x.dtype)), _stream(x.device))
    check(rc, "mf_mmd")


MF_LOSS_HUBER, MF_LOSS_MSE = 0, 1


def periodic_loss(inp, target, kind, huber_delta, mean, elem=None, deriv=None, total=None):
    """Wrapped-phi regression loss (mf_periodic_loss); inp, target contiguous with n elements."""
    n = inp.numel()
    L = lib()
    L.mf_periodic_loss_workspace_bytes.restype = ctypes.c_int64
    nbytes = int(L.mf_periodic_loss_workspace_bytes(ctypes.c_int64(n)))
    with torch.cuda.device(inp.device):
        ws = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=inp.device)
        rc = L.mf_periodic_loss(_ptr(inp), _ptr(target), ctypes.c_int64(n), ctypes.c_int32(kind), ctypes.c_double(huber_delta),
                                ctypes.c_int32(1 if mean else 0), _ptr(ws), _ptr(elem), _ptr(deriv), _ptr(total),
                                ctypes.c_int(_dtype_code(inp.dtype)), _stream(inp.device))
    check(rc, "mf_periodic_loss")


PROBE_LIB_PATH = os.path.join(_HERE, "_lib", "libmemflow_b200_probe.so")
_probe = None


def probe_lib():
    """Test-only library with the tcgen05 building-block checks and micro-benchmarks (csrc/tc_probe.cu); not part of
    the product ABI."""
    global _probe
    if _probe is None:
        if not os.path.exists(PROBE_LIB_PATH):
            raise MemflowB200Error(f"{PROBE_LIB_PATH} not found: build it with `python -m memflow_msci_project_b200.build`")
        _probe = ctypes.CDLL(PROBE_LIB_PATH)
        _probe.mf_last_error.restype = ctypes.c_char_p
    return _probe


def tc_probe_gemm(a, w, mode):
    """out[128,N] = a[128,K] @ w[N,K]^T on the tensor cores (test hook of the probe library)."""
    out = torch.empty((128, w.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = probe_lib().mf_tc_probe_gemm(_ptr(a), _ptr(w), _ptr(out), ctypes.c_int32(a.shape[1]), ctypes.c_int32(w.shape[0]),
                                          ctypes.c_int32(mode), _stream(a.device))
    if rc != 0:
        raise MemflowB200Error(f"mf_tc_probe_gemm failed (rc={rc}): {probe_lib().mf_last_error().decode(errors='replace')}")
    return out


def rqs_backward(params, x, bins, bound, slope, grad_y, grad_ladj, grad_params, grad_x):
    """Gradients of the forward spline w.r.t. packed parameters [n, 3K-1] and x [n] (mf_rqs_backward)."""
    with torch.cuda.device(x.device):
        rc = lib().mf_rqs_backward(_ptr(params), _ptr(x), ctypes.c_int64(x.numel()), ctypes.c_int32(bins),
                                   ctypes.c_double(bound), ctypes.c_double(slope or 0.0), _ptr(grad_y), _ptr(grad_ladj),
                                   _ptr(grad_params), _ptr(grad_x), ctypes.c_int(_dtype_code(x.dtype)), _stream(x.device))
    check(rc, "mf_rqs_backward")


def rqs_transform(params, x, bins, univariate, bound, slope, inverse, y, ladj, bins_out=None, inside=None):
    """zuko MonotonicRQSTransform on packed parameters [n, 3K-1] (mf_rqs_transform)."""
    with torch.cuda.device(x.device):
        rc = lib().mf_rqs_transform(_ptr(params), _ptr(x), ctypes.c_int64(x.numel()), ctypes.c_int32(bins),
                                    ctypes.c_int32(univariate), ctypes.c_double(bound), ctypes.c_double(slope or 0.0),
                                    ctypes.c_int32(1 if inverse else 0), _ptr(y), _ptr(ladj), _ptr(bins_out), _ptr(inside),
                                    ctypes.c_int(_dtype_code(x.dtype)), _stream(x.device))
    check(rc, "mf_rqs_transform")
