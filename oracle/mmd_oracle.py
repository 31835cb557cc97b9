"""CPU oracle of the MMD loss.  TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product
package (memflow_msci_project_b200/ has no CPU path).

numpy restatement of memflow/unfolding_flow/mmd_loss.py:4-41, step by step in the reference's own order (Gram matrices,
squared distances from their diagonals, one accumulator per sample pair, bandwidth loops, mean).  Pinned: checked in
tests/test_mmd_cpu.py against tests/golden/mmd.npz, which tools/make_golden_mmd.py wrote by running the UNMODIFIED
reference function (value, and autograd gradients) in the build container.
"""
import numpy as np

MULTISCALE = (0.1, 0.2, 0.5, 0.9, 1.3)   # mmd_loss.py:27
RBF = (10, 15, 20, 50)                    # mmd_loss.py:35


def mmd(x, y, kernel):
    """mean(XX + YY - 2 XY); x, y [B, D] arrays of one floating dtype (arithmetic stays in that dtype)."""
    x = np.asarray(x)
    y = np.asarray(y)
    xx, yy, zz = x @ x.T, y @ y.T, x @ y.T                       # mmd_loss.py:13
    rx = np.broadcast_to(np.diag(xx)[None, :], xx.shape)         # :14
    ry = np.broadcast_to(np.diag(yy)[None, :], yy.shape)         # :15
    dxx = rx.T + rx - 2.0 * xx                                   # :17
    dyy = ry.T + ry - 2.0 * yy                                   # :18
    dxy = rx.T + ry - 2.0 * zz                                   # :19
    XX, YY, XY = (np.zeros(xx.shape, x.dtype) for _ in range(3))  # :21-23
    if kernel == "multiscale":                                   # :25-31
        for a in MULTISCALE:
            a2 = x.dtype.type(a ** 2)
            XX += a2 * (a2 + dxx) ** -1
            YY += a2 * (a2 + dyy) ** -1
            XY += a2 * (a2 + dxy) ** -1
    if kernel == "rbf":                                          # :33-39
        for a in RBF:
            XX += np.exp(-0.5 * dxx / a)
            YY += np.exp(-0.5 * dyy / a)
            XY += np.exp(-0.5 * dxy / a)
    return np.mean(XX + YY - 2.0 * XY)                           # :41


def mmd_grad(x, y, kernel):
    """Analytic d mmd / d x, d mmd / d y in fp64 (what autograd derives from the lines above):
    d/dz_i = 4 / B^2 * sum_j s_ij k'(|z_i - z_j|^2) (z_i - z_j), s = +1 inside a sample, -1 across."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    n = x.shape[0]
    z = np.concatenate((x, y), 0)
    diff = z[:, None, :] - z[None, :, :]
    d2 = (diff ** 2).sum(-1)
    if kernel == "multiscale":
        dk = sum(-(a ** 2) / (a ** 2 + d2) ** 2 for a in MULTISCALE)
    elif kernel == "rbf":
        dk = sum(-0.5 / a * np.exp(-0.5 * d2 / a) for a in RBF)
    else:
        dk = np.zeros_like(d2)
    s = np.ones((2 * n, 2 * n))
    s[:n, n:] = -1.0
    s[n:, :n] = -1.0
    g = 4.0 / n ** 2 * ((s * dk)[:, :, None] * diff).sum(1)
    return g[:n], g[n:]
