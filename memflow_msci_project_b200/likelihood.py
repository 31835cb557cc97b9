"""Event-level likelihood: the composition the headline metric times.

``event_log_prob`` evaluates, for a batch of events, the transfer-flow density of every reconstructed
object and the phase-space Jacobian of the parton configuration:

    log p(event) = sum_objects mask * flow(c_obj).log_prob(x_obj)  -  log(w_RAMBO(ps))

i.e. the per-event sum of memflow/transfer_flow/transfer_flow_paper_AllPartons_btag.py:210-216 and the
``flow().log_prob(ps) - log(rambo_jac)`` bookkeeping of notebooks/TestFlows_ImportanceSampling.ipynb
cell 52.  Three kernel launches: RAMBO forward, fused flow log_prob, per-event reduction.
"""
import torch

from . import _cabi


def event_log_prob(phasespace, flow, ps_points, x, context, exist_mask=None):
    """ps_points [B, 3n-2]; x [B, P, D]; context [B, P, C]; exist_mask [B, P] bool/uint8 or None.
    Returns (log_prob_event [B], momenta [B, n+2, 4], weight [B])."""
    B, P, D = x.shape
    mom, weight, x1, x2, logw = phasespace.generator.generateKinematics_batch(ps_points, want_logw=True)
    lp = flow(context).log_prob(x)  # [B, P]
    if torch.is_grad_enabled() and (lp.requires_grad or logw.requires_grad):
        # training: the per-event sum stays a differentiable torch expression (three element-wise ops on [B, P])
        lpm = lp if exist_mask is None else lp * exist_mask.to(lp.dtype)
        return lpm.sum(1) - logw.to(lp.dtype), mom, weight
    lp = lp.contiguous()
    out = torch.empty(B, dtype=lp.dtype, device=lp.device)
    m = None
    if exist_mask is not None:
        m = exist_mask.to(torch.uint8).contiguous()
    _cabi.event_reduce(lp, m, logw.to(lp.dtype), out)
    return out, mom, weight


class HostPipeline:
    """End-to-end evaluation from pinned HOST buffers: the batch is cut into chunks and the host->device copy of
    chunk i+1 overlaps the three kernels of chunk i (two CUDA streams, double-buffered device staging); the
    per-event result goes back to a pinned host buffer.  This is the `e2e` path of bench.py."""

    def __init__(self, phasespace, flow, n_events, n_objects, features, context, n_ps, device, chunks=8,
                 dtype=torch.float32):
        self.ps, self.flow, self.dev = phasespace, flow, device
        self.n, self.P = n_events, n_objects
        self.chunk = (n_events + chunks - 1) // chunks
        self.copy_stream = torch.cuda.Stream(device)
        self.compute_stream = torch.cuda.Stream(device)

        def buf(*shape, dt=dtype):
            return [torch.empty(shape, dtype=dt, device=device) for _ in range(2)]

        c = self.chunk
        self.d_r, self.d_x = buf(c, n_ps), buf(c, n_objects, features)
        self.d_c, self.d_m = buf(c, n_objects, context), buf(c, n_objects, dt=torch.uint8)
        self.d_out = torch.empty(n_events, dtype=dtype, device=device)
        self.h_out = torch.empty(n_events, dtype=dtype).pin_memory()
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.done = torch.cuda.Event()

    def run(self, h_r, h_x, h_c, h_m):
        """All arguments are pinned host tensors of the full batch.  Returns the pinned host result; the method blocks
        the calling host thread until the device-to-host copy of the result has landed (so the buffer can be read, and
        a following run() cannot overwrite it while it is being read)."""
        cur = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(cur)
        self.compute_stream.wait_stream(cur)
        n, c = self.n, self.chunk
        with torch.no_grad():
            for i, a in enumerate(range(0, n, c)):
                b = min(n, a + c)
                k, s = b - a, i & 1
                with torch.cuda.stream(self.copy_stream):
                    if i >= 2:
                        self.copy_stream.wait_event(self.consumed[s])  # staging buffer s is free again
                    self.d_r[s][:k].copy_(h_r[a:b], non_blocking=True)
                    self.d_x[s][:k].copy_(h_x[a:b], non_blocking=True)
                    self.d_c[s][:k].copy_(h_c[a:b], non_blocking=True)
                    self.d_m[s][:k].copy_(h_m[a:b], non_blocking=True)
                    self.copied[s].record(self.copy_stream)
                with torch.cuda.stream(self.compute_stream):
                    self.compute_stream.wait_event(self.copied[s])
                    out, _, _ = event_log_prob(self.ps, self.flow, self.d_r[s][:k], self.d_x[s][:k], self.d_c[s][:k],
                                               self.d_m[s][:k])
                    self.d_out[a:b].copy_(out)
                    self.consumed[s].record(self.compute_stream)
            with torch.cuda.stream(self.compute_stream):
                self.h_out.copy_(self.d_out, non_blocking=True)
                self.done.record(self.compute_stream)
        cur.wait_stream(self.compute_stream)
        cur.wait_stream(self.copy_stream)
        self.done.synchronize()
        return self.h_out
