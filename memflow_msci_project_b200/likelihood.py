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
    lp = lp.contiguous()
    out = torch.empty(B, dtype=lp.dtype, device=lp.device)
    m = None
    if exist_mask is not None:
        m = exist_mask.to(torch.uint8).contiguous()
    _cabi.event_reduce(lp, m, logw.to(lp.dtype), out)
    return out, mom, weight
