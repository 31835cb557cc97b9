"""Autoregressive masks of the hyper-network, built the way zuko builds them.

``zuko.flows.autoregressive.MaskedAutoregressiveTransform.__init__`` derives an adjacency matrix
from the feature order and hands it to ``zuko.nn.MaskedMLP``, which turns it into one fixed boolean
mask per ``MaskedLinear``.  A checkpoint carries those masks as buffers, so loading one overrides
whatever is built here; for freshly initialised flows the construction below is what decides which
weights exist, and it is checked against the oracle's independent numpy restatement.
"""
import math

import torch


def autoregressive_adjacency(features, context, passes, order, total):
    """-> (adjacency bool [features*total, features+context], effective order, effective passes)."""
    if passes is None:
        passes = features
    if order is None:
        order = torch.arange(features)
    else:
        order = torch.as_tensor(order, dtype=torch.long)
    passes = min(max(int(passes), 1), features)
    order = torch.div(order, math.ceil(features / passes), rounding_mode="floor")
    in_order = torch.cat((order, torch.full((context,), -1, dtype=torch.long)))
    out_order = torch.repeat_interleave(order, total)
    return out_order[:, None] > in_order, order, passes


def mlp_masks(adjacency, hidden_features):
    """One boolean mask [out_l, in_l] per linear layer of MaskedMLP(adjacency, hidden_features)."""
    adjacency = adjacency.to(torch.bool)
    uniq, inverse = torch.unique(adjacency, dim=0, return_inverse=True)
    a = uniq.to(torch.long)
    precedence = (a @ a.t()) == a.sum(dim=-1)  # P_ij: dependencies of row j are a subset of row i's
    masks, indices = [], None
    sizes = list(hidden_features) + [adjacency.shape[0]]
    for i, width in enumerate(sizes):
        mask = precedence[:, indices] if i > 0 else uniq
        if (~mask).all():
            raise ValueError("The adjacency matrix leads to a null Jacobian.")
        if i < len(sizes) - 1:
            reachable = mask.sum(dim=-1).nonzero().squeeze(dim=-1)
            indices = reachable[torch.arange(width) % len(reachable)]
            mask = mask[indices]
        else:
            mask = mask[inverse]
        masks.append(mask.clone())
    return masks
