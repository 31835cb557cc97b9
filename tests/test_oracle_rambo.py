"""CPU: the C oracle (oracle/rambo_oracle.c) against the golden vectors produced by the
UNMODIFIED reference (tools/make_golden_rambo.py), plus domain invariants of the map."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_RAMBO, momenta_rel_err, rel_err, wrap_diff
from memflow_msci_project_b200.phasespace._desc import build_rambo_desc
from oracle import rambo as orc


def load(golden_dir, name):
    return dict(np.load(f"{golden_dir}/rambo_{name}.npz"))


@pytest.mark.parametrize("name", GOLDEN_RAMBO)
def test_forward_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    desc = build_rambo_desc(torch.tensor(g["masses"]), 13000)
    mom, w, logw, x1, x2, flags = orc.forward(desc, g["r"].astype(np.float64))
    assert momenta_rel_err(mom, g["momenta"]).max() < 1e-10
    assert rel_err(x1, g["x1"]).max() < 1e-13 and rel_err(x2, g["x2"]).max() < 1e-13
    fin = np.isfinite(g["weight"])
    assert np.array_equal(np.isfinite(w), fin)            # same fp32-overflow pattern (n = 8)
    assert rel_err(w[fin], g["weight"][fin]).max() < 1e-6  # fp32 in-place chain of the reference
    assert rel_err(np.exp(logw[fin]), g["weight"][fin]).max() < 1e-6
    assert np.all(np.isfinite(logw))
    assert not flags.any()


@pytest.mark.parametrize("name", GOLDEN_RAMBO)
def test_inverse_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    desc = build_rambo_desc(torch.tensor(g["masses"]), 13000)
    r, prob, flags = orc.inverse(desc, g["momenta"][:, 2:], g["x1"], g["x2"])
    assert wrap_diff(r, g["r_inv"]).max() < 1e-7
    ok = np.isfinite(g["detjacinv"]) & (g["detjacinv"] != 0)
    assert rel_err(prob[ok], g["detjacinv"][ok]).max() < 1e-6
    assert np.array_equal(prob == 0, g["detjacinv"] == 0)
    assert not flags.any()


@pytest.mark.parametrize("name", ["n3_Htt", "n4_HttG"])
def test_cuts_match_reference(golden_dir, name):
    g = load(golden_dir, name)
    pt, dr, eta = g["cuts"]
    desc = build_rambo_desc(torch.tensor(g["masses"]), 13000, pT_mincut=pt, delR_mincut=dr,
                            rap_maxcut=eta)
    _, w, *_ = orc.forward(desc, g["r"].astype(np.float64))
    assert np.array_equal(w == 0, g["weight_cuts"] == 0)
    keep = g["weight_cuts"] != 0
    assert 0.05 < keep.mean() < 0.95
    assert rel_err(w[keep], g["weight_cuts"][keep]).max() < 1e-6


def test_get_ps_matches_reference(golden_dir):
    g = dict(np.load(f"{golden_dir}/rambo_get_PS.npz"))
    tm = torch.Tensor([[125.25, 172.5, 172.5, 1e-5]])
    desc = build_rambo_desc(tm[0], 13000, sum_mass=tm.sum())
    ps, mask = orc.get_ps(desc, g["momenta"], g["boost"])
    assert np.array_equal(mask, g["mask"]) and mask.sum() >= 2
    good = ~mask
    assert wrap_diff(ps[good], g["ps"][good]).max() < 1e-7
    perm = [int(i) for i in g["perm"]]
    desc_p = build_rambo_desc(tm[0, perm], 13000, sum_mass=tm.sum())
    ps2, mask2 = orc.get_ps(desc_p, g["momenta"], g["boost"], perm=perm)
    assert np.array_equal(mask2, g["mask_perm"])
    good = ~mask2
    assert wrap_diff(ps2[good], g["ps_perm"][good]).max() < 1e-7


@pytest.mark.parametrize("masses", [[125.25, 172.5, 172.5, 1e-5], [4.7, 4.7, 4.7, 1e-3, 1e-3, 4.7, 1e-3, 1e-3],
                                    [80.4, 91.2, 1.0, 2.0, 3.0]])
def test_invariants(masses):
    """Size-independent properties: momentum conservation, on-shell masses,
    weight * detjacinv = 1, round trip r -> p -> r."""
    n = len(masses)
    rng = np.random.default_rng(7)
    r = rng.uniform(1e-6, 1 - 1e-6, size=(20000, 3 * n - 2)).astype(np.float32).astype(np.float64)
    desc = build_rambo_desc(torch.tensor(masses), 13000, quirks=0)
    mom, w, logw, x1, x2, _ = orc.forward(desc, r)
    tot_in = mom[:, :2].sum(1)
    tot_out = mom[:, 2:].sum(1)
    assert np.abs(tot_in - tot_out).max() < 2e-3  # float(E_cm) incoming partons, reference quirk
    m2 = mom[:, 2:, 0] ** 2 - (mom[:, 2:, 1:] ** 2).sum(-1)
    heavy = np.asarray(masses) > 1.0
    assert np.abs(np.sqrt(np.abs(m2[:, heavy])) - np.asarray(masses)[heavy]).max() < 1e-5
    r2, prob, _ = orc.inverse(desc, mom[:, 2:], x1, x2)
    # the inverse recovers light-particle masses from E^2 - p^2 (fp64 cancellation), so the round
    # trip of the n = 8, m = 1e-3 case is only good to ~1e-3 in the worst event (reference likewise)
    d = wrap_diff(r2, r)
    assert np.median(d) < 1e-7 and np.quantile(d, 0.999) < 1e-4 and d.max() < (5e-6 if n == 4 else 5e-3)
    e = np.abs(w * prob - 1)
    assert np.median(e) < 1e-6 and np.quantile(e, 0.99) < 1e-4 and e.max() < (1e-5 if n == 4 else 2e-2)
