"""Rows a6 / f2 of SURVEY.md section 8: the four-vector helpers of memflow/phasespace/utils.py and the coordinate glue
of memflow/unfolding_flow/utils.py that the reference's wrappers call directly, against golden values produced by the
UNMODIFIED reference (tools/make_golden_ps_utils.py).  The helpers are device-agnostic torch expressions, so this
parity test runs on the CPU."""
import os

import numpy as np
import pytest
import torch

from memflow_msci_project_b200.phasespace import utils as pu
from memflow_msci_project_b200.unfolding_flow.utils import Compute_ParticlesTensor as CPT


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ps_utils.npz"))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(a, b, tol=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) < tol


def test_four_vector_helpers_match_the_reference(gold):
    p4, q4 = _t(gold["p4"]), _t(gold["q4"])
    _close(pu.rho2_t(q4), gold["rho2_t"])
    _close(pu.rho2_tt(p4), gold["rho2_tt"])
    _close(pu.boostVector_t(q4), gold["boostVector_t"])
    _close(pu.square_t(p4[:, 1]), gold["square_t"])
    _close(pu.mass_t(p4[:, 1]), gold["mass_t"])
    _close(pu.dot_t(p4[:, 0], p4[:, 2]), gold["dot_t"])
    _close(pu.dot_s(p4[:, 0, 1:], p4[:, 1, 1:]), gold["dot_s"])
    _close(pu.set_square_t(p4[:, 0], 91.19 ** 2), gold["set_square_t"])
    beta = pu.boostVector_t(q4)
    _close(pu.boost_t(p4[:, 0], -beta), gold["boost_t"])
    _close(pu.boost_tt(p4, -beta.unsqueeze(1)), gold["boost_tt"])
    # boost_tt(p, -boostVector_t(total)) brings the total to rest: (M, 0, 0, 0)
    tot = pu.boost_tt(p4, -beta.unsqueeze(1)).sum(1)
    assert tot[:, 1:].abs().max().item() < 1e-9


def test_lab_frame_boost_and_x1x2_maps_match_the_reference(gold):
    _close(pu.boost_to_lab_frame(_t(gold["mom6"]), _t(gold["x1"]), _t(gold["x2"])), gold["boost_to_lab_frame"])
    a, b, w = pu.get_x1x2_from_uniform(_t(gold["unif"]), torch.tensor(470.25001), 13000.0)
    _close(torch.stack((a, b, w), 1), gold["x1x2_from_uniform"])             # q, A in fp32 like the reference
    u, jac = pu.get_uniform_from_x1x2(a, b, torch.tensor(470.25001), 13000.0)
    _close(torch.cat((u, jac.unsqueeze(1)), 1), gold["uniform_from_x1x2"])
    _close(u, gold["unif"], tol=1e-5)                                         # and the two maps invert each other


def test_coordinate_glue_matches_the_reference(gold):
    p = _t(gold["p4"])[:, 1]
    pte = CPT.get_ptetaphi_comp(p)
    _close(pte, gold["ptetaphi"])
    _close(CPT.get_ptetaphi_comp_batch(_t(gold["p4"]))[:, 1], gold["ptetaphi"])
    _close(CPT.get_cartesian_comp(pte[:, 1:], 172.5), gold["cartesian"])
    _close(CPT.get_cartesian_comp(pte[:, 1:], 172.5), p.numpy(), tol=1e-9)      # round trip to the input four-vector
