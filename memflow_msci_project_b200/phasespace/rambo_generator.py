"""B200 drop-in for ``memflow/phasespace/rambo_generator.py``.

Same class names, method names, argument meaning and error behaviour as the reference's
``FlatInvertiblePhasespace`` (rambo_generator.py:86-736); every batched method is ONE launch of a
hand-written sm_100a kernel through the C ABI (include/memflow_b200.h) instead of thousands of
ATen element-wise ops.  No CPU fallback: tensors must live on a CUDA device.

Deliberate, documented differences from the reference:
  * outputs take the dtype of the input uniforms/momenta (fp32 in -> fp32 out); the reference
    always returns fp64 momenta/weights.  ``reference_dtypes=True`` restores its dtypes.
  * ``getPSpoint_batch`` does not print to stdout (the reference prints two values per call, :633).
  * a batch of exactly 4 events works (the reference's ``square_t`` mis-dispatches on it,
    phasespace/utils.py:45-51).
"""
import math

import torch

from .. import _cabi
from ._desc import MF_QUIRK_F32_VOLUME, build_rambo_desc

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 0.0  # sic: rambo_generator.py:14 (phasespace.py:31 uses 1e-5)


class PhaseSpaceGeneratorError(Exception):
    pass


class VirtualPhaseSpaceGenerator(object):
    """Bookkeeping base class, mirrors rambo_generator.py:29-83."""

    def __init__(self, initial_masses, final_masses, collider_energy, pdf=None, pdf_active=False,
                 uniform_x1x2=True, lhapdf_dir=None, dev=None):
        if dev is None:
            dev = torch.device("cuda:0")
        self.dev = torch.device(dev)
        self.initial_masses = initial_masses
        self.masses_t = final_masses.clone().to(self.dev)
        self.tot_final_state_masses = torch.sum(self.masses_t)
        self.n_initial = len(initial_masses)
        self.n_final = len(final_masses)
        self.collider_energy = collider_energy
        self.pdf = pdf
        self.pdf_active = pdf_active
        self.uniform_x1x2 = uniform_x1x2

    def generateKinematics(self, E_cm, random_variables):
        raise NotImplementedError

    @property
    def nDimPhaseSpace(self):
        if self.n_final == 1:
            return 0
        return 3 * self.n_final - 4


class FlatInvertiblePhasespace(VirtualPhaseSpaceGenerator):
    """RAMBO on diet (S. Platzer, arXiv:1308.2922) on sm_100a."""

    epsilon_border = 1e-10
    absolute_Ecm_min = 1.0

    def __init__(self, *args, compute="f64", reference_dtypes=False, reference_quirks=True,
                 check_inputs=True, **opts):
        """``compute``: "f64" (reference arithmetic, parity mode) or "f32" (throughput mode).
        ``reference_quirks``: reproduce the fp32 overflow of ``get_flatWeights`` (weight = inf,
        rambo_generator.py:132).  ``check_inputs``: True keeps the reference's NaN / CM checks at the call (one
        device->host sync per call); "deferred" lets the kernels record the same conditions in per-event flags
        (MF_FLAG_*) without any sync and raises the reference's exception when ``raise_if_flagged()`` is called
        (e.g. once per step or per epoch); False drops the checks."""
        super().__init__(*args, **opts)
        if self.n_initial == 1:
            raise PhaseSpaceGeneratorError("This basic generator does not support decay topologies.")
        if self.n_initial > 2:
            raise PhaseSpaceGeneratorError(
                "This basic generator does not support more than 2 incoming particles.")
        if not (self.pdf_active and self.uniform_x1x2):
            raise PhaseSpaceGeneratorError(
                "only the pdf_active=True, uniform_x1x2=True mode used by PhaseSpace "
                "(phasespace.py:53-61) is implemented on the B200 path")
        if any(float(m) != 0.0 for m in self.initial_masses):
            raise PhaseSpaceGeneratorError("massive incoming partons are not implemented")
        self.compute = {"f64": _cabi.MF_COMPUTE_F64, "f32": _cabi.MF_COMPUTE_F32}[compute]
        self.reference_dtypes = reference_dtypes
        self.quirks = MF_QUIRK_F32_VOLUME if reference_quirks else 0
        self.reference_quirks = bool(reference_quirks)  # also selects the autograd graph (phasespace/_autograd.py)
        self.check_inputs = check_inputs
        self._desc_cache = {}
        self._pending_flags = []   # check_inputs="deferred": flag buffers of the calls since the last raise_if_flagged()
        self._folded_flags = None

    # ---- deferred input checks ---------------------------------------------------------------
    def _defer(self, flags):
        self._pending_flags.append(flags)
        if len(self._pending_flags) >= 32:   # bound the memory held: fold to one word on the device, still no sync
            self._fold()

    def _fold(self):
        if self._pending_flags:
            # OR of all flag words, bit by bit: amax over events of (word & bit) is that bit if any event carries it
            dev = self._pending_flags[0].device
            bits = torch.tensor([_cabi.MF_FLAG_NAN_INPUT, _cabi.MF_FLAG_NOT_CM, _cabi.MF_FLAG_BELOW_THRESHOLD],
                                dtype=torch.int32, device=dev)
            word = None
            for f in self._pending_flags:
                if f.numel() == 0:
                    continue
                m = (f.unsqueeze(1) & bits).amax(0).sum()   # the three bits are disjoint: their sum is their OR
                word = m if word is None else (word | m)
            if word is not None:
                self._folded_flags = word if self._folded_flags is None else (self._folded_flags | word)
            self._pending_flags = []

    def raise_if_flagged(self):
        """check_inputs="deferred": one synchronisation, then the exceptions the reference raises at the call."""
        self._fold()
        if self._folded_flags is None:
            return
        word = int(self._folded_flags.item())
        self._folded_flags = None
        if word & _cabi.MF_FLAG_NAN_INPUT:
            raise PhaseSpaceGeneratorError(
                "Some of the random variables passed to the phase-space generator are NaN")
        if word & _cabi.MF_FLAG_NOT_CM:
            raise Exception("Momenta batch not in the CM, failing to convert back to PS point")

    # ---- small host helpers kept for API parity (rambo_generator.py:111-149) ----------------
    @staticmethod
    def get_flatWeights(E_cm, n):
        if n == 1:
            return 1.0
        c = math.pow(2 * math.pi, 4 - 3 * n) * math.pow(math.pi / 2.0, n - 1)
        f = math.factorial(n - 1) * math.factorial(n - 2)
        if not torch.is_tensor(E_cm):
            return c * (math.pow(E_cm ** 2, n - 2) / f)
        E_cm = E_cm.to(torch.float)
        return c * (torch.pow(E_cm ** 2, n - 2) / f)

    def massless_map(self, x, exp):
        return (x ** exp) * ((exp + 1) - exp * x)

    @staticmethod
    def rho(M, N, m):
        Msqr = M ** 2
        return ((Msqr - (N + m) ** 2) * (Msqr - (N - m) ** 2)) ** 0.5 / (8.0 * Msqr)

    # ---- descriptor cache --------------------------------------------------------------------
    def _desc(self, key, masses, cuts=(-1, -1, -1)):
        k = (key, tuple(float(c) for c in cuts))
        d = self._desc_cache.get(k)
        if d is None:
            d = build_rambo_desc(masses, self.collider_energy, quirks=self.quirks,
                                 pT_mincut=cuts[0], delR_mincut=cuts[1], rap_maxcut=cuts[2])
            self._desc_cache[k] = d
        return d

    # ---- forward: generateKinematics_batch, rambo_generator.py:168-398 -----------------------
    def generateKinematics_batch(self, random_variables_full, pT_mincut=-1, delR_mincut=-1,
                                 rap_maxcut=-1, pdgs=[0, 0], requires_grad=False, want_logw=False):
        r = random_variables_full
        _cabi.require_cuda(r, "random_variables_full")
        if r.dim() != 2 or r.shape[1] != self.nDimPhaseSpace + 2:
            # the reference asserts on the column count (:251)
            raise AssertionError(
                f"expected [N, {self.nDimPhaseSpace + 2}] uniforms, got {tuple(r.shape)}")
        if self.check_inputs is True and torch.isnan(r).any():
            raise PhaseSpaceGeneratorError(
                "Some of the random variables passed to the phase-space generator are NaN")
        if requires_grad or r.requires_grad:
            from ._autograd import RamboForwardFn
            out = RamboForwardFn.apply(self, r, pT_mincut, delR_mincut, rap_maxcut)
            if want_logw:  # differentiable log of the weight (the kernel's overflow-safe log w is a no-grad output)
                return (*out, torch.log(out[1]))
            return out
        in_dtype = r.dtype
        if self.reference_dtypes and r.dtype != torch.float64:
            r = r.double()
        if r.dtype not in (torch.float32, torch.float64):
            r = r.float()
        r = r.contiguous()
        N, n = r.shape[0], self.n_final
        desc = self._desc("fwd", self.masses_t, (pT_mincut, delR_mincut, rap_maxcut))
        mom = torch.empty((N, n + 2, 4), dtype=r.dtype, device=r.device)
        weight = torch.empty(N, dtype=r.dtype, device=r.device)
        x1 = torch.empty(N, dtype=r.dtype, device=r.device)
        x2 = torch.empty(N, dtype=r.dtype, device=r.device)
        logw = torch.empty(N, dtype=r.dtype, device=r.device) if want_logw else None
        compute = self.compute if r.dtype == torch.float32 else _cabi.MF_COMPUTE_F64
        flags = torch.empty(N, dtype=torch.int32, device=r.device) if self.check_inputs == "deferred" else None
        _cabi.rambo_forward(desc, r, mom, weight, logw, x1, x2, flags, compute)
        if flags is not None:
            self._defer(flags)
        if self.reference_dtypes and in_dtype != torch.float64:
            x1, x2 = x1.to(in_dtype), x2.to(in_dtype)  # the reference's x1/x2 follow the uniforms' dtype
        if want_logw:
            return mom, weight, x1, x2, logw
        return mom, weight, x1, x2

    # ---- inverse: getPSpoint_batch, rambo_generator.py:615-736 -------------------------------
    def getPSpoint_batch(self, momenta_batch, x1, x2, order=[0, 1, 2, 3],
                         target_mass=torch.Tensor([[M_HIGGS, M_TOP, M_TOP, M_GLUON]]),
                         ensure_CM=True, ensure_onShell=True):
        P = momenta_batch
        _cabi.require_cuda(P, "momenta_batch")
        n = self.n_final
        perm = [int(i) for i in order]
        if len(perm) == P.shape[1] and sorted(perm) == list(range(P.shape[1])):
            use_perm = None if perm == list(range(n)) else perm  # permutation: done in-kernel
        else:
            P = P[:, perm, :]                                    # sub-selection (:624-625)
            use_perm = None
        if P.shape[1] != n or len(perm) != n:
            raise IndexError(f"order {order} does not select n_final={n} particles")
        if self.reference_dtypes and P.dtype != torch.float64:
            P = P.double()
        if P.dtype not in (torch.float32, torch.float64):
            P = P.float()
        P = P.contiguous()
        x1 = x1.to(P.dtype).contiguous()
        x2 = x2.to(P.dtype).contiguous()
        if ensure_onShell:
            key = ("inv", tuple(perm))
            desc = self._desc_cache.get(key)
            if desc is None:
                desc = build_rambo_desc(self.masses_t[perm], self.collider_energy,
                                        sum_mass=self.tot_final_state_masses, quirks=self.quirks)
                self._desc_cache[key] = desc
        else:
            # masses_t = target_mass, un-permuted (:651-652); the x1/x2 map keeps the generator's
            # own total mass (:699)
            desc = build_rambo_desc(target_mass.reshape(-1), self.collider_energy,
                                    sum_mass=self.tot_final_state_masses, quirks=self.quirks)
        N = P.shape[0]
        r = torch.empty((N, 3 * n - 2), dtype=P.dtype, device=P.device)
        prob = torch.empty(N, dtype=P.dtype, device=P.device)
        flags = torch.empty(N, dtype=torch.int32, device=P.device) if ensure_CM else None
        compute = self.compute if P.dtype == torch.float32 else _cabi.MF_COMPUTE_F64
        _cabi.rambo_inverse(desc, P, x1, x2, use_perm, r, prob, flags, compute)
        if ensure_CM and self.check_inputs == "deferred":
            self._defer(flags)
        elif ensure_CM and self.check_inputs and (flags & _cabi.MF_FLAG_NOT_CM).any():
            raise Exception("Momenta batch not in the CM, failing to convert back to PS point")
        return r, prob
