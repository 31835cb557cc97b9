"""B200 drop-in for ``memflow/phasespace/phasespace.py`` (``PhaseSpace``, phasespace.py:35-109)."""
import torch

from . import rambo_generator
from . import utils  # noqa: F401  (the reference module exposes it as well)

M_HIGGS = 125.25
M_TOP = 172.5
M_GLUON = 1e-5


class PhaseSpace:
    """Same constructor and methods as the reference class; every call is one fused kernel.

    Extra keyword arguments (all optional): ``compute`` ("f64" parity mode / "f32" throughput
    mode), ``reference_dtypes`` (return fp64 momenta/weights like the reference even for fp32
    uniforms), ``reference_quirks``, ``check_inputs`` -- see FlatInvertiblePhasespace.
    """

    def __init__(self, collider_energy, initial_pdgs, final_pdgs,
                 final_masses=torch.tensor([M_HIGGS, M_TOP, M_TOP, M_GLUON]), pdf=None, dev=None,
                 **generator_opts):
        if dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError("memflow_msci_project_b200 needs a CUDA device (no CPU fallback)")
            dev = torch.device("cuda:0")
        self.dev = torch.device(dev)
        self.collider_energy = collider_energy
        self.initial_pdgs = initial_pdgs
        self.final_pdgs = final_pdgs
        self.final_masses = final_masses.to(self.dev)
        self.final_state_mass = torch.sum(self.final_masses)
        self.pdf = pdf
        self.generator = rambo_generator.FlatInvertiblePhasespace(
            [0.0, 0.0], self.final_masses, self.collider_energy, pdf=pdf, pdf_active=True,
            uniform_x1x2=True, dev=self.dev, **generator_opts)

    def generate_random_phase_space_points(self, N, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1):
        """phasespace.py:64-80: N uniform points and their image."""
        ps_rand = torch.rand(N, self.generator.nDimPhaseSpace + 2, device=self.dev)
        momenta, weight, x1, x2 = self.generator.generateKinematics_batch(
            ps_rand, pdgs=self.initial_pdgs, pT_mincut=pT_mincut, delR_mincut=delR_mincut,
            rap_maxcut=rap_maxcut)
        return ps_rand, momenta, weight, x1, x2

    def get_momenta_from_ps(self, points, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1,
                            requires_grad=False):
        """phasespace.py:82-103: uniforms [N,3n-2] -> (momenta [N,n+2,4] in the partonic CM frame,
        weight [N], x1 [N], x2 [N])."""
        return self.generator.generateKinematics_batch(
            points, pdgs=self.initial_pdgs, pT_mincut=pT_mincut, delR_mincut=delR_mincut,
            rap_maxcut=rap_maxcut, requires_grad=requires_grad)

    def get_ps_from_momenta(self, momenta, x1, x2, ensure_CM=True):
        """phasespace.py:105-109: final-state momenta [N,n,4] -> (ps [N,3n-2], detjacinv [N])."""
        n = self.generator.n_final
        order = [0, 1, 2, 3] if n == 4 else list(range(n))  # the reference hard-codes n = 4 here
        return self.generator.getPSpoint_batch(momenta, x1, x2, order=order)
