"""Configuration objects of the sweep path -- mirror of renormalizer/utils/configs.py
(CompressConfig :128-260, OptimizeConfig :263-304, EvolveConfig :342-416), restricted to what
the accelerated path consumes."""
from enum import Enum

import numpy as np
import scipy.linalg


class CompressCriteria(Enum):
    threshold = "threshold"
    fixed = "fixed"
    both = "both"


class EvolveMethod(Enum):
    tdvp_ps = "TDVP_PS"
    tdvp_ps2 = "TDVP_PS2"
    prop_and_compress = "P&C"
    prop_and_compress_tdrk4 = "P&C TD RK4"        # configs.py:309-311: Runge-Kutta propagators, the
    prop_and_compress_tdrk = "P&C TD RK"          # Hamiltonian may depend on time
    tdvp_mu_vmf = "TDVP_MU_VMF"
    tdvp_vmf = "TDVP_VMF"
    tdvp_mu_cmf = "TDVP_MU_CMF"


class CompressConfig:
    def __init__(self, criteria=CompressCriteria.threshold, threshold=1e-3, max_bonddim=32,
                 vmethod="2site", vprocedure=None, vrtol=1e-5, vguess_m=(5, 5)):
        if isinstance(criteria, str):
            criteria = CompressCriteria[criteria]
        self.criteria = criteria
        self.threshold = threshold
        self.bond_dim_max_value = max_bonddim
        self.max_dims = None
        self.ofs = None
        # variational compression (configs.py:156-170)
        self.vmethod = vmethod
        if vprocedure is None:
            head = [[max_bonddim, 1.0], [max_bonddim, 0.7]] if vmethod == "1site" else []
            vprocedure = head + [[max_bonddim, 0.5], [max_bonddim, 0.3], [max_bonddim, 0.1]] \
                + [[max_bonddim, 0]] * 10
        self.vprocedure = vprocedure
        self.vrtol = vrtol
        self.vguess_m = vguess_m

    @property
    def threshold(self):
        return self._threshold

    @threshold.setter
    def threshold(self, v):
        if v <= 0:
            raise ValueError("non-positive threshold")
        if v == 1:
            raise ValueError("1 is an ambiguous threshold")
        if 1 < v:
            raise ValueError("Can't set threshold to be larger than 1")
        self._threshold = v

    @property
    def bonddim_should_set(self):
        return self.criteria is not CompressCriteria.threshold and self.max_dims is None

    def set_bonddim(self, length):
        if self.max_dims is None:
            self.max_dims = np.full(length, self.bond_dim_max_value, dtype=int)

    def _threshold_m_trunc(self, sigma):
        normed = sigma / scipy.linalg.norm(sigma)
        return int(np.sum(normed > self.threshold))

    def _fixed_m_trunc(self, sigma, idx, left):
        bond_idx = idx + 1 if left else idx
        return min(int(self.max_dims[bond_idx]), len(sigma))

    def compute_m_trunc(self, sigma, idx, left):
        if self.criteria is CompressCriteria.threshold:
            return self._threshold_m_trunc(sigma)
        if self.criteria is CompressCriteria.fixed:
            return self._fixed_m_trunc(sigma, idx, left)
        return min(self._threshold_m_trunc(sigma), self._fixed_m_trunc(sigma, idx, left))

    def copy(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__ = self.__dict__.copy()
        if self.max_dims is not None:
            new.max_dims = self.max_dims.copy()
        return new

    def update(self, other):
        """configs.py:221-235: the stricter of the two."""
        if self.criteria != other.criteria:
            raise ValueError("Can't update configs with different standard")
        self.threshold = min(self.threshold, other.threshold)
        if self.max_dims is None:
            self.max_dims = other.max_dims
        elif other.max_dims is not None:
            self.max_dims = np.maximum(self.max_dims, other.max_dims)


class OptimizeConfig:
    def __init__(self, procedure=None):
        self.procedure = procedure if procedure is not None else \
            [[10, 0.4], [20, 0.2], [30, 0.1], [40, 0], [40, 0]]
        self.method = "2site"
        self.algo = "davidson"
        self.nroots = 1
        self.e_rtol = 1e-6
        self.e_atol = 1e-8
        self.inverse = 1.0

    def copy(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__ = self.__dict__.copy()
        new.procedure = list(self.procedure)
        return new


class EvolveConfig:
    def __init__(self, method=EvolveMethod.prop_and_compress, adaptive=False, guess_dt=1e-1,
                 adaptive_rtol=5e-4, taylor_order=None, ivp_solver="krylov", rk_solver="C_RK4"):
        if isinstance(method, str):
            method = EvolveMethod[method]
        self.method = method
        self.adaptive = adaptive
        from .rk import RungeKutta
        self.rk_config = RungeKutta(rk_solver)      # configs.py:363
        if taylor_order is None:                    # configs.py:364-368
            taylor_order = 5 if adaptive else 4
        self.taylor_order = taylor_order
        self.guess_dt = guess_dt
        self.adaptive_rtol = adaptive_rtol
        self.ivp_solver = ivp_solver
        self.stat = None

    @property
    def is_tdvp(self):
        return self.method not in (EvolveMethod.prop_and_compress, EvolveMethod.prop_and_compress_tdrk4,
                                   EvolveMethod.prop_and_compress_tdrk)

    def check_valid_dt(self, evolve_dt):
        """configs.py:394-402."""
        info = f"in config: {self.guess_dt}, in arg: {evolve_dt}"
        if np.iscomplex(evolve_dt) ^ np.iscomplex(self.guess_dt):
            raise ValueError("real and imag not compatible. " + info)
        if (np.iscomplex(evolve_dt) and evolve_dt.imag * self.guess_dt.imag < 0) or \
                (not np.iscomplex(evolve_dt) and evolve_dt * self.guess_dt < 0):
            raise ValueError("evolve into wrong direction. " + info)

    def copy(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__ = self.__dict__.copy()
        return new
