"""Host-side mirror of the reference's potential dataclasses for the hot path
(reference: timemachine/potentials/potential.py:20-70, potentials.py:17-304).

Same class names, field order (= custom_ops constructor positional order, potential.py:28-37), `.to_gpu(precision)`,
`.bind(params)`, `GpuImplWrapper.unbound_impl`, `BoundGpuImplWrapper.bound_impl`.  The reference's `__call__` on the
dataclasses evaluates its JAX CPU implementation; that CPU path is deliberately NOT reproduced here (this package has
no CPU fallback) - call `.to_gpu(...)`.
"""

from __future__ import annotations

import warnings
from dataclasses import astuple, dataclass
from typing import Any, Optional, Sequence

import numpy as np

from . import custom_ops

Precision = Any


def get_custom_ops_class_name_suffix(precision: Precision) -> str:
    if precision == np.float32:
        return "f32"
    if precision == np.float64:
        return "f64"
    raise ValueError("invalid precision")


@dataclass
class GpuImplWrapper:
    unbound_impl: custom_ops.Potential

    def __call__(self, conf, params, box) -> float:
        return self.unbound_impl.execute(conf, params, box, False, False, True)[2]

    def bind(self, params) -> "BoundGpuImplWrapper":
        return BoundGpuImplWrapper(custom_ops.BoundPotential(self.unbound_impl, np.asarray(params)))


@dataclass
class BoundGpuImplWrapper:
    bound_impl: custom_ops.BoundPotential

    def __call__(self, conf, box) -> float:
        return self.bound_impl.execute(conf, box, False, True)[1]


@dataclass
class Potential:
    def __call__(self, conf, params, box):
        raise NotImplementedError("timemachine_b200 has no CPU evaluation path; use .to_gpu(precision)")

    def bind(self, params) -> "BoundPotential":
        return BoundPotential(self, params)

    def to_gpu(self, precision: Precision) -> GpuImplWrapper:
        ctor = getattr(custom_ops, f"{type(self).__name__}_{get_custom_ops_class_name_suffix(precision)}")
        # shallow field tuple: dataclasses.astuple would deep-copy / recurse into numpy arrays' containers
        args = tuple(getattr(self, f) for f in self.__dataclass_fields__)
        return GpuImplWrapper(ctor(*args))


@dataclass
class BoundPotential:
    potential: Potential
    params: Any

    def to_gpu(self, precision: Precision) -> BoundGpuImplWrapper:
        return self.potential.to_gpu(precision).bind(np.asarray(self.params))


@dataclass
class HarmonicBond(Potential):
    idxs: np.ndarray


@dataclass
class HarmonicAngle(Potential):
    idxs: np.ndarray


class HarmonicAngleStable(HarmonicAngle):
    """potentials.py:36-46: deprecated name kept so that old pickles load."""

    def __init__(self, *args, **kwargs):
        warnings.warn("HarmonicAngleStable is deprecated and will be removed in a future release.", DeprecationWarning)
        super().__init__(*args, **kwargs)

    def __setstate__(self, state):
        warnings.warn("HarmonicAngleStable is deprecated and will be removed in a future release.", DeprecationWarning)
        self.__dict__ = state

    def __getstate__(self):
        raise NotImplementedError("HarmonicAngleStable is deprecated. Serialization is disabled.")


@dataclass
class PeriodicTorsion(Potential):
    idxs: np.ndarray


@dataclass
class CentroidRestraint(Potential):
    """potentials.py:49-57"""

    group_a_idxs: np.ndarray
    group_b_idxs: np.ndarray
    kb: float
    b0: float


@dataclass
class FlatBottomBond(Potential):
    """potentials.py:77-83"""

    idxs: np.ndarray


@dataclass
class LogFlatBottomBond(Potential):
    """potentials.py:85-91"""

    idxs: np.ndarray
    beta: float


@dataclass
class ChiralAtomRestraint(Potential):
    """potentials.py:60-66"""

    idxs: np.ndarray


@dataclass
class ChiralBondRestraint(Potential):
    """potentials.py:68-75"""

    idxs: np.ndarray
    signs: np.ndarray


@dataclass
class NonbondedPairListPrecomputed(Potential):
    """potentials.py:218-237: per-pair parameters (q_ij, sig_ij, eps_ij, w_offset_ij), combining rules already applied."""

    idxs: np.ndarray
    beta: float
    cutoff: float


@dataclass
class NonbondedAllPairs(Potential):
    num_atoms: int
    beta: float
    cutoff: float
    atom_idxs: Optional[np.ndarray] = None
    disable_hilbert_sort: bool = False
    nblist_padding: float = 0.1


@dataclass
class NonbondedInteractionGroup(Potential):
    num_atoms: int
    row_atom_idxs: np.ndarray
    beta: float
    cutoff: float
    col_atom_idxs: Optional[np.ndarray] = None
    disable_hilbert_sort: bool = False
    nblist_padding: float = 0.1


@dataclass
class NonbondedPairList(Potential):
    idxs: np.ndarray
    rescale_mask: np.ndarray
    beta: float
    cutoff: float


@dataclass
class NonbondedExclusions(Potential):
    idxs: np.ndarray
    rescale_mask: np.ndarray
    beta: float
    cutoff: float


def filter_exclusions(atom_idxs, exclusion_idxs, scale_factors):
    """Drop exclusions that touch atoms outside atom_idxs (reference potentials/nonbonded.py:176-218)."""
    keep_set = set(int(a) for a in atom_idxs)
    exclusion_idxs = np.asarray(exclusion_idxs, dtype=np.int32).reshape(-1, 2)
    scale_factors = np.asarray(scale_factors, dtype=np.float64).reshape(-1, 2)
    keep = np.array([int(i) in keep_set and int(j) in keep_set for i, j in exclusion_idxs], dtype=bool)
    return exclusion_idxs[keep], scale_factors[keep]


@dataclass
class Nonbonded(Potential):
    num_atoms: int
    exclusion_idxs: np.ndarray
    scale_factors: np.ndarray
    beta: float
    cutoff: float
    atom_idxs: Optional[np.ndarray] = None
    disable_hilbert_sort: bool = False
    nblist_padding: float = 0.1

    def to_gpu(self, precision: Precision) -> GpuImplWrapper:
        """All pairs + negated exclusions under one fan-out (reference potentials.py:126-138)."""
        all_pairs = NonbondedAllPairs(
            self.num_atoms, self.beta, self.cutoff, atom_idxs=self.atom_idxs, disable_hilbert_sort=self.disable_hilbert_sort,
            nblist_padding=self.nblist_padding,
        )
        atom_idxs = self.atom_idxs if self.atom_idxs is not None else np.arange(self.num_atoms, dtype=np.int32)
        exclusion_idxs, scale_factors = filter_exclusions(atom_idxs, self.exclusion_idxs, self.scale_factors)
        exclusions = NonbondedExclusions(exclusion_idxs, scale_factors, self.beta, self.cutoff)
        return FanoutSummedPotential([all_pairs, exclusions]).to_gpu(precision)


@dataclass
class SummedPotentialGpuImplWrapper(GpuImplWrapper):
    def call_with_params_list(self, conf, params: Sequence[np.ndarray], box) -> float:
        return self(conf, np.concatenate([np.asarray(ps).reshape(-1) for ps in params]), box)

    def bind_params_list(self, params: Sequence[np.ndarray]) -> BoundGpuImplWrapper:
        flat = np.concatenate([np.asarray(ps).reshape(-1) for ps in params])
        return BoundGpuImplWrapper(custom_ops.BoundPotential(self.unbound_impl, flat))


@dataclass
class SummedPotential(Potential):
    potentials: Sequence[Potential]
    params_init: Sequence[np.ndarray]
    parallel: bool = True

    def __post_init__(self):
        if len(self.potentials) != len(self.params_init):
            raise ValueError("number of potentials != number of parameter arrays")

    def to_gpu(self, precision: Precision) -> SummedPotentialGpuImplWrapper:
        impls = [p.to_gpu(precision).unbound_impl for p in self.potentials]
        sizes = [int(np.asarray(ps).size) for ps in self.params_init]
        return SummedPotentialGpuImplWrapper(custom_ops.SummedPotential(impls, sizes, self.parallel))

    def bind_params_list(self, params: Sequence[np.ndarray]) -> BoundPotential:
        return BoundPotential(self, np.concatenate([np.asarray(ps).reshape(-1) for ps in params]))

    @property
    def params_shapes(self):
        return [np.asarray(ps).shape for ps in self.params_init]


def make_summed_potential(bps: Sequence[BoundPotential]) -> BoundPotential:
    potentials = [bp.potential for bp in bps]
    params = [np.asarray(bp.params) for bp in bps]
    return SummedPotential(potentials, params).bind_params_list(params)


@dataclass
class FanoutSummedPotential(Potential):
    potentials: Sequence[Potential]
    parallel: bool = True

    def to_gpu(self, precision: Precision) -> GpuImplWrapper:
        impls = [p.to_gpu(precision).unbound_impl for p in self.potentials]
        return GpuImplWrapper(custom_ops.FanoutSummedPotential(impls, self.parallel))
