"""Record what the reference's UNMODIFIED Python wrappers ask of `timemachine.lib.custom_ops`.

Run here (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden_dropin.py

`timemachine/potentials/potential.py`, `potentials.py`, `jax_interface.py` and `timemachine/lib/__init__.py` are imported
from /root/reference as they are (with the numpy stand-in for jax of make_golden.py: jax is not installable here), on top
of a RECORDING stand-in for `timemachine.lib.custom_ops`: every class the wrappers look up by name
(`getattr(custom_ops, f"{cls.__name__}_{suffix}")`, potential.py:28-37) and every constructor call (`ctor(*astuple(self))`,
`custom_ops.SummedPotential(impls, sizes, parallel)`, `custom_ops.BoundPotential(impl, params)`,
`custom_ops.LangevinIntegrator(masses, T, dt, friction, seed)`, `custom_ops.MonteCarloBarostat(...)`) is written down
with its arguments.  The trace goes to tests/golden/dropin_calls.{json,npz}:

  * tests/test_dropin_cpu.py checks that this repo's custom_ops module has every class of the trace and that its
    constructor signatures bind the recorded arguments (and, where /root/reference exists, that the trace is current);
  * tests/test_dropin_gpu.py replays the trace against the real module on the GPU - so "the reference's wrappers run
    unchanged on top of this module" is tested on the GPU box without the GPU box ever reading /root/reference.
"""

from __future__ import annotations

import importlib.util
import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF = Path("/root/reference")
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


class Trace:
    def __init__(self):
        self.calls = []   # {"id", "cls", "args": [...], "kwargs": {...}}
        self.arrays = {}  # name -> ndarray

    def encode(self, v, where):
        if isinstance(v, Recorded):
            return {"obj": v._id}
        if isinstance(v, np.ndarray):
            key = f"a{len(self.arrays)}"
            self.arrays[key] = np.array(v)
            return {"array": key, "dtype": str(v.dtype), "shape": list(v.shape)}
        if isinstance(v, (list, tuple)):
            return {"list": [self.encode(x, where) for x in v]}
        if isinstance(v, (bool, np.bool_)):
            return {"bool": bool(v)}
        if isinstance(v, (int, np.integer)):
            return {"int": int(v)}
        if isinstance(v, (float, np.floating)):
            return {"float": float(v)}
        if v is None:
            return {"none": True}
        if isinstance(v, set):
            return {"list": [self.encode(x, where) for x in sorted(v)]}
        raise TypeError(f"{where}: cannot record {type(v)}")


TRACE = Trace()


class Recorded:
    """An object constructed through the recording module."""

    def __init__(self, cls_name, args, kwargs):
        self._id = len(TRACE.calls)
        self._cls = cls_name
        TRACE.calls.append({
            "id": self._id, "cls": cls_name, "args": [TRACE.encode(a, cls_name) for a in args],
            "kwargs": {k: TRACE.encode(v, cls_name) for k, v in kwargs.items()},
        })


def make_recording_module():
    mod = types.ModuleType("timemachine.lib.custom_ops")

    class Potential:  # the annotations in potential.py name these two
        pass

    class BoundPotentialBase:
        pass

    def factory(name):
        def ctor(*args, **kwargs):
            return Recorded(name, args, kwargs)

        ctor.__name__ = name
        return ctor

    mod.Potential = Potential

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return factory(name)

    mod.__getattr__ = __getattr__
    return mod


def install_reference_wrappers():
    """jax stand-in + recording custom_ops + the reference's wrapper modules, unmodified, by path."""
    sys.path.insert(0, str(HERE))
    import make_golden as G

    G.install_jax_standin()
    jax = sys.modules["jax"]
    core = types.ModuleType("jax.core")

    class Tracer:  # nothing is ever traced through the stand-in
        pass

    core.Tracer = Tracer
    jax.core = core
    sys.modules["jax.core"] = core

    class custom_jvp:
        def __init__(self, fn, nondiff_argnums=()):
            self.fn = fn

        def __call__(self, *a, **k):
            return self.fn(*a, **k)

        def defjvp(self, rule):
            self.rule = rule
            return rule

    jax.custom_jvp = custom_jvp
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)

    rec = make_recording_module()
    pkg = types.ModuleType("timemachine")
    pkg.__path__ = [str(REF / "timemachine")]
    sys.modules["timemachine"] = pkg
    sys.modules["timemachine.lib.custom_ops"] = rec

    def load(name, path, is_pkg=False):
        spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[str(Path(path).parent)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("timemachine.constants", REF / "timemachine/constants.py")
    lib = load("timemachine.lib", REF / "timemachine/lib/__init__.py", is_pkg=True)
    pots = load("timemachine.potentials", REF / "timemachine/potentials/__init__.py", is_pkg=True)
    return pots, lib, rec


def scenario(pots, lib):
    """Every potential class of the path, in both precisions, bound and unbound, + the lib dataclasses."""
    from tests.common import water_box

    s = water_box(64, seed=5)
    N = s["N"]
    rng = np.random.default_rng(11)
    beta, cutoff = 2.0, 1.2
    lig = np.arange(N - 9, N, dtype=np.int32)
    env = np.arange(0, N - 9, dtype=np.int32)
    chain = np.array([(i, i + 1, i + 2, i + 3) for i in range(0, 20)], dtype=np.int32)
    pairs = np.array([(i, i + 7) for i in range(0, 40)], dtype=np.int32)
    plist = [
        pots.HarmonicBond(s["bond_idxs"]),
        pots.HarmonicAngle(s["angle_idxs"]),
        pots.PeriodicTorsion(chain),
        pots.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], beta, cutoff),
        pots.Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], beta, cutoff, atom_idxs=env, disable_hilbert_sort=True, nblist_padding=0.05),
        pots.NonbondedAllPairs(N, beta, cutoff),
        pots.NonbondedInteractionGroup(N, lig, beta, cutoff, col_atom_idxs=env),
        pots.NonbondedInteractionGroup(N, lig, beta, cutoff),
        pots.NonbondedPairList(pairs, rng.uniform(0, 1, (len(pairs), 2)), beta, cutoff),
        pots.NonbondedExclusions(s["exclusion_idxs"], s["scale_factors"], beta, cutoff),
        pots.NonbondedPairListPrecomputed(pairs, beta, cutoff),
        pots.FlatBottomBond(pairs),
        pots.LogFlatBottomBond(pairs, 1.0 / (0.008314462618 * 300.0)),
        pots.ChiralAtomRestraint(chain),
        pots.ChiralBondRestraint(chain, np.where(rng.random(len(chain)) < 0.5, -1, 1).astype(np.int32)),
        pots.CentroidRestraint(lig, env[:30], 250.0, 0.4),
    ]
    for p in plist:
        for precision in (np.float32, np.float64):
            p.to_gpu(precision)
    # the composite the free-energy code builds (fe/free_energy.py:1436: one SummedPotential per system) and its binding
    params_init = [s["bond_params"], s["angle_params"], s["params"]]
    summed = pots.SummedPotential([plist[0], plist[1], plist[3]], params_init)
    wrapper = summed.to_gpu(np.float32)
    bound = wrapper.bind_params_list(params_init)
    plist[0].bind(s["bond_params"]).to_gpu(np.float64)
    pots.FanoutSummedPotential([plist[5], plist[9]], parallel=False).to_gpu(np.float32)
    intg = lib.LangevinIntegrator(300.0, 1.5e-3, 1.0, s["masses"], 2024).impl()
    lib.VelocityVerletIntegrator(1.5e-3, s["masses"]).impl()
    groups = [np.arange(i, i + 3) for i in range(0, N, 3)]
    lib.MonteCarloBarostat(N, 1.013, 300.0, groups, 15, 7).impl([bound.bound_impl])
    lib.MonteCarloBarostat(N, 1.013, 300.0, groups, 5, 9, adaptive_scaling_enabled=False, initial_volume_scale_factor=0.02).impl([bound.bound_impl])
    return intg


def record():
    pots, lib, rec = install_reference_wrappers()
    scenario(pots, lib)
    return TRACE


def main():
    trace = record()
    (HERE / "dropin_calls.json").write_text(json.dumps({
        "source": "reference timemachine/potentials/{potential,potentials,jax_interface}.py and lib/__init__.py, executed unmodified by "
                  "tests/golden/make_golden_dropin.py on a recording custom_ops module",
        "calls": trace.calls,
    }, indent=1))
    np.savez_compressed(HERE / "dropin_calls.npz", **trace.arrays)
    names = sorted({c["cls"] for c in trace.calls})
    print(f"{len(trace.calls)} constructor calls, {len(names)} classes: {names}")


if __name__ == "__main__":
    main()
