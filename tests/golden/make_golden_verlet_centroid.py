"""Generate tests/golden/verlet_centroid.npz from the reference's own Python code:

* `centroid_restraint` (timemachine/potentials/bonded.py:8-31), imported unmodified with make_golden.py's numpy stand-in
  for jax: energies for b0 != 0 and b0 == 0, central finite differences of them as the gradient reference;
* `VelocityVerletIntegrator.multiple_steps` (timemachine/integrator.py:153-201): the class is exec'd from the reference's
  source text (the module imports jax.random and custom_ops at module scope) with lib/fixed_point.py's two helpers evaluated
  literally on numpy, and run on an anharmonic bonded chain.

    python tests/golden/make_golden_verlet_centroid.py     (here, where /root/reference exists)
"""

import sys
from pathlib import Path

import numpy as np

import make_golden as G  # noqa: E402  (same directory)

OUT = Path(__file__).resolve().parent
FIXED_EXPONENT = 0x1000000000


def main():
    G.install_jax_standin()
    nb, bd, ju = G.load_reference_modules()
    rng = np.random.default_rng(2024)

    # ---- centroid restraint ------------------------------------------------------------------------------------------
    n = 40
    x = rng.uniform(0, 3.0, (n, 3))
    group_a = rng.choice(n, 9, replace=False).astype(np.int32)
    group_b = np.setdiff1d(np.arange(n), group_a)[:14].astype(np.int32)
    kb, b0 = 123.4, 0.35
    box = np.eye(3) * 3.0
    out = {"x": x, "group_a": group_a, "group_b": group_b, "kb": kb, "b0": b0}
    for tag, b in (("b0", b0), ("zero", 0.0)):

        def u_of(xx, b=b):
            return float(bd.centroid_restraint(G.J(xx), None, G.J(box), group_a, group_b, kb, b))

        out[f"u_{tag}"] = u_of(x)
        out[f"du_dx_fd_{tag}"] = G.fd_grad(u_of, x, h=1e-6)
        print(f"centroid_restraint[{tag}]: u={out[f'u_{tag}']:.10f}")

    # ---- velocity Verlet --------------------------------------------------------------------------------------------
    src = (G.REF / "timemachine/integrator.py").read_text()
    start = src.index("class Integrator(ABC)")
    end = src.index("def _fori_steps")
    from abc import ABC, abstractmethod

    def fixed_to_float(v):  # lib/fixed_point.py:8-11
        return np.float64(np.int64(np.uint64(v))) / FIXED_EXPONENT

    def float_to_fixed(v):  # lib/fixed_point.py:13-16 (jnp.int64 truncates toward zero)
        return np.uint64(np.int64(np.asarray(v) * FIXED_EXPONENT))

    ns = {
        "np": np, "jnp": sys.modules["jax.numpy"], "jax": sys.modules["jax"], "ABC": ABC, "abstractmethod": abstractmethod,
        "fixed_to_float": fixed_to_float, "float_to_fixed": float_to_fixed, "Any": __import__("typing").Any,
        "Optional": __import__("typing").Optional,
        "time": __import__("time"), "partial": __import__("functools").partial, "jrandom": None, "BOLTZ": 0.0083144626,
        "langevin_coefficients": None,
    }
    exec(src[start:end], ns)
    m = 12
    masses = rng.uniform(1.0, 16.0, m)
    x0 = np.cumsum(rng.normal(0.1, 0.02, (m, 3)), axis=0)
    v0 = rng.normal(0, 0.5, (m, 3))
    k2, k4, r0 = 3.0e4, 5.0e5, 0.15

    def force(xx):  # chain of quartic-harmonic springs: F = -dU/dx, U = sum k2/2 (r - r0)^2 + k4/4 (r - r0)^4
        xx = np.asarray(xx, dtype=np.float64)
        d = xx[1:] - xx[:-1]
        r = np.linalg.norm(d, axis=1)
        du_dr = k2 * (r - r0) + k4 * (r - r0) ** 3
        g = (du_dr / r)[:, None] * d
        f = np.zeros_like(xx)
        f[1:] -= g
        f[:-1] += g
        return f

    dt, n_steps = 1.5e-3, 25
    intg = ns["VelocityVerletIntegrator"](force, masses, dt)
    with np.errstate(over="ignore"):
        xs, vs = intg.multiple_steps(x0, v0, n_steps=n_steps)
    out.update(vv_masses=masses, vv_x0=x0, vv_v0=v0, vv_k2=k2, vv_k4=k4, vv_r0=r0, vv_dt=dt, vv_n_steps=n_steps,
               vv_xs=np.asarray(xs), vv_vs=np.asarray(vs))
    print(f"velocity verlet: |x_T - x_0| = {np.abs(xs[-1] - xs[0]).max():.6f}")
    np.savez(OUT / "verlet_centroid.npz", **out)


if __name__ == "__main__":
    main()
