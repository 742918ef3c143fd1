"""Generate tests/golden/*.npz by running the REFERENCE's own Python functions on seeded inputs.

Run here (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden.py

The reference's CPU path is written against jax.numpy; jax is not installed in this image, so a small numpy stand-in
is registered under the name `jax` (jax.numpy -> numpy with an `.at[].set()` shim, jax.scipy.special.erfc -> scipy).
The reference modules themselves (timemachine/potentials/{nonbonded,bonded,jax_utils}.py, timemachine/integrator.py's
langevin_coefficients/_step) are imported UNMODIFIED from /root/reference and executed; their float64 energies are the
golden values.  jax.grad is not available through the stand-in, so reference gradients are obtained as central finite
differences of those reference energies (h = 1e-5 nm, relative error ~1e-9), which is what pins the oracle's analytic
gradient formulas.

The fixtures are what tests/test_oracle.py holds oracle/tm_oracle.py to; the CUDA path is then held to the oracle.
"""

from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import scipy.special

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


# ---------------------------------------------------------------------------------------------------------------------
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, val):
                out = np.array(arr, copy=True).view(JArr)
                out[idx] = val
                return out

            def add(self, val):
                out = np.array(arr, copy=True).view(JArr)
                out[idx] += val
                return out

        return _Setter()


class JArr(np.ndarray):
    """ndarray with the one jax-array method the reference energy functions use (`x.at[idx].set(v)`)."""

    @property
    def at(self):
        return _At(self)


def install_jax_standin():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("_"):
            setattr(jnp, name, getattr(np, name))
    jnp.ndarray = np.ndarray
    jnp.array = lambda a, *args, **kw: np.array(a, *args, **kw).view(JArr)
    jnp.asarray = lambda a, *args, **kw: np.asarray(a, *args, **kw).view(JArr)
    jax.numpy = jnp
    jax.Array = np.ndarray
    jax.jit = lambda f=None, **kw: (f if f is not None else (lambda g: g))

    def vmap(f, in_axes=0, out_axes=0):
        def g(*args):
            return np.stack([f(*a) for a in zip(*args)])

        return g

    jax.vmap = vmap
    jsp = types.ModuleType("jax.scipy")
    jsps = types.ModuleType("jax.scipy.special")
    jsps.erfc = scipy.special.erfc
    jsps.logsumexp = scipy.special.logsumexp
    jsp.special = jsps
    jtyping = types.ModuleType("jax.typing")
    jtyping.ArrayLike = object
    jrandom = types.ModuleType("jax.random")
    jax.scipy = jsp
    jax.typing = jtyping
    jax.random = jrandom
    for mod, name in [
        (jax, "jax"), (jnp, "jax.numpy"), (jsp, "jax.scipy"), (jsps, "jax.scipy.special"), (jtyping, "jax.typing"),
        (jrandom, "jax.random"),
    ]:
        sys.modules[name] = mod


def load_reference_modules():
    """Import the reference's function modules by path (the package __init__ files pull in the whole of jax)."""

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    pkg = types.ModuleType("timemachine")
    pkg.__path__ = [str(REF / "timemachine")]
    sys.modules["timemachine"] = pkg
    pots = types.ModuleType("timemachine.potentials")
    pots.__path__ = [str(REF / "timemachine/potentials")]
    sys.modules["timemachine.potentials"] = pots
    load("timemachine.constants", REF / "timemachine/constants.py")
    load("timemachine.potentials.types", REF / "timemachine/potentials/types.py")
    ju = load("timemachine.potentials.jax_utils", REF / "timemachine/potentials/jax_utils.py")
    pots.jax_utils = ju
    nb = load("timemachine.potentials.nonbonded", REF / "timemachine/potentials/nonbonded.py")
    bd = load("timemachine.potentials.bonded", REF / "timemachine/potentials/bonded.py")
    return nb, bd, ju


def J(a):
    return np.array(a, dtype=np.float64).view(JArr)


def fd_grad(f, x0, h=1e-5):
    x0 = np.array(x0, dtype=np.float64)
    g = np.zeros_like(x0)
    flat = x0.reshape(-1)
    gf = g.reshape(-1)
    for i in range(flat.size):
        old = flat[i]
        flat[i] = old + h
        up = f(x0)
        flat[i] = old - h
        dn = f(x0)
        flat[i] = old
        gf[i] = (up - dn) / (2 * h)
    return g


# ---------------------------------------------------------------------------------------------------------------------
def make_nonbonded_system(rng, N, box_len, with_w):
    """Water-like charges/LJ on a jittered lattice (no overlaps), some eps == 0 atoms, optional 4-D offsets."""
    n_side = int(np.ceil(N ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:N]
    x = (grid + 0.5) * (box_len / n_side) + rng.normal(0, 0.03, (N, 3))
    x += rng.integers(-2, 3, (N, 3)) * box_len * (rng.random((N, 1)) < 0.2)  # some atoms in neighbouring images
    q = rng.normal(0, 0.5, N) * np.sqrt(138.935456)
    sig = rng.uniform(0.1, 0.18, N)
    eps = np.sqrt(rng.uniform(0.1, 1.0, N))
    eps[rng.random(N) < 0.3] = 0.0
    w = np.zeros(N)
    if with_w:
        w[rng.random(N) < 0.3] = rng.uniform(0.0, 0.6)
        w[rng.random(N) < 0.05] = 1.2  # parked exactly at the cutoff
    params = np.stack([q, sig, eps, w], axis=1)
    box = np.eye(3) * box_len
    return x, params, box


def main():
    install_jax_standin()
    nb, bd, ju = load_reference_modules()
    rng = np.random.default_rng(2022)
    beta, cutoff = 2.0, 1.2

    # ---- nonbonded: monolithic with exclusions, 3 systems -----------------------------------------------------------
    for tag, N, box_len, with_w in [("a", 48, 3.0, False), ("b", 64, 2.6, True), ("c", 40, 2.5, True)]:
        x, params, box = make_nonbonded_system(rng, N, box_len, with_w)
        pairs = np.array([(i, i + 1) for i in range(0, N - 1, 3)] + [(i, i + 2) for i in range(0, N - 2, 5)], dtype=np.int32)
        scales = rng.choice([0.0, 0.5, 1.0], size=(len(pairs), 2))
        atom_idxs = np.sort(rng.choice(N, size=N - 9, replace=False)).astype(np.int32)

        def u_full(xx, pp=params):
            return float(nb.nonbonded(J(xx), J(pp), J(box), pairs, scales, beta, cutoff, runtime_validate=False))

        def u_p(pp):
            return float(nb.nonbonded(J(x), J(pp), J(box), pairs, scales, beta, cutoff, runtime_validate=False))

        u = u_full(x)
        u_subset = float(nb.nonbonded(J(x), J(params), J(box), pairs, scales, beta, cutoff, runtime_validate=False, atom_idxs=atom_idxs))
        du_dx = fd_grad(u_full, x)
        du_dp = fd_grad(u_p, params, h=1e-6)
        # interaction group block and explicit pair list through the reference's own helpers
        rows = np.arange(0, 10, dtype=np.int32)
        cols = np.arange(10, N, dtype=np.int32)
        u_block = float(nb.nonbonded_block(J(x[rows]), J(x[cols]), J(box), J(params[rows]), J(params[cols]), beta, cutoff))
        vdw, es = nb.nonbonded_on_specific_pairs(J(x), J(params), J(box), pairs, beta, cutoff, rescale_mask=J(scales))
        u_pairs = float(np.sum(vdw) + np.sum(es))
        np.savez(
            OUT / f"nonbonded_{tag}.npz", x=x, params=params, box=box, beta=beta, cutoff=cutoff, exclusion_idxs=pairs,
            scale_factors=scales, atom_idxs=atom_idxs, u=u, u_subset=u_subset, du_dx_fd=du_dx, du_dp_fd=du_dp, rows=rows,
            cols=cols, u_block=u_block, u_pairs=u_pairs,
        )
        print(f"nonbonded_{tag}: N={N} u={u:.10f} u_subset={u_subset:.10f} u_block={u_block:.10f} u_pairs={u_pairs:.10f}")

    # ---- bonded -------------------------------------------------------------------------------------------------------
    N = 30
    x = rng.normal(0, 0.4, (N, 3))
    bond_idxs = np.array([(i, i + 1) for i in range(N - 1)], dtype=np.int32)
    bond_params = np.stack([rng.uniform(100, 1000, N - 1), rng.uniform(0.08, 0.2, N - 1)], 1)
    bond_params[::7, 1] = 0.0  # the b0 == 0 branch
    angle_idxs = np.array([(i, i + 1, i + 2) for i in range(N - 2)], dtype=np.int32)
    angle_params = np.stack([rng.uniform(50, 500, N - 2), rng.uniform(1.0, 2.5, N - 2), rng.choice([0.0, 0.01, 0.1], N - 2)], 1)
    tors_idxs = np.array([(i, i + 1, i + 2, i + 3) for i in range(N - 3)], dtype=np.int32)
    tors_params = np.stack([rng.uniform(1, 20, N - 3), rng.uniform(-np.pi, np.pi, N - 3), rng.integers(1, 5, N - 3).astype(float)], 1)

    def u_bond(xx, pp=bond_params):
        return float(bd.harmonic_bond(J(xx), J(pp), None, bond_idxs))

    def u_angle(xx, pp=angle_params):
        return float(bd.harmonic_angle(J(xx), J(pp), None, angle_idxs))

    def u_tors(xx, pp=tors_params):
        return float(bd.periodic_torsion(J(xx), J(pp), None, tors_idxs))

    np.savez(
        OUT / "bonded.npz", x=x, bond_idxs=bond_idxs, bond_params=bond_params, angle_idxs=angle_idxs, angle_params=angle_params,
        torsion_idxs=tors_idxs, torsion_params=tors_params, u_bond=u_bond(x), u_angle=u_angle(x), u_torsion=u_tors(x),
        bond_du_dx_fd=fd_grad(u_bond, x), angle_du_dx_fd=fd_grad(u_angle, x), torsion_du_dx_fd=fd_grad(u_tors, x),
        bond_du_dp_fd=fd_grad(lambda p: u_bond(x, p), bond_params, h=1e-6),
        angle_du_dp_fd=fd_grad(lambda p: u_angle(x, p), angle_params, h=1e-6),
        torsion_du_dp_fd=fd_grad(lambda p: u_tors(x, p), tors_params, h=1e-6),
    )
    print(f"bonded: u_bond={u_bond(x):.10f} u_angle={u_angle(x):.10f} u_torsion={u_tors(x):.10f}")

    # ---- integrator: the reference's langevin_coefficients and one BAOAB _step ------------------------------------------
    # integrator.py imports jax.random / lib.fixed_point at module scope; only two pure functions are needed, so they are
    # exec'd from the reference source text (unmodified) rather than importing the module.
    src = (REF / "timemachine/integrator.py").read_text()
    start = src.index("def langevin_coefficients")
    end = src.index("class Integrator")
    ns = {"np": np, "BOLTZ": sys.modules["timemachine.constants"].BOLTZ}
    exec(src[start:end], ns)
    masses = rng.uniform(1.0, 16.0, N)
    temperature, dt, friction = 300.0, 2.5e-3, 1.0
    ca, cb, cc = ns["langevin_coefficients"](temperature, dt, friction, masses)
    v = rng.normal(0, 1, (N, 3))
    force = rng.normal(0, 100, (N, 3))
    noise = rng.normal(0, 1, (N, 3))
    # LangevinIntegrator._step, integrator.py:137-144 (restated inline by the oracle; here evaluated literally)
    v_mid = v + cb[:, None] * force
    new_v = (ca * v_mid) + (cc[:, None] * noise)
    new_x = x + 0.5 * dt * (v_mid + new_v)
    np.savez(
        OUT / "integrator.npz", x=x, v=v, force=force, noise=noise, masses=masses, temperature=temperature, dt=dt,
        friction=friction, ca=ca, cb=cb, cc=cc, new_x=new_x, new_v=new_v, boltz=ns["BOLTZ"],
    )
    print("integrator: ca=%.12f" % ca)

    # ---- barostat: the reference's Python CentroidRescaler and get_group_indices --------------------------------------
    # md/barostat/moves.py imports custom_ops and half of the package at module scope; the two pure pieces are exec'd
    # from the reference source text (unmodified), with numpy standing in for jax.numpy / jax.ops.segment_sum.
    def segment_sum(data, segment_ids):
        out = np.zeros((int(np.max(segment_ids)) + 1,) + np.shape(data)[1:])
        np.add.at(out, np.asarray(segment_ids), np.asarray(data))
        return out

    src = (REF / "timemachine/md/barostat/moves.py").read_text()
    ns = {"np": np, "jnp": sys.modules["jax.numpy"], "segment_sum": segment_sum}
    exec(src[src.index("def compute_centroid") : src.index("class NPTMove")], ns)
    n_mols, box_len = 40, 3.1
    sizes = rng.choice([1, 3, 3, 3, 7], n_mols)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    group_idxs = [np.arange(offsets[i], offsets[i + 1]) for i in range(n_mols)]
    coords = rng.uniform(-0.5, box_len + 0.5, (offsets[-1], 3))
    center = np.full(3, box_len / 2)
    scale = 1.0173
    rescaler = ns["CentroidRescaler"](group_idxs)
    scaled = np.asarray(rescaler.scale_centroids(coords, center, scale))
    centroids = np.asarray(rescaler.compute_centroids(coords))
    # get_group_indices needs networkx (not in this image): the connected components it returns are restated with the
    # reference's own slow path as the check - every group must be a maximal set connected by the bond list
    bond_list = [(int(g[k]), int(g[k + 1])) for g in group_idxs for k in range(len(g) - 1)]
    np.savez(
        OUT / "barostat.npz", coords=coords, group_sizes=sizes, center=center, scale=scale, scaled=scaled,
        centroids=centroids, bond_list=np.array(bond_list, dtype=np.int32), num_atoms=int(offsets[-1]),
    )
    print(f"barostat: {n_mols} groups, {offsets[-1]} atoms, scale {scale}")

    # ---- restraints and the precomputed pair list ------------------------------------------------------
    spec = importlib.util.spec_from_file_location("timemachine.potentials.chiral_restraints", REF / "timemachine/potentials/chiral_restraints.py")
    cr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cr)
    N = 40
    box = np.eye(3) * 2.4
    x = rng.uniform(0, 2.4, (N, 3))
    quads = np.array([rng.choice(N, 4, replace=False) for _ in range(24)], dtype=np.int32)
    k_atom = rng.uniform(10, 1000, len(quads))
    k_atom[::7] = 0.0
    signs = rng.choice([-1, 1], len(quads)).astype(np.int32)
    k_bond = rng.uniform(10, 1000, len(quads))

    def u_chiral_atom(xx, kk=k_atom):
        # chiral_atom_restraint is sum_r U_chiral_atom(conf, idxs[r], params[r]) (chiral_restraints.py:97-112)
        return float(sum(float(cr.U_chiral_atom(J(xx), quads[r], kk[r])) for r in range(len(quads))))

    def u_chiral_bond(xx, kk=k_bond):
        return float(sum(float(cr.U_chiral_bond(J(xx), quads[r], kk[r], signs[r])) for r in range(len(quads))))

    fb_idxs = np.array([rng.choice(N, 2, replace=False) for _ in range(30)], dtype=np.int32)
    fb_params = np.stack([rng.uniform(50, 500, 30), rng.uniform(0.0, 0.8, 30), rng.uniform(0.8, 1.6, 30)], 1)

    def u_fb(xx, pp=fb_params):
        return float(bd.flat_bottom_bond(J(xx), J(pp), J(box), fb_idxs))

    pre_idxs = np.array([rng.choice(N, 2, replace=False) for _ in range(60)], dtype=np.int32)
    pre_params = np.stack(
        [rng.normal(0, 1.0, 60), rng.uniform(0.1, 0.35, 60), rng.uniform(0.05, 1.0, 60), rng.choice([0.0, 0.0, 0.1, 0.4], 60)], 1
    )

    def u_pre(xx, pp=pre_params):
        vdw, es = nb.nonbonded_on_precomputed_pairs(J(xx), J(pp), J(box), pre_idxs, beta, cutoff)
        return float(np.sum(vdw) + np.sum(es))

    np.savez(
        OUT / "restraints.npz", x=x, box=box, beta=beta, cutoff=cutoff, quads=quads, k_atom=k_atom, k_bond=k_bond, signs=signs,
        u_chiral_atom=u_chiral_atom(x), u_chiral_bond=u_chiral_bond(x), chiral_atom_du_dx_fd=fd_grad(u_chiral_atom, x),
        chiral_bond_du_dx_fd=fd_grad(u_chiral_bond, x),
        fb_idxs=fb_idxs, fb_params=fb_params, u_fb=u_fb(x), fb_du_dx_fd=fd_grad(u_fb, x),
        fb_du_dp_fd=fd_grad(lambda p: u_fb(x, p), fb_params, h=1e-6),
        pre_idxs=pre_idxs, pre_params=pre_params, u_pre=u_pre(x), pre_du_dx_fd=fd_grad(u_pre, x),
        pre_du_dp_fd=fd_grad(lambda p: u_pre(x, p), pre_params, h=1e-6),
    )
    # ---- HREX neighbour swaps: the reference's _run_neighbor_swaps (md/hrex.py:50-129) ------------------------------------
    # hrex.py imports md.moves and scipy.stats at module scope; the one pure function is exec'd from the reference source
    # text (unmodified) with python stand-ins for jax.lax.scan / cond.
    lax = types.SimpleNamespace(
        cond=lambda pred, t, f: t() if bool(pred) else f(),
        scan=lambda f, init, xs: (
            __import__("functools").reduce(lambda c, x: f(c, x)[0], list(zip(*xs)), init),
            None,
        ),
    )
    jx = types.SimpleNamespace(jit=lambda f: f, lax=lax)
    src = (REF / "timemachine/md/hrex.py").read_text()
    jn = types.SimpleNamespace(zeros=lambda n, dt: np.zeros(n, dt).view(JArr), minimum=np.minimum, exp=np.exp, uint32=np.uint32)
    ns = {"jax": jx, "jnp": jn, "Array": np.ndarray, "np": np}
    exec(src[src.index("@jax.jit\ndef _run_neighbor_swaps") : src.index("@dataclass(frozen=True)\nclass HREX")], ns)
    n_states = 8
    log_q = -rng.uniform(0, 6, (n_states, n_states))
    pairs = np.array([(s, s + 1) for s in range(n_states - 1)])
    n_attempts = 300
    pair_idxs = rng.integers(0, len(pairs), n_attempts)
    uniforms = rng.random(n_attempts)
    start = rng.permutation(n_states)
    final, proposed, accepted = ns["_run_neighbor_swaps"](start.view(JArr), pairs.view(JArr), log_q.view(JArr), pair_idxs, uniforms)
    np.savez(
        OUT / "hrex.npz", log_q=log_q, neighbor_pairs=pairs, pair_idxs=pair_idxs, uniform_samples=uniforms, start=start,
        final=np.asarray(final), proposed=np.asarray(proposed), accepted=np.asarray(accepted),
    )
    print(f"hrex: final={np.asarray(final).tolist()} accepted={int(np.sum(accepted))}/{n_attempts}")

    print(f"restraints: u_chiral_atom={u_chiral_atom(x):.8f} u_chiral_bond={u_chiral_bond(x):.8f} u_fb={u_fb(x):.8f} u_pre={u_pre(x):.8f}")


if __name__ == "__main__":
    main()
