"""Generate tests/golden/log_flat_bottom_bond.npz from the reference's own `log_flat_bottom_bond`
(timemachine/potentials/bonded.py:245-253), imported unmodified with make_golden.py's numpy stand-in for jax.

    python tests/golden/make_golden_logfb.py     (here, where /root/reference exists)
"""

from pathlib import Path

import numpy as np

import make_golden as G  # noqa: E402  (same directory)

OUT = Path(__file__).resolve().parent


def main():
    G.install_jax_standin()
    nb, bd, ju = G.load_reference_modules()
    rng = np.random.default_rng(77)
    n, B = 24, 30
    box = np.eye(3) * 2.5
    x = rng.uniform(0, 2.5, (n, 3))
    idxs = np.array([rng.choice(n, 2, replace=False) for _ in range(B)], dtype=np.int32)
    # every bond outside its flat region (the potential is +inf inside): r_min above or r_max below the distance
    d = x[idxs[:, 0]] - x[idxs[:, 1]]
    d -= 2.5 * np.rint(d / 2.5)
    r = np.linalg.norm(d, axis=1)
    params = np.stack([rng.uniform(50, 2000, B), np.zeros(B), r * rng.uniform(0.3, 0.9, B)], 1)
    flip = rng.random(B) < 0.3  # some bonds are compressed instead: r < r_min
    params[flip, 1] = r[flip] * rng.uniform(1.1, 1.5, flip.sum())
    params[flip, 2] = params[flip, 1] + 0.5
    beta = 1.0 / (0.008314462618 * 300.0)

    def u_of(xx, pp=params):
        return float(bd.log_flat_bottom_bond(G.J(xx), G.J(pp), G.J(box), idxs, beta))

    u = u_of(x)
    np.savez(
        OUT / "log_flat_bottom_bond.npz", x=x, box=box, idxs=idxs, params=params, beta=beta, u=u,
        du_dx_fd=G.fd_grad(u_of, x, h=1e-6), du_dp_fd=G.fd_grad(lambda p: u_of(x, p), params, h=1e-6),
    )
    print(f"log_flat_bottom_bond: u={u:.8f}")


if __name__ == "__main__":
    main()
