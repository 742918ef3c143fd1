"""Generate tests/golden/exchange.npz from the reference's own Python for the water-exchange path:

  * per-molecule energies and log weights through `nonbonded_block_unsummed` (timemachine/potentials/nonbonded.py:82-150),
    exactly as BDExchangeMove.U_fn_unsummed / batch_log_weights use it (timemachine/md/exchange/exchange_mover.py:112-152);
  * `get_water_groups`, `compute_proposal_probabilities_given_counts`, `compute_raw_ratio_given_weights`, `delta_r_np`
    (exchange_mover.py:237-323), executed from the reference's source text (the module itself imports jax.random / scipy.stats
    machinery that the numpy stand-in does not provide, so the four pure functions are compiled on their own).

    python tests/golden/make_golden_exchange.py     (here, where /root/reference exists)
"""

import ast
import sys
from pathlib import Path

import numpy as np
from scipy.special import logsumexp

import make_golden as G  # noqa: E402  (same directory)

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from tests.common import water_box  # noqa: E402

OUT = Path(__file__).resolve().parent
BOLTZ = 0.008314462618


def reference_exchange_functions():
    src = (G.REF / "timemachine/md/exchange/exchange_mover.py").read_text()
    tree = ast.parse(src)
    wanted = {"delta_r_np", "get_water_groups", "compute_proposal_probabilities_given_counts", "compute_raw_ratio_given_weights"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert {n.name for n in body} == wanted
    ns = {"np": np, "logsumexp": logsumexp}
    exec(compile(ast.Module(body=body, type_ignores=[]), "exchange_mover.py (reference, 4 functions)", "exec"), ns)
    return ns


def main():
    G.install_jax_standin()
    nb, bd, ju = G.load_reference_modules()
    ex = reference_exchange_functions()
    s = water_box(60, seed=5)
    x, params, box = s["x"], s["params"], s["box"]
    rng = np.random.default_rng(5)
    x = x + rng.integers(-1, 2, (len(x), 1)) * box[0, 0] * (rng.random((len(x), 1)) < 0.2)  # some atoms in other images
    params = params.copy()
    params[7 * 3 : 9 * 3, 3] = 0.3  # two molecules lifted into the 4th dimension
    n = len(x)
    mols = np.arange(n).reshape(-1, 3)
    beta, cutoff, temperature = 2.0, 1.2, 300.0
    # U_fn_unsummed(conf, box, a_idxs, b_idxs) -> [3, n - 3] pair energies, NaN -> inf (exchange_mover.py:112-126)
    mol_energy = []
    pair_rows = None
    for m, a_idxs in enumerate(mols):
        b_idxs = np.delete(np.arange(n), a_idxs)
        u = np.asarray(nb.nonbonded_block_unsummed(G.J(x[a_idxs]), G.J(x[b_idxs]), G.J(box), G.J(params[a_idxs]), G.J(params[b_idxs]), beta, cutoff))
        u = np.where(np.isnan(u), np.inf, u)
        mol_energy.append(float(np.sum(u)))
        if m == 4:
            pair_rows = u
    mol_energy = np.array(mol_energy)
    log_weights = mol_energy / (BOLTZ * temperature)
    center = x[:6].mean(0)
    radius = 0.5
    inner, outer = ex["get_water_groups"](x, box, center, mols, radius)
    # a proposal: molecule `moved` goes from the outer to the inner region; weights after = some perturbed weights
    after = log_weights + rng.normal(0, 0.5, len(log_weights))
    moved = int(outer[3])
    vol_inner = 4.0 / 3.0 * np.pi * radius**3
    vol_outer = np.prod(np.diag(box)) - vol_inner
    raw_in = ex["compute_raw_ratio_given_weights"](log_weights[outer], after[np.append(inner, moved)], outer, inner, vol_outer, vol_inner)
    raw_out = ex["compute_raw_ratio_given_weights"](
        log_weights[inner], after[np.append(outer, inner[0])], inner, outer, vol_inner, vol_outer
    )
    # an empty destination region (proposal probability 1 instead of 1/2) and a single-molecule source
    raw_empty_dest = ex["compute_raw_ratio_given_weights"](log_weights[:5], after[:1], np.arange(5), np.arange(0), 2.0, 3.0)
    raw_single_src = ex["compute_raw_ratio_given_weights"](log_weights[:1], after[:4], np.arange(1), np.arange(3), 2.0, 3.0)
    np.savez(
        OUT / "exchange.npz", x=x, params=params, box=box, mols=mols, beta=beta, cutoff=cutoff, temperature=temperature,
        mol_energy=mol_energy, log_weights=log_weights, pair_rows_mol4=pair_rows, center=center, radius=radius, inner=inner, outer=outer,
        after=after, moved=moved, vol_inner=vol_inner, vol_outer=vol_outer, raw_in=raw_in, raw_out=raw_out, raw_empty_dest=raw_empty_dest,
        raw_single_src=raw_single_src, delta_r=ex["delta_r_np"](x[:10], x[10:20], box),
    )
    print("exchange.npz:", mol_energy[:3], len(inner), len(outer), raw_in, raw_out, raw_empty_dest, raw_single_src)


if __name__ == "__main__":
    main()
