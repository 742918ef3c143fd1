"""BASELINE.json configs[1]: the reference's DHFR benchmark geometry (tests/golden/dhfr_5dfr.npz, made from
timemachine/testsystems/data/5dfr_solv_equil.pdb by tests/golden/make_golden_dhfr.py) with a topology and protein-like
parameters derived from the geometry - OpenMM's amber99sbildn is not available here.  Deterministic: the same arrays feed
this repo and the compiled reference, so the parity tests are statements about the kernels on the real DHFR atom density,
protein / water mix and atom order."""

from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
ONE_4PI_EPS0 = 138.935456
# sigma / 2 (nm) and eps (kJ/mol) by element, amber-like
LJ = {1: (0.0535, 0.0657), 6: (0.17, 0.36), 7: (0.1625, 0.711), 8: (0.148, 0.879), 16: (0.178, 1.046)}
CHARGE = {1: 0.25, 6: 0.05, 7: -0.45, 8: -0.55, 16: -0.2}
MASS = {1: 1.008, 6: 12.011, 7: 14.007, 8: 15.999, 16: 32.06}


def _wave(idx, scale, salt):
    """Deterministic pseudo-random numbers in [-scale, scale] (no RNG state to keep in step between the two builds)."""
    return scale * np.sin(12.9898 * (np.asarray(idx, dtype=np.float64) + salt) + 78.233 * salt)


def load_dhfr():
    from scipy.spatial import cKDTree

    g = np.load(GOLDEN / "dhfr_5dfr.npz")
    x = g["xyz_milliangstrom"].astype(np.float64) * 1e-4  # nm
    elem, res, water = g["element"].astype(int), g["residue"], g["is_water"]
    L = float(g["box_angstrom"]) * 0.1
    N = len(x)
    box = np.eye(3) * L
    prot = np.flatnonzero(~water)
    wat_o = np.flatnonzero(water & (elem == 8))
    assert len(wat_o) * 3 == int(water.sum())

    # ---- bonds ----------------------------------------------------------------------------------------------------
    bonds = set()
    tree = cKDTree(x[prot])
    for a, b in tree.query_pairs(0.215):
        i, j = int(prot[a]), int(prot[b])
        if abs(int(res[i]) - int(res[j])) > 1:
            if not (elem[i] == 16 and elem[j] == 16):
                continue
        d = np.linalg.norm(x[i] - x[j])
        ei, ej = elem[i], elem[j]
        if ei == 1 and ej == 1:
            continue
        limit = 0.125 if 1 in (ei, ej) else (0.215 if 16 in (ei, ej) else 0.17)
        if d < limit:
            bonds.add((min(i, j), max(i, j)))
    # every hydrogen keeps its nearest partner only
    nearest = {}
    for i, j in bonds:
        for h, o in ((i, j), (j, i)):
            if elem[h] == 1:
                d = np.linalg.norm(x[h] - x[o])
                if h not in nearest or d < nearest[h][0]:
                    nearest[h] = (d, o)
    bonds = {(i, j) for i, j in bonds if not ((elem[i] == 1 and nearest[i][1] != j) or (elem[j] == 1 and nearest[j][1] != i))}
    # an atom the distance rules left alone is tied to its nearest heavy neighbour
    bonded = {a for b in bonds for a in b}
    heavy = prot[elem[prot] != 1]
    heavy_tree = cKDTree(x[heavy])
    for i in prot:
        if int(i) not in bonded:
            _, near = heavy_tree.query(x[i], k=2)
            j = int(heavy[near[0]]) if int(heavy[near[0]]) != int(i) else int(heavy[near[1]])
            bonds.add((min(int(i), j), max(int(i), j)))
    prot_bonds = np.array(sorted(bonds), dtype=np.int32)
    wat_bonds = np.concatenate([np.stack([wat_o, wat_o + 1], 1), np.stack([wat_o, wat_o + 2], 1)]).astype(np.int32)
    bond_idxs = np.concatenate([prot_bonds, wat_bonds]).astype(np.int32)
    d0 = np.linalg.norm(x[bond_idxs[:, 0]] - x[bond_idxs[:, 1]], axis=1)
    kb = np.where((elem[bond_idxs[:, 0]] == 1) | (elem[bond_idxs[:, 1]] == 1), 3.6e5, 2.6e5)
    kb[len(prot_bonds):] = 462750.4
    b0 = d0 * (1.0 + _wave(np.arange(len(d0)), 0.02, 1.0))
    bond_params = np.stack([kb, b0], 1)

    # ---- angles / torsions from the protein bond graph ----------------------------------------------------------------
    nbrs = {}
    for i, j in prot_bonds:
        nbrs.setdefault(int(i), []).append(int(j))
        nbrs.setdefault(int(j), []).append(int(i))
    angles = []
    for j, ns in nbrs.items():
        ns = sorted(ns)
        for a in range(len(ns)):
            for b in range(a + 1, len(ns)):
                angles.append((ns[a], j, ns[b]))
    prot_angles = np.array(sorted(angles), dtype=np.int32)
    wat_angles = np.stack([wat_o + 1, wat_o, wat_o + 2], 1).astype(np.int32)
    angle_idxs = np.concatenate([prot_angles, wat_angles]).astype(np.int32)
    v1 = x[angle_idxs[:, 0]] - x[angle_idxs[:, 1]]
    v2 = x[angle_idxs[:, 2]] - x[angle_idxs[:, 1]]
    th = np.arccos(np.clip(np.sum(v1 * v2, 1) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1)), -1, 1))
    ka = np.full(len(th), 480.0)
    ka[len(prot_angles):] = 836.8
    angle_params = np.stack([ka, th + _wave(np.arange(len(th)), 0.06, 2.0), np.zeros(len(th))], 1)

    propers = []
    for j, k in prot_bonds:
        j, k = int(j), int(k)
        for i in nbrs[j]:
            if i == k:
                continue
            for l in nbrs[k]:
                if l == j or l == i:
                    continue
                propers.append((i, j, k, l))
    proper_idxs = np.array(sorted(propers), dtype=np.int32)
    n_p = len(proper_idxs)
    proper_params = np.stack([
        4.0 + _wave(np.arange(n_p), 3.5, 3.0), np.where(_wave(np.arange(n_p), 1.0, 4.0) > 0, np.pi, 0.0),
        1.0 + np.floor(1.5 + _wave(np.arange(n_p), 1.49, 5.0)),
    ], 1)
    impropers = [(ns[0], ns[1], c, ns[2]) for c, ns in sorted(nbrs.items()) if len(ns) == 3 and elem[c] in (6, 7)]
    improper_idxs = np.array(impropers, dtype=np.int32)
    n_i = len(improper_idxs)
    improper_params = np.stack([np.full(n_i, 43.9), np.full(n_i, np.pi), np.full(n_i, 2.0)], 1)

    # ---- nonbonded parameters (timemachine encoding: q sqrt(k_e), sigma / 2, sqrt(eps), w) -----------------------------------
    q = np.array([CHARGE[e] for e in elem], dtype=np.float64) + _wave(np.arange(N), 0.08, 6.0)
    for r in np.unique(res[prot]):  # neutral residues
        m = np.flatnonzero(res == r)
        q[m] -= q[m].mean()
    sig = np.array([LJ[e][0] for e in elem])
    eps = np.array([LJ[e][1] for e in elem])
    q[wat_o], q[wat_o + 1], q[wat_o + 2] = -0.834, 0.417, 0.417
    sig[wat_o], eps[wat_o] = 0.1575375, 0.635968
    for h in (wat_o + 1, wat_o + 2):
        sig[h], eps[h] = 0.05, 0.0
    params = np.stack([q * np.sqrt(ONE_4PI_EPS0), sig, np.sqrt(eps), np.zeros(N)], 1)

    # ---- exclusions: 1-2 and 1-3 in full, 1-4 scaled (amber: charges / 1.2, LJ / 2) ----------------------------------------
    excl = {}
    for i, j in prot_bonds:
        excl[(int(i), int(j))] = (1.0, 1.0)
    for i, _, k in prot_angles:
        excl[(min(int(i), int(k)), max(int(i), int(k)))] = (1.0, 1.0)
    for i, _, _, l in proper_idxs:
        key = (min(int(i), int(l)), max(int(i), int(l)))
        if key not in excl and key[0] != key[1]:
            excl[key] = (1.0 - 1.0 / 1.2, 0.5)
    for o in wat_o:
        excl[(int(o), int(o) + 1)] = (1.0, 1.0)
        excl[(int(o), int(o) + 2)] = (1.0, 1.0)
        excl[(int(o) + 1, int(o) + 2)] = (1.0, 1.0)
    keys = sorted(excl)
    exclusion_idxs = np.array(keys, dtype=np.int32)
    scale_factors = np.array([excl[k] for k in keys], dtype=np.float64)

    masses = np.array([MASS[e] for e in elem])
    # hydrogen mass repartitioning like the reference's benchmark (tests/test_benchmark.py: apply_hmr): H x 3 from its partner
    for a, b in bond_idxs:
        h, o = (a, b) if elem[a] == 1 else ((b, a) if elem[b] == 1 else (None, None))
        if h is not None:
            masses[h] += 2 * 1.008
            masses[o] -= 2 * 1.008
    return dict(
        N=N, x=x, box=box, bond_idxs=bond_idxs, bond_params=bond_params, angle_idxs=angle_idxs, angle_params=angle_params,
        proper_idxs=proper_idxs, proper_params=proper_params, improper_idxs=improper_idxs, improper_params=improper_params,
        params=params, exclusion_idxs=exclusion_idxs, scale_factors=scale_factors, masses=masses, n_protein=len(prot),
    )
