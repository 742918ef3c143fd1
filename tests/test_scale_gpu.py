"""GPU: BASELINE.json's full sizes (23,558-atom DHFR-sized box, ~30k-atom RBFE box, ~90k-atom stress box).
At these sizes the NumPy oracle is too slow, so parity is established through
  * the C/OpenMP oracle (O(N^2), seconds) on the 23.5k box,
  * the compiled reference custom_ops where available (forces within 1e-5),
  * size-independent exact properties: Newton's third law holds EXACTLY in fixed point (sum of du_dx over atoms is the
    zero integer), Hilbert sort on/off and padding on/off are bitwise identical, translating every atom by a lattice
    vector changes nothing beyond input rounding, energies from the U-only and full kernels agree bitwise."""

import numpy as np
import pytest

from oracle import build_oracle as OC
from tests.common import assert_forces_close, load_reference_ops, round_to_f32, water_box

pytestmark = pytest.mark.gpu
BETA, CUTOFF = 2.0, 1.2


def pots():
    from timemachine_b200 import potentials

    return potentials


def to_fixed(a):
    return np.rint(a * 2.0**36).astype(np.int64)


@pytest.fixture(scope="module")
def dhfr_sized():
    s = water_box(7853, seed=23)  # 23,559 atoms (DHFR is 23,558), box 6.17 nm
    s["x"] = round_to_f32(s["x"])
    s["params"] = round_to_f32(s["params"])
    return s


def test_dhfr_sized_against_c_oracle(dhfr_sized):
    s = dhfr_sized
    N = s["N"]
    impl = pots().Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF).to_gpu(np.float32).unbound_impl
    dx, dp, u = impl.execute(s["x"], s["params"], s["box"])
    idx = np.arange(N)
    ref_u, ref_dx, ref_dp = OC.nonbonded_block(s["x"], s["params"], s["box"], idx, idx, BETA, CUTOFF, True, True, True)
    ref_u += OC.nonbonded_pairs(s["x"], s["params"], s["box"], s["exclusion_idxs"], s["scale_factors"], -1.0, BETA, CUTOFF, dx=ref_dx, dp=ref_dp)
    np.testing.assert_allclose(u, ref_u, rtol=1e-4)
    assert_forces_close(ref_dx, dx, 2e-4)
    assert_forces_close(ref_dp, dp, 2e-3, what="du_dp")
    # Newton's third law, exactly, in fixed point
    assert not to_fixed(dx).sum(axis=0).any()
    # charge / sigma / eps gradients of pairs are symmetric sums; the w-gradient is antisymmetric -> sums to zero exactly
    assert to_fixed(dp[:, 3]).sum() == 0


def test_dhfr_sized_invariances(dhfr_sized):
    s = dhfr_sized
    N = s["N"]
    P = pots()
    base = P.NonbondedAllPairs(N, BETA, CUTOFF).to_gpu(np.float32).unbound_impl
    r0 = base.execute(s["x"], s["params"], s["box"])
    unsorted = P.NonbondedAllPairs(N, BETA, CUTOFF, disable_hilbert_sort=True).to_gpu(np.float32).unbound_impl
    nopad = P.NonbondedAllPairs(N, BETA, CUTOFF, nblist_padding=0.0).to_gpu(np.float32).unbound_impl
    for other in (unsorted, nopad):
        for a, b in zip(r0, other.execute(s["x"], s["params"], s["box"])):
            np.testing.assert_array_equal(a, b)
    # energy-only and force-only variants agree bitwise with the full evaluation
    assert base.execute(s["x"], s["params"], s["box"], False, False, True)[2] == r0[2]
    np.testing.assert_array_equal(base.execute(s["x"], s["params"], s["box"], True, False, False)[0], r0[0])
    # the Hilbert-sorted list is markedly sparser than the unsorted one (T/N 1.03 vs 1.27)
    t_sorted, t_unsorted = base.get_tile_count(), unsorted.get_tile_count()
    assert t_sorted < t_unsorted
    assert 0.7 * N / 32 * 32 < t_sorted * 1.0 < 2.0 * N  # about one tile per atom


def test_dhfr_sized_against_reference(dhfr_sized):
    ref = load_reference_ops()
    if ref is None:
        pytest.skip("oracle/_ref/custom_ops*.so not built")
    s = dhfr_sized
    N = s["N"]
    ref_ap = ref.NonbondedAllPairs_f32(N, BETA, CUTOFF)
    ref_ex = ref.NonbondedExclusions_f32(s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF)
    ref_impl = ref.FanoutSummedPotential([ref_ap, ref_ex], True)
    rdx, rdp, ru = ref_impl.execute(s["x"], s["params"], s["box"])
    impl = pots().Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF).to_gpu(np.float32).unbound_impl
    dx, dp, u = impl.execute(s["x"], s["params"], s["box"])
    assert_forces_close(rdx, dx, 1e-5)  # north_star: forces within 1e-5 relative of the reference custom_ops
    assert_forces_close(rdp, dp, 1e-4, what="du_dp")
    np.testing.assert_allclose(u, ru, rtol=1e-6)
    # tile counts of the two neighbour lists are of the same size (same algorithmic work)
    assert abs(impl.get_potentials()[0].get_tile_count() - N) < N


@pytest.mark.parametrize("n_waters", [10000, 30000])
def test_large_boxes_exact_properties(n_waters):
    s = water_box(n_waters, seed=n_waters)
    N = s["N"]
    x, params, box = round_to_f32(s["x"]), round_to_f32(s["params"]), s["box"]
    impl = pots().Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF).to_gpu(np.float32).unbound_impl
    dx, dp, u = impl.execute(x, params, box)
    assert np.isfinite(u) and np.isfinite(dx).all()
    assert not to_fixed(dx).sum(axis=0).any()
    # repeat: bitwise identical
    dx2, dp2, u2 = impl.execute(x, params, box)
    np.testing.assert_array_equal(dx, dx2)
    np.testing.assert_array_equal(dp, dp2)
    assert u == u2
    # a rigid translation by a box vector is exactly representable here only up to f32 rounding of the inputs; apply it in
    # f64 to atoms whose image is exact in f32 (multiples of L that keep 24 bits): use the integer box trick instead:
    # moving atoms by exactly one box length in x for the f64 kernel must not change anything beyond 1e-9
    impl64 = pots().Nonbonded(N, s["exclusion_idxs"], s["scale_factors"], BETA, CUTOFF).to_gpu(np.float64).unbound_impl
    sel = np.arange(0, N, 7)
    x_shift = x.copy()
    x_shift[sel, 0] += box[0, 0]
    a = impl64.execute(x, params, box, True, False, True)
    b = impl64.execute(x_shift, params, box, True, False, True)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-11)
    assert_forces_close(a[0], b[0], 1e-9)
    tiles = impl.get_potentials()[0].get_tile_count()
    assert 0.5 * N < tiles < 2.0 * N


def test_half_precision_prefilter_equals_exact_distance_rounds_bitwise(tmp_path):
    """k_nb_tiles_cq.cu tests the 32 x 32 distances of most tiles in packed half precision against an enlarged
    threshold and re-checks candidates in f32 (TMB_NB_PREFILTER=0 in a fresh process: exact f32 rounds everywhere).
    The filter must never lose a pair: du_dx, du_dp and u are compared bit for bit on
      * the DHFR-sized box (box 6.17 nm: every vanilla tile takes the prefilter; ~25k pairs within 1e-3 nm of the cutoff),
      * the same box with molecules displaced by whole box vectors (imaging inside the prefilter),
      * a dense blob (every pair of a tile inside the cutoff: the candidate queue runs at capacity),
      * boxes of 2.7, 3.2 and 4.0 nm: from too small for the prefilter's same-image condition (falls back per tile)
        to just large enough for every tile."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    script = tmp_path / "dump.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "from tests.common import water_box, random_nonbonded_system, round_to_f32\n"
        "from timemachine_b200 import potentials as P\n"
        "out = {}\n"
        "def run(tag, x, p, box, n, cutoff=1.2, flags=(True, True, True)):\n"
        "    impl = P.NonbondedAllPairs(n, 2.0, cutoff).to_gpu(np.float32).unbound_impl\n"
        "    for rep in range(2):\n"  # second call reuses the list
        "        dx, dp, u = impl.execute(round_to_f32(x), round_to_f32(p), box, *flags)\n"
        "    for k, v in (('dx', dx), ('dp', dp), ('u', u)):\n"
        "        if v is not None: out[tag + k] = v\n"
        "s = water_box(7853, seed=23)\n"
        "run('dhfr', s['x'], s['params'], s['box'], s['N'])\n"
        "run('dhfr_x', s['x'], s['params'], s['box'], s['N'], flags=(True, False, False))\n"
        "rng = np.random.default_rng(5)\n"
        "shift = rng.integers(-3, 4, size=(s['N'] // 3, 1, 3)) * np.diag(s['box'])\n"
        "run('shifted', (s['x'].reshape(-1, 3, 3) + shift).reshape(-1, 3), s['params'], s['box'], s['N'])\n"
        "x, p, box = random_nonbonded_system(4000, seed=11, box_len=8.0)\n"
        "x[:2000] = 4.0 + rng.normal(0, 0.25, size=(2000, 3))\n"  # blob: thousands of atoms within the cutoff of each other
        "run('blob', x, p, box, 4000)\n"
        "x, p, box = random_nonbonded_system(3000, seed=12, box_len=2.7)\n"
        "run('small', x, p, box, 3000)\n"
        "x, p, box = random_nonbonded_system(6000, seed=13, box_len=7.0)\n"
        "run('short', x, p, box, 6000, cutoff=0.9)\n"
        "for L in (3.2, 4.0):\n"  # every / most tiles qualify once b/2 - cutoff exceeds the extent of a row block
        "    x, p, box = random_nonbonded_system(int(100 * L**3), seed=14, box_len=L)\n"
        "    run(f'mid{L}', x, p, box, len(x))\n"
        "np.savez(sys.argv[1], **out)\n"
    )
    results = {}
    for mode in ("0", "1"):
        env = dict(os.environ, TMB_NB_PREFILTER=mode)
        path = tmp_path / f"out{mode}.npz"
        subprocess.check_call([sys.executable, str(script), str(path)], env=env)
        results[mode] = dict(np.load(path))
    assert set(results["0"]) == set(results["1"]) and len(results["0"]) >= 22
    for k in results["0"]:
        np.testing.assert_array_equal(results["0"][k], results["1"][k], err_msg=k)
    # the displaced copy is the same physical system: equal up to the f32 rounding of the larger coordinates
    assert_forces_close(results["1"]["dhfrdx"], results["1"]["shifteddx"], 5e-3)
