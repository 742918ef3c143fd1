"""tests/golden/dhfr_5dfr.npz: the atoms of the reference's DHFR benchmark system (BASELINE.json configs[1]).

Run here (the container that has /root/reference), never on the GPU box:

    python tests/golden/make_golden_dhfr.py

The reference benchmarks "DHFR" = timemachine/testsystems/data/5dfr_solv_equil.pdb (23,558 atoms, cubic box 6.223 nm,
tests/test_benchmark.py:506-518; testsystems/dhfr.py:9-23) parameterised by OpenMM's amber99sbildn + tip3p - OpenMM is not
part of this image.  The fixture keeps what a force evaluation needs from the file: coordinates (milli-angstrom integers:
exactly what the PDB holds), the element of every atom, its residue number and whether it belongs to a water.  Bonds,
angles, torsions, exclusions and protein-like parameters are then DERIVED from the geometry by tests/dhfr_system.py,
identically for this repo and for the compiled reference, so the parity statements are about the kernels on the real
DHFR geometry (atom density, protein/water mix, atom ordering), not about a force field.
"""

from pathlib import Path

import numpy as np

PDB = Path("/root/reference/timemachine/testsystems/data/5dfr_solv_equil.pdb")
OUT = Path(__file__).resolve().parent / "dhfr_5dfr.npz"

ELEMENTS = {"H": 1, "C": 6, "N": 7, "O": 8, "S": 16, "NA": 11, "CL": 17}


def main():
    xyz, elem, resid, water = [], [], [], []
    box = None
    last_key, res_counter = None, -1
    for line in PDB.read_text().splitlines():
        if line.startswith("CRYST1"):
            box = float(line[6:15])
        if not line.startswith(("ATOM", "HETATM")):
            continue
        name = line[12:16].strip()
        resname = line[17:21].strip()
        key = (line[21:27], resname, line[72:76])
        if key != last_key:
            res_counter += 1
            last_key = key
        x, y, z = float(line[30:38]), float(line[38:46]), float(line[46:54])
        xyz.append((round(x * 1000), round(y * 1000), round(z * 1000)))
        is_water = resname in ("HOH", "TIP3", "WAT", "TIP")
        if is_water:
            e = 8 if name.startswith("O") else 1
        elif resname in ("SOD", "NA+", "Na+") or name in ("SOD", "NA"):
            e = 11
        elif resname in ("CLA", "CL-", "Cl-") or name in ("CLA", "CL"):
            e = 17
        else:
            first = name.lstrip("0123456789")[0]
            e = ELEMENTS[first]
        elem.append(e)
        resid.append(res_counter)
        water.append(is_water)
    xyz = np.array(xyz, dtype=np.int32)
    assert len(xyz) == 23558 and box is not None
    np.savez_compressed(
        OUT, xyz_milliangstrom=xyz, element=np.array(elem, dtype=np.uint8), residue=np.array(resid, dtype=np.int32),
        is_water=np.array(water, dtype=bool), box_angstrom=np.float64(box),
    )
    print(OUT, OUT.stat().st_size, "bytes;", int(np.sum(water)) // 3, "waters;", {int(e): int(np.sum(np.array(elem) == e)) for e in set(elem)})


if __name__ == "__main__":
    main()
