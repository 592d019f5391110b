"""Convert the reference's test and bench data files into one compressed numpy archive.

Run once in the build container (the reference checkout is not present on the GPU box):

    python tests/golden/make_fixtures.py /root/reference

Inputs (data, not code): the NIST Lennard-Jones and SPC/E configurations and the LAMMPS force dumps under
tests/data/nist-lj and tests/data/nist-spce, the criterion bench configurations benches/data/*.pdb, and
the MD test configurations tests/data/md-{helium,nacl,water}.  Output: tests/golden/lumol_fixtures.npz with,
per configuration <name>: <name>/names (unicode array), <name>/positions (n, 3), <name>/cell (3,),
<name>/bonds (nb, 2) where the file has CONECT records, and <name>/forces (n, 3) for the LAMMPS dumps.
The expected values the reference's tests assert on these files are kept next to the tests that use them
(tests/test_nist.py), each with its file:line.
"""

import os
import sys

import numpy as np


def read_xyz(path):
    with open(path) as fd:
        lines = fd.read().splitlines()
    natoms = int(lines[0])
    cell = None
    comment = lines[1]
    if "cell:" in comment:
        values = [float(v) for v in comment.split("cell:")[1].split()]
        cell = values * 3 if len(values) == 1 else values[:3]
    names, positions = [], []
    for line in lines[2 : 2 + natoms]:
        fields = line.split()
        names.append(fields[0])
        positions.append([float(v) for v in fields[1:4]])
    return names, np.array(positions), cell


def read_forces(path):
    with open(path) as fd:
        lines = fd.read().splitlines()
    natoms = int(lines[0])
    forces = np.zeros((natoms, 3))
    for i, line in enumerate(lines[2 : 2 + natoms]):
        fields = line.split()
        assert int(fields[0]) == i + 1
        forces[i] = [float(v) for v in fields[1:4]]
    return forces


def read_pdb(path):
    names, positions, bonds, cell = [], [], set(), None
    with open(path) as fd:
        for line in fd:
            record = line[:6].strip()
            if record == "CRYST1":
                cell = [float(line[6:15]), float(line[15:24]), float(line[24:33])]
            elif record in ("ATOM", "HETATM"):
                names.append(line[12:16].strip())
                positions.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            elif record == "CONECT":
                fields = [int(line[k : k + 5]) for k in range(6, len(line.rstrip()), 5) if line[k : k + 5].strip()]
                for other in fields[1:]:
                    i, j = fields[0] - 1, other - 1
                    bonds.add((min(i, j), max(i, j)))
    return names, np.array(positions), cell, np.array(sorted(bonds), dtype=np.int64).reshape(-1, 2)


def main(reference):
    out = {}

    def put(name, names, positions, cell, bonds=None):
        out[name + "/names"] = np.array(names)
        out[name + "/positions"] = positions
        out[name + "/cell"] = np.array(cell, dtype=np.float64)
        if bonds is not None:
            out[name + "/bonds"] = bonds

    data = os.path.join(reference, "tests", "data")
    lj_cells = {1: 10.0, 2: 8.0, 3: 10.0, 4: 8.0}  # tests/data/nist-lj/lj-N.toml: cell = ...
    for k in range(1, 5):
        names, positions, cell = read_xyz(os.path.join(data, "nist-lj", f"lj-{k}.xyz"))
        with open(os.path.join(data, "nist-lj", f"lj-{k}.toml")) as fd:
            toml_cell = [float(line.split("=")[1]) for line in fd if line.strip().startswith("cell")][0]
        assert toml_cell == lj_cells[k] and cell[0] == toml_cell
        put(f"nist-lj-{k}", names, positions, [toml_cell] * 3)
        names, positions, cell = read_xyz(os.path.join(data, "nist-spce", f"spce-{k}.xyz"))
        put(f"nist-spce-{k}", names, positions, cell)
        for cutoff in (9, 10):
            out[f"lammps-forces-{cutoff}-{k}/forces"] = read_forces(os.path.join(data, "nist-spce", f"forces-{cutoff}-{k}.xyz"))

    for name in ("argon", "nacl", "water", "propane"):
        names, positions, cell, bonds = read_pdb(os.path.join(reference, "benches", "data", name + ".pdb"))
        put("bench-" + name, names, positions, cell, bonds)

    names, positions, _ = read_xyz(os.path.join(data, "md-helium", "helium.xyz"))
    put("md-helium", names, positions, [10.0] * 3)  # nve-velocity-verlet.toml: cell = 10
    names, positions, _ = read_xyz(os.path.join(data, "md-nacl", "small.xyz"))
    put("md-nacl-small", names, positions, [11.2804] * 3)  # nve-ewald-small.toml: cell = 11.2804
    names, positions, cell, bonds = read_pdb(os.path.join(data, "md-water", "small.pdb"))
    put("md-water-small", names, positions, cell, bonds)

    target = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lumol_fixtures.npz")
    np.savez_compressed(target, **out)
    print(f"wrote {target}: {len(out)} arrays, {os.path.getsize(target)} bytes")

    # second archive: the flexible-molecule MD inputs tests/data/md-{butane,methane} (cell = 20 in their nve.toml).
    # The reference guesses the bonds with chemfiles (guess_bonds = true); both files list whole molecules one
    # after the other (4 united-atom carbons; C H H H H), so the bonds are written down and checked by distance.
    molecules = {}
    names, positions, _ = read_xyz(os.path.join(data, "md-butane", "butane.xyz"))
    bonds = [(4 * m + k, 4 * m + k + 1) for m in range(len(names) // 4) for k in range(3)]
    lengths = [np.linalg.norm(positions[i] - positions[j]) for (i, j) in bonds]
    assert len(names) == 200 and max(lengths) < 1.7 and min(lengths) > 1.3
    molecules["md-butane/names"], molecules["md-butane/positions"] = np.array(names), positions
    molecules["md-butane/cell"], molecules["md-butane/bonds"] = np.array([20.0] * 3), np.array(bonds, dtype=np.int64)
    names, positions, _ = read_xyz(os.path.join(data, "md-methane", "methane.xyz"))
    assert len(names) == 750 and names[:5] == ["C", "H", "H", "H", "H"]
    bonds = [(5 * m, 5 * m + k) for m in range(len(names) // 5) for k in range(1, 5)]
    lengths = [np.linalg.norm(positions[i] - positions[j]) for (i, j) in bonds]
    assert max(lengths) < 1.3 and min(lengths) > 0.9
    molecules["md-methane/names"], molecules["md-methane/positions"] = np.array(names), positions
    molecules["md-methane/cell"], molecules["md-methane/bonds"] = np.array([20.0] * 3), np.array(bonds, dtype=np.int64)
    target = os.path.join(os.path.dirname(os.path.abspath(__file__)), "md_molecules.npz")
    np.savez_compressed(target, **molecules)
    print(f"wrote {target}: {len(molecules)} arrays, {os.path.getsize(target)} bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
