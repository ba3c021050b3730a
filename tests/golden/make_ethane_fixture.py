"""Generates tests/golden/ethane_water.npz from the reference's example input
(/root/reference/examples/ethane_water_charge_only/solvated_ethane_from_openmm_setup.pdb: ethane + 1383 waters,
CRYST1 35 A).  Stores integer milli-Angstrom coordinates exactly as printed, the element of every atom and the
ethane bonds from the CONECT records (water bonds follow from the O H1 H2 residue layout).  Run in the build
container only (the reference tree does not exist on the GPU box)."""
import numpy as np

src = "/root/reference/examples/ethane_water_charge_only/solvated_ethane_from_openmm_setup.pdb"
xyz, elem, resname, serial = [], [], [], []
bonds = set()
box = None
for line in open(src):
    if line.startswith("CRYST1"):
        box = [float(line[6:15]), float(line[15:24]), float(line[24:33])]
    elif line.startswith(("HETATM", "ATOM")):
        serial.append(int(line[6:11]))
        resname.append(line[17:20].strip())
        elem.append(line[76:78].strip())
        xyz.append([int(round(float(line[30:38])*1000)), int(round(float(line[38:46])*1000)), int(round(float(line[46:54])*1000))])
    elif line.startswith("CONECT"):
        f = [int(line[k:k+5]) for k in range(6, len(line.rstrip()), 5)]
        for b in f[1:]:
            bonds.add((min(f[0], b), max(f[0], b)))
index = {s: i for i, s in enumerate(serial)}
xyz = np.array(xyz, dtype=np.int32)
n = len(xyz)
assert n == 8 + 3*1383, n
assert resname[:8] == ["UNK"]*8 and all(r == "HOH" for r in resname[8:])
assert elem[8:11] == ["O", "H", "H"]
eth = np.array(sorted((index[a], index[b]) for a, b in bonds if index[a] < 8 and index[b] < 8), dtype=np.int32)
assert len(eth) == 7
code = np.array([{"C": 6, "H": 1, "O": 8}[e] for e in elem], dtype=np.int8)
np.savez_compressed("/root/repo/tests/golden/ethane_water.npz", milli_angstrom=xyz, box_angstrom=np.array(box),
                    atomic_number=code, ethane_bonds=eth)
print(n, box, eth.tolist())
