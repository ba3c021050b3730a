"""Generates tests/golden/waterbox_31ang.npz from the reference's example input
(/root/reference/examples/waterbox/waterbox_31ang.pdb; 996 waters, O H1 H2 order, CRYST1 31.289 A).
Coordinates are stored as integer milli-Angstrom exactly as printed in the PDB.  Run in the build
container only (the reference tree does not exist on the GPU box)."""
import numpy as np

src = "/root/reference/examples/waterbox/waterbox_31ang.pdb"
xyz, names = [], []
box = None
for line in open(src):
    if line.startswith("CRYST1"):
        box = [float(line[6:15]), float(line[15:24]), float(line[24:33])]
    if line.startswith(("HETATM", "ATOM")):
        names.append(line[12:16].strip())
        xyz.append([int(round(float(line[30:38])*1000)), int(round(float(line[38:46])*1000)), int(round(float(line[46:54])*1000))])
xyz = np.array(xyz, dtype=np.int32)
assert len(xyz) == 2988 and names[:3] == ["O", "H1", "H2"], (len(xyz), names[:3])
np.savez_compressed("/root/repo/tests/golden/waterbox_31ang.npz", milli_angstrom=xyz, box_angstrom=np.array(box))
print(xyz.shape, box)
