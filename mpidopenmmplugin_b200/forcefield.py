"""Force-field XML -> MPIDForce: what the reference's python layer does with an `<MPIDForce>` section.

The reference registers `MPIDGenerator` with openmm.app.ForceField (python/mpidplugin.i:534-1052): `parseElement` reads
the `<Multipole>` / `<Polarize>` entries, `createForce` assigns every atom its entry, local-frame anchors and covalent
maps from the bonded topology and copies the `createSystem` keyword arguments into the force.  openmm.app is not
available here, so this module carries the small part of it the generator leans on -- atom types, residue
templates, a PDB reader with CONECT records, template matching by names or by element-labelled graph -- and restates
the generator's rules on top of the Python `MPIDForce` of api.py.  Units are the file's (nm, e, e.nm^k, nm^3); the
reference applies no conversion either (`conversion = 1.0`, :640).  Host code only.

    ff = ForceField("swm6.xml")
    top = Topology.from_pdb(open("waterbox_31ang.pdb").read())
    force = ff.create_mpid_force(top, nonbondedMethod=LJPME, nonbondedCutoff=0.8, defaultTholeWidth=8)
"""
import warnings
import xml.etree.ElementTree as ET

from .api import MPIDForce

NoCutoff, PME, LJPME = "NoCutoff", "PME", "LJPME"
_METHODS = {NoCutoff: MPIDForce.NoCutoff, PME: MPIDForce.PME, LJPME: MPIDForce.PME}      # mpidplugin.i:732-734

_DIPOLE = ("dX", "dY", "dZ")
_QUADRUPOLE = ("qXX", "qXY", "qYY", "qXZ", "qYZ", "qZZ")
_OCTOPOLE = ("oXXX", "oXXY", "oXYY", "oYYY", "oXXZ", "oXYZ", "oYYZ", "oXZZ", "oYZZ", "oZZZ")


# ---------------------------------------------------------------------------------------------------------------------
# topology
# ---------------------------------------------------------------------------------------------------------------------
class Topology:
    """Atoms (name, element, residue), residues and bonds; positions in nm and box lengths when read from a PDB file."""

    def __init__(self):
        self.atom_names, self.elements, self.residue_of = [], [], []
        self.residues = []            # (name, [atom indices])
        self.bonds = []               # (i, j), i < j, in the order they were added
        self.positions = []
        self.box = None

    def add_residue(self, name, atoms):
        """atoms: [(name, element, (x, y, z) or None)]; returns the indices of the new atoms."""
        first = len(self.atom_names)
        idx = list(range(first, first + len(atoms)))
        for nm, el, pos in atoms:
            self.atom_names.append(nm); self.elements.append(el); self.residue_of.append(len(self.residues))
            self.positions.append(pos)
        self.residues.append((name, idx))
        return idx

    def add_bond(self, i, j):
        b = (min(i, j), max(i, j))
        if i != j and b not in self.bonds:
            self.bonds.append(b)

    @property
    def num_atoms(self):
        return len(self.atom_names)

    @staticmethod
    def from_pdb(text):
        """ATOM/HETATM, TER, CRYST1 (orthorhombic lengths) and CONECT records; coordinates Angstrom -> nm."""
        top = Topology()
        serial_to_index, current, key = {}, None, None
        conect = []
        for line in text.splitlines():
            rec = line[:6]
            if rec == "CRYST1":
                top.box = tuple(float(line[6 + 9*k:15 + 9*k])*0.1 for k in range(3))
            elif rec in ("ATOM  ", "HETATM"):
                k = (line[21], line[22:27], line[17:20])
                if k != key or current is None:
                    current = []
                    top.residues.append((line[17:20].strip(), current))
                    key = k
                el = line[76:78].strip() if len(line) >= 78 and line[76:78].strip() else line[12:16].strip()[0]
                serial_to_index[int(line[6:11])] = top.num_atoms
                current.append(top.num_atoms)
                top.atom_names.append(line[12:16].strip()); top.elements.append(el.capitalize())
                top.residue_of.append(len(top.residues) - 1)
                top.positions.append((float(line[30:38])*0.1, float(line[38:46])*0.1, float(line[46:54])*0.1))
            elif rec.startswith("TER"):
                key = None
            elif rec == "CONECT":
                f = [int(line[c:c + 5]) for c in range(6, len(line.rstrip()), 5) if line[c:c + 5].strip()]
                conect += [(f[0], b) for b in f[1:]]
            elif rec.startswith("ENDMDL"):
                break
        for a, b in conect:
            if a in serial_to_index and b in serial_to_index:
                top.add_bond(serial_to_index[a], serial_to_index[b])
        top.apply_water_conventions()
        return top

    # openmm.app.PDBFile renames residues / atoms through its table of alternative names and adds the standard bonds of
    # the residues it knows, so a water written as WAT / SOL / TIP3 with OW, HW1, HW2 and no CONECT records arrives at the
    # force field as HOH with O-H1 and O-H2 bonds.  The examples of the reference are read through that class
    # (examples/waterbox/run.py), so its behaviour for water is part of what the generator sees.
    WATER_NAMES = ("HOH", "H2O", "HH0", "OHH", "OH2", "SOL", "WAT", "TIP", "TIP2", "TIP3", "TIP4")
    WATER_ATOMS = {"O": "O", "OW": "O", "OH2": "O", "H1": "H1", "HW1": "H1", "1H": "H1", "H2": "H2", "HW2": "H2", "2H": "H2"}

    def apply_water_conventions(self):
        for r, (name, atoms) in enumerate(self.residues):
            if name not in self.WATER_NAMES:
                continue
            renamed = [self.WATER_ATOMS.get(self.atom_names[a]) for a in atoms]
            if sorted(n for n in renamed if n) != ["H1", "H2", "O"]:
                continue                                    # not a plain three-site water: left as written
            self.residues[r] = ("HOH", atoms)
            where = {}
            for a, n in zip(atoms, renamed):
                if n:
                    self.atom_names[a] = n
                    where[n] = a
            self.add_bond(where["O"], where["H1"])
            self.add_bond(where["O"], where["H2"])


# ---------------------------------------------------------------------------------------------------------------------
# force field file: atom types, residue templates, the <MPIDForce> section
# ---------------------------------------------------------------------------------------------------------------------
class _Template:
    def __init__(self, name):
        self.name, self.atoms, self.bonds = name, [], []     # atoms: (name, type)


def axis_type_from_k(kz, kx, ky):
    """Axis type from the kz/kx/ky attributes (a leading '-' marks a bisector-style anchor); returns the axis type and
    the three type names without their signs.  reference: MPIDGenerator.setAxisType, mpidplugin.i:553-601."""
    def split(k):
        k = k or ""
        return (k[1:], True) if k.startswith("-") else (k, False)
    (kz, zneg), (kx, xneg), (ky, yneg) = split(kz), split(kx), split(ky)
    axis = MPIDForce.ZThenX
    if not kz:
        axis = MPIDForce.NoAxisType
    if kz and not kx:
        axis = MPIDForce.ZOnly
    if (kz and zneg) or (kx and xneg):
        axis = MPIDForce.Bisector
    if kx and xneg and ky and yneg:
        axis = MPIDForce.ZBisect
    if kz and zneg and kx and xneg and ky and yneg:
        axis = MPIDForce.ThreeFold
    return axis, kz, kx, ky


class ForceField:
    """The parts of openmm.app.ForceField the MPID generator needs, plus the generator itself."""

    def __init__(self, *sources):
        self.atom_types = {}          # name -> (class, element)
        self.templates = {}
        self.entries = {}             # atom type -> [multipole entry dict], file order (typeMap, mpidplugin.i:545)
        self.coulomb14scale = None    # attributes of the <MPIDForce> tag, as strings
        self.default_thole_width = None
        self._have_section = False
        for src in sources:
            self.load(src)

    def load(self, src):
        text = src if src.lstrip().startswith("<") else open(src).read()
        root = ET.fromstring(text)
        for t in root.findall("./AtomTypes/Type"):
            self.atom_types[t.attrib["name"]] = (t.attrib.get("class", t.attrib["name"]), t.attrib.get("element", ""))
        for r in root.findall("./Residues/Residue"):
            tpl = _Template(r.attrib["name"])
            for a in r.findall("Atom"):
                tpl.atoms.append((a.attrib["name"], a.attrib["type"]))
            names = [a[0] for a in tpl.atoms]
            for b in r.findall("Bond"):
                if "from" in b.attrib:
                    tpl.bonds.append((int(b.attrib["from"]), int(b.attrib["to"])))
                else:
                    tpl.bonds.append((names.index(b.attrib["atomName1"]), names.index(b.attrib["atomName2"])))
            self.templates[tpl.name] = tpl
        for section in root.findall("./MPIDForce"):
            self._parse_section(section)

    def _types_of(self, attrib):
        """`type="name"` or `class="name"` (every type of that class); ForceField._findAtomTypes in openmm.app."""
        if "type" in attrib:
            if attrib["type"] not in self.atom_types:
                raise ValueError("MPIDGenerator: error getting type for multipole: %s" % attrib["type"])
            return [attrib["type"]]
        if "class" in attrib:
            found = [t for t, (cls, _) in self.atom_types.items() if cls == attrib["class"]]
            if not found:
                raise ValueError("MPIDGenerator: error getting type for multipole: %s" % attrib["class"])
            return found
        raise ValueError("MPIDGenerator: a Multipole/Polarize entry needs a type or class attribute")

    def _parse_section(self, element):
        """reference: MPIDGenerator.parseElement, mpidplugin.i:605-728."""
        c14, thole = element.get("coulomb14scale"), element.get("defaultTholeWidth")
        if self._have_section:
            if c14 != self.coulomb14scale:
                raise ValueError("Found multiple MPIDForce tags with different coulomb14scale arguments")
            if thole != self.default_thole_width:
                raise ValueError("Found multiple MPIDForce tags with different defaultTholeWidth arguments")
        self.coulomb14scale, self.default_thole_width, self._have_section = c14, thole, True
        for m in element.findall("Multipole"):
            # The reference reads `type` (a <Multipole> keyed by class only raises KeyError there, mpidplugin.i:632) and
            # collects the k attributes POSITIONALLY: missing or empty ones are skipped and the rest move up, so
            # kx="H" without kz is treated as kz="H" (:634-640).  Mirror both.
            if "type" not in m.attrib:
                raise KeyError("type")
            label = m.attrib["type"]
            present = [m.attrib[k] for k in ("kz", "kx", "ky") if m.attrib.get(k)]
            present += [None]*(3 - len(present))
            axis, kz, kx, ky = axis_type_from_k(present[0], present[1], present[2])
            entry = dict(label=label, kz=kz, kx=kx, ky=ky, axisType=axis, charge=float(m.attrib["c0"]),
                         dipole=[float(m.get(k, 0.0)) for k in _DIPOLE], quadrupole=[float(m.get(k, 0.0)) for k in _QUADRUPOLE],
                         octopole=[float(m.get(k, 0.0)) for k in _OCTOPOLE])
            for t in self._types_of(m.attrib):
                self.entries.setdefault(t, []).append(dict(entry))
        for p in element.findall("Polarize"):
            label = p.attrib.get("type", p.attrib.get("class"))
            alphas = [float(p.attrib["polarizabilityXX"]), float(p.attrib["polarizabilityYY"]), float(p.attrib["polarizabilityZZ"])]
            for t in self._types_of(p.attrib):
                if t not in self.entries:
                    raise ValueError("MPIDGenerator: polarize type not present: %s" % label)
                hit = False
                for e in self.entries[t]:
                    if e["label"] == label:
                        e["polarizability"], e["thole"], hit = alphas, float(p.attrib["thole"]), True
                if not hit:
                    raise ValueError("MPIDGenerator: error getting type for polarize: class index=%s not in multipole list?" % label)

    # ---- template matching -------------------------------------------------------------------------------------------
    def assign_types(self, top):
        """Atom type of every atom and the bonds the templates add (water has no CONECT records in a PDB file).  A residue is
        matched by template name + atom names when both agree, otherwise by its element-labelled bond graph."""
        types = [None]*top.num_atoms
        internal = {}
        for i, j in top.bonds:
            if top.residue_of[i] == top.residue_of[j]:
                internal.setdefault(top.residue_of[i], []).append((i, j))
        for r, (name, atoms) in enumerate(top.residues):
            tpl = self.templates.get(name)
            if tpl is not None and sorted(top.atom_names[a] for a in atoms) == sorted(n for n, _ in tpl.atoms):
                where = {top.atom_names[a]: a for a in atoms}
                order = [where[n] for n, _ in tpl.atoms]
            else:
                tpl, order = self._match_graph(top, atoms, internal.get(r, []))
                if tpl is None:
                    raise ValueError("No template found for residue %d (%s)" % (r + 1, name))
            for a, (_, t) in zip(order, tpl.atoms):
                types[a] = t
            for i, j in tpl.bonds:
                top.add_bond(order[i], order[j])
        return types

    def _match_graph(self, top, atoms, bonds):
        local = {a: k for k, a in enumerate(atoms)}
        adj = [set() for _ in atoms]
        for i, j in bonds:
            adj[local[i]].add(local[j]); adj[local[j]].add(local[i])
        elems = [top.elements[a] for a in atoms]
        found = []
        for tpl in self.templates.values():
            if len(tpl.atoms) != len(atoms) or len(tpl.bonds) != len(bonds):
                continue
            telems = [self.atom_types[t][1] for _, t in tpl.atoms]
            if sorted(telems) != sorted(elems):
                continue
            tadj = [set() for _ in tpl.atoms]
            for i, j in tpl.bonds:
                tadj[i].add(j); tadj[j].add(i)
            assign = self._isomorphism(telems, tadj, elems, adj)
            if assign is not None:
                found.append((tpl, [atoms[k] for k in assign]))
        if not found:
            return None, None
        # several templates with the same graph are fine only if they give every atom the same type (openmm.app refuses
        # to choose between templates as well)
        def typing(match):
            tpl, order = match
            return sorted(zip(order, (t for _, t in tpl.atoms)))
        if any(typing(m) != typing(found[0]) for m in found[1:]):
            raise ValueError("Multiple matching templates found for a residue: " + ", ".join(m[0].name for m in found))
        return found[0]

    @staticmethod
    def _isomorphism(telems, tadj, elems, adj):
        """assign[t] = residue-local atom matched to template atom t (same element, same bonds), or None."""
        n = len(telems)
        assign, used = [-1]*n, [False]*n

        def place(t):
            if t == n:
                return True
            for a in range(n):
                if used[a] or elems[a] != telems[t] or len(adj[a]) != len(tadj[t]):
                    continue
                if any(assign[u] >= 0 and ((assign[u] in adj[a]) != (u in tadj[t])) for u in range(t)):
                    continue
                assign[t], used[a] = a, True
                if place(t + 1):
                    return True
                assign[t], used[a] = -1, False
            return False
        return assign if place(0) else None

    # ---- the generator ---------------------------------------------------------------------------------------------------
    def create_mpid_force(self, top, nonbondedMethod=NoCutoff, nonbondedCutoff=1.0, **args):
        """reference: MPIDGenerator.createForce, mpidplugin.i:730-1050 (keyword arguments are createSystem's)."""
        if not self._have_section:
            raise ValueError("the force field has no <MPIDForce> section")
        if nonbondedMethod not in _METHODS:
            raise ValueError("MPIDForce: input cutoff method not available.")
        force = MPIDForce()
        force.setNonbondedMethod(_METHODS[nonbondedMethod])
        force.setCutoffDistance(float(nonbondedCutoff))
        if "ewaldErrorTolerance" in args:
            force.setEwaldErrorTolerance(float(args["ewaldErrorTolerance"]))
        force.setPolarizationType(MPIDForce.Extrapolated)
        if "polarization" in args:
            kinds = {"direct": MPIDForce.Direct, "mutual": MPIDForce.Mutual, "extrapolated": MPIDForce.Extrapolated}
            if str(args["polarization"]).lower() not in kinds:
                raise ValueError("MPIDForce: invalide polarization type: " + str(args["polarization"]))
            force.setPolarizationType(kinds[str(args["polarization"]).lower()])
        for key, mine, setter in (("coulomb14scale", self.coulomb14scale, force.set14ScaleFactor),
                                  ("defaultTholeWidth", self.default_thole_width, force.setDefaultTholeWidth)):
            theirs = float(args[key]) if key in args else None
            mine = float(mine) if mine else None
            if theirs is not None:
                if mine is not None and mine != theirs:
                    warnings.warn("Conflicting %s values found in forcefield file (%s) and createSystem args (%s).  "
                                  "Using the value from createSystem's arguments" % (key, mine, theirs))
                setter(theirs)
            elif mine is not None:
                setter(mine)
        if "aEwald" in args:
            force.setAEwald(float(args["aEwald"]))
        if "pmeGridDimensions" in args:
            force.setPmeGridDimensions(list(args["pmeGridDimensions"]))
        if "mutualInducedMaxIterations" in args:
            force.setMutualInducedMaxIterations(int(args["mutualInducedMaxIterations"]))
        if "mutualInducedTargetEpsilon" in args:
            force.setMutualInducedTargetEpsilon(float(args["mutualInducedTargetEpsilon"]))

        types = self.assign_types(top)
        n = top.num_atoms
        b12 = [set() for _ in range(n)]                     # AmoebaVdwGenerator.getBondedParticleSets (openmm.app)
        for i, j in top.bonds:
            b12[i].add(j); b12[j].add(i)
        b13, b14 = [], []
        for i in range(n):                                  # mpidplugin.i:799-843
            s = set()
            for j in b12[i]:
                s |= b12[j]
            b13.append(s - b12[i] - {i})
        for i in range(n):
            s = set()
            for j in b13[i]:
                s |= b12[j]
            b14.append(s - b12[i] - b13[i] - {i})

        for i in range(n):
            t = types[i]
            where = "%s %s %d" % (top.atom_names[i], top.residues[top.residue_of[i]][0], top.residue_of[i])
            if t not in self.entries:
                raise ValueError("No multipole type for atom " + where)
            found = self._choose_entry(self.entries[t], i, types, b12, b13)
            if found is None:
                raise ValueError("Atom %s was not assigned." % where)
            e, z, x, y = found
            force.addMultipole(e["charge"], e["dipole"], e["quadrupole"], e["octopole"], e["axisType"], z, x, y,
                               e.get("thole", 0.0), e.get("polarizability", [0.0, 0.0, 0.0]))
            force.setCovalentMap(i, MPIDForce.Covalent12, tuple(b12[i]))
            force.setCovalentMap(i, MPIDForce.Covalent13, tuple(b13[i]))
            force.setCovalentMap(i, MPIDForce.Covalent14, tuple(b14[i]))
        return force

    @staticmethod
    def _choose_entry(entries, i, types, b12, b13):
        """Entry and anchors (z, x, y) of atom i; four searches in the reference's order, the first hit wins
        (mpidplugin.i:845-1023): z and x among the 1-2 partners; z 1-2 and x a 1-3 partner bonded to z; z only; no frame."""
        near, second = b12[i], b13[i]
        for e in entries:                                   # (1) anchors among the bonded partners
            for z in near:
                if types[z] != e["kz"]:
                    continue
                for x in near:
                    if x == z or types[x] != e["kx"]:
                        continue
                    if not e["ky"]:
                        if types[x] == types[z] and x < z:
                            return e, x, z, -1              # same type: the lower index is the z anchor
                        lowest = min([x] + [c for c in near if types[c] == e["kx"] and c != z and c < x])
                        return e, z, lowest, -1
                    for y in near:
                        if y != z and y != x and types[y] == e["ky"]:
                            return e, z, x, y
        for e in entries:                                   # (2) x (and y) two bonds away, through z
            for z in near:
                if types[z] != e["kz"]:
                    continue
                for x in second:
                    if x == z or types[x] != e["kx"] or z not in b12[x]:
                        continue
                    if not e["ky"]:
                        lowest = min([x] + [c for c in second if types[c] == e["kx"] and c != z and z in b12[c] and c < x])
                        return e, z, lowest, -1
                    for y in second:
                        if y != z and y != x and types[y] == e["ky"] and z in b12[y]:
                            return e, z, x, y
        for e in entries:                                   # (3) a z anchor only
            for z in near:
                if not e["kx"] and e["kz"] == types[z]:
                    return e, z, -1, -1
        for e in entries:                                   # (4) no frame
            if not e["kz"]:
                return e, -1, -1, -1
        return None
