"""Force-field XML -> MPIDForce: the generator rules of python/mpidplugin.i:534-1052 restated in
mpidopenmmplugin_b200/forcefield.py.  CPU only; the reference's example inputs are read when the tree is present."""
import os
import warnings

import pytest

from mpidopenmmplugin_b200 import MPIDForce
from mpidopenmmplugin_b200 import forcefield as FF
from mpidopenmmplugin_b200.workloads import ethane_box, water_box

EXAMPLES = "/root/reference/examples"

# a made-up force field that exercises every anchor search and axis type (numbers are arbitrary)
XML = """<ForceField>
 <AtomTypes>
  <Type name="C" class="CT" element="C" mass="12"/> <Type name="O" class="OH" element="O" mass="16"/>
  <Type name="HC" class="HX" element="H" mass="1"/> <Type name="HO" class="HX" element="H" mass="1"/>
  <Type name="N" class="NT" element="N" mass="14"/> <Type name="HN" class="HN" element="H" mass="1"/>
  <Type name="X" class="X" element="Ar" mass="40"/>
 </AtomTypes>
 <Residues>
  <Residue name="MOH">
   <Atom name="C" type="C"/><Atom name="O" type="O"/><Atom name="HO" type="HO"/>
   <Atom name="H1" type="HC"/><Atom name="H2" type="HC"/><Atom name="H3" type="HC"/>
   <Bond from="0" to="1"/><Bond from="1" to="2"/><Bond from="0" to="3"/><Bond from="0" to="4"/><Bond from="0" to="5"/>
  </Residue>
  <Residue name="NH3">
   <Atom name="N" type="N"/><Atom name="H1" type="HN"/><Atom name="H2" type="HN"/><Atom name="H3" type="HN"/>
   <Bond atomName1="N" atomName2="H1"/><Bond atomName1="N" atomName2="H2"/><Bond atomName1="N" atomName2="H3"/>
  </Residue>
  <Residue name="AR"><Atom name="AR" type="X"/></Residue>
 </Residues>
 <MPIDForce coulomb14scale="0.5" defaultTholeWidth="6.0">
  <Multipole type="C" kz="O" kx="HC" c0="0.1" dZ="0.01"/>
  <Multipole type="O" kz="-C" kx="-HO" c0="-0.6" qXX="0.001" qYY="-0.0005" qZZ="-0.0005"/>
  <Multipole type="HO" kz="O" kx="C" c0="0.4"/>
  <Multipole type="HC" kz="C" c0="0.03" dZ="-0.002"/>
  <Multipole type="N" kz="-HN" kx="-HN" ky="-HN" c0="-0.9" oZZZ="0.0001"/>
  <Multipole type="HN" kz="N" kx="HN" c0="0.3"/>
  <Multipole type="X" c0="0.0"/>
  <Polarize type="O" polarizabilityXX="0.0009" polarizabilityYY="0.0008" polarizabilityZZ="0.0007" thole="0.39"/>
  <Polarize type="N" polarizabilityXX="0.001" polarizabilityYY="0.001" polarizabilityZZ="0.001" thole="0.39"/>
 </MPIDForce>
</ForceField>"""


def _topology():
    top = FF.Topology()
    top.add_residue("MOH", [(n, e, None) for n, e in (("C", "C"), ("O", "O"), ("HO", "H"), ("H1", "H"), ("H2", "H"), ("H3", "H"))])
    top.add_residue("NH3", [(n, e, None) for n, e in (("N", "N"), ("H1", "H"), ("H2", "H"), ("H3", "H"))])
    top.add_residue("AR", [("AR", "Ar", None)])
    return top


def test_axis_type_rules():
    A = MPIDForce
    cases = {("", "", ""): A.NoAxisType, ("a", "", ""): A.ZOnly, ("a", "b", ""): A.ZThenX, ("-a", "-b", ""): A.Bisector,
             ("a", "-b", ""): A.Bisector, ("a", "-b", "-c"): A.ZBisect, ("-a", "-b", "-c"): A.ThreeFold, ("a", "b", "c"): A.ZThenX}
    for (kz, kx, ky), want in cases.items():
        axis, z, x, y = FF.axis_type_from_k(kz or None, kx or None, ky or None)
        assert axis == want, (kz, kx, ky)
        assert (z, x, y) == (kz.lstrip("-"), kx.lstrip("-"), ky.lstrip("-"))


def test_anchor_searches_and_covalent_maps():
    ff = FF.ForceField(XML)
    f = ff.create_mpid_force(_topology(), nonbondedMethod=FF.NoCutoff)
    assert f.getNumMultipoles() == 11 and f.getPolarizationType() == MPIDForce.Extrapolated
    assert f.get14ScaleFactor() == 0.5 and f.getDefaultTholeWidth() == 6.0          # attributes of the <MPIDForce> tag
    frames = [f.getMultipoleParameters(i)[4:8] for i in range(11)]
    A = MPIDForce
    assert frames[0] == (A.ZThenX, 1, 3, -1)            # C: z = O, x = the lowest-index HC among its bonded partners
    assert frames[1] == (A.Bisector, 0, 2, -1)          # O: z = C, x = HO, both bonded
    assert frames[2] == (A.ZThenX, 1, 0, -1)            # HO: z = O (bonded), x = C two bonds away through O
    assert frames[3] == frames[4] == frames[5] == (A.ZOnly, 0, -1, -1)
    # N: three hydrogens; which is z, x, y follows the iteration order of a Python set in the reference as well (the three
    # anchors of a ThreeFold frame are interchangeable)
    assert frames[6][0] == A.ThreeFold and sorted(frames[6][1:]) == [7, 8, 9]
    assert frames[7] == (A.ZThenX, 6, 8, -1)            # HN by class: z = N, x = lowest other HN through N
    assert frames[8] == (A.ZThenX, 6, 7, -1) and frames[9] == (A.ZThenX, 6, 7, -1)
    assert frames[10] == (A.NoAxisType, -1, -1, -1)
    assert f.getMultipoleParameters(1)[8:] == (0.39, [0.0009, 0.0008, 0.0007])       # Polarize joined to its Multipole entry
    assert f.getMultipoleParameters(0)[8:] == (0.0, [0.0, 0.0, 0.0])
    assert sorted(f.getCovalentMap(2, A.Covalent12)) == [1] and sorted(f.getCovalentMap(2, A.Covalent13)) == [0]
    assert sorted(f.getCovalentMap(2, A.Covalent14)) == [3, 4, 5] and f.getCovalentMap(2, A.Covalent15) == []
    assert sorted(f.getCovalentMap(7, A.Covalent13)) == [8, 9] and f.getCovalentMap(10, A.Covalent12) == []


def test_create_system_arguments():
    ff = FF.ForceField(XML)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        f = ff.create_mpid_force(_topology(), nonbondedMethod=FF.LJPME, nonbondedCutoff=0.7, polarization="Mutual", coulomb14scale=0.8,
                                 defaultTholeWidth=6.0, aEwald=3.1, pmeGridDimensions=[24, 24, 30], mutualInducedMaxIterations=77,
                                 mutualInducedTargetEpsilon=1e-7, ewaldErrorTolerance=1e-4)
    assert len(w) == 1 and "Conflicting coulomb14scale" in str(w[0].message)       # 0.5 in the file, 0.8 asked: argument wins
    assert f.getNonbondedMethod() == MPIDForce.PME and f.getCutoffDistance() == 0.7 and f.getPolarizationType() == MPIDForce.Mutual
    assert f.get14ScaleFactor() == 0.8 and f.getDefaultTholeWidth() == 6.0
    assert f.getPMEParameters() == (3.1, 24, 24, 30) and f.getMutualInducedMaxIterations() == 77
    assert f.getMutualInducedTargetEpsilon() == 1e-7 and f.getEwaldErrorTolerance() == 1e-4
    with pytest.raises(ValueError, match="invalide polarization type"):
        ff.create_mpid_force(_topology(), polarization="sor")
    with pytest.raises(ValueError, match="input cutoff method not available"):
        ff.create_mpid_force(_topology(), nonbondedMethod="Ewald")


def test_errors():
    with pytest.raises(ValueError, match="polarize type not present"):
        FF.ForceField(XML.replace('<Multipole type="N" kz="-HN" kx="-HN" ky="-HN" c0="-0.9" oZZZ="0.0001"/>', ""))
    ff = FF.ForceField(XML.replace('<Multipole type="X" c0="0.0"/>', ""))
    with pytest.raises(ValueError, match="No multipole type for atom AR"):
        ff.create_mpid_force(_topology())
    ff = FF.ForceField(XML.replace('<Multipole type="HC" kz="C" c0="0.03" dZ="-0.002"/>', '<Multipole type="HC" kz="O" c0="0.03"/>'))
    with pytest.raises(ValueError, match="was not assigned"):
        ff.create_mpid_force(_topology())
    top = _topology()
    top.add_residue("UNK", [("Q", "Xe", None)])
    with pytest.raises(ValueError, match="No template found"):
        FF.ForceField(XML).create_mpid_force(top)


def test_residue_matched_by_its_bond_graph():
    """A residue whose name and atom names match no template (PDB files written by other tools: the ethane of the
    reference's example is `UNK` with atoms C1 C2 H H2 ...) is matched by elements and CONECT bonds."""
    pdb = "\n".join([
        "CRYST1   20.000   20.000   20.000  90.00  90.00  90.00 P 1           1",
        "HETATM    1  Q1  UNK     1       0.000   0.000   0.000  1.00  0.00           H",
        "HETATM    2  Q2  UNK     1       1.000   0.000   0.000  1.00  0.00           N",
        "HETATM    3  Q3  UNK     1       1.300   0.900   0.000  1.00  0.00           H",
        "HETATM    4  Q4  UNK     1       1.300  -0.900   0.000  1.00  0.00           H",
        "CONECT    2    1    3    4", "END"])
    top = FF.Topology.from_pdb(pdb)
    assert top.box == (2.0, 2.0, 2.0) and top.bonds == [(0, 1), (1, 2), (1, 3)]
    f = FF.ForceField(XML).create_mpid_force(top)
    assert [f.getMultipoleParameters(i)[0] for i in range(4)] == [0.3, -0.9, 0.3, 0.3]
    assert f.getMultipoleParameters(1)[4] == MPIDForce.ThreeFold and sorted(f.getMultipoleParameters(1)[5:8]) == [0, 2, 3]


def test_water_written_with_other_names_is_read_like_openmm_app_reads_it():
    """WAT / OW HW1 HW2 without CONECT records: openmm.app.PDBFile renames it to HOH / O H1 H2 and adds the standard O-H
    bonds, so the template matches (by name when the force field calls it HOH, by its bond graph otherwise)."""
    pdb = "\n".join([
        "CRYST1   20.000   20.000   20.000  90.00  90.00  90.00 P 1           1",
        "ATOM      1  OW  WAT A   1       0.000   0.000   0.000  1.00  0.00           O",
        "ATOM      2  HW1 WAT A   1       0.957   0.000   0.000  1.00  0.00           H",
        "ATOM      3  HW2 WAT A   1      -0.240   0.927   0.000  1.00  0.00           H",
        "ATOM      4  O   SOL A   2       5.000   0.000   0.000  1.00  0.00           O",
        "ATOM      5  H1  SOL A   2       5.957   0.000   0.000  1.00  0.00           H",
        "ATOM      6  H2  SOL A   2       4.760   0.927   0.000  1.00  0.00           H",
        "END"])
    top = FF.Topology.from_pdb(pdb)
    assert [n for n, _ in top.residues] == ["HOH", "HOH"] and top.atom_names[:3] == ["O", "H1", "H2"]
    assert top.bonds == [(0, 1), (0, 2), (3, 4), (3, 5)]
    xml = """<ForceField>
     <AtomTypes><Type name="OT" class="OT" element="O" mass="16"/><Type name="HT" class="HT" element="H" mass="1"/></AtomTypes>
     <Residues><Residue name="%s"><Atom name="O" type="OT"/><Atom name="H1" type="HT"/><Atom name="H2" type="HT"/>
       <Bond from="0" to="1"/><Bond from="0" to="2"/></Residue></Residues>
     <MPIDForce coulomb14scale="1.0" defaultTholeWidth="8.0">
      <Multipole type="OT" kz="-HT" kx="-HT" c0="-0.8"/><Multipole type="HT" kz="OT" kx="HT" c0="0.4"/>
     </MPIDForce></ForceField>"""
    for template_name in ("HOH", "SWM"):                     # matched by name / by the bond graph
        f = FF.ForceField(xml % template_name).create_mpid_force(FF.Topology.from_pdb(pdb))
        assert [f.getMultipoleParameters(i)[0] for i in range(6)] == [-0.8, 0.4, 0.4]*2
        assert f.getMultipoleParameters(0)[4] == MPIDForce.Bisector


def test_ambiguous_templates_are_refused():
    """Two templates with the same bond graph but different atom types: nothing decides between them."""
    twin = XML.replace('<Residue name="AR"><Atom name="AR" type="X"/></Residue>',
                       '<Residue name="AR"><Atom name="AR" type="X"/></Residue>'
                       '<Residue name="ND3"><Atom name="N" type="N"/><Atom name="D1" type="HO"/><Atom name="D2" type="HN"/><Atom name="D3" type="HN"/>'
                       '<Bond from="0" to="1"/><Bond from="0" to="2"/><Bond from="0" to="3"/></Residue>')
    top = FF.Topology()
    idx = top.add_residue("UNK", [("Q2", "N", None), ("Q1", "H", None), ("Q3", "H", None), ("Q4", "H", None)])
    for h in idx[1:]:
        top.add_bond(idx[0], h)
    with pytest.raises(ValueError, match="Multiple matching templates"):
        FF.ForceField(twin).create_mpid_force(top)
    FF.ForceField(XML).create_mpid_force(top)          # with one candidate the same residue is fine


def _same_force(f, g):
    assert f.getNumMultipoles() == g.getNumMultipoles()
    for i in range(f.getNumMultipoles()):
        assert f.getMultipoleParameters(i) == g.getMultipoleParameters(i), i
        for t in range(8):
            assert sorted(f.getCovalentMap(i, t)) == sorted(g.getCovalentMap(i, t)), (i, t)


@pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="reference examples not present")
def test_reference_examples_reproduce_the_benchmark_workloads():
    """examples/waterbox (swm6.xml + 996 waters, run.py's arguments) and examples/ethane_water_charge_only give exactly the
    forces that workloads.water_box / ethane_box build by hand -- the ones pinned against the oracle by the GPU parity tests."""
    ff = FF.ForceField(EXAMPLES + "/parameters/swm6.xml")
    top = FF.Topology.from_pdb(open(EXAMPLES + "/waterbox/waterbox_31ang.pdb").read())
    f = ff.create_mpid_force(top, nonbondedMethod=FF.LJPME, nonbondedCutoff=0.8, defaultTholeWidth=8)
    assert abs(top.box[0] - 3.1289) < 1e-12 and f.getPolarizationType() == MPIDForce.Extrapolated and f.getDefaultTholeWidth() == 8.0
    _same_force(f, water_box((1, 1, 1), polarization=2).to_force())
    ff = FF.ForceField(EXAMPLES + "/ethane_water_charge_only/ethane_water.xml")
    top = FF.Topology.from_pdb(open(EXAMPLES + "/ethane_water_charge_only/solvated_ethane_from_openmm_setup.pdb").read())
    f = ff.create_mpid_force(top, nonbondedMethod=FF.PME, nonbondedCutoff=0.8, polarization="direct", defaultTholeWidth=8)
    assert f.get14ScaleFactor() == 1.0 and f.getPolarizationType() == MPIDForce.Direct
    _same_force(f, ethane_box().to_force())


def test_k_attributes_are_collected_positionally_like_the_reference():
    """mpidplugin.i:632-640 appends the non-empty kz/kx/ky values in that order, so a missing kz makes kx the z anchor;
    and it indexes attrib['type'], so a <Multipole> without a type attribute is an error."""
    import io
    xml = """<ForceField>
 <AtomTypes><Type name="A" class="A" element="O" mass="16"/><Type name="B" class="B" element="H" mass="1"/></AtomTypes>
 <Residues><Residue name="XX"><Atom name="A1" type="A"/><Atom name="B1" type="B"/><Bond from="0" to="1"/></Residue></Residues>
 <MPIDForce>
  <Multipole type="A" kx="B" c0="-0.5"/>
  <Multipole type="B" kz="" kx="A" ky="" c0="0.5"/>
 </MPIDForce>
</ForceField>"""
    ff = FF.ForceField(xml)
    a, b = ff.entries["A"][0], ff.entries["B"][0]
    assert (a["kz"], a["kx"], a["axisType"]) == ("B", "", FF.MPIDForce.ZOnly)
    assert (b["kz"], b["kx"], b["axisType"]) == ("A", "", FF.MPIDForce.ZOnly)
    bad = xml.replace('<Multipole type="A" kx="B"', '<Multipole class="A" kx="B"')
    with pytest.raises(KeyError):
        FF.ForceField(bad)
