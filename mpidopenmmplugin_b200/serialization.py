"""XML round trip of an MPIDForce in the reference's schema.

Mirror of MPIDForceProxy (reference: serialization/src/MPIDForceProxy.cpp:70-143 serialize, :146-245 deserialize) on
top of the Python MPIDForce of api.py: same element and attribute names, same nesting, attributes in key order and
doubles with 17 significant digits like the tree the proxy fills (OpenMM's XmlSerializer itself is third party and not
in the reference tree; compat/src/openmm_serialization.cpp restates the part the proxy needs, and the golden document
tests/golden/mpidforce_serialized.xml is what the reference's proxy wrote through it).

Quirks kept on purpose: `defaultTholeWidth` is neither written nor read, so it comes back as
the constructor default; the proxy writes a `damp` attribute from an uninitialised variable and never reads it -- it is
written as 0 here and ignored on input.  Host code only: nothing here touches the device."""
import xml.etree.ElementTree as ET

from .api import MPIDForce, MPIDB200Error

COVALENT_TYPES = ["Covalent12", "Covalent13", "Covalent14", "Covalent15", "PolarizationCovalent11", "PolarizationCovalent12",
                  "PolarizationCovalent13", "PolarizationCovalent14"]          # MPIDForceProxy.cpp:44-55
_DIPOLE = ["dX", "dY", "dZ"]
_QUADRUPOLE = ["qXX", "qXY", "qYY", "qXZ", "qYZ", "qZZ"]                       # API order (MPIDForceProxy.cpp:116-122)
_OCTOPOLE = ["oXXX", "oXXY", "oXYY", "oYYY", "oXXZ", "oXYZ", "oYYZ", "oXZZ", "oYZZ", "oZZZ"]


def _num(v):
    return str(int(v)) if isinstance(v, (int,)) and not isinstance(v, bool) else "%.17g" % float(v)


def _element(name, props, children=(), depth=0):
    pad = "\t"*depth
    attrs = "".join(' %s="%s"' % (k, _num(props[k]) if not isinstance(props[k], str) else props[k]) for k in sorted(props))
    if not children:
        return "%s<%s%s/>\n" % (pad, name, attrs)
    return "%s<%s%s>\n%s%s</%s>\n" % (pad, name, attrs, "".join(children), pad, name)


def serialize(force, root_name="Force"):
    """XmlSerializer::serialize<MPIDForce>(&force, root_name, stream) as text."""
    alpha, nx, ny, nz = force.getPMEParameters()
    props = dict(version=0, type="MPIDForce", forceGroup=int(force.getForceGroup()), nonbondedMethod=int(force.getNonbondedMethod()),
                 polarizationType=int(force.getPolarizationType()), mutualInducedMaxIterations=int(force.getMutualInducedMaxIterations()),
                 cutoffDistance=float(force.getCutoffDistance()), aEwald=float(alpha),
                 mutualInducedTargetEpsilon=float(force.getMutualInducedTargetEpsilon()),
                 ewaldErrorTolerance=float(force.getEwaldErrorTolerance()), scaleFactor14=float(force.get14ScaleFactor()))
    children = [_element("MultipoleParticleGridDimension", dict(d0=int(nx), d1=int(ny), d2=int(nz)), depth=1),
                _element("ExtrapolationCoefficients", {"c%d" % i: float(c) for i, c in enumerate(force.getExtrapolationCoefficients())}, depth=1)]
    particles = []
    for i in range(force.getNumMultipoles()):
        charge, d, q, o, axis, z, x, y, thole, alphas = force.getMultipoleParameters(i)
        sub = [_element("Dipole", dict(zip(_DIPOLE, d)), depth=3), _element("Quadrupole", dict(zip(_QUADRUPOLE, q)), depth=3),
               _element("Octopole", dict(zip(_OCTOPOLE, o)), depth=3)]
        for t, name in enumerate(COVALENT_TYPES):
            sub.append(_element(name, {}, [_element("Cv", dict(v=int(a)), depth=4) for a in force.getCovalentMap(i, t)], depth=3))
        particles.append(_element("Particle", dict(axisType=int(axis), multipoleAtomZ=int(z), multipoleAtomX=int(x), multipoleAtomY=int(y),
                                                   charge=float(charge), thole=float(thole), damp=0.0, polarizabilityXX=float(alphas[0]),
                                                   polarizabilityYY=float(alphas[1]), polarizabilityZZ=float(alphas[2])), sub, depth=2))
    children.append(_element("MultipoleParticles", {}, particles, depth=1))
    return '<?xml version="1.0" ?>\n' + _element(root_name, props, children)


def _get(node, key, conv, default=None):
    if key not in node.attrib:
        if default is not None:
            return default
        raise MPIDB200Error("Unknown property '%s' in node '%s'" % (key, node.tag))
    return conv(node.attrib[key])


def _child(node, name):
    c = node.find(name)
    if c is None:
        raise MPIDB200Error("Unknown child '%s' in node '%s'" % (name, node.tag))
    return c


def deserialize(text):
    """XmlSerializer::deserialize<MPIDForce>(stream): errors follow the proxy ("Unsupported version number") and the
    serialization tree (missing property / child)."""
    root = ET.fromstring(text)
    if root.attrib.get("type") != "MPIDForce":
        raise MPIDB200Error("There is no serialization proxy registered for type %s" % root.attrib.get("type"))
    if _get(root, "version", int) != 0:
        raise MPIDB200Error("Unsupported version number")
    f = MPIDForce()
    f.setForceGroup(_get(root, "forceGroup", int, 0))
    f.setNonbondedMethod(_get(root, "nonbondedMethod", int))
    f.setPolarizationType(_get(root, "polarizationType", int))
    f.setMutualInducedMaxIterations(_get(root, "mutualInducedMaxIterations", int))
    f.setCutoffDistance(_get(root, "cutoffDistance", float))
    f.setMutualInducedTargetEpsilon(_get(root, "mutualInducedTargetEpsilon", float))
    f.setEwaldErrorTolerance(_get(root, "ewaldErrorTolerance", float))
    f.set14ScaleFactor(_get(root, "scaleFactor14", float))
    g = _child(root, "MultipoleParticleGridDimension")
    f.setPMEParameters(_get(root, "aEwald", float), _get(g, "d0", int), _get(g, "d1", int), _get(g, "d2", int))
    cnode = _child(root, "ExtrapolationCoefficients")
    coefs = []
    while "c%d" % len(coefs) in cnode.attrib:
        coefs.append(float(cnode.attrib["c%d" % len(coefs)]))
    f.setExtrapolationCoefficients(coefs)
    for i, p in enumerate(_child(root, "MultipoleParticles")):
        d = [_get(_child(p, "Dipole"), k, float) for k in _DIPOLE]
        q = [_get(_child(p, "Quadrupole"), k, float) for k in _QUADRUPOLE]
        o = [_get(_child(p, "Octopole"), k, float) for k in _OCTOPOLE]
        a = [_get(p, k, float) for k in ("polarizabilityXX", "polarizabilityYY", "polarizabilityZZ")]
        f.addMultipole(_get(p, "charge", float), d, q, o, _get(p, "axisType", int), _get(p, "multipoleAtomZ", int),
                       _get(p, "multipoleAtomX", int), _get(p, "multipoleAtomY", int), _get(p, "thole", float), a)
        for t, name in enumerate(COVALENT_TYPES):
            f.setCovalentMap(i, t, [_get(cv, "v", int) for cv in _child(p, name)])
    return f
