"""CPU model of the index algebra of the fused reciprocal pass (mpidopenmmplugin_b200/csrc/mpid_fft.cuh, "fused2"): the
two-pass Cooley-Tukey split every 1-D transform uses, the in-register decimation-in-time DFT, and the half-length
real-to-complex / complex-to-real tricks.  The CUDA kernels themselves are checked against cuFFT on the GPU
(tests/test_gpu_parity.py::test_fused_reciprocal_pass_matches_cufft); this keeps the derivation executable."""
import numpy as np
import pytest

RNG = np.random.default_rng(5)


def dft_reg(v, inverse):
    """dftReg<R, INV>: radix-2 decimation in time, natural order in and out."""
    n = len(v)
    if n == 2:
        return np.array([v[0] + v[1], v[0] - v[1]])
    e, o = dft_reg(v[0::2], inverse), dft_reg(v[1::2], inverse)
    w = np.exp((2j if inverse else -2j)*np.pi*np.arange(n//2)/n)
    return np.concatenate([e + w*o, e - w*o])


def two_pass(x, r1, r2, inverse):
    """fft2Pass1 + fft2Pass2: n = r2 r + j, k = q + r1 p; pass 1 in place (slots r2 q + j), pass 2 to natural order."""
    length = r1*r2
    buf = x.astype(complex).copy()
    tw = np.exp(-2j*np.pi*np.arange(length)/length)
    for j in range(r2):
        v = dft_reg(buf[j::r2].copy(), inverse)
        w = tw[j*np.arange(r1)]
        buf[j::r2] = v*(np.conj(w) if inverse else w)
    out = np.empty(length, dtype=complex)
    for q in range(r1):
        out[q::r1] = dft_reg(buf[r2*q:r2*q + r2].copy(), inverse)
    return out


@pytest.mark.parametrize("r1,r2", [(4, 4), (8, 4), (8, 8), (16, 8), (16, 16)])      # every split Fft2Plan uses
def test_two_pass_transform_is_the_dft(r1, r2):
    x = RNG.normal(size=r1*r2) + 1j*RNG.normal(size=r1*r2)
    assert np.allclose(two_pass(x, r1, r2, False), np.fft.fft(x), atol=1e-11)
    assert np.allclose(two_pass(x, r1, r2, True), np.fft.ifft(x)*r1*r2, atol=1e-11)


@pytest.mark.parametrize("nz", [32, 64, 128])
def test_half_length_real_transforms(nz):
    """k_fft2_planes_forward / _backward: a length-nz real row as a length-nz/2 complex transform plus an untangling pass."""
    m = nz//2
    x = RNG.normal(size=nz)
    z = np.fft.fft(x[0::2] + 1j*x[1::2])
    spec = np.empty(m + 1, dtype=complex)
    for k in range(m + 1):
        zk, zr = z[0 if k == m else k], np.conj(z[0 if k == 0 else m - k])
        e, d = 0.5*(zk + zr), 0.5*(zk - zr)
        w = -1.0 if k == m else np.exp(-2j*np.pi*k/nz)
        spec[k] = e + w*(-1j*d)
    assert np.allclose(spec, np.fft.rfft(x), atol=1e-11)
    zb = np.empty(m, dtype=complex)
    for k in range(m):
        a, b = spec[k], np.conj(spec[m - k])
        zb[k] = (a + b) + 1j*np.conj(np.exp(-2j*np.pi*k/nz))*(a - b)
    back = np.fft.ifft(zb)*m
    rec = np.empty(nz)
    rec[0::2], rec[1::2] = back.real, back.imag
    assert np.allclose(rec, np.fft.irfft(spec, n=nz)*nz, atol=1e-10)          # unnormalised, like cufftExecC2R


# ---- third generation (k_fft3_planes_*): one buffer, pass 2 in place, factors of 7 ------------------------------------
def dft7(v, inverse):
    """dft7<INV>: the (n, 7-n) symmetry form used in mpid_fft.cuh."""
    c = np.cos(2*np.pi*np.arange(1, 4)/7)
    s = np.sin(2*np.pi*np.arange(1, 4)/7)
    cc = [[c[0], c[1], c[2]], [c[1], c[2], c[0]], [c[2], c[0], c[1]]]
    ss = [[s[0], s[1], s[2]], [s[1], -s[2], -s[0]], [s[2], -s[0], s[1]]]
    a = [v[1] + v[6], v[2] + v[5], v[3] + v[4]]
    b = [v[1] - v[6], v[2] - v[5], v[3] - v[4]]
    out = np.empty(7, dtype=complex)
    out[0] = v[0] + sum(a)
    for k in range(3):
        r = v[0] + sum(cc[k][n]*a[n] for n in range(3))
        q = sum(ss[k][n]*b[n] for n in range(3))
        miq = -1j*q
        if inverse:
            out[k+1], out[6-k] = r - miq, r + miq
        else:
            out[k+1], out[6-k] = r + miq, r - miq
    return out


def dft_reg3(v, inverse):
    n = len(v)
    if n == 7:
        return dft7(v, inverse)
    if n == 14:
        e, o = dft7(v[0::2], inverse), dft7(v[1::2], inverse)
        w = np.exp((2j if inverse else -2j)*np.pi*np.arange(7)/14)
        return np.concatenate([e + w*o, e - w*o])
    return dft_reg(v, inverse)


def slot_of(k, r1, r2):
    return (k % r1)*r2 + k//r1


def elem_of(s, r1, r2):
    return (s % r2)*r1 + s//r2


def pass1(buf, r1, r2, inverse):
    length = r1*r2
    tw = np.exp(-2j*np.pi*np.arange(length)/length)
    for j in range(r2):
        v = dft_reg3(buf[j::r2].copy(), inverse)
        w = tw[j*np.arange(r1)]
        buf[j::r2] = v*(np.conj(w) if inverse else w)


def pass2_in_place(buf, r1, r2, inverse):
    for q in range(r1):
        buf[r2*q:r2*q + r2] = dft_reg3(buf[r2*q:r2*q + r2].copy(), inverse)      # element q + r1 p now at slot q r2 + p


@pytest.mark.parametrize("r1,r2", [(16, 14), (16, 7), (8, 14), (4, 7)])
def test_in_place_two_pass_transform_with_factor_seven(r1, r2):
    n = r1*r2
    x = RNG.normal(size=n) + 1j*RNG.normal(size=n)
    for inverse, ref in ((False, np.fft.fft(x)), (True, np.fft.ifft(x)*n)):
        buf = x.copy()
        pass1(buf, r1, r2, inverse)
        pass2_in_place(buf, r1, r2, inverse)
        assert np.allclose([buf[slot_of(k, r1, r2)] for k in range(n)], ref, atol=1e-10)
        assert np.allclose(buf, [ref[elem_of(s, r1, r2)] for s in range(n)], atol=1e-10)


@pytest.mark.parametrize("ny,r1y,r2y,nz,r1z,r2z", [(28, 4, 7, 56, 4, 7), (32, 8, 4, 28, 2, 7), (28, 2, 14, 32, 4, 4)])
def test_single_buffer_plane_kernels(ny, r1y, r2y, nz, r1z, r2z):
    """k_fft3_planes_forward / _backward, statement by statement: rows in place, pairwise untangle on permuted slots,
    columns addressed through slotOf, final stores through elemOf."""
    m, mc = nz//2, nz//2 + 1
    plane = RNG.normal(size=(ny, nz))
    # ---- forward
    buf = np.zeros((ny, mc), dtype=complex)
    buf[:, :m] = plane[:, 0::2] + 1j*plane[:, 1::2]
    for y in range(ny):
        row = buf[y, :m]
        pass1(row, r1z, r2z, False)
        pass2_in_place(row, r1z, r2z, False)
    tw_u = np.exp(-2j*np.pi*np.arange(m)/nz)
    for y in range(ny):
        row = buf[y]
        for k in range(m//2 + 1):
            if k == 0:
                z0 = row[0]
                row[0], row[m] = z0.real + z0.imag, z0.real - z0.imag
                continue
            k2, sk, sk2 = m - k, slot_of(k, r1z, r2z), slot_of(m - k, r1z, r2z)
            zk, zk2 = row[sk], row[sk2]
            e, d = 0.5*(zk + np.conj(zk2)), 0.5*(zk - np.conj(zk2))
            row[sk] = e + tw_u[k]*(-1j*d)
            if k2 != k:
                e, d = 0.5*(zk2 + np.conj(zk)), 0.5*(zk2 - np.conj(zk))
                row[sk2] = e + tw_u[k2]*(-1j*d)
    out = np.zeros((ny, mc), dtype=complex)
    for kz in range(mc):
        col = buf[:, m if kz == m else slot_of(kz, r1z, r2z)].copy()
        pass1(col, r1y, r2y, False)
        for q in range(r1y):
            v = dft_reg3(col[r2y*q:r2y*q + r2y].copy(), False)
            for p in range(r2y):
                out[q + r1y*p, kz] = v[p]
    assert np.allclose(out, np.fft.rfft2(plane), atol=1e-9)
    # ---- backward
    buf = out.copy()
    for kz in range(mc):
        col = buf[:, kz]
        pass1(col, r1y, r2y, True)
        pass2_in_place(col, r1y, r2y, True)
    for r in range(ny):
        row = buf[r]
        for k in range(m//2 + 1):
            k2 = m - k
            xk, xk2 = row[k], row[k2]
            b = np.conj(xk2)
            s, d = xk + b, xk - b
            wd = d if k == 0 else np.conj(tw_u[k])*d
            row[k] = s + 1j*wd
            if k != 0 and k2 != k:
                b = np.conj(xk)
                s, d = xk2 + b, xk2 - b
                row[k2] = s + 1j*np.conj(tw_u[k2])*d
    rec = np.zeros((ny, nz))
    for r in range(ny):
        row = buf[r, :m]
        pass1(row, r1z, r2z, True)
        y = elem_of(r, r1y, r2y)
        for q in range(r1z):
            v = dft_reg3(row[r2z*q:r2z*q + r2z].copy(), True)
            for p in range(r2z):
                j = q + r1z*p
                rec[y, 2*j], rec[y, 2*j + 1] = v[p].real, v[p].imag
    assert np.allclose(rec, np.fft.irfft2(out, s=(ny, nz))*ny*nz, atol=1e-8)
