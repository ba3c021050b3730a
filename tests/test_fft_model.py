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
