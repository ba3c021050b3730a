#ifndef OPENMM_COMPAT_FFTPACK_H_
#define OPENMM_COMPAT_FFTPACK_H_
// Stand-in for OpenMM's fftpack wrapper: plain unnormalised complex 3-D DFT (both directions),
// written from scratch as a recursive mixed-radix Cooley-Tukey transform (any n; prime factors
// are handled by direct O(p^2) butterflies).  Only the entry points the MPID Reference platform
// calls are provided.
#include <cmath>
#include <vector>

struct t_complex {
    double re, im;
    t_complex() : re(0.0), im(0.0) {}
    t_complex(double re, double im) : re(re), im(im) {}
};

enum fftpack_direction { FFTPACK_BACKWARD = -1, FFTPACK_FORWARD = 1 };

struct fftpack_plan_3d {
    int n[3];
};
typedef fftpack_plan_3d* fftpack_t;

namespace fftpack_compat {

inline void fft1d(int n, const t_complex* tw, int twstride, const t_complex* in, int istride, t_complex* out,
                  std::vector<t_complex>& scratch, size_t soff) {
    // out[k] = sum_j in[j*istride] * w^(jk), w = tw[twstride] = exp(-/+ 2 pi i / n)
    if (n == 1) { out[0] = in[0]; return; }
    int p = n;
    for (int f = 2; f*f <= n; f++) if (n % f == 0) { p = f; break; }
    int m = n/p;
    // caller pre-sizes scratch to >= 2*n_top (sum of n + n/p + ... < 2 n), so pointers into it stay valid
    for (int r = 0; r < p; r++)
        fft1d(m, tw, twstride*p, in + (size_t) r*istride, istride*p, &scratch[soff + (size_t) r*m], scratch, soff + n);
    for (int k = 0; k < m; k++) {
        for (int q = 0; q < p; q++) {
            int kk = k + q*m;
            double sr = 0.0, si = 0.0;
            for (int r = 0; r < p; r++) {
                const t_complex& w = tw[(size_t) ((long long) r*kk % n)*twstride];
                const t_complex& v = scratch[soff + (size_t) r*m + k];
                sr += v.re*w.re - v.im*w.im;
                si += v.re*w.im + v.im*w.re;
            }
            out[kk] = t_complex(sr, si);
        }
    }
}

inline void twiddles(int n, int sign, std::vector<t_complex>& tw) {
    tw.resize(n);
    for (int j = 0; j < n; j++) {
        double ang = -sign*2.0*M_PI*j/n;
        tw[j] = t_complex(std::cos(ang), std::sin(ang));
    }
}

} // namespace fftpack_compat

inline int fftpack_init_3d(fftpack_t* plan, int nx, int ny, int nz) {
    *plan = new fftpack_plan_3d;
    (*plan)->n[0] = nx; (*plan)->n[1] = ny; (*plan)->n[2] = nz;
    return 0;
}

inline void fftpack_destroy(fftpack_t plan) { delete plan; }

inline int fftpack_exec_3d(fftpack_t plan, fftpack_direction dir, t_complex* in, t_complex* out) {
    const int nx = plan->n[0], ny = plan->n[1], nz = plan->n[2];
    const int sign = (dir == FFTPACK_FORWARD) ? 1 : -1;
    const size_t total = (size_t) nx*ny*nz;
    if (out != in) for (size_t i = 0; i < total; i++) out[i] = in[i];
    std::vector<t_complex> line, res, scratch, tw;
    int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    line.resize(nmax); res.resize(nmax); scratch.resize(2*(size_t) nmax + 8);
    // z lines (contiguous)
    fftpack_compat::twiddles(nz, sign, tw);
    for (int x = 0; x < nx; x++) for (int y = 0; y < ny; y++) {
        t_complex* base = out + ((size_t) x*ny + y)*nz;
        for (int z = 0; z < nz; z++) line[z] = base[z];
        fftpack_compat::fft1d(nz, tw.data(), 1, line.data(), 1, res.data(), scratch, 0);
        for (int z = 0; z < nz; z++) base[z] = res[z];
    }
    // y lines
    fftpack_compat::twiddles(ny, sign, tw);
    for (int x = 0; x < nx; x++) for (int z = 0; z < nz; z++) {
        t_complex* base = out + (size_t) x*ny*nz + z;
        for (int y = 0; y < ny; y++) line[y] = base[(size_t) y*nz];
        fftpack_compat::fft1d(ny, tw.data(), 1, line.data(), 1, res.data(), scratch, 0);
        for (int y = 0; y < ny; y++) base[(size_t) y*nz] = res[y];
    }
    // x lines
    fftpack_compat::twiddles(nx, sign, tw);
    for (int y = 0; y < ny; y++) for (int z = 0; z < nz; z++) {
        t_complex* base = out + (size_t) y*nz + z;
        for (int x = 0; x < nx; x++) line[x] = base[(size_t) x*ny*nz];
        fftpack_compat::fft1d(nx, tw.data(), 1, line.data(), 1, res.data(), scratch, 0);
        for (int x = 0; x < nx; x++) base[(size_t) x*ny*nz] = res[x];
    }
    return 0;
}
#endif
