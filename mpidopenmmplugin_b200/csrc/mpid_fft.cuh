// MPIDB200 -- fused reciprocal-space pass for power-of-two PME grids (single precision).
//
// One reciprocal pass of the reference is forward 3-D FFT -> multiply by the Ewald influence function -> backward
// 3-D FFT (reference: fftpack_exec_3d + performMPIDReciprocalConvolution, MPIDReferenceForce.cpp:2931-2933,
// 3329-3366, 4066-4068).  With a library FFT that is seven launches of short kernels per pass, and the pass sits on
// the critical path of every solver iteration.  Here it is three launches, each keeping a whole 2-D slab in shared
// memory:
//   k_fft_planes_forward   one CTA per x plane : real [ny][nz] -> half-complex [ny][nz/2+1]   (z: R2C, y: C2C)
//   k_fft_x_convolve       one CTA per ky row  : x forward, * eterm, x backward, in place       (x: C2C both ways)
//   k_fft_planes_backward  one CTA per x plane : half-complex -> real                           (y: C2C, z: C2R)
// Same data layouts and the same unnormalised transforms as the cuFFT R2C/C2R path it replaces, so the rest of the
// engine cannot tell the difference.  Transforms are Stockham radix-2 passes in shared memory with twiddles from a
// table computed in double precision on the host.
#ifndef MPIDB200_FFT_CUH_
#define MPIDB200_FFT_CUH_

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <algorithm>

namespace mpid {

#define MPID_FFT_MAXLEN 512            // longest 1-D transform the table serves
#define MPID_FFT_THREADS 512

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// twiddle[t] = exp(-2 pi i t / MPID_FFT_MAXLEN) for the full circle; exp(-2 pi i k / len) = twiddle[k * MAXLEN/len].
// Every kernel copies the table into shared memory first (twS).
__device__ __forceinline__ void loadTwiddles(float2* twS, const float2* __restrict__ tw) {
    for (int t = threadIdx.x; t < MPID_FFT_MAXLEN; t += blockDim.x) twS[t] = tw[t];
}
__device__ __forceinline__ float2 twiddleOf(const float2* twS, int k, int shift, bool inverse) {
    float2 w = twS[k << shift];                 // shift = log2(MAXLEN/len)
    if (inverse) w.y = -w.y;
    return w;
}
__device__ __forceinline__ int ilog2(int v) { return 31 - __clz(v); }

// One Stockham autosort pass of radix R (2 or 4) over `count` transforms of length 2^LOGLEN in shared memory; NS is
// the product of the radices of the passes already done.  Element e of transform b sits at buf[b*strideB + e*strideE].
// BFAST: consecutive threads walk consecutive transforms (column transforms, any count); otherwise consecutive
// butterflies of one transform (row transforms, count a power of two).
template <int LOGLEN, int R, int NS, bool BFAST>
__device__ __forceinline__ void fftPass(const float2* src, float2* dst, int count, unsigned magic, int strideE, int strideB,
                                        bool inverse, const float2* twS) {
    constexpr int LEN = 1 << LOGLEN, PER = LEN/R;            // PER butterflies per transform
    constexpr int LOGPER = LOGLEN - (R == 4 ? 2 : 1);
    constexpr int TWSTEP = (MPID_FFT_MAXLEN/LEN)*(LEN/(NS*R));   // table stride of exp(-2 pi i k/(NS R))
    const int work = count*PER;
    for (int t = threadIdx.x; t < work; t += blockDim.x) {
        int bidx, j;
        if (BFAST) { j = (int) __umulhi((unsigned) t, magic); bidx = t - j*count; }
        else { j = t & (PER - 1); bidx = t >> LOGPER; }
        const int k = j & (NS - 1);
        const float2* in = src + bidx*strideB + j*strideE;
        float2* out = dst + bidx*strideB + ((j - k)*R + k)*strideE;
        if (R == 2) {
            const float2 a = in[0];
            float2 w = twS[k*TWSTEP];
            if (inverse) w.y = -w.y;
            const float2 c = cmul(in[PER*strideE], w);
            out[0] = make_float2(a.x + c.x, a.y + c.y);
            out[NS*strideE] = make_float2(a.x - c.x, a.y - c.y);
        } else {
            float2 w1 = twS[k*TWSTEP], w2 = twS[2*k*TWSTEP], w3 = twS[3*k*TWSTEP];
            if (inverse) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            const float2 v0 = in[0];
            const float2 v1 = cmul(in[PER*strideE], w1);
            const float2 v2 = cmul(in[2*PER*strideE], w2);
            const float2 v3 = cmul(in[3*PER*strideE], w3);
            const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
            // forward: y1 = d02 - i d13, y3 = d02 + i d13 ; backward: signs of i swapped
            const float2 id13 = inverse ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);     // (-+ i) d13
            out[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
            out[NS*strideE] = make_float2(d02.x + id13.x, d02.y + id13.y);
            out[2*NS*strideE] = make_float2(s02.x - s13.x, s02.y - s13.y);
            out[3*NS*strideE] = make_float2(d02.x - id13.x, d02.y - id13.y);
        }
    }
    __syncthreads();
}

// All passes of a length-2^LOGLEN transform: radix 4 while possible, one radix-2 pass when LOGLEN is odd.  Ping-pongs
// between the two buffers and returns the one that holds the result.
template <int LOGLEN, bool BFAST>
__device__ __forceinline__ float2* fftSharedT(float2* src, float2* dst, int count, unsigned magic, int strideE, int strideB,
                                              bool inverse, const float2* twS) {
#define MPID_FFT_P4(NS) { fftPass<LOGLEN, 4, NS, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS); float2* sw = src; src = dst; dst = sw; }
    if (LOGLEN >= 2) MPID_FFT_P4(1)
    if (LOGLEN >= 4) MPID_FFT_P4(4)
    if (LOGLEN >= 6) MPID_FFT_P4(16)
    if (LOGLEN >= 8) MPID_FFT_P4(64)
#undef MPID_FFT_P4
    if (LOGLEN & 1) {
        fftPass<LOGLEN, 2, (1 << (LOGLEN - 1)), BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        float2* sw = src; src = dst; dst = sw;
    }
    return src;
}

// run-time length -> compile-time instantiation (lengths 4 .. 512)
template <bool BFAST>
__device__ __forceinline__ float2* fftShared(float2* src, float2* dst, int len, int count, int strideE, int strideB,
                                             bool inverse, const float2* twS) {
    const unsigned magic = (unsigned) ((0x100000000ull + (unsigned) count - 1u)/(unsigned) count);   // t/count for t < 2^16
    switch (ilog2(len)) {
        case 2: return fftSharedT<2, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        case 3: return fftSharedT<3, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        case 4: return fftSharedT<4, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        case 5: return fftSharedT<5, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        case 6: return fftSharedT<6, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        case 7: return fftSharedT<7, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        case 8: return fftSharedT<8, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
        default: return fftSharedT<9, BFAST>(src, dst, count, magic, strideE, strideB, inverse, twS);
    }
}

// real grid plane x -> half-complex plane x.  Shared memory: 2 * ny * (nz/2 + 1) float2.
__global__ void __launch_bounds__(MPID_FFT_THREADS)
k_fft_planes_forward(int ny, int nz, const float* __restrict__ grid, float2* __restrict__ out, const float2* __restrict__ tw) {
    extern __shared__ float2 fftsm[];
    __shared__ float2 twS[MPID_FFT_MAXLEN];
    const int m = nz >> 1, mc = m + 1;
    float2* bufA = fftsm;
    float2* bufB = fftsm + ny*mc;
    loadTwiddles(twS, tw);
    const int nzShift = ilog2(MPID_FFT_MAXLEN) - ilog2(nz);
    const float2* plane = reinterpret_cast<const float2*>(grid + (size_t) blockIdx.x*ny*nz);
    // rows of nz reals read as m complex numbers z_j = x_{2j} + i x_{2j+1}
    for (int t = threadIdx.x; t < ny*m; t += blockDim.x) {
        const int y = t / m, j = t - y*m;
        bufA[y*mc + j] = plane[t];
    }
    __syncthreads();
    float2* z = fftShared<false>(bufA, bufB, m, ny, 1, mc, false, twS);
    float2* other = z == bufA ? bufB : bufA;
    // untangle: X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[m-k])/2, O = -i (Z[k] - conj Z[m-k])/2, k = 0..m  (Z[m] = Z[0])
    for (int t = threadIdx.x; t < ny*mc; t += blockDim.x) {
        const int y = t / mc, k = t - y*mc;
        const float2 zk = z[y*mc + (k == m ? 0 : k)];
        const float2 zr = cconj(z[y*mc + (k == 0 ? 0 : m - k)]);
        const float2 e = make_float2(0.5f*(zk.x + zr.x), 0.5f*(zk.y + zr.y));
        const float2 d = make_float2(0.5f*(zk.x - zr.x), 0.5f*(zk.y - zr.y));
        const float2 o = make_float2(d.y, -d.x);                                  // -i d
        float2 w = k == m ? make_float2(-1.f, 0.f) : twiddleOf(twS, k, nzShift, false);
        const float2 wo = cmul(w, o);
        other[y*mc + k] = make_float2(e.x + wo.x, e.y + wo.y);
    }
    __syncthreads();
    float2* res = fftShared<true>(other, z, ny, mc, mc, 1, false, twS);
    float2* dstp = out + (size_t) blockIdx.x*ny*mc;
    for (int t = threadIdx.x; t < ny*mc; t += blockDim.x) dstp[t] = res[t];
}

// For one ky: forward transform along x, multiply by the influence function, backward transform along x, in place.
// Shared memory: 2 * nx * (nz/2 + 1) float2.
__global__ void __launch_bounds__(MPID_FFT_THREADS)
k_fft_x_convolve(int nx, int ny, int nzc, const float* __restrict__ eterm, float2* __restrict__ data, const float2* __restrict__ tw) {
    extern __shared__ float2 fftsm[];
    __shared__ float2 twS[MPID_FFT_MAXLEN];
    float2* bufA = fftsm;
    float2* bufB = fftsm + nx*nzc;
    loadTwiddles(twS, tw);
    const int ky = blockIdx.x;
    for (int t = threadIdx.x; t < nx*nzc; t += blockDim.x) {
        const int x = t / nzc, k = t - x*nzc;
        bufA[t] = data[((size_t) x*ny + ky)*nzc + k];
    }
    __syncthreads();
    float2* f = fftShared<true>(bufA, bufB, nx, nzc, nzc, 1, false, twS);
    float2* other = f == bufA ? bufB : bufA;
    for (int t = threadIdx.x; t < nx*nzc; t += blockDim.x) {
        const int x = t / nzc, k = t - x*nzc;
        const float e = eterm[((size_t) x*ny + ky)*nzc + k];
        f[t] = make_float2(f[t].x*e, f[t].y*e);
    }
    __syncthreads();
    float2* r = fftShared<true>(f, other, nx, nzc, nzc, 1, true, twS);
    for (int t = threadIdx.x; t < nx*nzc; t += blockDim.x) {
        const int x = t / nzc, k = t - x*nzc;
        data[((size_t) x*ny + ky)*nzc + k] = r[t];
    }
}

// half-complex plane x -> real grid plane x (unnormalised, like cufftExecC2R).
__global__ void __launch_bounds__(MPID_FFT_THREADS)
k_fft_planes_backward(int ny, int nz, const float2* __restrict__ in, float* __restrict__ grid, const float2* __restrict__ tw) {
    extern __shared__ float2 fftsm[];
    __shared__ float2 twS[MPID_FFT_MAXLEN];
    const int m = nz >> 1, mc = m + 1;
    float2* bufA = fftsm;
    float2* bufB = fftsm + ny*mc;
    loadTwiddles(twS, tw);
    const int nzShift = ilog2(MPID_FFT_MAXLEN) - ilog2(nz);
    const float2* src = in + (size_t) blockIdx.x*ny*mc;
    for (int t = threadIdx.x; t < ny*mc; t += blockDim.x) bufA[t] = src[t];
    __syncthreads();
    float2* xk = fftShared<true>(bufA, bufB, ny, mc, mc, 1, true, twS);
    float2* other = xk == bufA ? bufB : bufA;
    // Z[k] = (X[k] + conj X[m-k]) + i w^-k (X[k] - conj X[m-k]), k = 0..m-1 ; then a length-m backward transform gives
    // x_{2j} + i x_{2j+1} scaled by nz, which is the unnormalised C2R result
    for (int t = threadIdx.x; t < ny*m; t += blockDim.x) {
        const int y = t / m, k = t - y*m;
        const float2 a = xk[y*mc + k];
        const float2 b = cconj(xk[y*mc + (m - k)]);
        const float2 s = make_float2(a.x + b.x, a.y + b.y);
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const float2 wd = cmul(twiddleOf(twS, k, nzShift, true), d);
        other[y*mc + k] = make_float2(s.x - wd.y, s.y + wd.x);                  // s + i wd
    }
    __syncthreads();
    float2* z = fftShared<false>(other, xk, m, ny, 1, mc, true, twS);
    float2* plane = reinterpret_cast<float2*>(grid + (size_t) blockIdx.x*ny*nz);
    for (int t = threadIdx.x; t < ny*m; t += blockDim.x) {
        const int y = t / m, j = t - y*m;
        plane[t] = z[y*mc + j];
    }
}

// =====================================================================================================
// Second generation ("fused2"): the same three kernels with register-resident radix-4/8/16 butterflies.
//
// The Stockham version above spends its time on index arithmetic and barriers: a radix-4 pass moves two butterflies
// per thread between barriers, ten barriers per plane (ncu: 28 % issue utilisation, profiles/r01_fft_experiment.md).
// Here every 1-D transform of length L = R1*R2 is exactly two passes (Cooley-Tukey, n = R2 r + j, k = q + R1 p):
//   pass 1, task (j):  A[q] = DFT_R1 over r of x[R2 r + j];  B[j][q] = A[q] w_L^(j q)   -- in place (slots R2 q + j)
//   pass 2, task (q):  X[q + R1 p] = DFT_R2 over j of B[j][q]                          -- to the next stage
// with the R-point DFTs unrolled in registers (constant twiddles), so a plane costs five barriers and the x
// direction three.  Lanes always walk the transform index, and rows are MC = nz/2+1 (odd) complex numbers apart, so
// shared-memory accesses are conflict free in both directions.
// =====================================================================================================
#define MPID_FFT2_MAX_THREADS 576

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// R-point DFT in registers (R = 2, 4, 8, 16), natural order in and out, decimation in time; INV conjugates.
// 7-point DFT from the (n, 7-n) symmetry: three sums and three differences, 18 real multiplies per output pair.
template <bool INV>
__device__ __forceinline__ void dft7(float2* v) {
    constexpr float c1 = 0.62348980185873353f, c2 = -0.22252093395631440f, c3 = -0.90096886790241913f;     // cos(2 pi k/7)
    constexpr float s1 = 0.78183148246802981f, s2 = 0.97492791218182361f, s3 = 0.43388373911755812f;       // sin(2 pi k/7)
    const float2 a1 = cadd(v[1], v[6]), a2 = cadd(v[2], v[5]), a3 = cadd(v[3], v[4]);
    const float2 b1 = csub(v[1], v[6]), b2 = csub(v[2], v[5]), b3 = csub(v[3], v[4]);
    const float2 x0 = v[0];
    v[0] = make_float2(x0.x + a1.x + a2.x + a3.x, x0.y + a1.y + a2.y + a3.y);
    // X_k = R_k -+ i I_k,  X_{7-k} = R_k +- i I_k  (upper sign: forward)
    const float cc[3][3] = {{c1, c2, c3}, {c2, c3, c1}, {c3, c1, c2}};
    const float ss[3][3] = {{s1, s2, s3}, {s2, -s3, -s1}, {s3, -s1, s2}};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float2 r = make_float2(x0.x + cc[k][0]*a1.x + cc[k][1]*a2.x + cc[k][2]*a3.x, x0.y + cc[k][0]*a1.y + cc[k][1]*a2.y + cc[k][2]*a3.y);
        const float2 q = make_float2(ss[k][0]*b1.x + ss[k][1]*b2.x + ss[k][2]*b3.x, ss[k][0]*b1.y + ss[k][1]*b2.y + ss[k][2]*b3.y);
        const float2 miq = make_float2(q.y, -q.x);          // -i q
        if (INV) { v[k+1] = csub(r, miq); v[6-k] = cadd(r, miq); }
        else     { v[k+1] = cadd(r, miq); v[6-k] = csub(r, miq); }
    }
}

template <int R, bool INV>
__device__ __forceinline__ void dftReg(float2* v) {
    if constexpr (R == 7) {
        dft7<INV>(v);
    } else if constexpr (R == 14) {
        float2 e[7], o[7];
#pragma unroll
        for (int k = 0; k < 7; k++) { e[k] = v[2*k]; o[k] = v[2*k+1]; }
        dft7<INV>(e);
        dft7<INV>(o);
        // exp(-2 pi i k/14), k = 0..6
        const float c14[7] = {1.f, 0.90096886790241913f, 0.62348980185873353f, 0.22252093395631440f,
                              -0.22252093395631440f, -0.62348980185873353f, -0.90096886790241913f};
        const float s14[7] = {0.f, -0.43388373911755812f, -0.78183148246802981f, -0.97492791218182361f,
                              -0.97492791218182361f, -0.78183148246802981f, -0.43388373911755812f};
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const float wr = c14[k], wi = INV ? -s14[k] : s14[k];
            const float2 t = k == 0 ? o[0] : make_float2(o[k].x*wr - o[k].y*wi, o[k].x*wi + o[k].y*wr);
            v[k] = cadd(e[k], t);
            v[k + 7] = csub(e[k], t);
        }
    } else if constexpr (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else {
        constexpr int H = R/2;
        float2 e[H], o[H];
#pragma unroll
        for (int k = 0; k < H; k++) { e[k] = v[2*k]; o[k] = v[2*k+1]; }
        dftReg<H, INV>(e);
        dftReg<H, INV>(o);
        // exp(-2 pi i k/16), k = 0..7
        const float c16[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                              0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
        const float s16[8] = {0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                              -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
#pragma unroll
        for (int k = 0; k < H; k++) {
            float2 t;
            if (k == 0) t = o[0];
            else if (4*k == R) t = INV ? make_float2(-o[k].y, o[k].x) : make_float2(o[k].y, -o[k].x);     // -+ i
            else {
                const float wr = c16[k*(16/R)], wi = INV ? -s16[k*(16/R)] : s16[k*(16/R)];
                t = make_float2(o[k].x*wr - o[k].y*wi, o[k].x*wi + o[k].y*wr);
            }
            v[k] = cadd(e[k], t);
            v[k + H] = csub(e[k], t);
        }
    }
}

// element e of transform b sits at buf[e*sE + b*sB]; twL[t] = exp(-2 pi i t/L), t < L
template <int R1, int R2, bool INV>
__device__ __forceinline__ void fft2Pass1(float2* buf, int count, int sE, int sB, const float2* twL) {
    const int tasks = count*R2;
    for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
        const int j = t / count, b = t - j*count;
        float2* base = buf + b*sB + j*sE;
        float2 v[R1];
#pragma unroll
        for (int r = 0; r < R1; r++) v[r] = base[r*R2*sE];
        dftReg<R1, INV>(v);
#pragma unroll
        for (int q = 1; q < R1; q++) {
            float2 w = twL[j*q];
            if (INV) w.y = -w.y;
            v[q] = cmul(v[q], w);
        }
#pragma unroll
        for (int q = 0; q < R1; q++) base[q*R2*sE] = v[q];
    }
}
template <int R1, int R2, bool INV, typename Store>
__device__ __forceinline__ void fft2Pass2(const float2* buf, int count, int sE, int sB, Store store) {
    const int tasks = count*R1;
    for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
        const int q = t / count, b = t - q*count;
        const float2* base = buf + b*sB + q*R2*sE;
        float2 v[R2];
#pragma unroll
        for (int j = 0; j < R2; j++) v[j] = base[j*sE];
        dftReg<R2, INV>(v);
#pragma unroll
        for (int p = 0; p < R2; p++) store(b, q + R1*p, v[p]);
    }
}
// The device table holds exp(-2 pi i t/512), t < 512, followed by exp(-2 pi i t/448), t < 448: the first serves the
// power-of-two lengths, the second the lengths that divide 448 = 2^6 x 7 (224, 112, 56, ...).
#define MPID_FFT_LEN7 448
#define MPID_FFT_TABLE (MPID_FFT_MAXLEN + MPID_FFT_LEN7)
__device__ __forceinline__ void fft2LoadTablePart(float2* dst, int len, int count, const float2* __restrict__ tw) {
    // dst[t] = exp(-2 pi i t/len), t < count
    if (MPID_FFT_MAXLEN % len == 0) {
        const int step = MPID_FFT_MAXLEN/len;
        for (int t = threadIdx.x; t < count; t += blockDim.x) dst[t] = tw[t*step];
    } else {
        const int step = MPID_FFT_LEN7/len;
        for (int t = threadIdx.x; t < count; t += blockDim.x) dst[t] = tw[MPID_FFT_MAXLEN + t*step];
    }
}
__device__ __forceinline__ void fft2LoadTable(float2* dst, int len, const float2* __restrict__ tw) { fft2LoadTablePart(dst, len, len, tw); }

// real grid plane x -> half-complex plane x.  Dynamic shared memory: (2 NY MC + NY + 2 M) float2, M = NZ/2, MC = M+1.
// (Clearing each plane here after it is read, so that the next spreading pass needs no memset in front of it and the
// backward transform writes to a second grid, was measured: 1.639 ms per evaluation against 1.546 ms with the memset.)
// Slab-decomposed passes (several ranks) hand in nxl > 0: the half-complex result of plane xl then goes straight into the
// send layout of the all-to-all, [destination rank q][xl][ky - q nyl][kz] with nyl = NY/ranks rows per rank, and the
// backward kernels read that layout -- no separate transpose kernels.
// Peer-to-peer all-to-all fused into the producers (several ranks on one node): dst[q] is rank q's RECEIVE buffer, mapped
// into this process; a block of nxl x nyl x mc numbers sent by rank r lands at offset slotOfSender * blk there.
struct SlabPeers {
    float2* dst[16];
    int ranks;          // 0 = not peer-to-peer (pack into the local send buffer / write the local array)
    int slot;           // block position of THIS rank's data in the receivers' buffers
    int rot;            // block position b of the x transform belongs to rank (b - rot) mod ranks (halo mode: ranks/2)
};
__device__ __forceinline__ size_t slabIndex(int nxl, int nyl, int xl, int ky, int mc) {
    const int q = ky / nyl;
    return (((size_t) q*nxl + xl)*nyl + (ky - q*nyl))*mc;
}
template <int NY, int R1Y, int R2Y, int NZ, int R1Z, int R2Z>
__global__ void __launch_bounds__(MPID_FFT2_MAX_THREADS)
k_fft2_planes_forward(const float* __restrict__ grid, float2* __restrict__ out, const float2* __restrict__ tw, int nxl, int nyl, SlabPeers peers) {
    constexpr int M = NZ/2, MC = M + 1;
    static_assert(R1Y*R2Y == NY && R1Z*R2Z == M, "radix split");
    extern __shared__ float2 fftsm[];
    float2* buf0 = fftsm;
    float2* buf1 = buf0 + NY*MC;
    float2* twY = buf1 + NY*MC;
    float2* twZ = twY + NY;
    float2* twU = twZ + M;
    fft2LoadTable(twY, NY, tw);
    fft2LoadTable(twZ, M, tw);
    for (int t = threadIdx.x; t < M; t += blockDim.x) twU[t] = tw[t*(MPID_FFT_MAXLEN/NZ)];
    // rows of NZ reals read as M complex numbers z_j = x_{2j} + i x_{2j+1}
    const float2* plane = reinterpret_cast<const float2*>(grid + (size_t) blockIdx.x*NY*NZ);
    for (int t = threadIdx.x; t < NY*M; t += blockDim.x) buf0[(t / M)*MC + (t % M)] = plane[t];
    __syncthreads();
    fft2Pass1<R1Z, R2Z, false>(buf0, NY, 1, MC, twZ);
    __syncthreads();
    fft2Pass2<R1Z, R2Z, false>(buf0, NY, 1, MC, [&](int y, int k, float2 v) { buf1[y*MC + k] = v; });
    __syncthreads();
    // untangle: X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[M-k])/2, O = -i (Z[k] - conj Z[M-k])/2, k = 0..M (Z[M] = Z[0])
    for (int t = threadIdx.x; t < NY*MC; t += blockDim.x) {
        const int k = t / NY, y = t - k*NY;
        const float2 zk = buf1[y*MC + (k == M ? 0 : k)];
        const float2 zr = cconj(buf1[y*MC + (k == 0 ? 0 : M - k)]);
        const float2 e = make_float2(0.5f*(zk.x + zr.x), 0.5f*(zk.y + zr.y));
        const float2 d = make_float2(0.5f*(zk.x - zr.x), 0.5f*(zk.y - zr.y));
        const float2 o = make_float2(d.y, -d.x);                                  // -i d
        const float2 w = k == M ? make_float2(-1.f, 0.f) : twU[k];
        buf0[y*MC + k] = cadd(e, cmul(w, o));
    }
    __syncthreads();
    fft2Pass1<R1Y, R2Y, false>(buf0, MC, MC, 1, twY);
    __syncthreads();
    if (peers.ranks > 0) {
        // the all-to-all of the slab transform, done by the producer: row ky of this plane goes to rank ky / nyl
        fft2Pass2<R1Y, R2Y, false>(buf0, MC, MC, 1, [&](int kz, int ky, float2 v) {
            const int q = ky / nyl;
            peers.dst[q][(((size_t) peers.slot*nxl + blockIdx.x)*nyl + (ky - q*nyl))*MC + kz] = v;
        });
    } else if (nxl > 0) {
        fft2Pass2<R1Y, R2Y, false>(buf0, MC, MC, 1, [&](int kz, int ky, float2 v) { out[slabIndex(nxl, nyl, blockIdx.x, ky, MC) + kz] = v; });
    } else {
        float2* dstp = out + (size_t) blockIdx.x*NY*MC;
        fft2Pass2<R1Y, R2Y, false>(buf0, MC, MC, 1, [&](int kz, int ky, float2 v) { dstp[ky*MC + kz] = v; });
    }
}

// half-complex plane x -> real grid plane x (unnormalised, like cufftExecC2R).  Same shared memory as the forward kernel.
template <int NY, int R1Y, int R2Y, int NZ, int R1Z, int R2Z>
__global__ void __launch_bounds__(MPID_FFT2_MAX_THREADS)
k_fft2_planes_backward(const float2* __restrict__ in, float* __restrict__ grid, const float2* __restrict__ tw, int nxl, int nyl) {
    constexpr int M = NZ/2, MC = M + 1;
    extern __shared__ float2 fftsm[];
    float2* buf0 = fftsm;
    float2* buf1 = buf0 + NY*MC;
    float2* twY = buf1 + NY*MC;
    float2* twZ = twY + NY;
    float2* twU = twZ + M;
    fft2LoadTable(twY, NY, tw);
    fft2LoadTable(twZ, M, tw);
    for (int t = threadIdx.x; t < M; t += blockDim.x) twU[t] = tw[t*(MPID_FFT_MAXLEN/NZ)];
    if (nxl > 0) {
        for (int t = threadIdx.x; t < NY*MC; t += blockDim.x) { const int ky = t / MC; buf0[t] = in[slabIndex(nxl, nyl, blockIdx.x, ky, MC) + (t - ky*MC)]; }
    } else {
        const float2* src = in + (size_t) blockIdx.x*NY*MC;
        for (int t = threadIdx.x; t < NY*MC; t += blockDim.x) buf0[t] = src[t];
    }
    __syncthreads();
    fft2Pass1<R1Y, R2Y, true>(buf0, MC, MC, 1, twY);
    __syncthreads();
    fft2Pass2<R1Y, R2Y, true>(buf0, MC, MC, 1, [&](int kz, int y, float2 v) { buf1[y*MC + kz] = v; });
    __syncthreads();
    // Z[k] = (X[k] + conj X[M-k]) + i w^-k (X[k] - conj X[M-k]), k = 0..M-1; a length-M backward transform then gives
    // x_{2j} + i x_{2j+1} scaled by NZ, the unnormalised C2R result
    for (int t = threadIdx.x; t < NY*M; t += blockDim.x) {
        const int k = t / NY, y = t - k*NY;
        const float2 a = buf1[y*MC + k];
        const float2 b = cconj(buf1[y*MC + (M - k)]);
        const float2 s = cadd(a, b), d = csub(a, b);
        const float2 wd = cmul(cconj(twU[k]), d);
        buf0[y*MC + k] = make_float2(s.x - wd.y, s.y + wd.x);                   // s + i wd
    }
    __syncthreads();
    fft2Pass1<R1Z, R2Z, true>(buf0, NY, 1, MC, twZ);
    __syncthreads();
    fft2Pass2<R1Z, R2Z, true>(buf0, NY, 1, MC, [&](int y, int j, float2 v) { buf1[y*MC + j] = v; });
    __syncthreads();
    float2* plane = reinterpret_cast<float2*>(grid + (size_t) blockIdx.x*NY*NZ);
    for (int t = threadIdx.x; t < NY*M; t += blockDim.x) plane[t] = buf1[(t / M)*MC + (t % M)];
}

// For one ky and a chunk of kz: forward transform along x, multiply by the influence function, backward transform
// along x, in place.  grid = (ny, ceil(nzc/chunk)).  Dynamic shared memory: (2 NX S + NX) float2, S = chunk | 1.
// (Slab-decomposed passes hand in this rank's ky rows only: data holds nyData rows per x, the influence function all
// nyEterm of them, and row 0 of the data is row ky0 of the influence function.)
template <int NX, int R1, int R2>
__global__ void __launch_bounds__(256)
k_fft2_x_convolve(int ny, int nzc, int chunk, const float* __restrict__ eterm, float2* __restrict__ data, const float2* __restrict__ tw,
                  int nyEterm, int ky0, SlabPeers peers) {
    static_assert(R1*R2 == NX, "radix split");
    extern __shared__ float2 fftsm[];
    const int S = chunk | 1;
    float2* buf0 = fftsm;
    float2* buf1 = buf0 + NX*S;
    float2* twX = buf1 + NX*S;
    fft2LoadTable(twX, NX, tw);
    const int ky = blockIdx.x, kz0 = blockIdx.y*chunk;
    const int count = min(chunk, nzc - kz0);
    const size_t line = (size_t) ky*nzc + kz0, xStride = (size_t) ny*nzc;
    const size_t lineE = (size_t) (ky0 + ky)*nzc + kz0, xStrideE = (size_t) nyEterm*nzc;
    for (int t = threadIdx.x; t < NX*count; t += blockDim.x) {
        const int x = t / count, c = t - x*count;
        buf0[x*S + c] = data[x*xStride + line + c];
    }
    __syncthreads();
    fft2Pass1<R1, R2, false>(buf0, count, S, 1, twX);
    __syncthreads();
    fft2Pass2<R1, R2, false>(buf0, count, S, 1, [&](int c, int kx, float2 v) {
        const float e = eterm[kx*xStrideE + lineE + c];
        buf1[kx*S + c] = make_float2(v.x*e, v.y*e);
    });
    __syncthreads();
    fft2Pass1<R1, R2, true>(buf1, count, S, 1, twX);
    __syncthreads();
    if (peers.ranks > 0) {
        // the all-to-all back, done by the producer: the planes x of block position b belong to rank (b - R/2) mod R, whose
        // receive buffer takes this rank's ky rows at block offset peers.slot (= this rank's number)
        const int nxl = NX/peers.ranks;
        fft2Pass2<R1, R2, true>(buf1, count, S, 1, [&](int c, int x, float2 v) {
            const int b = x / nxl, r = (b + peers.ranks - peers.rot) % peers.ranks;
            peers.dst[r][(((size_t) peers.slot*nxl + (x - b*nxl))*ny + ky)*nzc + kz0 + c] = v;
        });
    } else {
        fft2Pass2<R1, R2, true>(buf1, count, S, 1, [&](int c, int x, float2 v) { data[x*xStride + line + c] = v; });
    }
}

// =====================================================================================================
// Third generation: the plane kernels with ONE shared-memory buffer (a 224 x 224 plane is 200 KB: two do not fit), and
// lengths with a factor 7 (224 = 16 x 14, 112 = 16 x 7) through the 7- and 14-point register DFTs above.
//
// Pass 2 of a transform writes its R2 results back into the R2 slots it read (task-local, so no second buffer and no
// extra barrier): element k = q + R1 p of the result then sits at slot q R2 + p.  Whoever consumes the result maps
// element -> slot (slotOf) or slot -> element (elemOf); the real/half-complex "untangle" step works on the pair
// (k, M - k) at once, which is what makes it safe in place.
// =====================================================================================================
template <int R1, int R2> __device__ __forceinline__ int slotOf(int k) { return (k % R1)*R2 + k / R1; }
template <int R1, int R2> __device__ __forceinline__ int elemOf(int s) { return (s % R2)*R1 + s / R2; }

// transform b's element at slot e sits at buf[off(b) + e*sE]
template <int R1, int R2, bool INV, typename Off>
__device__ __forceinline__ void fft3Pass1(float2* buf, int count, int sE, Off off, const float2* twL) {
    const int tasks = count*R2;
    for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
        const int j = t / count, b = t - j*count;
        float2* base = buf + off(b) + j*sE;
        float2 v[R1];
#pragma unroll
        for (int r = 0; r < R1; r++) v[r] = base[r*R2*sE];
        dftReg<R1, INV>(v);
#pragma unroll
        for (int q = 1; q < R1; q++) {
            float2 w = twL[j*q];
            if (INV) w.y = -w.y;
            v[q] = cmul(v[q], w);
        }
#pragma unroll
        for (int q = 0; q < R1; q++) base[q*R2*sE] = v[q];
    }
}
template <int R1, int R2, bool INV, typename Off>
__device__ __forceinline__ void fft3Pass2InPlace(float2* buf, int count, int sE, Off off) {
    const int tasks = count*R1;
    for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
        const int q = t / count, b = t - q*count;
        float2* base = buf + off(b) + q*R2*sE;
        float2 v[R2];
#pragma unroll
        for (int j = 0; j < R2; j++) v[j] = base[j*sE];
        dftReg<R2, INV>(v);
#pragma unroll
        for (int p = 0; p < R2; p++) base[p*sE] = v[p];
    }
}
// QFAST: consecutive threads take consecutive q of one transform (results q + R1 p are then contiguous per p);
// otherwise consecutive transforms.
template <int R1, int R2, bool INV, bool QFAST, typename Off, typename Store>
__device__ __forceinline__ void fft3Pass2Store(const float2* buf, int count, int sE, Off off, Store store) {
    const int tasks = count*R1;
    for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
        int q, b;
        if (QFAST) { b = t / R1; q = t - b*R1; } else { q = t / count; b = t - q*count; }
        const float2* base = buf + off(b) + q*R2*sE;
        float2 v[R2];
#pragma unroll
        for (int j = 0; j < R2; j++) v[j] = base[j*sE];
        dftReg<R2, INV>(v);
#pragma unroll
        for (int p = 0; p < R2; p++) store(b, q + R1*p, v[p]);
    }
}

// real grid plane x -> half-complex plane x.  Dynamic shared memory: (NY MC + NY + 2 M) float2, M = NZ/2, MC = M + 1.
template <int NY, int R1Y, int R2Y, int NZ, int R1Z, int R2Z>
__global__ void __launch_bounds__(MPID_FFT2_MAX_THREADS)
k_fft3_planes_forward(const float* __restrict__ grid, float2* __restrict__ out, const float2* __restrict__ tw, int nxl, int nyl, SlabPeers peers) {
    constexpr int M = NZ/2, MC = M + 1;
    static_assert(R1Y*R2Y == NY && R1Z*R2Z == M, "radix split");
    extern __shared__ float2 fftsm[];
    float2* buf = fftsm;
    float2* twY = buf + NY*MC;
    float2* twZ = twY + NY;
    float2* twU = twZ + M;
    fft2LoadTable(twY, NY, tw);
    fft2LoadTable(twZ, M, tw);
    fft2LoadTablePart(twU, NZ, M, tw);
    // rows of NZ reals read as M complex numbers z_j = x_{2j} + i x_{2j+1}
    const float2* plane = reinterpret_cast<const float2*>(grid + (size_t) blockIdx.x*NY*NZ);
    for (int t = threadIdx.x; t < NY*M; t += blockDim.x) buf[(t / M)*MC + (t % M)] = plane[t];
    __syncthreads();
    auto rowOff = [](int y) { return y*MC; };
    fft3Pass1<R1Z, R2Z, false>(buf, NY, 1, rowOff, twZ);
    __syncthreads();
    fft3Pass2InPlace<R1Z, R2Z, false>(buf, NY, 1, rowOff);
    __syncthreads();
    // untangle, pairwise in place: X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[M-k])/2, O = -i (Z[k] - conj Z[M-k])/2;
    // Z[k] sits at slot slotOf(k); X[M] (from Z[0]) takes the extra column M
    for (int t = threadIdx.x; t < NY*(M/2 + 1); t += blockDim.x) {
        const int k = t / NY, y = t - k*NY;
        float2* row = buf + y*MC;
        if (k == 0) {
            const float2 z0 = row[0];
            row[0] = make_float2(z0.x + z0.y, 0.f);
            row[M] = make_float2(z0.x - z0.y, 0.f);
        } else {
            const int k2 = M - k, sk = slotOf<R1Z, R2Z>(k), sk2 = slotOf<R1Z, R2Z>(k2);
            const float2 zk = row[sk], zk2 = row[sk2];
            {
                const float2 zr = cconj(zk2);
                const float2 e = make_float2(0.5f*(zk.x + zr.x), 0.5f*(zk.y + zr.y));
                const float2 d = make_float2(0.5f*(zk.x - zr.x), 0.5f*(zk.y - zr.y));
                row[sk] = cadd(e, cmul(twU[k], make_float2(d.y, -d.x)));
            }
            if (k2 != k) {
                const float2 zr = cconj(zk);
                const float2 e = make_float2(0.5f*(zk2.x + zr.x), 0.5f*(zk2.y + zr.y));
                const float2 d = make_float2(0.5f*(zk2.x - zr.x), 0.5f*(zk2.y - zr.y));
                row[sk2] = cadd(e, cmul(twU[k2], make_float2(d.y, -d.x)));
            }
        }
    }
    __syncthreads();
    // y transform of the MC columns; column kz lives at slot slotOf(kz) (column M at M)
    auto colOff = [](int kz) { return kz == M ? M : slotOf<R1Z, R2Z>(kz); };
    fft3Pass1<R1Y, R2Y, false>(buf, MC, MC, colOff, twY);
    __syncthreads();
    if (peers.ranks > 0) {
        fft3Pass2Store<R1Y, R2Y, false, false>(buf, MC, MC, colOff, [&](int kz, int ky, float2 v) {
            const int q = ky / nyl;
            peers.dst[q][(((size_t) peers.slot*nxl + blockIdx.x)*nyl + (ky - q*nyl))*MC + kz] = v;
        });
    } else if (nxl > 0) {
        fft3Pass2Store<R1Y, R2Y, false, false>(buf, MC, MC, colOff, [&](int kz, int ky, float2 v) { out[slabIndex(nxl, nyl, blockIdx.x, ky, MC) + kz] = v; });
    } else {
        float2* dstp = out + (size_t) blockIdx.x*NY*MC;
        fft3Pass2Store<R1Y, R2Y, false, false>(buf, MC, MC, colOff, [&](int kz, int ky, float2 v) { dstp[ky*MC + kz] = v; });
    }
}

// half-complex plane x -> real grid plane x (unnormalised, like cufftExecC2R).  Same shared memory as the forward kernel.
template <int NY, int R1Y, int R2Y, int NZ, int R1Z, int R2Z>
__global__ void __launch_bounds__(MPID_FFT2_MAX_THREADS)
k_fft3_planes_backward(const float2* __restrict__ in, float* __restrict__ grid, const float2* __restrict__ tw, int nxl, int nyl) {
    constexpr int M = NZ/2, MC = M + 1;
    extern __shared__ float2 fftsm[];
    float2* buf = fftsm;
    float2* twY = buf + NY*MC;
    float2* twZ = twY + NY;
    float2* twU = twZ + M;
    fft2LoadTable(twY, NY, tw);
    fft2LoadTable(twZ, M, tw);
    fft2LoadTablePart(twU, NZ, M, tw);
    if (nxl > 0) {
        for (int t = threadIdx.x; t < NY*MC; t += blockDim.x) { const int ky = t / MC; buf[t] = in[slabIndex(nxl, nyl, blockIdx.x, ky, MC) + (t - ky*MC)]; }
    } else {
        const float2* src = in + (size_t) blockIdx.x*NY*MC;
        for (int t = threadIdx.x; t < NY*MC; t += blockDim.x) buf[t] = src[t];
    }
    __syncthreads();
    auto colOff = [](int kz) { return kz; };
    fft3Pass1<R1Y, R2Y, true>(buf, MC, MC, colOff, twY);
    __syncthreads();
    fft3Pass2InPlace<R1Y, R2Y, true>(buf, MC, MC, colOff);          // row slot r now holds y = elemOf<R1Y, R2Y>(r)
    __syncthreads();
    // Z[k] = (X[k] + conj X[M-k]) + i w^-k (X[k] - conj X[M-k]), k = 0..M-1, pairwise in place (k = 0 pairs with column M);
    // a length-M backward transform then gives x_{2j} + i x_{2j+1} scaled by NZ, the unnormalised C2R result
    for (int t = threadIdx.x; t < NY*(M/2 + 1); t += blockDim.x) {
        const int k = t / NY, r = t - k*NY;
        float2* row = buf + r*MC;
        const int k2 = M - k;
        const float2 xk = row[k], xk2 = row[k2];
        {
            const float2 b = cconj(xk2);
            const float2 s = cadd(xk, b), d = csub(xk, b);
            const float2 wd = k == 0 ? d : cmul(cconj(twU[k]), d);
            row[k] = make_float2(s.x - wd.y, s.y + wd.x);
        }
        if (k != 0 && k2 != k) {
            const float2 b = cconj(xk);
            const float2 s = cadd(xk2, b), d = csub(xk2, b);
            const float2 wd = cmul(cconj(twU[k2]), d);
            row[k2] = make_float2(s.x - wd.y, s.y + wd.x);
        }
    }
    __syncthreads();
    auto rowOff = [](int r) { return r*MC; };
    fft3Pass1<R1Z, R2Z, true>(buf, NY, 1, rowOff, twZ);
    __syncthreads();
    float2* plane = reinterpret_cast<float2*>(grid + (size_t) blockIdx.x*NY*NZ);
    fft3Pass2Store<R1Z, R2Z, true, true>(buf, NY, 1, rowOff, [&](int r, int j, float2 v) { plane[elemOf<R1Y, R2Y>(r)*M + j] = v; });
}

// =====================================================================================================
// Fourth generation: one x plane per thread-block CLUSTER.  The C CTAs of a cluster split the plane's rows for the z
// transforms and its kz columns for the y transforms; the transposition between the two goes through distributed shared
// memory (every CTA reads the column block it owns out of the row blocks of its C-1 neighbours with ld.shared::cluster).
// A 224 x 224 plane then needs 2 x 52 KB per CTA instead of 206 KB in one, four times as many warps work on a plane, and
// a rank that holds only 28 planes of a slab-decomposed pass still fills 112 SMs.
//   forward :  rows -> A, z passes in A, untangle A -> B (natural kz order) | cluster.sync | columns of all B -> A,
//              y passes in A -> global | cluster.sync (nobody leaves while its B is still being read)
//   backward:  columns -> A, y passes in A (row slots) | cluster.sync | rows of all A -> B, untangle in B, z passes -> global
//              | cluster.sync
// =====================================================================================================
template <int NY, int R1Y, int R2Y, int NZ, int R1Z, int R2Z, int C>
__global__ void __launch_bounds__(MPID_FFT2_MAX_THREADS)
k_fft4_planes_forward(const float* __restrict__ grid, float2* __restrict__ out, const float2* __restrict__ tw, int nxl, int nyl, SlabPeers peers) {
    namespace cg = cooperative_groups;
    constexpr int M = NZ/2, MC = M + 1, NYC = NY/C, W = (MC + C - 1)/C;
    constexpr int REGION = (NYC*MC > NY*W ? NYC*MC : NY*W);
    static_assert(R1Y*R2Y == NY && R1Z*R2Z == M && NY % C == 0, "radix split");
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int) cluster.block_rank();
    const int plane = blockIdx.x / C;
    extern __shared__ float2 fftsm[];
    float2* A = fftsm;
    float2* B = A + REGION;
    float2* twY = B + REGION;
    float2* twZ = twY + NY;
    float2* twU = twZ + M;
    fft2LoadTable(twY, NY, tw);
    fft2LoadTable(twZ, M, tw);
    fft2LoadTablePart(twU, NZ, M, tw);
    // this CTA's rows y = c NYC .. (c+1) NYC - 1, NZ reals read as M complex numbers
    const float2* src = reinterpret_cast<const float2*>(grid + ((size_t) plane*NY + (size_t) c*NYC)*NZ);
    for (int t = threadIdx.x; t < NYC*M; t += blockDim.x) A[(t / M)*MC + (t % M)] = src[t];
    __syncthreads();
    auto rowOff = [](int y) { return y*MC; };
    fft3Pass1<R1Z, R2Z, false>(A, NYC, 1, rowOff, twZ);
    __syncthreads();
    fft3Pass2InPlace<R1Z, R2Z, false>(A, NYC, 1, rowOff);
    __syncthreads();
    // untangle A (element k at slot slotOf(k)) -> B in natural kz order
    for (int t = threadIdx.x; t < NYC*MC; t += blockDim.x) {
        const int k = t / NYC, y = t - k*NYC;
        const float2* row = A + y*MC;
        const float2 zk = row[k == M ? 0 : slotOf<R1Z, R2Z>(k)];
        const float2 zr = cconj(row[k == 0 ? 0 : slotOf<R1Z, R2Z>(M - k)]);
        const float2 e = make_float2(0.5f*(zk.x + zr.x), 0.5f*(zk.y + zr.y));
        const float2 d = make_float2(0.5f*(zk.x - zr.x), 0.5f*(zk.y - zr.y));
        const float2 w = k == M ? make_float2(-1.f, 0.f) : twU[k];
        B[y*MC + k] = cadd(e, cmul(w, make_float2(d.y, -d.x)));
    }
    cluster.sync();
    // this CTA's columns kz = c W .. : gather them from the row blocks of all C CTAs (distributed shared memory)
    const int kz0 = c*W, ncol = min(W, MC - kz0);
    for (int r = 0; r < C; r++) {
        const float2* rb = cluster.map_shared_rank(B, (unsigned) r);
        for (int t = threadIdx.x; t < NYC*ncol; t += blockDim.x) {
            const int y = t / ncol, q = t - y*ncol;
            A[(r*NYC + y)*W + q] = rb[y*MC + kz0 + q];
        }
    }
    __syncthreads();
    auto colOff = [](int q) { return q; };
    fft3Pass1<R1Y, R2Y, false>(A, ncol, W, colOff, twY);
    __syncthreads();
    if (peers.ranks > 0) {
        fft3Pass2Store<R1Y, R2Y, false, false>(A, ncol, W, colOff, [&](int q, int ky, float2 v) {
            const int dq = ky / nyl;
            peers.dst[dq][(((size_t) peers.slot*nxl + plane)*nyl + (ky - dq*nyl))*MC + kz0 + q] = v;
        });
    } else if (nxl > 0) {
        fft3Pass2Store<R1Y, R2Y, false, false>(A, ncol, W, colOff, [&](int q, int ky, float2 v) { out[slabIndex(nxl, nyl, plane, ky, MC) + kz0 + q] = v; });
    } else {
        float2* dstp = out + (size_t) plane*NY*MC;
        fft3Pass2Store<R1Y, R2Y, false, false>(A, ncol, W, colOff, [&](int q, int ky, float2 v) { dstp[ky*MC + kz0 + q] = v; });
    }
    cluster.sync();
}

template <int NY, int R1Y, int R2Y, int NZ, int R1Z, int R2Z, int C>
__global__ void __launch_bounds__(MPID_FFT2_MAX_THREADS)
k_fft4_planes_backward(const float2* __restrict__ in, float* __restrict__ grid, const float2* __restrict__ tw, int nxl, int nyl) {
    namespace cg = cooperative_groups;
    constexpr int M = NZ/2, MC = M + 1, NYC = NY/C, W = (MC + C - 1)/C;
    constexpr int REGION = (NYC*MC > NY*W ? NYC*MC : NY*W);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int) cluster.block_rank();
    const int plane = blockIdx.x / C;
    extern __shared__ float2 fftsm[];
    float2* A = fftsm;
    float2* B = A + REGION;
    float2* twY = B + REGION;
    float2* twZ = twY + NY;
    float2* twU = twZ + M;
    fft2LoadTable(twY, NY, tw);
    fft2LoadTable(twZ, M, tw);
    fft2LoadTablePart(twU, NZ, M, tw);
    // this CTA's columns kz = c W .., all ky
    const int kz0 = c*W, ncol = min(W, MC - kz0);
    if (nxl > 0) {
        for (int t = threadIdx.x; t < NY*ncol; t += blockDim.x) { const int ky = t / ncol, q = t - ky*ncol; A[ky*W + q] = in[slabIndex(nxl, nyl, plane, ky, MC) + kz0 + q]; }
    } else {
        const float2* src = in + (size_t) plane*NY*MC;
        for (int t = threadIdx.x; t < NY*ncol; t += blockDim.x) { const int ky = t / ncol, q = t - ky*ncol; A[ky*W + q] = src[ky*MC + kz0 + q]; }
    }
    __syncthreads();
    auto colOff = [](int q) { return q; };
    fft3Pass1<R1Y, R2Y, true>(A, ncol, W, colOff, twY);
    __syncthreads();
    fft3Pass2InPlace<R1Y, R2Y, true>(A, ncol, W, colOff);           // row slot s now holds y = elemOf<R1Y, R2Y>(s)
    cluster.sync();
    // this CTA's row slots s = c NYC .. : gather them from the column blocks of all C CTAs
    for (int r = 0; r < C; r++) {
        const float2* cb = cluster.map_shared_rank(A, (unsigned) r);
        const int rk0 = r*W, rn = min(W, MC - rk0);
        for (int t = threadIdx.x; t < NYC*rn; t += blockDim.x) {
            const int y = t / rn, q = t - y*rn;
            B[y*MC + rk0 + q] = cb[(c*NYC + y)*W + q];
        }
    }
    __syncthreads();
    // Z[k] = (X[k] + conj X[M-k]) + i w^-k (X[k] - conj X[M-k]), pairwise in place (k = 0 pairs with column M)
    for (int t = threadIdx.x; t < NYC*(M/2 + 1); t += blockDim.x) {
        const int k = t / NYC, r = t - k*NYC;
        float2* row = B + r*MC;
        const int k2 = M - k;
        const float2 xk = row[k], xk2 = row[k2];
        {
            const float2 b = cconj(xk2);
            const float2 s = cadd(xk, b), d = csub(xk, b);
            const float2 wd = k == 0 ? d : cmul(cconj(twU[k]), d);
            row[k] = make_float2(s.x - wd.y, s.y + wd.x);
        }
        if (k != 0 && k2 != k) {
            const float2 b = cconj(xk);
            const float2 s = cadd(xk2, b), d = csub(xk2, b);
            const float2 wd = cmul(cconj(twU[k2]), d);
            row[k2] = make_float2(s.x - wd.y, s.y + wd.x);
        }
    }
    __syncthreads();
    auto rowOff = [](int r) { return r*MC; };
    fft3Pass1<R1Z, R2Z, true>(B, NYC, 1, rowOff, twZ);
    __syncthreads();
    float2* dst = reinterpret_cast<float2*>(grid + (size_t) plane*NY*NZ);
    fft3Pass2Store<R1Z, R2Z, true, true>(B, NYC, 1, rowOff, [&](int r, int j, float2 v) { dst[elemOf<R1Y, R2Y>(c*NYC + r)*M + j] = v; });
    cluster.sync();
}

// ---- host-side dispatch over the supported sizes: x, y in {32, 64, 128, 224, 256}, z in {32, 64, 128, 224} ----------------
// Power-of-two planes use the two-buffer kernels (k_fft2_*); a plane with a 224 edge uses the single-buffer ones (k_fft3_*).
#define MPID_FFT_CLUSTER 4
struct Fft2Plan {
    bool ok = false;
    int cluster = 1;                    // CTAs per plane (k_fft4_*: thread-block cluster with DSMEM transposition)
    int nx = 0, ny = 0, nz = 0, chunk = 0, chunks = 0;
    int planeThreads = 0, xThreads = 0;
    size_t planeSmem = 0, xSmem = 0;
    void (*fwd)(const float*, float2*, const float2*, int, int, SlabPeers) = nullptr;
    void (*bwd)(const float2*, float*, const float2*, int, int) = nullptr;
    void (*xcv)(int, int, int, const float*, float2*, const float2*, int, int, SlabPeers) = nullptr;
};
template <int NY, int R1Y, int R2Y> inline bool fft2PickPlanes(Fft2Plan& p, int nz) {
    if (nz == 32)  { p.fwd = k_fft2_planes_forward<NY, R1Y, R2Y, 32, 4, 4>;  p.bwd = k_fft2_planes_backward<NY, R1Y, R2Y, 32, 4, 4>;  return true; }
    if (nz == 64)  { p.fwd = k_fft2_planes_forward<NY, R1Y, R2Y, 64, 8, 4>;  p.bwd = k_fft2_planes_backward<NY, R1Y, R2Y, 64, 8, 4>;  return true; }
    if (nz == 128) { p.fwd = k_fft2_planes_forward<NY, R1Y, R2Y, 128, 8, 8>; p.bwd = k_fft2_planes_backward<NY, R1Y, R2Y, 128, 8, 8>; return true; }
    return false;
}
template <int NY, int R1Y, int R2Y> inline bool fft3PickPlanes(Fft2Plan& p, int nz) {
    if (nz == 32)  { p.fwd = k_fft3_planes_forward<NY, R1Y, R2Y, 32, 4, 4>;   p.bwd = k_fft3_planes_backward<NY, R1Y, R2Y, 32, 4, 4>;   return true; }
    if (nz == 64)  { p.fwd = k_fft3_planes_forward<NY, R1Y, R2Y, 64, 8, 4>;   p.bwd = k_fft3_planes_backward<NY, R1Y, R2Y, 64, 8, 4>;   return true; }
    if (nz == 128) { p.fwd = k_fft3_planes_forward<NY, R1Y, R2Y, 128, 8, 8>;  p.bwd = k_fft3_planes_backward<NY, R1Y, R2Y, 128, 8, 8>;  return true; }
    if (nz == 224) { p.fwd = k_fft3_planes_forward<NY, R1Y, R2Y, 224, 16, 7>; p.bwd = k_fft3_planes_backward<NY, R1Y, R2Y, 224, 16, 7>; return true; }
    return false;
}
template <int NY, int R1Y, int R2Y> inline bool fft4PickPlanes(Fft2Plan& p, int nz) {
    constexpr int C = MPID_FFT_CLUSTER;
    if (nz == 32)  { p.fwd = k_fft4_planes_forward<NY, R1Y, R2Y, 32, 4, 4, C>;   p.bwd = k_fft4_planes_backward<NY, R1Y, R2Y, 32, 4, 4, C>;   return true; }
    if (nz == 64)  { p.fwd = k_fft4_planes_forward<NY, R1Y, R2Y, 64, 8, 4, C>;   p.bwd = k_fft4_planes_backward<NY, R1Y, R2Y, 64, 8, 4, C>;   return true; }
    if (nz == 128) { p.fwd = k_fft4_planes_forward<NY, R1Y, R2Y, 128, 8, 8, C>;  p.bwd = k_fft4_planes_backward<NY, R1Y, R2Y, 128, 8, 8, C>;  return true; }
    if (nz == 224) { p.fwd = k_fft4_planes_forward<NY, R1Y, R2Y, 224, 16, 7, C>; p.bwd = k_fft4_planes_backward<NY, R1Y, R2Y, 224, 16, 7, C>; return true; }
    return false;
}
// mode: 0 = default (two-buffer planes for power-of-two grids, single-buffer for 224), 3 = single-buffer everywhere,
// 4 = cluster planes everywhere
inline Fft2Plan fft2MakePlan(int nx, int ny, int nz, int mode = 0) {
    Fft2Plan p;
    p.nx = nx; p.ny = ny; p.nz = nz;
    int r2y = 0, r1x = 0, r2z = 0;
    bool okP = false;
    const bool gen4 = mode == 4;
    const bool gen3 = !gen4 && (mode == 3 || ny == 224 || nz == 224);
    if (gen4) {
        p.cluster = MPID_FFT_CLUSTER;
        if (ny == 32)  { okP = fft4PickPlanes<32, 8, 4>(p, nz);    r2y = 4; }
        if (ny == 64)  { okP = fft4PickPlanes<64, 8, 8>(p, nz);    r2y = 8; }
        if (ny == 128) { okP = fft4PickPlanes<128, 16, 8>(p, nz);  r2y = 8; }
        if (ny == 224) { okP = fft4PickPlanes<224, 16, 14>(p, nz); r2y = 14; }
        if (ny == 256) { okP = fft4PickPlanes<256, 16, 16>(p, nz); r2y = 16; }
        r2z = nz == 224 ? 7 : (nz == 128 ? 8 : 4);
    } else if (gen3) {
        if (ny == 32)  { okP = fft3PickPlanes<32, 8, 4>(p, nz);    r2y = 4; }
        if (ny == 64)  { okP = fft3PickPlanes<64, 8, 8>(p, nz);    r2y = 8; }
        if (ny == 128) { okP = fft3PickPlanes<128, 16, 8>(p, nz);  r2y = 8; }
        if (ny == 224) { okP = fft3PickPlanes<224, 16, 14>(p, nz); r2y = 14; }
        if (ny == 256) { okP = fft3PickPlanes<256, 16, 16>(p, nz); r2y = 16; }
    } else {
        if (ny == 32)  { okP = fft2PickPlanes<32, 8, 4>(p, nz);    r2y = 4; }
        if (ny == 64)  { okP = fft2PickPlanes<64, 8, 8>(p, nz);    r2y = 8; }
        if (ny == 128) { okP = fft2PickPlanes<128, 16, 8>(p, nz);  r2y = 8; }
        if (ny == 256) { okP = fft2PickPlanes<256, 16, 16>(p, nz); r2y = 16; }
    }
    if (nx == 32)  { p.xcv = k_fft2_x_convolve<32, 8, 4>;    r1x = 8; }
    if (nx == 64)  { p.xcv = k_fft2_x_convolve<64, 8, 8>;    r1x = 8; }
    if (nx == 128) { p.xcv = k_fft2_x_convolve<128, 16, 8>;  r1x = 16; }
    if (nx == 224) { p.xcv = k_fft2_x_convolve<224, 16, 14>; r1x = 16; }
    if (nx == 256) { p.xcv = k_fft2_x_convolve<256, 16, 16>; r1x = 16; }
    if (!okP || !p.xcv) return p;
    const int m = nz/2, mc = m + 1;
    if (gen4) {
        const int C = MPID_FFT_CLUSTER, nyc = ny/C, w = (mc + C - 1)/C;
        const size_t region = std::max((size_t) nyc*mc, (size_t) ny*w);
        p.planeSmem = (2*region + ny + 2*m)*sizeof(float2);
        const int tasks = std::max(std::max(nyc*r2z, w*r2y), 64);
        p.planeThreads = std::min(MPID_FFT2_MAX_THREADS, (tasks + 31)/32*32);
    } else {
        p.planeSmem = ((size_t) (gen3 ? 1 : 2)*ny*mc + ny + 2*m)*sizeof(float2);
        // one round of the widest y pass per block when it fits, else two
        int tasks = mc*r2y;
        if (tasks > MPID_FFT2_MAX_THREADS) tasks = (tasks + 1)/2;
        p.planeThreads = std::min(MPID_FFT2_MAX_THREADS, (tasks + 31)/32*32);
    }
    if (p.planeSmem > 226*1024) return p;
    p.chunks = (mc + 11)/12;
    p.chunk = (mc + p.chunks - 1)/p.chunks;
    p.chunks = (mc + p.chunk - 1)/p.chunk;
    const int S = p.chunk | 1;
    p.xSmem = ((size_t) 2*nx*S + nx)*sizeof(float2);
    p.xThreads = std::min(256, (p.chunk*r1x + 31)/32*32);
    if (cudaFuncSetAttribute((const void*) p.fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p.planeSmem) != cudaSuccess) return p;
    if (cudaFuncSetAttribute((const void*) p.bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p.planeSmem) != cudaSuccess) return p;
    if (cudaFuncSetAttribute((const void*) p.xcv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) p.xSmem) != cudaSuccess) return p;
    p.ok = true;
    return p;
}

} // namespace mpid
#endif
