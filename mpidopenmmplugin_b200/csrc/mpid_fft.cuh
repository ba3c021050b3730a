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

} // namespace mpid
#endif
