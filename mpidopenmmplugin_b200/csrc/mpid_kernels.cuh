// MPIDB200 -- sm_100a kernels of the MPIDForce hot path (one kernel per stage).
//
// Layout conventions (all per-atom arrays are in the SORTED order produced by the cell sort unless a
// name ends in "Orig"):
//   posS      double4  wrapped position (x,y,z, site class)      -- pair kernels, PME
//   posF      float4   same, single precision                    -- neighbour-list pre-test only
//   cart      real[20] lab Cartesian moments   q d(3) Q(6) O(10) -- permanent-field kernel, PME spread
//   pk        real[16] packed traceless moments (packPairMoments) -- energy kernel
//   mud       real4    (mu_x, mu_y, mu_z, 1/damping factor or 0) -- induced-field and energy kernels
//   field/... double   accumulators written by exactly one thread per atom (no atomics)
//   force/torque/energy  64-bit fixed point (2^32), atomics, order independent => deterministic
// Neighbour list: full CSR list for the gather-style field kernels, flat i-major half list for the
// energy kernel.  Entries pack the sorted index of j (26 bits), a bare-charge flag and the periodic image code (5 bits).
#ifndef MPIDB200_KERNELS_CUH_
#define MPIDB200_KERNELS_CUH_

#include "mpid_math.h"
#include <cuda_runtime.h>

namespace mpid {

#define MPID_JMASK 0x03FFFFFFu        // sorted index of j: 26 bits (67 M atoms)
#define MPID_SIMPLE_BIT 0x04000000u   // bit 26: j is a bare charge ("simple" site), so list consumers need not look it up
#define MPID_CODE_SHIFT 27            // bits 27-31: periodic image code (0..26)
#define MPID_FIXED_SCALE 4294967296.0
#define MPID_MAX_HISTORY 20

struct DevParams {
    int n;
    int method, polarization;
    int ncell[3];
    int reach[3];                // neighbour cells scanned on each side per dimension (0: single cell)
    int nbrCap;                  // per-atom capacity of the neighbour list
    int grid[3];
    int numRanks, rank;          // multi-GPU row partition
    int rowBegin, rowEnd;        // sorted-atom range owned by this rank
    double cutoff, cutoff2;
    double alpha, defaultThole, scale14;
    double selfFieldTerm;        // (4/3) alpha^3 / sqrt(pi)
    Box box;
    PmeGeom geom;
    double shift[27][3];         // lattice translation of each image code
};

template <typename real> struct Real4;
template <> struct Real4<float>  { typedef float4 type; };
template <> struct Real4<double> { typedef double4 type; };

__device__ __forceinline__ void atomicAddFixed(unsigned long long* p, double v) {
    atomicAdd(p, (unsigned long long) __double2ll_rn(v*MPID_FIXED_SCALE));
}
__device__ __forceinline__ double fixedToDouble(unsigned long long v) {
    return (double) ((long long) v)*(1.0/MPID_FIXED_SCALE);
}

// ---------------------------------------------------------------------------------------------------
// Stage 0: wrap, bin, sort support
// ---------------------------------------------------------------------------------------------------
__global__ void k_wrap_cells(DevParams P, const double* __restrict__ posOrig, double* __restrict__ poswOrig,
                             int* __restrict__ cellKey, int* __restrict__ atomIdx) {
    int o = blockIdx.x*blockDim.x + threadIdx.x;
    if (o >= P.n) return;
    double x = posOrig[3*o], y = posOrig[3*o+1], z = posOrig[3*o+2];
    int key = 0;
    if (P.method == PME) {
        double s = floor(z*P.box.rc[2]);
        x -= P.box.c[0]*s; y -= P.box.c[1]*s; z -= P.box.c[2]*s;
        s = floor(y*P.box.rb[1]);
        x -= P.box.b[0]*s; y -= P.box.b[1]*s;
        s = floor(x*P.box.ra[0]);
        x -= P.box.a[0]*s;
        double f[3];
        f[0] = x*P.box.ra[0] + y*P.box.rb[0] + z*P.box.rc[0];
        f[1] = x*P.box.ra[1] + y*P.box.rb[1] + z*P.box.rc[1];
        f[2] = x*P.box.ra[2] + y*P.box.rb[2] + z*P.box.rc[2];
        int c[3];
        for (int d = 0; d < 3; d++) {
            int v = (int) floor(f[d]*P.ncell[d]);
            c[d] = min(max(v, 0), P.ncell[d]-1);
        }
        key = (c[0]*P.ncell[1] + c[1])*P.ncell[2] + c[2];
    }
    poswOrig[3*o] = x; poswOrig[3*o+1] = y; poswOrig[3*o+2] = z;
    cellKey[o] = key;
    atomIdx[o] = o;
}

// Everything the neighbour search needs from the sort, in one pass over the sorted atoms: the inverse permutation,
// wrapped positions with the site class (double and float), the packed class counters for the scan, damping
// parameters, and cellStart[c] = first sorted atom whose cell key is >= c (c = 0..numCells) from the key boundaries.
// The site class is static (it follows from the parameters, flagOrig is filled by set_particles), so the
// lab-frame moments are NOT on this path: k_lab_frame runs beside the neighbour search on the second stream.
template <typename real>
__global__ void k_sorted_sites(int n, int numCells, const int* __restrict__ order, const int* __restrict__ sortedKey,
                               const double* __restrict__ poswOrig, const int* __restrict__ flagOrig,
                               const double* __restrict__ damp, const double* __restrict__ thole,
                               int* __restrict__ inv, double4* __restrict__ posS, float4* __restrict__ posF,
                               int* __restrict__ flagS, unsigned long long* __restrict__ classPacked,
                               double2* __restrict__ dampTholeD, typename Real4<real>::type* __restrict__ mud,
                               int* __restrict__ cellStart) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = order[s];
    inv[o] = s;
    const int key = sortedKey[s];
    const int prev = s > 0 ? sortedKey[s-1] : -1;
    for (int c = prev + 1; c <= key; c++) cellStart[c] = s;
    if (s == n - 1) {
        for (int c = key + 1; c <= numCells; c++) cellStart[c] = n;
        classPacked[n] = 0ull;
    }
    const double x = poswOrig[3*o], y = poswOrig[3*o+1], z = poswOrig[3*o+2];
    // site class: bit 0 = polarizable (non-zero lab polarizability), bit 1 = "simple" (charge only, never polarized)
    const int flag = flagOrig[o];
    flagS[s] = flag;
    // low word counts polarizable sites, high word bare-charge sites: one 64-bit scan ranks both classes
    classPacked[s] = (unsigned long long) (flag & 1) | ((unsigned long long) ((flag >> 1) & 1) << 32);
    posS[s] = make_double4(x, y, z, (double) flag);
    posF[s] = make_float4((float) x, (float) y, (float) z, (float) flag);
    const double dmp = damp[o];
    dampTholeD[s] = make_double2(dmp, thole[o]);
    typename Real4<real>::type m;
    m.x = 0; m.y = 0; m.z = 0; m.w = dmp != 0.0 ? (real) (1.0/dmp) : real(0);   // inverse damping factor
    mud[s] = m;
}

// ---------------------------------------------------------------------------------------------------
// Stage 1: lab-frame moments (one thread per sorted atom)
// ---------------------------------------------------------------------------------------------------
struct ParticleParams {       // original order, as handed to mpidb200_set_particles
    const double* charge; const double* dipole; const double* quadrupole; const double* octopole;
    const int* axis; const int* atomZ; const int* atomX; const int* atomY;
    const double* thole; const double* alpha; const double* damp;
};

// Per-atom records are written through a per-warp shared-memory tile so that the 32 consecutive sorted atoms of a
// warp store each output array as one contiguous, coalesced block (a thread-per-record store touches 32 lines per
// instruction and throttles the LSU).
template <typename TD, int W>
__device__ __forceinline__ void storeWarpRows(double (*tile)[21], int lane, int nValid, const double* vals, TD* dst) {
#pragma unroll
    for (int k = 0; k < W; k++) tile[lane][k] = vals[k];
    __syncwarp();
    for (int idx = lane; idx < nValid*W; idx += 32) dst[idx] = (TD) tile[idx/W][idx % W];
    __syncwarp();
}

template <typename real>
__global__ void __launch_bounds__(128)
k_lab_frame(DevParams P, ParticleParams pp, int framelessFix, const int* __restrict__ order,
            const double* __restrict__ posOrig,
            double* __restrict__ cartD, double* __restrict__ pkD, real* __restrict__ cartR, real* __restrict__ pkR,
            double* __restrict__ sphD, double* __restrict__ alphaLab, int* __restrict__ aniso, int sBegin, int sEnd) {
    // sorted atoms [sBegin, sEnd): everything on one rank; with several ranks the rows a rank owns plus the cell columns
    // its neighbour search reaches into (nobody else's moments are ever read there)
    __shared__ double tiles[4][32][21];
    const int s = sBegin + blockIdx.x*blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = s - lane;                                  // first atom of this warp
    if (s0 >= sEnd) return;
    const int nValid = min(32, sEnd - s0);
    const bool valid = s < sEnd;
    double c[20], pk[16], sph[16], alpha[6];
    if (valid) {
        const int o = order[s];
        const int az = pp.atomZ[o], ax = pp.atomX[o], ay = pp.atomY[o];
        const double* pi = posOrig + 3*o;
        const double* pz = az >= 0 ? posOrig + 3*az : pi;
        const double* px = ax >= 0 ? posOrig + 3*ax : pi;
        const double* py = ay >= 0 ? posOrig + 3*ay : pi;
        LabAtom a;
        labFrameAtom(pi, pz, px, py, pp.axis[o], az, ax, ay, pp.charge[o], pp.dipole + 3*o, pp.quadrupole + 6*o,
                     pp.octopole + 10*o, pp.alpha + 3*o, a);
        if (framelessFix && az < 0) {
            a.alpha[0] = pp.alpha[3*o]; a.alpha[3] = pp.alpha[3*o+1]; a.alpha[5] = pp.alpha[3*o+2];
        }
        c[0] = a.charge;
        for (int k = 0; k < 3; k++) c[1+k] = a.dip[k];
        for (int k = 0; k < 6; k++) c[4+k] = a.quad[k];
        for (int k = 0; k < 10; k++) c[10+k] = a.oct[k];
        packPairMoments(a, pk);
        for (int k = 0; k < 16; k++) sph[k] = a.sph[k];
        for (int k = 0; k < 6; k++) alpha[k] = a.alpha[k];
        aniso[s] = a.aniso;
    }
    double (*tile)[21] = tiles[warp];
    storeWarpRows<double, 20>(tile, lane, nValid, c, cartD + 20*(size_t) s0);
    storeWarpRows<double, 16>(tile, lane, nValid, pk, pkD + 16*(size_t) s0);
    if ((void*) cartR != (void*) cartD) {
        storeWarpRows<real, 20>(tile, lane, nValid, c, cartR + 20*(size_t) s0);
        storeWarpRows<real, 16>(tile, lane, nValid, pk, pkR + 16*(size_t) s0);
    }
    storeWarpRows<double, 16>(tile, lane, nValid, sph, sphD + 16*(size_t) s0);
    storeWarpRows<double, 6>(tile, lane, nValid, alpha, alphaLab + 6*(size_t) s0);
}

// site-class bookkeeping from the exclusive scan of the packed flags: rank of every sorted atom within the
// polarizable / bare-charge ("simple") / full (= not simple) classes, and the three compact lists
__global__ void k_class_lists(int n, const int* __restrict__ flagS, const unsigned long long* __restrict__ scanned,
                              int* __restrict__ polRank, int* __restrict__ simpleRank, int* __restrict__ fullRank,
                              int* __restrict__ polList, int* __restrict__ simpleList, int* __restrict__ fullList,
                              const int* __restrict__ order, const int* __restrict__ inv, const int* __restrict__ spStart,
                              const int* __restrict__ spPartner, int4* __restrict__ spSorted) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s > n) return;
    const unsigned long long v = scanned[s];
    const int rp = (int) (v & 0xffffffffull), rs = (int) (v >> 32), rf = s - rs;
    polRank[s] = rp; simpleRank[s] = rs; fullRank[s] = rf;
    if (s == n) return;
    {   // sorted indices of up to four covalently scaled partners (x = -2 flags "more than four: use the list")
        const int o = order[s];
        const int k0 = spStart[o], k1 = spStart[o+1];
        int4 q = make_int4(-1, -1, -1, -1);
        if (k1 - k0 > 4) q.x = -2;
        else {
            if (k1 - k0 > 0) q.x = inv[spPartner[k0]];
            if (k1 - k0 > 1) q.y = inv[spPartner[k0+1]];
            if (k1 - k0 > 2) q.z = inv[spPartner[k0+2]];
            if (k1 - k0 > 3) q.w = inv[spPartner[k0+3]];
        }
        spSorted[s] = q;
    }
    const int flag = flagS[s];
    if (flag & 1) polList[rp] = s;
    if (flag & 2) simpleList[rs] = s; else fullList[rf] = s;
}

// ---------------------------------------------------------------------------------------------------
// Stage 2: neighbour list (one warp per sorted atom, single pass)
// ---------------------------------------------------------------------------------------------------
// Layout: atom i owns nbr[i*cap .. (i+1)*cap).  Neighbours with a higher sorted index ("upper") are
// packed from the front, the others from the back, so the front run doubles as the half list of the
// energy kernel and the field kernels walk both runs.  counts[2*i] = upper, counts[2*i+1] = lower.
// A cheap FP32 test settles every candidate that is not within 1e-4 nm of the cutoff sphere; the few
// that are get the reference's own FP64 test on the raw positions, so the pair set is exactly
// { i<j : |minimg(r_j - r_i)|^2 <= rc^2 } as decided by MPIDReferencePmeForce (:2829, :4178, :4350).
// Pairs with a covalent scale (1-2, 1-3, 1-4) are left out: they live in the static special list.
// ROUND = false: every periodic dimension has >= 2*reach+1 cells, so the image of a neighbour cell is
// known from the cell wrap and no per-candidate rounding is needed.  ROUND = true: generic path
// (small boxes, no-cutoff all-pairs) with a per-candidate minimum-image search.
template <bool ROUND>
__global__ void __launch_bounds__(256)
k_neighbor_list(DevParams P, const float4* __restrict__ posF, const double* __restrict__ posOrig,
                const int* __restrict__ order, const int* __restrict__ sortedKey, const int* __restrict__ cellStart,
                const int* __restrict__ spStart, const int* __restrict__ spPartner, const int4* __restrict__ spSorted,
                const int* __restrict__ polRank, int polBegin,
                unsigned* __restrict__ nbr, uint4* __restrict__ counts, unsigned* __restrict__ polNbr, unsigned* __restrict__ polCount,
                unsigned* __restrict__ maxCount) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x*blockDim.x + threadIdx.x)/32;
    const int i = P.rowBegin + row;
    if (i >= P.rowEnd) return;
    const float4 pi = posF[i];
    const int oi = order[i];
    const int4 sp = spSorted[i];
    const bool spMany = sp.x == -2;
    int sp0 = 0, sp1 = 0;
    if (spMany) { sp0 = spStart[oi]; sp1 = spStart[oi+1]; }
    const bool pme = P.method == PME;
    int cx, cy, cz;
    {
        int key = sortedKey[i];
        cz = key % P.ncell[2]; key /= P.ncell[2];
        cy = key % P.ncell[1]; cx = key / P.ncell[1];
    }
    const float rcLo = (float) P.cutoff - 1.0e-4f, rcHi = (float) P.cutoff + 1.0e-4f;
    const float rcLo2 = rcLo > 0.f ? rcLo*rcLo : 0.f, rcHi2 = rcHi*rcHi;
    const float rax = (float) P.box.ra[0], rby = (float) P.box.rb[1], rcz = (float) P.box.rc[2];
    const float ax = (float) P.box.a[0], bx = (float) P.box.b[0], by = (float) P.box.b[1];
    const float ccx = (float) P.box.c[0], ccy = (float) P.box.c[1], ccz = (float) P.box.c[2];
    const unsigned cap = (unsigned) P.nbrCap;
    unsigned* base = nbr + (size_t) row*cap;
    unsigned nUp = 0, nLow = 0, nUpSimple = 0, nPol = 0;
    const bool iPol = ((int) pi.w & 1) != 0;
    unsigned* polBase = polNbr + (iPol ? (size_t) (polRank[i] - polBegin)*cap : 0);
    const int Rx = P.reach[0], Ry = P.reach[1], Rz = P.reach[2];
    const int wy1 = 2*Ry + 1, ncols = (2*Rx + 1)*wy1;      // <= 25 neighbour columns, one per lane
    // Lane l describes column l: its (wrapped) cell column and the periodic image that wrap implies.
    int colBase = 0, wx = 0, wy = 0;
    if (lane < ncols) {
        int X = cx + lane/wy1 - Rx, Y = cy + lane % wy1 - Ry;
        if (X < 0) { X += P.ncell[0]; wx = -1; } else if (X >= P.ncell[0]) { X -= P.ncell[0]; wx = 1; }
        if (Y < 0) { Y += P.ncell[1]; wy = -1; } else if (Y >= P.ncell[1]) { Y -= P.ncell[1]; wy = 1; }
        colBase = (X*P.ncell[1] + Y)*P.ncell[2];
    }
    // Two flat passes over the concatenated candidate ranges of all columns: pass 0 the z cells that need no wrap
    // (contiguous in the sorted order), pass 1 the wrapped remainder.  Flattening keeps all 32 lanes busy.
    for (int pass = 0; pass < 2; pass++) {
        int z0 = 0, z1 = -1, wz = 0;
        if (pass == 0) { z0 = max(cz - Rz, 0); z1 = min(cz + Rz, P.ncell[2] - 1); }
        else if (cz - Rz < 0) { z0 = cz - Rz + P.ncell[2]; z1 = P.ncell[2] - 1; wz = -1; }
        else if (cz + Rz >= P.ncell[2]) { z0 = 0; z1 = cz + Rz - P.ncell[2]; wz = 1; }
        if (z1 < z0) continue;
        int jb = 0, len = 0;
        if (lane < ncols) { jb = cellStart[colBase + z0]; len = cellStart[colBase + z1 + 1] - jb; }
        // image of the column: r_j(image) = r_j + wx a + wy b + wz c
        const float shx = wx*ax + wy*bx + wz*ccx, shy = wy*by + wz*ccy, shz = wz*ccz;
        const unsigned colCode = (unsigned) ((1 - wx)*9 + (1 - wy)*3 + (1 - wz));
        int end = len;                                   // inclusive scan of the range lengths
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL, end, off);
            if (lane >= off) end += t;
        }
        const int total = __shfl_sync(FULL, end, 31);
        const int begin = end - len;
        for (int base0 = 0; base0 < total; base0 += 32) {
            const int idx = base0 + lane;
            // r = first lane whose inclusive end exceeds idx
            int r = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int t = __shfl_sync(FULL, end, r + step - 1);
                if (t <= idx) r += step;
            }
            r = min(r, 31);
            const int rjb = __shfl_sync(FULL, jb, r), rbegin = __shfl_sync(FULL, begin, r);
            const float rsx = __shfl_sync(FULL, shx, r), rsy = __shfl_sync(FULL, shy, r), rsz = __shfl_sync(FULL, shz, r);
            unsigned code = __shfl_sync(FULL, colCode, r);
            const int j = rjb + (idx - rbegin);
            bool in = (idx < total) && (j != i);
            float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) pj = posF[j];
            if (in && pme) {
                float ddx = pj.x - pi.x, ddy = pj.y - pi.y, ddz = pj.z - pi.z;
                if (ROUND) {
                    float sz = floorf(ddz*rcz + 0.5f);
                    ddx -= ccx*sz; ddy -= ccy*sz; ddz -= ccz*sz;
                    float sy = floorf(ddy*rby + 0.5f);
                    ddx -= bx*sy; ddy -= by*sy;
                    float sx = floorf(ddx*rax + 0.5f);
                    ddx -= ax*sx;
                    code = (unsigned) (((int) sx + 1)*9 + ((int) sy + 1)*3 + ((int) sz + 1));
                } else {
                    ddx += rsx; ddy += rsy; ddz += rsz;
                }
                const float r2 = ddx*ddx + ddy*ddy + ddz*ddz;
                if (r2 > rcHi2) in = false;
                else if (r2 >= rcLo2) {
                    // borderline: the oracle's test, bit for bit, on the raw positions
                    const int oj = order[j];
                    const int lo = min(oi, oj), hi = max(oi, oj);
                    double ex = posOrig[3*hi] - posOrig[3*lo], ey = posOrig[3*hi+1] - posOrig[3*lo+1], ez = posOrig[3*hi+2] - posOrig[3*lo+2];
                    periodicDelta(P.box, ex, ey, ez);
                    in = !(dist2Exact(ex, ey, ez) > P.cutoff2);
                }
                if (code > 26u) in = false;   // cannot happen for wrapped positions; keeps the table index safe
            }
            if (!pme) code = 13;
            if (j == sp.x || j == sp.y || j == sp.z || j == sp.w) in = false;
            if (spMany && in) {
                const int oj = order[j];
                for (int k = sp0; k < sp1; k++) if (spPartner[k] == oj) in = false;
            }
            const int jflag = (int) pj.w;
            const bool upper = in && (j > i);
            const unsigned maskU = __ballot_sync(FULL, upper);
            const unsigned maskL = __ballot_sync(FULL, in && !upper);
            const unsigned maskS = __ballot_sync(FULL, upper && (jflag & 2));
            const unsigned maskP = __ballot_sync(FULL, in && iPol && (jflag & 1));
            const unsigned cu = __popc(maskU), cl = __popc(maskL);
            if (nUp + nLow + cu + cl <= cap) {
                const unsigned lt = (1u << lane) - 1u;
                const unsigned entry = (unsigned) j | (code << MPID_CODE_SHIFT) | ((jflag & 2) ? MPID_SIMPLE_BIT : 0u);
                if (upper) base[nUp + __popc(maskU & lt)] = entry;
                else if (in) base[cap - 1 - (nLow + __popc(maskL & lt))] = entry;
                if (in && iPol && (jflag & 1)) polBase[nPol + __popc(maskP & lt)] = entry;
            }
            nUp += cu; nLow += cl; nUpSimple += __popc(maskS); nPol += __popc(maskP);
        }
    }
    if (lane == 0) {
        // a row that did not fit is published as empty (the evaluation is repeated with a larger capacity; maxCount
        // carries the true size), so that no consumer ever walks past the entries that were stored
        const bool fits = nUp + nLow <= cap;
        counts[row] = fits ? make_uint4(nUp, nLow, nUpSimple, nPol) : make_uint4(0u, 0u, 0u, 0u);
        if (iPol) polCount[polRank[i] - polBegin] = fits ? nPol : 0u;
        atomicMax(maxCount, nUp + nLow);
    }
}

// Shared-memory variant for the regular periodic case (every dimension has >= 2*reach+1 cells): one CTA per
// cell.  All atoms of a cell share the same candidate set, so the CTA stages the (image-shifted) candidate
// positions of the neighbouring cells in shared memory once and each warp then scans them for one atom of the
// cell with full lane utilisation.  Output layout and pair set are identical to k_neighbor_list.
// (Eight warps per CTA: a cell of liquid water holds 6.6 atoms, 78 % of the cells finish in one round.  Twelve warps --
// 98 % in one round, queue in dynamic shared memory -- was measured slower, 1.619 against 1.546 ms per evaluation: the
// extra warps mostly idle at the barrier and cost a resident CTA per SM.)
#define MPID_NL_MAXC 1280
#define MPID_NL_MAXI 64
__global__ void __launch_bounds__(256)
k_neighbor_list_cell(DevParams P, const float4* __restrict__ posF, const double* __restrict__ posOrig,
                     const int* __restrict__ order, const int* __restrict__ cellStart,
                     const int* __restrict__ spStart, const int* __restrict__ spPartner, const int4* __restrict__ spSorted,
                     const int* __restrict__ polRank, int polBegin,
                     unsigned* __restrict__ nbr, uint4* __restrict__ counts, unsigned* __restrict__ polNbr, unsigned* __restrict__ polCount,
                     unsigned* __restrict__ maxCount) {
    __shared__ float4 cand[MPID_NL_MAXC];
    __shared__ unsigned char cflag[MPID_NL_MAXC];
    __shared__ int rJb[64], rBegin[65];
    __shared__ float rShift[64][3];
    __shared__ unsigned rCode[64];
    __shared__ unsigned ctr[MPID_NL_MAXI][4];
    __shared__ unsigned short hitq[8][MPID_NL_MAXC];      // per-warp queue of candidate slots inside the pre-test sphere
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cell = blockIdx.x;
    int cz = cell % P.ncell[2], tmp = cell / P.ncell[2];
    const int cy = tmp % P.ncell[1], cx = tmp / P.ncell[1];
    const int iBeg = max(cellStart[cell], P.rowBegin), iEnd = min(cellStart[cell+1], P.rowEnd);
    if (iBeg >= iEnd) return;
    const int Rx = P.reach[0], Ry = P.reach[1], Rz = P.reach[2];
    const int wy1 = 2*Ry + 1, ncols = (2*Rx + 1)*wy1;
    const float ax = (float) P.box.a[0], bx = (float) P.box.b[0], by = (float) P.box.b[1];
    const float ccx = (float) P.box.c[0], ccy = (float) P.box.c[1], ccz = (float) P.box.c[2];
    // candidate ranges: (pass 0: z cells that need no wrap, pass 1: wrapped remainder) x neighbour columns
    if (tid < 64) {
        int jb = 0, len = 0, wx = 0, wy = 0, wz = 0;
        const int col = tid % ncols, pass = tid / ncols;
        if (pass < 2) {
            int X = cx + col/wy1 - Rx, Y = cy + col % wy1 - Ry;
            if (X < 0) { X += P.ncell[0]; wx = -1; } else if (X >= P.ncell[0]) { X -= P.ncell[0]; wx = 1; }
            if (Y < 0) { Y += P.ncell[1]; wy = -1; } else if (Y >= P.ncell[1]) { Y -= P.ncell[1]; wy = 1; }
            const int colBase = (X*P.ncell[1] + Y)*P.ncell[2];
            int z0 = 0, z1 = -1;
            if (pass == 0) { z0 = max(cz - Rz, 0); z1 = min(cz + Rz, P.ncell[2] - 1); }
            else if (cz - Rz < 0) { z0 = cz - Rz + P.ncell[2]; z1 = P.ncell[2] - 1; wz = -1; }
            else if (cz + Rz >= P.ncell[2]) { z0 = 0; z1 = cz + Rz - P.ncell[2]; wz = 1; }
            if (z1 >= z0) { jb = cellStart[colBase + z0]; len = cellStart[colBase + z1 + 1] - jb; }
        }
        rJb[tid] = jb;
        rBegin[tid] = len;                  // lengths for now, scanned below
        rShift[tid][0] = wx*ax + wy*bx + wz*ccx; rShift[tid][1] = wy*by + wz*ccy; rShift[tid][2] = wz*ccz;
        rCode[tid] = (unsigned) ((1 - wx)*9 + (1 - wy)*3 + (1 - wz));
    }
    __syncthreads();
    if (tid < 32) {       // exclusive scan of the 64 range lengths, two per lane
        const int l0 = rBegin[2*tid], l1 = rBegin[2*tid + 1];
        int incl = l0 + l1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, off);
            if (tid >= off) incl += t;
        }
        rBegin[2*tid] = incl - l0 - l1;
        rBegin[2*tid + 1] = incl - l1;
        if (tid == 31) rBegin[64] = incl;
    }
    __syncthreads();
    const int total = rBegin[64];
    const float rcLo = (float) P.cutoff - 1.0e-4f, rcHi = (float) P.cutoff + 1.0e-4f;
    const float rcLo2 = rcLo > 0.f ? rcLo*rcLo : 0.f, rcHi2 = rcHi*rcHi;
    const unsigned cap = (unsigned) P.nbrCap;
    for (int ib0 = iBeg; ib0 < iEnd; ib0 += MPID_NL_MAXI) {          // batches of atoms of this cell (normally one)
        const int ib1 = min(ib0 + MPID_NL_MAXI, iEnd);
        for (int q = tid; q < MPID_NL_MAXI*4; q += 256) ctr[q >> 2][q & 3] = 0u;
        for (int chunk0 = 0; chunk0 < total; chunk0 += MPID_NL_MAXC) { // chunks of candidates (normally one)
            const int cnt = min(MPID_NL_MAXC, total - chunk0);
            __syncthreads();
            for (int c = tid; c < cnt; c += 256) {
                const int g = chunk0 + c;
                int r = 0;
#pragma unroll
                for (int step = 32; step > 0; step >>= 1) if (r + step < 64 && rBegin[r + step] <= g) r += step;
                const int j = rJb[r] + (g - rBegin[r]);
                const float4 p = posF[j];
                cand[c] = make_float4(p.x + rShift[r][0], p.y + rShift[r][1], p.z + rShift[r][2], __uint_as_float((unsigned) j | (rCode[r] << MPID_CODE_SHIFT) | (((int) p.w & 2) ? MPID_SIMPLE_BIT : 0u)));
                cflag[c] = (unsigned char) (int) p.w;
            }
            // pad to a multiple of 64 slots with far-away sentinels so that the pre-test loop needs no bounds check
            const int cntPad = (cnt + 63) & ~63;
            for (int c = cnt + tid; c < cntPad; c += 256) cand[c] = make_float4(1.0e18f, 0.f, 0.f, 0.f);
            __syncthreads();
            for (int i = ib0 + warp; i < ib1; i += 8) {
                const int row = i - P.rowBegin;
                const float4 pi = posF[i];
                const int4 sp = spSorted[i];
                const bool spMany = sp.x == -2;
                const int oi = order[i];
                int sp0 = 0, sp1 = 0;
                if (spMany) { sp0 = spStart[oi]; sp1 = spStart[oi+1]; }
                const bool iPol = ((int) pi.w & 1) != 0;
                unsigned* base = nbr + (size_t) row*cap;
                unsigned* polBase = polNbr + (iPol ? (size_t) (polRank[i] - polBegin)*cap : 0);
                unsigned nUp = ctr[i - ib0][0], nLow = ctr[i - ib0][1], nUpSimple = ctr[i - ib0][2], nPol = ctr[i - ib0][3];
                // phase 1: bare distance pre-test of every candidate; survivors (about a quarter) are queued in order
                unsigned short* const myq = hitq[warp];
                int qn = 0;
                const unsigned ltMask = (1u << lane) - 1u;
                for (int c0 = 0; c0 < cntPad; c0 += 64) {          // two candidates per lane and trip, branch free
                    const int ca = c0 + lane, cb = ca + 32;
                    const float4 qa = cand[ca], qb = cand[cb];
                    const float ax_ = qa.x - pi.x, ay_ = qa.y - pi.y, az_ = qa.z - pi.z;
                    const float bx_ = qb.x - pi.x, by_ = qb.y - pi.y, bz_ = qb.z - pi.z;
                    const bool hitA = ax_*ax_ + ay_*ay_ + az_*az_ <= rcHi2;
                    const bool hitB = bx_*bx_ + by_*by_ + bz_*bz_ <= rcHi2;
                    const unsigned ma = __ballot_sync(FULL, hitA), mb = __ballot_sync(FULL, hitB);
                    const int na = __popc(ma);
                    if (hitA) myq[qn + __popc(ma & ltMask)] = (unsigned short) ca;
                    if (hitB) myq[qn + na + __popc(mb & ltMask)] = (unsigned short) cb;
                    qn += na + __popc(mb);
                }
                __syncwarp();
                // phase 2: classify the survivors (self, covalent partners, borderline distances, list membership)
                for (int h0 = 0; h0 < qn; h0 += 32) {
                    const int h = h0 + lane;
                    const bool valid = h < qn;
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    int jflag = 0;
                    if (valid) { const int c = myq[h]; q = cand[c]; jflag = cflag[c]; }
                    const unsigned entry = __float_as_uint(q.w);
                    const int j = (int) (entry & MPID_JMASK);
                    bool in = valid && (j != i);
                    const float ddx = q.x - pi.x, ddy = q.y - pi.y, ddz = q.z - pi.z;
                    const float r2 = ddx*ddx + ddy*ddy + ddz*ddz;
                    if (in && r2 >= rcLo2) {
                        // borderline: the oracle's test, bit for bit, on the raw positions
                        const int oj = order[j];
                        const int lo = min(oi, oj), hi = max(oi, oj);
                        double ex = posOrig[3*hi] - posOrig[3*lo], ey = posOrig[3*hi+1] - posOrig[3*lo+1], ez = posOrig[3*hi+2] - posOrig[3*lo+2];
                        periodicDelta(P.box, ex, ey, ez);
                        in = !(dist2Exact(ex, ey, ez) > P.cutoff2);
                    }
                    if (j == sp.x || j == sp.y || j == sp.z || j == sp.w) in = false;
                    if (spMany && in) {
                        const int oj = order[j];
                        for (int k = sp0; k < sp1; k++) if (spPartner[k] == oj) in = false;
                    }
                    const bool upper = in && (j > i);
                    const unsigned maskU = __ballot_sync(FULL, upper);
                    const unsigned maskL = __ballot_sync(FULL, in && !upper);
                    const unsigned maskS = __ballot_sync(FULL, upper && (jflag & 2));
                    const unsigned maskP = __ballot_sync(FULL, in && iPol && (jflag & 1));
                    const unsigned cu = __popc(maskU), cl = __popc(maskL);
                    if (nUp + nLow + cu + cl <= cap) {
                        const unsigned lt = (1u << lane) - 1u;
                        if (upper) base[nUp + __popc(maskU & lt)] = entry;
                        else if (in) base[cap - 1 - (nLow + __popc(maskL & lt))] = entry;
                        if (in && iPol && (jflag & 1)) polBase[nPol + __popc(maskP & lt)] = entry;
                    }
                    nUp += cu; nLow += cl; nUpSimple += __popc(maskS); nPol += __popc(maskP);
                }
                __syncwarp();
                if (lane == 0) {
                    ctr[i - ib0][0] = nUp; ctr[i - ib0][1] = nLow; ctr[i - ib0][2] = nUpSimple; ctr[i - ib0][3] = nPol;
                    if (chunk0 + MPID_NL_MAXC >= total) {
                        const bool fits = nUp + nLow <= cap;       // see k_neighbor_list
                        counts[row] = fits ? make_uint4(nUp, nLow, nUpSimple, nPol) : make_uint4(0u, 0u, 0u, 0u);
                        if (iPol) polCount[polRank[i] - polBegin] = fits ? nPol : 0u;
                        atomicMax(maxCount, nUp + nLow);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// Neighbour-list reuse (Verlet skin).  The cell search above is run with the cutoff enlarged by a skin and its output
// kept as the CANDIDATE list; every evaluation -- the one that built it and the ones that reuse it -- derives the exact
// list from the candidates with k_filter_list: same FP32 pre-test, same FP64 re-test of the borderline pairs on the raw
// positions with the oracle's own formula, so the pair set is still exactly { i<j : |minimg(r_j - r_i)|^2 <= rc^2 }.
// While no atom has moved more than skin/2 since the search, every pair inside the cutoff is among the candidates, the
// sorted order is kept (no sort, no cell bookkeeping) and the periodic image recorded for a candidate is still the
// minimum image (|d| <= rc + skin < L/2).
// ---------------------------------------------------------------------------------------------------
// Sorted positions of a reuse step: the atom keeps the lattice translation it was wrapped with when the list was built,
// so pair vectors stay continuous.  Also the largest squared displacement since the build (atomicMax on float bits).
template <typename real>
__global__ void k_regather_sites(int n, const int* __restrict__ order, const double* __restrict__ posNow, const double* __restrict__ posBuild,
                                 const double* __restrict__ poswBuild, const int* __restrict__ flagS, const double2* __restrict__ dampTholeD,
                                 double4* __restrict__ posS, float4* __restrict__ posF, typename Real4<real>::type* __restrict__ mud,
                                 unsigned* __restrict__ maxDisp2Bits) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    float d2 = 0.f;
    if (s < n) {
        const int o = order[s];
        const double dx = posNow[3*(size_t) o] - posBuild[3*(size_t) o], dy = posNow[3*(size_t) o+1] - posBuild[3*(size_t) o+1], dz = posNow[3*(size_t) o+2] - posBuild[3*(size_t) o+2];
        const double x = poswBuild[3*(size_t) o] + dx, y = poswBuild[3*(size_t) o+1] + dy, z = poswBuild[3*(size_t) o+2] + dz;
        const int flag = flagS[s];
        posS[s] = make_double4(x, y, z, (double) flag);
        posF[s] = make_float4((float) x, (float) y, (float) z, (float) flag);
        const double dmp = dampTholeD[s].x;
        typename Real4<real>::type m;
        m.x = 0; m.y = 0; m.z = 0; m.w = dmp != 0.0 ? (real) (1.0/dmp) : real(0);
        mud[s] = m;
        d2 = __double2float_ru(dx*dx + dy*dy + dz*dz);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(0xffffffffu, d2, off));
    if ((threadIdx.x & 31) == 0 && d2 > 0.f) atomicMax(maxDisp2Bits, __float_as_uint(d2));
}

// One warp per row: walk the row's candidates (upper run from the front, lower run from the back, the layout of
// k_neighbor_list) and keep the pairs inside the cutoff, in the same layout and order.  The kernel is issue bound
// (ncu, profiles/r02e: 65 % issue utilisation), so the inner loop is kept lean: image shifts come from a float table in
// shared memory, the rare borderline re-test sits behind a warp-uniform branch, the bare-charge count is only taken on
// the upper run and the polarizable list only for polarizable rows (both warp-uniform conditions).
template <int RUN, bool IPOL>
__device__ __forceinline__ void filterRun(const DevParams& P, const float (*shiftF)[4], const float4* __restrict__ posF, const double* __restrict__ posOrig,
                                          const int* __restrict__ order, int i, const float4 pi, float rcLo2, float rcHi2,
                                          const unsigned* __restrict__ src, int step, unsigned len, unsigned cap, unsigned* __restrict__ base,
                                          unsigned* __restrict__ polBase, int lane, unsigned lt,
                                          unsigned& nOut, unsigned otherCount, unsigned& nUpSimple, unsigned& nPol) {
    const unsigned FULL = 0xffffffffu;
    constexpr int U = 4;
    for (unsigned k0 = 0; k0 < len; k0 += 32*U) {
        unsigned e[U];
        float4 pj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned k = k0 + 32*u + lane;
            e[u] = k < len ? src[(long long) step*(long long) k] : (31u << MPID_CODE_SHIFT);   // padding: atom 0 with image code 31 = far away
        }
#pragma unroll
        for (int u = 0; u < U; u++) pj[u] = posF[e[u] & MPID_JMASK];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (k0 + 32*u >= len) break;            // warp-uniform: the rest of this trip is padding
            const unsigned code = e[u] >> MPID_CODE_SHIFT;
            const float4 sh = *reinterpret_cast<const float4*>(shiftF[code]);
            const float ddx = (pj[u].x - pi.x) - sh.x, ddy = (pj[u].y - pi.y) - sh.y, ddz = (pj[u].z - pi.z) - sh.z;
            const float r2 = ddx*ddx + ddy*ddy + ddz*ddz;
            bool in = r2 <= rcHi2;
            if (__any_sync(FULL, in && r2 >= rcLo2)) {
                if (in && r2 >= rcLo2) {
                    // borderline: the oracle's test, bit for bit, on the raw positions
                    const int oi = order[i], oj = order[e[u] & MPID_JMASK];
                    const int lo = min(oi, oj), hi = max(oi, oj);
                    double ex = posOrig[3*(size_t) hi] - posOrig[3*(size_t) lo], ey = posOrig[3*(size_t) hi+1] - posOrig[3*(size_t) lo+1], ez = posOrig[3*(size_t) hi+2] - posOrig[3*(size_t) lo+2];
                    periodicDelta(P.box, ex, ey, ez);
                    in = !(dist2Exact(ex, ey, ez) > P.cutoff2);
                }
            }
            const unsigned maskIn = __ballot_sync(FULL, in);
            const unsigned cnt = __popc(maskIn);
            const bool room = nOut + otherCount + cnt <= cap;
            if (in && room) {
                const unsigned slot = nOut + __popc(maskIn & lt);
                if (RUN == 0) base[slot] = e[u]; else base[cap - 1 - slot] = e[u];
            }
            nOut += cnt;
            if (RUN == 0) nUpSimple += __popc(__ballot_sync(FULL, in && (e[u] & MPID_SIMPLE_BIT)));
            if (IPOL) {
                const bool jp = in && (((int) pj[u].w) & 1);
                const unsigned maskP = __ballot_sync(FULL, jp);
                if (jp && room) polBase[nPol + __popc(maskP & lt)] = e[u];
                nPol += __popc(maskP);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_filter_list(DevParams P, int candCap, const float4* __restrict__ posF, const double* __restrict__ posOrig, const int* __restrict__ order,
              const unsigned* __restrict__ cand, const uint4* __restrict__ candCounts, const int* __restrict__ polRank, int polBegin,
              unsigned* __restrict__ nbr, uint4* __restrict__ counts, unsigned* __restrict__ polNbr, unsigned* __restrict__ polCount,
              unsigned* __restrict__ maxCount) {
    __shared__ __align__(16) float shiftF[32][4];
    if (threadIdx.x < 32) {
        const int c = threadIdx.x;
        const bool real = c < 27 && P.method == PME;
        // entries 27..31 are never produced by the search; 31 marks the padding lanes and throws them far outside the cutoff
        shiftF[c][0] = real ? (float) P.shift[c][0] : (c == 31 ? 1.0e18f : 0.f);
        shiftF[c][1] = real ? (float) P.shift[c][1] : 0.f;
        shiftF[c][2] = real ? (float) P.shift[c][2] : 0.f;
        shiftF[c][3] = 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int rows = P.rowEnd - P.rowBegin;
    if (row >= rows) return;
    const int i = P.rowBegin + row;
    const float4 pi = posF[i];
    const bool pme = P.method == PME;
    const float rcLo = (float) P.cutoff - 1.0e-4f, rcHi = (float) P.cutoff + 1.0e-4f;
    // without a cutoff everything is inside and nothing is borderline
    const float rcLo2 = pme ? (rcLo > 0.f ? rcLo*rcLo : 0.f) : 3.0e38f, rcHi2 = pme ? rcHi*rcHi : 1.0e30f;
    const unsigned cap = (unsigned) P.nbrCap;
    const uint4 cc = candCounts[row];
    const unsigned* cbase = cand + (size_t) row*candCap;
    unsigned* base = nbr + (size_t) row*cap;
    const bool iPol = ((int) pi.w & 1) != 0;
    unsigned* polBase = polNbr + (iPol ? (size_t) (polRank[i] - polBegin)*cap : 0);
    unsigned nUp = 0, nLow = 0, nUpSimple = 0, nPol = 0;
    const unsigned lt = (1u << lane) - 1u;
    if (iPol) {
        filterRun<0, true>(P, shiftF, posF, posOrig, order, i, pi, rcLo2, rcHi2, cbase, 1, cc.x, cap, base, polBase, lane, lt, nUp, 0u, nUpSimple, nPol);
        filterRun<1, true>(P, shiftF, posF, posOrig, order, i, pi, rcLo2, rcHi2, cbase + candCap - 1, -1, cc.y, cap, base, polBase, lane, lt, nLow, nUp, nUpSimple, nPol);
    } else {
        filterRun<0, false>(P, shiftF, posF, posOrig, order, i, pi, rcLo2, rcHi2, cbase, 1, cc.x, cap, base, polBase, lane, lt, nUp, 0u, nUpSimple, nPol);
        filterRun<1, false>(P, shiftF, posF, posOrig, order, i, pi, rcLo2, rcHi2, cbase + candCap - 1, -1, cc.y, cap, base, polBase, lane, lt, nLow, nUp, nUpSimple, nPol);
    }
    if (lane == 0) {
        const bool fits = nUp + nLow <= cap;       // see k_neighbor_list
        counts[row] = fits ? make_uint4(nUp, nLow, nUpSimple, nPol) : make_uint4(0u, 0u, 0u, 0u);
        if (iPol) polCount[polRank[i] - polBegin] = fits ? nPol : 0u;
        atomicMax(maxCount, nUp + nLow);
    }
}

// Pair classes of the energy kernel: type = 2*simple(i) + simple(j).  typeCount[t*(rows+1) + r] = number of
// upper neighbours of row r that fall in class t (scanned per class to place the runs of the four flat lists).
__global__ void k_half_counts(DevParams P, int rows, const uint4* __restrict__ counts, const int* __restrict__ flagS,
                              unsigned* __restrict__ typeCount) {
    const int r = blockIdx.x*blockDim.x + threadIdx.x;
    if (r > rows) return;
    unsigned c[5] = {0u, 0u, 0u, 0u, 0u};
    if (r < rows) {
        const uint4 q = counts[r];
        const int si = (flagS[P.rowBegin + r] >> 1) & 1;
        c[2*si] = q.x - q.z;
        // only class 0 (full-full) is compacted into the flat list: simple-simple pairs go to k_simple_pairs and
        // simple-full pairs to k_charge_site_pairs (both gather over the per-atom lists); classes 1, 2, 4 only count
        if (si) c[4] = q.z; else c[1] = q.z;
    }
    for (int t = 0; t < 5; t++) typeCount[(size_t) t*(rows + 1) + r] = c[t];
}
// the eight numbers the host wants from a neighbour search, gathered for a single device-to-host copy:
// out[0] = largest row, out[1..5] = start of each pair class, out[6] = total number of ordinary pairs
__global__ void k_collect_totals(int rows, const unsigned* __restrict__ maxCount, const unsigned* __restrict__ typeStart, unsigned* __restrict__ out) {
    const int t = threadIdx.x;
    if (t == 0) out[0] = maxCount[0];
    else if (t <= 5) out[t] = typeStart[(size_t) (t - 1)*(rows + 1)];
    else if (t == 6) out[6] = typeStart[(size_t) 5*(rows + 1) - 1];
}
// four flat i-major half lists (one warp per atom distributes its upper run by the class of j)
__global__ void k_half_compact(DevParams P, const unsigned* __restrict__ nbr, const uint4* __restrict__ counts,
                               const float4* __restrict__ posF, const unsigned* __restrict__ typeStart, int rows,
                               unsigned listBase1, unsigned listBase2, unsigned listBase3, unsigned pairCap,
                               unsigned* __restrict__ pairI, unsigned* __restrict__ pairJ) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // virtual blocks, see k_simple_pairs
    const int warpsPerBlock = blockDim.x >> 5, numBlocks = (rows + warpsPerBlock - 1)/warpsPerBlock;
    for (int vb = blockIdx.x; vb < numBlocks; vb += gridDim.x) {
    const int row = vb*warpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) continue;
    const int i = P.rowBegin + row;
    const unsigned nUp = counts[row].x;
    const int si = ((int) posF[i].w >> 1) & 1;
    const unsigned bases[4] = {0u, listBase1, listBase2, listBase3};
    unsigned dstF = bases[2*si] + typeStart[(size_t) (2*si)*(rows + 1) + row];          // j full
    unsigned dstS = bases[2*si + 1] + typeStart[(size_t) (2*si + 1)*(rows + 1) + row];  // j simple
    const unsigned* base = nbr + (size_t) row*P.nbrCap;
    for (unsigned k0 = 0; k0 < nUp; k0 += 32) {
        const unsigned k = k0 + lane;
        const bool valid = k < nUp;
        unsigned e = 0; bool sj = false;
        if (valid) { e = base[k]; sj = (e & MPID_SIMPLE_BIT) != 0; }
        const unsigned mF = __ballot_sync(FULL, valid && !sj), mS = __ballot_sync(FULL, valid && sj);
        const unsigned lt = (1u << lane) - 1u;
        if (valid && !si && !sj) {
            const unsigned d = dstF + __popc(mF & lt);
            if (d < pairCap) { pairI[d] = (unsigned) i; pairJ[d] = e; }
        }
        dstF += __popc(mF); dstS += __popc(mS);
    }
    }
}

// ---------------------------------------------------------------------------------------------------
// Stage 3: real-space fields, gather over the full neighbour list (8 lanes per atom, no atomics)
// ---------------------------------------------------------------------------------------------------
#define MPID_LANES 8

template <typename real>
__device__ __forceinline__ void pairDelta(const DevParams& P, const double4& pi, const double4& pj, unsigned code,
                                          real& dx, real& dy, real& dz) {
    dx = (real) ((pj.x - pi.x) - P.shift[code][0]);
    dy = (real) ((pj.y - pi.y) - P.shift[code][1]);
    dz = (real) ((pj.z - pi.z) - P.shift[code][2]);
}

// Permanent-multipole field of the ordinary pairs.   reference stage: :911-934 + :2812-2920
// FROMCAND: walk the skin-padded CANDIDATE rows (layout of k_neighbor_list, stride listCap = candidate capacity) and
// apply the cutoff here -- FP32 test, the oracle's FP64 test on the raw positions for pairs within 1e-4 nm of rc, the
// same decisions k_filter_list takes.  Rows are 1.4x longer, but the kernel then does not wait for the filter.  Opt-in
// (MPIDB200_EARLY_FIXED=1): measured slower, the two issue-bound kernels only share the SMs (profiles/r02_list_reuse.md).
template <typename real, bool EWALD, bool FROMCAND>
__global__ void __launch_bounds__(256)
k_fixed_field(DevParams P, int numPol, const int* __restrict__ polList, const double4* __restrict__ posS, const real* __restrict__ cart,
              const typename Real4<real>::type* __restrict__ mud,
              const uint4* __restrict__ counts, const unsigned* __restrict__ nbr, double* __restrict__ field,
              int listCap, const int* __restrict__ order, const double* __restrict__ posOrig) {
    // only polarizable sites need the permanent field (mu = alpha.E); polList holds this rank's polarizable rows
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int rp = t/MPID_LANES;
    const int sub = t % MPID_LANES;
    const bool act = rp < numPol;
    const int i = act ? polList[rp] : 0;
    double ex = 0, ey = 0, ez = 0;
    if (act) {
        const double4 pi = posS[i];
        const real invDampI = mud[i].w;
        const uint4 cnt = counts[i - P.rowBegin];
        const unsigned nUp = cnt.x, nAll = cnt.x + cnt.y;
        const unsigned cap = FROMCAND ? (unsigned) listCap : (unsigned) P.nbrCap;
        const unsigned* base = nbr + (size_t) (i - P.rowBegin)*cap;
        const real rcLo = (real) P.cutoff - real(1.0e-4), rcHi = (real) P.cutoff + real(1.0e-4);
        const real rcLo2 = rcLo > real(0) ? rcLo*rcLo : real(0), rcHi2 = rcHi*rcHi;
        // the list entry of the next trip is fetched one trip ahead so that its latency overlaps the arithmetic
        unsigned eNext = sub < nAll ? (sub < nUp ? base[sub] : base[cap - 1 - (sub - nUp)]) : 0u;
        for (unsigned k = sub; k < nAll; k += MPID_LANES) {
            const unsigned e = eNext;
            const unsigned kn = k + MPID_LANES;
            if (kn < nAll) eNext = kn < nUp ? base[kn] : base[cap - 1 - (kn - nUp)];
            const unsigned j = e & MPID_JMASK;
            const double4 pj = posS[j];
            real dx, dy, dz;
            pairDelta<real>(P, pi, pj, e >> MPID_CODE_SHIFT, dx, dy, dz);
            const real r2 = dx*dx + dy*dy + dz*dz;
            if (FROMCAND && EWALD) {
                if (r2 > rcHi2) continue;
                if (r2 >= rcLo2) {
                    // borderline: the oracle's test, bit for bit, on the raw positions
                    const int oi = order[i], oj = order[j];
                    const int lo = min(oi, oj), hi = max(oi, oj);
                    double fx_ = posOrig[3*(size_t) hi] - posOrig[3*(size_t) lo], fy_ = posOrig[3*(size_t) hi+1] - posOrig[3*(size_t) lo+1], fz_ = posOrig[3*(size_t) hi+2] - posOrig[3*(size_t) lo+2];
                    periodicDelta(P.box, fx_, fy_, fz_);
                    if (dist2Exact(fx_, fy_, fz_) > P.cutoff2) continue;
                }
            }
            // (a one-coefficient shortcut for bare-charge partners was tried and measured slower: the 8-lane groups of a
            // warp then diverge between the two partner kinds and pay for both paths, profiles/r01r_ncu_full_96k.md)
            const typename Real4<real>::type* src = reinterpret_cast<const typename Real4<real>::type*>(cart + 20*(size_t) j);
            real c[4];
            fieldCoefficientsOrdinary<real, EWALD, 4>(r2, (real) P.alpha, (real) P.defaultThole, invDampI*mud[j].w, c);
            real m[20];
#pragma unroll
            for (int q = 0; q < 5; q++) {
                typename Real4<real>::type v = src[q];
                m[4*q] = v.x; m[4*q+1] = v.y; m[4*q+2] = v.z; m[4*q+3] = v.w;
            }
            real fx = 0, fy = 0, fz = 0;
            fixedFieldDirected<real>(m, dx, dy, dz, c, fx, fy, fz);
            ex += fx; ey += fy; ez += fz;
        }
    }
#pragma unroll
    for (int off = MPID_LANES/2; off > 0; off >>= 1) {
        ex += __shfl_xor_sync(0xffffffffu, ex, off);
        ey += __shfl_xor_sync(0xffffffffu, ey, off);
        ez += __shfl_xor_sync(0xffffffffu, ez, off);
    }
    if (act && sub == 0) { field[3*(size_t) i] = ex; field[3*(size_t) i+1] = ey; field[3*(size_t) i+2] = ez; }
}

// Field (and, for the extrapolated solver, field gradient) of the induced dipoles, ordinary pairs.
//   reference stage: :4084-4088 + :4161-4281 (PME), :1037-1048 + :962-1035 (no cutoff)
template <typename real, bool EWALD, bool GRAD>
__global__ void __launch_bounds__(256)
k_induced_field(DevParams P, int numPol, const int* __restrict__ polList, const double4* __restrict__ posS,
                const typename Real4<real>::type* __restrict__ mud,
                const unsigned* __restrict__ polCount, const unsigned* __restrict__ polNbr,
                double* __restrict__ field, double* __restrict__ grad) {
    // induced dipoles live on polarizable sites only and only polarizable sites use their field, so this kernel
    // walks the polarizable x polarizable neighbour list (mu = 0 elsewhere contributes exactly nothing)
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int rp = t/MPID_LANES;
    const int sub = t % MPID_LANES;
    const bool act = rp < numPol;
    const int i = act ? polList[rp] : 0;
    // per-lane partial sums stay in `real` (<= ~30 terms each); lanes are combined in double below
    real ax_ = 0, ay_ = 0, az_ = 0;
    real ga[6] = {0, 0, 0, 0, 0, 0};
    if (act) {
        const double4 pi = posS[i];
        const real invDampI = mud[i].w;
        const unsigned nAll = polCount[rp];
        const unsigned* base = polNbr + (size_t) rp*P.nbrCap;
        // software pipeline, two trips deep: the list entry is fetched two trips ahead and the partner's position and
        // dipole one trip ahead, so both dependent global-memory latencies overlap the arithmetic of earlier pairs
        typedef typename Real4<real>::type R4;
        unsigned e1 = sub < nAll ? base[sub] : 0u;
        unsigned e2 = sub + MPID_LANES < nAll ? base[sub + MPID_LANES] : 0u;
        double4 pjN = posS[e1 & MPID_JMASK];
        R4 mjN = mud[e1 & MPID_JMASK];
        for (unsigned k = sub; k < nAll; k += MPID_LANES) {
            const unsigned e = e1;
            const double4 pj = pjN;
            const R4 mj = mjN;
            e1 = e2;
            if (k + 2*MPID_LANES < nAll) e2 = base[k + 2*MPID_LANES];
            if (k + MPID_LANES < nAll) { pjN = posS[e1 & MPID_JMASK]; mjN = mud[e1 & MPID_JMASK]; }
            real dx, dy, dz;
            pairDelta<real>(P, pi, pj, e >> MPID_CODE_SHIFT, dx, dy, dz);
            const real r2 = dx*dx + dy*dy + dz*dz;
            real c[4];
            fieldCoefficientsOrdinary<real, EWALD, (GRAD ? 3 : 2)>(r2, (real) P.alpha, (real) P.defaultThole, invDampI*mj.w, c);
            inducedFieldDirected<real>(mj.x, mj.y, mj.z, dx, dy, dz, c, ax_, ay_, az_);
            if (GRAD) inducedFieldGradientDirected<real>(mj.x, mj.y, mj.z, dx, dy, dz, c, ga);
        }
    }
    double ex = ax_, ey = ay_, ez = az_;
    double g[6];
#pragma unroll
    for (int q = 0; q < 6; q++) g[q] = ga[q];
#pragma unroll
    for (int off = MPID_LANES/2; off > 0; off >>= 1) {
        ex += __shfl_xor_sync(0xffffffffu, ex, off);
        ey += __shfl_xor_sync(0xffffffffu, ey, off);
        ez += __shfl_xor_sync(0xffffffffu, ez, off);
        if (GRAD) {
#pragma unroll
            for (int q = 0; q < 6; q++) g[q] += __shfl_xor_sync(0xffffffffu, g[q], off);
        }
    }
    if (act && sub == 0) {
        field[3*(size_t) i] = ex; field[3*(size_t) i+1] = ey; field[3*(size_t) i+2] = ez;
        if (GRAD) {
#pragma unroll
            for (int q = 0; q < 6; q++) grad[6*(size_t) i + q] += g[q];
        }
    }
}

// Covalently scaled pairs (1-2, 1-3, 1-4): always FP64, exact reference cutoff test, one thread per
// atom over its (static) partner list, added on top of what the gather kernels wrote.
//   MODE 0: permanent field   MODE 1: induced field   MODE 2: induced field + gradient
// per-atom solver helpers (used from here on)
__device__ __forceinline__ void applyAlphaLab(const double* a, double fx, double fy, double fz, double& ox, double& oy, double& oz) {
    ox = a[0]*fx + a[1]*fy + a[2]*fz;
    oy = a[1]*fx + a[3]*fy + a[4]*fz;
    oz = a[2]*fx + a[4]*fy + a[5]*fz;
}

template <typename real>
__device__ __forceinline__ void reciprocalFieldOf(const DevParams& P, const real* __restrict__ phi, int s, double& fx, double& fy, double& fz) {
    const double p1 = phi[(size_t) 1*P.n + s], p2 = phi[(size_t) 2*P.n + s], p3 = phi[(size_t) 3*P.n + s];
    fx = -(p1*P.geom.A[0][0] + p2*P.geom.A[1][0] + p3*P.geom.A[2][0]);
    fy = -(p1*P.geom.A[0][1] + p2*P.geom.A[1][1] + p3*P.geom.A[2][1]);
    fz = -(p1*P.geom.A[0][2] + p2*P.geom.A[1][2] + p3*P.geom.A[2][2]);
}

// Field (MODE 0: of the permanent moments, 1: of the induced dipoles, 2: + its gradient) that the covalently scaled
// partners of sorted atom s produce at s, in FP64; returns false when s has no such partner.
template <int MODE>
__device__ __forceinline__ bool specialFieldAt(const DevParams& P, int s, const int* __restrict__ order, const int* __restrict__ inv,
                                               const double* __restrict__ posOrig, const int* __restrict__ spStart,
                                               const int* __restrict__ spPartner, const int* __restrict__ spClass,
                                               const double* __restrict__ cartD, const double2* __restrict__ dampTholeD,
                                               const int* __restrict__ flagS, const double* __restrict__ mu,
                                               double& ex, double& ey, double& ez, double* g) {
    const int o = order[s];
    const int k0 = spStart[o], k1 = spStart[o+1];
    if (k1 == k0) return false;
    const double2 dtI = dampTholeD[s];
    for (int k = k0; k < k1; k++) {
        const int oj = spPartner[k];
        const int cls = spClass[k];
        const int sj = inv[oj];
        // induced dipoles exist on polarizable sites only: a partner without one (e.g. the hydrogens of a water
        // oxygen) contributes exactly zero to the induced field
        if (MODE != 0 && !(flagS[sj] & 1)) continue;
        const int lo = min(o, oj), hi = max(o, oj);
        double dx = posOrig[3*hi] - posOrig[3*lo], dy = posOrig[3*hi+1] - posOrig[3*lo+1], dz = posOrig[3*hi+2] - posOrig[3*lo+2];
        if (P.method == PME) periodicDelta(P.box, dx, dy, dz);
        const double r2 = dist2Exact(dx, dy, dz);
        if (P.method == PME && r2 > P.cutoff2) continue;
        if (o == hi) { dx = -dx; dy = -dy; dz = -dz; }      // d = r_other - r_me
        const double2 dtJ = dampTholeD[sj];
        const double scale = cls == 1 ? 0.0 : P.scale14;
        const double r = sqrt(r2);
        double tc[4], c[4];
        tholeComplements<double>(dtI.x, dtJ.x, dtI.y + dtJ.y, P.defaultThole, scale == 0.0, r, tc);
        if (MODE == 0) {
            if (P.method == PME) fieldCoefficients<double, true>(r, P.alpha, scale, tc, 4, c);
            else fieldCoefficients<double, false>(r, 0.0, scale, tc, 4, c);
            fixedFieldDirected<double>(cartD + 20*(size_t) sj, dx, dy, dz, c, ex, ey, ez);
        } else {
            if (P.method == PME) fieldCoefficients<double, true>(r, P.alpha, 1.0, tc, 3, c);
            else fieldCoefficients<double, false>(r, 0.0, 1.0, tc, 3, c);
            const double mx = mu[3*(size_t) sj], my = mu[3*(size_t) sj+1], mz = mu[3*(size_t) sj+2];
            inducedFieldDirected<double>(mx, my, mz, dx, dy, dz, c, ex, ey, ez);
            if (MODE == 2) inducedFieldGradientDirected<double>(mx, my, mz, dx, dy, dz, c, g);
        }
    }
    return true;
}

template <int MODE>
__global__ void k_special_field(DevParams P, const int* __restrict__ order, const int* __restrict__ inv,
                                const double* __restrict__ posOrig, const int* __restrict__ spStart,
                                const int* __restrict__ spPartner, const int* __restrict__ spClass,
                                const double* __restrict__ cartD, const double2* __restrict__ dampTholeD, const int* __restrict__ flagS,
                                const double* __restrict__ mu, double* __restrict__ field, double* __restrict__ grad) {
    const int s = P.rowBegin + blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.rowEnd) return;
    if (!(flagS[s] & 1)) return;            // fields are only consumed at polarizable sites
    double ex = 0, ey = 0, ez = 0, g[6] = {0, 0, 0, 0, 0, 0};
    if (!specialFieldAt<MODE>(P, s, order, inv, posOrig, spStart, spPartner, spClass, cartD, dampTholeD, flagS, mu, ex, ey, ez, g)) return;
    field[3*(size_t) s] += ex; field[3*(size_t) s+1] += ey; field[3*(size_t) s+2] += ez;
    if (MODE == 2) for (int q = 0; q < 6; q++) grad[6*(size_t) s + q] += g[q];
}

// Single rank: the covalent-partner part of the permanent field and everything that follows it per atom in one pass
// -- + reciprocal field + self term, efix = alpha.E, mu0 = efix (k_special_field<0> + k_fixed_recip_mu).  The short
// per-atom kernel otherwise sits on the critical path behind the resident CTAs of the side-stream pair kernels.
template <typename real>
__global__ void k_special_field_finish(DevParams P, const int* __restrict__ order, const int* __restrict__ inv,
                                       const double* __restrict__ posOrig, const int* __restrict__ spStart,
                                       const int* __restrict__ spPartner, const int* __restrict__ spClass,
                                       const double* __restrict__ cartD, const double2* __restrict__ dampTholeD, const int* __restrict__ flagS,
                                       const real* __restrict__ phi, const double* __restrict__ alphaLab, const double* __restrict__ field,
                                       double* __restrict__ efix, double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.n) return;
    double ox = 0, oy = 0, oz = 0;
    if (flagS[s] & 1) {
        double ex = 0, ey = 0, ez = 0;
        specialFieldAt<0>(P, s, order, inv, posOrig, spStart, spPartner, spClass, cartD, dampTholeD, flagS, (const double*) nullptr, ex, ey, ez, (double*) nullptr);
        double fx = field[3*(size_t) s] + ex, fy = field[3*(size_t) s+1] + ey, fz = field[3*(size_t) s+2] + ez;
        if (P.method == PME) {
            double rx, ry, rz;
            reciprocalFieldOf<real>(P, phi, s, rx, ry, rz);
            fx += rx + P.selfFieldTerm*cartD[20*(size_t) s+1];
            fy += ry + P.selfFieldTerm*cartD[20*(size_t) s+2];
            fz += rz + P.selfFieldTerm*cartD[20*(size_t) s+3];
        }
        applyAlphaLab(alphaLab + 6*(size_t) s, fx, fy, fz, ox, oy, oz);
    }
    efix[3*(size_t) s] = ox; efix[3*(size_t) s+1] = oy; efix[3*(size_t) s+2] = oz;
    mu[3*(size_t) s] = ox; mu[3*(size_t) s+1] = oy; mu[3*(size_t) s+2] = oz;
    typename Real4<real>::type m = mud[s];
    m.x = (real) ox; m.y = (real) oy; m.z = (real) oz;
    mud[s] = m;
}

// ---------------------------------------------------------------------------------------------------
// Stage 4: pair energy / force / torque over the i-major half list (one thread per pair)
// ---------------------------------------------------------------------------------------------------
//   reference stage: :4932-4946 + :4335-4920 (PME), :2140-2158 + :1331-1893 (no cutoff)
template <typename real, bool EWALD, bool MUTUAL, bool SI, bool SJ>
__global__ void __launch_bounds__(128)
k_electrostatics(DevParams P, long long numPairs, const unsigned* __restrict__ dynCount, const unsigned* __restrict__ pairI, const unsigned* __restrict__ pairJ,
                 const double4* __restrict__ posS, const real* __restrict__ pk, const typename Real4<real>::type* __restrict__ mud,
                 const int* __restrict__ aniso,
                 unsigned long long* __restrict__ force, unsigned long long* __restrict__ torque, unsigned long long* __restrict__ energy) {
    const long long p = (long long) blockIdx.x*blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int i = -1 - lane;      // distinct dummies so idle lanes never merge with a real segment
    unsigned j = 0;
    double e = 0;
    real f[3] = {0, 0, 0}, ti[3] = {0, 0, 0}, tj[3] = {0, 0, 0};
    // dynCount: number of pairs as counted on the device (the launch is sized from the previous evaluation's count)
    if (dynCount) numPairs = min((long long) *dynCount, numPairs);
    const bool active = p < numPairs;
    if (active) {
        i = (int) pairI[p];
        const unsigned ej = pairJ[p];
        j = ej & MPID_JMASK;
        real dx, dy, dz;
        pairDelta<real>(P, posS[i], posS[j], ej >> MPID_CODE_SHIFT, dx, dy, dz);
        real qi[16], qj[16];
        typedef typename Real4<real>::type R4;
        const R4* si = reinterpret_cast<const R4*>(pk + 16*(size_t) i);
        const R4* sj = reinterpret_cast<const R4*>(pk + 16*(size_t) j);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            R4 a = si[q], b = sj[q];
            qi[4*q] = a.x; qi[4*q+1] = a.y; qi[4*q+2] = a.z; qi[4*q+3] = a.w;
            qj[4*q] = b.x; qj[4*q+1] = b.y; qj[4*q+2] = b.z; qj[4*q+3] = b.w;
        }
        const R4 mi = mud[i], mj = mud[j];
        real uI[3] = {mi.x, mi.y, mi.z}, uJ[3] = {mj.x, mj.y, mj.z};
        const real r2 = dx*dx + dy*dy + dz*dz;
        const real dampI = mi.w != real(0) ? real(1)/mi.w : real(0), dampJ = mj.w != real(0) ? real(1)/mj.w : real(0);
        e = (double) pairElectrostatics<real, EWALD, MUTUAL, SI, SJ>(qi, qj, uI, uJ, dampI, dampJ, real(0), real(0), aniso[i] != 0, aniso[j] != 0,
                                                           dx, dy, dz, r2, (real) P.alpha, (real) P.defaultThole, real(1), real(1), f, ti, tj);
    }
    // j side: scattered fixed-point atomics (a simple site feels no torque)
    if (active) {
        atomicAddFixed(&force[3*(size_t) j], (double) f[0]); atomicAddFixed(&force[3*(size_t) j+1], (double) f[1]); atomicAddFixed(&force[3*(size_t) j+2], (double) f[2]);
        if (!SJ) { atomicAddFixed(&torque[3*(size_t) j], (double) tj[0]); atomicAddFixed(&torque[3*(size_t) j+1], (double) tj[1]); atomicAddFixed(&torque[3*(size_t) j+2], (double) tj[2]); }
    }
    // i side: pairs of one i are contiguous, so a segmented warp reduction leaves one atomic per run.  The partial
    // sums (<= 32 terms) are carried in `real`: one shuffle per value and round; the cross-warp sum is fixed point.
    constexpr int NV = SI ? 3 : 6;      // a simple site feels no torque
    real v[6] = {-f[0], -f[1], -f[2], ti[0], ti[1], ti[2]};
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int io = __shfl_down_sync(0xffffffffu, i, off);
        const bool take = (lane + off < 32) && (io == i);
#pragma unroll
        for (int q = 0; q < NV; q++) {
            const real w = __shfl_down_sync(0xffffffffu, v[q], off);
            if (take) v[q] += w;
        }
    }
    const int iprev = __shfl_up_sync(0xffffffffu, i, 1);
    if (active && (lane == 0 || iprev != i)) {
        atomicAddFixed(&force[3*(size_t) i], (double) v[0]); atomicAddFixed(&force[3*(size_t) i+1], (double) v[1]); atomicAddFixed(&force[3*(size_t) i+2], (double) v[2]);
        if (!SI) { atomicAddFixed(&torque[3*(size_t) i], (double) v[3]); atomicAddFixed(&torque[3*(size_t) i+1], (double) v[4]); atomicAddFixed(&torque[3*(size_t) i+2], (double) v[5]); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if (lane == 0 && e != 0.0) atomicAddFixed(energy, e);
}

// Charge-only x charge-only pairs (e.g. H-H in water): 45 % of all pairs but a few dozen flops each, so the
// cost of the flat-list kernel would be its atomics.  Instead every simple site gathers over its own full
// neighbour list (8 lanes per site, each pair is seen from both ends) and takes half of the pair energy.
//   E = k q_i q_j B1/r,  F_i = -k q_i q_j B2 d/r^3  with  B1 = erfc(ar), B2 = B1 + 2 ar exp(-(ar)^2)/sqrt(pi)   (:4527-4534)
template <typename real, bool EWALD>
__global__ void __launch_bounds__(256)
k_simple_pairs(DevParams P, int numSimple, const int* __restrict__ simpleList, const double4* __restrict__ posS,
               const real* __restrict__ pk, const uint4* __restrict__ counts, const unsigned* __restrict__ nbr,
               unsigned long long* __restrict__ force, unsigned long long* __restrict__ energy) {
    // virtual blocks: the launch may hold fewer CTAs than the work has blocks (residency cap of the side stream)
    const int numBlocks = (numSimple*MPID_LANES + blockDim.x - 1)/blockDim.x;
    for (int vb = blockIdx.x; vb < numBlocks; vb += gridDim.x) {
    const int t = vb*blockDim.x + threadIdx.x;
    const int rs = t/MPID_LANES;
    const int sub = t % MPID_LANES;
    const bool act = rs < numSimple;
    const int i = act ? simpleList[rs] : 0;
    real fx = 0, fy = 0, fz = 0, en = 0;
    if (act) {
        const double4 pi = posS[i];
        const real qi = pk[16*(size_t) i];
        const uint4 cnt = counts[i - P.rowBegin];
        const unsigned nUp = cnt.x, nAll = cnt.x + cnt.y;
        const unsigned* base = nbr + (size_t) (i - P.rowBegin)*P.nbrCap;
        unsigned eNext = sub < nAll ? (sub < nUp ? base[sub] : base[P.nbrCap - 1 - (sub - nUp)]) : 0u;     // fetched one trip ahead
        for (unsigned k = sub; k < nAll; k += MPID_LANES) {
            const unsigned e = eNext;
            const unsigned kn = k + MPID_LANES;
            if (kn < nAll) eNext = kn < nUp ? base[kn] : base[P.nbrCap - 1 - (kn - nUp)];
            const unsigned j = e & MPID_JMASK;
            const double4 pj = posS[j];
            if (!(((int) pj.w) & 2)) continue;
            real dx, dy, dz;
            pairDelta<real>(P, pi, pj, e >> MPID_CODE_SHIFT, dx, dy, dz);
            const real r2 = dx*dx + dy*dy + dz*dz;
            const real rinv = t_rsqrt(r2);
            const real qq = real(MPID_ELECTRIC)*qi*pk[16*(size_t) j];
            real B1 = real(1), B2 = real(1);
            if (EWALD) {
                const real x = (real) P.alpha*r2*rinv;
                const real ex = t_expneg(-(x*x));
                B1 = t_erfc_ex(x, ex);
                B2 = B1 + real(2.0/MPID_SQRT_PI)*x*ex;
            }
            en += qq*B1*rinv;
            const real fr = -qq*B2*rinv*rinv*rinv;      // force on i = fr * d
            fx += fr*dx; fy += fr*dy; fz += fr*dz;
        }
    }
    double dfx = fx, dfy = fy, dfz = fz, de = 0.5*(double) en;
#pragma unroll
    for (int off = MPID_LANES/2; off > 0; off >>= 1) {
        dfx += __shfl_xor_sync(0xffffffffu, dfx, off);
        dfy += __shfl_xor_sync(0xffffffffu, dfy, off);
        dfz += __shfl_xor_sync(0xffffffffu, dfz, off);
    }
    if (act && sub == 0) {
        atomicAddFixed(&force[3*(size_t) i], dfx); atomicAddFixed(&force[3*(size_t) i+1], dfy); atomicAddFixed(&force[3*(size_t) i+2], dfz);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) de += __shfl_xor_sync(0xffffffffu, de, off);
    if ((threadIdx.x & 31) == 0 && de != 0.0) atomicAddFixed(energy, de);    }
}

// Full site x bare-charge site pairs (e.g. O-H between waters: 44 % of all pairs).  Every full site A gathers over
// its own full neighbour list (8 lanes per site) and handles the partners that are bare charges with the Cartesian
// form chargeSitePair: A's moments stay in registers, a partner costs one position and one charge, A's force and
// torque need no atomics until the end, the partner's force is three fixed-point atomics.  Each such pair is seen
// exactly once (from its full site), whichever index is larger.
//   reference stage: :4932-4946 + :4335-4920 (PME), :2140-2158 + :1331-1893 (no cutoff)
template <typename real, bool EWALD>
__global__ void __launch_bounds__(256)
k_charge_site_pairs(DevParams P, int numFull, const int* __restrict__ fullList, const double4* __restrict__ posS,
                    const real* __restrict__ pk, const typename Real4<real>::type* __restrict__ mud, const int* __restrict__ aniso,
                    const uint4* __restrict__ counts, const unsigned* __restrict__ nbr,
                    unsigned long long* __restrict__ force, unsigned long long* __restrict__ torque, unsigned long long* __restrict__ energy) {
    typedef typename Real4<real>::type R4;
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int rf = t/MPID_LANES;
    const int sub = t % MPID_LANES;
    const bool act = rf < numFull;
    const int i = act ? fullList[rf] : 0;
    real fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0, en = 0;
    if (act) {
        const double4 pi = posS[i];
        real mA[20];
        {
            real q[16];
            const R4* src = reinterpret_cast<const R4*>(pk + 16*(size_t) i);
#pragma unroll
            for (int k = 0; k < 4; k++) { const R4 v = src[k]; q[4*k] = v.x; q[4*k+1] = v.y; q[4*k+2] = v.z; q[4*k+3] = v.w; }
            unpackPairMoments<real>(q, mA);
        }
        const R4 mi = mud[i];
        const bool anisoA = aniso[i] != 0;
        const uint4 cnt = counts[i - P.rowBegin];
        const unsigned nUp = cnt.x, nAll = cnt.x + cnt.y;
        const unsigned* base = nbr + (size_t) (i - P.rowBegin)*P.nbrCap;
        unsigned eNext = sub < nAll ? (sub < nUp ? base[sub] : base[P.nbrCap - 1 - (sub - nUp)]) : 0u;     // fetched one trip ahead
        for (unsigned k = sub; k < nAll; k += MPID_LANES) {
            const unsigned e = eNext;
            const unsigned kn = k + MPID_LANES;
            if (kn < nAll) eNext = kn < nUp ? base[kn] : base[P.nbrCap - 1 - (kn - nUp)];
            const unsigned j = e & MPID_JMASK;
            const double4 pj = posS[j];
            if (!(((int) pj.w) & 2)) continue;          // partner is not a bare charge
            real rx, ry, rz;
            pairDelta<real>(P, pi, pj, e >> MPID_CODE_SHIFT, rx, ry, rz);     // r_j - r_i ; chargeSitePair wants d = r_A - r_B
            const real r2 = rx*rx + ry*ry + rz*rz;
            const real kq = real(MPID_ELECTRIC)*pk[16*(size_t) j];
            real fB[3], tA[3];
            const real phi = chargeSitePair<real, EWALD>(mA, mi.x, mi.y, mi.z, mi.w*mud[j].w, anisoA, -rx, -ry, -rz, r2,
                                                         (real) P.alpha, (real) P.defaultThole, fB, tA);
            en += kq*phi;
            const real bx = kq*fB[0], by = kq*fB[1], bz = kq*fB[2];
            fx -= bx; fy -= by; fz -= bz;
            tx += kq*tA[0]; ty += kq*tA[1]; tz += kq*tA[2];
            atomicAddFixed(&force[3*(size_t) j], (double) bx); atomicAddFixed(&force[3*(size_t) j+1], (double) by); atomicAddFixed(&force[3*(size_t) j+2], (double) bz);
        }
    }
    double v[6] = {fx, fy, fz, tx, ty, tz};
    double de = en;
#pragma unroll
    for (int off = MPID_LANES/2; off > 0; off >>= 1) {
#pragma unroll
        for (int q = 0; q < 6; q++) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    }
    if (act && sub == 0) {
        atomicAddFixed(&force[3*(size_t) i], v[0]); atomicAddFixed(&force[3*(size_t) i+1], v[1]); atomicAddFixed(&force[3*(size_t) i+2], v[2]);
        atomicAddFixed(&torque[3*(size_t) i], v[3]); atomicAddFixed(&torque[3*(size_t) i+1], v[4]); atomicAddFixed(&torque[3*(size_t) i+2], v[5]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) de += __shfl_xor_sync(0xffffffffu, de, off);
    if ((threadIdx.x & 31) == 0 && de != 0.0) atomicAddFixed(energy, de);
}

// Covalently scaled pairs: FP64, one thread per static special pair (original indices lo < hi).
template <bool MUTUAL>
__global__ void k_special_electrostatics(DevParams P, int numSpecial, const int* __restrict__ spPairLo, const int* __restrict__ spPairHi,
                                         const int* __restrict__ spPairClass, const int* __restrict__ inv,
                                         const double* __restrict__ posOrig, const double* __restrict__ pkD,
                                         const double2* __restrict__ dampTholeD, const double* __restrict__ mu, const int* __restrict__ aniso,
                                         unsigned long long* __restrict__ force, unsigned long long* __restrict__ torque,
                                         unsigned long long* __restrict__ energy, const int* __restrict__ ownList) {
    // ownList (several ranks): the pairs whose lower atom sits in this rank's rows (k_own_special), numSpecial of them
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    // no early return: the pair energies of a warp are summed with shuffles and leave through ONE atomic per warp
    // (one atomic per pair on the single energy word serialises at the L2: 95,616 of them at 95,616 atoms)
    const bool active = t < numSpecial;
    const int k = active ? (ownList ? ownList[t] : t) : 0;
    double e = 0.0;
    if (active) {
        const int lo = spPairLo[k], hi = spPairHi[k];
        double dx = posOrig[3*hi] - posOrig[3*lo], dy = posOrig[3*hi+1] - posOrig[3*lo+1], dz = posOrig[3*hi+2] - posOrig[3*lo+2];
        if (P.method == PME) periodicDelta(P.box, dx, dy, dz);
        const double r2 = dist2Exact(dx, dy, dz);
        if (!(P.method == PME && r2 > P.cutoff2)) {
            const int si = inv[lo], sj = inv[hi];
            const double scale = spPairClass[k] == 1 ? 0.0 : P.scale14;
            const double2 dtI = dampTholeD[si], dtJ = dampTholeD[sj];
            double f[3], ti[3], tj[3];
            if (P.method == PME)
                e = pairElectrostatics<double, true, MUTUAL>(pkD + 16*(size_t) si, pkD + 16*(size_t) sj, mu + 3*(size_t) si, mu + 3*(size_t) sj,
                        dtI.x, dtJ.x, dtI.y, dtJ.y, aniso[si] != 0, aniso[sj] != 0, dx, dy, dz, r2, P.alpha, P.defaultThole, scale, scale, f, ti, tj);
            else
                e = pairElectrostatics<double, false, MUTUAL>(pkD + 16*(size_t) si, pkD + 16*(size_t) sj, mu + 3*(size_t) si, mu + 3*(size_t) sj,
                        dtI.x, dtJ.x, dtI.y, dtJ.y, aniso[si] != 0, aniso[sj] != 0, dx, dy, dz, r2, 0.0, P.defaultThole, scale, scale, f, ti, tj);
            for (int q = 0; q < 3; q++) {
                atomicAddFixed(&force[3*(size_t) si + q], -f[q]);
                atomicAddFixed(&force[3*(size_t) sj + q], f[q]);
                atomicAddFixed(&torque[3*(size_t) si + q], ti[q]);
                atomicAddFixed(&torque[3*(size_t) sj + q], tj[q]);
            }
        }
    }
    // fixed point before the reduction, so that the sum does not depend on which pairs share a warp
    long long ef = __double2ll_rn(e*MPID_FIXED_SCALE);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ef += __shfl_xor_sync(0xffffffffu, ef, off);
    if ((threadIdx.x & 31) == 0 && ef != 0) atomicAdd(energy, (unsigned long long) ef);
}

// ---------------------------------------------------------------------------------------------------
// Stage 5: PME
// ---------------------------------------------------------------------------------------------------
// Scaled-fractional multipoles of the permanent moments (20 per atom).  reference: :3077-3169
template <typename real>
__global__ void k_fractional_multipoles(DevParams P, const real* __restrict__ cart, real* __restrict__ frac) {
    const int s = P.rowBegin + blockIdx.x*blockDim.x + threadIdx.x;      // this rank's rows: only they are spread
    if (s >= P.rowEnd) return;
    real m[20], f[20];
    for (int k = 0; k < 20; k++) m[k] = cart[20*(size_t) s + k];
    multipolesToFractional<real>(P.geom.A, m, f);
    for (int k = 0; k < 20; k++) frac[20*(size_t) s + k] = f[k];
}

// Six consecutive grid points of one (x,y) line as 16-byte vector reductions (red.global.add.v4.f32, sm_90+) on the
// aligned quads that cover them: two requests for three of the four alignments, three for the fourth, the unused
// lanes adding 0.0f (x + 0 = x).  The kernel is bound by the number of reduction requests the LSU queues
// (`lg_throttle`), not by the adds: against exact v4/v2/scalar pieces (2.75 requests per line) this took the
// evaluation from 1.577 to 1.546 ms at 95,616 atoms.  `room` = floats from p to the end of the row: a quad that
// would reach past the row end (and, on the last row, past the grid) is issued as scalars instead.
__device__ __forceinline__ void redLine6(float* p, const float* v, int room) {
    const unsigned mis = (unsigned) ((reinterpret_cast<size_t>(p) >> 2) & 3);
    float* q = p - mis;                                   // first aligned quad
    const int quads = mis == 3 ? 3 : 2;
    if ((int) (4*quads - mis) > room) {                   // the padded tail would leave the row
#pragma unroll
        for (int k = 0; k < 6; k++) atomicAdd(p + k, v[k]);
        return;
    }
    float w[12];
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        // w[mis + k] = v[k] without dynamic register indexing
        w[k]     = mis == 0 ? v[k] : w[k];
        w[k + 1] = mis == 1 ? v[k] : w[k + 1];
        w[k + 2] = mis == 2 ? v[k] : w[k + 2];
        w[k + 3] = mis == 3 ? v[k] : w[k + 3];
    }
    atomicAdd(reinterpret_cast<float4*>(q), make_float4(w[0], w[1], w[2], w[3]));
    atomicAdd(reinterpret_cast<float4*>(q + 4), make_float4(w[4], w[5], w[6], w[7]));
    if (mis == 3) atomicAdd(reinterpret_cast<float4*>(q + 8), make_float4(w[8], w[9], w[10], w[11]));
}
__device__ __forceinline__ void redLine6(double* p, const double* v, int) {
#pragma unroll
    for (int k = 0; k < 6; k++) atomicAdd(p + k, v[k]);
}

// B-spline weights of the polarizable sites, once per evaluation: their positions do not change between the
// reciprocal passes of the solver iterations, and each pass used to rebuild the weights in every thread of its spread
// and gather launch.   thetaPol[r][axis][k][8] : k-th derivative (k = 0..2) of the six weights of polarizable site r
// (rank among the polarizable sites) along `axis`, two pad floats so that every (axis, k) row is one 32-byte sector;
// igridPol[r] = first grid point per axis.  The permanent-moment pass (all atoms, once per evaluation) builds its
// weights in the kernel: a record for every atom costs more HBM traffic than the arithmetic it saves (measured at
// 1,024,884 atoms: 0.39 ms to write it, gathers 0.1 ms slower each, spread 0.13 ms faster).
//   reference: computeMPIDBsplines (:3049-3075), computeBSplinePoint (:2956-3044)
#define MPID_THETA_POL (3*3*8)
template <typename real>
__global__ void __launch_bounds__(128)
k_spline_weights(DevParams P, int numPol, const int* __restrict__ polList, int recBase, const double4* __restrict__ posS,
                 real* __restrict__ thetaPol, int4* __restrict__ igridPol) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= numPol) return;
    const int r = recBase + t;
    const double4 p = posS[polList[t]];
    int ig[3]; double w[3];
    pmeAtomCell(P.box, P.geom, p.x, p.y, p.z, ig, w);
    igridPol[r] = make_int4(ig[0], ig[1], ig[2], 0);
    typedef typename Real4<real>::type real4;
#pragma unroll
    for (int axis = 0; axis < 3; axis++) {
        real th[6][5];
        bsplineWeights<real>((real) w[axis], th);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            real4 lo, hi;
            lo.x = th[0][k]; lo.y = th[1][k]; lo.z = th[2][k]; lo.w = th[3][k];
            hi.x = th[4][k]; hi.y = th[5][k]; hi.z = 0; hi.w = 0;
            real4* dp = reinterpret_cast<real4*>(thetaPol + (size_t) r*MPID_THETA_POL + (axis*3 + k)*8);
            dp[0] = lo; dp[1] = hi;
        }
    }
}
// rows (axis, k = 0..NK-1) of one site's weight record -> t[point][k]
template <typename real, int NK>
__device__ __forceinline__ void loadTheta(const real* __restrict__ rec, int axis, real (*t)[5]) {
    typedef typename Real4<real>::type real4;
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const real4* src = reinterpret_cast<const real4*>(rec + (axis*3 + k)*8);
        const real4 lo = src[0], hi = src[1];
        t[0][k] = lo.x; t[1][k] = lo.y; t[2][k] = lo.z; t[3][k] = lo.w; t[4][k] = hi.x; t[5][k] = hi.y;
    }
}

// B-spline spreading: 6 threads per atom (one per x plane), 36 grid points each.  FIXED: permanent moments of rows
// rowBegin.., weights built here from the positions; otherwise induced dipoles of the listed polarizable rows with
// the weights of k_spline_weights (record recBase + t/6).
//   reference: spreadFixedMultipolesOntoGrid (:3269-3327), spreadInducedDipolesOnGrid (:3532-3573)
template <typename real, bool FIXED>
__global__ void __launch_bounds__(192)
k_spread(DevParams P, int numRows, const int* __restrict__ rowList, int recBase, const real* __restrict__ theta,
         const int4* __restrict__ igrid, const double4* __restrict__ posS, const real* __restrict__ frac,
         const double* __restrict__ mu, real* __restrict__ grid) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int ix = t % 6;
    if (t/6 >= numRows) return;
    const int s = rowList ? rowList[t/6] : P.rowBegin + t/6;
    int4 ig;
    real ty[6][5], tz[6][5], txr[5];
    if (FIXED) {
        const double4 p = posS[s];
        int g[3]; double w[3];
        pmeAtomCell(P.box, P.geom, p.x, p.y, p.z, g, w);
        ig = make_int4(g[0], g[1], g[2], 0);
        real tx[6][5];
        bsplineWeights<real>((real) w[0], tx);
        bsplineWeights<real>((real) w[1], ty);
        bsplineWeights<real>((real) w[2], tz);
#pragma unroll
        for (int k = 0; k < 5; k++) txr[k] = tx[0][k];
#pragma unroll
        for (int a = 1; a < 6; a++)
            if (a == ix)
#pragma unroll
                for (int k = 0; k < 5; k++) txr[k] = tx[a][k];
    } else {
        const int rec = recBase + t/6;
        const real* th = theta + (size_t) rec*MPID_THETA_POL;
        ig = igrid[rec];
        loadTheta<real, 2>(th, 1, ty);
        loadTheta<real, 2>(th, 2, tz);
#pragma unroll
        for (int k = 0; k < 2; k++) txr[k] = th[k*8 + ix];
    }
    real f[20];
    if (FIXED) {
        for (int k = 0; k < 20; k++) f[k] = frac[20*(size_t) s + k];
    } else {
        const double mx = mu[3*(size_t) s], my = mu[3*(size_t) s+1], mz = mu[3*(size_t) s+2];
        f[0] = 0;
        for (int k = 0; k < 3; k++) f[1+k] = (real) (P.geom.A[k][0]*mx + P.geom.A[k][1]*my + P.geom.A[k][2]*mz);
    }
    const int nx = P.grid[0], ny = P.grid[1], nz = P.grid[2];
    int x = ig.x + ix; x -= (x >= nx) ? nx : 0;
#pragma unroll
    for (int iy = 0; iy < 6; iy++) {
        int y = ig.y + iy; y -= (y >= ny) ? ny : 0;
        real* row = grid + ((size_t) x*ny + y)*nz;
        real v[6];
#pragma unroll
        for (int iz = 0; iz < 6; iz++) v[iz] = spreadTerm<real, FIXED>(f, txr, ty[iy], tz[iz]);
        if (ig.z + 5 < nz) redLine6(row + ig.z, v, nz - ig.z);
        else {
#pragma unroll
            for (int iz = 0; iz < 6; iz++) {
                int z = ig.z + iz; z -= (z >= nz) ? nz : 0;
                atomicAdd(row + z, v[iz]);
            }
        }
    }
}

// Reciprocal convolution on the half-complex grid.   reference: performMPIDReciprocalConvolution (:3329-3366)
template <typename cplx, typename real>
__global__ void k_convolution(size_t count, const real* __restrict__ eterm, cplx* __restrict__ g) {
    const size_t k = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= count) return;
    const real e = eterm[k];
    cplx v = g[k];
    v.x *= e; v.y *= e;
    g[k] = v;
}

// Slab-decomposed reciprocal pass (several ranks), see Engine::slabReciprocalPass.
// PACK: [xl][r][e] -> [r][xl][e] (e = one rank's ky rows x kz of a plane, contiguous) so that the block for rank r is
// one contiguous send buffer; !PACK: the inverse, after the transpose back.
template <typename cplx, bool PACK>
__global__ void k_slab_transpose(int nxl, int R, int rowLen, const cplx* __restrict__ src, cplx* __restrict__ dst) {
    const size_t idx = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    const size_t total = (size_t) nxl*R*rowLen;
    if (idx >= total) return;
    const int e = (int) (idx % rowLen);
    const size_t q = idx / rowLen;
    const int r = (int) (q % R), xl = (int) (q / R);          // idx enumerates [xl][r][e]
    const size_t packed = ((size_t) r*nxl + xl)*rowLen + e;   // [r][xl][e]
    if (PACK) dst[packed] = src[idx]; else dst[idx] = src[packed];
}
// influence function on this rank's ky rows: data[x][kyl][kz] *= eterm[x][ky0 + kyl][kz]
// halo reduce of the partitioned reciprocal pass: what the neighbours spread into this rank's block
//   top[0..nLo) += in[0..nLo)   (low halo of the rank above)      bottom[0..nHi) += in[nLo..nLo+nHi)   (high halo of the rank below)
template <typename real>
__global__ void k_halo_add(size_t nLo, size_t nHi, const real* __restrict__ in, real* __restrict__ top, real* __restrict__ bottom) {
    const size_t t = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (t < nLo) top[t] += in[t];
    else if (t < nLo + nHi) bottom[t - nLo] += in[t];
}

template <typename cplx, typename real>
__global__ void k_slab_convolution(int nx, int ny, int nyl, int ky0, int nzc, const real* __restrict__ eterm, cplx* __restrict__ data) {
    const size_t idx = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= (size_t) nx*nyl*nzc) return;
    const int kz = (int) (idx % nzc);
    const size_t q = idx / nzc;
    const int kyl = (int) (q % nyl), x = (int) (q / nyl);
    const real e = eterm[((size_t) x*ny + ky0 + kyl)*nzc + kz];
    cplx v = data[idx];
    v.x *= e; v.y *= e;
    data[idx] = v;
}

// eterm[kx][ky][kz<=nz/2] = exp(-pi^2 m^2/alpha^2) / (pi V m^2 Bx By Bz), zero at the origin
template <typename real>
__global__ void k_eterm_table(DevParams P, const double* __restrict__ modX, const double* __restrict__ modY,
                              const double* __restrict__ modZ, real* __restrict__ eterm) {
    const int nx = P.grid[0], ny = P.grid[1], nz = P.grid[2], nzc = nz/2 + 1;
    const size_t k = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= (size_t) nx*ny*nzc) return;
    const int kz = (int) (k % nzc);
    const int ky = (int) ((k / nzc) % ny);
    const int kx = (int) (k / ((size_t) nzc*ny));
    if (kx == 0 && ky == 0 && kz == 0) { eterm[k] = 0; return; }
    const int mx = (kx < (nx+1)/2) ? kx : kx - nx;
    const int my = (ky < (ny+1)/2) ? ky : ky - ny;
    const int mz = (kz < (nz+1)/2) ? kz : kz - nz;
    const double hx = mx*P.box.ra[0];
    const double hy = mx*P.box.rb[0] + my*P.box.rb[1];
    const double hz = mx*P.box.rc[0] + my*P.box.rc[1] + mz*P.box.rc[2];
    const double m2 = hx*hx + hy*hy + hz*hz;
    const double expFactor = MPID_PI*MPID_PI/(P.alpha*P.alpha);
    const double scaleFactor = 1.0/(MPID_PI*P.box.a[0]*P.box.b[1]*P.box.c[2]);
    eterm[k] = (real) (scaleFactor*exp(-expFactor*m2)/(m2*modX[kx]*modY[ky]*modZ[kz]));
}

// Potential derivatives at the atoms up to total order LEVEL (4 -> all 35), SoA output phi[idx*n + s]; one thread
// per atom contracts its 6x6x6 support z -> y -> x.  POL: the listed polarizable rows with the weights of
// k_spline_weights (record recBase + t, LEVEL <= 2); otherwise weights are built here from the positions.
// (A six-lanes-per-atom variant, one x plane per lane with a shuffle reduction, was measured slower on B200 -- 22 us
// against 16 us for the field-only gather, 79 us against 40 us for all 35 derivatives -- although it exposes six
// times the loads: the kernel is bound by L1 sector traffic, 36 sectors per atom whichever way they are issued.
// Reading each row as the two or three aligned quads that cover it, with z weights shifted to the quad origin, was
// measured too: 1.628 ms per evaluation against 1.546 ms -- 94 instead of 80 registers for the field-only gather.)
//   reference: computeFixedPotentialFromGrid (:3368-3530), computeInducedPotentialFromGrid (:3575-3737)
template <typename real, int LEVEL, bool POL>
__global__ void __launch_bounds__(128)
k_gather(DevParams P, int numRows, const int* __restrict__ rowList, int recBase, const real* __restrict__ theta,
         const int4* __restrict__ igrid, const double4* __restrict__ posS, const real* __restrict__ grid, real* __restrict__ phi) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= numRows) return;
    const int s = rowList ? rowList[t] : P.rowBegin + t;
    constexpr int NV = LEVEL + 1;
    static_assert(!POL || LEVEL <= 2, "the polarizable-site record holds derivatives 0..2");
    int4 ig;
    real tx[6][5], ty[6][5], tz[6][5];
    if (POL) {
        const int rec = recBase + t;
        const real* th = theta + (size_t) rec*MPID_THETA_POL;
        ig = igrid[rec];
        loadTheta<real, NV>(th, 0, tx);
        loadTheta<real, NV>(th, 1, ty);
        loadTheta<real, NV>(th, 2, tz);
    } else {
        const double4 p = posS[s];
        int g[3]; double w[3];
        pmeAtomCell(P.box, P.geom, p.x, p.y, p.z, g, w);
        ig = make_int4(g[0], g[1], g[2], 0);
        bsplineWeights<real>((real) w[0], tx);
        bsplineWeights<real>((real) w[1], ty);
        bsplineWeights<real>((real) w[2], tz);
    }
    const int nx = P.grid[0], ny = P.grid[1], nz = P.grid[2];
    real acc[NV][NV][NV];
#pragma unroll
    for (int a = 0; a < NV; a++)
#pragma unroll
        for (int u = 0; u < NV; u++)
#pragma unroll
            for (int v = 0; v < NV; v++) acc[a][u][v] = 0;
    const bool zWrap = ig.z + 5 >= nz;
#pragma unroll
    for (int ix = 0; ix < 6; ix++) {
        int x = ig.x + ix; x -= (x >= nx) ? nx : 0;
        real yz[NV][NV];
#pragma unroll
        for (int u = 0; u < NV; u++)
#pragma unroll
            for (int v = 0; v < NV; v++) yz[u][v] = 0;
#pragma unroll
        for (int iy = 0; iy < 6; iy++) {
            int y = ig.y + iy; y -= (y >= ny) ? ny : 0;
            const real* row = grid + ((size_t) x*ny + y)*nz;
            real zs[NV];
#pragma unroll
            for (int v = 0; v < NV; v++) zs[v] = 0;
#pragma unroll
            for (int iz = 0; iz < 6; iz++) {
                int z = ig.z + iz; z -= (zWrap && z >= nz) ? nz : 0;
                const real q = row[z];
#pragma unroll
                for (int v = 0; v < NV; v++) zs[v] += q*tz[iz][v];
            }
#pragma unroll
            for (int u = 0; u < NV; u++)
#pragma unroll
                for (int v = 0; v < NV; v++)
                    if (u + v <= LEVEL) yz[u][v] += zs[v]*ty[iy][u];
        }
#pragma unroll
        for (int a = 0; a < NV; a++)
#pragma unroll
            for (int u = 0; u < NV; u++)
#pragma unroll
                for (int v = 0; v < NV; v++)
                    if (a + u + v <= LEVEL) acc[a][u][v] += yz[u][v]*tx[ix][a];
    }
#pragma unroll
    for (int a = 0; a < NV; a++)
#pragma unroll
        for (int u = 0; u < NV; u++)
#pragma unroll
            for (int v = 0; v < NV; v++)
                if (a + u + v <= LEVEL) phi[(size_t) phiIndex(a, u, v)*P.n + s] = acc[a][u][v];
}

// ---------------------------------------------------------------------------------------------------
// Stage 6: per-atom solver arithmetic
// ---------------------------------------------------------------------------------------------------
// E_fixed += reciprocal + self for the rows this rank owns   (:2922-2949)
template <typename real>
__global__ void k_fixed_recip(DevParams P, const real* __restrict__ phi, const double* __restrict__ cartD, double* __restrict__ field) {
    const int s = P.rowBegin + blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.rowEnd) return;
    if (P.method != PME) return;
    double rx, ry, rz;
    reciprocalFieldOf<real>(P, phi, s, rx, ry, rz);
    field[3*(size_t) s]   += rx + P.selfFieldTerm*cartD[20*(size_t) s+1];
    field[3*(size_t) s+1] += ry + P.selfFieldTerm*cartD[20*(size_t) s+2];
    field[3*(size_t) s+2] += rz + P.selfFieldTerm*cartD[20*(size_t) s+3];
}
// efix = alpha.E_fixed ; mu = efix       (:1305-1316, :936-946)
template <typename real>
__global__ void k_fixed_mu(DevParams P, const double* __restrict__ alphaLab, const double* __restrict__ field,
                           double* __restrict__ efix, double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud, int sBegin, int sEnd) {
    const int s = sBegin + blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= sEnd) return;
    double ox, oy, oz;
    applyAlphaLab(alphaLab + 6*(size_t) s, field[3*(size_t) s], field[3*(size_t) s+1], field[3*(size_t) s+2], ox, oy, oz);
    efix[3*(size_t) s] = ox; efix[3*(size_t) s+1] = oy; efix[3*(size_t) s+2] = oz;
    mu[3*(size_t) s] = ox; mu[3*(size_t) s+1] = oy; mu[3*(size_t) s+2] = oz;
    typename Real4<real>::type m = mud[s];
    m.x = (real) ox; m.y = (real) oy; m.z = (real) oz;
    mud[s] = m;
}

// Single-rank fusion of k_fixed_recip and k_fixed_mu (no collective between them): one pass over the atoms.
template <typename real>
__global__ void k_fixed_recip_mu(DevParams P, const real* __restrict__ phi, const double* __restrict__ cartD,
                                 const double* __restrict__ alphaLab, const double* __restrict__ field,
                                 double* __restrict__ efix, double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.n) return;
    double fx = field[3*(size_t) s], fy = field[3*(size_t) s+1], fz = field[3*(size_t) s+2];
    if (P.method == PME) {
        double rx, ry, rz;
        reciprocalFieldOf<real>(P, phi, s, rx, ry, rz);
        fx += rx + P.selfFieldTerm*cartD[20*(size_t) s+1];
        fy += ry + P.selfFieldTerm*cartD[20*(size_t) s+2];
        fz += rz + P.selfFieldTerm*cartD[20*(size_t) s+3];
    }
    double ox, oy, oz;
    applyAlphaLab(alphaLab + 6*(size_t) s, fx, fy, fz, ox, oy, oz);
    efix[3*(size_t) s] = ox; efix[3*(size_t) s+1] = oy; efix[3*(size_t) s+2] = oz;
    mu[3*(size_t) s] = ox; mu[3*(size_t) s+1] = oy; mu[3*(size_t) s+2] = oz;
    typename Real4<real>::type m = mud[s];
    m.x = (real) ox; m.y = (real) oy; m.z = (real) oz;
    mud[s] = m;
}

// Induced field: add the reciprocal part, the self term and (GRAD) the reciprocal field gradient.
//   reference: :4046-4058, :4094-4129, :4133-4140
template <typename real, bool GRAD>
__global__ void k_induced_finish(DevParams P, int numPol, const int* __restrict__ polList, const real* __restrict__ phidp,
                                 const double* __restrict__ mu, double* __restrict__ field, double* __restrict__ grad) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= numPol) return;
    if (P.method != PME) return;
    const int s = polList[t];
    double rx, ry, rz;
    reciprocalFieldOf<real>(P, phidp, s, rx, ry, rz);
    field[3*(size_t) s]   += rx + P.selfFieldTerm*mu[3*(size_t) s];
    field[3*(size_t) s+1] += ry + P.selfFieldTerm*mu[3*(size_t) s+1];
    field[3*(size_t) s+2] += rz + P.selfFieldTerm*mu[3*(size_t) s+2];
    if (GRAD) {
        const size_t n = P.n;
        const double pxx = phidp[4*n + s], pyy = phidp[5*n + s], pzz = phidp[6*n + s];
        const double pxy = phidp[7*n + s], pxz = phidp[8*n + s], pyz = phidp[9*n + s];
        const double E[3][3] = {{pxx, pxy, pxz}, {pxy, pyy, pyz}, {pxz, pyz, pzz}};
        const int gi[6] = {0, 1, 2, 0, 0, 1}, gj[6] = {0, 1, 2, 1, 2, 2};
        for (int c = 0; c < 6; c++) {
            double sum = 0;
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) sum += P.geom.A[k][gi[c]]*E[k][l]*P.geom.A[l][gj[c]];
            grad[6*(size_t) s + c] -= sum;
        }
    }
}

// DIIS bookkeeping (:1195-1218): newDip = efix + alpha.E_ind, err = newDip - mu, both into history slot.
__global__ void k_diis_record(DevParams P, const double* __restrict__ alphaLab, const double* __restrict__ efix,
                              const double* __restrict__ ifield, const double* __restrict__ mu,
                              double* __restrict__ histDip, double* __restrict__ histErr) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.n) return;
    double ox, oy, oz;
    applyAlphaLab(alphaLab + 6*(size_t) s, ifield[3*(size_t) s], ifield[3*(size_t) s+1], ifield[3*(size_t) s+2], ox, oy, oz);
    const double nx = efix[3*(size_t) s] + ox, ny = efix[3*(size_t) s+1] + oy, nz = efix[3*(size_t) s+2] + oz;
    histDip[3*(size_t) s] = nx; histDip[3*(size_t) s+1] = ny; histDip[3*(size_t) s+2] = nz;
    histErr[3*(size_t) s] = nx - mu[3*(size_t) s]; histErr[3*(size_t) s+1] = ny - mu[3*(size_t) s+1]; histErr[3*(size_t) s+2] = nz - mu[3*(size_t) s+2];
}

// dots[k] = <vec, hist_k> for k < m : fixed block partition + ordered second pass => deterministic
struct VecList { const double* v[MPID_MAX_HISTORY + 1]; };
__global__ void __launch_bounds__(256)
k_dots_partial(size_t len, int m, const double* __restrict__ vec, VecList hist, double* __restrict__ partial) {
    __shared__ double sh[256/32][MPID_MAX_HISTORY + 1];
    double acc[MPID_MAX_HISTORY + 1];
    for (int k = 0; k < m; k++) acc[k] = 0;
    for (size_t idx = (size_t) blockIdx.x*blockDim.x + threadIdx.x; idx < len; idx += (size_t) gridDim.x*blockDim.x) {
        const double a = vec[idx];
        for (int k = 0; k < m; k++) acc[k] += a*hist.v[k][idx];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < m; k++) {
        double v = acc[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sh[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < m) {
        double v = 0;
        for (int w = 0; w < 256/32; w++) v += sh[w][threadIdx.x];
        partial[(size_t) blockIdx.x*(MPID_MAX_HISTORY + 1) + threadIdx.x] = v;
    }
}
__global__ void __launch_bounds__(128)
k_dots_final(int numBlocks, int m, const double* __restrict__ partial, double* __restrict__ out) {
    // block k reduces the per-block partials of dot product k with a fixed-shape tree (deterministic)
    __shared__ double sh[128];
    const int k = blockIdx.x;
    double v = 0;
    for (int b = threadIdx.x; b < numBlocks; b += 128) v += partial[(size_t) b*(MPID_MAX_HISTORY + 1) + k];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 64; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0 && k < m) out[k] = sh[0];
}

// ---- device-resident DIIS ---------------------------------------------------------------------------
// The mutual solver runs without host round trips: the error-overlap matrix, the convergence test and the
// small linear solve live on the device, and every kernel of an iteration returns at once when `done` is set,
// so the host may enqueue several iterations ahead (it predicts the count from the previous evaluation) and only
// then read the status back.  Extra iterations are no-ops on mu, hence idempotent.
struct DiisStatus {
    int done;                     // eps < target reached
    unsigned ticket;              // CTAs of the fused step kernel that have finished (last one solves)
    int iter;                     // device-side iteration counter (graph-replayed solver steps read it instead of a launch argument)
    int iterations;               // index of the iteration that converged / last one run (:1219-1231)
    double eps;
    double coef[MPID_MAX_HISTORY + 1];
    double B[MPID_MAX_HISTORY*MPID_MAX_HISTORY];     // <err_a, err_b>, indexed by history slot
};
struct SlotList { int s[MPID_MAX_HISTORY + 1]; };

// newDip = efix + alpha.E_ind, err = newDip - mu into history slot (:1195-1218), fused with the partial dot
// products <err_new, err_k> over the m vectors of the history (the last one being err_new itself).
// dst[3 idx + c] = src[3 siteList[idx] + c]: the polarizable entries of a per-atom vector, contiguous -- what the
// per-iteration all-reduce of the partial induced field moves (a third of the atoms in water).
// Covalently scaled pairs this rank evaluates in the energy stage: those whose lower atom is one of its rows (their
// moments are then among the ones it builds).  Order is whatever the atomics give: every consumer accumulates in fixed point.
__global__ void k_own_special(int numSpecial, const int* __restrict__ spPairLo, const int* __restrict__ inv, int rowBegin, int rowEnd,
                              int* __restrict__ ownList, unsigned* __restrict__ count) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= numSpecial) return;
    const int s = inv[spPairLo[k]];
    if (s >= rowBegin && s < rowEnd) ownList[atomicAdd(count, 1u)] = k;
}

// compact dipoles of ALL polarizable sites (order of siteList) -> per-atom mu and the float copy the field kernels read
template <typename real>
__global__ void k_unpack_mu(int numSites, const int* __restrict__ siteList, const double* __restrict__ compact,
                            double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud) {
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= numSites) return;
    const int s = siteList[idx];
    const double x = compact[3*(size_t) idx], y = compact[3*(size_t) idx+1], z = compact[3*(size_t) idx+2];
    mu[3*(size_t) s] = x; mu[3*(size_t) s+1] = y; mu[3*(size_t) s+2] = z;
    typename Real4<real>::type m = mud[s];
    m.x = (real) x; m.y = (real) y; m.z = (real) z;
    mud[s] = m;
}

// column sums of the per-block partial error overlaps, ready for the all-reduce: warp k sums column k (lanes stride
// over the blocks, then a shuffle tree -- a fixed order, so deterministic).  Launch with 32*m threads.
// Peer-to-peer mode (peers.p[0] != nullptr): instead of one local vector for an NCCL all-reduce, every rank writes its sums
// into row `rank` of an [R][MPID_MAX_HISTORY+1] table in EVERY rank's memory (remote stores); after the barrier each rank
// adds the R rows itself (k_diis_solve with numBlocks = R), in the same order everywhere.
struct PeerPtrs { void* p[16]; };      // one device pointer per rank (peer mappings of the same buffer; own entry = local)
__global__ void k_sum_partials(int numBlocks, int m, const double* __restrict__ partial, double* __restrict__ out, int numRanks, int rank, PeerPtrs peerTables, int usePeers) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (k >= m) return;
    double v = 0;
    for (int b = lane; b < numBlocks; b += 32) v += partial[(size_t) b*(MPID_MAX_HISTORY + 1) + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (usePeers) {
        // lanes 0..R-1 each serve one destination rank
        if (lane < numRanks) reinterpret_cast<double*>(peerTables.p[lane])[(size_t) rank*(MPID_MAX_HISTORY + 1) + k] = v;
    } else if (lane == 0) out[k] = v;
}
// new dipoles of this rank's polarizable sites -> the compact dipole array of EVERY rank (remote stores), own piece
// [3 polBegin, 3 (polBegin + numPol)): the dipole exchange of the owner-computes solver without a collective call
__global__ void k_push_dipoles(int numPol, int polBegin, const int* __restrict__ siteList, const double* __restrict__ mu, int numRanks, PeerPtrs peerCompact) {
    const size_t t = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= 3*(size_t) numPol) return;
    const size_t idx = t/3;
    const int c = (int) (t - 3*idx);
    const double v = mu[3*(size_t) siteList[idx] + c];
    for (int r = 0; r < numRanks; r++) reinterpret_cast<double*>(peerCompact.p[r])[3*(size_t) polBegin + t] = v;
}

// ---- peer-to-peer pieces of the partitioned reciprocal pass (all ranks on one NVLink/NVSwitch node) ------------------
// The transform kernels write their results straight into the receive buffers of the other ranks (remote stores through
// peer mappings of their memory), so the all-to-all happens tile by tile inside the kernel that produces the data; what
// is left of the collective is this barrier: every rank tells every other that its stores are out, and waits to hear the
// same from all of them.  flags[r] (in this rank's memory) is written by rank r with a monotonically increasing epoch.
// Halo planes pushed into a neighbour's memory (remote float4 stores over NVLink): two contiguous plane ranges, each to
// its own destination.  Used for the halo reduce (into the neighbours' staging buffers) and the halo gather (into the halo
// regions of the neighbours' grids) of the partitioned reciprocal pass.
__global__ void k_halo_push(size_t nA4, size_t nB4, const float4* __restrict__ srcA, float4* __restrict__ dstA,
                            const float4* __restrict__ srcB, float4* __restrict__ dstB) {
    const size_t t = (size_t) blockIdx.x*blockDim.x + threadIdx.x;
    if (t < nA4) dstA[t] = srcA[t];
    else if (t < nA4 + nB4) dstB[t - nA4] = srcB[t - nA4];
}
__global__ void k_cross_barrier(int numRanks, int rank, int epoch, PeerPtrs peerFlags, volatile int* __restrict__ myFlags, int* __restrict__ timedOut) {
    const int r = threadIdx.x;
    if (r >= numRanks) return;
    __threadfence_system();                                   // this GPU's earlier stores (previous kernels) before the flag
    int* remote = reinterpret_cast<int*>(peerFlags.p[r]) + rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(remote), "r"(epoch) : "memory");
    long long spins = 0;
    for (;;) {
        int seen;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(myFlags + r) : "memory");
        if (seen - epoch >= 0) break;
        if (++spins > 400000000LL) { *timedOut = 1; break; }   // a rank never arrived: report instead of hanging the GPU
        __nanosleep(100);
    }
}

__global__ void k_pack_sites(int numSites, const int* __restrict__ siteList, const double* __restrict__ src, double* __restrict__ dst) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= 3*numSites) return;
    const int idx = t/3, c = t - 3*idx;
    dst[t] = src[3*(size_t) siteList[idx] + c];
}

// Fixed block partition + ordered second pass => deterministic.
__global__ void __launch_bounds__(256)
k_diis_record_dots(DevParams P, const double* __restrict__ alphaLab, const double* __restrict__ efix,
                   const double* __restrict__ ifield, const double* __restrict__ mu,
                   double* __restrict__ histDip, double* __restrict__ histErr, int m, VecList errs,
                   const DiisStatus* __restrict__ status, double* __restrict__ partial,
                   int numSites, const int* __restrict__ siteList, int fieldIsCompact) {
    // siteList: the polarizable sites (every other entry of the solver vectors is and stays zero), nullptr = all atoms;
    // fieldIsCompact: ifield is indexed by the position in siteList (k_pack_sites) instead of by atom
    if (status->done) return;
    __shared__ double sh[256/32][MPID_MAX_HISTORY + 1];
    double acc[MPID_MAX_HISTORY + 1];
    for (int k = 0; k < m; k++) acc[k] = 0;
    for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < numSites; idx += gridDim.x*blockDim.x) {
        const int s = siteList ? siteList[idx] : idx;
        const size_t fi = fieldIsCompact ? (size_t) idx : (size_t) s;
        double ox, oy, oz;
        applyAlphaLab(alphaLab + 6*(size_t) s, ifield[3*fi], ifield[3*fi+1], ifield[3*fi+2], ox, oy, oz);
        const double nx = efix[3*(size_t) s] + ox, ny = efix[3*(size_t) s+1] + oy, nz = efix[3*(size_t) s+2] + oz;
        histDip[3*(size_t) s] = nx; histDip[3*(size_t) s+1] = ny; histDip[3*(size_t) s+2] = nz;
        const double e0 = nx - mu[3*(size_t) s], e1 = ny - mu[3*(size_t) s+1], e2 = nz - mu[3*(size_t) s+2];
        histErr[3*(size_t) s] = e0; histErr[3*(size_t) s+1] = e1; histErr[3*(size_t) s+2] = e2;
        for (int k = 0; k < m - 1; k++) {
            const double* h = errs.v[k] + 3*(size_t) s;
            acc[k] += e0*h[0] + e1*h[1] + e2*h[2];
        }
        acc[m-1] += e0*e0 + e1*e1 + e2*e2;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < m; k++) {
        double v = acc[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sh[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < m) {
        double v = 0;
        for (int w = 0; w < 256/32; w++) v += sh[w][threadIdx.x];
        partial[(size_t) blockIdx.x*(MPID_MAX_HISTORY + 1) + threadIdx.x] = v;
    }
}

// One CTA: finish the dot products, update B, test convergence (eps = 48.033324 sqrt(<e,e>/N), :1219) and, if not
// converged, solve the (m+1)x(m+1) DIIS system (:1254-1291 does it through an SVD) by Gauss-Jordan elimination
// with partial pivoting, one thread per matrix element.
__device__ __forceinline__ void diisSolveBlock(int numBlocks, int m, const SlotList& slots, int iteration, int numAtoms, double targetEps,
                                               const double* __restrict__ partial, DiisStatus* __restrict__ status) {
    // requires blockDim.x == 512
    constexpr int H = MPID_MAX_HISTORY, R = MPID_MAX_HISTORY + 1, W = MPID_MAX_HISTORY + 2;
    __shared__ double red[512];
    __shared__ double dotv[R];
    __shared__ double a[R][W];
    __shared__ int pivRow;
    __shared__ int converged;
    const int tid = threadIdx.x;
    // dot product k is finished by the 20 threads of group k = tid/20, in a fixed order
    {
        const int k = tid/20, q = tid % 20;
        double v = 0;
        if (k < m) for (int b = q; b < numBlocks; b += 20) v += partial[(size_t) b*R + k];
        red[tid] = v;
        __syncthreads();
        if (k < m && q == 0) {
            double t = 0;
            for (int u = 0; u < 20; u++) t += red[tid + u];
            dotv[k] = t;
        }
        __syncthreads();
    }
    const int newSlot = slots.s[m-1];
    if (tid < m) {
        const int sk = slots.s[tid];
        status->B[newSlot*H + sk] = dotv[tid];
        status->B[sk*H + newSlot] = dotv[tid];
    }
    if (tid == 0) {
        const double eps = MPID_DEBYE*sqrt(dotv[m-1]/numAtoms);
        status->iterations = iteration;
        status->eps = eps;
        converged = eps < targetEps ? 1 : 0;
        if (converged) status->done = 1;
    }
    __syncthreads();
    if (converged) return;
    if (m == 1) { if (tid == 0) status->coef[0] = 1.0; return; }
    __threadfence_block();
    // The error overlaps span 20+ orders of magnitude near convergence, so the overlap block is scaled to a unit diagonal
    // (B'_ij = B_ij d_i d_j, d_i = 1/sqrt(B_ii); the border -1 becomes -d_i) before the elimination, and pivots below 1e-12
    // are dropped (their coefficient is 0): the rank truncation the reference gets from its SVD (:1276-1289).
    // tests/host_emul/emul.cpp carries the same algorithm and is checked against the oracle down to eps = 1e-12.
    const int rank = m + 1, w = rank + 1;
    const int r = tid / W, c = tid % W;
    const bool inside = r < rank && c < w && tid < R*W;
    __shared__ double dscale[R];
    __shared__ int dropped[R];
    if (tid < rank) {
        dropped[tid] = 0;
        if (tid == 0) dscale[0] = 1.0;
        else {
            const double bii = status->B[slots.s[tid-1]*H + slots.s[tid-1]];
            dscale[tid] = bii > 0.0 ? rsqrt(bii) : 1.0;
        }
    }
    __syncthreads();
    if (inside) {
        double v;
        if (c == rank) v = r == 0 ? -1.0 : 0.0;
        else if (r == 0 && c == 0) v = 0.0;
        else if (r == 0) v = -dscale[c];
        else if (c == 0) v = -dscale[r];
        else v = status->B[slots.s[r-1]*H + slots.s[c-1]]*dscale[r]*dscale[c];
        a[r][c] = v;
    }
    __syncthreads();
    for (int col = 0; col < rank; col++) {
        if (tid == 0) {
            int piv = col;
            for (int q = col + 1; q < rank; q++) if (fabs(a[q][col]) > fabs(a[piv][col])) piv = q;
            pivRow = piv;
        }
        __syncthreads();
        const int piv = pivRow;
        // swap rows col <-> piv through registers
        const bool sw = piv != col && inside && (r == col || r == piv);
        double other = 0.0;
        if (sw) other = a[r == col ? piv : col][c];
        __syncthreads();
        if (sw) a[r][c] = other;
        __syncthreads();
        const double d = a[col][col];
        const bool usable = fabs(d) >= 1e-12;
        if (!usable && tid == 0) dropped[col] = 1;
        double f = 0.0, pc = 0.0;
        if (inside && r != col && usable) { f = a[r][col]/d; pc = a[col][c]; }
        __syncthreads();
        if (inside && r != col && usable && f != 0.0 && c >= col) a[r][c] -= f*pc;
        __syncthreads();
    }
    if (tid < m) {
        const double d = a[tid+1][tid+1];
        status->coef[tid] = (!dropped[tid+1] && d != 0.0) ? dscale[tid+1]*a[tid+1][rank]/d : 0.0;
    }
}

__global__ void __launch_bounds__(512)
k_diis_solve(int numBlocks, int m, SlotList slots, int iteration, int numAtoms, double targetEps,
             const double* __restrict__ partial, DiisStatus* __restrict__ status) {
    if (status->done) return;
    diisSolveBlock(numBlocks, m, slots, iteration, numAtoms, targetEps, partial, status);
}

// Single-GPU fusion of one solver step: (1) finish the induced field at polarizable sites (reciprocal part from
// phidp + self term, what k_induced_finish does), (2) newDip / err / history and the partial error overlaps (what
// k_diis_record_dots does), (3) the CTA that finishes last reduces the partials and runs the convergence test and
// the DIIS solve (k_diis_solve).  One launch instead of three on the solver's critical path.
// History bookkeeping as a function of the iteration index (what the host loop of the reference does with its
// vectors, :1232-1236): iteration `it` writes ring slot it % H; the m = min(it+1, H) live vectors in age order are
// the slots (it-m+1 .. it) % H.
__device__ __forceinline__ int diisHistory(int it, SlotList& sl) {
    const int H = MPID_MAX_HISTORY;
    const int m = min(it + 1, H);
    for (int a = 0; a < m; a++) sl.s[a] = (it - (m - 1) + a) % H;
    return m;
}

// (Folding the field-only k_gather<1> of the iteration into this kernel was measured: 1.577 ms per evaluation against
// 1.546 ms with the separate launch -- the 216 grid reads per site at 128 registers cost more than the launch saves.)
template <typename real>
__global__ void __launch_bounds__(512)
k_diis_step(DevParams P, const int* __restrict__ flagS, const real* __restrict__ phidp,
            const double* __restrict__ alphaLab, const double* __restrict__ efix,
            const double* __restrict__ ifield, const double* __restrict__ mu,
            double* __restrict__ histDipBase, double* __restrict__ histErrBase,
            int itHost, double targetEps, DiisStatus* __restrict__ status, double* __restrict__ partial,
            int numSites, const int* __restrict__ siteList) {
    // siteList: the polarizable sites (every other entry of the solver vectors is and stays zero); nullptr = all atoms
    if (status->done) return;
    __shared__ double sh[512/32][MPID_MAX_HISTORY + 1];
    __shared__ int isLast;
    // itHost < 0: the launch is a replayed graph node and the iteration index lives in the status block
    const int it = itHost >= 0 ? itHost : status->iter;
    SlotList slots;
    const int m = diisHistory(it, slots);
    const size_t vlen = 3*(size_t) P.n;
    double* histDip = histDipBase + (size_t) slots.s[m-1]*vlen;
    double* histErr = histErrBase + (size_t) slots.s[m-1]*vlen;
    double acc[MPID_MAX_HISTORY + 1];
    for (int k = 0; k < m; k++) acc[k] = 0;
    const bool pme = P.method == PME;
    for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < numSites; idx += gridDim.x*blockDim.x) {
        const int s = siteList ? siteList[idx] : idx;
        double fx = ifield[3*(size_t) s], fy = ifield[3*(size_t) s+1], fz = ifield[3*(size_t) s+2];
        const double ux = mu[3*(size_t) s], uy = mu[3*(size_t) s+1], uz = mu[3*(size_t) s+2];
        if (pme && (flagS[s] & 1)) {
            double rx, ry, rz;
            reciprocalFieldOf<real>(P, phidp, s, rx, ry, rz);
            fx += rx + P.selfFieldTerm*ux; fy += ry + P.selfFieldTerm*uy; fz += rz + P.selfFieldTerm*uz;
        }
        double ox, oy, oz;
        applyAlphaLab(alphaLab + 6*(size_t) s, fx, fy, fz, ox, oy, oz);
        const double nx = efix[3*(size_t) s] + ox, ny = efix[3*(size_t) s+1] + oy, nz = efix[3*(size_t) s+2] + oz;
        histDip[3*(size_t) s] = nx; histDip[3*(size_t) s+1] = ny; histDip[3*(size_t) s+2] = nz;
        const double e0 = nx - ux, e1 = ny - uy, e2 = nz - uz;
        histErr[3*(size_t) s] = e0; histErr[3*(size_t) s+1] = e1; histErr[3*(size_t) s+2] = e2;
        for (int k = 0; k < m - 1; k++) {
            const double* h = histErrBase + (size_t) slots.s[k]*vlen + 3*(size_t) s;
            acc[k] += e0*h[0] + e1*h[1] + e2*h[2];
        }
        acc[m-1] += e0*e0 + e1*e1 + e2*e2;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < m; k++) {
        double v = acc[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sh[wid][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < m) {
        double v = 0;
        for (int w = 0; w < 512/32; w++) v += sh[w][threadIdx.x];
        partial[(size_t) blockIdx.x*(MPID_MAX_HISTORY + 1) + threadIdx.x] = v;
    }
    // last CTA done: every CTA publishes its partials, then takes a ticket
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&status->ticket, 1u);
        isLast = t == gridDim.x - 1;
        if (isLast) status->ticket = 0u;
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    diisSolveBlock((int) gridDim.x, m, slots, it, P.n, targetEps, partial, status);
    __syncthreads();
    if (threadIdx.x == 0) status->iter = it + 1;
}

// mu = sum_k coef[k] * histDip_k with the coefficients the solve left in the status block (:1240-1249).
// itHost < 0: replayed graph node, the iteration just solved is status->iter - 1.
template <typename real>
__global__ void k_diis_combine_ring(int n, int itHost, const double* __restrict__ histDipBase, const DiisStatus* __restrict__ status,
                                    double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud,
                                    int numSites, const int* __restrict__ siteList) {
    if (status->done) return;
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= numSites) return;
    const int s = siteList ? siteList[idx] : idx;
    const int it = itHost >= 0 ? itHost : status->iter - 1;
    SlotList slots;
    const int m = diisHistory(it, slots);
    const size_t vlen = 3*(size_t) n;
    double x = 0, y = 0, z = 0;
    for (int k = 0; k < m; k++) {
        const double c = status->coef[k];
        const double* v = histDipBase + (size_t) slots.s[k]*vlen + 3*(size_t) s;
        x += c*v[0]; y += c*v[1]; z += c*v[2];
    }
    mu[3*(size_t) s] = x; mu[3*(size_t) s+1] = y; mu[3*(size_t) s+2] = z;
    typename Real4<real>::type v = mud[s];
    v.x = (real) x; v.y = (real) y; v.z = (real) z;
    mud[s] = v;
}

// mu = sum_k coef[k] * vec_k   (DIIS extrapolation :1240-1249, OPT combination :1172-1177); repacks mud
struct CoefList { double c[MPID_MAX_HISTORY + 1]; };
template <typename real>
__global__ void k_combine(int n, int m, VecList vecs, CoefList coef, double* __restrict__ mu,
                          typename Real4<real>::type* __restrict__ mud) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= n) return;
    double x = 0, y = 0, z = 0;
    for (int k = 0; k < m; k++) {
        const double c = coef.c[k];
        x += c*vecs.v[k][3*(size_t) s]; y += c*vecs.v[k][3*(size_t) s+1]; z += c*vecs.v[k][3*(size_t) s+2];
    }
    mu[3*(size_t) s] = x; mu[3*(size_t) s+1] = y; mu[3*(size_t) s+2] = z;
    typename Real4<real>::type v = mud[s];
    v.x = (real) x; v.y = (real) y; v.z = (real) z;
    mud[s] = v;
}

// ---- preconditioned conjugate gradient (the alternative mutual solver, MPIDB200_SOLVER_CG) -----------------
// Solves A mu = E with A = alpha^-1 - T (symmetric positive definite on the polarizable sites) and the diagonal-block
// preconditioner M^-1 = alpha_lab, starting from the reference's own first guess mu0 = alpha E (:936-946).  One field
// pass per iteration gives T p.  alpha^-1 p never needs an inverse: p = z + beta p with z = alpha r, so
// w := alpha^-1 p obeys w = r + beta w.  Convergence is tested on the same quantity as the reference's DIIS loop,
// eps = 48.033324 sqrt(sum |mu_new - mu|^2 / N) (:1219): for a Jacobi update mu_new - mu = alpha (E - A mu) = z.
// The scalars (gamma, beta, r.z) live in the status block; all kernels return at once when `done` is set.
struct CgStatus {
    int done;
    unsigned ticket;
    int iter;                     // field evaluations used so far - 1
    int iterations;
    double eps;
    double rz;                    // r.z of the current residual
    double gamma, beta;
};

// block-level sum of two doubles per thread, then last-CTA-done reduction over the grid; returns true in the last CTA
// with the totals in out[0..1] (block size 512)
__device__ __forceinline__ bool cgGridReduce2(double a, double b, double* __restrict__ partial, unsigned* ticket, double* out) {
    __shared__ double sh2[512/32][2];
    __shared__ int last2;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
    if (lane == 0) { sh2[wid][0] = a; sh2[wid][1] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0, tb = 0;
        for (int w = 0; w < 512/32; w++) { ta += sh2[w][0]; tb += sh2[w][1]; }
        partial[2*blockIdx.x] = ta; partial[2*blockIdx.x + 1] = tb;
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        last2 = t == gridDim.x - 1;
        if (last2) *ticket = 0u;
    }
    __syncthreads();
    if (!last2) return false;
    __threadfence();
    if (threadIdx.x == 0) {
        double ta = 0, tb = 0;
        for (unsigned b2 = 0; b2 < gridDim.x; b2++) { ta += partial[2*b2]; tb += partial[2*b2 + 1]; }     // fixed order: deterministic
        out[0] = ta; out[1] = tb;
    }
    __syncthreads();
    return true;
}

// Start: r0 = T mu0 (the induced field of the first guess, reciprocal + self part added here on a single GPU),
// z0 = alpha r0, p0 = z0 (into mu / mud, the input of the next field pass), w0 = r0, muAcc = mu0; eps0, r.z.
template <typename real, bool FINISH>
__global__ void __launch_bounds__(512)
k_cg_init(DevParams P, const int* __restrict__ flagS, const real* __restrict__ phidp, const double* __restrict__ alphaLab,
          const double* __restrict__ ifield, double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud,
          double* __restrict__ r, double* __restrict__ w, double* __restrict__ muAcc, double targetEps,
          CgStatus* __restrict__ status, double* __restrict__ partial) {
    double rz = 0, zz = 0;
    for (int s = blockIdx.x*blockDim.x + threadIdx.x; s < P.n; s += gridDim.x*blockDim.x) {
        double fx = ifield[3*(size_t) s], fy = ifield[3*(size_t) s+1], fz = ifield[3*(size_t) s+2];
        const double ux = mu[3*(size_t) s], uy = mu[3*(size_t) s+1], uz = mu[3*(size_t) s+2];
        const bool pol = (flagS[s] & 1) != 0;
        if (FINISH && P.method == PME && pol) {
            double rx, ry, rzc;
            reciprocalFieldOf<real>(P, phidp, s, rx, ry, rzc);
            fx += rx + P.selfFieldTerm*ux; fy += ry + P.selfFieldTerm*uy; fz += rzc + P.selfFieldTerm*uz;
        }
        if (!pol) { fx = fy = fz = 0; }
        double zx, zy, zzc;
        applyAlphaLab(alphaLab + 6*(size_t) s, fx, fy, fz, zx, zy, zzc);
        muAcc[3*(size_t) s] = ux; muAcc[3*(size_t) s+1] = uy; muAcc[3*(size_t) s+2] = uz;
        r[3*(size_t) s] = fx; r[3*(size_t) s+1] = fy; r[3*(size_t) s+2] = fz;
        w[3*(size_t) s] = fx; w[3*(size_t) s+1] = fy; w[3*(size_t) s+2] = fz;
        mu[3*(size_t) s] = zx; mu[3*(size_t) s+1] = zy; mu[3*(size_t) s+2] = zzc;
        typename Real4<real>::type v = mud[s];
        v.x = (real) zx; v.y = (real) zy; v.z = (real) zzc;
        mud[s] = v;
        rz += fx*zx + fy*zy + fz*zzc;
        zz += zx*zx + zy*zy + zzc*zzc;
    }
    double tot[2];
    if (!cgGridReduce2(rz, zz, partial, &status->ticket, tot)) return;
    if (threadIdx.x == 0) {
        status->rz = tot[0];
        status->eps = MPID_DEBYE*sqrt(tot[1]/P.n);
        status->iterations = 0; status->iter = 0;
        if (status->eps < targetEps) status->done = 1;
    }
}

// Ap = w - T p ; p.Ap ; gamma = r.z / p.Ap
template <typename real, bool FINISH>
__global__ void __launch_bounds__(512)
k_cg_ap(DevParams P, const int* __restrict__ flagS, const real* __restrict__ phidp, const double* __restrict__ ifield,
        const double* __restrict__ p, const double* __restrict__ w, double* __restrict__ ap,
        CgStatus* __restrict__ status, double* __restrict__ partial) {
    if (status->done) return;
    double pap = 0;
    for (int s = blockIdx.x*blockDim.x + threadIdx.x; s < P.n; s += gridDim.x*blockDim.x) {
        double fx = ifield[3*(size_t) s], fy = ifield[3*(size_t) s+1], fz = ifield[3*(size_t) s+2];
        const double px = p[3*(size_t) s], py = p[3*(size_t) s+1], pz = p[3*(size_t) s+2];
        const bool pol = (flagS[s] & 1) != 0;
        if (FINISH && P.method == PME && pol) {
            double rx, ry, rzc;
            reciprocalFieldOf<real>(P, phidp, s, rx, ry, rzc);
            fx += rx + P.selfFieldTerm*px; fy += ry + P.selfFieldTerm*py; fz += rzc + P.selfFieldTerm*pz;
        }
        double ax = 0, ay = 0, az = 0;
        if (pol) { ax = w[3*(size_t) s] - fx; ay = w[3*(size_t) s+1] - fy; az = w[3*(size_t) s+2] - fz; }
        ap[3*(size_t) s] = ax; ap[3*(size_t) s+1] = ay; ap[3*(size_t) s+2] = az;
        pap += px*ax + py*ay + pz*az;
    }
    double tot[2];
    if (!cgGridReduce2(pap, 0.0, partial, &status->ticket, tot)) return;
    if (threadIdx.x == 0) status->gamma = tot[0] != 0.0 ? status->rz/tot[0] : 0.0;
}

// muAcc += gamma p ; r -= gamma Ap ; z = alpha r (kept in `z`) ; new r.z and eps ; beta
__global__ void __launch_bounds__(512)
k_cg_update(DevParams P, const double* __restrict__ alphaLab, const double* __restrict__ p, const double* __restrict__ ap,
            double* __restrict__ muAcc, double* __restrict__ r, double* __restrict__ z, double targetEps,
            CgStatus* __restrict__ status, double* __restrict__ partial) {
    if (status->done) return;
    const double gamma = status->gamma;
    double rz = 0, zz = 0;
    for (int s = blockIdx.x*blockDim.x + threadIdx.x; s < P.n; s += gridDim.x*blockDim.x) {
        double rx = r[3*(size_t) s], ry = r[3*(size_t) s+1], rzc = r[3*(size_t) s+2];
        for (int k = 0; k < 3; k++) muAcc[3*(size_t) s + k] += gamma*p[3*(size_t) s + k];
        rx -= gamma*ap[3*(size_t) s]; ry -= gamma*ap[3*(size_t) s+1]; rzc -= gamma*ap[3*(size_t) s+2];
        r[3*(size_t) s] = rx; r[3*(size_t) s+1] = ry; r[3*(size_t) s+2] = rzc;
        double zx, zy, zzc;
        applyAlphaLab(alphaLab + 6*(size_t) s, rx, ry, rzc, zx, zy, zzc);
        z[3*(size_t) s] = zx; z[3*(size_t) s+1] = zy; z[3*(size_t) s+2] = zzc;
        rz += rx*zx + ry*zy + rzc*zzc;
        zz += zx*zx + zy*zy + zzc*zzc;
    }
    double tot[2];
    if (!cgGridReduce2(rz, zz, partial, &status->ticket, tot)) return;
    if (threadIdx.x == 0) {
        status->beta = status->rz != 0.0 ? tot[0]/status->rz : 0.0;
        status->rz = tot[0];
        status->eps = MPID_DEBYE*sqrt(tot[1]/P.n);
        status->iter += 1;
        status->iterations = status->iter;
        if (status->eps < targetEps) status->done = 1;
    }
}

// p = z + beta p (into mu / mud for the next field pass) ; w = r + beta w.   When converged: mu = muAcc instead.
template <typename real>
__global__ void k_cg_direction(int n, const double* __restrict__ z, const double* __restrict__ r, const double* __restrict__ muAcc,
                               double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud, double* __restrict__ w,
                               const CgStatus* __restrict__ status, int finalOnly) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= n) return;
    double x, y, zc;
    if (status->done) {
        x = muAcc[3*(size_t) s]; y = muAcc[3*(size_t) s+1]; zc = muAcc[3*(size_t) s+2];
    } else {
        if (finalOnly) return;
        const double beta = status->beta;
        x = z[3*(size_t) s] + beta*mu[3*(size_t) s]; y = z[3*(size_t) s+1] + beta*mu[3*(size_t) s+1]; zc = z[3*(size_t) s+2] + beta*mu[3*(size_t) s+2];
        for (int k = 0; k < 3; k++) w[3*(size_t) s + k] = r[3*(size_t) s + k] + beta*w[3*(size_t) s + k];
    }
    mu[3*(size_t) s] = x; mu[3*(size_t) s+1] = y; mu[3*(size_t) s+2] = zc;
    typename Real4<real>::type v = mud[s];
    v.x = (real) x; v.y = (real) y; v.z = (real) zc;
    mud[s] = v;
}

// OPT recursion step (:1149-1167): mu = alpha.E_ind ; store mu, field (gradient is stored by the caller's buffer)
template <typename real>
__global__ void k_opt_step(DevParams P, const double* __restrict__ alphaLab, const double* __restrict__ ifield,
                           double* __restrict__ mu, typename Real4<real>::type* __restrict__ mud,
                           double* __restrict__ ptDip, double* __restrict__ ptField) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.n) return;
    double ox, oy, oz;
    const double fx = ifield[3*(size_t) s], fy = ifield[3*(size_t) s+1], fz = ifield[3*(size_t) s+2];
    applyAlphaLab(alphaLab + 6*(size_t) s, fx, fy, fz, ox, oy, oz);
    mu[3*(size_t) s] = ox; mu[3*(size_t) s+1] = oy; mu[3*(size_t) s+2] = oz;
    ptDip[3*(size_t) s] = ox; ptDip[3*(size_t) s+1] = oy; ptDip[3*(size_t) s+2] = oz;
    ptField[3*(size_t) s] = fx; ptField[3*(size_t) s+1] = fy; ptField[3*(size_t) s+2] = fz;
    typename Real4<real>::type v = mud[s];
    v.x = (real) ox; v.y = (real) oy; v.z = (real) oz;
    mud[s] = v;
}

// ---------------------------------------------------------------------------------------------------
// Stage 7: reciprocal-space energy/force/torque, self terms, OPT response, torque mapping, output
// ---------------------------------------------------------------------------------------------------
//   reference: :3739-3866, :3871-4024, :4283-4333
template <typename real>
__global__ void __launch_bounds__(128)
k_reciprocal_terms(DevParams P, const real* __restrict__ phi, const real* __restrict__ phidp, const double* __restrict__ cartD,
                   const double* __restrict__ sphD, const double* __restrict__ mu, const int* __restrict__ aniso,
                   unsigned long long* __restrict__ force, unsigned long long* __restrict__ torque, unsigned long long* __restrict__ energy) {
    const int s = P.rowBegin + blockIdx.x*blockDim.x + threadIdx.x;
    double e = 0;
    if (s < P.rowEnd) {
        const size_t n = P.n;
        const double ke = MPID_ELECTRIC;
        const bool mutual = P.polarization == Mutual;
        double c[20], frac[20], find[4], p[35], pd[35], m[20], cp[20], tq[3];
        for (int k = 0; k < 20; k++) c[k] = cartD[20*(size_t) s + k];
        for (int k = 0; k < 35; k++) { p[k] = phi[k*n + s]; pd[k] = phidp[k*n + s]; }
        const double ux = mu[3*(size_t) s], uy = mu[3*(size_t) s+1], uz = mu[3*(size_t) s+2];
        multipolesToFractional<double>(P.geom.A, c, frac);
        find[0] = 0;
        for (int k = 0; k < 3; k++) find[1+k] = P.geom.A[k][0]*ux + P.geom.A[k][1]*uy + P.geom.A[k][2]*uz;
        const bool an = aniso[s] != 0;
        double tx = 0, ty = 0, tz = 0;
        // induced part
        const bool addU = mutual && an;
        torqueMultipoles<double>(c, addU ? ux : 0.0, addU ? uy : 0.0, addU ? uz : 0.0, m);
        potentialToCartesian<double>(P.geom.A, pd, cp);
        reciprocalTorque<double>(m, cp, tq);
        tx += ke*tq[0]; ty += ke*tq[1]; tz += ke*tq[2];
        double eInd = 2.0*(find[1]*p[1] + find[2]*p[2] + find[3]*p[3]);
        double f[3];
        for (int d = 0; d < 3; d++) {
            const int dt = d == 0, du = d == 1, dv = d == 2;
            double v = 2.0*contractFractional<double>(find, 4, p, dt, du, dv);
            if (mutual) v += 2.0*contractFractional<double>(find, 4, pd, dt, du, dv);
            v += 2.0*contractFractional<double>(frac, 20, pd, dt, du, dv);
            f[d] = 0.5*ke*v;
        }
        // permanent part
        torqueMultipoles<double>(c, an ? ux : 0.0, an ? uy : 0.0, an ? uz : 0.0, m);
        potentialToCartesian<double>(P.geom.A, p, cp);
        reciprocalTorque<double>(m, cp, tq);
        tx += ke*tq[0]; ty += ke*tq[1]; tz += ke*tq[2];
        const double ePerm = contractFractional<double>(frac, 20, p, 0, 0, 0);
        for (int d = 0; d < 3; d++) f[d] += ke*contractFractional<double>(frac, 20, p, d == 0, d == 1, d == 2);
        double Fx = -(f[0]*P.geom.A[0][0] + f[1]*P.geom.A[1][0] + f[2]*P.geom.A[2][0]);
        double Fy = -(f[0]*P.geom.A[0][1] + f[1]*P.geom.A[1][1] + f[2]*P.geom.A[2][1]);
        double Fz = -(f[0]*P.geom.A[0][2] + f[1]*P.geom.A[1][2] + f[2]*P.geom.A[2][2]);
        // self torque on isotropic sites
        if (!an) {
            const double term = (2.0/3.0)*ke*P.alpha*P.alpha*P.alpha/MPID_SQRT_PI;
            const double dx = c[1], dy = c[2], dz = c[3];
            tx += term*2.0*(dy*uz - dz*uy); ty += term*2.0*(dz*ux - dx*uz); tz += term*2.0*(dx*uy - dy*ux);
        }
        // self energy
        const double* sp = sphD + 16*(size_t) s;
        double qii = 0, oii = 0;
        for (int k = 4; k < 9; k++) qii += sp[k]*sp[k];
        for (int k = 9; k < 16; k++) oii += sp[k]*sp[k];
        const double cii = sp[0]*sp[0];
        const double dii = sp[2]*(sp[2] + ux) + sp[3]*(sp[3] + uy) + sp[1]*(sp[1] + uz);
        const double a2 = P.alpha*P.alpha;
        e = 0.25*ke*eInd + 0.5*ke*ePerm
          - P.alpha*ke/MPID_SQRT_PI*(cii + (2.0/3.0)*a2*dii + (4.0/15.0)*a2*a2*qii + (8.0/105.0)*a2*a2*a2*oii);
        atomicAddFixed(&force[3*(size_t) s], Fx); atomicAddFixed(&force[3*(size_t) s+1], Fy); atomicAddFixed(&force[3*(size_t) s+2], Fz);
        atomicAddFixed(&torque[3*(size_t) s], tx); atomicAddFixed(&torque[3*(size_t) s+1], ty); atomicAddFixed(&torque[3*(size_t) s+2], tz);
    }
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAddFixed(energy, e);
}

// OPT dipole-response force and torque (:4956-4984, :2160-2188)
struct OptLists { const double* dip[8]; const double* field[8]; const double* grad[8]; double part[8]; int K; };
__global__ void k_opt_force(DevParams P, OptLists L, const int* __restrict__ aniso,
                            unsigned long long* __restrict__ force, unsigned long long* __restrict__ torque) {
    const int s = P.rowBegin + blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= P.rowEnd) return;
    const double ke = MPID_ELECTRIC;
    double fx = 0, fy = 0, fz = 0, tx = 0, ty = 0, tz = 0;
    const bool an = aniso[s] != 0;
    for (int l = 0; l < L.K-1; l++)
        for (int m = 0; m < L.K-1-l; m++) {
            const double p = L.part[l+m+1];
            if (fabs(p) < 1e-6) continue;
            const double* u = L.dip[l] + 3*(size_t) s;
            const double* g = L.grad[m] + 6*(size_t) s;
            fx += p*ke*(u[0]*g[0] + u[1]*g[3] + u[2]*g[4]);
            fy += p*ke*(u[0]*g[3] + u[1]*g[1] + u[2]*g[5]);
            fz += p*ke*(u[0]*g[4] + u[1]*g[5] + u[2]*g[2]);
            if (an) {
                const double* fl = L.field[m] + 3*(size_t) s;
                tx += p*ke*(u[1]*fl[2] - u[2]*fl[1]);
                ty += p*ke*(u[2]*fl[0] - u[0]*fl[2]);
                tz += p*ke*(u[0]*fl[1] - u[1]*fl[0]);
            }
        }
    atomicAddFixed(&force[3*(size_t) s], fx); atomicAddFixed(&force[3*(size_t) s+1], fy); atomicAddFixed(&force[3*(size_t) s+2], fz);
    if (an) { atomicAddFixed(&torque[3*(size_t) s], tx); atomicAddFixed(&torque[3*(size_t) s+1], ty); atomicAddFixed(&torque[3*(size_t) s+2], tz); }
}

// torque -> forces on the frame atoms (:2112-2131, :1895-2110); one thread per sorted atom
__global__ void k_torque_to_force(DevParams P, ParticleParams pp, const int* __restrict__ order, const int* __restrict__ inv,
                                  const double* __restrict__ posOrig, const unsigned long long* __restrict__ torque,
                                  unsigned long long* __restrict__ force, int sBegin, int sEnd) {
    // The map torque -> forces is linear in the torque, so with several ranks it is applied to each rank's PARTIAL torques
    // (atoms [sBegin, sEnd): the only ones it can have touched) before the forces are summed across ranks.
    const int s = sBegin + blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= sEnd) return;
    const int o = order[s];
    const int axis = pp.axis[o];
    if (axis == NoAxisType) return;
    const int az = pp.atomZ[o], ax = pp.atomX[o], ay = pp.atomY[o];
    if (az < 0) return;
    const double tq[3] = {fixedToDouble(torque[3*(size_t) s]), fixedToDouble(torque[3*(size_t) s+1]), fixedToDouble(torque[3*(size_t) s+2])};
    const double* pi = posOrig + 3*o;
    const double* pz = posOrig + 3*az;
    const double* px = ax >= 0 ? posOrig + 3*ax : pz;
    const double* py = ay >= 0 ? posOrig + 3*ay : pz;
    double fI[3], fZ[3], fX[3], fY[3];
    torqueToForce(axis, pi, pz, px, py, ax >= 0, ay >= 0, tq, fI, fZ, fX, fY);
    const int sz = inv[az];
    for (int k = 0; k < 3; k++) {
        atomicAddFixed(&force[3*(size_t) s + k], fI[k]);
        atomicAddFixed(&force[3*(size_t) sz + k], fZ[k]);
    }
    if (ax >= 0 && axis != ZOnly) { const int sx = inv[ax]; for (int k = 0; k < 3; k++) atomicAddFixed(&force[3*(size_t) sx + k], fX[k]); }
    if (ay >= 0) { const int sy = inv[ay]; for (int k = 0; k < 3; k++) atomicAddFixed(&force[3*(size_t) sy + k], fY[k]); }
}

// ---- OpenMM CUDA-platform data conventions (reference: platforms/cuda/src/MPIDCudaKernels.cpp) -------------------------
// posq: real4 per atom in the CudaContext's REORDERED atom order (slot i holds atom atomIndex[i]: cu.getAtomIndex(),
// MPIDCudaKernels.cpp:1089); float4 in single/mixed precision -- mixed adds the low bits from posqCorrection -- or double4
// in double precision (cu.getPosq(), :216).  Forces: cu.getForce(), signed 64-bit fixed point, scale 2^32, laid out
// [x: paddedNumAtoms][y: paddedNumAtoms][z: paddedNumAtoms] by slot and ACCUMULATED with atomicAdd (kernels/
// multipoleElectrostatics.cu:708-710).
template <typename T4>
__global__ void k_positions_from_cuda_context(int n, const T4* __restrict__ posq, const float4* __restrict__ posqCorrection,
                                              const int* __restrict__ atomIndex, double* __restrict__ posOrig) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T4 p = posq[i];
    double x = p.x, y = p.y, z = p.z;
    if (posqCorrection) { const float4 c = posqCorrection[i]; x += (double) c.x; y += (double) c.y; z += (double) c.z; }
    const int o = atomIndex[i];
    posOrig[3*(size_t) o] = x; posOrig[3*(size_t) o+1] = y; posOrig[3*(size_t) o+2] = z;
}
__global__ void k_forces_to_cuda_context(int n, int paddedNumAtoms, const int* __restrict__ atomIndex, const double* __restrict__ forcesOrig,
                                         unsigned long long* __restrict__ forceBuffer) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = atomIndex[i];
    for (int k = 0; k < 3; k++)
        atomicAdd(&forceBuffer[i + (size_t) k*paddedNumAtoms], (unsigned long long) __double2ll_rn(forcesOrig[3*(size_t) o + k]*MPID_FIXED_SCALE));
}

// forcesOrig[o] += fixed-point force of the sorted slot
__global__ void k_output_forces(int n, const int* __restrict__ order, const unsigned long long* __restrict__ force, double* __restrict__ forcesOrig) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = order[s];
    forcesOrig[3*(size_t) o]   += fixedToDouble(force[3*(size_t) s]);
    forcesOrig[3*(size_t) o+1] += fixedToDouble(force[3*(size_t) s+1]);
    forcesOrig[3*(size_t) o+2] += fixedToDouble(force[3*(size_t) s+2]);
}

// out[o] = vec[s] (3 components), optionally adding the lab permanent dipole
__global__ void k_unsort_vec3(int n, const int* __restrict__ order, const double* __restrict__ vec, const double* __restrict__ cartD,
                              int which, double* __restrict__ out) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int o = order[s];
    for (int k = 0; k < 3; k++) {
        double v = 0;
        if (which == 0 || which == 2) v += vec[3*(size_t) s + k];
        if (which == 1 || which == 2) v += cartD[20*(size_t) s + 1 + k];
        out[3*(size_t) o + k] = v;
    }
}

} // namespace mpid
#endif
