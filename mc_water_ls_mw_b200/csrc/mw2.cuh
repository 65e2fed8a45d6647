// mw2.cuh -- the walker kernel, second generation: one warp per LATTICE, one shared-memory block per lattice.
//
// The reference marks the parallelism itself: the `do ils = 1,num_lattices` loops around the local-energy calls of
// a translation (mc_moves.F90:1007-1018, :1076-1090) are independent per lattice.  Warp L of a walker's CTA owns
// lattice L.  Everything that warp touches in the hot loop -- positions, image vectors, cell, Verlet rows, bond
// masks, the bond records and the item table of the move in flight -- sits in ONE contiguous block at a fixed
// offset from a single base register, so the addressing of the hot code is "base + immediate (+ index)".
// Per trial move the two warps meet twice:
//     barrier A   both local-energy pairs (old, new) are in shared memory
//     warp 0      acceptance, weights, histograms, lattice switch          (mc_moves.F90:1104-1213, :1536-1689)
//     barrier B   the decision (accepted, active lattice) is in shared memory; each warp commits / restores its lattice
// Move generation (mc_moves.F90:1001-1067: molecule, direction, magnitude, both fractional transforms) is done for
// GB moves at a time with the moves spread over the lanes of warp 0, in the reference's exact arithmetic.
//
// The local energies of one lattice are evaluated as ONE list of items (formulation: mw_device.cuh):
//     own bonds   (imol at its old / trial position, list slot)        -> pair energy + bond record (u, g)
//     candidates  (neighbour j, in-range list slot of j not pointing back at imol) -> j-centred triplets of BOTH variants
// every item is "a list slot of a row, seen from a centre position": one inlined copy of the geometry, the radial
// functions and the exponential serves both kinds; the i-centred triplets pair the bond records afterwards.
//
// One-lattice boxes (single_box) run the same code with one warp.  Boxes of more than 64 molecules have no
// reverse-slot field in their list entries and stay with the first-generation kernel (mw_mc.cuh).
#pragma once
#include "mw_mc.cuh"

#ifndef MW2_BLOCKS
#define MW2_BLOCKS 14        // resident walkers (CTAs) per SM the register allocation is bounded for
#endif

#ifndef MW2_CHUNK
#define MW2_CHUNK 8          // cycles per unit of work when a batch is larger than the resident blocks of the GPU
#endif

namespace mw {
namespace v2 {

constexpr int RC2 = 32;      // bond records per lattice (old + new bonds of the moved molecule share them)
constexpr int IT2 = 96;      // items per lattice and round (own bonds first, then candidates)
constexpr int GB  = 7;       // trial moves generated per batch (7 x 8 draws fit one 64-number refill from any parity)
constexpr int GF  = 9;       // doubles per generated move: displacement, and its image in the other lattice when
                             // lattice 1 / lattice 2 is the active one

// item descriptor: slot | row << 5 | ra << 11 | rb << 16 | type << 21
//   own bond : ra = its record, type 1 (old position) / 2 (trial position)
//   candidate: ra / rb = record of the centre's bond to imol in the old / new variant (31 = not bonded), type 0
constexpr uint32_t IT_OLD = 1u << 21, IT_NEW = 2u << 21, IT_NONE = 31u;

// ---------------------------------------------------------------- shared-memory layout (byte offsets)
template <int NT>
struct Lay {
    int n;
    __host__ __device__ explicit Lay(int N) : n(N) {}
    __host__ __device__ __forceinline__ int N() const { return NT > 0 ? NT : n; }
    // one lattice block
    __host__ __device__ __forceinline__ int oP()  const { return 0; }                          // [3][N] fp64: x | y | z
    __host__ __device__ __forceinline__ int oV()  const { return 24 * N(); }                   // [3][IVC] image vectors
    __host__ __device__ __forceinline__ int oH()  const { return oV() + 24 * IVC; }            // [9] hmatrix, column-major
    __host__ __device__ __forceinline__ int oR()  const { return oH() + 72; }                  // [9] recip_matrix
    __host__ __device__ __forceinline__ int oT()  const { return oR() + 72; }                  // [6] trial position, displacement
    __host__ __device__ __forceinline__ int oQ()  const { return oT() + 48; }                  // [4][RC2] bond records ux|uy|uz|g
    __host__ __device__ __forceinline__ int oL()  const { return oQ() + 32 * RC2; }            // [N][LC] uint16 Verlet rows
    __host__ __device__ __forceinline__ int oBM() const { return oL() + 2 * LC * N(); }        // [N] uint32 in-range slots
    __host__ __device__ __forceinline__ int oIT() const { return oBM() + 4 * N(); }            // [IT2] uint32 items
    __host__ __device__ __forceinline__ int oNN() const { return oIT() + 4 * IT2; }            // [N] uint8 row lengths
    __host__ __device__ __forceinline__ int oRJ() const { return oNN() + ((N() + 3) & ~3); }   // [RC2] uint8 molecule of a record
    __host__ __device__ __forceinline__ int oNIV() const { return oRJ() + RC2; }               // int
    __host__ __device__ __forceinline__ int LB()  const { return (oNIV() + 4 + 15) & ~15; }
    // shared block (after nlat lattice blocks)
    static constexpr int sSC = 0;                                    // WalkerScalars
    static constexpr int sRB = 208;                                  // uint64 draw index of rngbuf[0]
    static constexpr int sRNG = 224;                                 // [RB] fp64
    static constexpr int sGEN = sRNG + 8 * RB;                       // [GB][GF] fp64 generated moves | [36] volume-move scratch
    static constexpr int sGI = sGEN + 512;                           // [GB] int molecule of a generated move
    static constexpr int sCTL = sGI + 32;                            // [8] int
    static constexpr int sXCH = sCTL + 32;                           // [8] fp64 energies exchanged between the warps
    static constexpr int sLV = sXCH + 64;                            // [2] log(V1/V2), log(V2/V1)
    static constexpr int sKV = sLV + 16;                             // [4] volume terms of the switch / of ls_mu (refresh_kv)
    static constexpr int sSWAP = sKV + 24;                           // int: warp 0 takes lattice 2 (the 4th kv slot is free)
    static constexpr int sUNIT = sKV + 28;                           // int: the unit of work the block took from the queue
    static constexpr int SB = sKV + 32;
    __host__ __device__ __forceinline__ size_t bytes(int nlat) const { return (size_t)nlat * LB() + SB; }
};
static_assert(sizeof(WalkerScalars) == 208, "layout");
static_assert(GB * GF * 8 <= 512 && 36 * 8 <= 512, "layout");
constexpr int CTL_STOP = 6, CTL_DEC = 7;

__host__ inline size_t walker_bytes(int N, int nlat) { return Lay<0>(N).bytes(nlat); }

template <typename T> __device__ __forceinline__ T* at(unsigned char* b, int off) { return (T*)(b + off); }
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---------------------------------------------------------------- staging: global memory <-> the walker's image
template <int NT>
__device__ __forceinline__ void load_walker(const Lay<NT> Y, const DeviceState& S, int wi, unsigned char* smem, int tid, int nt)
{
    const int N = Y.N(), nlat = S.nlat;
    for (int lat = 0; lat < nlat; ++lat) {
        unsigned char* lb = smem + lat * Y.LB();
        const double* gp = S.pos + ((size_t)wi * nlat + lat) * 3 * N;
        double* P = at<double>(lb, Y.oP());
        for (int t = tid; t < 3 * N; t += nt) P[t] = gp[t];
        const double* gi = S.iv + ((size_t)wi * nlat + lat) * 3 * IVC;
        double* V = at<double>(lb, Y.oV());
        for (int t = tid; t < 3 * IVC; t += nt) V[t] = gi[t];
        if (tid < 9) {
            at<double>(lb, Y.oH())[tid] = S.cell[((size_t)wi * nlat + lat) * 9 + tid];
            at<double>(lb, Y.oR())[tid] = S.recip[((size_t)wi * nlat + lat) * 9 + tid];
        }
        if (tid == 0) *at<int>(lb, Y.oNIV()) = S.niv[wi * 2 + lat];
        const uint4* gl = (const uint4*)(S.list + ((size_t)wi * nlat + lat) * N * LC);
        uint4* sl = at<uint4>(lb, Y.oL());
        for (int t = tid; t < N * LC / 8; t += nt) sl[t] = gl[t];
        const uint8_t* gn = S.nn + ((size_t)wi * nlat + lat) * N;
        uint8_t* NN = at<uint8_t>(lb, Y.oNN());
        for (int t = tid; t < N; t += nt) NN[t] = gn[t];
    }
    unsigned char* sb = smem + nlat * Y.LB();
    const uint32_t* gs = (const uint32_t*)(S.scal + wi);
    uint32_t* ss = at<uint32_t>(sb, Lay<NT>::sSC);
    for (int t = tid; t < (int)(sizeof(WalkerScalars) / 4); t += nt) ss[t] = gs[t];
}

template <int NT>
__device__ __forceinline__ void store_walker(const Lay<NT> Y, const DeviceState& S, int wi, unsigned char* smem, int tid, int nt)
{
    const int N = Y.N(), nlat = S.nlat;
    for (int lat = 0; lat < nlat; ++lat) {
        unsigned char* lb = smem + lat * Y.LB();
        double* gp = S.pos + ((size_t)wi * nlat + lat) * 3 * N;
        const double* P = at<double>(lb, Y.oP());
        for (int t = tid; t < 3 * N; t += nt) gp[t] = P[t];
        double* gi = S.iv + ((size_t)wi * nlat + lat) * 3 * IVC;
        const double* V = at<double>(lb, Y.oV());
        for (int t = tid; t < 3 * IVC; t += nt) gi[t] = V[t];
        if (tid < 9) {
            S.cell[((size_t)wi * nlat + lat) * 9 + tid] = at<double>(lb, Y.oH())[tid];
            S.recip[((size_t)wi * nlat + lat) * 9 + tid] = at<double>(lb, Y.oR())[tid];
        }
        if (tid == 0) S.niv[wi * 2 + lat] = *at<int>(lb, Y.oNIV());
        uint4* gl = (uint4*)(S.list + ((size_t)wi * nlat + lat) * N * LC);
        const uint4* sl = at<uint4>(lb, Y.oL());
        for (int t = tid; t < N * LC / 8; t += nt) gl[t] = sl[t];
        uint8_t* gn = S.nn + ((size_t)wi * nlat + lat) * N;
        const uint8_t* NN = at<uint8_t>(lb, Y.oNN());
        for (int t = tid; t < N; t += nt) gn[t] = NN[t];
    }
    unsigned char* sb = smem + nlat * Y.LB();
    uint32_t* gs = (uint32_t*)(S.scal + wi);
    const uint32_t* ss = at<uint32_t>(sb, Lay<NT>::sSC);
    for (int t = tid; t < (int)(sizeof(WalkerScalars) / 4); t += nt) gs[t] = ss[t];
}

// ---------------------------------------------------------------- cold paths on one lattice block (one warp)
// molint.F90:174-217.  Lane k builds vector k.
template <int NT>
__device__ __noinline__ int compute_ivects(const Lay<NT> Y, unsigned char* lb)
{
    const double* h = at<double>(lb, Y.oH());
    double* V = at<double>(lb, Y.oV());
    const double l1 = xsqrt(xa(xa(xm(h[0], h[0]), xm(h[1], h[1])), xm(h[2], h[2])));
    const double l2 = xsqrt(xa(xa(xm(h[3], h[3]), xm(h[4], h[4])), xm(h[5], h[5])));
    const double l3 = xsqrt(xa(xa(xm(h[6], h[6]), xm(h[7], h[7])), xm(h[8], h[8])));
    const int im = (int)floor(xd(RC, l1)) + 1;
    const int jm = (int)floor(xd(RC, l2)) + 1;
    const int km = (int)floor(xd(RC, l3)) + 1;
    const int nj = 2 * jm + 1, nk = 2 * km + 1;
    const int nv = (2 * im + 1) * nj * nk;
    const int k = lane_id();
    if (k == 0) *at<int>(lb, Y.oNIV()) = nv;
    if (nv > IVC) { __syncwarp(); return ERR_IVECT_OVERFLOW; }
    const int f0 = (nv - 1) / 2;
    if (k < nv) {
        double vx = 0.0, vy = 0.0, vz = 0.0;
        if (k > 0) {
            const int f = (k <= f0) ? k - 1 : k;
            const int ic = f / (nj * nk) - im;
            const int jc = (f / nk) % nj - jm;
            const int kc = f % nk - km;
            const double a = (double)ic, b = (double)jc, c = (double)kc;
            vx = xa(xa(xm(a, h[0]), xm(b, h[3])), xm(c, h[6]));
            vy = xa(xa(xm(a, h[1]), xm(b, h[4])), xm(c, h[7]));
            vz = xa(xa(xm(a, h[2]), xm(b, h[5])), xm(c, h[8]));
        }
        V[k] = vx; V[IVC + k] = vy; V[2 * IVC + k] = vz;
    }
    __syncwarp();
    return 0;
}

// molint.F90:501-559 with the candidate pruning of mw_device.cuh; entries j | rev << 6 | image << 11
template <int NT>
__device__ __noinline__ int compute_neighbours(const Lay<NT> Y, unsigned char* lb)
{
    int err = compute_ivects<NT>(Y, lb);                        // molint.F90:518
    if (err) return err;
    const int N = Y.N(), lane = lane_id();
    const int nv = *at<int>(lb, Y.oNIV());
    const double* P = at<double>(lb, Y.oP());
    const double* V = at<double>(lb, Y.oV());
    const double* h = at<double>(lb, Y.oH());
    uint16_t* L = at<uint16_t>(lb, Y.oL());
    uint8_t* NN = at<uint8_t>(lb, Y.oNN());
    double b[9];
    b[0] = h[4] * h[8] - h[5] * h[7]; b[1] = h[5] * h[6] - h[3] * h[8]; b[2] = h[3] * h[7] - h[4] * h[6];
    b[3] = h[7] * h[2] - h[8] * h[1]; b[4] = h[8] * h[0] - h[6] * h[2]; b[5] = h[6] * h[1] - h[7] * h[0];
    b[6] = h[1] * h[5] - h[2] * h[4]; b[7] = h[2] * h[3] - h[0] * h[5]; b[8] = h[0] * h[4] - h[1] * h[3];
    const double idet = 1.0 / (h[0] * b[0] + h[1] * b[1] + h[2] * b[2]);
#pragma unroll
    for (int k = 0; k < 9; ++k) b[k] *= idet;
    const double R0 = RN * (1.0 + 1e-9) * sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) + 1e-12;
    const double R1 = RN * (1.0 + 1e-9) * sqrt(b[3] * b[3] + b[4] * b[4] + b[5] * b[5]) + 1e-12;
    const double R2 = RN * (1.0 + 1e-9) * sqrt(b[6] * b[6] + b[7] * b[7] + b[8] * b[8]) + 1e-12;
    const bool boxed = (nv == 27);
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const double ix = P[i], iy = P[N + i], iz = P[2 * N + i];
        int total = 0;
#pragma unroll 1
        for (int jb = 0; jb < N; jb += 32) {
            const int j = jb + lane;
            uint32_t m = 0;
            if (j < N) {
                const double vx = xs(P[j], ix), vy = xs(P[N + j], iy), vz = xs(P[2 * N + j], iz);
                uint32_t cand = lowbits(nv);
                if (boxed) {
                    const double s0 = b[0] * vx + b[1] * vy + b[2] * vz;
                    const double s1 = b[3] * vx + b[4] * vy + b[5] * vz;
                    const double s2 = b[6] * vx + b[7] * vy + b[8] * vz;
                    constexpr uint32_t A0 = image_axis_mask(0, -1), A1 = image_axis_mask(0, 0), A2 = image_axis_mask(0, 1);
                    constexpr uint32_t B0 = image_axis_mask(1, -1), B1 = image_axis_mask(1, 0), B2 = image_axis_mask(1, 1);
                    constexpr uint32_t C0 = image_axis_mask(2, -1), C1 = image_axis_mask(2, 0), C2 = image_axis_mask(2, 1);
                    cand = axis_candidates(s0, R0, A0, A1, A2) & axis_candidates(s1, R1, B0, B1, B2) & axis_candidates(s2, R2, C0, C1, C2);
                }
#pragma unroll 1
                while (cand) {
                    const int k = __ffs(cand) - 1; cand &= cand - 1;
                    const double tx = xa(vx, V[k]), ty = xa(vy, V[IVC + k]), tz = xa(vz, V[2 * IVC + k]);
                    const double r2 = xa(xa(xm(tx, tx), xm(ty, ty)), xm(tz, tz));
                    if (r2 < RN2) m |= 1u << k;
                }
                if (j == i) {
                    m &= ~1u;                                   // (k==1).and.(jmol==imol) cycle
                    if (m) err |= ERR_SELF_IMAGE;
                }
            }
            const int cnt = __popc(m);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            int off = total + incl - cnt;
            uint16_t* row = L + i * LC;
#pragma unroll 1
            while (m) {
                const int k = __ffs(m) - 1; m &= m - 1;
                if (off < LC) row[off] = (uint16_t)((k << 11) | j);
                ++off;
            }
            total += __shfl_sync(FULL, incl, 31);
        }
        if (total > LC) { err |= ERR_LIST_OVERFLOW; total = LC; }
        if (lane == 0) NN[i] = (uint8_t)total;
    }
    err = (int)__reduce_or_sync(FULL, (unsigned)err);
    __syncwarp();
    // reverse slots: lanes = the slots of row i; a row is sorted by (j, image): bisection in row j
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        const int nni = NN[i];
        if (lane < nni) {
            uint16_t* row = L + i * LC;
            const uint32_t e = row[lane];
            const int j = e & 63, img = e >> 11;
            const uint32_t want = ((uint32_t)i << 5) | (uint32_t)inverse_image(img, nv);
            const uint16_t* rj = L + j * LC;
            int lo = 0, hi = (int)NN[j] - 1;
#pragma unroll 1
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint32_t e2 = rj[mid];
                const uint32_t have = ((e2 & 63u) << 5) | (e2 >> 11);
                if (have < want) lo = mid + 1; else hi = mid;
            }
            row[lane] = (uint16_t)(e | ((uint32_t)lo << 6));
        }
    }
    __syncwarp();
    return err;
}

// bmask[a] bit s <=> slot s of a's row is a SIGNIFICANT bond: inside RSIG < a*sigma (molint.F90:276/454 test
// r^2 < rcsq; the shell between RSIG and the cut-off carries three-body factors below 1e-17, mw_device.cuh)
template <int NT>
__device__ __noinline__ void compute_bond_masks(const Lay<NT> Y, unsigned char* lb)
{
    const int N = Y.N(), lane = lane_id();
    const double* P = at<double>(lb, Y.oP());
    const double* V = at<double>(lb, Y.oV());
    const uint16_t* L = at<uint16_t>(lb, Y.oL());
    const uint8_t* NN = at<uint8_t>(lb, Y.oNN());
    uint32_t* BM = at<uint32_t>(lb, Y.oBM());
    for (int a = 0; a < N; ++a) {
        const bool has = lane < NN[a];
        const uint32_t e = has ? L[a * LC + lane] : 0u;
        const int j = e & 63, img = e >> 11;
        const double r2 = dist2((P[j] + V[img]) - P[a], (P[N + j] + V[IVC + img]) - P[N + a], (P[2 * N + j] + V[2 * IVC + img]) - P[2 * N + a]);
        const uint32_t m = __ballot_sync(FULL, has && r2 < CK.rsig2);
        if (lane == 0) BM[a] = m;
    }
    __syncwarp();
}

// compute_model_energy (molint.F90:407-499) of one lattice block, molecule-chunked through the block's record
// table (rotation pairing as in mw_device.cuh); refreshes the bond masks.  Returns E (uniform).
template <int NT>
__device__ __noinline__ double full_energy(const Lay<NT> Y, unsigned char* lb)
{
    const int N = Y.N(), lane = lane_id();
    const unsigned lt = lt_mask();
    const double* P = at<double>(lb, Y.oP());
    const double* V = at<double>(lb, Y.oV());
    const uint16_t* L = at<uint16_t>(lb, Y.oL());
    const uint8_t* NN = at<uint8_t>(lb, Y.oNN());
    uint32_t* BM = at<uint32_t>(lb, Y.oBM());
    double* q = at<double>(lb, Y.oQ());
    uint32_t* qmeta = at<uint32_t>(lb, Y.oIT());
    double acc = 0.0;
    int a = 0;
    while (a < N) {
        int nq = 0;
        int a1 = a;
        for (; a1 < N; ++a1) {
            const bool has = lane < NN[a1];
            const uint32_t e = has ? L[a1 * LC + lane] : 0u;
            const int j = e & 63, img = e >> 11;
            const double tx = (P[j] + V[img]) - P[a1];
            const double ty = (P[N + j] + V[IVC + img]) - P[N + a1];
            const double tz = (P[2 * N + j] + V[2 * IVC + img]) - P[2 * N + a1];
            const double r2 = dist2(tx, ty, tz);
            const bool f = has && r2 < RCSQ;
            const uint32_t bm = __ballot_sync(FULL, f);
            const uint32_t bs = __ballot_sync(FULL, has && r2 < CK.rsig2);
            const int cnt = __popc(bm);
            if (nq + cnt > RC2) break;                // cnt <= LC == RC2: a chunk always holds >= 1 molecule
            if (lane == 0) BM[a1] = bs;
            if (f) {
                const int io = nq + __popc(bm & lt);
                q[io] = tx; q[RC2 + io] = ty; q[2 * RC2 + io] = tz; q[3 * RC2 + io] = r2;
                qmeta[io] = (uint32_t)cnt | ((uint32_t)__popc(bm & lt) << 8);
            }
            nq += cnt;
        }
        __syncwarp();
        {   // bond evaluation: 0.5 * pair energy (molint.F90:464); one pass (nq <= 32)
            const int r = lane;
            if (r < nq) {
                const double tx = q[r], ty = q[RC2 + r], tz = q[2 * RC2 + r], r2 = q[3 * RC2 + r];
                double ir, isr;
                bond_radial(r2, ir, isr);
                const double e1 = exp_fast(CK.sig02 * isr);
                const double e_2 = e1 * e1, e_4 = e_2 * e_2;
                const double s2 = CK.ss * ir * ir;
                q[r] = tx * ir; q[RC2 + r] = ty * ir; q[2 * RC2 + r] = tz * ir; q[3 * RC2 + r] = e_4 * e_2;
                acc += 0.5 * (CK.aeps * (CK.bigb * (s2 * s2) - 1.0) * (e_4 * e1));
            }
        }
        __syncwarp();
        {
            const int r = lane;
            const bool act = r < nq;
            const uint32_t qm = act ? qmeta[r] : 0u;
            const int n = qm & 255, pos = qm >> 8;
            const int half = n >> 1, send = r - pos + n;
            const bool even = !(n & 1);
            const double ux = act ? q[r] : 0.0, uy = act ? q[RC2 + r] : 0.0, uz = act ? q[2 * RC2 + r] : 0.0;
            const double g = act ? q[3 * RC2 + r] : 0.0;
            double tb = 0.0;
            const int maxd = __reduce_max_sync(FULL, half);
            for (int d = 1; d <= maxd; ++d) {
                int c = r + d;
                c = (c >= send) ? c - n : c;
                const bool on = (d <= half) && !(even && d == half && pos >= half);
                c = on ? c : r;
                const double ct = ux * q[c] + uy * q[RC2 + c] + uz * q[2 * RC2 + c];
                const double dd = ct - CK.cos0;                 // no k==i filter in compute_model_energy (molint.F90:480-483)
                tb += on ? q[3 * RC2 + c] * dd * dd : 0.0;
            }
            acc += CK.leps * g * tb;
        }
        __syncwarp();
        a = a1;
    }
    return warp_sum(acc);
}

template <int NT>
__device__ __noinline__ void rescale_all(const Lay<NT> Y, unsigned char* lb, double* R)
{
    const int N = Y.N(), lane = lane_id();
    double rm[9], hm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { rm[k] = at<double>(lb, Y.oR())[k]; hm[k] = at<double>(lb, Y.oH())[k]; }
    double* P = at<double>(lb, Y.oP());
    for (int i = lane; i < N; i += 32) {
        double x = P[i], y = P[N + i], z = P[2 * N + i];
        rescale_pos(x, y, z, rm, hm);
        P[i] = x; P[N + i] = y; P[2 * N + i] = z;
        x = R[i]; y = R[N + i]; z = R[2 * N + i];
        rescale_pos(x, y, z, rm, hm);
        R[i] = x; R[N + i] = y; R[2 * N + i] = z;
    }
    __syncwarp();
}

template <int NT>
__device__ __noinline__ void refresh_recip(const Lay<NT> Y, unsigned char* lb)
{
    double hm[9], rm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = at<double>(lb, Y.oH())[k];
    recipmatrix3(hm, rm);
    __syncwarp();
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) at<double>(lb, Y.oR())[k] = rm[k];
    }
    __syncwarp();
}

template <int NT>
__device__ __forceinline__ double cell_volume(const Lay<NT> Y, unsigned char* lb)
{
    double hm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = at<double>(lb, Y.oH())[k];
    return fabs(determinant3(hm));
}

// mc_volume (mc_moves.F90:1216-1534), executed by warp 0 on both lattice blocks.  Returns the new position in
// the random-number buffer.
template <int NLAT, int NT>
__device__ __noinline__ int volume_move(const Lay<NT> Y, unsigned char* smem, const DeviceState& S, const McParams& p, int wi, int rng_pos)
{
    const int N = Y.N(), lane = lane_id();
    unsigned char* sb = smem + NLAT * Y.LB();
    WalkerScalars* sc = at<WalkerScalars>(sb, Lay<NT>::sSC);
    const double* rngbuf = at<double>(sb, Lay<NT>::sRNG);
    double* lv = at<double>(sb, Lay<NT>::sLV);
    double* save = at<double>(sb, Lay<NT>::sGEN);           // [NLAT][18]: old cell | old recip
    const double Nd = (double)N;
    const double* wgt = S.weight + (size_t)wi * S.NB;
    double backupE[2] = {0.0, 0.0}, old_vol[2] = {0.0, 0.0}, newE[2] = {0.0, 0.0};
    int err = 0;
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        backupE[lat] = sc->E[lat];
        old_vol[lat] = sc->vol[lat];
        refresh_recip<NT>(Y, smem + lat * Y.LB());                          // :1260-1262
    }
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat)
        if (lane < 9) {
            save[lat * 18 + lane] = at<double>(smem + lat * Y.LB(), Y.oH())[lane];
            save[lat * 18 + 9 + lane] = at<double>(smem + lat * Y.LB(), Y.oR())[lane];
        }
    __syncwarp();
    double x = rngbuf[rng_pos++];
    const int idim = (int)xm(x, 3.0) + 1;
    x = rngbuf[rng_pos++];
    const int jdim = (int)xm(x, 3.0) + 1;
    x = rngbuf[rng_pos++];
    const double dh = xm(xs(xm(2.0, x), 1.0), sc->dv_max);
    if (lane < NLAT) {
        double* hm = at<double>(smem + lane * Y.LB(), Y.oH());
        const double v = xa(MW_H(hm, idim, jdim), dh);
        if (idim != jdim) MW_H(hm, jdim, idim) = xa(MW_H(hm, jdim, idim), dh);
        MW_H(hm, idim, jdim) = v;
    }
    __syncwarp();
    double* refpos = S.ref + (size_t)wi * NLAT * 3 * N;
    double vol_new[2] = {0.0, 0.0};
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        unsigned char* lb = smem + lat * Y.LB();
        rescale_all<NT>(Y, lb, refpos + (size_t)lat * 3 * N);               // recip = old cell's, h = new cell
        vol_new[lat] = cell_volume<NT>(Y, lb);
        refresh_recip<NT>(Y, lb);
        err |= compute_ivects<NT>(Y, lb);
        newE[lat] = full_energy<NT>(Y, lb);
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) { sc->vol[lat] = vol_new[lat]; sc->E[lat] = newE[lat]; }
    }
    __syncwarp();
    double old_eta = 0.0, new_eta = 0.0, old_mu = 0.0, new_mu = 0.0;
    const bool one = (sc->ls == 1);
    double nlv12 = lv[0], nlv21 = lv[1];
    if (NLAT == 2) {
        old_eta = eta_bin(p, S.mubin, S.ginv, sc, wgt, sc->mu).eta;
        old_mu = sc->mu;
        nlv12 = log(sc->vol[0] / sc->vol[1]); nlv21 = log(sc->vol[1] / sc->vol[0]);
        new_mu = mu_paren(p, sc, Nd, nlv12);
        new_eta = eta_bin(p, S.mubin, S.ginv, sc, wgt, new_mu).eta;
    }
    x = rngbuf[rng_pos++];
    const double dE = one ? newE[0] - backupE[0] : newE[1] - backupE[1];
    const double Vs = one ? vol_new[0] : vol_new[1], Vo = one ? old_vol[0] : old_vol[1];
    const double diffkT = p.beta * dE + new_eta - old_eta + p.beta * p.pressure * (Vs - Vo) - Nd * log(Vs / Vo);
    const double compare = fmin(1.0, exp(-diffkT));
    __syncwarp();
    if (x < compare) {
        if (lane == 0) {
            sc->acc_v += 1;
            if (NLAT == 2) {
                sc->mu = new_mu;
                const double dmu = fabs(old_mu - new_mu);
                if (dmu < sc->min_dmu) sc->min_dmu = dmu;
                if (dmu > sc->max_dmu) sc->max_dmu = dmu;
            }
            lv[0] = nlv12; lv[1] = nlv21;
        }
        __syncwarp();
    } else {
        // :1434-1528: V,h <- old; rescale with recip(NEW) and h(OLD); recip <- old; ivects; E <- backup
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat)
            if (lane < 9) at<double>(smem + lat * Y.LB(), Y.oH())[lane] = save[lat * 18 + lane];
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) rescale_all<NT>(Y, smem + lat * Y.LB(), refpos + (size_t)lat * 3 * N);
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat)
            if (lane < 9) at<double>(smem + lat * Y.LB(), Y.oR())[lane] = save[lat * 18 + 9 + lane];
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int lat = 0; lat < NLAT; ++lat) { sc->vol[lat] = old_vol[lat]; sc->E[lat] = backupE[lat]; }
        }
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            err |= compute_ivects<NT>(Y, smem + lat * Y.LB());
            compute_bond_masks<NT>(Y, smem + lat * Y.LB());    // positions moved by rounding; keep masks fresh
        }
        if (NLAT == 2) {
            const double mu_back = mu_paren(p, sc, Nd, lv[0]);
            __syncwarp();
            if (lane == 0) sc->mu = mu_back;
            __syncwarp();
        }
    }
    if (err) { if (lane == 0) sc->error |= err; __syncwarp(); }
    return rng_pos;
}

// The parts of the lattice-switch exponent (mc_moves.F90:1562-1574) and of ls_mu (:1583-1585) that only change
// with the cell: -diffkT(switch 1->2) = -(beta*(E2-E1) + kv[0]), -diffkT(2->1) = -(beta*(E1-E2) + kv[1]),
// ls_mu = beta*(E1-E2) + kv[2].  Energy-class arithmetic (free association): lane 0 of warp 0 refreshes
// them at kernel start and after every volume move.
__device__ __forceinline__ void refresh_kv(const McParams& p, const WalkerScalars* sc, const double* lv, double* kv, double Nd)
{
    const double bp = p.beta * p.pressure;
    double sh = 0.0;
    if (p.leshift) sh = p.beta * (sc->refH[0] - sc->refH[1]);
    const double dv = sc->vol[0] - sc->vol[1];
    kv[0] = (p.npt ? (-bp * dv - Nd * lv[1]) : 0.0) + sh;
    kv[1] = (p.npt ? (bp * dv - Nd * lv[0]) : 0.0) - sh;
    kv[2] = bp * dv - sh - Nd * lv[0];
}

// sum two per-lane accumulators over the warp and broadcast both totals
__device__ __forceinline__ void reduce2(double& a0, double& a1)
{
    const bool up = lane_id() & 16;
    const double s = up ? a0 : a1;                      // lanes 0-15 keep a0, lanes 16-31 keep a1
    const double r = __shfl_xor_sync(FULL, s, 16);
    double a = (up ? a1 : a0) + r;
    a += __shfl_xor_sync(FULL, a, 8);
    a += __shfl_xor_sync(FULL, a, 4);
    a += __shfl_xor_sync(FULL, a, 2);
    a += __shfl_xor_sync(FULL, a, 1);
    a0 = __shfl_sync(FULL, a, 0);
    a1 = __shfl_sync(FULL, a, 16);
}

// ---------------------------------------------------------------- local energies of one lattice block (one warp)
// compute_local_real_energy(imol) at the old position and -- when with_new -- at the trial position T[0..2] of the
// block (molint.F90:220-404; mc_moves.F90:1010,1083).  Returns the energies (uniform), the in-range slot masks of
// imol's row for both positions, and error bits.
// WPL = warps per lattice.  WPL == 2 (small ensembles, where a step lasts as long as one walker's serial chain):
// the two warps of a lattice run stage 1 and the item table redundantly (same values, no exchange), then split
// the work: warp `sub` takes the item passes sub, sub + 2, ... (the bond records of pass 0 cross through the
// lattice's named barrier `bar`) and the i-centred pairs of variant `sub` (0 old, 1 new).  The energies returned
// are this warp's PARTIAL sums; the caller adds the two warps' parts.
__device__ __forceinline__ void lat_bar(int bar)
{
    if (bar == 1) asm volatile("bar.sync 1, 64;" ::: "memory");         // immediates: the block reserves 3 barriers, not 16
    else asm volatile("bar.sync 2, 64;" ::: "memory");
}

// ILP = item passes in flight per loop turn (one warp per lattice; registers to spare and nothing else to issue when
// the GPU holds few walkers): 3 in the instantiation for at most four walkers per SM, 1 everywhere else;
// the dependent chains of the passes (geometry -> 1/r -> exponential) interleave.
template <int NT, int WPL, int ILP>
__device__ __forceinline__ int local_energies(const Lay<NT> Y, unsigned char* lb, int imol, bool with_new, int sub, int bar,
                                              double& eo, double& en, uint32_t& mo, uint32_t& mn)
{
    const int N = Y.N(), lane = lane_id();
    const unsigned lt = lt_mask();
    const double* P = at<double>(lb, Y.oP());
    const double* V = at<double>(lb, Y.oV());
    const double* T = at<double>(lb, Y.oT());
    const uint16_t* L = at<uint16_t>(lb, Y.oL());
    const uint32_t* BM = at<uint32_t>(lb, Y.oBM());
    double* q = at<double>(lb, Y.oQ());
    uint32_t* items = at<uint32_t>(lb, Y.oIT());
    uint8_t* recj = at<uint8_t>(lb, Y.oRJ());
    int err = 0;

    // ---- stage 1: lanes = slots of imol's row.  Distance tests at both positions: in range (own bonds, pair
    // energies) and significant (legs of j-centred triplets)
    const int nni = at<uint8_t>(lb, Y.oNN())[imol];
    const bool has = lane < nni;
    const uint32_t e = has ? L[imol * LC + lane] : 0u;
    const int j = e & 63;
    uint32_t bo, bn, so, sn;
    {
        const int img = e >> 11;
        const double pjx = P[j] + V[img], pjy = P[N + j] + V[IVC + img], pjz = P[2 * N + j] + V[2 * IVC + img];
        const double r2o = dist2(pjx - P[imol], pjy - P[N + imol], pjz - P[2 * N + imol]);
        const double r2n = dist2(pjx - T[0], pjy - T[1], pjz - T[2]);
        bo = __ballot_sync(FULL, has && r2o < CK.rcc2);
        so = __ballot_sync(FULL, has && r2o < CK.rsig2);
        bn = with_new ? __ballot_sync(FULL, has && r2n < CK.rcc2) : 0u;
        sn = with_new ? __ballot_sync(FULL, has && r2n < CK.rsig2) : 0u;
    }
    mo = so; mn = sn;
    // slots of row j that point back at imol (any image) are no candidates: the k == i entries of the reference's
    // list B are either filtered (cos = 1) or covered by the factor 3 of the i-centred pairs.  Cells narrower than
    // twice the list radius hold two images of one molecule in most rows; a row is sorted by molecule, so the
    // images sit in neighbouring lanes.
    uint32_t excl = 1u << ((e >> 6) & 31u);
    {
        const uint32_t el = __shfl_up_sync(FULL, e, 1), er = __shfl_down_sync(FULL, e, 1), el2 = __shfl_up_sync(FULL, e, 2);
        if (has && lane > 0 && ((el ^ e) & 63u) == 0u) excl |= 1u << ((el >> 6) & 31u);
        if (lane + 1 < nni && ((er ^ e) & 63u) == 0u) excl |= 1u << ((er >> 6) & 31u);
        if (__any_sync(FULL, has && lane > 1 && ((el2 ^ e) & 63u) == 0u)) {   // three or more images: general form
#pragma unroll 1
            for (int l2 = 0; l2 < nni; ++l2) {
                const uint32_t e2 = __shfl_sync(FULL, e, l2);
                if (((e2 ^ e) & 63u) == 0u) excl |= 1u << ((e2 >> 6) & 31u);
            }
        }
    }
    double ao = 0.0, an = 0.0;
    // Both variants share one pass (the j-k geometry of a candidate serves the old and the new position) unless
    // their bonds together exceed the record table -- compressed cells after a large volume move --: then the
    // old and the new variant are evaluated one after the other.
    const int npass = (__popc(bo) + __popc(bn) < RC2) ? 1 : 2;
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
        const bool useo = (npass == 1 || pass == 0), usen = (npass == 1 || pass == 1);
        const uint32_t bop = useo ? bo : 0u, bnp = usen ? bn : 0u;
        const bool fo = (bop >> lane) & 1u, fn = (bnp >> lane) & 1u;
        const bool go = useo && ((so >> lane) & 1u), gn = usen && ((sn >> lane) & 1u);   // significant legs
        const int no = __popc(bop), nw = __popc(bnp), nown = no + nw;
        const uint32_t ro = __popc(bop & lt), rn = no + __popc(bnp & lt);
        if (nown >= RC2) { err |= ERR_BOND_OVERFLOW; break; }            // every slot of the row in range: flagged
        const uint32_t own = (uint32_t)lane | ((uint32_t)imol << 5);
        if (fo) { items[ro] = own | (ro << 11) | IT_OLD; recj[ro] = (uint8_t)j; }
        if (fn) { items[rn] = own | (rn << 11) | IT_NEW; recj[rn] = (uint8_t)j; }
        const uint32_t bmj0 = (go || gn) ? (BM[j] & ~excl) : 0u;
        const uint32_t dbase = ((uint32_t)j << 5) | ((go ? ro : IT_NONE) << 11) | ((gn ? rn : IT_NONE) << 16);

        // ---- rounds: one, when the candidates fit the table beside the own bonds (nearly always); a dense walker
        // (ten bonds per molecule: > 96 items) takes its candidates in windows of the table's free part, two or
        // three rounds (windows of the candidate enumeration, not of the centres: a centre's candidates may straddle
        // two rounds)
        int incl0 = __popc(bmj0);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl0, d);
            if (lane >= d) incl0 += t;
        }
        const int ncand = __shfl_sync(FULL, incl0, 31);
        const int cap = IT2 - nown;                            // nown < RC2: at least 65 entries
        const int nrounds = (ncand <= cap) ? 1 : (ncand + cap - 1) / cap;
#pragma unroll 1
        for (int rd = 0; rd < nrounds; ++rd) {
            const int first = (rd == 0) ? 0 : nown;            // own bonds are evaluated in the first round only
            int nitems = nown + ncand;
            if (nrounds == 1) {
                uint32_t bmj = bmj0;
                uint32_t* it = items + (nown + incl0 - __popc(bmj0));
                if (ILP >= 2) {
                    // few walkers per SM (latency, not issue slots): lowest slot from the front and highest from the
                    // back in one turn -- half the dependent turns, the same table (an odd one out is written twice)
                    uint32_t* ie = it + __popc(bmj0) - 1;
#pragma unroll 1
                    while (bmj) {
                        const int lo = __ffs(bmj) - 1, hi = 31 - __clz(bmj);
                        *it++ = dbase | (uint32_t)lo;
                        *ie-- = dbase | (uint32_t)hi;
                        bmj &= bmj - 1;
                        bmj &= ~(1u << hi);
                    }
                } else {
#pragma unroll 1
                    while (bmj) {
                        const int s2 = __ffs(bmj) - 1; bmj &= bmj - 1;
                        *it++ = dbase | (uint32_t)s2;
                    }
                }
            } else {
                const int lo = rd * cap, hi = min(ncand, lo + cap);
                nitems = nown + (hi - lo);
                uint32_t bmj = bmj0;
                int gi = incl0 - __popc(bmj0);
#pragma unroll 1
                while (bmj) {
                    const int s2 = __ffs(bmj) - 1; bmj &= bmj - 1;
                    if (gi >= lo && gi < hi) items[nown + gi - lo] = dbase | (uint32_t)s2;
                    ++gi;
                }
            }
            __syncwarp();
            // ---- items: geometry, radial functions, one exponential; own bonds leave their record and pair
            // energy, candidates close the j-centred triplets of both variants
            bool synced = (WPL == 1);
            // one item: the geometry of "list slot of a row, seen from a centre", 1/r, the exponential
            struct Item { uint32_t d; double ux, uy, uz, g, ir, e1, e_4; bool ok; };
            auto item_at = [&](int t0_) -> Item {
                Item a;
                const int t = t0_ + lane;
                const bool in = t < nitems;
                a.d = items[in ? t : first];
                const uint32_t ty = a.d >> 21;
                const uint32_t e2 = L[((a.d >> 5) & 63u) * LC + (a.d & 31u)];
                const int k = e2 & 63, im2 = e2 >> 11;
                // centre: a molecule of the block, or the trial position
                const double* cp = (ty == 2u) ? T : P + ((a.d >> 5) & 63u);
                const int cs = (ty == 2u) ? 1 : N;
                const double tx = (P[k] + V[im2]) - cp[0];
                const double ty_ = (P[N + k] + V[IVC + im2]) - cp[cs];
                const double tz = (P[2 * N + k] + V[2 * IVC + im2]) - cp[2 * cs];
                const double sq0 = dist2(tx, ty_, tz);
                a.ok = in && (sq0 < CK.rcc2);
                const double sq = a.ok ? sq0 : CK.ss;                  // any length inside the cut-off
                double isr;
                bond_radial(sq, a.ir, isr);
                a.e1 = exp_nc(CK.sig02 * isr);                         // exp(sigma*isr) = e1^5, exp(gamma*sigma*isr) = e1^6
                const double e_2 = a.e1 * a.e1;
                a.e_4 = e_2 * e_2;
                a.g = a.e_4 * e_2;
                a.ux = tx * a.ir; a.uy = ty_ * a.ir; a.uz = tz * a.ir;
                return a;
            };
            // own bonds (all in the first 32 items) leave their record and pair energy
            auto own_bond = [&](const Item& a) {
                const uint32_t ty = a.d >> 21;
                if (a.ok && ty != 0u) {
                    const int r = (a.d >> 11) & 31;
                    q[r] = a.ux; q[RC2 + r] = a.uy; q[2 * RC2 + r] = a.uz; q[3 * RC2 + r] = a.g;
                    const double s2 = CK.ss * a.ir * a.ir;
                    const double pe = CK.aeps * (CK.bigb * (s2 * s2) - 1.0) * (a.e_4 * a.e1);
                    if (ty == 1u) ao += pe; else an += pe;
                }
            };
            // candidates close the j-centred triplets of both variants
            auto candidate = [&](const Item& a) {
                const uint32_t ra = (a.d >> 11) & 31u, rb = (a.d >> 16) & 31u;
                const bool cand = a.ok && (a.d >> 21) == 0u;
                const bool ho = cand && ra != IT_NONE, hn = cand && rb != IT_NONE;
                const int io = ho ? (int)ra : 0, in_ = hn ? (int)rb : 0;
                const double ex = CK.leps * a.g;
                const double cto = -(q[io] * a.ux + q[RC2 + io] * a.uy + q[2 * RC2 + io] * a.uz);
                const double d_o = cto - CK.cos0;
                const double vo = q[3 * RC2 + io] * ex * (d_o * d_o);
                if (ho && cto < CK.c099) ao += vo;                      // the cos < 0.99 filter of molint.F90:367-371
                const double ctn = -(q[in_] * a.ux + q[RC2 + in_] * a.uy + q[2 * RC2 + in_] * a.uz);
                const double dn = ctn - CK.cos0;
                const double vn = q[3 * RC2 + in_] * ex * (dn * dn);
                if (hn && ctn < CK.c099) an += vn;
            };
            if (ILP >= 2) {
                // as many passes in flight as the items left need, up to ILP (three: a dense walker's 65-96 items)
#pragma unroll 1
                for (int t0 = first; t0 < nitems;) {
                    const int left = nitems - t0;
                    if (ILP >= 3 && left > 64) {
                        const Item a = item_at(t0), b = item_at(t0 + 32), c = item_at(t0 + 64);
                        if (t0 == 0) { own_bond(a); __syncwarp(); }
                        candidate(a);
                        candidate(b);
                        candidate(c);
                        t0 += 96;
                    } else if (left > 32) {
                        const Item a = item_at(t0), b = item_at(t0 + 32);
                        if (t0 == 0) { own_bond(a); __syncwarp(); }
                        candidate(a);
                        candidate(b);
                        t0 += 64;
                    } else {
                        const Item a = item_at(t0);
                        if (t0 == 0) { own_bond(a); __syncwarp(); }
                        candidate(a);
                        t0 += 32;
                    }
                }
            } else {
#pragma unroll 1
                for (int t0 = first + 32 * sub; t0 < nitems; t0 += 32 * WPL) {
                    const Item a = item_at(t0);
                    if (t0 == 0) { own_bond(a); __syncwarp(); }
                    if (WPL == 2 && !synced) { lat_bar(bar); synced = true; }   // the records of pass 0 reach the other warp
                    candidate(a);
                }
            }
            __syncwarp();
            if (WPL == 2) {
                if (!synced) lat_bar(bar);                               // a warp without a pass in this round
                if (nrounds > 1) lat_bar(bar);                           // the item table is rewritten by the next round
            }
        }

        // ---- triplets centred on imol: all unordered pairs of bond records of one variant (rotation pairing:
        // record at position pos of a segment of n pairs with (pos + d) mod n, d = 1 .. n/2).  Up to 16 records
        // (nearly always) take two lanes each: lanes 0-15 the odd steps d, lanes 16-31 the even ones.
        if (WPL == 2) {
            // this warp: the records of variant `sub`; 4 / 2 / 1 lanes per record take the steps d = d0, d0 + 4 / 2 / 1, ...
            const int nv = sub ? nw : no, base = sub ? no : 0;
            const int sh = (nv <= 8) ? 3 : (nv <= 16) ? 4 : 5;
            const int stride = 32 >> sh;
            const int pos = lane & ((1 << sh) - 1);
            const bool act = pos < nv;
            const int n = act ? nv : 0;
            const int r = base + pos;
            const int half = n >> 1, send = base + n;
            const bool even = !(n & 1);
            const int rr = act ? r : 0;
            const double ux = q[rr], uy = q[RC2 + rr], uz = q[2 * RC2 + rr];
            const double g = act ? q[3 * RC2 + rr] : 0.0;
            const uint32_t jr = recj[rr];
            const int maxd = nv >> 1;
            double tb = 0.0;
#pragma unroll 1
            for (int d = 1 + (lane >> sh); d <= maxd; d += stride) {
                int c = r + d;
                c = (c >= send) ? c - n : c;
                const bool on = (d <= half) && !(even && d == half && pos >= half);
                c = on ? c : rr;
                const double ct = ux * q[c] + uy * q[RC2 + c] + uz * q[2 * RC2 + c];
                const double mult = (recj[c] == jr) ? 3.0 : 1.0;
                const double dd = ct - CK.cos0;
                if (on && ct < CK.c099) tb += q[3 * RC2 + c] * (dd * dd) * mult;
            }
            tb *= CK.leps * g;
            if (sub) an += tb; else ao += tb;
            if (npass == 2) lat_bar(bar);                                // the records are rewritten by the second pass
        } else {
            const int stride = (nown <= 16) ? 2 : 1;
            const int r = (stride == 2) ? (lane & 15) : lane;
            const bool act = r < nown;
            const bool sg = r >= no;
            const int n = act ? (sg ? nw : no) : 0, pos = sg ? r - no : r;
            const int half = n >> 1, send = r - pos + n;
            const bool even = !(n & 1);
            const int rr = act ? r : 0;
            const double ux = q[rr], uy = q[RC2 + rr], uz = q[2 * RC2 + rr];
            const double g = act ? q[3 * RC2 + rr] : 0.0;
            const uint32_t jr = recj[rr];
            const int maxd = max(no, nw) >> 1;
            double tb = 0.0;
            if (ILP >= 2) {
                // few walkers per SM: two steps d per turn, their loads and dot products independent; the terms are
                // added in the same order (a masked term adds +0.0: same bits)
                auto term = [&](int d) -> double {
                    int c = r + d;
                    c = (c >= send) ? c - n : c;
                    const bool on = (d <= half) && !(even && d == half && pos >= half);
                    c = on ? c : rr;
                    const double ct = ux * q[c] + uy * q[RC2 + c] + uz * q[2 * RC2 + c];
                    const double mult = (recj[c] == jr) ? 3.0 : 1.0;
                    const double dd = ct - CK.cos0;
                    const double v = q[3 * RC2 + c] * (dd * dd) * mult;
                    return (on && ct < CK.c099) ? v : 0.0;
                };
                int d = (stride == 2) ? 1 + (lane >> 4) : 1;
#pragma unroll 1
                for (; d + stride <= maxd; d += 2 * stride) {
                    const double t1 = term(d), t2 = term(d + stride);
                    tb += t1;
                    tb += t2;
                }
                if (d <= maxd) tb += term(d);
            } else {
#pragma unroll 1
                for (int d = (stride == 2) ? 1 + (lane >> 4) : 1; d <= maxd; d += stride) {
                    int c = r + d;
                    c = (c >= send) ? c - n : c;
                    const bool on = (d <= half) && !(even && d == half && pos >= half);
                    c = on ? c : rr;
                    const double ct = ux * q[c] + uy * q[RC2 + c] + uz * q[2 * RC2 + c];
                    const double mult = (recj[c] == jr) ? 3.0 : 1.0;
                    const double dd = ct - CK.cos0;
                    if (on && ct < CK.c099) tb += q[3 * RC2 + c] * (dd * dd) * mult;
                }
            }
            tb *= CK.leps * g;
            if (sg) an += tb; else ao += tb;
        }
        __syncwarp();
    }
    reduce2(ao, an);
    eo = ao; en = an;
    return err;
}

// ---------------------------------------------------------------- move generation, GB moves per call (warp 0)
// mc_moves.F90:1001-1067 in the reference's exact arithmetic.  Lane = 4*move + role; every lane of a move's quad
// derives the direction and the magnitude (identical instructions for all moves of the batch); roles 0 / 1 apply the
// fractional transform for "lattice 1 active" / "lattice 2 active", role 2 stores the plain displacement and the
// molecule.  D = draws per translation move incl. acceptance and switch.  Returns the index (0..nb) of the first
// move of the batch that is NOT a translation.
template <int NLAT, int NT>
__device__ __forceinline__ int generate_moves(const Lay<NT> Y, unsigned char* smem, const McParams& p, int pos, int D, int nb)
{
    const int N = Y.N(), lane = lane_id(), m = lane >> 2, role = lane & 3;
    unsigned char* sb = smem + NLAT * Y.LB();
    const bool act = m < nb;
    const double* u = at<double>(sb, Lay<NT>::sRNG) + pos + (act ? m : 0) * D;
    const double xi = u[0];
    int imol = (int)xm(u[1], (double)N) + 1;
    if (imol > N) imol = N;
    imol -= 1;
    double vx = xs(xm(2.0, u[2]), 1.0), vy = xs(xm(2.0, u[3]), 1.0), vz = xs(xm(2.0, u[4]), 1.0);
    const double norm = xd(1.0, xsqrt(xa(xa(xm(vx, vx), xm(vy, vy)), xm(vz, vz))));
    vx = xm(vx, norm); vy = xm(vy, norm); vz = xm(vz, norm);
    const double r = xs(xm(u[5], 2.0), 1.0);
    const double mt = at<WalkerScalars>(sb, Lay<NT>::sSC)->max_trans;
    vx = xm(xm(vx, mt), r); vy = xm(xm(vy, mt), r); vz = xm(xm(vz, mt), r);
    double* rec = at<double>(sb, Lay<NT>::sGEN) + (act ? m : 0) * GF;
    if (NLAT == 2 && role < 2) {
        // role 0: lattice 1 active -> image of the displacement in lattice 2: recip(1), hmatrix(2); role 1: the reverse
        const double* rm = at<double>(smem + (role == 0 ? 0 : Y.LB()), Y.oR());
        const double* hm = at<double>(smem + (role == 0 ? Y.LB() : 0), Y.oH());
        double sx = xa(xa(xm(MW_H(rm,1,1), vx), xm(MW_H(rm,2,1), vy)), xm(MW_H(rm,3,1), vz));
        double sy = xa(xa(xm(MW_H(rm,1,2), vx), xm(MW_H(rm,2,2), vy)), xm(MW_H(rm,3,2), vz));
        double sz = xa(xa(xm(MW_H(rm,1,3), vx), xm(MW_H(rm,2,3), vy)), xm(MW_H(rm,3,3), vz));
        sx = xm(xm(sx, 0.5), INV_PI); sy = xm(xm(sy, 0.5), INV_PI); sz = xm(xm(sz, 0.5), INV_PI);
        const double bx = xa(xa(xm(MW_H(hm,1,1), sx), xm(MW_H(hm,1,2), sy)), xm(MW_H(hm,1,3), sz));
        const double by = xa(xa(xm(MW_H(hm,2,1), sx), xm(MW_H(hm,2,2), sy)), xm(MW_H(hm,2,3), sz));
        const double bz = xa(xa(xm(MW_H(hm,3,1), sx), xm(MW_H(hm,3,2), sy)), xm(MW_H(hm,3,3), sz));
        if (act) { rec[3 + 3 * role] = bx; rec[4 + 3 * role] = by; rec[5 + 3 * role] = bz; }
    } else if (role == 2 && act) {
        rec[0] = vx; rec[1] = vy; rec[2] = vz;
        at<int>(sb, Lay<NT>::sGI)[m] = imol;
    }
    const uint32_t rare = __ballot_sync(FULL, act && role == 0 && !(xi < p.transP));
    return rare ? ((__ffs(rare) - 1) >> 2) : nb;
}

// ---------------------------------------------------------------- one turn of a walker on a block
// NLAT warps; loads the walker's image, runs the move loop of mc_cycle (mc_moves.F90:160-260) and stores the image
// back.  `first` = the walker's first turn of this launch (fixes the cycle at which its part of the launch ends).
// With chunk > 0 the walker may give up the block every `chunk` cycles: it does when it is not behind the average
// progress of the batch (qctr[4..5] = cycles completed by all walkers), so that slow walkers -- more bonds per
// molecule -- hold their blocks longer and all walkers reach the end of the launch together.
// Returns (uniform over the block) whether the walker has cycles left in this launch.
template <int NLAT, int NT, int WPL, int ILP>
__device__ __forceinline__ bool run_walker(const DeviceState& S, const McParams& p, unsigned char* smem, int wi, int ncycles_launch,
                                           int chunk, bool first)
{
    static_assert(WPL == 1 || (WPL == 2 && NLAT == 2), "two warps per lattice: lattice-switch boxes only");
    static_assert(ILP == 1 || WPL == 1, "several passes in flight: one warp per lattice");
    constexpr int NTHR = 32 * NLAT * WPL;
    const Lay<NT> Y(S.N);
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = Y.N();
    unsigned char* sb = smem + NLAT * Y.LB();
    load_walker<NT>(Y, S, wi, smem, tid, NTHR);
    // Which warp takes lattice 1 -- and with it the serial acceptance -- alternates with the hardware warp slot:
    // the two warps of a walker sit on neighbouring schedulers (slot % 4), and with a fixed assignment every
    // scheduler pair would carry all its acceptance warps on one side (measured: 77 % / 46 % issue-active).
    if (NLAT == 2 && tid == 0) {
        unsigned wslot;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
        *at<int>(sb, Lay<NT>::sSWAP) = (int)((wslot >> 2) & 1u);
    }
    __syncthreads();
    const int sub = (WPL == 2) ? ((tid >> 5) & 1) : 0;                 // WPL == 2: warps 2L, 2L+1 share lattice L
    const int lat = (NLAT == 2) ? ((tid >> (WPL == 2 ? 6 : 5)) ^ *at<int>(sb, Lay<NT>::sSWAP)) : 0;
    const bool prim = (sub == 0);                                       // the lattice's warp that commits and rebuilds
    const bool accw = (lat == 0) && prim;                               // the warp that carries the serial acceptance
    const int lbar = 1 + lat;                                           // named barrier of the lattice's two warps
    unsigned char* lb = smem + lat * Y.LB();
    WalkerScalars* sc = at<WalkerScalars>(sb, Lay<NT>::sSC);
    uint64_t* rngbase = at<uint64_t>(sb, Lay<NT>::sRB);
    double* rngbuf = at<double>(sb, Lay<NT>::sRNG);
    double* lv = at<double>(sb, Lay<NT>::sLV);
    double* kv = at<double>(sb, Lay<NT>::sKV);
    int* ctl = at<int>(sb, Lay<NT>::sCTL);
    const double Nd = (double)N;
    double* wgt = S.weight + (size_t)wi * S.NB;
    double* hist = S.hist + (size_t)wi * S.NB;
    double* uhist = S.uhist + (size_t)wi * S.NB;
    double* P = at<double>(lb, Y.oP());
    double* T = at<double>(lb, Y.oT());
    int err = 0;
    const int cycle0 = sc->cycle;
    const int cyc_end = first ? cycle0 + ncycles_launch : S.cyc_end[wi];
    if (tid == 0 && first) { S.cyc_end[wi] = cyc_end; S.wtime[2 * wi] = globaltimer_ns(); }
    const int ncycles = cyc_end - cycle0;
    int rng_pos = 0;

    if (prim) compute_bond_masks<NT>(Y, lb);
    if (accw) {
        if (p.prob_error) err |= ERR_PROB;
        const uint64_t idx = sc->rng_index;
        __syncwarp();
        if (lane == 0) {
            *rngbase = idx & ~(uint64_t)1;
            if (NLAT == 2) {
                lv[0] = log(sc->vol[0] / sc->vol[1]); lv[1] = log(sc->vol[1] / sc->vol[0]);
                refresh_kv(p, sc, lv, kv, Nd);
            }
            ctl[CTL_STOP] = (err & ERR_PROB) ? 1 : 0;
        }
        rng_pos = (int)(idx & 1);
        __syncwarp();
    }
    __syncthreads();

    bool stop = ctl[CTL_STOP] != 0;
    int bpar = 0;                                                       // batch parity (both warps count alike)
    int klast = p.nbins / 2 + 1;                                        // bin of the previous move's order parameter
    for (int cyc = 0; cyc < ncycles && !stop; ++cyc) {
        const int cycle = cycle0 + cyc + 1;
        if (accw) {
            __syncwarp();
            if (lane == 0) {
                sc->cycle = cycle;
                if (p.dd) {                                            // mc_moves.F90:181-210
                    if (cycle < p.eq_mc_cycles) sc->in_window = (sc->mu > sc->mu_lo) && (sc->mu < sc->mu_hi);
                    else if (cycle == p.eq_mc_cycles) { if (!sc->in_window) ctl[CTL_STOP] = 1; }
                    else sc->in_window = 1;
                }
            }
            __syncwarp();
            if (ctl[CTL_STOP]) err |= ERR_WINDOW;
        }
        if (cycle % p.list_update_int == 0) {                          // :218-222, each lattice by its (first) warp
            if (prim) {
                err |= compute_neighbours<NT>(Y, lb);
                compute_bond_masks<NT>(Y, lb);
            }
            if (WPL == 2) lat_bar(lbar);
        }
        const bool dd_eq = p.dd && (cycle < p.eq_mc_cycles);
        const bool bins_on = !(cycle < p.eq_mc_cycles);                // mc_update_wl_bins: :1615
        const bool do_switch = (NLAT == 2) && p.always_switch && !dd_eq;
        const bool fuse_switch = do_switch && p.samplerun;             // weights fixed: eta of the switch is already known
        const int D = 7 + (do_switch ? 1 : 0);                         // draws per translation move

        int imove = 0;
        while (imove < N) {                                            // :224-250, GB moves per batch
            if (accw) {
                // every batch starts with a refill at the current draw index (the buffer starts at an even index)
                const uint64_t next = *rngbase + (uint64_t)rng_pos;
                __syncwarp();
                if (lane == 0) *rngbase = next & ~(uint64_t)1;
                __syncwarp();
                rng_pos = (int)(next & 1);
                rng_refill_at(rngbase, rngbuf, S, p, wi);
                const int nb = min(GB, N - imove);
                const int nr = generate_moves<NLAT, NT>(Y, smem, p, rng_pos, D, nb);
                if (lane == 0) { ctl[bpar * 3] = nb; ctl[bpar * 3 + 1] = nr; ctl[bpar * 3 + 2] = (sc->ls == 1); }
            }
            __syncthreads();
            if (ctl[CTL_STOP]) { stop = true; break; }
            const int nb = ctl[bpar * 3], nr = ctl[bpar * 3 + 1];
            bool one = ctl[bpar * 3 + 2] != 0;
            bpar ^= 1;

            for (int m = 0; m < nr; ++m) {
                // ====================== mc_water_translation (mc_moves.F90:966-1213) ======================
                const int imol = at<int>(sb, Lay<NT>::sGI)[m];
                {
                    // displacement of my lattice: the plain one when it is the active lattice, else its image
                    const int off = (NLAT == 1) ? 0 : (lat == 0) ? (one ? 0 : 6) : (one ? 3 : 0);
                    if (lane < 3) {
                        const double tv = at<double>(sb, Lay<NT>::sGEN)[m * GF + off + lane];
                        T[lane] = xa(P[lane * N + imol], tv);
                        T[3 + lane] = tv;
                    }
                    __syncwarp();
                }
                double eo, en;
                uint32_t mo, mn;
                err |= local_energies<NT, WPL, ILP>(Y, lb, imol, true, sub, lbar, eo, en, mo, mn);
                double* xch = at<double>(sb, Lay<NT>::sXCH);
                if (NLAT == 2) {
                    if (lane == 0) { xch[(lat * WPL + sub) * 2] = eo; xch[(lat * WPL + sub) * 2 + 1] = en; }
                    __syncthreads();                                    // A: both lattices' energies
                }
                if (accw) {
                    if (WPL == 2) { eo = xch[0] + xch[2]; en = xch[1] + xch[3]; }     // the two warps' parts
                    if (lane == 0) atomicAdd(S.transcount + (size_t)wi * N + imol, 1);
                    const double* u = rngbuf + rng_pos + m * D;
                    // model_energy bookkeeping exactly as :1013-1016, :1087-1090
                    const double Eb0 = sc->E[0], Ea0 = (Eb0 - eo) + en, dE0 = en - eo;
                    double Eb1 = 0.0, Ea1 = 0.0, dE1 = 0.0;
                    const double mu_old = sc->mu;
                    double diffkT, mu_acc = mu_old, mu_rej = mu_old, eta_acc = 0.0, eta_rej = 0.0;
                    int k_acc = 0, k_rej = 0;
                    double cinc = 0.0;       // histogram increment of the lane's bin, requested as soon as the bin is known
                    if (NLAT == 1) {
                        diffkT = p.beta * dE0;
                        if (bins_on) {       // single box: ls_mu is never assigned (0) but the bins are still updated
                            const EtaBin eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, mu_old);
                            eta_acc = eta_rej = eb.eta; k_acc = k_rej = eb.k;
                            if (eb.k >= 1 && eb.k <= p.nbins) cinc = __ldg(S.hinc + eb.k - 1);
                        }
                    } else {
                        const double eo1 = (WPL == 2) ? xch[4] + xch[6] : xch[2], en1 = (WPL == 2) ? xch[5] + xch[7] : xch[3];
                        Eb1 = sc->E[1]; Ea1 = (Eb1 - eo1) + en1; dE1 = en1 - eo1;
                        const double dm = (dE0 - dE1) * p.beta;
                        mu_acc = mu_old + dm;                           // :1113
                        mu_rej = mu_acc - dm;                           // :1195 -- (mu + d) - d, not a copy
                        // three weight look-ups in parallel lanes: eta(mu), eta(mu_acc), eta(mu_rej)
                        const double mine = (lane == 0) ? mu_old : (lane == 1) ? mu_acc : mu_rej;
                        EtaBin eb; eb.eta = 0.0; eb.k = 0;
                        if (lane < 3) {
                            eb = eta_bin_near(p, S, sc, wgt, mine, klast);
                            if (bins_on && eb.k >= 1 && eb.k <= p.nbins) cinc = __ldg(S.hinc + eb.k - 1);
                        }
                        const double eta_old = __shfl_sync(FULL, eb.eta, 0);
                        eta_acc = __shfl_sync(FULL, eb.eta, 1); eta_rej = __shfl_sync(FULL, eb.eta, 2);
                        k_acc = __shfl_sync(FULL, eb.k, 1); k_rej = __shfl_sync(FULL, eb.k, 2);
                        diffkT = (one ? dE0 : dE1) * p.beta + eta_acc - eta_old;
                    }
                    // one exponential pass: lane 0 acceptance, lanes 1/2 switch probability if accepted / rejected,
                    // lanes 3/4 unbiased-histogram factor if accepted / rejected
                    double arg = -diffkT;
                    if (fuse_switch && (lane == 1 || lane == 2)) {
                        // -diffkT of mc_lattice_switch (:1562-1574) for the energies after an accepted / a rejected move
                        const double dEs = (lane == 1) ? (Ea1 - Ea0) : (Eb1 - Eb0);     // E(2) - E(1)
                        arg = one ? -(p.beta * dEs + kv[0]) : (p.beta * dEs - kv[1]);
                    }
                    if (lane == 3) arg = eta_acc - p.log_unbiased_norm;
                    if (lane == 4) arg = eta_rej - p.log_unbiased_norm;
                    const double ex = (arg > 0.0 && lane < 3) ? 1.0 : exp_fast(fmin(arg, 700.0));
                    const bool accepted = u[6] < __shfl_sync(FULL, ex, 0);                      // :1145-1146
                    const double En0 = accepted ? Ea0 : Eb0, En1 = accepted ? Ea1 : Eb1;
                    bool sw = false;
                    double mu_new = accepted ? mu_acc : mu_rej;
                    if (fuse_switch) {
                        // ====================== mc_lattice_switch (mc_moves.F90:1536-1594) ======================
                        sw = u[7] < __shfl_sync(FULL, ex, accepted ? 1 : 2);
                        if (sw) mu_new = p.beta * (En0 - En1) + kv[2];          // ls_mu from scratch (:1583-1585)
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (accepted) {
                            sc->acc_r += 1;
                            const double dmu = fabs(dE0 - dE1) * p.beta;
                            if (dmu < sc->min_dmu) sc->min_dmu = dmu;
                            if (dmu > sc->max_dmu) sc->max_dmu = dmu;
                            sc->E[0] = Ea0;
                            if (NLAT == 2) sc->E[1] = Ea1;
                        }
                        if (NLAT == 2) sc->mu = mu_new;
                        sc->att_r += 1;
                        if (fuse_switch) {
                            if (sw) { sc->acc_s += 1; sc->ls = 3 - sc->ls; }
                            sc->att_s += 1;
                        }
                    }
                    __syncwarp();
                    // ====================== mc_update_wl_bins (mc_moves.F90:1597-1689) ======================
                    // (the bin of the order parameter BEFORE the switch of this move, as in the reference's order)
                    const int kb = accepted ? k_acc : k_rej;
                    klast = kb;
                    const double c = (NLAT == 1) ? cinc : __shfl_sync(FULL, cinc, accepted ? 1 : 2);
                    if (bins_on && kb >= 1 && kb <= p.nbins) {
                        if (lane == 0) atomicAdd(hist + kb - 1, c);
                        if (p.samplerun) {
                            const double uf = __shfl_sync(FULL, ex, accepted ? 3 : 4);
                            if (lane == 0) atomicAdd(uhist + kb - 1, c * uf);
                        } else {
                            update_weights_at(sc, N, p, S.binwidth, wgt, hist, kb);
                        }
                    }
                    if (do_switch && !fuse_switch) {
                        // weights may have moved in update_weights: the reference looks eta up again
                        lattice_switch_at(sc, lv, rngbuf, S, p, wi, rng_pos + m * D + 7);
                    }
                    if (lane == 0) ctl[CTL_DEC] = (accepted ? 1 : 0) | (sc->ls == 1 ? 2 : 0);
                    __syncwarp();
                }
                if (NLAT == 2) __syncthreads();                         // B: the decision
                const int dec = ctl[CTL_DEC];
                one = (dec & 2) != 0;
                if (!prim) {
                    // the lattice's first warp commits / restores
                } else if (dec & 1) {
                    // commit: new position, own bond mask, and the reverse bits of the bonds that formed / broke
                    uint32_t* BM = at<uint32_t>(lb, Y.oBM());
                    if (lane < 3) P[lane * N + imol] = T[lane];
                    if (lane == 0) BM[imol] = mn;
                    const uint32_t changed = mo ^ mn;
                    if ((changed >> lane) & 1u) {
                        const uint32_t e = at<uint16_t>(lb, Y.oL())[imol * LC + lane];
                        const uint32_t bit = 1u << ((e >> 6) & 31u);
                        if ((mn >> lane) & 1u) atomicOr(BM + (e & 63u), bit); else atomicAnd(BM + (e & 63u), ~bit);
                    }
                } else if (lane < 3) {
                    // reject: the reference restores by (x+t)-t, not by copy (mc_moves.F90:1186)
                    P[lane * N + imol] = xs(T[lane], T[3 + lane]);
                }
                __syncwarp();
                if (WPL == 2) lat_bar(lbar);
            }
            imove += nr;
            if (accw) rng_pos += nr * D;
            if (nr < nb) {
                // ---------------- rare move types (warp 0; warp 1 waits at the next batch barrier) ----------------
                // a volume move works on both lattice blocks: warp 1 must be through with its commit / restore
                if (NLAT == 2) __syncthreads();
                if (accw) {
                    const double xi = rngbuf[rng_pos];
                    rng_pos += 1;
                    if (xi < p.volP) {
                        rng_pos = volume_move<NLAT, NT>(Y, smem, S, p, wi, rng_pos);
                        if (NLAT == 2) { if (lane == 0) refresh_kv(p, sc, lv, kv, Nd); __syncwarp(); }
                        const EtaBin eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, sc->mu);
                        if (bins_on && eb.k >= 1 && eb.k <= p.nbins) {
                            const double c = __ldg(S.hinc + eb.k - 1);
                            if (lane == 0) atomicAdd(hist + eb.k - 1, c);
                            if (p.samplerun) {
                                if (lane == 0) atomicAdd(uhist + eb.k - 1, c * exp(eb.eta - p.log_unbiased_norm));
                            } else {
                                update_weights_at(sc, N, p, S.binwidth, wgt, hist, eb.k);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) sc->att_v += 1;
                        __syncwarp();
                    } else if (xi < p.swP) {
                        if (NLAT == 2 && !dd_eq) rng_pos = lattice_switch_at(sc, lv, rngbuf, S, p, wi, rng_pos);
                    }
                    if (do_switch) rng_pos = lattice_switch_at(sc, lv, rngbuf, S, p, wi, rng_pos);
                }
                if (NLAT == 2) __syncthreads();          // ... and must not start its next list rebuild before the move is over
                imove += 1;
            }
        }
        if (stop) break;
        if (accw) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int l = 0; l < NLAT; ++l) {                       // :253-255
                    double a = sc->avgE[l] + sc->E[l];
                    if (p.npt) a = a + p.pressure * sc->vol[l];
                    sc->avgE[l] = a;
                }
            }
            __syncwarp();
            if (S.therm_int > 0 && cycle % S.therm_int == 0) {         // main.f90:200-223 (values only)
                const int n = S.therm_n[wi];
                __syncwarp();
                if (n < S.therm_cap && lane < THERM_ROW) {
                    double v;
                    switch (lane) {
                    case 0: v = (double)cycle; break;
                    case 1: v = (double)sc->ls; break;
                    case 2: v = sc->E[0]; break;
                    case 3: v = sc->E[1]; break;
                    case 4: v = sc->mu; break;
                    case 5: v = sc->vol[0]; break;
                    case 6: v = sc->vol[1]; break;
                    default: v = at<double>(smem, Y.oH())[lane - 7]; break;
                    }
                    S.therm[((size_t)wi * S.therm_cap + n) * THERM_ROW + lane] = v;
                }
                __syncwarp();
                if (lane == 0) S.therm_n[wi] = n + 1;
            }
        }
        if (chunk > 0 && (cyc + 1) % chunk == 0 && cyc + 1 < ncycles) {
            // end of a unit: keep the block while this walker is behind the average progress of the batch
            __syncthreads();
            if (tid == 0) {
                const long long tot = (long long)atomicAdd((unsigned long long*)(S.qctr + 4), (unsigned long long)chunk) + chunk;
                const long long mine = cycle - (cyc_end - ncycles_launch);
                ctl[CTL_DEC] = (mine * S.W >= tot) ? 1 : 0;
            }
            __syncthreads();
            if (ctl[CTL_DEC]) break;
        }
    }
    __syncthreads();
    if (accw && lane == 0) {
        const uint64_t idx = *rngbase + (uint64_t)rng_pos;
        if (p.rng_mode == 1 && idx > S.fifo_len) err |= ERR_RNG_UNDERRUN;
        sc->rng_index = idx;
    }
    err = (int)__reduce_or_sync(FULL, (unsigned)err);
    if (lane == 0 && err) atomicOr(&sc->error, err);
    __syncthreads();
    const bool more = !ctl[CTL_STOP] && sc->cycle < cyc_end;
    store_walker<NT>(Y, S, wi, smem, tid, NTHR);
    if (tid == 0 && !more) S.wtime[2 * wi + 1] = globaltimer_ns();
    return more;
}

// ---------------------------------------------------------------- the kernel: persistent blocks of NLAT warps
// A batch larger than the resident blocks of the GPU would run in waves of whole walkers and end on its slowest
// one: measured on 4096 walkers, 2072 resident, a 250-cycle launch spends 40 of 191 ms below 90 % residency (walkers
// differ by +-10 % in bonds per molecule, a few by 50 %; scripts/diag/tail_probe.py).  Walkers are independent
// (the reference runs one per MPI rank), so the launch is cut into units of `chunk` cycles and the resident blocks
// take walkers from a queue: entry q < W is walker q's first turn; at the end of a unit a walker that is not behind
// the average progress of the batch gives up its block: the block publishes it at the tail of the queue (release)
// and the block that takes that entry acquires it.  All walkers advance at the same rate in cycles -- the slow
// ones simply hold a block for a larger share of the time -- and the tail of the launch is about one unit long.
// A launch cut into turns is the same computation as consecutive launches (the state of a walker between cycles is
// its stored image), so the results do not depend on chunk or on the number of blocks.
// BL = resident walkers per SM the register allocation is bounded for: MW2_BLOCKS (72 registers) when the batch
// fills the GPU, MW2_BLOCKS / 2 (no spills, 122 registers) for small ensembles, where a step lasts as long as one
// walker's chain and more registers shorten it by 8 % (profiles/README.md).  Same PTX, same results.
template <int NLAT, int NT, int BL, int WPL, int ILP>
__global__ void __launch_bounds__(32 * NLAT * WPL, BL * (3 - NLAT)) k_mc_run2(const __grid_constant__ DeviceState S,
                                                                         const __grid_constant__ McParams p, int ncycles, int chunk)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, W = S.W;
    int* s_unit = at<int>(smem + NLAT * Lay<NT>(S.N).LB(), Lay<NT>::sUNIT);
    const bool single = chunk >= ncycles;                   // one unit per walker: nothing is ever published
    for (;;) {
        if (tid == 0) {
            int q = atomicAdd(S.qctr, 1), wi = -1;
            if (q < W) wi = q;
            else if (!single) {
                const int* slot = S.queue + (q - W);
                for (;;) {
                    int v;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
                    if (v) { wi = v - 1; break; }
                    int done;
                    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(S.qctr + 2) : "memory");
                    if (done >= W) break;                   // every walker finished: no entry will follow
                    __nanosleep(500);
                }
            }
            *s_unit = (wi < 0) ? -1 : (wi | (q < W ? (1 << 30) : 0));
        }
        __syncthreads();
        const int u = *s_unit;
        if (u < 0) return;
        const int wi = u & ((1 << 30) - 1);
        const bool more = run_walker<NLAT, NT, WPL, ILP>(S, p, smem, wi, ncycles, single ? 0 : chunk, (u >> 30) != 0);
        __threadfence();
        __syncthreads();                                    // the image is stored; s_unit and smem may be reused
        if (tid == 0) {
            if (more) {
                const int t = atomicAdd(S.qctr + 1, 1);
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(S.queue + t), "r"(wi + 1) : "memory");
            } else {
                __threadfence();
                atomicAdd(S.qctr + 2, 1);
            }
        }
    }
}

}  // namespace v2
}  // namespace mw
