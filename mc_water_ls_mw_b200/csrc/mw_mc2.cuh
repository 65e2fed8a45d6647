// mw_mc2.cuh -- the walker kernel with TWO warps per walker (one per lattice), for two-lattice boxes of up to
// 64 molecules (every lattice-switch deck of the reference has 48).
//
// The reference marks this parallelism itself: the `do ils = 1,num_lattices` loops around the local-energy calls
// of a translation (mc_moves.F90:1007-1018, :1076-1090) are independent per lattice.  Warp L of a walker's CTA owns
// lattice L: its positions, Verlet rows, image vectors and bond masks are read and written by that warp only.
// Per trial move the two warps meet twice:
//     barrier A   both local-energy pairs (old, new) are in shared memory
//     warp 0      acceptance, weights, histograms, lattice switch          (mc_moves.F90:1104-1213, :1536-1689)
//     barrier B   the decision (accepted, active lattice) is in shared memory; each warp commits / restores its lattice
// Move generation (mc_moves.F90:1001-1067: molecule, direction, magnitude, both fractional transforms) is done for
// GB moves at a time with the moves spread over the lanes of warp 0, in the reference's exact arithmetic, so the
// ~200 lane-uniform instructions it used to cost per move are paid once per batch.
//
// The local energies of one lattice are evaluated as ONE list of items (mw_device.cuh explains the formulation):
//     own bonds   (imol at its old / trial position, list slot)        -> pair energy + bond record (u, g)
//     candidates  (neighbour j, in-range list slot of j not pointing back at imol) -> j-centred triplets of BOTH variants
// every item is "a list slot of a row, seen from a centre position", so one inlined copy of the geometry, the
// radial functions and the exponential serves both kinds; the i-centred triplets pair the bond records afterwards.
#pragma once
#include "mw_mc.cuh"

#ifndef MW_MC2_BLOCKS
#define MW_MC2_BLOCKS 14     // resident walkers (2-warp CTAs) per SM the register allocation is bounded for
#endif

namespace mw {

constexpr int IT2 = 96;      // items per lattice and round (own bonds first, then candidates)
constexpr int RC2 = 32;      // bond records per lattice (old + new bonds of the moved molecule; > 31 -> ERR_BOND_OVERFLOW)
constexpr int GB  = 7;       // trial moves generated per batch (7 x 8 draws fit one 64-number refill from any parity)
constexpr int GF  = 9;       // doubles per generated move: displacement d, and its image in the other lattice when
                             // lattice 1 / lattice 2 is the active one

// item descriptor: slot | row << 5 | ra << 11 | rb << 16 | type << 21
//   own bond : ra = its record, type 1 (old position) / 2 (trial position)
//   candidate: ra / rb = record of the centre's bond to imol in the old / new variant (31 = not bonded), type 0
constexpr uint32_t IT_OLD = 1u << 21, IT_NEW = 2u << 21, IT_NONE = 31u;

// extra views into the walker's shared-memory image: the tables of the one-warp kernel that this kernel does not
// use (cq, cpre, cand; qmeta + cmeta; save) hold its own scratch, so both kernels share one layout and all cold paths
struct W2 {
    uint32_t* items;    // [2][IT2]
    uint8_t*  recj;     // [2][RC2]  molecule of a bond record (images of one molecule: factor 3, mw_device.cuh)
    int*      gimol;    // [GB]
    int*      ctl;      // [8] per batch parity: {length, first rare move, active lattice is 1} | stop flag | decision
    double*   gen;      // [GB][GF]
    double*   xch;      // [2][2]    local energies (old, new) of the two lattices
};
static_assert(sizeof(uint16_t) * (CC * 2 + CC + 2 + KC) >= sizeof(uint32_t) * 2 * IT2 + 2 * RC2 + 4 * GB + 32 + 4, "scratch (cq..cand)");
constexpr int CTL_STOP = 6, CTL_DEC = 7;
static_assert(sizeof(uint32_t) * (QC + CC) >= sizeof(double) * GB * GF, "scratch (qmeta, cmeta)");

__device__ __forceinline__ W2 carve2(const WalkerView& w)
{
    W2 x;
    unsigned char* p = (unsigned char*)w.cq;            // 16-byte aligned block: cq | cpre | cand
    x.items = (uint32_t*)p;            p += sizeof(uint32_t) * 2 * IT2;
    x.ctl   = (int*)p;                 p += 32;
    x.gimol = (int*)p;                 p += 4 * ((GB + 1) & ~1);
    x.recj  = (uint8_t*)p;
    x.gen   = (double*)w.qmeta;                         // qmeta | cmeta (16-byte aligned, 8-byte multiple)
    x.xch   = w.save;                                   // volume-move scratch, idle during translations
    return x;
}

// bar.sync over the two warps of a walker
__device__ __forceinline__ void pair_sync() { __syncthreads(); }

// ---------------------------------------------------------------- staging with all threads of the CTA
__device__ __forceinline__ void load_walker_cta(const DeviceState& S, int wi, const WalkerView& w, int tid, int nt)
{
    const int N = S.N, nlat = S.nlat;
    const double* gp = S.pos + (size_t)wi * nlat * 3 * N;
    for (int t = tid; t < nlat * 3 * N; t += nt) w.pos[t] = gp[t];
    const double* gi = S.iv + (size_t)wi * nlat * 3 * IVC;
    for (int t = tid; t < nlat * 3 * IVC; t += nt) w.iv[t] = gi[t];
    if (tid < nlat * 9) {
        w.cell[tid] = S.cell[(size_t)wi * nlat * 9 + tid];
        w.recip[tid] = S.recip[(size_t)wi * nlat * 9 + tid];
    }
    if (tid < 2) w.niv[tid] = S.niv[wi * 2 + tid];
    const uint4* gl = (const uint4*)(S.list + (size_t)wi * nlat * N * LC);
    uint4* sl = (uint4*)w.list;
    for (int t = tid; t < nlat * N * LC / 8; t += nt) sl[t] = gl[t];
    const uint8_t* gn = S.nn + (size_t)wi * nlat * N;
    for (int t = tid; t < nlat * N; t += nt) w.nn[t] = gn[t];
    const uint32_t* gs = (const uint32_t*)(S.scal + wi);
    uint32_t* ss = (uint32_t*)w.sc;
    for (int t = tid; t < (int)(sizeof(WalkerScalars) / 4); t += nt) ss[t] = gs[t];
}

__device__ __forceinline__ void store_walker_cta(const DeviceState& S, int wi, const WalkerView& w, int tid, int nt)
{
    const int N = S.N, nlat = S.nlat;
    double* gp = S.pos + (size_t)wi * nlat * 3 * N;
    for (int t = tid; t < nlat * 3 * N; t += nt) gp[t] = w.pos[t];
    double* gi = S.iv + (size_t)wi * nlat * 3 * IVC;
    for (int t = tid; t < nlat * 3 * IVC; t += nt) gi[t] = w.iv[t];
    if (tid < nlat * 9) {
        S.cell[(size_t)wi * nlat * 9 + tid] = w.cell[tid];
        S.recip[(size_t)wi * nlat * 9 + tid] = w.recip[tid];
    }
    if (tid < 2) S.niv[wi * 2 + tid] = w.niv[tid];
    uint4* gl = (uint4*)(S.list + (size_t)wi * nlat * N * LC);
    const uint4* sl = (const uint4*)w.list;
    for (int t = tid; t < nlat * N * LC / 8; t += nt) gl[t] = sl[t];
    uint8_t* gn = S.nn + (size_t)wi * nlat * N;
    for (int t = tid; t < nlat * N; t += nt) gn[t] = w.nn[t];
    uint32_t* gs = (uint32_t*)(S.scal + wi);
    const uint32_t* ss = (const uint32_t*)w.sc;
    for (int t = tid; t < (int)(sizeof(WalkerScalars) / 4); t += nt) gs[t] = ss[t];
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

// ---------------------------------------------------------------- local energies of one lattice (one warp)
// compute_local_real_energy(imol, lat) at the old position and at the trial position w.mv[lat*6 .. +2]
// (molint.F90:220-404; mc_moves.F90:1010,1083).  Returns the two energies (uniform), the in-range slot masks
// of imol's row for both positions, and error bits.
template <int NT>
__device__ __forceinline__ int local_energies_lat(const WalkerView& w, const W2& x, int lat, int imol,
                                                  double& eo, double& en, uint32_t& mo, uint32_t& mn)
{
    const int N = (NT > 0) ? NT : w.N, lane = lane_id();
    const unsigned lt = lt_mask();
    const double* P = w.pos + lat * 3 * N;
    const double* V = w.iv + lat * 3 * IVC;
    const uint16_t* L = w.list + (size_t)lat * N * LC;
    const uint32_t* BM = w.bmask + lat * N;
    const double* T = w.mv + lat * 6;
    double* q = w.q + lat * RC2;                        // records [4][QC], columns RC2*lat ..
    uint32_t* items = x.items + lat * IT2;
    uint8_t* recj = x.recj + lat * RC2;
    int err = 0;

    // ---- stage 1: lanes = slots of imol's row.  Distance tests at both positions, bond records, centres
    const int nni = w.nn[lat * N + imol];
    const bool has = lane < nni;
    const uint32_t e = has ? L[imol * LC + lane] : 0u;
    const int j = e & 63, rev = (e >> 6) & 31, img = e >> 11;
    uint32_t bo, bn;
    {
        const double pjx = P[j] + V[img], pjy = P[N + j] + V[IVC + img], pjz = P[2 * N + j] + V[2 * IVC + img];
        const double r2o = dist2(pjx - P[imol], pjy - P[N + imol], pjz - P[2 * N + imol]);
        const double r2n = dist2(pjx - T[0], pjy - T[1], pjz - T[2]);
        bo = __ballot_sync(FULL, has && r2o < CK.rcsq);
        bn = __ballot_sync(FULL, has && r2n < CK.rcsq);
    }
    mo = bo; mn = bn;
    // slots of row j that point back at imol (any image) are no candidates: the k == i entries of the reference's
    // list B are either filtered (cos = 1) or covered by the factor 3 of the i-centred pairs
    // (cells narrower than twice the list radius hold two images of one molecule in most rows; a row is sorted by
    // molecule, so the images are neighbouring lanes)
    uint32_t excl = 1u << rev;
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
    const uint32_t bop = (npass == 1 || pass == 0) ? bo : 0u, bnp = (npass == 1 || pass == 1) ? bn : 0u;
    const bool fo = (bop >> lane) & 1u, fn = (bnp >> lane) & 1u;
    const int no = __popc(bop), nw = __popc(bnp), nown = no + nw;
    const int ro = __popc(bop & lt), rn = no + __popc(bnp & lt);
    if (nown >= RC2) { err |= ERR_BOND_OVERFLOW; break; }            // every slot of the row in range: flagged
    if (fo) { items[ro] = (uint32_t)lane | ((uint32_t)imol << 5) | ((uint32_t)ro << 11) | IT_OLD; recj[ro] = (uint8_t)j; }
    if (fn) { items[rn] = (uint32_t)lane | ((uint32_t)imol << 5) | ((uint32_t)rn << 11) | IT_NEW; recj[rn] = (uint8_t)j; }
    const uint32_t bmj0 = (fo || fn) ? (BM[j] & ~excl) : 0u;
    const uint32_t dbase = ((uint32_t)j << 5) | ((fo ? (uint32_t)ro : IT_NONE) << 11) | ((fn ? (uint32_t)rn : IT_NONE) << 16);

    // ---- rounds: all centres at once when their candidates fit the table (always, at physical densities),
    // else two centre lanes per round
    const int cnt0 = __popc(bmj0);
    int tot0 = cnt0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot0 += __shfl_xor_sync(FULL, tot0, d);
    const int step = (nown + tot0 <= IT2) ? 32 : 2;
#pragma unroll 1
    for (int c0 = 0; c0 < 32; c0 += step) {
        const bool mine = lane >= c0 && lane < c0 + step;
        uint32_t bmj = mine ? bmj0 : 0u;
        const int cnt = __popc(bmj);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        const int first = (c0 == 0) ? 0 : nown;           // own bonds are evaluated in the first round only
        const int nitems = nown + __shfl_sync(FULL, incl, 31);
        {
            int pos = nown + incl - cnt;
#pragma unroll 1
            while (bmj) {
                const int s2 = __ffs(bmj) - 1; bmj &= bmj - 1;
                items[pos++] = dbase | (uint32_t)s2;
            }
        }
        __syncwarp();
        // ---- items: geometry, radial functions, one exponential; own bonds leave their record and pair energy,
        // candidates close the j-centred triplets of both variants
#pragma unroll 1
        for (int t0 = first; t0 < nitems; t0 += 32) {
            const int t = t0 + lane;
            const bool in = t < nitems;
            const uint32_t d = items[in ? t : first];
            const int s = d & 31, row = (d >> 5) & 63;
            const uint32_t ty = d >> 21;
            const uint32_t e2 = L[row * LC + s];
            const int k = e2 & 63, im2 = e2 >> 11;
            const bool isnew = ty == 2u;
            const double* cp = isnew ? T : P + row;
            const int cs = isnew ? 1 : N;
            const double tx = (P[k] + V[im2]) - cp[0];
            const double ty_ = (P[N + k] + V[IVC + im2]) - cp[cs];
            const double tz = (P[2 * N + k] + V[2 * IVC + im2]) - cp[2 * cs];
            const double sq0 = dist2(tx, ty_, tz);
            const bool ok = in && (sq0 < CK.rcsq);
            const double sq = ok ? sq0 : CK.ss;                    // any length inside the cut-off
            double ir, isr;
            bond_radial(sq, ir, isr);
            const double e1 = exp_fast(CK.sig02 * isr);            // exp(sigma*isr) = e1^5, exp(gamma*sigma*isr) = e1^6
            const double e_2 = e1 * e1, e_4 = e_2 * e_2;
            const double g = e_4 * e_2;
            const double ux = tx * ir, uy = ty_ * ir, uz = tz * ir;
            if (t0 == 0) {                                          // all own bonds sit in the first 32 items
                if (ok && ty != 0u) {
                    const int r = (d >> 11) & 31;
                    q[r] = ux; q[QC + r] = uy; q[2 * QC + r] = uz; q[3 * QC + r] = g;
                    const double s2 = CK.ss * ir * ir;
                    const double pe = CK.aeps * (CK.bigb * (s2 * s2) - 1.0) * (e_4 * e1);
                    if (ty == 1u) ao += pe; else an += pe;
                }
                __syncwarp();
            }
            {
                const uint32_t ra = (d >> 11) & 31u, rb = (d >> 16) & 31u;
                const bool cand = ok && ty == 0u;
                const bool ho = cand && ra != IT_NONE, hn = cand && rb != IT_NONE;
                const int io = ho ? (int)ra : 0, in_ = hn ? (int)rb : 0;
                const double ex = CK.leps * g;
                const double cto = -(q[io] * ux + q[QC + io] * uy + q[2 * QC + io] * uz);
                const double d_o = cto - CK.cos0;
                const double vo = q[3 * QC + io] * ex * (d_o * d_o);
                if (ho && cto < CK.c099) ao += vo;                  // the cos < 0.99 filter of molint.F90:367-371
                const double ctn = -(q[in_] * ux + q[QC + in_] * uy + q[2 * QC + in_] * uz);
                const double dn = ctn - CK.cos0;
                const double vn = q[3 * QC + in_] * ex * (dn * dn);
                if (hn && ctn < CK.c099) an += vn;
            }
        }
        __syncwarp();
        if (step == 32) break;
    }

    // ---- triplets centred on imol: all unordered pairs of bond records of one variant (rotation pairing:
    // record at position pos of a segment of n pairs with (pos + d) mod n, d = 1 .. n/2)
    {
        const int r = lane;
        const bool act = r < nown;
        const bool sg = r >= no;
        const int n = act ? (sg ? nw : no) : 0, pos = sg ? r - no : r;
        const int half = n >> 1, send = r - pos + n;
        const bool even = !(n & 1);
        const int rr = act ? r : 0;
        const double ux = q[rr], uy = q[QC + rr], uz = q[2 * QC + rr];
        const double g = act ? q[3 * QC + rr] : 0.0;
        const uint32_t jr = recj[rr];
        const int maxd = max(no, nw) >> 1;
        double tb = 0.0;
#pragma unroll 1
        for (int d = 1; d <= maxd; ++d) {
            int c = r + d;
            c = (c >= send) ? c - n : c;
            const bool on = (d <= half) && !(even && d == half && pos >= half);
            c = on ? c : rr;
            const double ct = ux * q[c] + uy * q[QC + c] + uz * q[2 * QC + c];
            const double mult = (recj[c] == jr) ? 3.0 : 1.0;
            const double dd = ct - CK.cos0;
            if (on && ct < CK.c099) tb += q[3 * QC + c] * (dd * dd) * mult;
        }
        tb *= CK.leps * g;
        if (sg) an += tb; else ao += tb;
    }
    __syncwarp();
    }
    reduce2(ao, an);
    eo = ao; en = an;
    __syncwarp();
    return err;
}

// ---------------------------------------------------------------- move generation, GB moves per call (warp 0)
// mc_moves.F90:1001-1067 in the reference's exact arithmetic.  Lane = 4*move + role; every lane of a move's quad
// derives the direction and the magnitude (identical instructions for all moves of the batch); roles 0 / 1 apply the
// fractional transform for "lattice 1 active" / "lattice 2 active", role 2 stores the plain displacement and the
// molecule.  D = draws per translation move incl. acceptance and switch.  Returns the index (0..nb) of the first
// move of the batch that is NOT a translation.
__device__ __forceinline__ int generate_moves(const WalkerView& w, const W2& x, const McParams& p, int N, int pos, int D, int nb)
{
    const int lane = lane_id(), m = lane >> 2, role = lane & 3;
    const bool act = m < nb;
    const double* u = w.rngbuf + pos + (act ? m : 0) * D;
    const double xi = u[0];
    const double Nd = (double)N;
    int imol = (int)xm(u[1], Nd) + 1;
    if (imol > N) imol = N;
    imol -= 1;
    double vx = xs(xm(2.0, u[2]), 1.0), vy = xs(xm(2.0, u[3]), 1.0), vz = xs(xm(2.0, u[4]), 1.0);
    const double norm = xd(1.0, xsqrt(xa(xa(xm(vx, vx), xm(vy, vy)), xm(vz, vz))));
    vx = xm(vx, norm); vy = xm(vy, norm); vz = xm(vz, norm);
    const double r = xs(xm(u[5], 2.0), 1.0);
    const double mt = w.sc->max_trans;
    vx = xm(xm(vx, mt), r); vy = xm(xm(vy, mt), r); vz = xm(xm(vz, mt), r);
    double* rec = x.gen + (act ? m : 0) * GF;
    if (role < 2) {
        // role 0: lattice 1 active -> image of the displacement in lattice 2: recip(1), hmatrix(2); role 1: the reverse
        const double* rm = w.recip + (role == 0 ? 0 : 9);
        const double* hm = w.cell + (role == 0 ? 9 : 0);
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
        x.gimol[m] = imol;
    }
    const uint32_t rare = __ballot_sync(FULL, act && role == 0 && !(xi < p.transP));
    return rare ? ((__ffs(rare) - 1) >> 2) : nb;
}

// ---------------------------------------------------------------- the kernel
template <int NT>
__global__ void __launch_bounds__(64, MW_MC2_BLOCKS) k_mc_run2(const __grid_constant__ DeviceState S,
                                                              const __grid_constant__ McParams p, int ncycles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int wi = blockIdx.x;
    if (wi >= S.W) return;
    const int tid = threadIdx.x, lane = tid & 31, lat = tid >> 5;
    const int N = (NT > 0) ? NT : S.N;
    const WalkerView w = carve_walker(smem, N, 2);
    const W2 x = carve2(w);
    load_walker_cta(S, wi, w, tid, 64);
    __syncthreads();
    WalkerScalars* sc = w.sc;
    const double Nd = (double)N;
    double* wgt = S.weight + (size_t)wi * S.NB;
    double* hist = S.hist + (size_t)wi * S.NB;
    double* uhist = S.uhist + (size_t)wi * S.NB;
    double* P = w.pos + lat * 3 * N;
    double* T = w.mv + lat * 6;
    int err = 0;
    const int cycle0 = sc->cycle;
    int rng_pos = 0;

    compute_bond_masks_warp(smem, N, 2, lat);
    if (lat == 0) {
        if (p.prob_error) err |= ERR_PROB;
        const uint64_t idx = sc->rng_index;
        __syncwarp();
        if (lane == 0) {
            *w.rngbase = idx & ~(uint64_t)1;
            w.lv[0] = log(sc->vol[0] / sc->vol[1]); w.lv[1] = log(sc->vol[1] / sc->vol[0]);
            x.ctl[CTL_STOP] = (err & ERR_PROB) ? 1 : 0;
        }
        rng_pos = (int)(idx & 1);
        __syncwarp();
    }
    __syncthreads();

    bool stop = x.ctl[CTL_STOP] != 0;
    int bpar = 0;                                                       // batch parity (both warps count alike)
    for (int cyc = 0; cyc < ncycles && !stop; ++cyc) {
        const int cycle = cycle0 + cyc + 1;
        if (lat == 0) {
            __syncwarp();
            if (lane == 0) {
                sc->cycle = cycle;
                if (p.dd) {                                            // mc_moves.F90:181-210
                    if (cycle < p.eq_mc_cycles) sc->in_window = (sc->mu > sc->mu_lo) && (sc->mu < sc->mu_hi);
                    else if (cycle == p.eq_mc_cycles) { if (!sc->in_window) x.ctl[CTL_STOP] = 1; }
                    else sc->in_window = 1;
                }
            }
            __syncwarp();
            if (x.ctl[CTL_STOP]) err |= ERR_WINDOW;
        }
        if (cycle % p.list_update_int == 0) {                          // :218-222, each warp its lattice
            err |= compute_neighbours_warp(smem, N, 2, lat);
            compute_bond_masks_warp(smem, N, 2, lat);
        }
        const bool dd_eq = p.dd && (cycle < p.eq_mc_cycles);
        const bool bins_on = !(cycle < p.eq_mc_cycles);                // mc_update_wl_bins: :1615
        const bool do_switch = p.always_switch && !dd_eq;
        const bool fuse_switch = do_switch && p.samplerun;             // weights fixed: eta of the switch is already known
        const int D = 7 + (do_switch ? 1 : 0);                         // draws per translation move

        int imove = 0;
        while (imove < N) {                                            // :224-250, GB moves per batch
            if (lat == 0) {
                // every batch starts with a refill at the current draw index (the buffer starts at an even index)
                const uint64_t next = *w.rngbase + (uint64_t)rng_pos;
                __syncwarp();
                if (lane == 0) *w.rngbase = next & ~(uint64_t)1;
                __syncwarp();
                rng_pos = (int)(next & 1);
                rng_refill(smem, N, 2, S, p, wi);
                const int nb = min(GB, N - imove);
                const int nr = generate_moves(w, x, p, N, rng_pos, D, nb);
                if (lane == 0) { x.ctl[bpar * 3] = nb; x.ctl[bpar * 3 + 1] = nr; x.ctl[bpar * 3 + 2] = (sc->ls == 1); }
            }
            __syncthreads();
            if (x.ctl[CTL_STOP]) { stop = true; break; }
            const int nb = x.ctl[bpar * 3], nr = x.ctl[bpar * 3 + 1];
            bool one = x.ctl[bpar * 3 + 2] != 0;
            bpar ^= 1;

            for (int m = 0; m < nr; ++m) {
                // ====================== mc_water_translation (mc_moves.F90:966-1213) ======================
                const int imol = x.gimol[m];
                {
                    // displacement of my lattice: the plain one when it is the active lattice, else its image
                    const int off = (lat == 0) ? (one ? 0 : 6) : (one ? 3 : 0);
                    if (lane < 3) {
                        const double tv = x.gen[m * GF + off + lane];
                        T[lane] = xa(P[lane * N + imol], tv);
                        T[3 + lane] = tv;
                    }
                    __syncwarp();
                }
                double eo, en;
                uint32_t mo, mn;
                err |= local_energies_lat<NT>(w, x, lat, imol, eo, en, mo, mn);
                if (lane == 0) { x.xch[lat * 2] = eo; x.xch[lat * 2 + 1] = en; }
                __syncthreads();                                        // A: both lattices' energies
                if (lat == 0) {
                    if (lane == 0) atomicAdd(S.transcount + (size_t)wi * N + imol, 1);
                    const double eo1 = x.xch[2], en1 = x.xch[3];
                    // model_energy bookkeeping exactly as :1013-1016, :1087-1090
                    const double Eb0 = sc->E[0], Eb1 = sc->E[1];
                    const double Ea0 = (Eb0 - eo) + en, Ea1 = (Eb1 - eo1) + en1;
                    const double dE0 = en - eo, dE1 = en1 - eo1;
                    const double mu_old = sc->mu;
                    const double dm = (dE0 - dE1) * p.beta;
                    const double mu_acc = mu_old + dm;                  // :1113
                    const double mu_rej = mu_acc - dm;                  // :1195 -- (mu + d) - d, not a copy
                    // three weight look-ups in parallel lanes: eta(mu), eta(mu_acc), eta(mu_rej)
                    const double mine = (lane == 0) ? mu_old : (lane == 1) ? mu_acc : mu_rej;
                    EtaBin eb; eb.eta = 0.0; eb.k = 0;
                    if (lane < 3) eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, mine);
                    const double eta_old = __shfl_sync(FULL, eb.eta, 0);
                    const double eta_acc = __shfl_sync(FULL, eb.eta, 1), eta_rej = __shfl_sync(FULL, eb.eta, 2);
                    const int k_acc = __shfl_sync(FULL, eb.k, 1), k_rej = __shfl_sync(FULL, eb.k, 2);
                    const double diffkT = (one ? dE0 : dE1) * p.beta + eta_acc - eta_old;
                    // one exponential pass: lane 0 acceptance, lanes 1/2 switch probability if accepted / rejected,
                    // lanes 3/4 unbiased-histogram factor if accepted / rejected
                    double arg = -diffkT;
                    if (fuse_switch && (lane == 1 || lane == 2)) {
                        const bool a = (lane == 1);
                        arg = switch_arg(p, w, a ? Ea0 : Eb0, a ? Ea1 : Eb1, one, a ? eta_acc : eta_rej, Nd);
                    }
                    if (lane == 3) arg = eta_acc - p.log_unbiased_norm;
                    if (lane == 4) arg = eta_rej - p.log_unbiased_norm;
                    const double ex = (arg > 0.0 && lane < 3) ? 1.0 : exp_call(fmin(arg, 700.0));
                    const double* u = w.rngbuf + rng_pos + m * D;
                    const double zeta = u[6];
                    const bool accepted = zeta < __shfl_sync(FULL, ex, 0);                      // :1145-1146
                    __syncwarp();
                    if (lane == 0) {
                        if (accepted) {
                            sc->acc_r += 1;
                            const double dmu = fabs(dE0 - dE1) * p.beta;
                            if (dmu < sc->min_dmu) sc->min_dmu = dmu;
                            if (dmu > sc->max_dmu) sc->max_dmu = dmu;
                            sc->E[0] = Ea0; sc->E[1] = Ea1; sc->mu = mu_acc;
                        } else {
                            sc->mu = mu_rej;
                        }
                        sc->att_r += 1;
                    }
                    __syncwarp();
                    // ====================== mc_update_wl_bins (mc_moves.F90:1597-1689) ======================
                    const int kb = accepted ? k_acc : k_rej;
                    if (bins_on && kb >= 1 && kb <= p.nbins) {
                        const double c = __ldg(S.hinc + kb - 1);
                        if (lane == 0) atomicAdd(hist + kb - 1, c);
                        if (p.samplerun) {
                            const double uf = __shfl_sync(FULL, ex, accepted ? 3 : 4);
                            if (lane == 0) atomicAdd(uhist + kb - 1, c * uf);
                        } else {
                            update_weights(smem, N, 2, p, S.binwidth, wgt, hist, kb);
                        }
                    }
                    // ====================== mc_lattice_switch (mc_moves.F90:1536-1594) ======================
                    if (fuse_switch) {
                        const double compare = __shfl_sync(FULL, ex, accepted ? 1 : 2);
                        const bool sw = u[7] < compare;
                        if (sw) {
                            const double mu_sw = mu_paren(p, sc, Nd, w.lv[0]);
                            __syncwarp();
                            if (lane == 0) { sc->acc_s += 1; sc->mu = mu_sw; sc->ls = 3 - sc->ls; }
                        }
                        if (lane == 0) sc->att_s += 1;
                        __syncwarp();
                    } else if (do_switch) {
                        // weights may have moved in update_weights: the reference looks eta up again
                        lattice_switch_cold(smem, S, p, wi, 2, rng_pos + m * D + 7);
                        __syncwarp();
                    }
                    if (lane == 0) x.ctl[CTL_DEC] = (accepted ? 1 : 0) | (sc->ls == 1 ? 2 : 0);
                }
                __syncthreads();                                        // B: the decision
                const int dec = x.ctl[CTL_DEC];
                one = (dec & 2) != 0;
                if (dec & 1) {
                    // commit: new position, own bond mask, and the reverse bits of the bonds that formed / broke
                    if (lane < 3) P[lane * N + imol] = T[lane];
                    if (lane == 0) w.bmask[lat * N + imol] = mn;
                    const uint32_t changed = mo ^ mn;
                    if ((changed >> lane) & 1u) {
                        const uint32_t e = w.list[((size_t)lat * N + imol) * LC + lane];
                        uint32_t* bj = w.bmask + lat * N + (e & 63u);
                        const uint32_t bit = 1u << ((e >> 6) & 31u);
                        if ((mn >> lane) & 1u) atomicOr(bj, bit); else atomicAnd(bj, ~bit);
                    }
                } else if (lane < 3) {
                    // reject: the reference restores by (x+t)-t, not by copy (mc_moves.F90:1186)
                    P[lane * N + imol] = xs(T[lane], T[3 + lane]);
                }
                __syncwarp();
            }
            imove += nr;
            if (lat == 0) rng_pos += nr * D;
            if (nr < nb) {
                // ---------------- rare move types (warp 0; warp 1 waits at the next batch barrier) ----------------
                if (lat == 0) {
                    const double xi = w.rngbuf[rng_pos];
                    rng_pos += 1;
                    if (xi < p.volP) {
                        rng_pos = volume_move<2>(smem, S, p, wi, rng_pos);
                        const EtaBin eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, sc->mu);
                        if (bins_on && eb.k >= 1 && eb.k <= p.nbins) {
                            const double c = __ldg(S.hinc + eb.k - 1);
                            if (lane == 0) atomicAdd(hist + eb.k - 1, c);
                            if (p.samplerun) {
                                if (lane == 0) atomicAdd(uhist + eb.k - 1, c * exp(eb.eta - p.log_unbiased_norm));
                            } else {
                                update_weights(smem, N, 2, p, S.binwidth, wgt, hist, eb.k);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) sc->att_v += 1;
                        __syncwarp();
                    } else if (xi < p.swP) {
                        if (!dd_eq) rng_pos = lattice_switch_cold(smem, S, p, wi, 2, rng_pos);
                    }
                    if (do_switch) rng_pos = lattice_switch_cold(smem, S, p, wi, 2, rng_pos);
                }
                imove += 1;
            }
        }
        if (stop) break;
        if (lat == 0) {
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int l = 0; l < 2; ++l) {                          // :253-255
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
                    default: v = w.cell[lane - 7]; break;
                    }
                    S.therm[((size_t)wi * S.therm_cap + n) * THERM_ROW + lane] = v;
                }
                __syncwarp();
                if (lane == 0) S.therm_n[wi] = n + 1;
            }
        }
    }
    __syncthreads();
    if (lat == 0 && lane == 0) {
        const uint64_t idx = *w.rngbase + (uint64_t)rng_pos;
        if (p.rng_mode == 1 && idx > S.fifo_len) err |= ERR_RNG_UNDERRUN;
        sc->rng_index = idx;
    }
    err = (int)__reduce_or_sync(FULL, (unsigned)err);
    if (lane == 0 && err) atomicOr(&sc->error, err);
    __syncthreads();
    store_walker_cta(S, wi, w, tid, 64);
}

}  // namespace mw
