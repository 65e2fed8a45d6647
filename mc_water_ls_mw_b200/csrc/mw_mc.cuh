// mw_mc.cuh -- device side of the Monte-Carlo move loop (mc_moves.F90:217-255,
// :893-964, :966-1213, :1216-1534, :1536-1594, :1597-1689, :2187-2215) for one
// walker per warp, plus the global-memory layout of a context.
//
// Code layout matters here: the hot loop (translation move + weight look-up +
// histogram update + lattice switch) is kept small and inlined, everything that
// runs rarely (volume move, list rebuild, full energy, image vectors) is
// __noinline__ so that it does not share the instruction cache with the loop.
#pragma once
#include "mw_device.cuh"

#ifndef MW_MC_BLOCKS
#define MW_MC_BLOCKS 14      // resident walkers (1-warp CTAs) per SM the register allocation is bounded for
#endif

namespace mw {

// Run parameters shared by all walkers (kernel argument, by value).
struct McParams {
    double beta, pressure;
    double transP, volP, swP;
    double r_pos, r_neg, a_pos, a_neg, log_r_pos, log_r_neg, inv_log_r_pos, inv_log_r_neg;
    double c_pos, c_neg;    // (1 - r)/a of the two geometric progressions
    double av_binwidth, log_unbiased_norm;
    double mu_min, mu_max;
    double orig_wl_factor, wl_alpha;
    unsigned long long seed;
    unsigned int stream0;
    int rng_mode;           // 0 philox, 1 fifo
    int nbins;
    int npt, eta_interp, samplerun, leshift, always_switch, dd, wl_swetnam;
    int list_update_int, eq_mc_cycles;
    int prob_error;
};

// Global-memory arrays of a context (all walkers).
struct DeviceState {
    int N, nlat, W, NB;
    double* pos;        // [W][nlat][3][N]
    double* ref;        // [W][nlat][3][N]   ref_ljr
    double* cell;       // [W][nlat][9]
    double* recip;      // [W][nlat][9]
    double* refcell;    // [W][nlat][9]      ref_hmatrix
    double* iv;         // [W][nlat][3][IVC]
    int*    niv;        // [W][2]
    uint16_t* list;     // [W][nlat][N][LC]
    uint8_t*  nn;       // [W][nlat][N]
    WalkerScalars* scal;// [W]
    double* weight;     // [W][NB]
    double* hist;       // [W][NB]
    double* uhist;      // [W][NB]
    double* wbase;      // [W][NB]  eta_last_sync    comms_mpi.f90:86
    double* hbase;      // [W][NB]  hist_last_sync
    double* ubase;      // [W][NB]  uhist_last_sync
    int*    transcount; // [W][N]   mc_translations
    double* mubin;      // [NB]
    double* binwidth;   // [NB]
    double* ginv;       // [NB]  ginv[k-1] = 2/(binwidth(k) + binwidth(k+1)), k = 1..NB-1
    double* hinc;       // [NB]  av_binwidth/binwidth(k): histogram increment of bin k (mc_moves.F90:1621)
    double* edge;       // [NB+1] bin edges of the grid: bin k (1-based) spans edge[k-1] .. edge[k] (mc_moves.F90:570-656)
    const double* fifo; // host-supplied random numbers (walker 0 only)
    unsigned long long fifo_len;
    // therm rows (main.f90:200-223): what the reference writes every file_output_int cycles, recorded by
    // the walker kernel so that a launch can span many output intervals
    double* therm;      // [W][therm_cap][THERM_ROW]
    int*    therm_n;    // [W] rows recorded since the last drain (may exceed therm_cap: the excess was dropped)
    int     therm_int, therm_cap;
    // scheduling of the warp-per-lattice walker kernel (mw2.cuh: persistent blocks take (walker, chunk of cycles)
    // units from a queue in global memory)
    int*    queue;      // [units of the launch] 0 = not yet published, else walker + 1
    int*    qctr;       // [0] entries taken, [1] entries published after the initial W, [2] walkers finished,
                        // [4..5] (64 bit) cycles completed by all walkers at unit boundaries
    int*    cyc_end;    // [W] cycle number at which the walker's part of the launch ends
    unsigned long long* wtime;   // [W][2] %globaltimer at the start of the walker's first / the end of its last unit
};
constexpr int THERM_ROW = 16;   // icyc, ls, E(1:2), ls_mu, volume(1:2), hmatrix(:,:,1)

// ---------------------------------------------------------------- staging
__device__ __forceinline__ void load_walker(const DeviceState& S, int wi, const WalkerView& w)
{
    const int N = S.N, nlat = S.nlat, lane = lane_id();
    const double* gp = S.pos + (size_t)wi * nlat * 3 * N;
    for (int t = lane; t < nlat * 3 * N; t += 32) w.pos[t] = gp[t];
    const double* gi = S.iv + (size_t)wi * nlat * 3 * IVC;
    for (int t = lane; t < nlat * 3 * IVC; t += 32) w.iv[t] = gi[t];
    if (lane < nlat * 9) {
        w.cell[lane] = S.cell[(size_t)wi * nlat * 9 + lane];
        w.recip[lane] = S.recip[(size_t)wi * nlat * 9 + lane];
    }
    if (lane < 2) w.niv[lane] = S.niv[wi * 2 + lane];
    // lists: 16-byte vector copies (N*LC*2 bytes per lattice is a multiple of 16)
    const uint4* gl = (const uint4*)(S.list + (size_t)wi * nlat * N * LC);
    uint4* sl = (uint4*)w.list;
    for (int t = lane; t < nlat * N * LC / 8; t += 32) sl[t] = gl[t];
    const uint8_t* gn = S.nn + (size_t)wi * nlat * N;
    for (int t = lane; t < nlat * N; t += 32) w.nn[t] = gn[t];
    // scalars: word-wise copy
    const uint32_t* gs = (const uint32_t*)(S.scal + wi);
    uint32_t* ss = (uint32_t*)w.sc;
    for (int t = lane; t < (int)(sizeof(WalkerScalars) / 4); t += 32) ss[t] = gs[t];
    __syncwarp();
}

__device__ __forceinline__ void store_walker(const DeviceState& S, int wi, const WalkerView& w, bool lists)
{
    const int N = S.N, nlat = S.nlat, lane = lane_id();
    __syncwarp();
    double* gp = S.pos + (size_t)wi * nlat * 3 * N;
    for (int t = lane; t < nlat * 3 * N; t += 32) gp[t] = w.pos[t];
    double* gi = S.iv + (size_t)wi * nlat * 3 * IVC;
    for (int t = lane; t < nlat * 3 * IVC; t += 32) gi[t] = w.iv[t];
    if (lane < nlat * 9) {
        S.cell[(size_t)wi * nlat * 9 + lane] = w.cell[lane];
        S.recip[(size_t)wi * nlat * 9 + lane] = w.recip[lane];
    }
    if (lane < 2) S.niv[wi * 2 + lane] = w.niv[lane];
    if (lists) {
        uint4* gl = (uint4*)(S.list + (size_t)wi * nlat * N * LC);
        const uint4* sl = (const uint4*)w.list;
        for (int t = lane; t < nlat * N * LC / 8; t += 32) gl[t] = sl[t];
        uint8_t* gn = S.nn + (size_t)wi * nlat * N;
        for (int t = lane; t < nlat * N; t += 32) gn[t] = w.nn[t];
    }
    uint32_t* gs = (uint32_t*)(S.scal + wi);
    const uint32_t* ss = (const uint32_t*)w.sc;
    for (int t = lane; t < (int)(sizeof(WalkerScalars) / 4); t += 32) gs[t] = ss[t];
}

// ---------------------------------------------------------------- random numbers
// Per-walker stream of U[0,1) numbers (random.f90:87-102), buffered RB at a time in shared
// memory.  mode 0: Philox (draw n = half n&1 of block n>>1); mode 1: host FIFO (draw n = fifo[n]).
// `pos` (index of the next draw inside the buffer) is carried in a register by the caller.
__device__ __noinline__ void rng_refill_at(const uint64_t* rngbase, double* rngbuf, const DeviceState& S, const McParams& p, int wi)
{
    const int lane = lane_id();
    const uint64_t base = *rngbase;
    __syncwarp();
    double v0, v1;
    if (p.rng_mode == 0) {
        philox_block(p.seed, p.stream0 + (uint32_t)wi, (base >> 1) + (uint64_t)lane, v0, v1);
    } else {
        const uint64_t i0 = base + 2u * (uint64_t)lane;
        v0 = (i0 < S.fifo_len) ? S.fifo[i0] : 0.5;
        v1 = (i0 + 1 < S.fifo_len) ? S.fifo[i0 + 1] : 0.5;
    }
    rngbuf[2 * lane] = v0; rngbuf[2 * lane + 1] = v1;
    __syncwarp();
}

__device__ __forceinline__ void rng_refill(unsigned char* smem, int N, int nlat, const DeviceState& S, const McParams& p, int wi)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    rng_refill_at(w.rngbase, w.rngbuf, S, p, wi);
}

struct Rng {
    unsigned char* smem; int N, nlat, wi;
    const DeviceState* S; const McParams* p;
    double* buf; uint64_t* base;
    int pos;
    // make sure the next n draws are in the buffer (n <= RB - 1); called once per trial move, so that
    // draw() itself is branch-free.  The buffer always starts at an even draw index (Philox block).
    __device__ __forceinline__ void reserve(int n)
    {
        if (pos + n > RB) {
            __syncwarp();
            const uint64_t next = *base + (uint64_t)pos;
            __syncwarp();
            *base = next & ~(uint64_t)1;      // every lane stores the same value
            rng_refill(smem, N, nlat, *S, *p, wi);
            pos = (int)(next & 1);
        }
    }
    __device__ __forceinline__ double draw() { return buf[pos++]; }
};
constexpr int DRAWS_PER_MOVE = 8;            // at most 8 draws per trial move (+ switch)

// ---------------------------------------------------------------- order parameter / weights
struct EtaBin { double eta; int k; };

__device__ __noinline__ int bin_exact(double arg, double lr) { return (int)(log(arg) / lr); }

// mu_to_bin (mc_moves.F90:2187-2215, 1-based bin) and eta_weight (mc_moves.F90:893-964) in one
// call.  The two sign branches of mu_to_bin share one log: for mu > 0, mu - 0.5 == |mu| - 0.5.
// wgt is this walker's weight array (global memory, read through L2 because the same warp
// updates it when generating weights).
__device__ __forceinline__ int mu_to_bin_dev(const McParams& p, double mu)
{
    const int nb = p.nbins;
    if (fabs(mu) <= 0.5) return nb / 2 + 1;
    const bool pos = mu > 0.0;
    // (|mu| - 0.5)*(1 - r)/a with (1 - r)/a precomputed: identical to the reference's order for a == 1
    const double arg = 1.0 - (fabs(mu) - 0.5) * (pos ? p.c_pos : p.c_neg);
    // int(log(arg)/log(r)): fast logarithm; the library log and the true division decide only
    // when the quotient is within 1e-7 of an integer
    const double y = log_fast(arg) * (pos ? p.inv_log_r_pos : p.inv_log_r_neg);
    int t = (int)y;
    if (fabs(y - rint(y)) < 1e-7 || !(arg > 0.0)) t = bin_exact(arg, pos ? p.log_r_pos : p.log_r_neg);
    return pos ? nb / 2 + 2 + t : nb / 2 - t;
}

// eta_weight (mc_moves.F90:893-964) for the bin k = mu_to_bin(mu)
__device__ __forceinline__ double eta_of_bin(const McParams& p, const double* __restrict__ mubin,
                                             const double* __restrict__ binwidth /* = DeviceState::ginv */, const WalkerScalars* sc,
                                             const double* wgt, double mu, int k)
{
    const int nb = p.nbins;
    if (!sc->in_window) return 0.0;                          // undefined in the reference (:913); defined as 0
    if (mu < sc->mu_lo || mu > sc->mu_hi) return F_HUGE;
    k = min(max(k, 1), nb);                                  // memory safety at mu == mu_max (reference would overrun)
    const double* w = wgt - 1;
    const double* mb = mubin - 1;
    // fixed weights (sample runs) are read-only for the whole launch: take them through the L1; when the
    // same warp updates them every move (weight generation) they come from the L2
    const bool ro = p.samplerun != 0;
    if (!p.eta_interp) return ro ? __ldg(w + k) : __ldcg(w + k);
    int ka, kb, kr;                                          // gradient between bins ka<kb, anchored at kr
    if (k == sc->start_bin)      { ka = k; kb = k + 1; kr = k; }
    else if (k == sc->end_bin)   { ka = k - 1; kb = k; kr = k; }
    else if (mu > __ldg(mb + k)) { ka = k; kb = k + 1; kr = k; }
    else                         { ka = k - 1; kb = k; kr = k - 1; }
    ka = max(ka, 1); kb = min(kb, nb);
    const double wa = ro ? __ldg(w + ka) : __ldcg(w + ka), wb = ro ? __ldg(w + kb) : __ldcg(w + kb);
    const double g = (wb - wa) * __ldg(binwidth + ka - 1);       // binwidth = the 2/(bw(ka)+bw(kb)) table here
    const double wr = (kr == ka) ? wa : wb;
    return wr + (mu - __ldg(mb + kr)) * g;
}

__device__ __noinline__ EtaBin eta_bin(const McParams& p, const double* __restrict__ mubin,
                                       const double* __restrict__ binwidth /* = DeviceState::ginv */, const WalkerScalars* sc,
                                       const double* wgt, double mu)
{
    EtaBin r;
    r.k = mu_to_bin_dev(p, mu);
    r.eta = eta_of_bin(p, mubin, binwidth, sc, wgt, mu, r.k);
    return r;
}

// The same with the bin taken from the table of bin edges, starting at a guess (the bin of the previous move: mu
// moves by a few kT per move).  mu_to_bin's closed form and the edges agree to ~1e-12; whenever mu is within 1e-9
// of an edge -- or further than one bin from the guess -- the closed form decides, so the bin is always the
// reference's.
__device__ __noinline__ int mu_to_bin_call(const McParams& p, double mu) { return mu_to_bin_dev(p, mu); }

__device__ __forceinline__ EtaBin eta_bin_near(const McParams& p, const DeviceState& S, const WalkerScalars* sc,
                                               const double* wgt, double mu, int kguess)
{
    EtaBin r;
    const int nb = p.nbins;
    // Fast path: mu is still in the bin of the guess (the usual case) and that bin is an inner bin of the window.
    // Everything eta_weight needs around the guess is requested at once -- bin edges, mid-bin values, weights and
    // gradient factors of the bins g-1, g, g+1 -- so the look-up waits for ONE round trip to the L1 / L2 instead
    // of three dependent ones (edge -> mid-bin value -> weights); the arithmetic is eta_of_bin's, term for term.
    if (nb >= 3 && p.eta_interp) {
        const int g = min(max(kguess, 2), nb - 1);
        const bool ro = p.samplerun != 0;
        const double* w = wgt - 1;
        const double* mb = S.mubin - 1;
        const double e0 = __ldg(S.edge + g - 1), e1 = __ldg(S.edge + g);
        const double m_m = __ldg(mb + g - 1), m_0 = __ldg(mb + g);
        const double w_m = ro ? __ldg(w + g - 1) : __ldcg(w + g - 1), w_0 = ro ? __ldg(w + g) : __ldcg(w + g),
                     w_p = ro ? __ldg(w + g + 1) : __ldcg(w + g + 1);
        const double g_m = __ldg(S.ginv + g - 2), g_0 = __ldg(S.ginv + g - 1);
        if (sc->in_window && !(mu < sc->mu_lo || mu > sc->mu_hi) && (mu > e0 + 1e-9) && (mu < e1 - 1e-9) &&
            g != sc->start_bin && g != sc->end_bin) {
            const bool up = mu > m_0;                        // gradient between bins g, g+1 anchored at g; else g-1, g at g-1
            const double wa = up ? w_0 : w_m, wb = up ? w_p : w_0;
            const double gg = (wb - wa) * (up ? g_0 : g_m);
            r.k = g;
            r.eta = wa + (mu - (up ? m_0 : m_m)) * gg;
            return r;
        }
    }
    int k = min(max(kguess, 1), nb);
    double lo = __ldg(S.edge + k - 1), hi = __ldg(S.edge + k);
    if (mu >= hi && k < nb) { ++k; lo = hi; hi = __ldg(S.edge + k); }
    else if (mu < lo && k > 1) { --k; hi = lo; lo = __ldg(S.edge + k - 1); }
    if (!((mu > lo + 1e-9) && (mu < hi - 1e-9))) k = mu_to_bin_call(p, mu);
    r.k = k;
    r.eta = eta_of_bin(p, S.mubin, S.ginv, sc, wgt, mu, k);
    return r;
}

// mu recomputed from scratch, parenthesised association (mc_moves.F90:1370-1372, :1525-1527, :1583-1585)
__device__ __forceinline__ double mu_paren(const McParams& p, const WalkerScalars* sc, double N, double lv12)
{
    double mu = (sc->E[0] + p.pressure * sc->vol[0]) - (sc->E[1] + p.pressure * sc->vol[1]);
    if (p.leshift) mu = mu - sc->refH[0] + sc->refH[1];
    return mu * p.beta - N * lv12;
}

// -(diffkT) of mc_lattice_switch (mc_moves.F90:1562-1574) for model energies (E0,E1); eta enters
// as (x + eta) - eta exactly as in the reference
__device__ __forceinline__ double switch_arg(const McParams& p, const WalkerScalars* sc, const double* lv, double E0, double E1,
                                             bool one, double eta, double N)
{
    const double Es = one ? E0 : E1, En = one ? E1 : E0;
    double d;
    if (p.npt) {
        const double Vs = one ? sc->vol[0] : sc->vol[1], Vn = one ? sc->vol[1] : sc->vol[0];
        const double lvn = one ? lv[1] : lv[0];                    // log(volume(lsn)/volume(ls))
        d = p.beta * En - p.beta * Es + p.beta * p.pressure * (Vn - Vs) - N * lvn + eta - eta;
    } else {
        d = p.beta * En - p.beta * Es + eta - eta;
    }
    if (p.leshift) {
        const double Rs = one ? sc->refH[0] : sc->refH[1], Rn = one ? sc->refH[1] : sc->refH[0];
        d = d - p.beta * Rn + p.beta * Rs;
    }
    return -d;
}

// mc_lattice_switch (mc_moves.F90:1536-1594), stand-alone form (cold paths)
__device__ __noinline__ int lattice_switch_at(WalkerScalars* sc, const double* lv, const double* rngbuf, const DeviceState& S,
                                              const McParams& p, int wi, int rng_pos)
{
    const int N = S.N;
    const double eta = eta_bin(p, S.mubin, S.ginv, sc, S.weight + (size_t)wi * S.NB, sc->mu).eta;
    const double arg = switch_arg(p, sc, lv, sc->E[0], sc->E[1], sc->ls == 1, eta, (double)N);
    const double compare = (arg > 0.0) ? 1.0 : exp_call(arg);
    const double x = rngbuf[rng_pos];
    const double mu_new = mu_paren(p, sc, (double)N, lv[0]);
    __syncwarp();
    if (lane_id() == 0) {
        if (x < compare) {
            sc->acc_s += 1;
            sc->mu = mu_new;
            sc->ls = 3 - sc->ls;
        }
        sc->att_s += 1;
    }
    __syncwarp();
    return rng_pos + 1;
}

__device__ __forceinline__ int lattice_switch_cold(unsigned char* smem, const DeviceState& S, const McParams& p, int wi,
                                                   int nlat, int rng_pos)
{
    const WalkerView w = carve_walker(smem, S.N, nlat);
    return lattice_switch_at(w.sc, w.lv, w.rngbuf, S, p, wi, rng_pos);
}

// mc_moves.F90:1597-1689 for the weight-generation case (not samplerun): the histogram increment
// is done by the caller; this updates wl_factor (Swetnam / 1-over-t variants) and the weights.
__device__ __noinline__ void update_weights_at(WalkerScalars* sc, int N, const McParams& p,
                                               const double* __restrict__ binwidth, double* wgt, const double* hist, int k)
{
    const int nb = p.nbins, lane = lane_id();
    if (p.wl_swetnam) {
        // Swetnam's increment from the current histogram (:1636-1653); sequential order as in the reference
        __syncwarp();
        const double sumhist = sc->sumhist + 1.0;
        sc->sumhist = sumhist;
        double f = 0.0;
        for (int i = 0; i < nb; ++i) {
            const double bwi = __ldg(binwidth + i);
            const double binfrac = bwi / (p.mu_max - p.mu_min - 1.0);
            const double d = __ldcg(hist + i) * bwi / sumhist - binfrac;
            f = f + d * d;
        }
        f = sqrt(f / (double)nb);
        f = log(f);
        f = f * p.wl_alpha * (double)nb;
        sc->wl_factor = fmin(f, p.orig_wl_factor);
    } else if (sc->wl_invt_active) {
        sc->wl_factor = fmin(sc->wl_factor, (double)nb / (double)(sc->cycle * N));
    }
    const double wk_old = __ldcg(wgt + k - 1);
    // weight(k) = weight(k) + av_binwidth*incr/binwidth(k)   (:1680)
    const double wk = wk_old + p.av_binwidth * sc->wl_factor / __ldg(binwidth + k - 1);
    __syncwarp();
    if (lane == 0) wgt[k - 1] = wk;
    __syncwarp();
    // minbin = minval(weight(start:end)); weight -= minbin (:1682-1685).  Subtracting an exact 0 is a
    // no-op, and the minimum stays 0 unless bin k was a zero bin.
    if (sc->wmin_zero && wk_old > 0.0 && wk >= wk_old) return;      // a negative (Swetnam) increment may create a new minimum
    const int sb = sc->start_bin, eb = sc->end_bin;
    double mn = F_HUGE;
    for (int i = sb - 1 + lane; i < eb; i += 32) mn = fmin(mn, __ldcg(wgt + i));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL, mn, d));
    if (mn != 0.0)
        for (int i = sb - 1 + lane; i < eb; i += 32) wgt[i] = __ldcg(wgt + i) - mn;
    sc->wmin_zero = 1;
    __syncwarp();
}

__device__ __forceinline__ void update_weights(unsigned char* smem, int N, int nlat, const McParams& p,
                                               const double* __restrict__ binwidth, double* wgt, const double* hist, int k)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    update_weights_at(w.sc, N, p, binwidth, wgt, hist, k);
}

// fractional rescale of one position (mc_moves.F90:1290-1315 and its three copies): exact arithmetic
__device__ __forceinline__ void rescale_pos(double& x, double& y, double& z, const double* rm, const double* hm)
{
    const double o0 = x, o1 = y, o2 = z;
    double n0 = xa(xa(xm(MW_H(rm,1,1), o0), xm(MW_H(rm,2,1), o1)), xm(MW_H(rm,3,1), o2));
    double n1 = xa(xa(xm(MW_H(rm,1,2), o0), xm(MW_H(rm,2,2), o1)), xm(MW_H(rm,3,2), o2));
    double n2 = xa(xa(xm(MW_H(rm,1,3), o0), xm(MW_H(rm,2,3), o1)), xm(MW_H(rm,3,3), o2));
    n0 = xm(xm(n0, 0.5), INV_PI); n1 = xm(xm(n1, 0.5), INV_PI); n2 = xm(xm(n2, 0.5), INV_PI);
    double t0 = xa(xa(xm(MW_H(hm,1,1), n0), xm(MW_H(hm,1,2), n1)), xm(MW_H(hm,1,3), n2));
    double t1 = xa(xa(xm(MW_H(hm,2,1), n0), xm(MW_H(hm,2,2), n1)), xm(MW_H(hm,2,3), n2));
    double t2 = xa(xa(xm(MW_H(hm,3,1), n0), xm(MW_H(hm,3,2), n1)), xm(MW_H(hm,3,3), n2));
    t0 = xs(t0, o0); t1 = xs(t1, o1); t2 = xs(t2, o2);
    x = xa(x, t0); y = xa(y, t1); z = xa(z, t2);
}

__device__ __noinline__ void rescale_all(unsigned char* smem, int N, int nlat, double* refpos, int lat)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    const int lane = lane_id();
    double rm[9], hm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { rm[k] = w.recip[lat * 9 + k]; hm[k] = w.cell[lat * 9 + k]; }
    double* P = w.pos + lat * 3 * N;
    double* R = refpos + (size_t)lat * 3 * N;
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

// recip matrix of the cell in shared memory -> shared memory (uniform stores)
__device__ __noinline__ void refresh_recip(unsigned char* smem, int N, int nlat, int lat)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    double hm[9], rm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = w.cell[lat * 9 + k];
    recipmatrix3(hm, rm);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 9; ++k) w.recip[lat * 9 + k] = rm[k];
    __syncwarp();
}

__device__ __forceinline__ double cell_volume(const WalkerView& w, int lat)
{
    double hm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = w.cell[lat * 9 + k];
    return fabs(determinant3(hm));
}

// mc_moves.F90:1216-1534.  Cold path (0.26 % of the moves): not inlined.  Returns the new
// position in the random-number buffer.
template <int NLAT>
__device__ __noinline__ int volume_move(unsigned char* smem, const DeviceState& S, const McParams& p, int wi, int rng_pos)
{
    const int N = S.N;
    const WalkerView w = carve_walker(smem, N, NLAT);
    WalkerScalars* sc = w.sc;
    Rng rng{smem, N, NLAT, wi, &S, &p, w.rngbuf, w.rngbase, rng_pos};
    const int lane = lane_id();
    const double Nd = (double)N;
    const double* wgt = S.weight + (size_t)wi * S.NB;
    double backupE[2] = {0.0, 0.0}, old_vol[2] = {0.0, 0.0}, newE[2] = {0.0, 0.0};
    int err = 0;
    // old cell + recip are parked in shared memory: save[0..17] = h, save[18..35] = recip
    double* save = w.save;
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        backupE[lat] = sc->E[lat];
        old_vol[lat] = sc->vol[lat];
        refresh_recip(smem, N, NLAT, lat);                 // :1260-1262
    }
    if (lane < NLAT * 9) { save[lane] = w.cell[lane]; save[18 + lane] = w.recip[lane]; }
    __syncwarp();
    double x = rng.draw();
    const int idim = (int)xm(x, 3.0) + 1;
    x = rng.draw();
    const int jdim = (int)xm(x, 3.0) + 1;
    x = rng.draw();
    const double dh = xm(xs(xm(2.0, x), 1.0), sc->dv_max);
    if (lane < NLAT) {
        double* hm = w.cell + lane * 9;
        const double v = xa(MW_H(hm, idim, jdim), dh);
        if (idim != jdim) MW_H(hm, jdim, idim) = xa(MW_H(hm, jdim, idim), dh);
        MW_H(hm, idim, jdim) = v;
    }
    __syncwarp();
    double* refpos = S.ref + (size_t)wi * NLAT * 3 * N;
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        rescale_all(smem, N, NLAT, refpos, lat);           // recip = old cell's, h = new cell
        sc->vol[lat] = cell_volume(w, lat);
        refresh_recip(smem, N, NLAT, lat);
        err |= compute_ivects_warp(smem, N, NLAT, lat);
        newE[lat] = full_energy_warp(smem, N, NLAT, lat);
        sc->E[lat] = newE[lat];
    }
    double old_eta = 0.0, new_eta = 0.0, old_mu = 0.0;
    const bool one = (sc->ls == 1);
    double nlv12 = w.lv[0], nlv21 = w.lv[1];
    if (NLAT == 2) {
        old_eta = eta_bin(p, S.mubin, S.ginv, sc, wgt, sc->mu).eta;
        old_mu = sc->mu;
        nlv12 = log(sc->vol[0] / sc->vol[1]); nlv21 = log(sc->vol[1] / sc->vol[0]);
        sc->mu = mu_paren(p, sc, Nd, nlv12);
        new_eta = eta_bin(p, S.mubin, S.ginv, sc, wgt, sc->mu).eta;
    }
    x = rng.draw();
    const double dE = one ? newE[0] - backupE[0] : newE[1] - backupE[1];
    const double Vs = one ? sc->vol[0] : sc->vol[1], Vo = one ? old_vol[0] : old_vol[1];
    const double diffkT = p.beta * dE + new_eta - old_eta + p.beta * p.pressure * (Vs - Vo) - Nd * log(Vs / Vo);
    const double compare = fmin(1.0, exp(-diffkT));
    if (x < compare) {
        sc->acc_v += 1;
        if (NLAT == 2) {
            const double dmu = fabs(old_mu - sc->mu);
            if (dmu < sc->min_dmu) sc->min_dmu = dmu;
            if (dmu > sc->max_dmu) sc->max_dmu = dmu;
        }
        w.lv[0] = nlv12; w.lv[1] = nlv21;
    } else {
        // :1434-1528: V,h <- old; rescale with recip(NEW) and h(OLD); recip <- old; ivects; E <- backup
        __syncwarp();
        if (lane < NLAT * 9) w.cell[lane] = save[lane];
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            sc->vol[lat] = old_vol[lat];
            rescale_all(smem, N, NLAT, refpos, lat);
        }
        if (lane < NLAT * 9) w.recip[lane] = save[18 + lane];
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            err |= compute_ivects_warp(smem, N, NLAT, lat);
            sc->E[lat] = backupE[lat];
            compute_bond_masks_warp(smem, N, NLAT, lat);   // positions moved by rounding; keep masks fresh
        }
        if (NLAT == 2) sc->mu = mu_paren(p, sc, Nd, w.lv[0]);
    }
    sc->error |= err;
    return rng.pos;
}

// commit of an accepted translation: new position, own bond mask, and the reverse bits of the
// bonds that formed / broke (rare)
template <int NLAT>
__device__ __forceinline__ void commit_translation(const WalkerView& w, int imol, const uint32_t* mo, const uint32_t* mn)
{
    const int N = w.N, lane = lane_id();
    __syncwarp();
#pragma unroll 1
    for (int lat = 0; lat < NLAT; ++lat) {
        double* P = w.pos + lat * 3 * N;
        if (lane < 3) P[lane * N + imol] = w.mv[lat * 6 + lane];
        const uint32_t mol = (lat == 0) ? mo[0] : mo[NLAT - 1], mnl = (lat == 0) ? mn[0] : mn[NLAT - 1];
        uint32_t changed = mol ^ mnl;
        if (lane == 0) w.bmask[lat * N + imol] = mnl;
        const int nv = w.niv[lat];
        const EntFmt F = ent_fmt(N);
        const uint32_t keep = ~(ent_has_rev(N) ? (31u << 6) : 0u);     // compare without the reverse-slot field
        while (changed) {
            const int s = __ffs(changed) - 1; changed &= changed - 1;
            const uint32_t e = w.list[((size_t)lat * N + imol) * LC + s];
            const int j = e & F.jmask, img = e >> F.ishift;
            const uint32_t target = ((uint32_t)inverse_image(img, nv) << F.ishift) | (uint32_t)imol;
            const int nnj = w.nn[lat * N + j];
            const uint32_t e2 = (lane < nnj) ? (w.list[((size_t)lat * N + j) * LC + lane] & keep) : 0xffffffffu;
            const uint32_t hit = __ballot_sync(FULL, e2 == target);
            if (hit && lane == 0) {
                const int s2 = __ffs(hit) - 1;
                const uint32_t bit = (mnl >> s) & 1u;
                w.bmask[lat * N + j] = (w.bmask[lat * N + j] & ~(1u << s2)) | (bit << s2);
            }
            __syncwarp();
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------- the walker kernel
// One warp (= one CTA of 32 threads) per walker; ncycles MC cycles of the hot
// part of mc_cycle (mc_moves.F90:117-255).
// NT > 0: the number of molecules as a compile-time constant (every offset into the walker's shared
// memory image folds into the LDS/STS immediates: ~8 % fewer instructions); NT == 0: N = S.N at run time.
template <int NLAT, int NT>
__global__ void __launch_bounds__(32, MW_MC_BLOCKS) k_mc_run(const __grid_constant__ DeviceState S,
                                               const __grid_constant__ McParams p, int ncycles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int wi = blockIdx.x;
    if (wi >= S.W) return;
    const int lane = lane_id();
    const int N = (NT > 0) ? NT : S.N;
    const WalkerView w = carve_walker(smem, N, NLAT);
    load_walker(S, wi, w);
    WalkerScalars* sc = w.sc;
    const double Nd = (double)N;
    double* wgt = S.weight + (size_t)wi * S.NB;
    double* hist = S.hist + (size_t)wi * S.NB;
    double* uhist = S.uhist + (size_t)wi * S.NB;
    int err = 0;
    if (p.prob_error) err |= ERR_PROB;

#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) compute_bond_masks_warp(smem, N, NLAT, lat);

    Rng rng{smem, N, NLAT, wi, &S, &p, w.rngbuf, w.rngbase, 0};
    {
        const uint64_t idx = sc->rng_index;
        *w.rngbase = idx & ~(uint64_t)1;
        rng.pos = (int)(idx & 1);
        rng_refill(smem, N, NLAT, S, p, wi);
    }
    if (NLAT == 2) { w.lv[0] = log(sc->vol[0] / sc->vol[1]); w.lv[1] = log(sc->vol[1] / sc->vol[0]); }

    for (int cyc = 0; cyc < ncycles && !(err & (ERR_WINDOW | ERR_PROB)); ++cyc) {
        const int cycle = sc->cycle + 1;
        sc->cycle = cycle;
        if (p.dd) {                                            // mc_moves.F90:181-210
            if (cycle < p.eq_mc_cycles) sc->in_window = (sc->mu > sc->mu_lo) && (sc->mu < sc->mu_hi);
            else if (cycle == p.eq_mc_cycles) { if (!sc->in_window) { err |= ERR_WINDOW; break; } }
            else sc->in_window = 1;
        }
        if (cycle % p.list_update_int == 0) {                  // :218-222
#pragma unroll
            for (int lat = 0; lat < NLAT; ++lat) {
                err |= compute_neighbours_warp(smem, N, NLAT, lat);
                compute_bond_masks_warp(smem, N, NLAT, lat);
            }
        }
        const bool dd_eq = p.dd && (cycle < p.eq_mc_cycles);
        const bool bins_on = !(cycle < p.eq_mc_cycles);        // mc_update_wl_bins: :1615
        const bool do_switch = (NLAT == 2) && p.always_switch && !dd_eq;
        const bool fuse_switch = do_switch && p.samplerun;     // weights fixed: eta of the switch is already known

        for (int imove = 0; imove < N; ++imove) {              // :224-250
            rng.reserve(DRAWS_PER_MOVE);
            const double xi = rng.draw();
            if (xi < p.transP) {
                // ====================== mc_water_translation (mc_moves.F90:966-1213) ======================
                const bool one = (sc->ls == 1);
                double x = rng.draw();
                int imol = (int)xm(x, Nd) + 1;
                if (imol > N) imol = N;
                imol -= 1;
                if (lane == 0) atomicAdd(S.transcount + (size_t)wi * N + imol, 1);
                x = rng.draw();
                double y = rng.draw();
                double z = rng.draw();
                x = xs(xm(2.0, x), 1.0); y = xs(xm(2.0, y), 1.0); z = xs(xm(2.0, z), 1.0);
                const double norm = xd(1.0, xsqrt(xa(xa(xm(x, x), xm(y, y)), xm(z, z))));
                x = xm(x, norm); y = xm(y, norm); z = xm(z, norm);
                const double r = xs(xm(rng.draw(), 2.0), 1.0);
                const double mt = sc->max_trans;
                x = xm(xm(x, mt), r); y = xm(xm(y, mt), r); z = xm(xm(z, mt), r);
                // displacement in the active lattice (x,y,z) and, through the fractional
                // coordinates of the active cell, in the other lattice (:1042-1067)
                double bx = 0.0, by = 0.0, bz = 0.0;
                if (NLAT == 2) {
                    const double* rm = w.recip + (one ? 0 : 9);
                    double sx = xa(xa(xm(MW_H(rm,1,1), x), xm(MW_H(rm,2,1), y)), xm(MW_H(rm,3,1), z));
                    double sy = xa(xa(xm(MW_H(rm,1,2), x), xm(MW_H(rm,2,2), y)), xm(MW_H(rm,3,2), z));
                    double sz = xa(xa(xm(MW_H(rm,1,3), x), xm(MW_H(rm,2,3), y)), xm(MW_H(rm,3,3), z));
                    sx = xm(xm(sx, 0.5), INV_PI); sy = xm(xm(sy, 0.5), INV_PI); sz = xm(xm(sz, 0.5), INV_PI);
                    const double* hm = w.cell + (one ? 9 : 0);
                    bx = xa(xa(xm(MW_H(hm,1,1), sx), xm(MW_H(hm,1,2), sy)), xm(MW_H(hm,1,3), sz));
                    by = xa(xa(xm(MW_H(hm,2,1), sx), xm(MW_H(hm,2,2), sy)), xm(MW_H(hm,2,3), sz));
                    bz = xa(xa(xm(MW_H(hm,3,1), sx), xm(MW_H(hm,3,2), sy)), xm(MW_H(hm,3,3), sz));
                }
                // the move record: trial position and displacement of imol in every lattice (shared memory,
                // every lane stores the same values: "uniform registers in shared memory")
                __syncwarp();
#pragma unroll
                for (int lat = 0; lat < NLAT; ++lat) {
                    const bool act = (lat == 0) == one;            // lattice `lat` is the active one
                    const double tx = act ? x : bx, ty = act ? y : by, tz = act ? z : bz;
                    const double* P = w.pos + lat * 3 * N;
                    double* m = w.mv + lat * 6;
                    m[0] = xa(P[imol], tx); m[1] = xa(P[N + imol], ty); m[2] = xa(P[2 * N + imol], tz);
                    m[3] = tx; m[4] = ty; m[5] = tz;
                }
                __syncwarp();
                double eo[2] = {0.0, 0.0}, en[2] = {0.0, 0.0};
                uint32_t mo[2] = {0, 0}, mn[2] = {0, 0};
                local_energies_warp<NLAT, true>(w, imol, eo, en, mo, mn);

                // model_energy bookkeeping exactly as :1013-1016, :1087-1090
                const double Eb0 = sc->E[0], Eb1 = sc->E[1];
                const double Ea0 = (Eb0 - eo[0]) + en[0], Ea1 = (Eb1 - eo[1]) + en[1];
                const double dE0 = en[0] - eo[0], dE1 = en[1] - eo[1];
                const double mu_old = sc->mu;
                double diffkT, mu_acc = mu_old, mu_rej = mu_old;
                double eta_acc = 0.0, eta_rej = 0.0;
                int k_acc = 0, k_rej = 0;
                if (NLAT == 1) {
                    diffkT = p.beta * dE0;
                    if (bins_on) {       // single box: ls_mu is never assigned (0) but the bins are still updated
                        const EtaBin eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, mu_old);
                        eta_acc = eta_rej = eb.eta; k_acc = k_rej = eb.k;
                    }
                } else {
                    const double dm = (dE0 - dE1) * p.beta;
                    mu_acc = mu_old + dm;                         // :1113
                    mu_rej = mu_acc - dm;                         // :1195 -- (mu + d) - d, not a copy
                    // three weight look-ups in parallel lanes: eta(mu), eta(mu_acc), eta(mu_rej)
                    const double mine = (lane == 0) ? mu_old : (lane == 1) ? mu_acc : mu_rej;
                    EtaBin eb; eb.eta = 0.0; eb.k = 0;
                    if (lane < 3) eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, mine);
                    const double eta_old = __shfl_sync(FULL, eb.eta, 0);
                    eta_acc = __shfl_sync(FULL, eb.eta, 1); eta_rej = __shfl_sync(FULL, eb.eta, 2);
                    k_acc = __shfl_sync(FULL, eb.k, 1); k_rej = __shfl_sync(FULL, eb.k, 2);
                    diffkT = (one ? dE0 : dE1) * p.beta + eta_acc - eta_old;
                }
                // one exponential pass for everything the rest of this move needs:
                //   lane 0  : acceptance probability            exp(-diffkT)
                //   lane 1/2: lattice-switch probability if this move is accepted / rejected
                //   lane 3/4: unbiased-histogram factor exp(eta - log_unbiased_norm) if accepted / rejected
                double arg = -diffkT;
                if (fuse_switch && (lane == 1 || lane == 2)) {
                    const bool a = (lane == 1);
                    arg = switch_arg(p, w.sc, w.lv, a ? Ea0 : Eb0, a ? Ea1 : Eb1, one, a ? eta_acc : eta_rej, Nd);
                }
                if (lane == 3) arg = eta_acc - p.log_unbiased_norm;
                if (lane == 4) arg = eta_rej - p.log_unbiased_norm;
                // min(1, exp(.)) for the probabilities (lanes 0-2); plain exp for the histogram factors
                const double ex = (arg > 0.0 && lane < 3) ? 1.0 : exp_call(fmin(arg, 700.0));
                const double zeta = rng.draw();
                const bool accepted = zeta < __shfl_sync(FULL, ex, 0);                         // :1145-1146
                if (accepted) {
                    sc->acc_r += 1;
                    const double dmu = fabs(dE0 - dE1) * p.beta;
                    if (dmu < sc->min_dmu) sc->min_dmu = dmu;
                    if (dmu > sc->max_dmu) sc->max_dmu = dmu;
                    sc->E[0] = Ea0;
                    if (NLAT == 2) { sc->E[1] = Ea1; sc->mu = mu_acc; }
                    commit_translation<NLAT>(w, imol, mo, mn);
                } else {
                    // reject: the reference restores by (x+t)-t, not by copy (mc_moves.F90:1186)
                    __syncwarp();
                    if (lane < 3 * NLAT) {
                        const int lat = (lane >= 3) ? 1 : 0, d = lane - 3 * lat;
                        const double* m = w.mv + lat * 6;
                        w.pos[(lat * 3 + d) * N + imol] = xs(m[d], m[3 + d]);
                    }
                    if (NLAT == 2) sc->mu = mu_rej;
                    __syncwarp();
                }
                sc->att_r += 1;
                // ====================== mc_update_wl_bins (mc_moves.F90:1597-1689) ======================
                const int kb = accepted ? k_acc : k_rej;
                if (bins_on && kb >= 1 && kb <= p.nbins) {
                    const double c = __ldg(S.hinc + kb - 1);
                    if (lane == 0) atomicAdd(hist + kb - 1, c);
                    if (p.samplerun) {
                        const double uf = __shfl_sync(FULL, ex, accepted ? 3 : 4);
                        if (lane == 0) atomicAdd(uhist + kb - 1, c * uf);
                    } else {
                        update_weights(smem, N, NLAT, p, S.binwidth, wgt, hist, kb);
                    }
                }
                // ====================== mc_lattice_switch (mc_moves.F90:1536-1594) ======================
                if (fuse_switch) {
                    const double compare = __shfl_sync(FULL, ex, accepted ? 1 : 2);
                    const double xs_ = rng.draw();
                    if (xs_ < compare) {
                        sc->acc_s += 1;
                        sc->mu = mu_paren(p, sc, Nd, w.lv[0]);
                        sc->ls = 3 - sc->ls;
                    }
                    sc->att_s += 1;
                } else if (do_switch) {
                    // weights may have moved in update_weights: the reference looks eta up again
                    rng.pos = lattice_switch_cold(smem, S, p, wi, NLAT, rng.pos);
                }
                continue;
            }
            // ---------------- rare move types ----------------
            if (xi < p.volP) {
                rng.pos = volume_move<NLAT>(smem, S, p, wi, rng.pos);
                const EtaBin eb = eta_bin(p, S.mubin, S.ginv, sc, wgt, sc->mu);
                if (bins_on && eb.k >= 1 && eb.k <= p.nbins) {
                    const double c = __ldg(S.hinc + eb.k - 1);
                    if (lane == 0) atomicAdd(hist + eb.k - 1, c);
                    if (p.samplerun) {
                        if (lane == 0) atomicAdd(uhist + eb.k - 1, c * exp(eb.eta - p.log_unbiased_norm));
                    } else {
                        update_weights(smem, N, NLAT, p, S.binwidth, wgt, hist, eb.k);
                    }
                }
                sc->att_v += 1;
            } else if (xi < p.swP) {
                if (NLAT == 2 && !dd_eq) rng.pos = lattice_switch_cold(smem, S, p, wi, NLAT, rng.pos);
            }
            if (do_switch) rng.pos = lattice_switch_cold(smem, S, p, wi, NLAT, rng.pos);
        }
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {                 // :253-255
            double a = sc->avgE[lat] + sc->E[lat];
            if (p.npt) a = a + p.pressure * sc->vol[lat];
            sc->avgE[lat] = a;
        }
        if (S.therm_int > 0 && cycle % S.therm_int == 0) {     // main.f90:200-223 (values only; formatted by the host)
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
    const uint64_t idx = *w.rngbase + (uint64_t)rng.pos;
    if (p.rng_mode == 1 && idx > S.fifo_len) err |= ERR_RNG_UNDERRUN;
    sc->rng_index = idx;
    sc->error |= err;
    __syncwarp();
    store_walker(S, wi, w, true);
}

}  // namespace mw
