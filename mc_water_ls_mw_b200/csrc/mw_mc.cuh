// mw_mc.cuh -- device side of the Monte-Carlo move loop (mc_moves.F90:217-255,
// :893-964, :966-1213, :1216-1534, :1536-1594, :1597-1689, :2187-2215) for one
// walker per warp, plus the global-memory layout of a walker.
#pragma once
#include "mw_device.cuh"

namespace mw {

// Per-walker scalars that persist between launches.
struct WalkerScalars {
    double E[2];            // model_energy(1:2)                        molint.F90:41
    double vol[2];          // volume(1:2)                              data_structures.f90:48
    double mu;              // ls_mu                                    mc_moves.F90:63
    double max_trans;       // mc_max_trans (Bohr) -- per walker: eq_adjust_mc tunes it per rank
    double dv_max;          // mc_dv_max (Bohr)
    double wl_factor;
    double mu_lo, mu_hi;    // my_mu_min, my_mu_max                     mc_moves.F90:108
    double avgE[2];         // average_energy                           mc_moves.F90:88
    double min_dmu, max_dmu;
    double refH[2];         // ref_enthalpy                             mc_moves.F90:88
    double sumhist;
    unsigned long long rng_index;   // next draw index (Philox) / FIFO position
    int ls;                 // active lattice, 1-based                  data_structures.f90:51
    int cycle;              // mc_cycle_num
    int acc_r, acc_v, acc_s, att_r, att_v, att_s;
    int start_bin, end_bin; // my_start_bin, my_end_bin (1-based)
    int in_window;          // walker_in_window
    int wl_invt_active;
    int wmin_zero;          // invariant "min(weight(window)) == 0" established
    int error;
};

// Run parameters shared by all walkers (kernel argument, by value).
struct McParams {
    double beta, pressure;
    double transP, volP, swP;
    double r_pos, r_neg, a_pos, a_neg, log_r_pos, log_r_neg;
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
    const double* fifo; // host-supplied random numbers (walker 0 only)
    unsigned long long fifo_len;
};

// ---------------------------------------------------------------- staging
__device__ inline void load_walker(const DeviceState& S, int wi, const WalkerView& w)
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
    __syncwarp();
}

__device__ inline void store_walker(const DeviceState& S, int wi, const WalkerView& w, bool lists)
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
}

// ---------------------------------------------------------------- order parameter / weights
// mc_moves.F90:2187-2215 (1-based bin)
__device__ __forceinline__ int mu_to_bin(const McParams& p, double mu)
{
    const int nb = p.nbins;
    if (fabs(mu) <= 0.5) return nb / 2 + 1;
    if (mu > 0.0) {
        const double arg = 1.0 - (mu - 0.5) * (1.0 - p.r_pos) / p.a_pos;
        return nb / 2 + 2 + (int)(log(arg) / p.log_r_pos);
    }
    const double arg = 1.0 - (fabs(mu) - 0.5) * (1.0 - p.r_neg) / p.a_neg;
    return nb / 2 - (int)(log(arg) / p.log_r_neg);
}

// mc_moves.F90:893-964.  wgt is this walker's weight array (global memory,
// read through L2 because the same warp updates it when generating weights).
__device__ inline double eta_weight(const McParams& p, const DeviceState& S, const WalkerScalars& sc,
                                    const double* wgt, double mu)
{
    if (!sc.in_window) return 0.0;          // undefined in the reference (:913); defined as 0
    if (mu < sc.mu_lo) return F_HUGE;
    if (mu > sc.mu_hi) return F_HUGE;
    int k = mu_to_bin(p, mu);
    k = min(max(k, 1), p.nbins);            // memory safety at mu == mu_max (reference would overrun)
    const double* w = wgt - 1;
    const double* bw = S.binwidth - 1;
    const double* mb = S.mubin - 1;
    if (!p.eta_interp) return __ldcg(w + k);
    int ka, kb, kr;                          // gradient between bins ka<kb, anchored at kr
    if (k == sc.start_bin)      { ka = k; kb = k + 1; kr = k; }
    else if (k == sc.end_bin)   { ka = k - 1; kb = k; kr = k; }
    else if (mu > __ldg(mb + k)){ ka = k; kb = k + 1; kr = k; }
    else                        { ka = k - 1; kb = k; kr = k - 1; }
    ka = max(ka, 1); kb = min(kb, p.nbins);
    const double wa = __ldcg(w + ka), wb = __ldcg(w + kb);
    const double g = 2.0 * (wb - wa) / (__ldg(bw + ka) + __ldg(bw + kb));
    const double wr = (kr == ka) ? wa : wb;
    return wr + (mu - __ldg(mb + kr)) * g;
}

// mu recomputed from scratch, parenthesised association (mc_moves.F90:1370-1372, :1525-1527, :1583-1585)
__device__ __forceinline__ double mu_paren(const McParams& p, const WalkerScalars& sc, double N, double lv12)
{
    double mu = (sc.E[0] + p.pressure * sc.vol[0]) - (sc.E[1] + p.pressure * sc.vol[1]);
    if (p.leshift) mu = mu - sc.refH[0] + sc.refH[1];
    return mu * p.beta - N * lv12;
}

// mc_moves.F90:1597-1689
__device__ inline void update_wl_bins(const McParams& p, const DeviceState& S, WalkerScalars& sc,
                                      double* wgt, double* hist, double* uhist, double eta_mu)
{
    if (sc.cycle < p.eq_mc_cycles) return;
    const int nb = p.nbins, lane = lane_id();
    const int k = mu_to_bin(p, sc.mu);
    if (k < 1 || k > nb) return;
    const double c = p.av_binwidth / __ldg(S.binwidth + k - 1);
    if (p.samplerun) {
        if (lane == 0) {
            atomicAdd(hist + k - 1, c);
            atomicAdd(uhist + k - 1, c * exp(eta_mu - p.log_unbiased_norm));
        }
        return;
    }
    if (lane == 0) atomicAdd(hist + k - 1, c);
    if (p.wl_swetnam) {
        // Swetnam's increment from the current histogram (:1636-1653)
        __syncwarp();
        sc.sumhist = sc.sumhist + 1.0;
        double f = 0.0;
        for (int i = 0; i < nb; ++i) {       // sequential order as in the reference; every lane redundantly
            const double binfrac = __ldg(S.binwidth + i) / (p.mu_max - p.mu_min - 1.0);
            const double d = __ldcg(hist + i) * __ldg(S.binwidth + i) / sc.sumhist - binfrac;
            f = f + d * d;
        }
        f = sqrt(f / (double)nb);
        f = log(f);
        f = f * p.wl_alpha * (double)nb;
        sc.wl_factor = fmin(f, p.orig_wl_factor);
    } else if (sc.wl_invt_active) {
        sc.wl_factor = fmin(sc.wl_factor, (double)nb / (double)(sc.cycle * S.N));
    }
    const double wk_old = __ldcg(wgt + k - 1);
    // weight(k) = weight(k) + av_binwidth*incr/binwidth(k)   (:1680)
    const double wk = wk_old + p.av_binwidth * sc.wl_factor / __ldg(S.binwidth + k - 1);
    __syncwarp();
    if (lane == 0) wgt[k - 1] = wk;
    __syncwarp();
    // minbin = minval(weight(start:end)); weight -= minbin (:1682-1685).  Subtracting
    // an exact 0 is a no-op, and the minimum stays 0 unless bin k was a zero bin.
    if (sc.wmin_zero && wk_old > 0.0) return;
    double mn = F_HUGE;
    for (int i = sc.start_bin - 1 + lane; i < sc.end_bin; i += 32) mn = fmin(mn, __ldcg(wgt + i));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL, mn, d));
    if (mn != 0.0)
        for (int i = sc.start_bin - 1 + lane; i < sc.end_bin; i += 32) wgt[i] = __ldcg(wgt + i) - mn;
    sc.wmin_zero = 1;
    __syncwarp();
}

// mc_moves.F90:1536-1594
__device__ inline void lattice_switch(const McParams& p, WalkerScalars& sc, WarpRng& rng, double N,
                                      double eta, double lv12, double lv21)
{
    const int ls = sc.ls, lsn = 3 - ls;
    const double b = p.beta;
    double diffkT;
    const bool one = (ls == 1);
    const double Es = one ? sc.E[0] : sc.E[1], En = one ? sc.E[1] : sc.E[0];
    const double Vs = one ? sc.vol[0] : sc.vol[1], Vn = one ? sc.vol[1] : sc.vol[0];
    const double lvn = one ? lv21 : lv12;              // log(volume(lsn)/volume(ls))
    if (p.npt) {
        diffkT = b * En - b * Es + b * p.pressure * (Vn - Vs) - N * lvn + eta - eta;
    } else {
        diffkT = b * En - b * Es + eta - eta;
    }
    if (p.leshift) {
        const double Rs = one ? sc.refH[0] : sc.refH[1], Rn = one ? sc.refH[1] : sc.refH[0];
        diffkT = diffkT - b * Rn + b * Rs;
    }
    const double compare = fmin(1.0, exp(-diffkT));
    const double x = rng.draw();
    if (x < compare) {
        sc.acc_s += 1;
        sc.mu = mu_paren(p, sc, N, lv12);
        sc.ls = lsn;
    }
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

__device__ inline void rescale_all(const DeviceState& S, int wi, const WalkerView& w, int lat)
{
    const int N = w.N, lane = lane_id();
    double rm[9], hm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { rm[k] = w.recip[lat * 9 + k]; hm[k] = w.cell[lat * 9 + k]; }
    double* P = w.pos + lat * 3 * N;
    double* R = S.ref + ((size_t)wi * w.nlat + lat) * 3 * N;
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

// refresh volume, recip matrix, image vectors of a lattice from the cell in shared memory
__device__ inline void refresh_cell(const WalkerView& w, int lat, double& vol, bool set_volume, int& err)
{
    double hm[9], rm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = w.cell[lat * 9 + k];
    if (set_volume) vol = fabs(determinant3(hm));
    recipmatrix3(hm, rm);
    __syncwarp();
    if (lane_id() == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) w.recip[lat * 9 + k] = rm[k];
    }
    __syncwarp();
    compute_ivects_warp(w, lat, err);
}

__device__ __forceinline__ double sel2(bool first, double a, double b) { return first ? a : b; }

// mc_moves.F90:1216-1534
template <int NLAT>
__device__ inline void volume_move(const McParams& p, const DeviceState& S, int wi, const WalkerView& w,
                                   WalkerScalars& sc, WarpRng& rng, const double* wgt,
                                   double& lv12, double& lv21, int& err)
{
    const int lane = lane_id();
    const double Nd = (double)w.N;
    double backupE[2] = {0.0, 0.0}, old_vol[2] = {0.0, 0.0}, newE[2] = {0.0, 0.0};
    // old cell + recip are parked in shared memory: save[0..17] = h, save[18..35] = recip
    double* save = w.save;
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        backupE[lat] = sc.E[lat];
        old_vol[lat] = sc.vol[lat];
        double hm[9], rm[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) hm[k] = w.cell[lat * 9 + k];
        recipmatrix3(hm, rm);                              // :1260-1262
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                w.recip[lat * 9 + k] = rm[k];
                save[lat * 9 + k] = hm[k];
                save[18 + lat * 9 + k] = rm[k];
            }
        }
    }
    __syncwarp();
    double x = rng.draw();
    const int idim = (int)xm(x, 3.0) + 1;
    x = rng.draw();
    const int jdim = (int)xm(x, 3.0) + 1;
    x = rng.draw();
    const double dh = xm(xs(xm(2.0, x), 1.0), sc.dv_max);
    if (lane < NLAT) {
        double* hm = w.cell + lane * 9;
        const double v = xa(MW_H(hm, idim, jdim), dh);
        if (idim != jdim) MW_H(hm, jdim, idim) = xa(MW_H(hm, jdim, idim), dh);
        MW_H(hm, idim, jdim) = v;
    }
    __syncwarp();
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        rescale_all(S, wi, w, lat);                        // recip = old cell's, h = new cell
        refresh_cell(w, lat, sc.vol[lat], true, err);
        newE[lat] = full_energy_warp(w, lat, err);
        sc.E[lat] = newE[lat];
    }
    double old_eta = 0.0, new_eta = 0.0, old_mu = 0.0;
    const bool one = (sc.ls == 1);
    double nlv12 = lv12, nlv21 = lv21;
    if (NLAT == 2) {
        old_eta = eta_weight(p, S, sc, wgt, sc.mu);
        old_mu = sc.mu;
        nlv12 = log(sc.vol[0] / sc.vol[1]); nlv21 = log(sc.vol[1] / sc.vol[0]);
        sc.mu = mu_paren(p, sc, Nd, nlv12);
        new_eta = eta_weight(p, S, sc, wgt, sc.mu);
    }
    x = rng.draw();
    const double dE = sel2(one, newE[0] - backupE[0], newE[1] - backupE[1]);
    const double Vs = sel2(one, sc.vol[0], sc.vol[1]), Vo = sel2(one, old_vol[0], old_vol[1]);
    const double diffkT = p.beta * dE + new_eta - old_eta + p.beta * p.pressure * (Vs - Vo) - Nd * log(Vs / Vo);
    const double compare = fmin(1.0, exp(-diffkT));
    if (x < compare) {
        sc.acc_v += 1;
        if (NLAT == 2) {
            const double dmu = fabs(old_mu - sc.mu);
            if (dmu < sc.min_dmu) sc.min_dmu = dmu;
            if (dmu > sc.max_dmu) sc.max_dmu = dmu;
        }
        lv12 = nlv12; lv21 = nlv21;
    } else {
        // :1434-1528: V,h <- old; rescale with recip(NEW) and h(OLD); recip <- old; ivects; E <- backup
        __syncwarp();
        if (lane < NLAT * 9) w.cell[lane] = save[lane];
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            sc.vol[lat] = old_vol[lat];
            rescale_all(S, wi, w, lat);
        }
        if (lane < NLAT * 9) w.recip[lane] = save[18 + lane];
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            compute_ivects_warp(w, lat, err);
            sc.E[lat] = backupE[lat];
            compute_bond_masks_warp(w, lat);               // positions moved by rounding; keep masks fresh
        }
        if (NLAT == 2) sc.mu = mu_paren(p, sc, Nd, lv12);
    }
}

// mc_moves.F90:966-1213
template <int NLAT>
__device__ inline void translation_move(const McParams& p, const DeviceState& S, int wi, const WalkerView& w,
                                        WalkerScalars& sc, WarpRng& rng, const double* wgt,
                                        double& eta_final, int& err)
{
    const int N = w.N, lane = lane_id();
    const bool one = (sc.ls == 1);
    double x = rng.draw();
    int imol = (int)xm(x, (double)N) + 1;
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
    x = xm(xm(x, sc.max_trans), r);
    y = xm(xm(y, sc.max_trans), r);
    z = xm(xm(z, sc.max_trans), r);

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
    double tv[2][3];
    tv[0][0] = sel2(one, x, bx); tv[0][1] = sel2(one, y, by); tv[0][2] = sel2(one, z, bz);
    tv[1][0] = sel2(one, bx, x); tv[1][1] = sel2(one, by, y); tv[1][2] = sel2(one, bz, z);
    double pnew[2][3];
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        const double* P = w.pos + lat * 3 * N;
        pnew[lat][0] = xa(P[imol], tv[lat][0]);
        pnew[lat][1] = xa(P[N + imol], tv[lat][1]);
        pnew[lat][2] = xa(P[2 * N + imol], tv[lat][2]);
    }

    double eo[2] = {0.0, 0.0}, en[2] = {0.0, 0.0};
    uint32_t mo[2] = {0, 0}, mn[2] = {0, 0};
    local_energies_warp<NLAT, true>(w, imol, pnew, eo, en, mo, mn, err);

    double backup[2] = {0.0, 0.0}, dE[2] = {0.0, 0.0};
#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) {
        backup[lat] = sc.E[lat];
        sc.E[lat] = sc.E[lat] - eo[lat];
        sc.E[lat] = sc.E[lat] + en[lat];
        dE[lat] = en[lat] - eo[lat];
    }
    double diffkT, mu_acc = sc.mu, mu_rej = sc.mu, eta_acc = 0.0, eta_rej = 0.0;
    if (NLAT == 1) {
        diffkT = p.beta * dE[0];
    } else {
        const double dm = (dE[0] - dE[1]) * p.beta;
        mu_acc = sc.mu + dm;
        mu_rej = mu_acc - dm;
        // three weight look-ups in parallel lanes: eta(mu), eta(mu_acc), eta(mu_rej)
        const double mine = (lane == 0) ? sc.mu : (lane == 1) ? mu_acc : mu_rej;
        double e = 0.0;
        if (lane < 3) e = eta_weight(p, S, sc, wgt, mine);
        const double eta_old = __shfl_sync(FULL, e, 0);
        eta_acc = __shfl_sync(FULL, e, 1);
        eta_rej = __shfl_sync(FULL, e, 2);
        diffkT = sel2(one, dE[0], dE[1]) * p.beta + eta_acc - eta_old;
    }
    const double zeta = rng.draw();
    if (zeta < fmin(1.0, exp(-diffkT))) {
        sc.acc_r += 1;
        const double dmu = fabs(dE[0] - dE[1]) * p.beta;
        if (dmu < sc.min_dmu) sc.min_dmu = dmu;
        if (dmu > sc.max_dmu) sc.max_dmu = dmu;
        sc.mu = mu_acc; eta_final = eta_acc;
        // commit: position, own bond mask, and the reverse bits of bonds that formed / broke
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            double* P = w.pos + lat * 3 * N;
            if (lane < 3) P[lane * N + imol] = (lane == 0) ? pnew[lat][0] : (lane == 1) ? pnew[lat][1] : pnew[lat][2];
            uint32_t changed = mo[lat] ^ mn[lat];
            if (lane == 0) w.bmask[lat * N + imol] = mn[lat];
            const int nv = w.niv[lat];
            while (changed) {
                const int s = __ffs(changed) - 1; changed &= changed - 1;
                const uint32_t e = w.list[((size_t)lat * N + imol) * LC + s];
                const int j = e & 1023, img = e >> 10;
                const uint32_t target = ((uint32_t)inverse_image(img, nv) << 10) | (uint32_t)imol;
                const int nnj = w.nn[lat * N + j];
                const uint32_t e2 = (lane < nnj) ? w.list[((size_t)lat * N + j) * LC + lane] : 0xffffffffu;
                const uint32_t hit = __ballot_sync(FULL, e2 == target);
                if (hit && lane == 0) {
                    const int s2 = __ffs(hit) - 1;
                    const uint32_t bit = (mn[lat] >> s) & 1u;
                    w.bmask[lat * N + j] = (w.bmask[lat * N + j] & ~(1u << s2)) | (bit << s2);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // reject: the reference restores by (x+t)-t, not by copy (mc_moves.F90:1186)
        __syncwarp();
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            double* P = w.pos + lat * 3 * N;
            const double pn = (lane == 0) ? pnew[lat][0] : (lane == 1) ? pnew[lat][1] : pnew[lat][2];
            const double tt = (lane == 0) ? tv[lat][0] : (lane == 1) ? tv[lat][1] : tv[lat][2];
            if (lane < 3) P[lane * N + imol] = xs(pn, tt);
            sc.E[lat] = backup[lat];
        }
        if (NLAT == 2) { sc.mu = mu_rej; eta_final = eta_rej; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- the walker kernel
// One warp (= one CTA of 32 threads) per walker; ncycles MC cycles of the hot
// part of mc_cycle (mc_moves.F90:117-255).
template <int NLAT>
__global__ void __launch_bounds__(32) k_mc_run(DeviceState S, McParams p, int ncycles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int wi = blockIdx.x;
    if (wi >= S.W) return;
    const int lane = lane_id();
    const WalkerView w = carve_walker(smem, S.N, NLAT);
    load_walker(S, wi, w);
    WalkerScalars sc = S.scal[wi];
    const int N = S.N;
    const double Nd = (double)N;
    double* wgt = S.weight + (size_t)wi * S.NB;
    double* hist = S.hist + (size_t)wi * S.NB;
    double* uhist = S.uhist + (size_t)wi * S.NB;
    int err = sc.error;
    if (p.prob_error) err |= ERR_PROB;

#pragma unroll
    for (int lat = 0; lat < NLAT; ++lat) compute_bond_masks_warp(w, lat);

    WarpRng rng;
    rng.mode = p.rng_mode; rng.seed = p.seed; rng.stream = p.stream0 + (uint32_t)wi;
    rng.fifo = S.fifo; rng.fifo_len = S.fifo_len;
    rng.init(sc.rng_index);

    double lv12 = 0.0, lv21 = 0.0;
    if (NLAT == 2) { lv12 = log(sc.vol[0] / sc.vol[1]); lv21 = log(sc.vol[1] / sc.vol[0]); }

    for (int cyc = 0; cyc < ncycles && !(err & (ERR_WINDOW | ERR_PROB)); ++cyc) {
        sc.cycle += 1;
        if (p.dd) {                                            // mc_moves.F90:181-210
            if (sc.cycle < p.eq_mc_cycles) sc.in_window = (sc.mu > sc.mu_lo) && (sc.mu < sc.mu_hi);
            else if (sc.cycle == p.eq_mc_cycles) { if (!sc.in_window) { err |= ERR_WINDOW; break; } }
            else sc.in_window = 1;
        }
        if (sc.cycle % p.list_update_int == 0) {               // :218-222
#pragma unroll
            for (int lat = 0; lat < NLAT; ++lat) {
                compute_neighbours_warp(w, lat, err);
                compute_bond_masks_warp(w, lat);
            }
        }
        const bool dd_eq = p.dd && (sc.cycle < p.eq_mc_cycles);
        for (int imove = 0; imove < N; ++imove) {              // :224-250
            const double xi = rng.draw();
            double eta = 0.0;
            bool eta_known = false;
            if (xi < p.transP) {
                translation_move<NLAT>(p, S, wi, w, sc, rng, wgt, eta, err);
                eta_known = (NLAT == 2);
                if (p.samplerun && !eta_known) eta = eta_weight(p, S, sc, wgt, sc.mu);
                update_wl_bins(p, S, sc, wgt, hist, uhist, eta);
                sc.att_r += 1;
            } else if (xi < p.volP) {
                volume_move<NLAT>(p, S, wi, w, sc, rng, wgt, lv12, lv21, err);
                eta = eta_weight(p, S, sc, wgt, sc.mu);
                update_wl_bins(p, S, sc, wgt, hist, uhist, eta);
                sc.att_v += 1;
            } else if (xi < p.swP) {
                if (NLAT == 2 && !dd_eq) {
                    lattice_switch(p, sc, rng, Nd, eta_weight(p, S, sc, wgt, sc.mu), lv12, lv21);
                    sc.att_s += 1;
                }
            }
            if (NLAT == 2 && p.always_switch && !dd_eq) {
                // weights may have moved in update_wl_bins when generating them: look eta up again then
                if (!eta_known || !p.samplerun) eta = eta_weight(p, S, sc, wgt, sc.mu);
                lattice_switch(p, sc, rng, Nd, eta, lv12, lv21);
                sc.att_s += 1;
            }
        }
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {                 // :253-255
            sc.avgE[lat] = sc.avgE[lat] + sc.E[lat];
            if (p.npt) sc.avgE[lat] = sc.avgE[lat] + p.pressure * sc.vol[lat];
        }
    }
    if (rng.underrun) err |= ERR_RNG_UNDERRUN;
    sc.rng_index = rng.index();
    sc.error = err;
    store_walker(S, wi, w, true);
    if (lane == 0) S.scal[wi] = sc;
}

}  // namespace mw
