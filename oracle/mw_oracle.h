/*
 * mw_oracle.h -- CPU ORACLE for the mW lattice-switch Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (keb721/mc_water_ls_mw) algorithm for the hot path, used as the checker in
 * tests/, in __graft_entry__.smoke() and as the `cpu_baseline` / `--impl
 * reference` leg of bench.py.  Nothing in the product (mc_water_ls_mw_b200/,
 * include/) may include, link or call it.
 *
 * Pinning status: the reference cannot be compiled in this environment (no
 * Fortran compiler, no MPI).  The oracle is pinned against every
 * golden datum the reference ships for this path (bin grid = column 1 of
 * examples/ice1_sample/eta_weights.dat, the float32-promoted literal in that
 * file's header, neighbour-list lengths 16/17, and the 1d-10 Ha
 * delta(full)==delta(local) invariant of mc_moves.F90:1094-1102).  Energies,
 * neighbour-list contents and accept/reject counts are NOT stored anywhere in
 * the reference: for those quantities parity is "unpinned" and is defined
 * oracle <-> GPU.
 *
 * Array layouts are the reference's (Fortran column-major) ones:
 *   ljr(3,1,nwater,nlat)      -> ljr[(ils*nwater+imol)*3 + d]
 *   hmatrix(3,3,nlat)         -> h[ils*9 + col*3 + row]
 *   nn(nwater,nlat)           -> nn[ils*nwater+imol]
 *   jn(maxneigh,nwater,nlat)  -> jn[(ils*nwater+imol)*ORC_MAXNEIGH + ln]   (values 1-based)
 *   vn(...)                   -> same, 1-based image index into ivect
 *   ivect(3,maxnivect,nlat)   -> ivect[(ils*ORC_MAXIVECT + k)*3 + d]
 * All C-side indices (ils, imol) are 0-based; stored jn/vn values stay 1-based.
 */
#ifndef MW_ORACLE_H
#define MW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXNEIGH 50   /* molint.F90:79 */
#define ORC_MAXIVECT 125  /* (2*2+1)^3: room for im=jm=km=2 */

/* ---- random stream (random.f90:87-102 is the compiler's random_number, unpinned;
 *      the oracle defines its own documented stream, see DESIGN.md "RNG") ---- */
typedef struct orc_rng {
    int       mode;       /* 0 = Philox-4x32-10 counter stream, 1 = host FIFO */
    uint64_t  seed;       /* philox key */
    uint32_t  stream;     /* philox counter word 2 (walker id) */
    uint64_t  index;      /* next draw index (philox) */
    const double *fifo;   /* FIFO of U[0,1) numbers (mode 1) */
    int64_t   fifo_len;
    int64_t   fifo_pos;
    int       underrun;   /* set when the FIFO ran dry */
} orc_rng;

/* ---- run parameters (userparams.f90, internal units: Bohr, Hartree, a.u. pressure) ---- */
typedef struct orc_params {
    double temperature;      /* K */
    double pressure;         /* a.u. (atm / aup_to_atm) */
    int    npt;              /* mc_ensemble == 'npt' */
    double mc_max_trans;     /* Bohr */
    double mc_dv_max;        /* Bohr */
    double mc_target_ratio;
    double wl_factor;
    int    wl_swetnam;
    double wl_alpha;
    int    eta_interp;
    int    samplerun;
    int    leshift;
    int    nbins;
    double mu_min, mu_max;
    int    allow_switch, allow_vol, allow_trans;
    double mc_trans_prob, mc_vol_prob, mc_switch_prob;
    int    mc_always_switch;
    int    list_update_int;
    int    eq_mc_cycles;
    int    max_mc_cycles;
    int    eq_adjust_mc;
    int    monitor_int;
    int    dd;               /* parallel_strategy == 'dd' */
    int    window_overlap;
    double input_ref_enthalpy[2];
    int    ls;               /* initially active lattice, 1-based */
} orc_params;

typedef struct orc_system {
    int nwater, nlat;
    /* model (data_structures.f90:39-51) */
    double *ljr, *ref_ljr;
    double h[18], ref_h[18], recip[18];
    double volume[2];
    int    ls;                         /* active lattice, 1-based */
    /* energy module (molint.F90:41-81) */
    double model_energy[2];
    int    nivect[2];
    double *ivect;                     /* [nlat][ORC_MAXIVECT][3] */
    int   *nn, *jn, *vn;
    int    nn_warnings;                /* molint.F90:552-554 count of nn<16 */
    /* mc_moves module state */
    orc_params p;
    int    rank, size;
    int    mc_cycle_num;
    int    acc_r, acc_v, acc_s, att_r, att_v, att_s;   /* mc_moves.F90:45-52 */
    int   *mc_translations;
    double ls_mu;
    double ref_enthalpy[2], average_energy[2];
    double max_dmu, min_dmu;
    double *histogram, *weight, *unbiased_hist, *mu_bin, *binwidth;
    double av_binwidth, log_unbiased_norm;
    double a_pos, a_neg, r_pos, r_neg, s_pos, s_neg;
    double orig_wl_factor, wl_factor;
    int    wl_invt_active;
    double sumhist;
    double my_mu_max, my_mu_min;
    int    my_start_bin, my_end_bin;   /* 1-based */
    int    walker_in_window;
    double transP, volP, swP;
    int    firstpass;
    int    firstcycle;                 /* mc_moves.F90:85  "is this the original wl_factor" */
    int    histogram_reset;            /* mc_moves.F90:1957 (saved local of mc_check_flatness) */
    /* comms module state (comms_mpi.f90:73-104) */
    double *eta_last_sync, *hist_last_sync, *uhist_last_sync;
    orc_rng rng;
    int    error;                      /* non-zero == a reference `stop` was hit */
} orc_system;

/* constants (constants.f90, molint.F90:64-74) */
double orc_const(const char *name);

/* model / energy module */
orc_system *orc_create(int nwater, int nlat);
void   orc_destroy(orc_system *s);
void   orc_set_config(orc_system *s, const double *ljr_bohr, const double *hmatrix_bohr); /* init.f90:80-106 */
double orc_determinant(const double *m);                       /* util.f90:16-41 */
void   orc_recipmatrix(const double *h, double *recip);        /* util.f90:43-77 */
void   orc_energy_init(orc_system *s);                         /* molint.F90:91-153 */
void   orc_compute_ivects(orc_system *s, int ils);             /* molint.F90:174-217 */
void   orc_compute_neighbours(orc_system *s, int ils);         /* molint.F90:501-559 */
void   orc_compute_model_energy(orc_system *s, int ils);       /* molint.F90:407-499 */
double orc_compute_local_real_energy(orc_system *s, int imol, int ils); /* molint.F90:220-404 */

/* mc_moves module */
void   orc_params_default(orc_params *p);                      /* userparams.f90:14-79 */
int    orc_mc_init(orc_system *s, const orc_params *p, int rank, int size,
                   const double *file_weights, int n_file_weights, double file_wl_factor); /* main.f90:146-175 + mc_moves.F90:504-877 */
double orc_eta_weight(orc_system *s, double mu);               /* mc_moves.F90:893-964 */
int    orc_mu_to_bin(const orc_system *s, double mu);          /* mc_moves.F90:2187-2215 (1-based) */
void   orc_mc_water_translation(orc_system *s);                /* mc_moves.F90:966-1213 */
void   orc_mc_volume(orc_system *s);                           /* mc_moves.F90:1216-1534 */
void   orc_mc_lattice_switch(orc_system *s);                   /* mc_moves.F90:1536-1594 */
void   orc_mc_update_wl_bins(orc_system *s);                   /* mc_moves.F90:1597-1689 */
int    orc_mc_cycle(orc_system *s);                            /* mc_moves.F90:117-255 (hot part) */
int    orc_mc_run(orc_system *s, int ncycles);
void   orc_mc_monitor(orc_system *s);                          /* state effects of mc_moves.F90:1722-1732,1786-1810 */
void   orc_mc_chain_sync(orc_system *s);                       /* mc_moves.F90:2217-2416 */
void   orc_allreduce_bins(orc_system **walkers, int nwalkers); /* comms_mpi.f90:244-277,461-530 over in-process walkers */

/* mc_checkpoint_load (mc_moves.F90:403-501) + the restart refresh of mc_init (:842-862); ref_h and the
 * neighbour lists stay what the start-up sequence made them, as in the reference */
void   orc_mc_restore(orc_system *s, int mc_cycle_num, double mc_max_trans, double mc_dv_max, double wl_factor,
                      int wl_invt_active, int ls, const double *histogram, const double *weight,
                      const double *unbiased_hist, const double *hmatrix, const double *ref_ljr, const double *ljr);

/* ---- periodic bookkeeping that consumes the reduced arrays ---- */
typedef struct orc_flat_params {       /* userparams.f90:33-36 */
    int    wl_schedule;                /* 0 = within tol of mean, 1 = min visits, 2 = above (1-tol) of mean */
    int    wl_minhist;
    double wl_flattol;
    int    wl_useinvt;
} orc_flat_params;
typedef struct orc_flat_report {       /* what mc_check_flatness writes to the log, for walker (rank) 0 */
    int    checked;                    /* 0: returned at the samplerun / empty-histogram guard */
    int    hist_reset;                 /* the one-off histogram reset of :1973-1980 happened in this call */
    int    flat;
    int    invt_switched;
    double mean, max_pct, min_pct;
    double wl_factor;                  /* after the call */
} orc_flat_report;
/* mc_check_flatness (mc_moves.F90:1936-2185), state effects on every walker (= MPI rank); the file
 * dumps of :2067-2101 / :2147-2178 are left to the caller.  Returns non-zero on the reference's `stop`. */
int    orc_mc_check_flatness(orc_system **walkers, int nwalkers, const orc_flat_params *fp, orc_flat_report *rep);
/* mc_compute_deltaG_from_hist (mc_moves.F90:2498-2621): returns deltaG in kT (whole box), normP(nbins) */
double orc_mc_deltaG_from_hist(orc_system **walkers, int nwalkers, double *normP);
/* comms_join_uhist / comms_join_eta (comms_mpi.f90:299-375, :377-459): stitch the windows of dd runs */
void   orc_join_uhist(orc_system **walkers, int nwalkers, int overlap, double *joined);
void   orc_join_eta(orc_system **walkers, int nwalkers, int overlap, double *joined);

/* RNG */
void   orc_rng_philox(orc_rng *r, uint64_t seed, uint32_t stream, uint64_t start_index);
void   orc_rng_fifo(orc_rng *r, const double *u, int64_t n);
double orc_rng_draw(orc_rng *r);
void   orc_philox_block(uint64_t seed, uint32_t stream, uint64_t block, double out[2]);
void   orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* batch helpers for the CPU baseline (OpenMP over independent walkers) */
int    orc_mc_run_many(orc_system **walkers, int nwalkers, int ncycles, int nthreads);
void   orc_model_energy_many(orc_system **walkers, int nwalkers, int nthreads, double *e_out);
int    orc_max_threads(void);

/* name-based accessors for the ctypes test harness */
double *orc_ptr_d(orc_system *s, const char *name);
int    *orc_ptr_i(orc_system *s, const char *name);
double  orc_get_d(const orc_system *s, const char *name);
int64_t orc_get_i(const orc_system *s, const char *name);
void    orc_set_d(orc_system *s, const char *name, double v);
void    orc_set_i(orc_system *s, const char *name, int64_t v);
void    orc_set_rng_philox(orc_system *s, uint64_t seed, uint32_t stream, uint64_t start_index);
void    orc_set_rng_fifo(orc_system *s, const double *u, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
