/*
 * mwgpu.h -- C ABI of the B200-native hot path of keb721/mc_water_ls_mw.
 *
 * Drop-in boundary: the reference has no FFI layer; the seam is the Fortran
 * module `energy` (molint.F90:22-37) plus the move loop of mc_moves.F90 that
 * drives it.  Every entry point below names the reference interface it
 * replaces.  All arrays use the reference's own (Fortran column-major) layouts
 * and 1-based molecule / lattice / image numbers:
 *
 *   ljr(3,1,nwater,nlat), ref_ljr(...)   data_structures.f90:39,116
 *   hmatrix(3,3,nlat)                    data_structures.f90:42
 *   nn(nwater,nlat), jn(maxneigh,nwater,nlat), vn(maxneigh,nwater,nlat)   molint.F90:79-81
 *   nivect(nlat), ivect(3,maxnivect,nlat)                                  molint.F90:44-45
 *
 * A context owns `nwalkers` independent walker boxes on one CUDA device (one
 * walker == one MPI rank of the reference).  The fine-grained calls take a
 * 0-based `walker` index; the batched calls act on all walkers at once.
 * Every function returns 0 on success and a non-zero code on failure, with
 * text available from mwgpu_last_error() (the reference `stop`s with a
 * message; a Fortran shim does `if (ierr/=0) stop`).  A context is not
 * thread-safe; distinct contexts may be driven from distinct host threads.
 *
 * There is NO CPU fallback: every call fails loudly when CUDA is unavailable.
 */
#ifndef MWGPU_H
#define MWGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MWGPU_MAXNEIGH 50      /* molint.F90:79  leading dimension of jn/vn in the ABI (NOT the device capacity, see below) */
#define MWGPU_MAXIVECT 32      /* image vectors supported per lattice (27 in all reference decks) */
#define MWGPU_LIST_SLOTS 32    /* neighbours per molecule held on the device (reference decks: 16..23) */
/* Capacity envelope of the device path (the reference allocates maxneigh = 50 list slots and checks nothing,
 * molint.F90:79; here every limit is checked and reported through the MWGPU_ERR_* bits instead of a result):
 *   - 32 list neighbours per molecule within 1.18*a*sigma        -> MWGPU_ERR_LIST_OVERFLOW   (decks: <= 23)
 *   - 32 image vectors per lattice (cell wider than the cut-off) -> MWGPU_ERR_IVECT_OVERFLOW  (decks: 27)
 *   - a molecule must not neighbour its own periodic image       -> MWGPU_ERR_SELF_IMAGE
 *   - bonds inside the cut-off of the moved molecule per lattice: 31 per evaluated variant in the warp-per-lattice
 *     kernel (old and new position share a 32-record table and fall back to one variant at a time); 64 in total over
 *     both lattices and both variants in the one-warp kernel     -> MWGPU_ERR_BOND_OVERFLOW   (decks: <= 11 per variant)
 *   - boxes of up to 1024 molecules; the warp-per-lattice kernel serves boxes of up to 64 (all reference decks: 48)
 * The reference's compile-time variant MINU (lattice switch folded into every move, mc_moves.F90:1119-1140,
 * :1385-1401) is not implemented. */

typedef struct mwgpu_ctx mwgpu_ctx;

/* run parameters: userparams.f90:14-79 in INTERNAL units (Bohr, Hartree, a.u.
 * pressure) i.e. after the conversions of io.f90:165,185-186 */
typedef struct mwgpu_mc_params {
    double temperature;        /* K */
    double pressure;           /* a.u. */
    int    npt;                /* mc_ensemble == 'npt' */
    double mc_max_trans;       /* Bohr */
    double mc_dv_max;          /* Bohr */
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
    int    dd;                 /* parallel_strategy == 'dd' */
    int    window_overlap;
    double input_ref_enthalpy[2];
    int    ls;                 /* initially active lattice (1-based) */
} mwgpu_mc_params;

/* per-walker observables (what main.f90:200-223 and mc_moves.F90:1722-1792 read) */
typedef struct mwgpu_walker_state {
    double model_energy[2];
    double volume[2];
    double ls_mu;
    double mc_max_trans, mc_dv_max;
    double wl_factor;
    double my_mu_min, my_mu_max;
    double average_energy[2];
    double min_dmu, max_dmu;
    double ref_enthalpy[2];
    int64_t rng_index;         /* random numbers consumed so far (draw index) */
    int    ls;
    int    mc_cycle_num;
    int    accepted[3];        /* translations, volume moves, switches   mc_moves.F90:45-47 */
    int    attempted[3];       /*                                        mc_moves.F90:50-52 */
    int    my_start_bin, my_end_bin;
    int    walker_in_window;
    int    error;              /* MWGPU_ERR_* bits */
    int    wl_invt_active;     /* 1/t increment active (mc_moves.F90:87; part of the checkpoint, :366) */
} mwgpu_walker_state;

enum {
    MWGPU_ERR_LIST_OVERFLOW  = 1,
    MWGPU_ERR_IVECT_OVERFLOW = 2,
    MWGPU_ERR_BOND_OVERFLOW  = 4,
    MWGPU_ERR_ITEM_OVERFLOW  = 8,
    MWGPU_ERR_SELF_IMAGE     = 16,
    MWGPU_ERR_RNG_UNDERRUN   = 32,
    MWGPU_ERR_WINDOW         = 64,
    MWGPU_ERR_PROB           = 128
};

/* ---- lifecycle ------------------------------------------------------------------- */
/* replaces create_model / energy_init allocation (data_structures.f90:66-144, molint.F90:108-143) */
int  mwgpu_create(int nwater, int nlat, int nwalkers, int device, mwgpu_ctx **out);
/* replaces energy_deinit / destroy_model (molint.F90:155-171) */
void mwgpu_destroy(mwgpu_ctx *ctx);
const char *mwgpu_last_error(void);
int  mwgpu_device_count(void);
int  mwgpu_num_walkers(const mwgpu_ctx *ctx);

/* ---- model state: upload after read_xmol / checkpoint load, download before output --- */
/* walker >= 0: one walker; walker == -1: the same configuration is given to every walker */
int  mwgpu_upload(mwgpu_ctx *ctx, int walker, const double *ljr, const double *ref_ljr, const double *hmatrix);
int  mwgpu_download(mwgpu_ctx *ctx, int walker, double *ljr, double *ref_ljr, double *hmatrix);
/* all walkers at once: arrays carry a leading walker dimension, e.g. ljr(3,1,nwater,nlat,nwalkers) */
int  mwgpu_upload_all(mwgpu_ctx *ctx, const double *ljr, const double *ref_ljr, const double *hmatrix);
int  mwgpu_download_all(mwgpu_ctx *ctx, double *ljr, double *ref_ljr, double *hmatrix);

/* ---- module energy (molint.F90) --------------------------------------------------- */
/* energy_init, molint.F90:91-153: volume, recip matrix, image vectors, lists, energies -- all walkers.
 * When mwgpu_mc_init() has been called before (restart path, mc_moves.F90:842-862) the order
 * parameter ls_mu is recomputed from the fresh energies as well. */
int  mwgpu_energy_init(mwgpu_ctx *ctx);
/* compute_ivects(ils), molint.F90:174-217; nivect/ivect(3,MWGPU_MAXIVECT) may be NULL */
int  mwgpu_compute_ivects(mwgpu_ctx *ctx, int walker, int ils, int *nivect, double *ivect);
/* compute_neighbours(ils), molint.F90:501-559; nn(nwater), jn/vn(MWGPU_MAXNEIGH,nwater) may be NULL */
int  mwgpu_compute_neighbours(mwgpu_ctx *ctx, int walker, int ils, int *nn, int *jn, int *vn);
/* compute_model_energy(ils), molint.F90:407-499; result also kept as model_energy(ils) */
int  mwgpu_compute_model_energy(mwgpu_ctx *ctx, int walker, int ils, double *energy);
/* compute_local_real_energy(imol,ils), molint.F90:220-404 */
int  mwgpu_compute_local_real_energy(mwgpu_ctx *ctx, int walker, int imol, int ils, double *energy);
/* the same for every molecule of a lattice in one launch: energy(nwater) */
int  mwgpu_compute_local_real_energy_all(mwgpu_ctx *ctx, int walker, int ils, double *energy);
/* batched: compute_neighbours / compute_model_energy for every lattice of every walker.
 * energies(nlat,nwalkers) may be NULL */
int  mwgpu_compute_neighbours_all(mwgpu_ctx *ctx);
int  mwgpu_compute_model_energy_all(mwgpu_ctx *ctx, double *energies);
/* read the current lists without rebuilding them */
int  mwgpu_get_neighbours(mwgpu_ctx *ctx, int walker, int ils, int *nn, int *jn, int *vn);

/* ---- mc_moves: initialisation (main.f90:146-175, mc_moves.F90:504-877) ------------ */
/* walker w of this context is rank `first_rank + w` of `size` ranks (windows in dd mode,
 * log_unbiased_norm).  file_weights = column 2 of eta_weights.dat (or NULL), file_wl_factor
 * its header value.  Requires mwgpu_energy_init(). */
int  mwgpu_mc_init(mwgpu_ctx *ctx, const mwgpu_mc_params *p, int first_rank, int size,
                   const double *file_weights, int n_file_weights, double file_wl_factor);
/* random stream (random.f90:87-102 draws from the compiler's generator, unpinned):
 * Philox-4x32-10 keyed by seed, walker w uses stream first_stream + w, first draw = start_index
 * (main.f90:79-81 burns 1 000 000 draws) ... */
int  mwgpu_mc_set_rng_philox(mwgpu_ctx *ctx, uint64_t seed, uint32_t first_stream, uint64_t start_index);
/* next draw index of one walker (walker -1: all): a restart must advance the stream past what the first segment
 * consumed -- mwgpu_mc_get_state().rng_index at checkpoint time -- or use another seed; the reference re-seeds from
 * the clock (random.f90:62-63) and its checkpoint carries no generator state (mc_moves.F90:353-379) */
int  mwgpu_mc_set_rng_index(mwgpu_ctx *ctx, int walker, uint64_t index);
/* ... or a host FIFO of U[0,1) numbers for walker 0 of a 1-walker context: the device consumes
 * them in the reference's draw order; mwgpu_mc_get_state().rng_index tells how many were used */
int  mwgpu_mc_set_rng_fifo(mwgpu_ctx *ctx, const double *u, int64_t n);

/* ---- mc_moves: the hot loop -------------------------------------------------------- */
/* ncycles x mc_cycle (mc_moves.F90:117-255: list refresh, nwater trial moves incl.
 * mc_water_translation :966, mc_volume :1216, mc_lattice_switch :1536, mc_update_wl_bins :1597,
 * average-energy accumulation) for every walker.  The periodic bookkeeping of :257-316 is driven
 * by the host between calls through the functions below. */
int  mwgpu_mc_run(mwgpu_ctx *ctx, int ncycles);
/* Which walker kernel mwgpu_mc_run uses: 0 = automatic (default), 1 = one warp per walker, 2 = two warps per walker,
 * one per lattice -- the per-lattice loops of mc_moves.F90:1007-1018, :1076-1090 side by side (two lattices, up to 64
 * molecules), 4 = two warps per lattice, which also split the bond / triplet loops of compute_local_real_energy
 * (molint.F90:276-404) of one lattice (at most four walkers per SM; never chosen automatically: it shortens a walker's
 * step by ~1.5 %).  All produce the same chain: positions / lists / counters bit for bit, energies to 1e-11. */
int  mwgpu_mc_set_kernel(mwgpu_ctx *ctx, int warps_per_walker);
int  mwgpu_mc_run_async(mwgpu_ctx *ctx, int ncycles);     /* no host synchronisation */
int  mwgpu_synchronize(mwgpu_ctx *ctx);

int  mwgpu_mc_get_state(mwgpu_ctx *ctx, int walker, mwgpu_walker_state *out);
int  mwgpu_mc_get_states(mwgpu_ctx *ctx, mwgpu_walker_state *out /* [nwalkers] */);
int  mwgpu_mc_get_translations(mwgpu_ctx *ctx, int walker, int *mc_translations /* (nwater) */);
/* weight / histogram / unbiased_hist (nbins) of one walker; any pointer may be NULL */
int  mwgpu_mc_get_bins(mwgpu_ctx *ctx, int walker, double *weight, double *histogram, double *unbiased_hist);
int  mwgpu_mc_set_bins(mwgpu_ctx *ctx, int walker, const double *weight, const double *histogram, const double *unbiased_hist);
int  mwgpu_mc_get_grid(mwgpu_ctx *ctx, double *mu_bin, double *binwidth, double *scalars /* r_pos,r_neg,av_binwidth,log_unbiased_norm */);
int  mwgpu_mc_set_wl_factor(mwgpu_ctx *ctx, int walker, double wl_factor, int wl_invt_active);  /* walker -1: all */
int  mwgpu_mc_set_active_lattice(mwgpu_ctx *ctx, int walker, int ls);
/* state effects of mc_monitor_stats (mc_moves.F90:1722-1732 step-size adjustment during
 * equilibration, :1786-1792 energy re-synchronisation, :1797-1810 counter reset), all walkers.
 * mc_monitor_stats also re-synchronises histogram / weights / unbiased histogram in 'mw' runs (:1813-1821): follow
 * this call with mwgpu_comms_allreduce_bins() (a no-op right after the mpi_sync_int merge, a real merge whenever
 * monitor_int is not a multiple of mpi_sync_int). */
int  mwgpu_mc_monitor(mwgpu_ctx *ctx);
/* mc_check_chain_synchronisation (mc_moves.F90:2217-2416), all walkers */
int  mwgpu_mc_chain_sync(mwgpu_ctx *ctx);

/* ---- comms (comms_mpi.f90:244-277, :461-530): delta-since-last-sync all-reduce ----- */
/* Sums the increments of weight / histogram / (samplerun) unbiased_hist over all walkers of the
 * context and, when mwgpu_comms_init() was called, over all ranks with one NCCL all-reduce on
 * the context's stream; every walker ends with base + total and re-bases. */
int  mwgpu_comms_allreduce_bins(mwgpu_ctx *ctx);
/* comms_set_histogram / comms_set_uhistogram (comms_mpi.f90:533-565): re-base after a reset */
int  mwgpu_comms_set_hist_base(mwgpu_ctx *ctx, const double *histogram, const double *unbiased_hist);
/* NCCL bootstrap: rank 0 creates an id (128 bytes) and ships it to the other ranks out of band */
int  mwgpu_comms_get_unique_id(void *id128);
int  mwgpu_comms_init(mwgpu_ctx *ctx, int nranks, int rank, const void *id128);
/* the staging buffer [3][nbins_padded] of summed increments, for an external all-reduce
 * (e.g. torch.distributed): reduce_local -> all-reduce of `*count` doubles at `*dev_ptr` -> apply */
int  mwgpu_comms_reduce_local(mwgpu_ctx *ctx, void **dev_ptr, int *count);
int  mwgpu_comms_apply(mwgpu_ctx *ctx);

/* ---- periodic bookkeeping on the reduced arrays, device side ---- */
typedef struct mwgpu_flat_params {     /* userparams.f90:33-36 */
    int    wl_schedule;                /* 0: every bin within tol of the mean, 1: min visits, 2: above (1-tol) of the mean */
    int    wl_minhist;
    double wl_flattol;
    int    wl_useinvt;
} mwgpu_flat_params;
typedef struct mwgpu_flat_report {     /* what mc_check_flatness logs, for walker 0 of the context */
    int    checked;                    /* 0: returned at the samplerun / empty-histogram guard (:1961) */
    int    hist_reset;                 /* the one-off histogram reset of :1973-1980 happened in this call */
    int    flat;
    int    invt_switched;              /* switched to the 1/t increment (:2134-2142) */
    double mean, max_pct, min_pct;     /* window mean; most / least populated bin in % of it */
    double wl_factor;                  /* after the call */
} mwgpu_flat_report;
/* mc_check_flatness (mc_moves.F90:1936-2185), state effects for every walker: histogram delta
 * all-reduce ('mw'), one-off histogram reset, flatness test by wl_schedule, weight shift, histogram
 * reset + re-base, wl_factor halving, switch to 1/t.  The files of :2067-2101 stay with the caller
 * (mwgpu_mc_get_bins gives the arrays). */
int  mwgpu_mc_check_flatness(mwgpu_ctx *ctx, const mwgpu_flat_params *p, mwgpu_flat_report *report);
/* mc_compute_deltaG_from_hist (mc_moves.F90:2498-2621): all-reduces ('mw') or joins ('dd') the
 * unbiased histogram, returns G(lattice2)-G(lattice1) in kT for the whole box and normP(nbins). */
int  mwgpu_mc_deltag_from_hist(mwgpu_ctx *ctx, double *deltaG_kT, double *normP);
/* comms_join_uhist / comms_join_eta (comms_mpi.f90:299-375, :377-459): stitch the windows of a
 * 'dd' run (walker = window, rank order; NCCL all-gather when the windows span several GPUs) */
int  mwgpu_comms_join_uhist(mwgpu_ctx *ctx, int overlap, double *joined /* nbins */);
int  mwgpu_comms_join_eta(mwgpu_ctx *ctx, int overlap, double *joined /* nbins */);

/* ---- restart and therm rows ----------------------------------- */
/* State effects of mc_checkpoint_load (mc_moves.F90:403-501: cycle number, step sizes, wl_factor, bins,
 * comms_set_histogram, sumhist, firstcycle, cell, reference positions, positions, active lattice) followed
 * by the refresh of mc_init (:842-862: volumes, reciprocal cells, image vectors, chain synchronisation,
 * energies, ls_mu) for one walker of a context that went through the normal start-up (upload of the input
 * configuration, energy_init, mc_init).  As in the reference, ref_hmatrix stays the input cell and the
 * neighbour lists are NOT rebuilt here (they are refreshed at the next list_update_int cycle). */
int  mwgpu_mc_restore(mwgpu_ctx *ctx, int walker, int mc_cycle_num, double mc_max_trans, double mc_dv_max,
                      double wl_factor, int wl_invt_active, int ls,
                      const double *histogram, const double *weight, const double *unbiased_hist,
                      const double *hmatrix, const double *ref_ljr, const double *ljr);
/* the values of one row of <seed>RRR_therm.dat (main.f90:200-223), recorded by the walker kernel */
typedef struct mwgpu_therm_row {
    int64_t icyc;
    int64_t ls;
    double  model_energy[2];   /* Hartree */
    double  ls_mu;
    double  volume[2];         /* Bohr^3 */
    double  hmatrix1[9];       /* hmatrix(:,:,1), Bohr (single-lattice rows print a,b,c,alpha,beta,gamma) */
} mwgpu_therm_row;
/* record a row every file_output_int cycles, at most `capacity` rows per walker between two drains
 * (0, 0 switches recording off) */
int  mwgpu_mc_set_therm(mwgpu_ctx *ctx, int file_output_int, int capacity);
/* drain the rows of one walker; *ndropped = rows lost because the ring was full */
int  mwgpu_mc_get_therm(mwgpu_ctx *ctx, int walker, mwgpu_therm_row *rows, int max_rows, int *nrows, int *ndropped);

/* ---- measurement helpers ------------------------------------------------------------ */
/* elapsed milliseconds of the last mwgpu_mc_run / mwgpu_compute_model_energy_all kernel
 * (CUDA events on the context's stream) */
int  mwgpu_last_kernel_ms(mwgpu_ctx *ctx, float *ms);
/* CUDA-event stopwatch on the context's stream: start, enqueue work (e.g. mwgpu_mc_run_async),
 * stop (synchronises) -> elapsed device milliseconds */
int  mwgpu_timer_start(mwgpu_ctx *ctx);
int  mwgpu_timer_stop(mwgpu_ctx *ctx, float *ms);
/* dependent-free DFMA throughput of the device in TFLOP/s (roofline denominator) */
int  mwgpu_measure_fp64_peak(int device, double *tflops);
int  mwgpu_kernel_launches(mwgpu_ctx *ctx, int64_t *count);   /* kernels launched by this context */
/* Scheduling of one mwgpu_mc_run launch (warp-per-lattice kernel).  The walkers of a batch are independent
 * (mc_cycle, mc_moves.F90:160-260, advances one walker's state; the reference runs one walker per MPI rank), so a
 * batch larger than the GPU holds at once is cut into units of chunk_cycles MC cycles per walker that max_blocks
 * persistent blocks take from a queue; a result never depends on either.  0 = automatic (the default): as many
 * blocks as are resident at once, one unit per walker if the batch fits them, else units of 8 cycles.  A
 * non-zero chunk_cycles cuts every launch into such units. */
int  mwgpu_mc_set_schedule(mwgpu_ctx *ctx, int chunk_cycles, int max_blocks);
/* %globaltimer (ns) at the start and the end of every walker's part of the last mwgpu_mc_run launch:
 * start_end_ns[2*w], start_end_ns[2*w+1] (warp-per-lattice kernel only) */
int  mwgpu_mc_get_walker_times(mwgpu_ctx *ctx, uint64_t *start_end_ns);

#ifdef __cplusplus
}
#endif
#endif
