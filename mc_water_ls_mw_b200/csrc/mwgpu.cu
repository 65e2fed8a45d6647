// mwgpu.cu -- host side of the C ABI (include/mwgpu.h) and the small service kernels.
//
// The two hot kernels are k_mc_run (mw_mc.cuh) and k_model_energy_all (below);
// everything else here is start-up / bookkeeping plumbing around them.
#include "../../include/mwgpu.h"
#include "mw2.cuh"
#include "mw2_energy.cuh"

#include <cmath>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <dlfcn.h>

using namespace mw;

// ------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(const std::string& msg, int code = 1)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(e_), 100 + (int)e_);          \
    } while (0)

extern "C" const char* mwgpu_last_error(void) { return g_last_error.c_str(); }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct NcclApi;
struct mwgpu_ctx {
    int device = 0;
    int N = 0, nlat = 0, W = 0, NB = 0, NBP = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, tv0 = nullptr, tv1 = nullptr;
    DeviceState S{};
    McParams P{};
    mwgpu_mc_params user{};
    bool mc_ready = false, energy_ready = false;
    int walker_kernel = 0;         // 0: automatic, 1: one warp per walker, 2: one warp per lattice, 4: two warps per lattice
    int energy_kernel = 0;         // batched full energy: 0 flattened-entry kernel (mw2_energy.cuh), 1 first generation
    int first_rank = 0, size = 1;
    int num_sms = 148;
    double* stage = nullptr;       // device staging for layout conversion: [W][nlat][N][3] x2 + cells
    size_t stage_doubles = 0;
    double* out = nullptr;         // device scratch for results
    size_t out_doubles = 0;
    int* iout = nullptr;
    size_t iout_ints = 0;
    double* delta = nullptr;       // [3][NBP] summed increments
    double* fifo = nullptr;
    double* book = nullptr;        // scratch of the periodic bookkeeping kernels: report | joined | normP
    int* skip = nullptr;           // [W] guard flags of mc_check_flatness
    double* gather = nullptr;      // [size][NB] windows of all ranks (dd joins over several GPUs)
    size_t gather_doubles = 0;
    size_t queue_ints = 0;         // capacity of S.queue
    int chunk_cycles = 0;          // cycles per unit of work of the walker kernel; 0: automatic
    int max_blocks = 0;            // blocks of the walker kernel; 0: as many as the GPU holds at once
    int64_t launches = 0;
    float last_ms = 0.f;
    std::vector<double> h_mubin, h_binwidth;
    // NCCL
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
};

template <typename T>
static int dalloc(T** p, size_t n)
{
    CUDA_TRY(cudaMalloc((void**)p, sizeof(T) * (n ? n : 1)));
    CUDA_TRY(cudaMemset(*p, 0, sizeof(T) * (n ? n : 1)));
    return 0;
}

extern "C" int mwgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int mwgpu_num_walkers(const mwgpu_ctx* c) { return c ? c->W : 0; }

extern "C" int mwgpu_create(int nwater, int nlat, int nwalkers, int device, mwgpu_ctx** out)
{
    if (!out) return fail("mwgpu_create: out is NULL");
    *out = nullptr;
    if (nwater < 1 || nwater > NMAX) return fail("mwgpu_create: nwater must be in 1..1024");
    if (nlat != 1 && nlat != 2) return fail("Error num_lattices must equal 1 or 2!");
    if (nwalkers < 1) return fail("mwgpu_create: nwalkers must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return fail(std::string("mwgpu_create: no CUDA device available (") + cudaGetErrorString(e) +
                    "); this library has no CPU fallback", 2);
    if (device < 0 || device >= ndev) return fail("mwgpu_create: invalid device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    const size_t smem = walker_smem_bytes(nwater, nlat);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (smem > prop.sharedMemPerBlockOptin)
        return fail("mwgpu_create: a walker of this size does not fit in shared memory");
    const int num_sms = prop.multiProcessorCount;

    mwgpu_ctx* c = new mwgpu_ctx();
    c->device = device; c->N = nwater; c->nlat = nlat; c->W = nwalkers; c->num_sms = num_sms;
    const size_t W = nwalkers, N = nwater, L = nlat;
    DeviceState& S = c->S;
    S.N = nwater; S.nlat = nlat; S.W = nwalkers; S.NB = 0;
    int rc = 0;
    rc |= dalloc(&S.pos, W * L * 3 * N);
    rc |= dalloc(&S.ref, W * L * 3 * N);
    rc |= dalloc(&S.cell, W * L * 9);
    rc |= dalloc(&S.recip, W * L * 9);
    rc |= dalloc(&S.refcell, W * L * 9);
    rc |= dalloc(&S.iv, W * L * 3 * IVC);
    rc |= dalloc(&S.niv, W * 2);
    rc |= dalloc(&S.list, W * L * N * LC);
    rc |= dalloc(&S.nn, W * L * N);
    rc |= dalloc(&S.scal, W);
    rc |= dalloc(&S.transcount, W * N);
    rc |= dalloc(&S.qctr, 8);
    rc |= dalloc(&S.cyc_end, W);
    rc |= dalloc(&S.wtime, W * 2);
    c->stage_doubles = W * L * (2 * 3 * N + 9);
    rc |= dalloc(&c->stage, c->stage_doubles);
    c->out_doubles = W * 2 > N ? W * 2 : N;
    if (c->out_doubles < (size_t)3 * MWGPU_MAXIVECT) c->out_doubles = 3 * MWGPU_MAXIVECT;
    rc |= dalloc(&c->out, c->out_doubles);
    c->iout_ints = N * (2 * MWGPU_MAXNEIGH + 1) + 4;
    rc |= dalloc(&c->iout, c->iout_ints);
    if (rc) { mwgpu_destroy(c); return rc; }
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaEventCreate(&c->tv0));
    CUDA_TRY(cudaEventCreate(&c->tv1));
    *out = c;
    return 0;
}

static void nccl_destroy(mwgpu_ctx* c);

extern "C" void mwgpu_destroy(mwgpu_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    nccl_destroy(c);
    DeviceState& S = c->S;
    void* ptrs[] = {S.pos, S.ref, S.cell, S.recip, S.refcell, S.iv, S.niv, S.list, S.nn, S.scal,
                    S.weight, S.hist, S.uhist, S.wbase, S.hbase, S.ubase, S.transcount, S.mubin,
                    S.binwidth, S.ginv, S.hinc, S.edge, c->stage, c->out, c->iout, c->delta, c->fifo,
                    c->book, c->skip, c->gather, S.therm, S.therm_n, S.queue, S.qctr, S.cyc_end, S.wtime};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tv0) cudaEventDestroy(c->tv0);
    if (c->tv1) cudaEventDestroy(c->tv1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int check_ctx(mwgpu_ctx* c, int walker, bool allow_all)
{
    if (!c) return fail("mwgpu: NULL context");
    if (walker < (allow_all ? -1 : 0) || walker >= c->W) return fail("mwgpu: walker index out of range");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return 0;
}

static int check_ils(mwgpu_ctx* c, int ils)
{
    if (ils < 1 || ils > c->nlat) return fail("mwgpu: lattice index ils out of range (1-based)");
    return 0;
}

static int finish(mwgpu_ctx* c, bool sync)
{
    CUDA_TRY(cudaGetLastError());
    if (sync) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int mwgpu_synchronize(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int mwgpu_kernel_launches(mwgpu_ctx* c, int64_t* n)
{
    if (!c || !n) return fail("mwgpu_kernel_launches: NULL argument");
    *n = c->launches;
    return 0;
}

extern "C" int mwgpu_timer_start(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    CUDA_TRY(cudaEventRecord(c->tv0, c->stream));
    return 0;
}

extern "C" int mwgpu_timer_stop(mwgpu_ctx* c, float* ms)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!ms) return fail("mwgpu_timer_stop: NULL");
    CUDA_TRY(cudaEventRecord(c->tv1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->tv1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->tv0, c->tv1));
    return 0;
}

extern "C" int mwgpu_last_kernel_ms(mwgpu_ctx* c, float* ms)
{
    if (!c || !ms) return fail("mwgpu_last_kernel_ms: NULL argument");
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    *ms = c->last_ms;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layout conversion kernels: reference AoS ljr(3,1,N,nlat[,W])  <->  device SoA [W][nlat][3][N]
// ------------------------------------------------------------------------------------------------
__global__ void k_unpack(DeviceState S, const double* __restrict__ ljr, const double* __restrict__ ref,
                         const double* __restrict__ hm, int w0, int nw, int bcast, int keep_refcell)
{
    const int N = S.N, L = S.nlat;
    const size_t per = (size_t)L * N;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t < (size_t)nw * per) {
        const int w = (int)(t / per);
        const int rem = (int)(t % per);
        const int lat = rem / N, i = rem % N;
        const size_t src = ((bcast ? 0 : (size_t)w * per) + (size_t)lat * N + i) * 3;
        const size_t dst = ((size_t)(w0 + w) * L + lat) * 3 * N + i;
        S.pos[dst] = ljr[src]; S.pos[dst + N] = ljr[src + 1]; S.pos[dst + 2 * N] = ljr[src + 2];
        S.ref[dst] = ref[src]; S.ref[dst + N] = ref[src + 1]; S.ref[dst + 2 * N] = ref[src + 2];
    }
    if (t < (size_t)nw * L * 9) {
        const int w = (int)(t / (L * 9));
        const int rem = (int)(t % (L * 9));
        const double v = hm[(bcast ? 0 : (size_t)w * L * 9) + rem];
        S.cell[(size_t)(w0 + w) * L * 9 + rem] = v;
        if (!keep_refcell) S.refcell[(size_t)(w0 + w) * L * 9 + rem] = v;     // ref_hmatrix = hmatrix at input only
    }
}

__global__ void k_pack(DeviceState S, double* __restrict__ ljr, double* __restrict__ ref,
                       double* __restrict__ hm, int w0, int nw)
{
    const int N = S.N, L = S.nlat;
    const size_t per = (size_t)L * N;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t < (size_t)nw * per) {
        const int w = (int)(t / per);
        const int rem = (int)(t % per);
        const int lat = rem / N, i = rem % N;
        const size_t dst = ((size_t)w * per + (size_t)lat * N + i) * 3;
        const size_t src = ((size_t)(w0 + w) * L + lat) * 3 * N + i;
        ljr[dst] = S.pos[src]; ljr[dst + 1] = S.pos[src + N]; ljr[dst + 2] = S.pos[src + 2 * N];
        ref[dst] = S.ref[src]; ref[dst + 1] = S.ref[src + N]; ref[dst + 2] = S.ref[src + 2 * N];
    }
    if (t < (size_t)nw * L * 9) {
        const int w = (int)(t / (L * 9));
        const int rem = (int)(t % (L * 9));
        hm[(size_t)w * L * 9 + rem] = S.cell[(size_t)(w0 + w) * L * 9 + rem];
    }
}

static int upload_impl(mwgpu_ctx* c, int w0, int nw, int bcast, const double* ljr, const double* ref, const double* hm,
                       bool restart = false)
{
    if (!ljr || !hm) return fail("mwgpu_upload: ljr and hmatrix must not be NULL");
    if (!ref) ref = ljr;                               // init.f90:103: ref_ljr = ljr
    const size_t nsrc = bcast ? 1 : (size_t)nw;
    const size_t np = nsrc * c->nlat * c->N * 3, nh = nsrc * c->nlat * 9;
    double* d_ljr = c->stage;
    double* d_ref = d_ljr + (size_t)c->W * c->nlat * c->N * 3;
    double* d_hm = d_ref + (size_t)c->W * c->nlat * c->N * 3;
    CUDA_TRY(cudaMemcpyAsync(d_ljr, ljr, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ref, ref, np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(d_hm, hm, nh * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const size_t threads = (size_t)nw * c->nlat * c->N;
    const int blk = 256;
    k_unpack<<<(unsigned)((threads + blk - 1) / blk), blk, 0, c->stream>>>(c->S, d_ljr, d_ref, d_hm, w0, nw, bcast, restart ? 1 : 0);
    c->launches++;
    if (!restart) c->energy_ready = false;
    return finish(c, true);
}

extern "C" int mwgpu_upload(mwgpu_ctx* c, int walker, const double* ljr, const double* ref, const double* hm)
{
    if (int rc = check_ctx(c, walker, true)) return rc;
    if (walker < 0) return upload_impl(c, 0, c->W, 1, ljr, ref, hm);
    return upload_impl(c, walker, 1, 0, ljr, ref, hm);
}

extern "C" int mwgpu_upload_all(mwgpu_ctx* c, const double* ljr, const double* ref, const double* hm)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    return upload_impl(c, 0, c->W, 0, ljr, ref, hm);
}

static int download_impl(mwgpu_ctx* c, int w0, int nw, double* ljr, double* ref, double* hm)
{
    const size_t np = (size_t)nw * c->nlat * c->N * 3, nh = (size_t)nw * c->nlat * 9;
    double* d_ljr = c->stage;
    double* d_ref = d_ljr + (size_t)c->W * c->nlat * c->N * 3;
    double* d_hm = d_ref + (size_t)c->W * c->nlat * c->N * 3;
    const size_t threads = (size_t)nw * c->nlat * c->N;
    const int blk = 256;
    k_pack<<<(unsigned)((threads + blk - 1) / blk), blk, 0, c->stream>>>(c->S, d_ljr, d_ref, d_hm, w0, nw);
    c->launches++;
    if (ljr) CUDA_TRY(cudaMemcpyAsync(ljr, d_ljr, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (ref) CUDA_TRY(cudaMemcpyAsync(ref, d_ref, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (hm) CUDA_TRY(cudaMemcpyAsync(hm, d_hm, nh * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return finish(c, true);
}

extern "C" int mwgpu_download(mwgpu_ctx* c, int walker, double* ljr, double* ref, double* hm)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    return download_impl(c, walker, 1, ljr, ref, hm);
}

extern "C" int mwgpu_download_all(mwgpu_ctx* c, double* ljr, double* ref, double* hm)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    return download_impl(c, 0, c->W, ljr, ref, hm);
}

// ------------------------------------------------------------------------------------------------
// service kernel: one warp per walker, op selected at run time
// ------------------------------------------------------------------------------------------------
enum WalkerOp : int {
    OP_ENERGY_INIT = 0,   // molint.F90:91-153
    OP_IVECTS,            // compute_ivects(lat)
    OP_NEIGHBOURS,        // compute_neighbours(lat); lat < 0: all lattices
    OP_MODEL_ENERGY,      // compute_model_energy(lat); lat < 0: all lattices
    OP_LOCAL_ONE,         // compute_local_real_energy(imol, lat)
    OP_LOCAL_ALL,         // for every molecule
    OP_MONITOR,           // state effects of mc_monitor_stats
    OP_CHAIN_SYNC,        // mc_check_chain_synchronisation
    OP_RESTART,           // refresh after mc_checkpoint_load (mc_moves.F90:842-862)
};

struct OpArgs {
    int op, w0, lat, imol;
    double* out;          // per-op result buffer
    int eq_adjust, eq_mc_cycles;
    double target_ratio;
    double beta, pressure; int leshift;
    int refresh_mu;
};

template <int NLAT>
__global__ void __launch_bounds__(32) k_walker_op(const __grid_constant__ DeviceState S, const __grid_constant__ OpArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int wi = a.w0 + blockIdx.x;
    if (wi >= S.W) return;
    const int lane = lane_id(), N = S.N;
    const WalkerView w = carve_walker(smem, N, NLAT);
    load_walker(S, wi, w);
    WalkerScalars* sc = w.sc;
    int err = 0;
    bool store_lists = false;

    switch (a.op) {
    case OP_ENERGY_INIT: {
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            sc->vol[lat] = cell_volume(w, lat);                         // molint.F90:125
            refresh_recip(smem, N, NLAT, lat);                          // init.f90:90
        }
        sc->error = 0;
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            err |= compute_neighbours_warp(smem, N, NLAT, lat);         // includes compute_ivects
            sc->E[lat] = full_energy_warp(smem, N, NLAT, lat);
        }
        if (NLAT == 2 && a.refresh_mu) {      // mc_moves.F90:857-862 (left-to-right association)
            double mu = sc->E[0] + a.pressure * sc->vol[0] - sc->E[1] - a.pressure * sc->vol[1];
            if (a.leshift) mu = mu - sc->refH[0] + sc->refH[1];
            sc->mu = mu * a.beta - (double)N * log(sc->vol[0] / sc->vol[1]);
        }
        store_lists = true;
        break;
    }
    case OP_IVECTS:
        err |= compute_ivects_warp(smem, N, NLAT, a.lat);
        break;
    case OP_NEIGHBOURS:
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat)
            if (a.lat < 0 || a.lat == lat) err |= compute_neighbours_warp(smem, N, NLAT, lat);
        store_lists = true;
        break;
    case OP_MODEL_ENERGY:
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat)
            if (a.lat < 0 || a.lat == lat) {
                const double e = full_energy_warp(smem, N, NLAT, lat);
                sc->E[lat] = e;
                if (a.out && lane == 0) a.out[(size_t)(wi - a.w0) * NLAT + lat] = e;
            }
        break;
    case OP_LOCAL_ONE:
    case OP_LOCAL_ALL: {
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) compute_bond_masks_warp(smem, N, NLAT, lat);
        const int i0 = (a.op == OP_LOCAL_ONE) ? a.imol : 0;
        const int i1 = (a.op == OP_LOCAL_ONE) ? a.imol + 1 : N;
        for (int i = i0; i < i1; ++i) {
            double eo[2] = {0, 0}, en[2] = {0, 0};
            uint32_t mo[2], mn[2];
            local_energies_warp<NLAT, false>(w, i, eo, en, mo, mn);
            if (lane == 0) a.out[i - i0] = (a.lat == 0) ? eo[0] : eo[1];
        }
        break;
    }
    case OP_MONITOR: {
        // mc_moves.F90:1722-1732 (exact arithmetic: the step sizes feed the state arithmetic)
        const double atr = xd((double)sc->acc_r, (double)sc->att_r);
        const double avr = xd((double)sc->acc_v, (double)sc->att_v);
        if (a.eq_adjust && sc->cycle < a.eq_mc_cycles) {
            sc->max_trans = fmax(xd(xm(sc->max_trans, atr), a.target_ratio), 0.1);
            sc->dv_max = fmax(xd(xm(sc->dv_max, avr), a.target_ratio), 0.0001);
        }
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) sc->E[lat] = full_energy_warp(smem, N, NLAT, lat);   // :1786-1792
        sc->acc_r = 0; sc->acc_v = 0; sc->acc_s = 0; sc->att_r = 0; sc->att_v = 0; sc->att_s = 0; // :1797-1810
        for (int i = lane; i < N; i += 32) S.transcount[(size_t)wi * N + i] = 0;
        sc->avgE[0] = 0.0; sc->avgE[1] = 0.0;
        sc->max_dmu = 0.0; sc->min_dmu = F_HUGE;
        break;
    }
    case OP_RESTART:
        // mc_moves.F90:842-856: volume, recip, image vectors of the loaded cells -- NOT the neighbour lists
        // (the reference keeps the lists it built for the input configuration until the next refresh) --
        // then the chain synchronisation and the energies
#pragma unroll
        for (int lat = 0; lat < NLAT; ++lat) {
            sc->vol[lat] = cell_volume(w, lat);
            refresh_recip(smem, N, NLAT, lat);
            err |= compute_ivects_warp(smem, N, NLAT, lat);
        }
        if (NLAT == 1) sc->E[0] = full_energy_warp(smem, N, NLAT, 0);
        // fall through
    case OP_CHAIN_SYNC: {
        // mc_moves.F90:2217-2416 (two lattices only)
        if (NLAT == 2) {
            sc->E[0] = full_energy_warp(smem, N, NLAT, 0);
            sc->E[1] = full_energy_warp(smem, N, NLAT, 1);
            const double* rh = S.refcell + (size_t)wi * NLAT * 9;
            if (lane < 9) w.cell[9 + lane] = xa(rh[9 + lane], xs(w.cell[lane], rh[lane]));     // :2262,2277
            __syncwarp();
            refresh_recip(smem, N, NLAT, 0);
            refresh_recip(smem, N, NLAT, 1);
            const double* R = S.ref + (size_t)wi * NLAT * 3 * N;
            for (int i = lane; i < N; i += 32) {
                double sv[2][3], rsv[2][3];
#pragma unroll
                for (int lat = 0; lat < 2; ++lat) {
                    const double* rm = w.recip + 9 * lat;
                    const double* P = w.pos + lat * 3 * N;
                    const double* Q = R + lat * 3 * N;
                    const double a0 = P[i], a1 = P[N + i], a2 = P[2 * N + i];
                    const double b0 = Q[i], b1 = Q[N + i], b2 = Q[2 * N + i];
                    sv[lat][0] = xa(xa(xm(MW_H(rm,1,1), a0), xm(MW_H(rm,2,1), a1)), xm(MW_H(rm,3,1), a2));
                    sv[lat][1] = xa(xa(xm(MW_H(rm,1,2), a0), xm(MW_H(rm,2,2), a1)), xm(MW_H(rm,3,2), a2));
                    sv[lat][2] = xa(xa(xm(MW_H(rm,1,3), a0), xm(MW_H(rm,2,3), a1)), xm(MW_H(rm,3,3), a2));
                    rsv[lat][0] = xa(xa(xm(MW_H(rm,1,1), b0), xm(MW_H(rm,2,1), b1)), xm(MW_H(rm,3,1), b2));
                    rsv[lat][1] = xa(xa(xm(MW_H(rm,1,2), b0), xm(MW_H(rm,2,2), b1)), xm(MW_H(rm,3,2), b2));
                    rsv[lat][2] = xa(xa(xm(MW_H(rm,1,3), b0), xm(MW_H(rm,2,3), b1)), xm(MW_H(rm,3,3), b2));
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        sv[lat][d] = xm(xm(sv[lat][d], 0.5), INV_PI);
                        rsv[lat][d] = xm(xm(rsv[lat][d], 0.5), INV_PI);
                    }
                }
                double s2[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) s2[d] = xa(rsv[1][d], xs(sv[0][d], rsv[0][d]));
                const double* hm = w.cell + 9;
                double* P2 = w.pos + 3 * N;
                P2[i]         = xa(xa(xm(MW_H(hm,1,1), s2[0]), xm(MW_H(hm,1,2), s2[1])), xm(MW_H(hm,1,3), s2[2]));
                P2[N + i]     = xa(xa(xm(MW_H(hm,2,1), s2[0]), xm(MW_H(hm,2,2), s2[1])), xm(MW_H(hm,2,3), s2[2]));
                P2[2 * N + i] = xa(xa(xm(MW_H(hm,3,1), s2[0]), xm(MW_H(hm,3,2), s2[1])), xm(MW_H(hm,3,3), s2[2]));
            }
            __syncwarp();
#pragma unroll
            for (int lat = 0; lat < 2; ++lat) {
                sc->vol[lat] = cell_volume(w, lat);
                err |= compute_ivects_warp(smem, N, NLAT, lat);
            }
            sc->E[0] = full_energy_warp(smem, N, NLAT, 0);
            sc->E[1] = full_energy_warp(smem, N, NLAT, 1);
            // left-to-right association (:2400-2402)
            double mu = sc->E[0] + a.pressure * sc->vol[0] - sc->E[1] - a.pressure * sc->vol[1];
            if (a.leshift) mu = mu - sc->refH[0] + sc->refH[1];
            sc->mu = mu * a.beta - (double)N * log(sc->vol[0] / sc->vol[1]);
        }
        break;
    }
    default: break;
    }
    sc->error |= err;
    __syncwarp();
    store_walker(S, wi, w, store_lists);
}

static int launch_op(mwgpu_ctx* c, OpArgs a, int nw, bool sync = true)
{
    const size_t smem = walker_smem_bytes(c->N, c->nlat);
    if (c->nlat == 2) {
        CUDA_TRY(cudaFuncSetAttribute(k_walker_op<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_walker_op<2><<<nw, 32, smem, c->stream>>>(c->S, a);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(k_walker_op<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_walker_op<1><<<nw, 32, smem, c->stream>>>(c->S, a);
    }
    c->launches++;
    return finish(c, sync);
}

static OpArgs make_args(mwgpu_ctx* c, int op, int w0, int lat, int imol, double* out)
{
    OpArgs a{};
    a.op = op; a.w0 = w0; a.lat = lat; a.imol = imol; a.out = out;
    a.eq_adjust = c->user.eq_adjust_mc; a.eq_mc_cycles = c->user.eq_mc_cycles;
    a.target_ratio = c->user.mc_target_ratio;
    a.beta = c->P.beta; a.pressure = c->P.pressure; a.leshift = c->P.leshift;
    a.refresh_mu = c->mc_ready ? 1 : 0;
    return a;
}

// OR of the walkers' error bits and the first flagged walker, reduced on the device: 8 bytes come back
__global__ void k_collect_errors(DeviceState S, int* __restrict__ out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = (w < S.W) ? S.scal[w].error : 0;
    const unsigned any = __ballot_sync(0xffffffffu, e != 0);
    if (any == 0) return;
    const int all = (int)__reduce_or_sync(0xffffffffu, (unsigned)e);
    if ((threadIdx.x & 31) == __ffs(any) - 1) { atomicOr(out, all); atomicMin(out + 1, w); }
}

static int collect_errors(mwgpu_ctx* c, const char* what)
{
    int h[2] = {0, 0x7fffffff};
    CUDA_TRY(cudaMemcpyAsync(c->iout, h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    k_collect_errors<<<(c->W + 255) / 256, 256, 0, c->stream>>>(c->S, c->iout);
    c->launches++;
    CUDA_TRY(cudaMemcpyAsync(h, c->iout, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const int all = h[0], first = h[1];
    if (!all) return 0;
    std::string msg = std::string(what) + ": device error bits " + std::to_string(all) + " (first walker " +
                      std::to_string(first) + "):";
    if (all & ERR_LIST_OVERFLOW) msg += " a molecule has more than 32 list neighbours;";
    if (all & ERR_IVECT_OVERFLOW) msg += " more than 32 image vectors (cell narrower than the cut-off);";
    if (all & ERR_BOND_OVERFLOW) msg += " too many in-range bonds in one batch;";
    if (all & ERR_ITEM_OVERFLOW) msg += " too many triplet centres/items in one trial move;";
    if (all & ERR_SELF_IMAGE) msg += " a molecule neighbours its own periodic image (cell too small);";
    if (all & ERR_RNG_UNDERRUN) msg += " random-number FIFO ran dry;";
    if (all & ERR_WINDOW) msg += " Error : Not all walkers have reached their designated window;";
    if (all & ERR_PROB) msg += " Cumulative move type probability error;";
    return fail(msg, 3);
}

extern "C" int mwgpu_energy_init(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (int rc = launch_op(c, make_args(c, OP_ENERGY_INIT, 0, -1, 0, nullptr), c->W)) return rc;
    c->energy_ready = true;
    return collect_errors(c, "mwgpu_energy_init");
}

extern "C" int mwgpu_compute_ivects(mwgpu_ctx* c, int walker, int ils, int* nivect, double* ivect)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (int rc = check_ils(c, ils)) return rc;
    if (int rc = launch_op(c, make_args(c, OP_IVECTS, walker, ils - 1, 0, nullptr), 1)) return rc;
    int niv[2];
    CUDA_TRY(cudaMemcpy(niv, c->S.niv + walker * 2, sizeof(niv), cudaMemcpyDeviceToHost));
    if (nivect) *nivect = niv[ils - 1];
    if (niv[ils - 1] > IVC) return fail("mwgpu_compute_ivects: more than 32 image vectors", 3);
    if (ivect) {
        std::vector<double> h(3 * IVC);
        CUDA_TRY(cudaMemcpy(h.data(), c->S.iv + ((size_t)walker * c->nlat + (ils - 1)) * 3 * IVC,
                            sizeof(double) * 3 * IVC, cudaMemcpyDeviceToHost));
        for (int k = 0; k < IVC; ++k)
            for (int d = 0; d < 3; ++d) ivect[k * 3 + d] = (k < niv[ils - 1]) ? h[d * IVC + k] : 0.0;
    }
    return 0;
}

static int fetch_lists(mwgpu_ctx* c, int walker, int ils, int* nn, int* jn, int* vn)
{
    const int N = c->N;
    std::vector<uint16_t> hl((size_t)N * LC);
    std::vector<uint8_t> hn(N);
    CUDA_TRY(cudaMemcpy(hl.data(), c->S.list + ((size_t)walker * c->nlat + (ils - 1)) * N * LC,
                        sizeof(uint16_t) * N * LC, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(hn.data(), c->S.nn + ((size_t)walker * c->nlat + (ils - 1)) * N, N, cudaMemcpyDeviceToHost));
    const EntFmt F = ent_fmt(N);
    for (int i = 0; i < N; ++i) {
        if (nn) nn[i] = hn[i];
        for (int s = 0; s < MWGPU_MAXNEIGH; ++s) {
            const bool used = s < hn[i] && s < LC;
            if (jn) jn[i * MWGPU_MAXNEIGH + s] = used ? (hl[(size_t)i * LC + s] & F.jmask) + 1 : 0;
            if (vn) vn[i * MWGPU_MAXNEIGH + s] = used ? (hl[(size_t)i * LC + s] >> F.ishift) + 1 : 0;
        }
    }
    return 0;
}

extern "C" int mwgpu_compute_neighbours(mwgpu_ctx* c, int walker, int ils, int* nn, int* jn, int* vn)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (int rc = check_ils(c, ils)) return rc;
    if (int rc = launch_op(c, make_args(c, OP_NEIGHBOURS, walker, ils - 1, 0, nullptr), 1)) return rc;
    if (int rc = collect_errors(c, "mwgpu_compute_neighbours")) return rc;
    return fetch_lists(c, walker, ils, nn, jn, vn);
}

extern "C" int mwgpu_get_neighbours(mwgpu_ctx* c, int walker, int ils, int* nn, int* jn, int* vn)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (int rc = check_ils(c, ils)) return rc;
    return fetch_lists(c, walker, ils, nn, jn, vn);
}

extern "C" int mwgpu_compute_neighbours_all(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (int rc = launch_op(c, make_args(c, OP_NEIGHBOURS, 0, -1, 0, nullptr), c->W)) return rc;
    return collect_errors(c, "mwgpu_compute_neighbours_all");
}

extern "C" int mwgpu_compute_model_energy(mwgpu_ctx* c, int walker, int ils, double* energy)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (int rc = check_ils(c, ils)) return rc;
    if (int rc = launch_op(c, make_args(c, OP_MODEL_ENERGY, walker, ils - 1, 0, c->out), 1)) return rc;
    if (energy) CUDA_TRY(cudaMemcpy(energy, c->out + (ils - 1), sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mwgpu_compute_local_real_energy(mwgpu_ctx* c, int walker, int imol, int ils, double* energy)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (int rc = check_ils(c, ils)) return rc;
    if (imol < 1 || imol > c->N) return fail("mwgpu_compute_local_real_energy: imol out of range (1-based)");
    if (!energy) return fail("mwgpu_compute_local_real_energy: energy is NULL");
    if (int rc = launch_op(c, make_args(c, OP_LOCAL_ONE, walker, ils - 1, imol - 1, c->out), 1)) return rc;
    CUDA_TRY(cudaMemcpy(energy, c->out, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mwgpu_compute_local_real_energy_all(mwgpu_ctx* c, int walker, int ils, double* energy)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (int rc = check_ils(c, ils)) return rc;
    if (!energy) return fail("mwgpu_compute_local_real_energy_all: energy is NULL");
    if (int rc = launch_op(c, make_args(c, OP_LOCAL_ALL, walker, ils - 1, 0, c->out), 1)) return rc;
    CUDA_TRY(cudaMemcpy(energy, c->out, sizeof(double) * c->N, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched full energy: "full mW energy evals/s" kernel.  One warp per (walker, lattice);
// needs only positions, lists and image vectors of that lattice.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_model_energy_all(const __grid_constant__ DeviceState S, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int unit = blockIdx.x;                 // walker * nlat + lat
    if (unit >= S.W * S.nlat) return;
    const int wi = unit / S.nlat, lat = unit % S.nlat;
    const int lane = lane_id(), N = S.N;
    // a one-lattice view: lattice `lat` of the walker is staged as lattice 0
    const WalkerView w = carve_walker(smem, N, 1);
    const double* gp = S.pos + ((size_t)wi * S.nlat + lat) * 3 * N;
    for (int t = lane; t < 3 * N; t += 32) w.pos[t] = gp[t];
    const double* gi = S.iv + ((size_t)wi * S.nlat + lat) * 3 * IVC;
    for (int t = lane; t < 3 * IVC; t += 32) w.iv[t] = gi[t];
    const uint4* gl = (const uint4*)(S.list + ((size_t)wi * S.nlat + lat) * N * LC);
    uint4* sl = (uint4*)w.list;
    for (int t = lane; t < N * LC / 8; t += 32) sl[t] = gl[t];
    const uint8_t* gn = S.nn + ((size_t)wi * S.nlat + lat) * N;
    for (int t = lane; t < N; t += 32) w.nn[t] = gn[t];
    __syncwarp();
    const double e = full_energy_warp(smem, N, 1, 0);
    if (lane == 0) {
        S.scal[wi].E[lat] = e;
        if (out) out[unit] = e;
    }
}

extern "C" int mwgpu_compute_model_energy_all(mwgpu_ctx* c, double* energies)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    const size_t smem = (size_t)v2::ELay(c->N).bytes();
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    if (c->energy_kernel == 1) {
        const size_t smem1 = walker_smem_bytes(c->N, 1);
        CUDA_TRY(cudaFuncSetAttribute(k_model_energy_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        k_model_energy_all<<<c->W * c->nlat, 32, smem1, c->stream>>>(c->S, c->out);
    } else if (c->N <= v2::E3_THREADS && c->energy_kernel == 0) {
        // one lane per molecule, tensor-form three-body sum (mw2_energy.cuh)
        const v2::E3Lay Y3(c->N);
        const int nunits = c->W * c->nlat, grid = (nunits + Y3.upc() - 1) / Y3.upc();
        if (c->N == 48 && !getenv("MWGPU_ENERGY_NO_TMA")) {
            // persistent, TMA double-buffered form: one CTA per resident slot, looping over pairs of units
            const v2::E4Lay Y4(c->N);
            int per_sm = 0;
            CUDA_TRY(cudaFuncSetAttribute(v2::k_model_energy4<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, Y4.bytes()));
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v2::k_model_energy4<48>, v2::E3_THREADS, Y4.bytes()));
            const int g4 = std::min(grid, std::max(1, per_sm) * c->num_sms);
            v2::k_model_energy4<48><<<g4, v2::E3_THREADS, Y4.bytes(), c->stream>>>(c->S, c->out);
        } else if (c->N == 48) {
            CUDA_TRY(cudaFuncSetAttribute(v2::k_model_energy3<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, Y3.bytes()));
            v2::k_model_energy3<48><<<grid, v2::E3_THREADS, Y3.bytes(), c->stream>>>(c->S, c->out);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(v2::k_model_energy3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Y3.bytes()));
            v2::k_model_energy3<0><<<grid, v2::E3_THREADS, Y3.bytes(), c->stream>>>(c->S, c->out);
        }
    } else if (c->N == 48) {
        CUDA_TRY(cudaFuncSetAttribute(v2::k_model_energy2<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        v2::k_model_energy2<48><<<c->W * c->nlat, 32, smem, c->stream>>>(c->S, c->out);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(v2::k_model_energy2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        v2::k_model_energy2<0><<<c->W * c->nlat, 32, smem, c->stream>>>(c->S, c->out);
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->launches++;
    if (int rc = finish(c, false)) return rc;
    if (energies)
        CUDA_TRY(cudaMemcpyAsync(energies, c->out, sizeof(double) * c->W * c->nlat, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// mc_init (host): bin grid, windows, weights, log_unbiased_norm, initial order parameter
// ------------------------------------------------------------------------------------------------
static double powi(double x, int n)     // x**n, integer n (mc_moves.F90:588,626,642)
{
    unsigned m = (n < 0) ? (unsigned)(-n) : (unsigned)n;
    double y = (m & 1) ? x : 1.0;
    while (m >>= 1) { x = x * x; if (m & 1) y = y * x; }
    return (n < 0) ? 1.0 / y : y;
}

static double gp_ratio(double a, double ssum, int Ns)     // mc_moves.F90:584-594, :604-613
{
    double r = 1.1, r_new;
    int k = 0;
    for (;;) {
        ++k;
        const double tmpsum = a * (1.0 - powi(r, Ns)) / (1.0 - r);
        r_new = r * std::pow(ssum / tmpsum, 1.0 / (double)Ns);
        if (std::fabs(r_new - r) <= 2.0 * DBL_EPSILON) break;
        if (k > 1000000) break;
        r = r_new;
    }
    return r;
}

extern "C" int mwgpu_mc_init(mwgpu_ctx* c, const mwgpu_mc_params* up, int first_rank, int size,
                             const double* file_weights, int n_file_weights, double file_wl_factor)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!up) return fail("mwgpu_mc_init: params is NULL");
    if (!c->energy_ready) return fail("mwgpu_mc_init: call mwgpu_energy_init first");
    if (size < 1 || first_rank < 0 || first_rank + c->W > size)
        return fail("mwgpu_mc_init: walkers [first_rank, first_rank+nwalkers) must lie inside [0,size)");
    c->user = *up;
    mwgpu_mc_params& u = c->user;
    c->first_rank = first_rank; c->size = size;
    const int N = c->N, W = c->W;
    if (u.nbins % 2 == 0) u.nbins += 1;                                     // :557
    const int nb = u.nbins;
    c->NB = nb; c->NBP = (nb + 31) / 32 * 32;

    // ---- grid (:570-656)
    const double s_pos = std::fabs(u.mu_max) - 0.5, s_neg = std::fabs(u.mu_min) - 0.5;
    const double a_pos = 1.0, a_neg = 1.0;
    const int Ns = nb / 2;
    const double r_pos = gp_ratio(a_pos, s_pos, Ns), r_neg = gp_ratio(a_neg, s_neg, Ns);
    std::vector<double> mu_bin(nb), bw(nb), edge(nb + 1);
    {
        double mu_u = -0.5, mu_l;
        int k = 0;
        for (int ibin = nb / 2; ibin >= 1; --ibin) {
            mu_l = mu_u - a_neg * powi(r_neg, k);
            mu_bin[ibin - 1] = 0.5 * (mu_u + mu_l);
            bw[ibin - 1] = mu_u - mu_l;
            edge[ibin - 1] = mu_l; edge[ibin] = mu_u;
            mu_u = mu_l; ++k;
        }
        mu_bin[nb / 2] = 0.0; bw[nb / 2] = 1.0;
        mu_l = 0.5; k = 0;
        for (int ibin = nb / 2 + 2; ibin <= nb; ++ibin) {
            mu_u = mu_l + a_pos * powi(r_pos, k);
            mu_bin[ibin - 1] = 0.5 * (mu_u + mu_l);
            bw[ibin - 1] = mu_u - mu_l;
            edge[ibin - 1] = mu_l; edge[ibin] = mu_u;
            mu_l = mu_u; ++k;
        }
    }
    double av_bw = 0.0;
    for (int i = 0; i < nb; ++i) av_bw = av_bw + bw[i];
    av_bw = av_bw / (double)nb;
    c->h_mubin = mu_bin; c->h_binwidth = bw;

    // ---- weights (:734-776) and log_unbiased_norm (:781-806)
    std::vector<double> weight(nb, 0.0);
    double wl_factor = u.wl_factor;
    const double orig_wl_factor = u.wl_factor;
    double lun = 0.0;
    if (c->nlat == 2) {
        if (file_weights) {
            if (file_wl_factor > (double)1e-10f) {
                wl_factor = std::fmin(wl_factor, file_wl_factor);
                if (u.samplerun) wl_factor = 0.0;
            }
            for (int i = 0; i < n_file_weights && i < nb; ++i) weight[i] = file_weights[i];
        }
        double hits = (double)u.max_mc_cycles - (double)u.eq_mc_cycles;
        hits = hits * (double)(size * N) / (double)nb;
        double incr = hits * av_bw;
        lun = std::log(incr) + weight[0];
        for (int k = 1; k < nb; ++k) {
            incr = hits * av_bw;
            if (lun > weight[k] + std::log(incr)) lun = lun + std::log(1.0 + incr * std::exp(weight[k] - lun));
            else lun = std::log(incr) + weight[k] + std::log(1.0 + std::exp(lun - weight[k]) / incr);
        }
    }

    // ---- (re)allocate per-walker bin arrays
    DeviceState& S = c->S;
    void* old[] = {S.weight, S.hist, S.uhist, S.wbase, S.hbase, S.ubase, S.mubin, S.binwidth, S.ginv, S.hinc, S.edge, c->delta};
    for (void* p : old) if (p) cudaFree(p);
    S.NB = nb;
    int rc = 0;
    rc |= dalloc(&S.weight, (size_t)W * nb); rc |= dalloc(&S.hist, (size_t)W * nb); rc |= dalloc(&S.uhist, (size_t)W * nb);
    rc |= dalloc(&S.wbase, (size_t)W * nb); rc |= dalloc(&S.hbase, (size_t)W * nb); rc |= dalloc(&S.ubase, (size_t)W * nb);
    rc |= dalloc(&S.mubin, nb); rc |= dalloc(&S.binwidth, nb); rc |= dalloc(&S.ginv, nb); rc |= dalloc(&S.hinc, nb);
    rc |= dalloc(&S.edge, nb + 1);
    rc |= dalloc(&c->delta, (size_t)3 * c->NBP);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(S.mubin, mu_bin.data(), sizeof(double) * nb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.binwidth, bw.data(), sizeof(double) * nb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.edge, edge.data(), sizeof(double) * (nb + 1), cudaMemcpyHostToDevice));
    {
        std::vector<double> tab(nb, 0.0);
        for (int k = 0; k + 1 < nb; ++k) tab[k] = 2.0 / (bw[k] + bw[k + 1]);
        CUDA_TRY(cudaMemcpy(S.ginv, tab.data(), sizeof(double) * nb, cudaMemcpyHostToDevice));
        for (int k = 0; k < nb; ++k) tab[k] = av_bw / bw[k];
        CUDA_TRY(cudaMemcpy(S.hinc, tab.data(), sizeof(double) * nb, cudaMemcpyHostToDevice));
    }

    // ---- per-walker scalars: windows (:659-722), ref_enthalpy (main.f90:146-150), mu (:857-862)
    std::vector<WalkerScalars> h(W);
    CUDA_TRY(cudaMemcpy(h.data(), S.scal, sizeof(WalkerScalars) * W, cudaMemcpyDeviceToHost));
    std::vector<double> hw((size_t)W * nb), hb((size_t)W * nb);
    const double beta = 1.0 / (KB * u.temperature);
    for (int w = 0; w < W; ++w) {
        WalkerScalars& sc = h[w];
        const int rank = first_rank + w;
        sc.ls = u.ls;
        if (u.dd) {
            const int bpw = nb / size;
            const int ov = (size == 1) ? 0 : u.window_overlap;             // io.f90:249
            auto sumw = [&](int n) { double s = 0.0; for (int i = 0; i < n; ++i) s += bw[i]; return s; };
            if (rank == 0) {
                sc.start_bin = 1; sc.end_bin = bpw + ov;
                sc.mu_lo = u.mu_min; sc.mu_hi = u.mu_min + sumw(sc.end_bin);
            }
            if (size > 1) {
                if (rank >= 1 && rank <= size - 2) {
                    sc.start_bin = rank * bpw - ov; sc.end_bin = (rank + 1) * bpw + ov;
                    sc.mu_lo = u.mu_min + sumw(sc.start_bin - 1); sc.mu_hi = u.mu_min + sumw(sc.end_bin);
                }
                if (rank == size - 1) {
                    sc.start_bin = rank * bpw - ov; sc.end_bin = nb;
                    sc.mu_lo = u.mu_min + sumw(sc.start_bin - 1); sc.mu_hi = u.mu_max;
                }
            }
            if (sc.mu_hi < 0.0) sc.ls = 1;
            if (sc.mu_lo > 0.0) sc.ls = 2;
        } else {
            sc.start_bin = 1; sc.end_bin = nb; sc.mu_lo = u.mu_min; sc.mu_hi = u.mu_max;
        }
        for (int l = 0; l < c->nlat; ++l) {
            sc.refH[l] = sc.E[l];
            if (u.npt) sc.refH[l] = sc.refH[l] + u.pressure * sc.vol[l];
        }
        if (std::fabs(u.input_ref_enthalpy[0]) > DBL_MIN || std::fabs(u.input_ref_enthalpy[1]) > DBL_MIN) {
            sc.refH[0] = u.input_ref_enthalpy[0]; sc.refH[1] = u.input_ref_enthalpy[1];
        }
        if (c->nlat == 2) {
            double mu = sc.E[0] + u.pressure * sc.vol[0] - sc.E[1] - u.pressure * sc.vol[1];
            if (u.leshift) mu = mu - sc.refH[0] + sc.refH[1];
            sc.mu = mu * beta - (double)N * std::log(sc.vol[0] / sc.vol[1]);
        } else {
            sc.mu = 0.0;
        }
        sc.max_trans = u.mc_max_trans; sc.dv_max = u.mc_dv_max; sc.wl_factor = wl_factor;
        sc.avgE[0] = sc.avgE[1] = 0.0; sc.min_dmu = DBL_MAX; sc.max_dmu = 0.0; sc.sumhist = 0.0;
        sc.rng_index = 0; sc.cycle = 0;
        sc.acc_r = sc.acc_v = sc.acc_s = sc.att_r = sc.att_v = sc.att_s = 0;
        sc.in_window = u.dd ? 0 : 1;
        sc.wl_invt_active = 0; sc.wmin_zero = 0; sc.error = 0;
        sc.firstcycle = !(c->nlat == 2 && wl_factor < orig_wl_factor);       // :817-821
        sc.hist_reset = 0;
        for (int i = 0; i < nb; ++i) {
            hb[(size_t)w * nb + i] = weight[i];                            // eta_last_sync = weight (:776)
            double v = weight[i];
            if (u.dd && (i + 1 < sc.start_bin || i + 1 > sc.end_bin)) v = 0.0;   // :808-814
            hw[(size_t)w * nb + i] = v;
        }
    }
    CUDA_TRY(cudaMemcpy(S.scal, h.data(), sizeof(WalkerScalars) * W, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.weight, hw.data(), sizeof(double) * W * nb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(S.wbase, hb.data(), sizeof(double) * W * nb, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(S.transcount, 0, sizeof(int) * W * N));

    // ---- shared run parameters; move-type probabilities (mc_moves.F90:153-176)
    McParams& P = c->P;
    P = McParams{};
    P.beta = beta; P.pressure = u.pressure;
    double sw = u.mc_switch_prob, vp = u.mc_vol_prob, tp = u.mc_trans_prob;
    if (u.mc_always_switch) sw = 0.0;
    if (!u.allow_switch) sw = 0.0;
    if (!u.npt) vp = 0.0;
    if (!u.allow_vol) vp = 0.0;
    if (!u.allow_trans) tp = 0.0;
    const double sum_prob = tp + vp + sw;
    P.transP = tp / sum_prob; P.volP = vp / sum_prob; P.swP = sw / sum_prob;
    P.volP = P.volP + P.transP; P.swP = P.swP + P.volP;
    P.prob_error = (P.swP < 0.999) ? 1 : 0;
    P.r_pos = r_pos; P.r_neg = r_neg; P.a_pos = a_pos; P.a_neg = a_neg;
    P.log_r_pos = std::log(r_pos); P.log_r_neg = std::log(r_neg);
    P.inv_log_r_pos = 1.0 / P.log_r_pos; P.inv_log_r_neg = 1.0 / P.log_r_neg;
    P.c_pos = (1.0 - r_pos) / a_pos; P.c_neg = (1.0 - r_neg) / a_neg;
    P.av_binwidth = av_bw; P.log_unbiased_norm = lun;
    P.mu_min = u.mu_min; P.mu_max = u.mu_max;
    P.orig_wl_factor = orig_wl_factor; P.wl_alpha = u.wl_alpha;
    P.seed = 20141211ull; P.stream0 = (unsigned)first_rank; P.rng_mode = 0;
    P.nbins = nb;
    P.npt = u.npt; P.eta_interp = u.eta_interp; P.samplerun = u.samplerun; P.leshift = u.leshift;
    P.always_switch = (c->nlat == 2) ? u.mc_always_switch : 0; P.dd = u.dd; P.wl_swetnam = u.wl_swetnam;
    P.list_update_int = u.list_update_int; P.eq_mc_cycles = u.eq_mc_cycles;
    c->mc_ready = true;
    return 0;
}

// one field of the per-walker scalars, for one walker or (walker < 0) all of them: no host round trip
enum ScalarField : int { SF_RNG_INDEX = 0, SF_WL_FACTOR, SF_LS, SF_WMIN_ZERO };
__global__ void k_set_scalar(DeviceState S, int walker, int field, double dval, unsigned long long uval, int ival)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= S.W || (walker >= 0 && w != walker)) return;
    WalkerScalars& sc = S.scal[w];
    switch (field) {
    case SF_RNG_INDEX: sc.rng_index = uval; break;
    case SF_WL_FACTOR: sc.wl_factor = dval; sc.wl_invt_active = ival; break;
    case SF_LS: sc.ls = ival; break;
    case SF_WMIN_ZERO: sc.wmin_zero = ival; break;
    default: break;
    }
}

static int set_scalar(mwgpu_ctx* c, int walker, int field, double dval, unsigned long long uval, int ival)
{
    k_set_scalar<<<(c->W + 127) / 128, 128, 0, c->stream>>>(c->S, walker, field, dval, uval, ival);
    c->launches++;
    return finish(c, true);
}

extern "C" int mwgpu_mc_set_rng_philox(mwgpu_ctx* c, uint64_t seed, uint32_t first_stream, uint64_t start_index)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_set_rng_philox: call mwgpu_mc_init first");
    c->P.seed = seed; c->P.stream0 = first_stream; c->P.rng_mode = 0;
    return set_scalar(c, -1, SF_RNG_INDEX, 0.0, start_index, 0);
}

extern "C" int mwgpu_mc_set_rng_index(mwgpu_ctx* c, int walker, uint64_t index)
{
    if (int rc = check_ctx(c, walker, true)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_set_rng_index: call mwgpu_mc_init first");
    return set_scalar(c, walker, SF_RNG_INDEX, 0.0, index, 0);
}

extern "C" int mwgpu_mc_set_rng_fifo(mwgpu_ctx* c, const double* u, int64_t n)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_set_rng_fifo: call mwgpu_mc_init first");
    if (c->W != 1) return fail("mwgpu_mc_set_rng_fifo: the host FIFO serves single-walker contexts only");
    if (!u || n < 1) return fail("mwgpu_mc_set_rng_fifo: empty FIFO");
    if (c->fifo) { cudaFree(c->fifo); c->fifo = nullptr; }
    CUDA_TRY(cudaMalloc((void**)&c->fifo, sizeof(double) * n));
    CUDA_TRY(cudaMemcpy(c->fifo, u, sizeof(double) * n, cudaMemcpyHostToDevice));
    c->S.fifo = c->fifo; c->S.fifo_len = (unsigned long long)n;
    c->P.rng_mode = 1;
    return set_scalar(c, 0, SF_RNG_INDEX, 0.0, 0ull, 0);
}

// ------------------------------------------------------------------------------------------------
// the hot loop
// ------------------------------------------------------------------------------------------------
static int mc_run_impl(mwgpu_ctx* c, int ncycles, bool sync)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_run: call mwgpu_mc_init first");
    if (ncycles < 0) return fail("mwgpu_mc_run: ncycles must be >= 0");
    const size_t smem = walker_smem_bytes(c->N, c->nlat);
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    // the 48-molecule boxes of every reference deck get the kernel with N folded into its addressing
#define MW_LAUNCH_MC(NLAT_, NT_)                                                                                  \
    do {                                                                                                          \
        CUDA_TRY(cudaFuncSetAttribute(k_mc_run<NLAT_, NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_mc_run<NLAT_, NT_><<<c->W, 32, smem, c->stream>>>(c->S, c->P, ncycles);                                 \
    } while (0)
    // The warp-per-lattice kernel is persistent: as many blocks as the GPU holds at once (never more than walkers)
    // take (walker, chunk of cycles) units from a queue.  A batch that fits the GPU at once runs one unit per walker;
    // a larger one is cut into units of MW2_CHUNK cycles so that it does not end on its slowest walkers (mw2.cuh).
#define MW_LAUNCH_MC2(NLAT_, NT_, BL_, WPL_, ILP_)                                                                    \
    do {                                                                                                          \
        auto kern = v2::k_mc_run2<NLAT_, NT_, BL_, WPL_, ILP_>;                                                   \
        constexpr int nthr = 32 * NLAT_ * WPL_;                                                                   \
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));            \
        int per_sm = 0;                                                                                           \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthr, smem2));                      \
        if (per_sm < 1) return fail("mwgpu_mc_run: the walker kernel does not fit on an SM");                     \
        long long slots = (long long)per_sm * c->num_sms;                                                         \
        if (c->max_blocks > 0) slots = std::min<long long>(slots, c->max_blocks);                                 \
        const int grid = (int)std::min<long long>(c->W, slots);                                                   \
        int chunk = ncycles > 0 ? ncycles : 1;                                                                    \
        if (c->chunk_cycles > 0) chunk = std::min(chunk, c->chunk_cycles);                                        \
        else if (c->W > slots) chunk = std::min(chunk, MW2_CHUNK);                                                \
        chunk = std::max(chunk, (ncycles + 511) / 512);       /* at most 512 turns per walker: bounds the queue */        \
        const size_t units = (size_t)c->W * ((size_t)(ncycles + chunk - 1) / chunk + 1);                          \
        if (units > c->queue_ints) {                                                                              \
            if (c->S.queue) { CUDA_TRY(cudaStreamSynchronize(c->stream)); cudaFree(c->S.queue); c->S.queue = nullptr; } \
            if (int rc = dalloc(&c->S.queue, units)) return rc;                                                   \
            c->queue_ints = units;                                                                                \
        }                                                                                                         \
        if (chunk < ncycles) CUDA_TRY(cudaMemsetAsync(c->S.queue, 0, sizeof(int) * units, c->stream));            \
        CUDA_TRY(cudaMemsetAsync(c->S.qctr, 0, sizeof(int) * 8, c->stream));                                      \
        kern<<<grid, nthr, smem2, c->stream>>>(c->S, c->P, ncycles, chunk);                                       \
    } while (0)
    // boxes of up to 64 molecules: one warp per lattice on a per-lattice shared-memory block (mw2.cuh); larger
    // boxes (and walker_kernel == 1): the first-generation kernel, one warp per walker
    const bool gen2 = ent_has_rev(c->N) && c->walker_kernel != 1;
    const size_t smem2 = v2::walker_bytes(c->N, c->nlat);
    // small ensembles (at most MW2_BLOCKS / 2 walkers per SM): the instantiation with the larger register budget
    const bool small = (long long)c->W * 2 <= (long long)c->num_sms * MW2_BLOCKS;
    // Two warps per lattice (four per walker) on request only: measured on 512 walkers per GPU the split of the item
    // passes and pair sums of one lattice shortens a walker's step by 1.5 % -- the serial acceptance (35 % of a move)
    // and the dependent chains inside one pass bound it, not the number of passes (profiles/README.md).
    const bool quad = c->nlat == 2 && c->walker_kernel == 4;
    if (gen2 && quad) {
        if (c->N == 48) MW_LAUNCH_MC2(2, 48, 4, 2, 1); else MW_LAUNCH_MC2(2, 0, 4, 2, 1);
    } else if (gen2 && c->nlat == 2) {
        // at most four walkers per SM: up to three item passes in flight per warp (registers to spare, idle issue
        // slots; at up to seven per SM two in flight measured no gain: 64-67 ms per step at 1024 walkers either way);
        // automatic selection only -- an explicit mwgpu_mc_set_kernel(2) keeps one pass per turn, the code the full
        // GPU runs, so the parity tests hold both forms to the oracle
        const bool tiny = c->walker_kernel == 0 && (long long)c->W <= 4ll * c->num_sms;
        if (c->N == 48) {
            if (tiny) MW_LAUNCH_MC2(2, 48, 4, 1, 3);
            else if (small) MW_LAUNCH_MC2(2, 48, MW2_BLOCKS / 2, 1, 1);
            else MW_LAUNCH_MC2(2, 48, MW2_BLOCKS, 1, 1);
        }
        else MW_LAUNCH_MC2(2, 0, MW2_BLOCKS, 1, 1);
    } else if (gen2) {
        if (c->N == 48) MW_LAUNCH_MC2(1, 48, MW2_BLOCKS, 1, 1); else MW_LAUNCH_MC2(1, 0, MW2_BLOCKS, 1, 1);
    }
    else if (c->nlat == 2)    { if (c->N == 48) MW_LAUNCH_MC(2, 48); else MW_LAUNCH_MC(2, 0); }
    else                      { if (c->N == 48) MW_LAUNCH_MC(1, 48); else MW_LAUNCH_MC(1, 0); }
#undef MW_LAUNCH_MC
#undef MW_LAUNCH_MC2
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    c->launches++;
    if (int rc = finish(c, sync)) return rc;
    if (sync) return collect_errors(c, "mwgpu_mc_run");
    return 0;
}

extern "C" int mwgpu_mc_set_kernel(mwgpu_ctx* c, int warps_per_walker)
{
    if (!c) return fail("mwgpu_mc_set_kernel: NULL context");
    if (warps_per_walker != 0 && warps_per_walker != 1 && warps_per_walker != 2 && warps_per_walker != 4)
        return fail("mwgpu_mc_set_kernel: 0 (automatic), 1, 2 or 4 warps per walker");
    if (warps_per_walker >= 2 && !ent_has_rev(c->N))
        return fail("mwgpu_mc_set_kernel: the warp-per-lattice kernels need boxes of up to 64 molecules");
    if (warps_per_walker == 4 && c->nlat != 2)
        return fail("mwgpu_mc_set_kernel: four warps per walker = two per lattice of a lattice-switch box");
    c->walker_kernel = warps_per_walker;
    c->energy_kernel = (warps_per_walker == 1) ? 1 : 0;       // generation 1 keeps its own batched energy kernel
    return 0;
}

extern "C" int mwgpu_mc_set_schedule(mwgpu_ctx* c, int chunk_cycles, int max_blocks)
{
    if (!c) return fail("mwgpu_mc_set_schedule: NULL context");
    if (chunk_cycles < 0 || max_blocks < 0) return fail("mwgpu_mc_set_schedule: chunk_cycles and max_blocks must be >= 0 (0 = automatic)");
    c->chunk_cycles = chunk_cycles; c->max_blocks = max_blocks;
    return 0;
}

extern "C" int mwgpu_mc_get_walker_times(mwgpu_ctx* c, uint64_t* start_end_ns)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!start_end_ns) return fail("mwgpu_mc_get_walker_times: NULL");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(start_end_ns, c->S.wtime, sizeof(uint64_t) * 2 * (size_t)c->W, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mwgpu_mc_run(mwgpu_ctx* c, int ncycles) { return mc_run_impl(c, ncycles, true); }
extern "C" int mwgpu_mc_run_async(mwgpu_ctx* c, int ncycles) { return mc_run_impl(c, ncycles, false); }

static void fill_state(const WalkerScalars& s, mwgpu_walker_state* o)
{
    o->model_energy[0] = s.E[0]; o->model_energy[1] = s.E[1];
    o->volume[0] = s.vol[0]; o->volume[1] = s.vol[1];
    o->ls_mu = s.mu; o->mc_max_trans = s.max_trans; o->mc_dv_max = s.dv_max; o->wl_factor = s.wl_factor;
    o->my_mu_min = s.mu_lo; o->my_mu_max = s.mu_hi;
    o->average_energy[0] = s.avgE[0]; o->average_energy[1] = s.avgE[1];
    o->min_dmu = s.min_dmu; o->max_dmu = s.max_dmu;
    o->ref_enthalpy[0] = s.refH[0]; o->ref_enthalpy[1] = s.refH[1];
    o->rng_index = (int64_t)s.rng_index;
    o->ls = s.ls; o->mc_cycle_num = s.cycle;
    o->accepted[0] = s.acc_r; o->accepted[1] = s.acc_v; o->accepted[2] = s.acc_s;
    o->attempted[0] = s.att_r; o->attempted[1] = s.att_v; o->attempted[2] = s.att_s;
    o->my_start_bin = s.start_bin; o->my_end_bin = s.end_bin;
    o->walker_in_window = s.in_window; o->error = s.error; o->wl_invt_active = s.wl_invt_active;
}

extern "C" int mwgpu_mc_get_state(mwgpu_ctx* c, int walker, mwgpu_walker_state* out)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (!out) return fail("mwgpu_mc_get_state: out is NULL");
    WalkerScalars s;
    CUDA_TRY(cudaMemcpyAsync(&s, c->S.scal + walker, sizeof(s), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    fill_state(s, out);
    return 0;
}

extern "C" int mwgpu_mc_get_states(mwgpu_ctx* c, mwgpu_walker_state* out)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!out) return fail("mwgpu_mc_get_states: out is NULL");
    std::vector<WalkerScalars> h(c->W);
    CUDA_TRY(cudaMemcpyAsync(h.data(), c->S.scal, sizeof(WalkerScalars) * c->W, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int w = 0; w < c->W; ++w) fill_state(h[w], out + w);
    return 0;
}

extern "C" int mwgpu_mc_get_translations(mwgpu_ctx* c, int walker, int* t)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (!t) return fail("mwgpu_mc_get_translations: NULL");
    CUDA_TRY(cudaMemcpy(t, c->S.transcount + (size_t)walker * c->N, sizeof(int) * c->N, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mwgpu_mc_get_bins(mwgpu_ctx* c, int walker, double* weight, double* hist, double* uhist)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_get_bins: call mwgpu_mc_init first");
    const size_t off = (size_t)walker * c->NB, nbytes = sizeof(double) * c->NB;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (weight) CUDA_TRY(cudaMemcpy(weight, c->S.weight + off, nbytes, cudaMemcpyDeviceToHost));
    if (hist) CUDA_TRY(cudaMemcpy(hist, c->S.hist + off, nbytes, cudaMemcpyDeviceToHost));
    if (uhist) CUDA_TRY(cudaMemcpy(uhist, c->S.uhist + off, nbytes, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mwgpu_mc_set_bins(mwgpu_ctx* c, int walker, const double* weight, const double* hist, const double* uhist)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_set_bins: call mwgpu_mc_init first");
    const size_t off = (size_t)walker * c->NB, nbytes = sizeof(double) * c->NB;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (weight) CUDA_TRY(cudaMemcpy(c->S.weight + off, weight, nbytes, cudaMemcpyHostToDevice));
    if (hist) CUDA_TRY(cudaMemcpy(c->S.hist + off, hist, nbytes, cudaMemcpyHostToDevice));
    if (uhist) CUDA_TRY(cudaMemcpy(c->S.uhist + off, uhist, nbytes, cudaMemcpyHostToDevice));
    if (weight) return set_scalar(c, walker, SF_WMIN_ZERO, 0.0, 0ull, 0);    // the wl-bin update may no longer assume min(weight) == 0
    return 0;
}

extern "C" int mwgpu_mc_get_grid(mwgpu_ctx* c, double* mu_bin, double* binwidth, double* scalars)
{
    if (!c || !c->mc_ready) return fail("mwgpu_mc_get_grid: call mwgpu_mc_init first");
    if (mu_bin) memcpy(mu_bin, c->h_mubin.data(), sizeof(double) * c->NB);
    if (binwidth) memcpy(binwidth, c->h_binwidth.data(), sizeof(double) * c->NB);
    if (scalars) {
        scalars[0] = c->P.r_pos; scalars[1] = c->P.r_neg; scalars[2] = c->P.av_binwidth;
        scalars[3] = c->P.log_unbiased_norm;
    }
    return 0;
}

extern "C" int mwgpu_mc_set_wl_factor(mwgpu_ctx* c, int walker, double wl_factor, int wl_invt_active)
{
    if (int rc = check_ctx(c, walker, true)) return rc;
    return set_scalar(c, walker, SF_WL_FACTOR, wl_factor, 0ull, wl_invt_active);
}

extern "C" int mwgpu_mc_set_active_lattice(mwgpu_ctx* c, int walker, int ls)
{
    if (int rc = check_ctx(c, walker, true)) return rc;
    if (int rc = check_ils(c, ls)) return rc;
    return set_scalar(c, walker, SF_LS, 0.0, 0ull, ls);
}

extern "C" int mwgpu_mc_monitor(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_monitor: call mwgpu_mc_init first");
    if (int rc = launch_op(c, make_args(c, OP_MONITOR, 0, -1, 0, nullptr), c->W)) return rc;
    return collect_errors(c, "mwgpu_mc_monitor");
}

extern "C" int mwgpu_mc_chain_sync(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_chain_sync: call mwgpu_mc_init first");
    if (c->nlat != 2) return 0;
    if (int rc = launch_op(c, make_args(c, OP_CHAIN_SYNC, 0, -1, 0, nullptr), c->W)) return rc;
    return collect_errors(c, "mwgpu_mc_chain_sync");
}

// ------------------------------------------------------------------------------------------------
// restart and therm rows
// ------------------------------------------------------------------------------------------------
// State effects of mc_checkpoint_load (mc_moves.F90:403-501) + the refresh of mc_init (:842-862) for one walker.
// ref_hmatrix and the neighbour lists stay what the start-up sequence made them (the reference does not touch
// them on a restart either).
extern "C" int mwgpu_mc_restore(mwgpu_ctx* c, int walker, int mc_cycle_num, double mc_max_trans, double mc_dv_max,
                                double wl_factor, int wl_invt_active, int ls,
                                const double* histogram, const double* weight, const double* unbiased_hist,
                                const double* hmatrix, const double* ref_ljr, const double* ljr)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_restore: call mwgpu_mc_init first");
    if (!histogram || !weight) return fail("mwgpu_mc_restore: histogram / weight is NULL");
    if (!hmatrix || !ref_ljr || !ljr) return fail("mwgpu_mc_restore: hmatrix / ref_ljr / ljr is NULL");
    if (ls < 1 || ls > c->nlat) return fail("mwgpu_mc_restore: ls out of range");
    if (c->user.samplerun && !unbiased_hist) return fail("mwgpu_mc_restore: a sample run needs unbiased_hist");
    // every argument is checked before the walker is touched: a rejected call leaves it as it was
    if (int rc = upload_impl(c, walker, 1, 0, ljr, ref_ljr, hmatrix, true)) return rc;                 // :479-489
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const size_t off = (size_t)walker * c->NB, nbytes = sizeof(double) * c->NB;
    WalkerScalars s;
    CUDA_TRY(cudaMemcpy(&s, c->S.scal + walker, sizeof(s), cudaMemcpyDeviceToHost));
    s.cycle = mc_cycle_num; s.max_trans = mc_max_trans; s.dv_max = mc_dv_max;          // :449-455
    s.wl_factor = wl_factor; s.wl_invt_active = wl_invt_active ? 1 : 0; s.ls = ls;
    double sum = 0.0;
    for (int k = 0; k < c->NB; ++k) sum = sum + histogram[k];
    s.sumhist = sum;                                                                    // :470
    if (wl_factor < c->P.orig_wl_factor) s.firstcycle = 0;                              // :473-476
    s.wmin_zero = 0;
    CUDA_TRY(cudaMemcpy(c->S.scal + walker, &s, sizeof(s), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->S.hist + off, histogram, nbytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->S.weight + off, weight, nbytes, cudaMemcpyHostToDevice));
    if (c->user.samplerun) CUDA_TRY(cudaMemcpy(c->S.uhist + off, unbiased_hist, nbytes, cudaMemcpyHostToDevice));
    if (!c->user.dd) {                                                                  // :462-466 comms_set_(u)histogram
        CUDA_TRY(cudaMemcpy(c->S.hbase + off, histogram, nbytes, cudaMemcpyHostToDevice));
        if (c->user.samplerun) CUDA_TRY(cudaMemcpy(c->S.ubase + off, unbiased_hist, nbytes, cudaMemcpyHostToDevice));
    }
    if (int rc = launch_op(c, make_args(c, OP_RESTART, walker, -1, 0, nullptr), 1)) return rc;
    return collect_errors(c, "mwgpu_mc_restore");
}

extern "C" int mwgpu_mc_set_therm(mwgpu_ctx* c, int file_output_int, int capacity)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (file_output_int < 0 || capacity < 0) return fail("mwgpu_mc_set_therm: negative argument");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->S.therm) cudaFree(c->S.therm);
    if (c->S.therm_n) cudaFree(c->S.therm_n);
    c->S.therm = nullptr; c->S.therm_n = nullptr; c->S.therm_int = 0; c->S.therm_cap = 0;
    if (file_output_int == 0 || capacity == 0) return 0;
    if (int rc = dalloc(&c->S.therm, (size_t)c->W * capacity * THERM_ROW)) return rc;
    if (int rc = dalloc(&c->S.therm_n, (size_t)c->W)) return rc;
    c->S.therm_int = file_output_int; c->S.therm_cap = capacity;
    return 0;
}

extern "C" int mwgpu_mc_get_therm(mwgpu_ctx* c, int walker, mwgpu_therm_row* rows, int max_rows, int* nrows, int* ndropped)
{
    if (int rc = check_ctx(c, walker, false)) return rc;
    if (!nrows) return fail("mwgpu_mc_get_therm: nrows is NULL");
    *nrows = 0; if (ndropped) *ndropped = 0;
    if (!c->S.therm) return 0;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int n = 0;
    CUDA_TRY(cudaMemcpy(&n, c->S.therm_n + walker, sizeof(int), cudaMemcpyDeviceToHost));
    const int have = n < c->S.therm_cap ? n : c->S.therm_cap;
    if (have > max_rows) return fail("mwgpu_mc_get_therm: rows buffer too small");
    if (have > 0 && !rows) return fail("mwgpu_mc_get_therm: rows is NULL");
    static_assert(sizeof(mwgpu_therm_row) == sizeof(double) * THERM_ROW, "therm row layout");
    std::vector<double> tmp((size_t)have * THERM_ROW);
    if (have > 0)
        CUDA_TRY(cudaMemcpy(tmp.data(), c->S.therm + (size_t)walker * c->S.therm_cap * THERM_ROW,
                            sizeof(double) * have * THERM_ROW, cudaMemcpyDeviceToHost));
    for (int r = 0; r < have; ++r) {
        const double* t = tmp.data() + (size_t)r * THERM_ROW;
        mwgpu_therm_row& o = rows[r];
        o.icyc = (int64_t)t[0]; o.ls = (int64_t)t[1];
        o.model_energy[0] = t[2]; o.model_energy[1] = t[3]; o.ls_mu = t[4];
        o.volume[0] = t[5]; o.volume[1] = t[6];
        for (int k = 0; k < 9; ++k) o.hmatrix1[k] = t[7 + k];
    }
    const int zero = 0;
    CUDA_TRY(cudaMemcpy(c->S.therm_n + walker, &zero, sizeof(int), cudaMemcpyHostToDevice));
    *nrows = have; if (ndropped) *ndropped = n - have;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// comms: delta-since-last-sync all-reduce of weight / histogram / unbiased_hist
// (comms_mpi.f90:244-277, :461-530)
// ------------------------------------------------------------------------------------------------
// delta[a][k] = sum over walkers (in walker order) of arr_a[w][k] - base_a[w][k]
__global__ void k_reduce_bins(DeviceState S, double* __restrict__ delta, int NBP, int first, int narr)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= narr * NBP) return;
    t += first * NBP;
    const int a = t / NBP, k = t % NBP;
    double s = 0.0;
    if (k < S.NB) {
        const double* arr = (a == 0) ? S.weight : (a == 1) ? S.hist : S.uhist;
        const double* base = (a == 0) ? S.wbase : (a == 1) ? S.hbase : S.ubase;
        for (int w = 0; w < S.W; ++w) s = s + (arr[(size_t)w * S.NB + k] - base[(size_t)w * S.NB + k]);
    }
    delta[t] = s;
}

// The same sum for large batches: one CTA per (array, bin); thread t adds the walkers t, t+256, ... in
// walker order, then a fixed shared-memory tree.  Deterministic, but not the serial order (MPI leaves the
// order of its reduction unspecified as well); batches of up to REDUCE_SERIAL_MAX walkers keep the serial
// kernel, whose order the oracle reproduces bit for bit.
constexpr int REDUCE_SERIAL_MAX = 256;
__global__ void __launch_bounds__(256) k_reduce_bins_tree(DeviceState S, double* __restrict__ delta, int NBP, int first)
{
    __shared__ double part[256];
    const int a = first + blockIdx.x / NBP, k = blockIdx.x % NBP;
    double s = 0.0;
    if (k < S.NB) {
        const double* arr = (a == 0) ? S.weight : (a == 1) ? S.hist : S.uhist;
        const double* base = (a == 0) ? S.wbase : (a == 1) ? S.hbase : S.ubase;
        for (int w = threadIdx.x; w < S.W; w += 256) s = s + (arr[(size_t)w * S.NB + k] - base[(size_t)w * S.NB + k]);
    }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) part[threadIdx.x] += part[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) delta[a * NBP + k] = part[0];
}

static void launch_reduce_bins(mwgpu_ctx* c, int first, int narr);

// arr = total + base ; base = arr
__global__ void k_apply_bins(DeviceState S, const double* __restrict__ delta, int NBP, int first, int narr)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t per = (size_t)S.W * S.NB;
    if (t >= per * narr) return;
    const int a = first + (int)(t / per);
    const size_t r = t % per;
    const int k = (int)(r % S.NB);
    double* arr = (a == 0) ? S.weight : (a == 1) ? S.hist : S.uhist;
    double* base = (a == 0) ? S.wbase : (a == 1) ? S.hbase : S.ubase;
    const double v = delta[a * NBP + k] + base[r];
    arr[r] = v; base[r] = v;
}

__global__ void k_clear_wmin(DeviceState S)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < S.W) S.scal[w].wmin_zero = 0;
}

static int narr_of(const mwgpu_ctx* c) { return c->user.samplerun ? 3 : 2; }

static void launch_reduce_bins(mwgpu_ctx* c, int first, int narr)
{
    const int n = narr * c->NBP;
    if (c->W <= REDUCE_SERIAL_MAX) k_reduce_bins<<<(n + 127) / 128, 128, 0, c->stream>>>(c->S, c->delta, c->NBP, first, narr);
    else k_reduce_bins_tree<<<n, 256, 0, c->stream>>>(c->S, c->delta, c->NBP, first);
    c->launches++;
}

extern "C" int mwgpu_comms_reduce_local(mwgpu_ctx* c, void** dev_ptr, int* count)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_comms_reduce_local: call mwgpu_mc_init first");
    const int narr = narr_of(c), n = narr * c->NBP;
    launch_reduce_bins(c, 0, narr);
    if (dev_ptr) *dev_ptr = c->delta;
    if (count) *count = n;
    return finish(c, dev_ptr != nullptr);
}

extern "C" int mwgpu_comms_apply(mwgpu_ctx* c)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_comms_apply: call mwgpu_mc_init first");
    const int narr = narr_of(c);
    const size_t n = (size_t)c->W * c->NB * narr;
    k_apply_bins<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->S, c->delta, c->NBP, 0, narr);
    k_clear_wmin<<<(c->W + 127) / 128, 128, 0, c->stream>>>(c->S);
    c->launches += 2;
    return finish(c, true);
}

// base[w][k] = src[k] for every walker
__global__ void k_broadcast_bins(double* __restrict__ base, const double* __restrict__ src, int W, int NB)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t < (size_t)W * NB) base[t] = src[t % NB];
}

extern "C" int mwgpu_comms_set_hist_base(mwgpu_ctx* c, const double* hist, const double* uhist)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_comms_set_hist_base: call mwgpu_mc_init first");
    // one upload per array into the reduction staging buffer, one broadcast kernel: no per-walker copies
    const size_t n = (size_t)c->W * c->NB;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (hist) {
        CUDA_TRY(cudaMemcpyAsync(c->delta, hist, sizeof(double) * c->NB, cudaMemcpyHostToDevice, c->stream));
        k_broadcast_bins<<<grid, 256, 0, c->stream>>>(c->S.hbase, c->delta, c->W, c->NB);
        c->launches++;
    }
    if (uhist) {
        CUDA_TRY(cudaMemcpyAsync(c->delta + c->NBP, uhist, sizeof(double) * c->NB, cudaMemcpyHostToDevice, c->stream));
        k_broadcast_bins<<<grid, 256, 0, c->stream>>>(c->S.ubase, c->delta + c->NBP, c->W, c->NB);
        c->launches++;
    }
    return finish(c, true);
}

// ---- NCCL, loaded lazily so that the library has no link-time dependency on it -------------------
typedef struct { char internal[128]; } nccl_unique_id;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(void**, int, nccl_unique_id, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.handle) break; }
    if (!g_nccl.handle) return fail(std::string("mwgpu_comms: cannot load NCCL: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(nccl_unique_id*))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, nccl_unique_id, int))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail("mwgpu_comms: NCCL symbols missing");
    return 0;
}

static int nccl_fail(const char* what, int r)
{
    return fail(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"), 200 + r);
}

extern "C" int mwgpu_comms_get_unique_id(void* id128)
{
    if (!id128) return fail("mwgpu_comms_get_unique_id: NULL");
    if (int rc = nccl_load()) return rc;
    nccl_unique_id id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id128, &id, 128);
    return 0;
}

extern "C" int mwgpu_comms_init(mwgpu_ctx* c, int nranks, int rank, const void* id128)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail("mwgpu_comms_init: bad arguments");
    if (int rc = nccl_load()) return rc;
    nccl_unique_id id;
    memcpy(&id, id128, 128);
    const int r = g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank);
    if (r) return nccl_fail("ncclCommInitRank", r);
    c->nranks = nranks; c->rank = rank;
    return 0;
}

static void nccl_destroy(mwgpu_ctx* c)
{
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr;
}

extern "C" int mwgpu_comms_allreduce_bins(mwgpu_ctx* c)
{
    if (int rc = mwgpu_comms_reduce_local(c, nullptr, nullptr)) return rc;
    if (c->nccl_comm && c->nranks > 1) {
        const int n = narr_of(c) * c->NBP;
        // ncclDouble = 8, ncclSum = 0
        const int r = g_nccl.AllReduce(c->delta, c->delta, (size_t)n, 8, 0, c->nccl_comm, c->stream);
        if (r) return nccl_fail("ncclAllReduce", r);
    }
    return mwgpu_comms_apply(c);
}

// ------------------------------------------------------------------------------------------------
// Periodic bookkeeping that consumes the reduced arrays, on the device:
// mc_check_flatness (mc_moves.F90:1936-2185), mc_compute_deltaG_from_hist (:2498-2621),
// comms_join_uhist / comms_join_eta (comms_mpi.f90:299-459).  All sums run in the reference's
// order with explicit round-to-nearest operations (no FMA contraction), so the decisions and the
// new weights / histograms are bit-identical to the oracle; only log/exp differ by an ulp.
// ------------------------------------------------------------------------------------------------
// delta all-reduce of the arrays [first, first+narr) (0 weight, 1 histogram, 2 unbiased_hist)
static int allreduce_arrays(mwgpu_ctx* c, int first, int narr)
{
    const int n = narr * c->NBP;
    launch_reduce_bins(c, first, narr);
    if (c->nccl_comm && c->nranks > 1) {
        double* d = c->delta + (size_t)first * c->NBP;
        const int r = g_nccl.AllReduce(d, d, (size_t)n, 8, 0, c->nccl_comm, c->stream);
        if (r) return nccl_fail("ncclAllReduce", r);
    }
    const size_t m = (size_t)c->W * c->NB * narr;
    k_apply_bins<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->S, c->delta, c->NBP, first, narr);
    c->launches++;
    if (first == 0) { k_clear_wmin<<<(c->W + 127) / 128, 128, 0, c->stream>>>(c->S); c->launches++; }
    return finish(c, false);
}

struct FlatArgs {
    int wl_schedule, wl_minhist, wl_useinvt;
    double wl_flattol;
    int dd, wl_swetnam, nwater;
};

// :1961 guard of every walker (= rank) on its own histogram
__global__ void k_flat_guard(DeviceState S, int samplerun, int* __restrict__ skip)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= S.W) return;
    const double* hist = S.hist + (size_t)w * S.NB;
    double sum = 0.0;
    for (int k = 0; k < S.NB; ++k) sum = __dadd_rn(sum, hist[k]);
    skip[w] = samplerun || (sum < DBL_MIN);
}

__device__ __forceinline__ long long nint_dev(double x) { return (long long)(x < 0.0 ? __dadd_rn(x, -0.5) : __dadd_rn(x, 0.5)); }

// :1968-2142, one thread per walker.  In 'mw' runs every walker holds the same reduced histogram and
// the same window, so the flatness decision is identical on all of them (the reference broadcasts
// rank 0's, comms_bcastlog :2042).
__global__ void k_flat_decide(DeviceState S, FlatArgs a, const int* __restrict__ skip, mwgpu_flat_report* __restrict__ rep)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= S.W) return;
    const int nb = S.NB;
    WalkerScalars& sc = S.scal[w];
    double* hist = S.hist + (size_t)w * nb;
    double* hbase = S.hbase + (size_t)w * nb;
    double* wgt = S.weight + (size_t)w * nb;
    mwgpu_flat_report rr;
    rr.checked = 0; rr.hist_reset = 0; rr.flat = 0; rr.invt_switched = 0;
    rr.mean = 0.0; rr.max_pct = 0.0; rr.min_pct = 0.0; rr.wl_factor = sc.wl_factor;
    if (skip[w]) { if (w == 0) *rep = rr; return; }
    rr.checked = 1;
    double mn = hist[0], mx = hist[0];
    for (int k = 1; k < nb; ++k) { const double h = hist[k]; if (h < mn) mn = h; if (h > mx) mx = h; }
    if (sc.firstcycle && !sc.hist_reset && nint_dev(mn) > (long long)a.wl_minhist) {      // :1973-1980
        sc.hist_reset = 1;
        for (int k = 0; k < nb; ++k) { hist[k] = 0.0; hbase[k] = 0.0; }
        rr.hist_reset = 1;
        if (w == 0) *rep = rr;
        return;
    }
    const int sb = sc.start_bin, eb = sc.end_bin;
    double av = 0.0;
    for (int k = sb; k <= eb; ++k) av = __dadd_rn(av, hist[k - 1]);                       // :1983-1989
    av = __ddiv_rn(av, (double)(eb - sb + 1));
    rr.mean = av;
    rr.max_pct = __ddiv_rn(__dmul_rn(100.0, mx), av);
    rr.min_pct = __ddiv_rn(__dmul_rn(100.0, mn), av);
    if (!(sc.wl_invt_active || a.wl_swetnam)) {
        bool flat = true;
        if (a.wl_schedule == 0) {
            for (int k = sb; k <= eb; ++k)
                if (__ddiv_rn(fabs(__dsub_rn(hist[k - 1], av)), av) > a.wl_flattol) flat = false;
        } else if (a.wl_schedule == 1) {
            double m2 = hist[sb - 1];
            for (int k = sb; k <= eb; ++k) if (hist[k - 1] < m2) m2 = hist[k - 1];
            if (nint_dev(m2) < (long long)a.wl_minhist) flat = false;
        } else {
            const double thr = __dmul_rn(__dsub_rn(1.0, a.wl_flattol), av);
            for (int k = sb; k <= eb; ++k) if (hist[k - 1] < thr) flat = false;
        }
        if (flat) {
            if (!a.dd) {
                const double mid = wgt[nb / 2];                                           // weight(nbins/2+1)
                for (int k = 0; k < nb; ++k) wgt[k] = __dsub_rn(wgt[k], mid);
                for (int k = 0; k < nb; ++k) { hist[k] = 0.0; hbase[k] = 0.0; }
                sc.wmin_zero = 0;
            } else {
                for (int k = 0; k < nb; ++k) hist[k] = 0.0;
            }
            sc.wl_factor = __dmul_rn(sc.wl_factor, 0.5);
            sc.firstcycle = 0;
        }
        rr.flat = flat ? 1 : 0;
        const double wl_invt = __ddiv_rn((double)nb, (double)(sc.cycle * a.nwater));     // :2134-2142
        if (sc.wl_factor < wl_invt && sc.wl_factor > DBL_MIN && a.wl_useinvt) {
            sc.wl_invt_active = 1; sc.wl_factor = wl_invt; rr.invt_switched = 1;
        }
    }
    rr.wl_factor = sc.wl_factor;
    if (w == 0) *rep = rr;
}

static int book_alloc(mwgpu_ctx* c)
{
    if (!c->book) if (int rc = dalloc(&c->book, (size_t)4 * c->NBP + 64)) return rc;
    if (!c->skip) if (int rc = dalloc(&c->skip, (size_t)c->W)) return rc;
    return 0;
}

extern "C" int mwgpu_mc_check_flatness(mwgpu_ctx* c, const mwgpu_flat_params* fp, mwgpu_flat_report* rep)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_check_flatness: call mwgpu_mc_init first");
    if (!fp) return fail("mwgpu_mc_check_flatness: params is NULL");
    if (fp->wl_schedule < 0 || fp->wl_schedule > 2) return fail("Error - unknown wl_schedule value");   // :2036
    mwgpu_flat_report r0; memset(&r0, 0, sizeof(r0));
    if (c->nlat != 2 || c->user.samplerun) {                           // :293 (two lattices only), :1961
        if (rep) *rep = r0;
        return 0;
    }
    if (int rc = book_alloc(c)) return rc;
    k_flat_guard<<<(c->W + 127) / 128, 128, 0, c->stream>>>(c->S, c->user.samplerun, c->skip);
    c->launches++;
    int skip0 = 0;
    CUDA_TRY(cudaMemcpyAsync(&skip0, c->skip, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (!c->user.dd && !skip0)                                          // :1964-1966 comms_allreduce_hist
        if (int rc = allreduce_arrays(c, 1, 1)) return rc;
    FlatArgs a;
    a.wl_schedule = fp->wl_schedule; a.wl_minhist = fp->wl_minhist; a.wl_useinvt = fp->wl_useinvt;
    a.wl_flattol = fp->wl_flattol; a.dd = c->user.dd; a.wl_swetnam = c->user.wl_swetnam; a.nwater = c->N;
    mwgpu_flat_report* drep = (mwgpu_flat_report*)c->book;
    k_flat_decide<<<(c->W + 63) / 64, 64, 0, c->stream>>>(c->S, a, c->skip, drep);
    c->launches++;
    if (int rc = finish(c, false)) return rc;
    CUDA_TRY(cudaMemcpyAsync(&r0, drep, sizeof(r0), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (rep) *rep = r0;
    return 0;
}

// comms_join_uhist (is_eta = 0, comms_mpi.f90:299-375) / comms_join_eta (is_eta = 1, :377-459): rank 0's
// sequential stitch of the windows, one thread (size - 1 seams of 2*overlap+1 bins, nbins ~ 100)
__global__ void k_join_windows(const double* __restrict__ arrs, int size, int nb, int overlap, int is_eta,
                               double* __restrict__ joined)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int bpw = nb / size;
    for (int k = 0; k < nb; ++k) joined[k] = arrs[k];
    for (int ir = 1; ir < size; ++ir) {
        const double* recv = arrs + (size_t)ir * nb;
        const int my_end = ir * bpw;
        const int lo = max(my_end - overlap, 1), hi = min(my_end + overlap, nb);      // memory safety only
        double myave = 0.0, nextav = 0.0;
        for (int k = lo; k <= hi; ++k) myave = __dadd_rn(myave, is_eta ? joined[k - 1] : log(joined[k - 1]));
        myave = __ddiv_rn(myave, (double)(2 * overlap + 1));
        for (int k = lo; k <= hi; ++k) nextav = __dadd_rn(nextav, is_eta ? recv[k - 1] : log(recv[k - 1]));
        nextav = __ddiv_rn(nextav, (double)(2 * overlap + 1));
        double shift = __dsub_rn(myave, nextav);
        if (is_eta) {
            for (int k = my_end + 1; k <= nb; ++k) joined[k - 1] = __dadd_rn(recv[k - 1], shift);
        } else {
            if (isnan(shift)) shift = 0.0;
            const double f = exp(shift);
            for (int k = my_end + 1; k <= nb; ++k) joined[k - 1] = __dmul_rn(recv[k - 1], f);
        }
    }
    if (is_eta) {
        const double mid = joined[nb / 2];
        for (int k = 0; k < nb; ++k) joined[k] = __dsub_rn(joined[k], mid);
    }
}

// windows of all ranks as one [size][NB] array: the context's own array on one GPU, an NCCL
// all-gather (rank order = window order) over several
static int gather_windows(mwgpu_ctx* c, const double* mine, const double** all, int* size)
{
    *all = mine; *size = c->W;
    if (!(c->nccl_comm && c->nranks > 1)) return 0;
    if (!g_nccl.AllGather) return fail("mwgpu_comms_join: ncclAllGather missing");
    const size_t per = (size_t)c->W * c->NB, need = per * c->nranks;
    if (c->gather_doubles < need) {
        if (c->gather) cudaFree(c->gather);
        c->gather = nullptr; c->gather_doubles = 0;
        if (int rc = dalloc(&c->gather, need)) return rc;
        c->gather_doubles = need;
    }
    const int r = g_nccl.AllGather(mine, c->gather, per, 8, c->nccl_comm, c->stream);
    if (r) return nccl_fail("ncclAllGather", r);
    *all = c->gather; *size = c->W * c->nranks;
    return 0;
}

static int join_common(mwgpu_ctx* c, int overlap, int is_eta, double* d_joined)
{
    const double* all = nullptr; int size = 0;
    if (int rc = gather_windows(c, is_eta ? c->S.weight : c->S.uhist, &all, &size)) return rc;
    if (overlap < 0 || size < 1 || c->NB / size - overlap < 1 || (size - 1) * (c->NB / size) + overlap > c->NB)
        return fail("mwgpu_comms_join: windows too narrow for this overlap (bins_per_window = nbins/size)");
    k_join_windows<<<1, 32, 0, c->stream>>>(all, size, c->NB, overlap, is_eta, d_joined);
    c->launches++;
    return finish(c, false);
}

static int join_entry(mwgpu_ctx* c, int overlap, int is_eta, double* joined)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_comms_join: call mwgpu_mc_init first");
    if (!joined) return fail("mwgpu_comms_join: joined is NULL");
    if (int rc = book_alloc(c)) return rc;
    double* dj = c->book + 64;
    if (int rc = join_common(c, overlap, is_eta, dj)) return rc;
    CUDA_TRY(cudaMemcpyAsync(joined, dj, sizeof(double) * c->NB, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int mwgpu_comms_join_uhist(mwgpu_ctx* c, int overlap, double* joined) { return join_entry(c, overlap, 0, joined); }
extern "C" int mwgpu_comms_join_eta(mwgpu_ctx* c, int overlap, double* joined) { return join_entry(c, overlap, 1, joined); }

// :2540-2577 on the joined unbiased histogram; out[0] = deltaG (kT), out[1..nb] = normP
__global__ void k_deltaG(const double* __restrict__ joined, const double* __restrict__ binwidth, int nb,
                         int leshift, double beta, const WalkerScalars* __restrict__ sc0, double* __restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double Pnorm = 0.0;
    for (int i = 0; i < nb; ++i) Pnorm = __dadd_rn(Pnorm, __dmul_rn(joined[i], binwidth[i]));
    double* normP = out + 1;
    for (int i = 0; i < nb; ++i) normP[i] = __ddiv_rn(joined[i], Pnorm);
    double pA = 0.0, pB = 0.0;
    for (int i = 0; i < nb / 2; ++i) pA = __dadd_rn(pA, __dmul_rn(normP[i], binwidth[i]));
    for (int i = nb / 2; i < nb; ++i) pB = __dadd_rn(pB, __dmul_rn(normP[i], binwidth[i]));
    double dG = log(__ddiv_rn(pA, pB));
    if (leshift) dG = __dsub_rn(__dadd_rn(dG, __dmul_rn(beta, sc0->refH[1])), __dmul_rn(beta, sc0->refH[0]));
    out[0] = dG;
}

extern "C" int mwgpu_mc_deltag_from_hist(mwgpu_ctx* c, double* deltaG, double* normP)
{
    if (int rc = check_ctx(c, 0, false)) return rc;
    if (!c->mc_ready) return fail("mwgpu_mc_deltag_from_hist: call mwgpu_mc_init first");
    if (c->nlat != 2 || !c->user.samplerun)                              // :293, :305
        return fail("mwgpu_mc_deltag_from_hist: needs a two-lattice sample run (mc_moves.F90:305)");
    if (int rc = book_alloc(c)) return rc;
    const double* joined = c->S.uhist;                                   // walker 0 after the all-reduce
    if (!c->user.dd) {
        if (int rc = allreduce_arrays(c, 2, 1)) return rc;              // :2530 comms_allreduce_uhist
    } else {
        double* dj = c->book + 64;
        if (int rc = join_common(c, c->user.window_overlap, 0, dj)) return rc;   // :2533 comms_join_uhist
        joined = dj;
    }
    double* out = c->book + 64 + c->NBP;
    const double beta = 1.0 / (KB * c->user.temperature);
    k_deltaG<<<1, 32, 0, c->stream>>>(joined, c->S.binwidth, c->NB, c->user.leshift, beta, c->S.scal, out);
    c->launches++;
    if (int rc = finish(c, false)) return rc;
    std::vector<double> h((size_t)c->NB + 1);
    CUDA_TRY(cudaMemcpyAsync(h.data(), out, sizeof(double) * (c->NB + 1), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (deltaG) *deltaG = h[0];
    if (normP) memcpy(normP, h.data() + 1, sizeof(double) * c->NB);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak micro-benchmark: 8 independent DFMA chains per thread
// ------------------------------------------------------------------------------------------------
__global__ void k_dfma_peak(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int mwgpu_measure_fp64_peak(int device, double* tflops)
{
    if (!tflops) return fail("mwgpu_measure_fp64_peak: NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return fail("mwgpu_measure_fp64_peak: no CUDA device", 2);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return 0;
}
