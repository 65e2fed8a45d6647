// mw2_energy.cuh -- batched compute_model_energy (molint.F90:407-499): "full mW energy evaluations / s".
//
// One warp per unit = (walker, lattice).  The unit's positions, image vectors and Verlet rows are staged in
// 7.3 KB of shared memory (5.0 KB from HBM per unit; 27 units resident per SM -- the first version carried the
// whole walker image of the move kernel, 10 KB, and held 18).  Rows are tested one per pass (lanes = slots,
// molint.F90:438-455); bonds inside the cut-off are compacted into a table of RQ records per chunk of molecules,
// evaluated (pair energy, radial factor, unit vector: :456-468; no underflow clamp: bonds beyond RCC are exact
// zeros) and paired per centre molecule (:470-492, rotation pairing of mw_device.cuh).
// A flattened walk (32 list entries per pass whatever their molecule) was built and measured: its per-pass
// bookkeeping (molecule of an entry, segments that straddle passes) costs what the idle lanes cost here
// (7 900 vs 7 000 warp-instructions per unit; profiles/README.md).
//
// Summation order: per-lane partial sums in record order, then a 5-level xor-shuffle tree (parity tolerance 1e-11).
#pragma once
#include "mw_mc.cuh"

namespace mw {
namespace v2 {

constexpr int RQ = 64;            // bond records per chunk

struct ELay {                     // byte offsets of one unit's image
    int N;
    __host__ __device__ explicit ELay(int n) : N(n) {}
    __host__ __device__ int oP()  const { return 0; }                               // [3][N] fp64
    __host__ __device__ int oV()  const { return 24 * N; }                          // [3][IVC]
    __host__ __device__ int oQ()  const { return oV() + 24 * IVC; }                 // [4][RQ]
    __host__ __device__ int oL()  const { return oQ() + 32 * RQ; }                  // [N][LC] uint16
    __host__ __device__ int oQM() const { return oL() + 2 * LC * N; }               // [RQ] uint32: n | pos << 8
    __host__ __device__ int oNN() const { return oQM() + 4 * RQ; }                  // [N] uint8
    __host__ __device__ int bytes() const { return (oNN() + N + 15) & ~15; }
};

template <int NT>
__global__ void __launch_bounds__(32, 24) k_model_energy2(const __grid_constant__ DeviceState S, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int unit = blockIdx.x;                 // walker * nlat + lat
    if (unit >= S.W * S.nlat) return;
    const int wi = unit / S.nlat, lat = unit % S.nlat;
    const int lane = lane_id();
    const unsigned lt = lt_mask();
    const int N = (NT > 0) ? NT : S.N;
    const ELay Y(N);
    double* P = (double*)(smem + Y.oP());
    double* V = (double*)(smem + Y.oV());
    double* q = (double*)(smem + Y.oQ());
    uint16_t* L = (uint16_t*)(smem + Y.oL());
    uint32_t* qmeta = (uint32_t*)(smem + Y.oQM());
    uint8_t* NN = (uint8_t*)(smem + Y.oNN());

    // ---- stage the unit: coalesced 16-byte loads (3N*8, 3*IVC*8 and N*LC*2 are multiples of 16 for even N)
    {
        const double* gp = S.pos + ((size_t)wi * S.nlat + lat) * 3 * N;
        for (int t = lane; t < 3 * N; t += 32) P[t] = gp[t];
        const double* gi = S.iv + ((size_t)wi * S.nlat + lat) * 3 * IVC;
        for (int t = lane; t < 3 * IVC; t += 32) V[t] = gi[t];
        const uint4* gl = (const uint4*)(S.list + ((size_t)wi * S.nlat + lat) * N * LC);
        uint4* sl = (uint4*)L;
        for (int t = lane; t < N * LC / 8; t += 32) sl[t] = gl[t];
        const uint8_t* gn = S.nn + ((size_t)wi * S.nlat + lat) * N;
        for (int t = lane; t < N; t += 32) NN[t] = gn[t];
    }
    __syncwarp();
    const EntFmt F = ent_fmt(N);

    double acc = 0.0;
    int a = 0;                                    // first molecule of the chunk
    while (a < N) {
        // ---- fill: one row per pass (lanes = slots) while the bonds of the row still fit the table
        int nq = 0;
        int a_next = a;
#pragma unroll 1
        for (; a_next < N; ++a_next) {
            const bool has = lane < (int)NN[a_next];
            const uint32_t e = has ? L[a_next * LC + lane] : 0u;
            const int j = e & F.jmask, img = e >> F.ishift;
            const double tx = (P[j] + V[img]) - P[a_next];
            const double ty = (P[N + j] + V[IVC + img]) - P[N + a_next];
            const double tz = (P[2 * N + j] + V[2 * IVC + img]) - P[2 * N + a_next];
            const double r2 = dist2(tx, ty, tz);
            const bool inr = has && r2 < CK.rcc2;            // beyond RCC every term of the bond is an exact 0.0 (mw_device.cuh)
            const uint32_t bal = __ballot_sync(FULL, inr);
            const int cnt = __popc(bal);
            if (nq + cnt > RQ) break;                         // cnt <= LC < RQ: a chunk always holds >= 1 molecule
            if (inr) {
                const int pos = __popc(bal & lt), io = nq + pos;
                q[io] = tx; q[RQ + io] = ty; q[2 * RQ + io] = tz; q[3 * RQ + io] = r2;
                qmeta[io] = (uint32_t)cnt | ((uint32_t)pos << 8);
            }
            nq += cnt;
        }
        __syncwarp();
        const int nrec = nq;
        // ---- bond evaluation: 0.5 * pair energy (molint.F90:464), unit vector, radial factor
        for (int b = 0; b < nrec; b += 32) {
            const int r = b + lane;
            if (r < nrec) {
                const double tx = q[r], ty = q[RQ + r], tz = q[2 * RQ + r], r2 = q[3 * RQ + r];
                double ir, isr;
                bond_radial(r2, ir, isr);
                const double e1 = exp_nc(CK.sig02 * isr);
                const double e_2 = e1 * e1, e_4 = e_2 * e_2;
                const double s2 = CK.ss * ir * ir;
                q[r] = tx * ir; q[RQ + r] = ty * ir; q[2 * RQ + r] = tz * ir; q[3 * RQ + r] = e_4 * e_2;
                acc += 0.5 * (CK.aeps * (CK.bigb * (s2 * s2) - 1.0) * (e_4 * e1));
            }
        }
        __syncwarp();
        // ---- triplets centred on each molecule of the chunk (no k==i filter: molint.F90:480-483 has none)
        for (int b = 0; b < nrec; b += 32) {
            const int r = b + lane;
            const bool act = r < nrec;
            const uint32_t qm = act ? qmeta[r] : 0u;
            const int n = qm & 255, pos = qm >> 8;
            const int half = n >> 1, send = r - pos + n;
            const bool even = !(n & 1);
            const int rr = act ? r : 0;
            const double ux = q[rr], uy = q[RQ + rr], uz = q[2 * RQ + rr];
            const double g = act ? q[3 * RQ + rr] : 0.0;
            double tb = 0.0;
            const int maxd = __reduce_max_sync(FULL, half);
#pragma unroll 1
            for (int d = 1; d <= maxd; ++d) {
                int c = r + d;
                c = (c >= send) ? c - n : c;
                const bool on = (d <= half) && !(even && d == half && pos >= half);
                c = on ? c : rr;
                const double ct = ux * q[c] + uy * q[RQ + c] + uz * q[2 * RQ + c];
                const double dd = ct - CK.cos0;
                if (on) tb += q[3 * RQ + c] * (dd * dd);
            }
            acc += CK.leps * g * tb;
        }
        __syncwarp();
        a = a_next;
    }
    const double e = warp_sum(acc);
    if (lane == 0) {
        S.scal[wi].E[lat] = e;
        if (out) out[unit] = e;
    }
}


// ================================================================================================================
// k_model_energy3 -- the same energy with ONE LANE PER MOLECULE and the three-body sum in tensor form.
//
//   sum_{b<c} g_b g_c (u_b.u_c - c0)^2  =  1/2 [ T:T - 2 c0 |v|^2 + c0^2 s^2 - (1 - c0)^2 sum_b g_b^2 ]
//   T = sum_b g_b u_b (x) u_b,   v = sum_b g_b u_b,   s = sum_b g_b          (b, c: bonds of the centre inside the cut-off)
//
// -- exact for compute_model_energy, which has no cos < 0.99 filter (molint.F90:470-483) -- so a centre needs 11
// running sums instead of a table of bond records and a pairing loop.  A CTA of 3 warps holds TWO units (2 x 48
// molecules = 96 lanes, every lane busy).  Phase 1: every lane walks its own Verlet row (stored transposed in shared
// memory, [slot][molecule], so that the lanes of a warp read consecutive addresses) and keeps the slots that may be
// in range as a bit mask in a register; this SCREEN runs in fp32 on float4 copies of the positions (relative to the
// unit's first molecule) and of the image vectors -- two 16-byte loads per entry instead of six 8-byte ones, no fp64
// pipe -- against a radius widened by 1e-5, so it has no false negatives.  Phase 2: every lane walks its mask in fp64:
// geometry, the exact r < RCC test, radial functions, one exponential, pair energy, 11 FMAs.  Nothing is compacted and nothing is paired.  The molecules' energies are summed per unit in
// molecule order by one thread (fixed order; parity tolerance 1e-11).
// ================================================================================================================
constexpr int E3_THREADS = 96;

struct E3Lay {                    // byte offsets of one unit's image; units follow each other
    int N;
    __host__ __device__ explicit E3Lay(int n) : N(n) {}
    __host__ __device__ int oP()  const { return 0; }                               // [3][N] fp64
    __host__ __device__ int oV()  const { return 24 * N; }                          // [3][IVC]
    __host__ __device__ int oL()  const { return oV() + 24 * IVC; }                 // [LC][N] uint16, TRANSPOSED rows
    __host__ __device__ int oE()  const { return oL() + 2 * LC * N; }               // [N] fp64 energies of the molecules
    __host__ __device__ int oPF() const { return (oE() + 8 * N + 15) & ~15; }       // [N] float4 positions - origin (screen)
    __host__ __device__ int oVF() const { return oPF() + 16 * N; }                  // [IVC] float4 image vectors (screen)
    __host__ __device__ int unit() const { return oVF() + 16 * IVC; }
    __host__ __device__ int upc() const { return E3_THREADS / N; }                  // units per CTA
    __host__ __device__ int bytes() const { return upc() * unit(); }
};

// Energy of the terms centred on molecule i of a staged unit (see k_model_energy3): fp32 screen of the row, then the
// exact fp64 walk of the screened slots with the tensor-form running sums.
__device__ __forceinline__ double molecule_energy(const double* P, const double* V, const uint16_t* LT, const float4* PF,
                                                  const float4* VF, int N, int i, int nni, const EntFmt F)
{
    const double px = P[i], py = P[N + i], pz = P[2 * N + i];
    // ---- phase 1: fp32 screen of my row (no false negatives: radius widened by 1e-5; positions relative to
    // the unit's first molecule keep the fp32 error of a separation below 1e-6 of the cut-off)
    const float4 pf = PF[i];
    const float rscreen = (float)(RCC * RCC * (1.0 + 1e-5));
    uint32_t mask = 0;
#pragma unroll 2
    for (int s = 0; s < nni; ++s) {
        const uint32_t e = LT[s * N + i];
        const float4 a = PF[e & F.jmask], b = VF[e >> F.ishift];
        const float dx = (a.x + b.x) - pf.x, dy = (a.y + b.y) - pf.y, dz = (a.z + b.z) - pf.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (r2 < rscreen) mask |= 1u << s;
    }
    // ---- phase 2: my bonds: pair energy and the 11 running sums of the tensor form
    double txx = 0.0, tyy = 0.0, tzz = 0.0, txy = 0.0, txz = 0.0, tyz = 0.0, vx = 0.0, vy = 0.0, vz = 0.0, sg = 0.0, sg2 = 0.0;
    double pair = 0.0;
#pragma unroll 1
    while (mask) {
        const int s = __ffs(mask) - 1; mask &= mask - 1;
        const uint32_t e = LT[s * N + i];
        const int j = e & F.jmask, img = e >> F.ishift;
        const double tx = (P[j] + V[img]) - px, ty = (P[N + j] + V[IVC + img]) - py, tz = (P[2 * N + j] + V[2 * IVC + img]) - pz;
        const double r2 = dist2(tx, ty, tz);
        if (r2 < CK.rcc2) {                               // the exact test; beyond RCC every term of the bond is an exact 0.0
            double ir, isr;
            bond_radial(r2, ir, isr);
            const double e1 = exp_nc(CK.sig02 * isr);
            const double e_2 = e1 * e1, e_4 = e_2 * e_2;
            const double g = e_4 * e_2;
            const double s2 = CK.ss * ir * ir;
            pair += CK.aeps * (CK.bigb * (s2 * s2) - 1.0) * (e_4 * e1);
            const double ux = tx * ir, uy = ty * ir, uz = tz * ir;
            const double gx = g * ux, gy = g * uy, gz = g * uz;
            txx = fma(gx, ux, txx); tyy = fma(gy, uy, tyy); tzz = fma(gz, uz, tzz);
            txy = fma(gx, uy, txy); txz = fma(gx, uz, txz); tyz = fma(gy, uz, tyz);
            vx += gx; vy += gy; vz += gz; sg += g; sg2 = fma(g, g, sg2);
        }
    }
    const double tt = txx * txx + tyy * tyy + tzz * tzz + 2.0 * (txy * txy + txz * txz + tyz * tyz);
    const double vv = vx * vx + vy * vy + vz * vz;
    const double omc = 1.0 - CK.cos0;
    const double three = 0.5 * (tt - 2.0 * CK.cos0 * vv + CK.cos0 * CK.cos0 * sg * sg - omc * omc * sg2);
    return 0.5 * pair + CK.leps * three;                     // molint.F90:464 (half the pair term), :483
}

template <int NT>
__global__ void __launch_bounds__(E3_THREADS, 7) k_model_energy3(const __grid_constant__ DeviceState S, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = (NT > 0) ? NT : S.N;
    const E3Lay Y(N);
    const int upc = Y.upc();
    const int nunits = S.W * S.nlat;
    const int unit0 = blockIdx.x * upc;
    const int tid = threadIdx.x;
    const EntFmt F = ent_fmt(N);

    // ---- stage the CTA's units; the Verlet rows are transposed on the way in
    for (int u = 0; u < upc; ++u) {
        const int unit = unit0 + u;
        if (unit >= nunits) break;
        unsigned char* ub = smem + u * Y.unit();
        double* P = (double*)(ub + Y.oP());
        double* V = (double*)(ub + Y.oV());
        uint16_t* LT = (uint16_t*)(ub + Y.oL());
        const double* gp = S.pos + (size_t)unit * 3 * N;
        for (int t = tid; t < 3 * N; t += E3_THREADS) P[t] = gp[t];
        const double* gi = S.iv + (size_t)unit * 3 * IVC;
        for (int t = tid; t < 3 * IVC; t += E3_THREADS) V[t] = gi[t];
        float4* PF = (float4*)(ub + Y.oPF());
        float4* VF = (float4*)(ub + Y.oVF());
        const double ox = gp[0], oy = gp[N], oz = gp[2 * N];          // origin of the fp32 screen: the unit's first molecule
        for (int t = tid; t < N; t += E3_THREADS)
            PF[t] = make_float4((float)(gp[t] - ox), (float)(gp[N + t] - oy), (float)(gp[2 * N + t] - oz), 0.f);
        for (int t = tid; t < IVC; t += E3_THREADS) VF[t] = make_float4((float)gi[t], (float)gi[IVC + t], (float)gi[2 * IVC + t], 0.f);
        const uint4* gl = (const uint4*)(S.list + (size_t)unit * N * LC);
        for (int t = tid; t < N * LC / 8; t += E3_THREADS) {
            const uint4 v = gl[t];                       // entries 8*(t % (LC/8)) .. +7 of row t / (LC/8)
            const int row = t / (LC / 8), s0 = (t % (LC / 8)) * 8;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                LT[(s0 + 2 * k) * N + row] = (uint16_t)(w[k] & 0xffffu);
                LT[(s0 + 2 * k + 1) * N + row] = (uint16_t)(w[k] >> 16);
            }
        }
    }
    __syncthreads();

    const int u = tid / N, i = tid - u * N;              // my unit (within the CTA) and molecule
    const int unit = unit0 + u;
    const bool live = (u < upc) && (unit < nunits);
    unsigned char* ub = smem + (live ? u : 0) * Y.unit();
    const double* P = (const double*)(ub + Y.oP());
    const double* V = (const double*)(ub + Y.oV());
    const uint16_t* LT = (const uint16_t*)(ub + Y.oL());
    if (live)
        ((double*)(ub + Y.oE()))[i] = molecule_energy(P, V, LT, (const float4*)(ub + Y.oPF()), (const float4*)(ub + Y.oVF()),
                                                      N, i, S.nn[(size_t)unit * N + i], F);
    __syncthreads();
    if (tid < upc && unit0 + tid < nunits) {
        const double* E = (const double*)(smem + tid * Y.unit() + Y.oE());
        double acc = 0.0;
        for (int k = 0; k < N; ++k) acc += E[k];
        const int un = unit0 + tid;
        S.scal[un / S.nlat].E[un % S.nlat] = acc;
        if (out) out[un] = acc;
    }
}


// ================================================================================================================
// k_model_energy4 -- k_model_energy3 as a PERSISTENT kernel fed by the TMA: the loads of the NEXT pair of units are in
// flight while the current pair is computed, so the 5 KB per unit cost no LSU instructions and no exposed HBM latency
// (in k_model_energy3 a third of the stall samples sit on the staging loads of a freshly started CTA).
//   * positions, image vectors and row lengths land in one of TWO buffers (`cp.async.bulk` = 1-D TMA, completion on
//     an mbarrier with expect_tx); the pair after next is requested as soon as a buffer is free;
//   * the Verlet rows land in ONE buffer: they are only read by the transposition at the start of an iteration, so
//     the next pair's rows are requested right after it (fence.proxy.async orders the generic reads before the
//     asynchronous overwrite);
//   * a wait that does not complete traps instead of hanging the GPU.
// Needs N % 16 == 0 (bulk copies move multiples of 16 bytes): 48 in every deck.
// ================================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 24)) __trap();                  // ~seconds: a lost transaction must not hang the device
    }
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct E4Lay {                    // byte offsets inside the CTA's shared memory (two units per CTA iteration)
    int N;
    __host__ __device__ explicit E4Lay(int n) : N(n) {}
    __host__ __device__ int pv_unit() const { return 24 * N + 24 * IVC + N; }       // P | V | nn of one unit (16-byte multiples)
    __host__ __device__ int l_unit()  const { return 2 * LC * N; }
    __host__ __device__ int prep_unit() const { return 2 * LC * N + 16 * N + 16 * IVC + 8 * N; }   // LT | PF | VF | E
    __host__ __device__ int upc()   const { return E3_THREADS / N; }
    __host__ __device__ int oBAR()  const { return 0; }                             // 3 mbarriers
    __host__ __device__ int oPV(int b) const { return 32 + b * upc() * pv_unit(); }
    __host__ __device__ int oL()    const { return 32 + 2 * upc() * pv_unit(); }
    __host__ __device__ int oPREP() const { return oL() + upc() * l_unit(); }
    __host__ __device__ int bytes() const { return oPREP() + upc() * prep_unit(); }
};

template <int NT>
__global__ void __launch_bounds__(E3_THREADS, 7) k_model_energy4(const __grid_constant__ DeviceState S, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = (NT > 0) ? NT : S.N;
    const E4Lay Y(N);
    const int upc = Y.upc();
    const int nunits = S.W * S.nlat;
    const int npairs = (nunits + upc - 1) / upc;
    const int tid = threadIdx.x;
    const EntFmt F = ent_fmt(N);
    uint64_t* bar = (uint64_t*)(smem + Y.oBAR());        // [0], [1]: positions / image vectors / row lengths; [2]: rows

    // one thread requests the data of a pair: every existing unit of it, bytes announced first
    auto request_pv = [&](int pair, int b) {
        const int u0 = pair * upc, nu = min(upc, nunits - u0);
        mbar_expect_tx(bar + b, (uint32_t)(nu * Y.pv_unit()));
        for (int u = 0; u < nu; ++u) {
            unsigned char* dst = smem + Y.oPV(b) + u * Y.pv_unit();
            tma_load_1d(dst, S.pos + (size_t)(u0 + u) * 3 * N, 24 * N, bar + b);
            tma_load_1d(dst + 24 * N, S.iv + (size_t)(u0 + u) * 3 * IVC, 24 * IVC, bar + b);
            tma_load_1d(dst + 24 * N + 24 * IVC, S.nn + (size_t)(u0 + u) * N, N, bar + b);
        }
    };
    auto request_l = [&](int pair) {
        const int u0 = pair * upc, nu = min(upc, nunits - u0);
        mbar_expect_tx(bar + 2, (uint32_t)(nu * Y.l_unit()));
        for (int u = 0; u < nu; ++u)
            tma_load_1d(smem + Y.oL() + u * Y.l_unit(), S.list + (size_t)(u0 + u) * N * LC, Y.l_unit(), bar + 2);
    };

    if (tid == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < npairs) { request_pv(blockIdx.x, 0); request_l(blockIdx.x); }

    const int u = tid / N, i = tid - u * N;              // my unit within the pair and my molecule
    int it = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++it) {
        const int b = it & 1;
        const int next = pair + gridDim.x;
        // the other position buffer was last read two iterations ago (barrier at the end of every iteration)
        if (tid == 0 && next < npairs) request_pv(next, b ^ 1);
        mbar_wait(bar + b, (uint32_t)(it >> 1) & 1u);
        mbar_wait(bar + 2, (uint32_t)it & 1u);
        const int unit = pair * upc + u;
        const bool live = (u < upc) && (unit < nunits);
        // ---- prepare: transpose the rows, fp32 copies for the screen
        for (int uu = 0; uu < upc; ++uu) {
            if (pair * upc + uu >= nunits) break;
            const unsigned char* pv = smem + Y.oPV(b) + uu * Y.pv_unit();
            unsigned char* pr = smem + Y.oPREP() + uu * Y.prep_unit();
            const double* P = (const double*)pv;
            const double* V = (const double*)(pv + 24 * N);
            uint16_t* LT = (uint16_t*)pr;
            float4* PF = (float4*)(pr + 2 * LC * N);
            float4* VF = PF + N;
            const double ox = P[0], oy = P[N], oz = P[2 * N];
            for (int t = tid; t < N; t += E3_THREADS)
                PF[t] = make_float4((float)(P[t] - ox), (float)(P[N + t] - oy), (float)(P[2 * N + t] - oz), 0.f);
            for (int t = tid; t < IVC; t += E3_THREADS) VF[t] = make_float4((float)V[t], (float)V[IVC + t], (float)V[2 * IVC + t], 0.f);
            const uint4* sl = (const uint4*)(smem + Y.oL() + uu * Y.l_unit());
            for (int t = tid; t < N * LC / 8; t += E3_THREADS) {
                const uint4 v = sl[t];
                const int row = t / (LC / 8), s0 = (t % (LC / 8)) * 8;
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    LT[(s0 + 2 * k) * N + row] = (uint16_t)(w[k] & 0xffffu);
                    LT[(s0 + 2 * k + 1) * N + row] = (uint16_t)(w[k] >> 16);
                }
            }
        }
        __syncthreads();
        // the row buffer is free again: request the next pair's rows (generic reads ordered before the async write)
        if (tid == 0 && next < npairs) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            request_l(next);
        }
        // ---- compute
        if (live) {
            const unsigned char* pv = smem + Y.oPV(b) + u * Y.pv_unit();
            unsigned char* pr = smem + Y.oPREP() + u * Y.prep_unit();
            const double* P = (const double*)pv;
            const double* V = (const double*)(pv + 24 * N);
            const uint8_t* NN = (const uint8_t*)(pv + 24 * N + 24 * IVC);
            const float4* PF = (const float4*)(pr + 2 * LC * N);
            double* E = (double*)(pr + 2 * LC * N + 16 * N + 16 * IVC);
            E[i] = molecule_energy(P, V, (const uint16_t*)pr, PF, PF + N, N, i, NN[i], F);
        }
        __syncthreads();
        if (tid < upc && pair * upc + tid < nunits) {
            const double* E = (const double*)(smem + Y.oPREP() + tid * Y.prep_unit() + 2 * LC * N + 16 * N + 16 * IVC);
            double acc = 0.0;
            for (int k = 0; k < N; ++k) acc += E[k];
            const int un = pair * upc + tid;
            S.scal[un / S.nlat].E[un % S.nlat] = acc;
            if (out) out[un] = acc;
        }
        __syncthreads();
        // the position buffer b is rewritten by the TMA in the NEXT iteration (request_pv(.., b)): order this
        // iteration's generic reads before it
        if (tid == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

}  // namespace v2
}  // namespace mw
