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

}  // namespace v2
}  // namespace mw
