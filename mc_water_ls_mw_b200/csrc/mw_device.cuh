// mw_device.cuh -- warp-cooperative device routines for the mW hot path (sm_100a).
//
// One warp owns one walker box.  A walker's positions, packed Verlet lists,
// image vectors, cell matrices, "bond masks", scalar state and random-number
// buffer live in shared memory; every routine below is warp-synchronous (all
// 32 lanes call it).
//
// Two classes of arithmetic are kept strictly apart (DESIGN.md "Parity"):
//   * STATE arithmetic (positions, cell, fractional transforms, image vectors,
//     neighbour tests) uses the x*() helpers = explicit round-to-nearest
//     mul/add/sub/div/sqrt intrinsics that the compiler never contracts into
//     FMAs, in exactly the reference's operation order, so that it is
//     bit-identical to the oracle (and to an uncontracted build of the reference).
//   * ENERGY arithmetic is free-form fp64 (FMA, MUFU-seeded Newton reciprocals,
//     pairwise shuffle-tree summation); parity tolerance 1e-11 relative.
//
// Reference lines cited as file:line are into keb721/mc_water_ls_mw.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mw {

// ---------------------------------------------------------------- constants
// constants.f90:23-24,39,43 ; molint.F90:64-74,255,516
constexpr double PI          = 3.141592653589793238462643383279502884197;
constexpr double INV_PI      = 1.0 / 3.141592653589793238462643383279502884197;
constexpr double KB          = 1.0 / 3.1577465e5;
constexpr double ANG_TO_BOHR = 1.0 / 0.5291772108;
constexpr double SIGMA   = 2.3925 * ANG_TO_BOHR;
constexpr double EPSILON = 6.189 / 627.509469;
constexpr double LAMBDA  = 23.15;
constexpr double BIGA    = 7.049556277;
constexpr double BIGB    = 0.6022245584;
constexpr double GAMMA   = 1.2;
constexpr double SW_A    = 1.8;
constexpr double COS0    = (double)(-0.33331324756f);   // single-precision literal, molint.F90:74
constexpr double RC      = SIGMA * SW_A;
constexpr double RCSQ    = SIGMA * SW_A * SIGMA * SW_A;
constexpr double RN      = SW_A * SIGMA * 1.18;
constexpr double RN2     = RN * RN;
constexpr double GS      = GAMMA * SIGMA;
constexpr double AEPS    = BIGA * EPSILON;
constexpr double LEPS    = LAMBDA * EPSILON;
constexpr double SS      = SIGMA * SIGMA;
constexpr double F_HUGE  = 1.7976931348623157e308;
// Two radii just inside the cut-off, used by the second-generation walker kernel (mw2.cuh):
//  * RCC: exp(0.2*sigma/(r - a*sigma)) < exp(-700) beyond it, so every energy term of such a bond (its 5th / 6th
//    power) is an exact 0.0 in fp64 -- for the oracle's exp() as well; evaluating only r < RCC changes no bit and
//    frees the exponential from its underflow clamp;
//  * RSIG: beyond it the three-body radial factor g = exp(gamma*sigma/(r - a*sigma)) is below 1e-17, and a j-centred
//    triplet built on that bond is below 4e-18 Ha -- under the rounding of the sums it would be added to (local
//    energies are ~4e-2 Ha).  Such bonds are not enumerated as triplet legs; pair energies always use the full range.
constexpr double RCC     = RC - 0.2 * SIGMA / 700.0;
constexpr double RSIG    = RC - GS / 39.14394658089878;        // ln(1e17)

// ---------------------------------------------------------------- capacities
#ifndef MW_LC
#define MW_LC 32
#endif
#ifndef MW_QC
#define MW_QC 64
#endif
#ifndef MW_RB
#define MW_RB 64
#endif
#ifndef MW_KC
#define MW_KC 256
#endif
#ifndef MW_CAND_ILP
#define MW_CAND_ILP 1      // neighbour-bond candidates per lane and iteration in the j-centred triplet stage (2: measured 11 % slower)
#endif
constexpr int LC  = MW_LC;    // list slots per molecule held in shared memory (one lane per slot)
constexpr int IVC = 32;    // image vectors per lattice (27 in every BASELINE config)
constexpr int QC  = MW_QC;    // bond records per batch (a trial move has ~26; more than QC in-range bonds -> ERR_BOND_OVERFLOW)
constexpr int CC  = 64;    // triplet centres per trial move: 2 lattices x LC slots
constexpr int RB  = MW_RB;    // random numbers buffered per refill
constexpr int KC  = MW_KC;   // (centre, bond) candidates of the j-centred triplets of one trial move (~90)
constexpr int NMAX = 1024; // molecules (10 bits of a packed list entry)
constexpr unsigned FULL = 0xffffffffu;
constexpr uint16_t NONE16 = 0xffffu;

// error bits reported per walker
enum : int {
    ERR_LIST_OVERFLOW  = 1,    // a molecule has more than LC list neighbours
    ERR_IVECT_OVERFLOW = 2,    // more than IVC image vectors (cell shrank below the cut-off)
    ERR_BOND_OVERFLOW  = 4,    // more than QC bonds inside the cut-off in one trial move (unphysical density)
    ERR_ITEM_OVERFLOW  = 8,    // more than KC neighbour bonds around one trial move (unphysical density)
    ERR_SELF_IMAGE     = 16,   // a molecule is its own list neighbour (cell narrower than 1.18*a*sigma)
    ERR_RNG_UNDERRUN   = 32,   // host FIFO ran dry
    ERR_WINDOW         = 64,   // dd: walker not in its window at eq_mc_cycles (mc_moves.F90:191-201)
    ERR_PROB           = 128,  // cumulative move probability error (mc_moves.F90:174)
};

// ---------------------------------------------------------------- exact (state) arithmetic
__device__ __forceinline__ double xm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xa(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xs(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lowbits(int n) { return (n >= 32) ? 0xffffffffu : ((1u << n) - 1u); }
__device__ __forceinline__ unsigned lt_mask() { return (1u << (threadIdx.x & 31)) - 1u; }

// Packed Verlet-list entry (uint16).  Boxes of up to 64 molecules (every reference deck has 48) carry the
// REVERSE SLOT of the entry as well: entry s of row i = (j, image) has rev = the slot of row j that holds
// (i, inverse image) -- the neighbour test is symmetric (molint.F90:537 on exact negatives), so it always exists.
//     N <= 64 :  j | rev << 6 | image << 11          N > 64 :  j | image << 10
struct EntFmt { uint32_t jmask; int ishift; };
__host__ __device__ __forceinline__ EntFmt ent_fmt(int N) { return (N <= 64) ? EntFmt{63u, 11} : EntFmt{1023u, 10}; }
__host__ __device__ __forceinline__ bool ent_has_rev(int N) { return N <= 64; }
__device__ __forceinline__ double dist2(double x, double y, double z) { return fma(z, z, fma(y, y, x * x)); }

// Fortran m(i,j), 1-based, column-major 3x3
#define MW_H(m, i, j) ((m)[((j) - 1) * 3 + ((i) - 1)])

// util.f90:16-41
__device__ __forceinline__ double determinant3(const double* m)
{
    double det = xm(MW_H(m,1,1), xs(xm(MW_H(m,2,2), MW_H(m,3,3)), xm(MW_H(m,2,3), MW_H(m,3,2))));
    det = xs(det, xm(MW_H(m,1,2), xs(xm(MW_H(m,2,1), MW_H(m,3,3)), xm(MW_H(m,2,3), MW_H(m,3,1)))));
    det = xa(det, xm(MW_H(m,1,3), xs(xm(MW_H(m,2,1), MW_H(m,3,2)), xm(MW_H(m,2,2), MW_H(m,3,1)))));
    return det;
}

// util.f90:43-77 (every lane computes the same 9 numbers)
__device__ __forceinline__ void recipmatrix3(const double* h, double* r)
{
    MW_H(r,1,1) = xs(xm(MW_H(h,2,2), MW_H(h,3,3)), xm(MW_H(h,2,3), MW_H(h,3,2)));
    MW_H(r,1,2) = xs(xm(MW_H(h,2,3), MW_H(h,3,1)), xm(MW_H(h,2,1), MW_H(h,3,3)));
    MW_H(r,1,3) = xs(xm(MW_H(h,2,1), MW_H(h,3,2)), xm(MW_H(h,2,2), MW_H(h,3,1)));
    MW_H(r,2,1) = xs(xm(MW_H(h,1,3), MW_H(h,3,2)), xm(MW_H(h,1,2), MW_H(h,3,3)));
    MW_H(r,2,2) = xs(xm(MW_H(h,1,1), MW_H(h,3,3)), xm(MW_H(h,1,3), MW_H(h,3,1)));
    MW_H(r,2,3) = xs(xm(MW_H(h,1,2), MW_H(h,3,1)), xm(MW_H(h,1,1), MW_H(h,3,2)));
    MW_H(r,3,1) = xs(xm(MW_H(h,1,2), MW_H(h,2,3)), xm(MW_H(h,1,3), MW_H(h,2,2)));
    MW_H(r,3,2) = xs(xm(MW_H(h,1,3), MW_H(h,2,1)), xm(MW_H(h,1,1), MW_H(h,2,3)));
    MW_H(r,3,3) = xs(xm(MW_H(h,1,1), MW_H(h,2,2)), xm(MW_H(h,1,2), MW_H(h,2,1)));
    const double vol = xa(xa(xm(MW_H(h,1,1), MW_H(r,1,1)), xm(MW_H(h,1,2), MW_H(r,1,2))), xm(MW_H(h,1,3), MW_H(r,1,3)));
#pragma unroll 1
    for (int k = 0; k < 9; ++k) r[k] = xd(xm(xm(r[k], 2.0), PI), vol);
}

// 64-bit literals cost two UMOV instructions each at every use (10 % of the walker kernel's
// instruction stream before this table existed); operands taken from the constant bank cost nothing.
struct EnergyConsts {
    double log2e, magic, ln2hi, ln2lo, xmin;
    double e2, e3, e4, e5, e6, e7, e8, e9, e10, e11, e12, e13;   // 1/k!
    double q3125, q375;
    double rc, rc2, rcsq, sig02, ss, bigb, aeps, leps, gs, cos0, c099;
    double rcc2, rsig2;
};
__constant__ EnergyConsts CK = {
    1.4426950408889634, 6755399441055744.0, 6.93147180369123816490e-01, 1.90821492927058770002e-10, -708.0,
    0.5, 1.6666666666666666e-01, 4.1666666666666664e-02, 8.333333333333333e-03, 1.388888888888889e-03,
    1.984126984126984e-04, 2.48015873015873e-05, 2.755731922398589e-06, 2.755731922398589e-07,
    2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10,
    0.3125, 0.375,
    RC, RC * RC, RCSQ, 0.2 * SIGMA, SS, BIGB, AEPS, LEPS, GS, COS0, 0.99,
    RCC * RCC, RSIG * RSIG,
};

// ---------------------------------------------------------------- fast fp64 math for the ENERGY arithmetic
// The library routines (division, rsqrt, exp) carry special-case handling that costs 40-60 SASS
// instructions each; the energy terms only see normal, positive, moderate arguments, so a MUFU
// seed + Newton steps / a plain range reduction is enough.  Accuracy ~2e-16 relative.
__device__ __forceinline__ double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));       // MUFU.RCP64H seed (20 mantissa bits of x), e ~ 2^-20
    const double e = fma(-x, y, 1.0);                            // 1/x = y/(1 - e) = y (1 + e)(1 + e^2) + O(e^4)
    const double a = fma(y, e, y);                               // a and e*e are independent: 3 deep after the seed
    return fma(a, e * e, a);
}

__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));     // MUFU.RSQ64H seed (20 mantissa bits of x), e = 1 - x y^2 ~ 2^-19
    const double e = fma(-(x * y), y, 1.0);                      // x^-1/2 = y (1 - e)^-1/2 = y (1 + e/2 + 3e^2/8 + 5e^3/16 + O(e^4))
    const double p = fma(fma(CK.q3125, e, CK.q375), e, 0.5);         // one quartic step (error 35/128 e^4); p and y*e are independent
    return fma(y * e, p, y);
}

// 1/r and 1/(r - a*sigma) = (r + a*sigma)/(r^2 - (a*sigma)^2) of a squared bond length: written so that the
// reciprocal does not wait for the square root (two independent MUFU + Newton chains)
__device__ __forceinline__ void bond_radial(double r2, double& ir, double& isr)
{
    ir = rsqrt_fast(r2);
    const double id = rcp_fast(r2 - CK.rc2);
    isr = fma(r2, ir, CK.rc) * id;
}

// exp(x) for x <= ~700; returns 0 below -708 (the SW terms vanish at the cut-off: x -> -inf).
// Degree-13 Taylor polynomial on |r| <= ln2/2 in Estrin form (dependency depth 4 after r instead of 13:
// the walker kernel is bound by latency and instruction supply, not by fp64 throughput).
__device__ __forceinline__ double exp_fast(double x)
{
    const double xc = fmax(x, CK.xmin);
    const double t = fma(xc, CK.log2e, CK.magic);                          // round(x*log2 e) in the low word
    const int n = __double2loint(t);
    const double fn = t - CK.magic;
    double r = fma(-fn, CK.ln2hi, xc);
    r = fma(-fn, CK.ln2lo, r);                                             // |r| <= ln2/2
    const double r2 = r * r;
    const double a0 = 1.0 + r;                                             // 1/0! + r/1!
    const double a1 = fma(CK.e3, r, CK.e2);                                // 1/2! + r/3!
    const double a2 = fma(CK.e5, r, CK.e4);
    const double a3 = fma(CK.e7, r, CK.e6);
    const double a4 = fma(CK.e9, r, CK.e8);
    const double a5 = fma(CK.e11, r, CK.e10);
    const double a6 = fma(CK.e13, r, CK.e12);
    const double r4 = r2 * r2;
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(a5, r2, a4);
    const double r8 = r4 * r4;
    const double c0 = fma(b1, r4, b0);
    const double c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    const double s = __hiloint2double((n + 1023) << 20, 0);               // 2^n, n in [-1022, 1023]
    return (x < CK.xmin) ? 0.0 : p * s;
}

// the same for -708 <= x (no underflow clamp): arguments of bonds inside RCC
__device__ __forceinline__ double exp_nc(double x)
{
    const double t = fma(x, CK.log2e, CK.magic);
    const int n = __double2loint(t);
    const double fn = t - CK.magic;
    double r = fma(-fn, CK.ln2hi, x);
    r = fma(-fn, CK.ln2lo, r);
    const double r2 = r * r;
    const double a0 = 1.0 + r;
    const double a1 = fma(CK.e3, r, CK.e2);
    const double a2 = fma(CK.e5, r, CK.e4);
    const double a3 = fma(CK.e7, r, CK.e6);
    const double a4 = fma(CK.e9, r, CK.e8);
    const double a5 = fma(CK.e11, r, CK.e10);
    const double a6 = fma(CK.e13, r, CK.e12);
    const double r4 = r2 * r2;
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(a5, r2, a4);
    const double r8 = r4 * r4;
    const double c0 = fma(b1, r4, b0);
    const double c1 = fma(a6, r4, b2);
    const double p = fma(c1, r8, c0);
    return p * __hiloint2double((n + 1023) << 20, 0);
}

// one shared copy for the call sites outside the bond loops (acceptance, lattice switch): code size matters
// more than the call there
__device__ __noinline__ double exp_call(double x) { return exp_fast(x); }

// log(x) for normal positive x (no special cases), ~2e-16 relative: x = m*2^e, m in [sqrt(1/2), sqrt(2)),
// log m = 2 atanh((m-1)/(m+1))
__device__ __forceinline__ double log_fast(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                 // m in [1,2)
    if (hi >= 0x3ff6a09e) { hi -= 0x00100000; e += 1; }  // m >= sqrt(2) (high word): halve it
    const double m = __hiloint2double(hi, lo);
    const double s = (m - 1.0) * rcp_fast(m + 1.0);
    const double z = s * s;
    double p = 1.0 / 21.0;
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    p = p * z;
    const double lm = fma(2.0 * s, p, 2.0 * s);
    return fma((double)e, 6.93147180559945286e-01, lm);
}

// ---------------------------------------------------------------- Philox-4x32-10
// counter = (block_lo, block_hi, stream, 0), key = (seed_lo, seed_hi); two
// doubles per block (53 high bits of each 64-bit half) -- the stream documented in DESIGN.md "RNG".
__device__ __forceinline__ void philox_block(uint64_t seed, uint32_t stream, uint64_t block, double& o0, double& o1)
{
    uint32_t c0 = (uint32_t)block, c1 = (uint32_t)(block >> 32), c2 = stream, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const uint64_t a = ((uint64_t)c1 << 32) | c0;
    const uint64_t b = ((uint64_t)c3 << 32) | c2;
    o0 = (double)(a >> 11) * 0x1.0p-53;
    o1 = (double)(b >> 11) * 0x1.0p-53;
}

// ---------------------------------------------------------------- per-walker state
// Scalars that persist between launches (global memory) and live in shared memory inside a
// kernel.  Inside a kernel they are "uniform registers in shared memory": every lane stores the
// same value and reads back what it wrote itself, so no synchronisation is needed for them.
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
    int firstcycle;         // mc_moves.F90:85: wl_factor is still the original one
    int hist_reset;         // mc_moves.F90:1957: the one-off histogram reset of mc_check_flatness happened
};

// ---------------------------------------------------------------- shared-memory view of one walker
struct WalkerView {
    int N, nlat;
    double*   pos;     // [nlat][3][N]   SoA x|y|z
    double*   iv;      // [nlat][3][IVC]
    double*   cell;    // [nlat][9]      hmatrix, column-major
    double*   recip;   // [nlat][9]
    double*   q;       // [4][QC] scratch: tx,ty,tz,r2 -> ux,uy,uz,g
    double*   save;    // [36]    old cell + recip during a volume move
    double*   rngbuf;  // [RB]    buffered U[0,1) numbers
    double*   lv;      // [2]     log(V1/V2), log(V2/V1)
    double*   mv;      // [2][6]  trial position (x,y,z) and displacement (x,y,z) of the move in flight, per lattice
    WalkerScalars* sc;
    uint64_t* rngbase; // draw index of rngbuf[0]
    uint32_t* qmeta;   // [QC]   call | j<<8
    uint32_t* cmeta;   // [CC]   lat | j<<6
    uint32_t* bmask;   // [nlat][N] bit s: list slot s currently within the cut-off a*sigma
    int*      niv;     // [2]
    uint16_t* cq;      // [CC][2] bond record of the centre in the old / new variant
    uint16_t* cpre;    // [CC+2]  exclusive prefix of the per-centre bond counts
    uint16_t* cand;    // [KC]    centre<<5 | list slot of every bond of every centre
    uint16_t* list;    // [nlat][N][LC] packed entries img<<10 | j   (0-based)
    uint8_t*  nn;      // [nlat][N]
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
__host__ __device__ inline size_t smem_doubles(int N, int nlat)
{
    return (size_t)nlat * 3 * N + (size_t)nlat * 3 * IVC + (size_t)nlat * 18 + 4 * QC + 36 + RB + 2 + 12;
}

// Layout (every block 16-byte aligned): doubles | scalars | list | 32-bit words | 16-bit words | bytes
__host__ __device__ inline size_t walker_smem_bytes(int N, int nlat)
{
    size_t b = 0;
    b += align16(sizeof(double) * smem_doubles(N, nlat));
    b += align16(sizeof(WalkerScalars) + sizeof(uint64_t));
    b += align16(sizeof(uint16_t) * (size_t)nlat * N * LC);                       // list
    b += align16(sizeof(uint32_t) * (QC + CC + (size_t)nlat * N + 2));            // qmeta, cmeta, bmask, niv
    b += align16(sizeof(uint16_t) * (CC * 2 + CC + 2 + KC));                      // cq, cpre, cand
    b += align16((size_t)nlat * N);                                               // nn
    return b;
}

__device__ __forceinline__ WalkerView carve_walker(unsigned char* base, int N, int nlat)
{
    WalkerView w; w.N = N; w.nlat = nlat;
    unsigned char* p = base;
    w.pos    = (double*)p;
    w.iv     = w.pos + (size_t)nlat * 3 * N;
    w.cell   = w.iv + (size_t)nlat * 3 * IVC;
    w.recip  = w.cell + (size_t)nlat * 9;
    w.q      = w.recip + (size_t)nlat * 9;
    w.save   = w.q + 4 * QC;
    w.rngbuf = w.save + 36;
    w.lv     = w.rngbuf + RB;
    w.mv     = w.lv + 2;
    p += align16(sizeof(double) * smem_doubles(N, nlat));
    w.sc      = (WalkerScalars*)p;
    w.rngbase = (uint64_t*)(p + sizeof(WalkerScalars));
    p += align16(sizeof(WalkerScalars) + sizeof(uint64_t));
    w.list  = (uint16_t*)p;
    p += align16(sizeof(uint16_t) * (size_t)nlat * N * LC);
    w.qmeta = (uint32_t*)p;
    w.cmeta = w.qmeta + QC;
    w.bmask = w.cmeta + CC;
    w.niv   = (int*)(w.bmask + (size_t)nlat * N);
    p += align16(sizeof(uint32_t) * (QC + CC + (size_t)nlat * N + 2));
    w.cq    = (uint16_t*)p;
    w.cpre  = w.cq + CC * 2;
    w.cand  = w.cpre + CC + 2;
    p += align16(sizeof(uint16_t) * (CC * 2 + CC + 2 + KC));
    w.nn    = (uint8_t*)p;
    return w;
}

// ---------------------------------------------------------------- image vectors
// molint.F90:174-217.  Lane k builds vector k.  Returns error bits.
__device__ __noinline__ int compute_ivects_warp(unsigned char* smem, int N, int nlat, int lat)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    const double* h = w.cell + 9 * lat;
    const double l1 = xsqrt(xa(xa(xm(h[0], h[0]), xm(h[1], h[1])), xm(h[2], h[2])));
    const double l2 = xsqrt(xa(xa(xm(h[3], h[3]), xm(h[4], h[4])), xm(h[5], h[5])));
    const double l3 = xsqrt(xa(xa(xm(h[6], h[6]), xm(h[7], h[7])), xm(h[8], h[8])));
    const int im = (int)floor(xd(RC, l1)) + 1;
    const int jm = (int)floor(xd(RC, l2)) + 1;
    const int km = (int)floor(xd(RC, l3)) + 1;
    const int nj = 2 * jm + 1, nk = 2 * km + 1;
    const int nv = (2 * im + 1) * nj * nk;
    const int k = lane_id();
    if (k == 0) w.niv[lat] = nv;
    if (nv > IVC) { __syncwarp(); return ERR_IVECT_OVERFLOW; }
    const int f0 = (nv - 1) / 2;           // full-grid index of the (0,0,0) cell
    if (k < nv) {
        double vx = 0.0, vy = 0.0, vz = 0.0;
        if (k > 0) {
            const int f = (k <= f0) ? k - 1 : k;       // entry 1 (k=0) is the zero vector, skipped in the loop nest
            const int ic = f / (nj * nk) - im;
            const int jc = (f / nk) % nj - jm;
            const int kc = f % nk - km;
            const double a = (double)ic, b = (double)jc, c = (double)kc;
            vx = xa(xa(xm(a, h[0]), xm(b, h[3])), xm(c, h[6]));
            vy = xa(xa(xm(a, h[1]), xm(b, h[4])), xm(c, h[7]));
            vz = xa(xa(xm(a, h[2]), xm(b, h[5])), xm(c, h[8]));
        }
        double* V = w.iv + lat * 3 * IVC;
        V[k] = vx; V[IVC + k] = vy; V[2 * IVC + k] = vz;
    }
    __syncwarp();
    return 0;
}

// index of the image vector that is the negative of image `img`
__device__ __forceinline__ int inverse_image(int img, int nv)
{
    if (img == 0) return 0;
    const int f0 = (nv - 1) / 2;
    const int f = (img <= f0) ? img - 1 : img;
    const int g = nv - 1 - f;
    return (g < f0) ? g + 1 : g;
}

// ---------------------------------------------------------------- Verlet list
// molint.F90:501-559: brute force over (j, image k), emitted in j-ascending,
// k-ascending order.  Lanes own molecules j; every lane tests its candidate images and keeps a
// bit mask; entries are then emitted in lane order.
//
// Candidate images: with b_a the reciprocal vectors of the cell (b_a . h_b = delta_ab), the image
// displaced by n = (n1,n2,n3) cells has |t| >= |b_a.v + n_a| / |b_a| for every axis a, so only
// n_a with |b_a.v + n_a| < rn |b_a| (widened by 1e-9) can pass the exact test; the box of those
// n (typically 1-2 of the 27 images) is evaluated with the reference's exact arithmetic.
__host__ __device__ constexpr uint32_t image_axis_mask(int axis, int n)      // images k (0..26) whose cell offset along `axis` is n
{
    uint32_t m = 0;
    for (int k = 0; k < 27; ++k) {
        const int f = (k == 0) ? 13 : (k <= 13 ? k - 1 : k);
        const int c = (axis == 0) ? f / 9 - 1 : (axis == 1) ? (f / 3) % 3 - 1 : f % 3 - 1;
        if (c == n) m |= 1u << k;
    }
    return m;
}

__device__ __forceinline__ uint32_t axis_candidates(double s, double R, uint32_t mm, uint32_t m0, uint32_t mp)
{
    return (fabs(s - 1.0) < R ? mm : 0u) | (fabs(s) < R ? m0 : 0u) | (fabs(s + 1.0) < R ? mp : 0u);
}

__device__ __noinline__ int compute_neighbours_warp(unsigned char* smem, int N, int nlat, int lat)
{
    int err = compute_ivects_warp(smem, N, nlat, lat);          // molint.F90:518
    if (err) return err;
    const WalkerView w = carve_walker(smem, N, nlat);
    const int lane = lane_id();
    const int nv = w.niv[lat];
    const double* P = w.pos + lat * 3 * N;
    const double* V = w.iv + lat * 3 * IVC;
    // reciprocal vectors (plain arithmetic: they only select candidates)
    const double* h = w.cell + 9 * lat;
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
    const bool boxed = (nv == 27);                              // always, unless the image set is degenerate
    const EntFmt F = ent_fmt(N);
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
                    // the image k adds n = +cell offset; |b.v + n| < R  (n = -1, 0, +1)
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
            uint16_t* row = w.list + ((size_t)lat * N + i) * LC;
#pragma unroll 1
            while (m) {
                const int k = __ffs(m) - 1; m &= m - 1;
                if (off < LC) row[off] = (uint16_t)((k << F.ishift) | j);
                ++off;
            }
            total += __shfl_sync(FULL, incl, 31);
        }
        if (total > LC) { err |= ERR_LIST_OVERFLOW; total = LC; }
        if (lane == 0) w.nn[lat * N + i] = (uint8_t)total;
    }
    err = (int)__reduce_or_sync(FULL, (unsigned)err);
    __syncwarp();
    // reverse slots (boxes of up to 64 molecules): lanes = the slots of row i; a row is sorted by (j, image), so
    // the entry (i, inverse image) of row j is found by bisection.  After an overflow the rows are truncated and
    // the field is meaningless (the walker is flagged).
    if (ent_has_rev(N)) {
#pragma unroll 1
        for (int i = 0; i < N; ++i) {
            const int nni = w.nn[lat * N + i];
            if (lane < nni) {
                uint16_t* row = w.list + ((size_t)lat * N + i) * LC;
                const uint32_t e = row[lane];
                const int j = e & 63, img = e >> 11;
                const uint32_t key = ((uint32_t)inverse_image(img, nv) << 11) | (uint32_t)i;
                const uint16_t* rj = w.list + ((size_t)lat * N + j) * LC;
                // keys compare as (j, image) = low 6 bits major: order by (e & 63) << 5 | (e >> 11)
                const uint32_t want = ((key & 63u) << 5) | (key >> 11);
                int lo = 0, hi = (int)w.nn[lat * N + j] - 1;
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
    }
    return err;
}

// ---------------------------------------------------------------- bond masks
// bmask[lat][a] bit s <=> slot s of a's list is inside the cut-off (r^2 < rcsq,
// molint.F90:276/454).  Lanes are the slots.
__device__ __noinline__ void compute_bond_masks_warp(unsigned char* smem, int N, int nlat, int lat)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    const int lane = lane_id();
    const double* P = w.pos + lat * 3 * N;
    const double* V = w.iv + lat * 3 * IVC;
    const EntFmt F = ent_fmt(N);
    for (int a = 0; a < N; ++a) {
        const int nna = w.nn[lat * N + a];
        const bool has = lane < nna;
        const uint32_t e = has ? w.list[((size_t)lat * N + a) * LC + lane] : 0u;
        const int j = e & F.jmask, img = e >> F.ishift;
        const double tx = (P[j] + V[img]) - P[a];
        const double ty = (P[N + j] + V[IVC + img]) - P[N + a];
        const double tz = (P[2 * N + j] + V[2 * IVC + img]) - P[2 * N + a];
        const double r2 = dist2(tx, ty, tz);
        const uint32_t m = __ballot_sync(FULL, has && r2 < RCSQ);
        if (lane == 0) w.bmask[lat * N + a] = m;
    }
    __syncwarp();
}

// ---------------------------------------------------------------- energy kernels (free-form fp64)
// Pair record evaluation shared by the local and the full energy:
//   in : q[0..2][r] = separation vector t, q[3][r] = r^2
//   out: q[0..2][r] = unit vector u = t/r, q[3][r] = g = exp(gamma*sigma/(r - a*sigma))
//   returns the pair energy A*eps*(B*(sigma/r)^4 - 1)*exp(sigma/(r - a*sigma))   (molint.F90:278-297 / :456-461)
__device__ __forceinline__ double eval_bond(double* q, int r)
{
    const double tx = q[r], ty = q[QC + r], tz = q[2 * QC + r], r2 = q[3 * QC + r];
    double ir, isr;
    bond_radial(r2, ir, isr);
    // exp(sigma*isr) = e^5 and exp(gamma*sigma*isr) = e^6 with e = exp(0.2*sigma*isr)  (gamma = 1.2)
    const double e1 = exp_fast(CK.sig02 * isr);
    const double e_2 = e1 * e1, e_4 = e_2 * e_2;
    const double e2 = e_4 * e1;
    const double g = e_4 * e_2;
    const double s2 = CK.ss * ir * ir;
    q[r] = tx * ir; q[QC + r] = ty * ir; q[2 * QC + r] = tz * ir; q[3 * QC + r] = g;
    return CK.aeps * (CK.bigb * (s2 * s2) - 1.0) * e2;
}

// (cos(theta) - cos0)^2 with the reference's k==i filter (molint.F90:367-371)
__device__ __forceinline__ double hfun(double ct)
{
    const double d = ct - CK.cos0;
    return (ct < CK.c099) ? d * d : 0.0;
}

// Sum 4 per-lane accumulators over the warp (pairwise shuffle tree) and
// broadcast the 4 totals to every lane.
__device__ __forceinline__ void reduce4(double& a0, double& a1, double& a2, double& a3)
{
    const int lane = lane_id();
    {   // xor 16: lanes 0-15 keep (a0,a1), lanes 16-31 keep (a2,a3)
        const bool up = lane & 16;
        const double s0 = up ? a0 : a2, s1 = up ? a1 : a3;
        const double r0 = __shfl_xor_sync(FULL, s0, 16), r1 = __shfl_xor_sync(FULL, s1, 16);
        a0 = (up ? a2 : a0) + r0;
        a1 = (up ? a3 : a1) + r1;
    }
    {   // xor 8: within each half, lanes with bit 3 clear keep a0, the others keep a1
        const bool up = lane & 8;
        const double s = up ? a0 : a1;
        const double r = __shfl_xor_sync(FULL, s, 8);
        a0 = (up ? a1 : a0) + r;
    }
    a0 += __shfl_xor_sync(FULL, a0, 4);
    a0 += __shfl_xor_sync(FULL, a0, 2);
    a0 += __shfl_xor_sync(FULL, a0, 1);
    const double t0 = __shfl_sync(FULL, a0, 0), t1 = __shfl_sync(FULL, a0, 8);
    const double t2 = __shfl_sync(FULL, a0, 16), t3 = __shfl_sync(FULL, a0, 24);
    a0 = t0; a1 = t1; a2 = t2; a3 = t3;
}

__device__ __forceinline__ double warp_sum(double a)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(FULL, a, d);
    return a;
}

// position of the (rank+1)-th set bit of m (rank < popc(m))
__device__ __forceinline__ int nth_set_bit(uint32_t m, int rank)
{
    int pos = 0;
    int c = __popc(m & 0xffffu);
    if (rank >= c) { rank -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xffu);
    if (rank >= c) { rank -= c; pos += 8; m >>= 8; }
    c = __popc(m & 0xfu);
    if (rank >= c) { rank -= c; pos += 4; m >>= 4; }
    c = __popc(m & 0x3u);
    if (rank >= c) { rank -= c; pos += 2; m >>= 2; }
    if (rank >= (int)(m & 1u)) pos += 1;
    return pos;
}

// ------------------------------------------------------------------------------------------------
// Local energies of molecule imol in every lattice, for the current ("old")
// position and -- when WITH_NEW -- for a trial position pnew[lat] as well
// (4 evaluations of compute_local_real_energy, molint.F90:220-404, in one
// flattened pass; mc_moves.F90:1010,1083).
//
// Formulation (equal to the reference's sum up to fp64 rounding):
//   E(i) = sum_{b in bonds(i)} phi2(r_ib)
//        + lam*eps * sum_{b<c in bonds(i)} g_ib g_ic h(u_ib.u_ic) * (j_b==j_c ? 3 : 1)
//        + lam*eps * sum_{b in bonds(i)} g_ib sum_{k in bonds(j_b), k not an image of i} g_jk h(-u_ib.u_jk)
// where bonds(x) are the list entries inside the cut-off.  The factor 3 covers
// the two j-centred triplets whose third body is another periodic image of i
// (reference: list-B entries with kmol==imol that survive the cos<0.99 filter).
// Bonds of the neighbours j come from the cached bond masks; their geometry
// does not depend on the position of i, so one item evaluation serves both
// the old and the new variant.
//
// Outputs (uniform over the warp): eo[lat], en[lat]; mo[lat]/mn[lat] = in-range
// slot masks of imol for the old / new position.
template <int NLAT, bool WITH_NEW>
__device__ __forceinline__ void local_energies_warp(const WalkerView& w, int imol,
                                                    double* eo, double* en, uint32_t* mo, uint32_t* mn)
{
    const int N = w.N, lane = lane_id();
    const unsigned lt = lt_mask();
    double* q = w.q;
    int nq = 0, nc = 0;

    // ---- stage 1: distance tests over imol's own list (lanes = slots), compaction into bond records.
    // One copy of the code for both lattices (the loop is NOT unrolled: the kernel sits at the edge of
    // the instruction cache); the trial position comes from the walker's move record w.mv.
    uint32_t mo0 = 0, mo1 = 0, mn0 = 0, mn1 = 0;
    const EntFmt F = ent_fmt(N);
#pragma unroll 1
    for (int lat = 0; lat < NLAT; ++lat) {
        const double* P = w.pos + lat * 3 * N;
        const double* V = w.iv + lat * 3 * IVC;
        const int nni = w.nn[lat * N + imol];
        const bool has = lane < nni;
        const uint32_t e = has ? w.list[((size_t)lat * N + imol) * LC + lane] : 0u;
        const int j = e & F.jmask, img = e >> F.ishift;
        const double pjx = P[j] + V[img], pjy = P[N + j] + V[IVC + img], pjz = P[2 * N + j] + V[2 * IVC + img];
        const double tox = pjx - P[imol], toy = pjy - P[N + imol], toz = pjz - P[2 * N + imol];
        const double r2o = dist2(tox, toy, toz);
        const bool fo = has && r2o < CK.rcsq;
        const uint32_t bo = __ballot_sync(FULL, fo);
        if (lat == 0) mo0 = bo; else mo1 = bo;
        const int io = nq + __popc(bo & lt);
        nq += __popc(bo);
        bool fn = false; int in_ = 0; uint32_t bn = 0;
        double tnx = 0, tny = 0, tnz = 0, r2n = 0;
        if (WITH_NEW) {
            const double* pn = w.mv + lat * 6;
            tnx = pjx - pn[0]; tny = pjy - pn[1]; tnz = pjz - pn[2];
            r2n = dist2(tnx, tny, tnz);
            fn = has && r2n < CK.rcsq;
            bn = __ballot_sync(FULL, fn);
            if (lat == 0) mn0 = bn; else mn1 = bn;
            in_ = nq + __popc(bn & lt);
            nq += __popc(bn);
        }
        const uint32_t bu = bo | bn;
        const int ic = nc + __popc(bu & lt);
        nc += __popc(bu);
        // nc <= 2*LC == CC by construction; nq may exceed QC only at unphysical densities (flagged below)
        if (fo && io < QC) {
            q[io] = tox; q[QC + io] = toy; q[2 * QC + io] = toz; q[3 * QC + io] = r2o;
            w.qmeta[io] = (uint32_t)(lat * 2) | ((uint32_t)__popc(bo) << 2) | ((uint32_t)j << 8) | ((uint32_t)__popc(bo & lt) << 18);
        }
        if (WITH_NEW && fn && in_ < QC) {
            q[in_] = tnx; q[QC + in_] = tny; q[2 * QC + in_] = tnz; q[3 * QC + in_] = r2n;
            w.qmeta[in_] = (uint32_t)(lat * 2 + 1) | ((uint32_t)__popc(bn) << 2) | ((uint32_t)j << 8) | ((uint32_t)__popc(bn & lt) << 18);
        }
        if (fo || fn) {
            w.cmeta[ic] = (uint32_t)lat | ((uint32_t)j << 6);
            w.cq[ic * 2] = (fo && io < QC) ? (uint16_t)io : NONE16;
            w.cq[ic * 2 + 1] = (fn && in_ < QC) ? (uint16_t)in_ : NONE16;
        }
    }
    mo[0] = mo0; mn[0] = mn0;
    if (NLAT == 2) { mo[1] = mo1; mn[1] = mn1; }
    if (nq > QC) {                                  // results of this call are invalid; the walker is flagged
        w.sc->error |= ERR_BOND_OVERFLOW;
        nq = QC;
    }
    __syncwarp();

    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;     // per-lane partial sums of the 4 evaluations

    // ---- stage 2: bond evaluation (pair energy, g, unit vector)
    for (int b = 0; b < nq; b += 32) {
        const int r = b + lane;
        if (r < nq) {
            const double pe = eval_bond(q, r);
            const int c = w.qmeta[r] & 3;
            if (c == 0) a0 += pe;
            if (c == 1) a1 += pe;
            if (c == 2) a2 += pe;
            if (c == 3) a3 += pe;
        }
    }
    __syncwarp();

    // ---- stage 3: triplets centred on imol: all unordered pairs of bond records of one evaluation.
    // Lanes = records; a record at position `pos` of its segment of n records pairs with the records
    // (pos + d) mod n, d = 1 .. n/2: every pair exactly once (for even n the step d = n/2 meets each
    // pair from both ends, so only the lower half of the segment takes it), all lanes busy in every
    // step and half as many steps as a "later records" loop.
    const int nq3 = (w.sc->error & ERR_BOND_OVERFLOW) ? 0 : nq;     // overflowed records are incomplete: flagged, skipped
#pragma unroll 1
    for (int b0 = 0; b0 < nq3; b0 += 32) {
        const int r = b0 + lane;
        const bool act = r < nq3;
        const uint32_t qm = act ? w.qmeta[r] : 0u;                // ev | n << 2 | j << 8 | pos << 18
        const int ev = qm & 3, n = (qm >> 2) & 63, pos = (qm >> 18) & 63;
        const int half = n >> 1, send = r - pos + n;
        const bool even = !(n & 1);
        const double ux = act ? q[r] : 0.0, uy = act ? q[QC + r] : 0.0, uz = act ? q[2 * QC + r] : 0.0;
        const double g = act ? q[3 * QC + r] : 0.0;
        const int maxd = __reduce_max_sync(FULL, half);
        double tb = 0.0;
#pragma unroll 1
        for (int d = 1; d <= maxd; ++d) {
            int c = r + d;
            c = (c >= send) ? c - n : c;
            const bool on = (d <= half) && !(even && d == half && pos >= half);
            c = on ? c : r;
            const double ct = ux * q[c] + uy * q[QC + c] + uz * q[2 * QC + c];
            const double mult = (((qm ^ w.qmeta[c]) & 0x3ff00u) == 0u) ? 3.0 : 1.0;   // images of one molecule
            const double dd = ct - CK.cos0;
            if (on && ct < CK.c099) tb += q[3 * QC + c] * (dd * dd) * mult;      // the k==i filter as a predicate
        }
        tb *= CK.leps * g;
        if (ev == 0) a0 += tb;
        if (ev == 1) a1 += tb;
        if (ev == 2) a2 += tb;
        if (ev == 3) a3 += tb;
    }

    // ---- stages 4+5: j-centred triplets.  Centres are taken in groups (all of them, or 8 at a time
    // when their bonds would not fit the candidate table); per group: bond counts (lanes = centres),
    // exclusive prefix, expansion of every bond into the table as centre<<5 | list slot; then
    // lanes = candidates.  Images of imol are skipped (see above).  One evaluation of (j,k) serves
    // both variants.
#pragma unroll 1
    for (int cb = 0; cb < nc;) {
        const int c0 = cb + lane;
        int cnt = 0;
        uint32_t bm0 = 0;
        if (c0 < nc) {
            const uint32_t cm = w.cmeta[c0];
            bm0 = w.bmask[(cm & 1) * N + ((cm >> 6) & 1023)];
            cnt = __popc(bm0);
        }
        int take = 32;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (__shfl_sync(FULL, incl, 31) > KC) {        // unusually crowded: 8 centres (<= 256 bonds) at a time
            take = 8;
            if (lane >= 8) { cnt = 0; bm0 = 0; }
            incl = cnt;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            const int tot8 = __shfl_sync(FULL, incl, 7);
            incl = (lane < 8) ? incl : tot8;
        }
        const int ncand = __shfl_sync(FULL, incl, 31);
        {
            int pos = incl - cnt;
#pragma unroll 1
            while (bm0) {
                const int s2 = __ffs(bm0) - 1; bm0 &= bm0 - 1;
                w.cand[pos++] = (uint16_t)(((c0 - cb) << 5) | s2);
            }
        }
        __syncwarp();
        // Branch-free body: invalid lanes work on clamped indices and are discarded by selects (no
        // BSSY/BSYNC regions inside the loop).  More than one candidate per lane and iteration
        // (MW_CAND_ILP) executes empty tail passes and was measured slower: kernel time follows the
        // number of executed warp-instructions, not the length of the dependency chains.
        for (int tb0 = 0; tb0 < ncand; tb0 += 32 * MW_CAND_ILP) {
#pragma unroll
            for (int h = 0; h < MW_CAND_ILP; ++h) {
                const int t = tb0 + 32 * h + lane;
                const bool in = t < ncand;
                const uint32_t ce = w.cand[in ? t : 0];
                const int c = cb + (int)(ce >> 5), s2 = ce & 31;
                const uint32_t cm = w.cmeta[c];
                const int lat = cm & 1, j = (cm >> 6) & 1023;
                const uint32_t e2 = w.list[((size_t)lat * N + j) * LC + s2];
                const int k = e2 & F.jmask, img = e2 >> F.ishift;
                const double* P = w.pos + lat * 3 * N;
                const double* V = w.iv + lat * 3 * IVC;
                const double tx = (P[k] + V[img]) - P[j];
                const double ty = (P[N + k] + V[IVC + img]) - P[N + j];
                const double tz = (P[2 * N + k] + V[2 * IVC + img]) - P[2 * N + j];
                const double sq0 = dist2(tx, ty, tz);
                const bool ok = in && (k != imol) && (sq0 < CK.rcsq);
                const double sq = ok ? sq0 : CK.ss;                    // any length inside the cut-off
                double vi, isr;
                bond_radial(sq, vi, isr);
                const double ex = CK.leps * exp_fast(CK.gs * isr);
                const double ux = tx * vi, uy = ty * vi, uz = tz * vi;
                const uint32_t cq2 = *(const uint32_t*)(w.cq + c * 2);
                const uint32_t qo = cq2 & 0xffffu, qn = cq2 >> 16;
                const bool ho = ok && (qo != NONE16), hn = WITH_NEW && ok && (qn != NONE16);
                const int io = (qo != NONE16) ? (int)qo : 0, in_ = (qn != NONE16) ? (int)qn : 0;
                // validity, the cos < 0.99 filter (molint.F90:367-371) and the lattice all end up in the
                // predicates of the four accumulating adds
                const double cto = -(q[io] * ux + q[QC + io] * uy + q[2 * QC + io] * uz);
                const double d_o = cto - CK.cos0;
                const double vo = q[3 * QC + io] * ex * (d_o * d_o);
                const bool po = ho && (cto < CK.c099);
                if (po && lat == 0) a0 += vo;
                if (po && lat != 0) a2 += vo;
                if (WITH_NEW) {
                    const double ctn = -(q[in_] * ux + q[QC + in_] * uy + q[2 * QC + in_] * uz);
                    const double dn = ctn - CK.cos0;
                    const double vn = q[3 * QC + in_] * ex * (dn * dn);
                    const bool pn = hn && (ctn < CK.c099);
                    if (pn && lat == 0) a1 += vn;
                    if (pn && lat != 0) a3 += vn;
                }
            }
        }
        __syncwarp();
        cb += take;
    }

    reduce4(a0, a1, a2, a3);
    eo[0] = a0; en[0] = a1;
    if (NLAT == 2) { eo[1] = a2; en[1] = a3; }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Full energy of one lattice (compute_model_energy, molint.F90:407-499), molecule-
// chunked: bonds of a chunk of molecules are compacted (stage 1), evaluated
// (stage 2) and combined into i-centred triplets (stage 3).  Also refreshes the
// bond masks of the lattice.  Returns E (uniform).
__device__ __noinline__ double full_energy_warp(unsigned char* smem, int N, int nlat, int lat)
{
    const WalkerView w = carve_walker(smem, N, nlat);
    const int lane = lane_id();
    const unsigned lt = lt_mask();
    const double* P = w.pos + lat * 3 * N;
    const double* V = w.iv + lat * 3 * IVC;
    double* q = w.q;
    const EntFmt F = ent_fmt(N);
    double acc = 0.0;
    int a = 0;
    while (a < N) {
        // ---- gather bonds of molecules a.. while they fit in the record buffer
        int nq = 0;
        int a1 = a;
        for (; a1 < N; ++a1) {
            const int nna = w.nn[lat * N + a1];
            const bool has = lane < nna;
            const uint32_t e = has ? w.list[((size_t)lat * N + a1) * LC + lane] : 0u;
            const int j = e & F.jmask, img = e >> F.ishift;
            const double tx = (P[j] + V[img]) - P[a1];
            const double ty = (P[N + j] + V[IVC + img]) - P[N + a1];
            const double tz = (P[2 * N + j] + V[2 * IVC + img]) - P[2 * N + a1];
            const double r2 = dist2(tx, ty, tz);
            const bool f = has && r2 < RCSQ;
            const uint32_t bm = __ballot_sync(FULL, f);
            const int cnt = __popc(bm);
            if (nq + cnt > QC) break;                 // cnt <= LC < QC: a chunk always holds >= 1 molecule
            if (lane == 0) w.bmask[lat * N + a1] = bm;
            if (f) {
                const int io = nq + __popc(bm & lt);
                q[io] = tx; q[QC + io] = ty; q[2 * QC + io] = tz; q[3 * QC + io] = r2;
                w.qmeta[io] = (uint32_t)cnt | ((uint32_t)__popc(bm & lt) << 8);   // segment length | position in it
            }
            nq += cnt;
        }
        __syncwarp();
        // ---- bond evaluation: 0.5 * pair energy (molint.F90:464)
        for (int b = 0; b < nq; b += 32) {
            const int r = b + lane;
            if (r < nq) acc += 0.5 * eval_bond(q, r);
        }
        __syncwarp();
        // ---- triplets centred on each molecule of the chunk: lanes = records, rotation pairing inside
        // the molecule's segment as in local_energies_warp (every unordered pair once, all lanes busy)
        for (int b = 0; b < nq; b += 32) {
            const int r = b + lane;
            const bool act = r < nq;
            const uint32_t qm = act ? w.qmeta[r] : 0u;
            const int n = qm & 255, pos = qm >> 8;
            const int half = n >> 1, send = r - pos + n;
            const bool even = !(n & 1);
            const double ux = act ? q[r] : 0.0, uy = act ? q[QC + r] : 0.0, uz = act ? q[2 * QC + r] : 0.0;
            const double g = act ? q[3 * QC + r] : 0.0;
            double tb = 0.0;
            const int maxd = __reduce_max_sync(FULL, half);
            for (int d = 1; d <= maxd; ++d) {
                int c = r + d;
                c = (c >= send) ? c - n : c;
                const bool on = (d <= half) && !(even && d == half && pos >= half);
                c = on ? c : r;
                const double ct = ux * q[c] + uy * q[QC + c] + uz * q[2 * QC + c];
                // no k==i filter here: compute_model_energy has none (molint.F90:480-483)
                const double dd = ct - CK.cos0;
                tb += on ? q[3 * QC + c] * dd * dd : 0.0;
            }
            acc += CK.leps * g * tb;
        }
        __syncwarp();
        a = a1;
    }
    return warp_sum(acc);
}

}  // namespace mw
