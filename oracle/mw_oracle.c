/*
 * mw_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the hot path of keb721/mc_water_ls_mw:
 * molint.F90 (module energy) and the move routines of mc_moves.F90, in the
 * reference's operation and summation order.  Compile with
 * -ffp-contract=off: the state arithmetic (positions, cell, transforms) must be
 * free of FMA contraction so that the CUDA path, which uses explicit
 * round-to-nearest mul/add intrinsics for the same expressions, is bit-identical.
 *
 * Every function cites the reference file:line it follows.  See mw_oracle.h
 * for the pinning status ("parity unpinned" for energies / accept counts).
 */
#include "mw_oracle.h"

#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stddef.h>

/* ------------------------------------------------------------------ */
/* constants.f90:23-24,39,43,59 ; molint.F90:64-74                     */
/* ------------------------------------------------------------------ */
static const double Pi    = 3.141592653589793238462643383279502884197;
static const double invPi = 1.0 / 3.141592653589793238462643383279502884197;
static const double kB    = 1.0 / 3.1577465e5;
static const double ang_to_bohr = 1.0 / 0.5291772108;
static const double aup_to_atm  = 2.90363081e8;

#define MW_SIGMA   (2.3925 * ang_to_bohr)
#define MW_EPSILON (6.189 / 627.509469)
static const double mw_lambda = 23.15;
static const double sw_bigA   = 7.049556277;
static const double sw_B      = 0.6022245584;
static const double sw_gamma  = 1.2;
static const double sw_a      = 1.8;
/* molint.F90:74 -- a default-real (single precision) literal assigned to a dp parameter */
static const double cos0 = (double)(-0.33331324756f);

double orc_const(const char *name)
{
    if (!strcmp(name, "pi")) return Pi;
    if (!strcmp(name, "invpi")) return invPi;
    if (!strcmp(name, "kb")) return kB;
    if (!strcmp(name, "ang_to_bohr")) return ang_to_bohr;
    if (!strcmp(name, "aup_to_atm")) return aup_to_atm;
    if (!strcmp(name, "mw_sigma")) return MW_SIGMA;
    if (!strcmp(name, "mw_epsilon")) return MW_EPSILON;
    if (!strcmp(name, "mw_lambda")) return mw_lambda;
    if (!strcmp(name, "sw_bigA")) return sw_bigA;
    if (!strcmp(name, "sw_B")) return sw_B;
    if (!strcmp(name, "sw_gamma")) return sw_gamma;
    if (!strcmp(name, "sw_a")) return sw_a;
    if (!strcmp(name, "cos0")) return cos0;
    if (!strcmp(name, "wl_factor_default")) return (double)0.05f; /* userparams.f90:32 */
    return NAN;
}

/* Fortran huge(1.0_dp) */
#define F_HUGE DBL_MAX
/* Fortran tiny(1.0_dp) */
#define F_TINY DBL_MIN

/* x**n with integer n: repeated multiplication (binary powering, as the
 * compiler run-time does for integer exponents). mc_moves.F90:588,626,642 */
static double powi(double x, int n)
{
    unsigned m = (n < 0) ? (unsigned)(-n) : (unsigned)n;
    double y = (m & 1) ? x : 1.0;
    while (m >>= 1) {
        x = x * x;
        if (m & 1) y = y * x;
    }
    return (n < 0) ? 1.0 / y : y;
}

/* ------------------------------------------------------------------ */
/* RNG                                                                 */
/* ------------------------------------------------------------------ */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

/* Philox-4x32-10 (Salmon et al., SC'11), raw form: pinned by the published
 * known-answer vectors in tests/test_oracle_golden.py. */
void orc_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    uint32_t k[2] = { key[0], key[1] };
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* counter = (block_lo, block_hi, stream, 0), key = (seed_lo, seed_hi).
 * Two doubles per block: 53 high bits of (w1:w0) and of (w3:w2), times 2^-53. */
void orc_philox_block(uint64_t seed, uint32_t stream, uint64_t block, double out[2])
{
    const uint32_t ctr[4] = { (uint32_t)block, (uint32_t)(block >> 32), stream, 0u };
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t c[4];
    orc_philox_raw(ctr, key, c);
    const uint64_t a = ((uint64_t)c[1] << 32) | c[0];
    const uint64_t b = ((uint64_t)c[3] << 32) | c[2];
    out[0] = (double)(a >> 11) * 0x1.0p-53;
    out[1] = (double)(b >> 11) * 0x1.0p-53;
}

void orc_rng_philox(orc_rng *r, uint64_t seed, uint32_t stream, uint64_t start_index)
{
    memset(r, 0, sizeof(*r));
    r->mode = 0; r->seed = seed; r->stream = stream; r->index = start_index;
}

void orc_rng_fifo(orc_rng *r, const double *u, int64_t n)
{
    memset(r, 0, sizeof(*r));
    r->mode = 1; r->fifo = u; r->fifo_len = n; r->fifo_pos = 0;
}

/* random.f90:87-102 random_uniform_random */
double orc_rng_draw(orc_rng *r)
{
    if (r->mode == 1) {
        if (r->fifo_pos >= r->fifo_len) { r->underrun = 1; return 0.5; }
        return r->fifo[r->fifo_pos++];
    }
    double o[2];
    orc_philox_block(r->seed, r->stream, r->index >> 1, o);
    const double x = o[r->index & 1];
    r->index++;
    return x;
}

/* ------------------------------------------------------------------ */
/* util.f90                                                            */
/* ------------------------------------------------------------------ */
#define H(m, i, j) ((m)[((j) - 1) * 3 + ((i) - 1)])   /* Fortran m(i,j), 1-based */

/* util.f90:16-41 */
double orc_determinant(const double *m)
{
    double det;
    det = H(m,1,1) * (H(m,2,2) * H(m,3,3) - H(m,2,3) * H(m,3,2));
    det = det - H(m,1,2) * (H(m,2,1) * H(m,3,3) - H(m,2,3) * H(m,3,1));
    det = det + H(m,1,3) * (H(m,2,1) * H(m,3,2) - H(m,2,2) * H(m,3,1));
    return det;
}

/* util.f90:43-77 */
void orc_recipmatrix(const double *h, double *r)
{
    H(r,1,1) = H(h,2,2) * H(h,3,3) - H(h,2,3) * H(h,3,2);
    H(r,1,2) = H(h,2,3) * H(h,3,1) - H(h,2,1) * H(h,3,3);
    H(r,1,3) = H(h,2,1) * H(h,3,2) - H(h,2,2) * H(h,3,1);

    H(r,2,1) = H(h,1,3) * H(h,3,2) - H(h,1,2) * H(h,3,3);
    H(r,2,2) = H(h,1,1) * H(h,3,3) - H(h,1,3) * H(h,3,1);
    H(r,2,3) = H(h,1,2) * H(h,3,1) - H(h,1,1) * H(h,3,2);

    H(r,3,1) = H(h,1,2) * H(h,2,3) - H(h,1,3) * H(h,2,2);
    H(r,3,2) = H(h,1,3) * H(h,2,1) - H(h,1,1) * H(h,2,3);
    H(r,3,3) = H(h,1,1) * H(h,2,2) - H(h,1,2) * H(h,2,1);

    const double vol = H(h,1,1) * H(r,1,1) + H(h,1,2) * H(r,1,2) + H(h,1,3) * H(r,1,3);
    for (int k = 0; k < 9; ++k) r[k] = r[k] * 2.0 * Pi / vol;   /* (r*2)*Pi/vol, left to right */
}

/* ------------------------------------------------------------------ */
/* create / destroy / configuration                                    */
/* ------------------------------------------------------------------ */
orc_system *orc_create(int nwater, int nlat)
{
    orc_system *s = (orc_system *)calloc(1, sizeof(orc_system));
    s->nwater = nwater; s->nlat = nlat;
    s->ljr     = (double *)calloc((size_t)3 * nwater * nlat, sizeof(double));
    s->ref_ljr = (double *)calloc((size_t)3 * nwater * nlat, sizeof(double));
    s->ivect   = (double *)calloc((size_t)3 * ORC_MAXIVECT * nlat, sizeof(double));
    s->nn = (int *)calloc((size_t)nwater * nlat, sizeof(int));
    s->jn = (int *)calloc((size_t)ORC_MAXNEIGH * nwater * nlat, sizeof(int));
    s->vn = (int *)calloc((size_t)ORC_MAXNEIGH * nwater * nlat, sizeof(int));
    s->mc_translations = (int *)calloc((size_t)nwater, sizeof(int));
    s->ls = 1;
    s->min_dmu = F_HUGE;
    s->firstpass = 1;
    orc_rng_philox(&s->rng, 20141211ull, 0u, 0ull);
    return s;
}

void orc_destroy(orc_system *s)
{
    if (!s) return;
    free(s->ljr); free(s->ref_ljr); free(s->ivect);
    free(s->nn); free(s->jn); free(s->vn); free(s->mc_translations);
    free(s->histogram); free(s->weight); free(s->unbiased_hist);
    free(s->mu_bin); free(s->binwidth);
    free(s->eta_last_sync); free(s->hist_last_sync); free(s->uhist_last_sync);
    free(s);
}

/* init.f90:80-106 after the Angstrom->Bohr scaling: store cell + positions,
 * reciprocal matrix, and the reference copies. */
void orc_set_config(orc_system *s, const double *ljr, const double *hmatrix)
{
    const int n3 = 3 * s->nwater * s->nlat;
    memcpy(s->ljr, ljr, sizeof(double) * n3);
    memcpy(s->ref_ljr, ljr, sizeof(double) * n3);
    for (int ils = 0; ils < s->nlat; ++ils) {
        memcpy(s->h + 9 * ils, hmatrix + 9 * ils, 9 * sizeof(double));
        memcpy(s->ref_h + 9 * ils, hmatrix + 9 * ils, 9 * sizeof(double));
        orc_recipmatrix(s->h + 9 * ils, s->recip + 9 * ils);
    }
}

/* ------------------------------------------------------------------ */
/* molint.F90                                                          */
/* ------------------------------------------------------------------ */
#define LJR(s, d, imol, ils) ((s)->ljr[((ils) * (s)->nwater + (imol)) * 3 + (d)])
#define REF(s, d, imol, ils) ((s)->ref_ljr[((ils) * (s)->nwater + (imol)) * 3 + (d)])
#define IV(s, d, k, ils)     ((s)->ivect[((ils) * ORC_MAXIVECT + (k)) * 3 + (d)])
#define NN(s, imol, ils)     ((s)->nn[(ils) * (s)->nwater + (imol)])
#define JN(s, ln, imol, ils) ((s)->jn[((ils) * (s)->nwater + (imol)) * ORC_MAXNEIGH + (ln)])
#define VN(s, ln, imol, ils) ((s)->vn[((ils) * (s)->nwater + (imol)) * ORC_MAXNEIGH + (ln)])

/* molint.F90:174-217 */
void orc_compute_ivects(orc_system *s, int ils)
{
    const double *h = s->h + 9 * ils;
    const double rc = sw_a * MW_SIGMA;
    const int im = (int)floor(rc / sqrt(H(h,1,1)*H(h,1,1) + H(h,2,1)*H(h,2,1) + H(h,3,1)*H(h,3,1))) + 1;
    const int jm = (int)floor(rc / sqrt(H(h,1,2)*H(h,1,2) + H(h,2,2)*H(h,2,2) + H(h,3,2)*H(h,3,2))) + 1;
    const int km = (int)floor(rc / sqrt(H(h,1,3)*H(h,1,3) + H(h,2,3)*H(h,2,3) + H(h,3,3)*H(h,3,3))) + 1;

    s->nivect[ils] = (2 * im + 1) * (2 * jm + 1) * (2 * km + 1);
    if (s->nivect[ils] > ORC_MAXIVECT) { s->error = 10; s->nivect[ils] = ORC_MAXIVECT; return; }

    for (int d = 0; d < 3; ++d) IV(s, d, 0, ils) = 0.0;

    int k = 1;
    for (int ic = -im; ic <= im; ++ic) {
        double sx[3];
        for (int d = 0; d < 3; ++d) sx[d] = (double)ic * h[0 * 3 + d];
        for (int jc = -jm; jc <= jm; ++jc) {
            double sy[3];
            for (int d = 0; d < 3; ++d) sy[d] = (double)jc * h[1 * 3 + d];
            for (int kc = -km; kc <= km; ++kc) {
                double sz[3];
                for (int d = 0; d < 3; ++d) sz[d] = (double)kc * h[2 * 3 + d];
                if (abs(ic) + abs(jc) + abs(kc) == 0) continue;
                for (int d = 0; d < 3; ++d) IV(s, d, k, ils) = sx[d] + sy[d] + sz[d];
                ++k;
            }
        }
    }
}

/* molint.F90:501-559 */
void orc_compute_neighbours(orc_system *s, int ils)
{
    const double rn = sw_a * MW_SIGMA * 1.18;
    orc_compute_ivects(s, ils);
    for (int imol = 0; imol < s->nwater; ++imol) {
        const double ilj[3] = { LJR(s,0,imol,ils), LJR(s,1,imol,ils), LJR(s,2,imol,ils) };
        int ni = 0;
        for (int jmol = 0; jmol < s->nwater; ++jmol) {
            double v[3];
            for (int d = 0; d < 3; ++d) v[d] = LJR(s,d,jmol,ils) - ilj[d];
            for (int k = 0; k < s->nivect[ils]; ++k) {
                if (k == 0 && jmol == imol) continue;
                const double tx = v[0] + IV(s,0,k,ils);
                const double ty = v[1] + IV(s,1,k,ils);
                const double tz = v[2] + IV(s,2,k,ils);
                const double r2 = tx * tx + ty * ty + tz * tz;
                if (r2 < rn * rn) {
                    if (ni < ORC_MAXNEIGH) {       /* the reference has no bound check (maxneigh=50) */
                        JN(s, ni, imol, ils) = jmol + 1;
                        VN(s, ni, imol, ils) = k + 1;
                    } else {
                        s->error = 11;
                    }
                    ++ni;
                }
            }
        }
        NN(s, imol, ils) = (ni < ORC_MAXNEIGH) ? ni : ORC_MAXNEIGH;
        if (ni < 16) s->nn_warnings++;             /* molint.F90:552-554 */
    }
}

/* molint.F90:407-499 */
void orc_compute_model_energy(orc_system *s, int ils)
{
    double Evdw = 0.0;
    const double rcsq = MW_SIGMA * sw_a * MW_SIGMA * sw_a;
    for (int imol = 0; imol < s->nwater; ++imol) {
        const double ilj[3] = { LJR(s,0,imol,ils), LJR(s,1,imol,ils), LJR(s,2,imol,ils) };
        const int nni = NN(s, imol, ils);
        for (int ln = 0; ln < nni; ++ln) {
            const int jmol = JN(s, ln, imol, ils) - 1;
            const int ji   = VN(s, ln, imol, ils) - 1;
            double t[3];
            for (int d = 0; d < 3; ++d) t[d] = (LJR(s,d,jmol,ils) + IV(s,d,ji,ils)) - ilj[d];
            const double r2_ij = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
            if (r2_ij < rcsq) {
                const double r1_ij = sqrt(r2_ij);
                double exp2 = exp(MW_SIGMA / (r1_ij - MW_SIGMA * sw_a));
                double tmpE = sw_bigA * MW_EPSILON * (sw_B * ((MW_SIGMA * MW_SIGMA / r2_ij) * (MW_SIGMA * MW_SIGMA / r2_ij)) - 1.0);
                tmpE = tmpE * exp2;
                exp2 = exp(sw_gamma * MW_SIGMA / (r1_ij - MW_SIGMA * sw_a));
                Evdw = Evdw + 0.5 * tmpE;
                for (int ln2 = ln + 1; ln2 < nni; ++ln2) {
                    const int kmol = JN(s, ln2, imol, ils) - 1;
                    const int ki   = VN(s, ln2, imol, ils) - 1;
                    double t2[3];
                    for (int d = 0; d < 3; ++d) t2[d] = (LJR(s,d,kmol,ils) + IV(s,d,ki,ils)) - ilj[d];
                    const double r2_ik = t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2];
                    if (r2_ik < rcsq) {
                        const double r1_ik = sqrt(r2_ik);
                        const double ctheta = (t[0] * t2[0] + t[1] * t2[1] + t[2] * t2[2]) / (r1_ik * r1_ij);
                        const double csq = (ctheta - cos0) * (ctheta - cos0);
                        const double exp1 = exp(sw_gamma * MW_SIGMA / (r1_ik - MW_SIGMA * sw_a));
                        Evdw = Evdw + mw_lambda * MW_EPSILON * exp1 * exp2 * csq;
                    }
                }
            }
        }
    }
    s->model_energy[ils] = Evdw;
}

/* molint.F90:220-404 */
double orc_compute_local_real_energy(orc_system *s, int imol, int ils)
{
    double Evdw = 0.0, Etb = 0.0;
    const double rcsq = MW_SIGMA * sw_a * MW_SIGMA * sw_a;
    double sqlist[2 * ORC_MAXNEIGH], cthetalist[2 * ORC_MAXNEIGH];
    const double ilj[3] = { LJR(s,0,imol,ils), LJR(s,1,imol,ils), LJR(s,2,imol,ils) };
    const int nni = NN(s, imol, ils);

    for (int ln = 0; ln < nni; ++ln) {
        double iEtb = 0.0;
        const int jmol = JN(s, ln, imol, ils) - 1;
        const int ji   = VN(s, ln, imol, ils) - 1;
        double j_ivect[3], jlj[3], t[3];
        for (int d = 0; d < 3; ++d) {
            j_ivect[d] = IV(s, d, ji, ils);
            jlj[d] = LJR(s, d, jmol, ils) + j_ivect[d];
            t[d] = jlj[d] - ilj[d];
        }
        const double r2_ij = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];

        if (r2_ij < rcsq) {
            const double ir1_ij = 1.0 / sqrt(r2_ij);
            const double r1_ij = ir1_ij * r2_ij;
            const double isr1_ij = 1.0 / (r1_ij - MW_SIGMA * sw_a);
            const double exp2 = exp(MW_SIGMA * isr1_ij);
            const double exp3 = exp(sw_gamma * MW_SIGMA * isr1_ij);
            const double s2 = MW_SIGMA * MW_SIGMA * ir1_ij * ir1_ij;
            double tmpE = sw_bigA * MW_EPSILON * (sw_B * (s2 * s2) - 1.0);
            tmpE = tmpE * exp2;
            Evdw = Evdw + tmpE;

            int vl = 0;
            /* jmol--imol--kmol, ln2 > ln   (:302-318) */
            for (int ln2 = ln + 1; ln2 < nni; ++ln2) {
                const int kmol = JN(s, ln2, imol, ils) - 1;
                const int ki   = VN(s, ln2, imol, ils) - 1;
                double t2[3];
                for (int d = 0; d < 3; ++d) t2[d] = (LJR(s,d,kmol,ils) + IV(s,d,ki,ils)) - ilj[d];
                sqlist[vl] = t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2];
                cthetalist[vl] = (t[0] * t2[0] + t[1] * t2[1] + t[2] * t2[2]) * ir1_ij;
                ++vl;
            }
            /* imol--jmol--kmol, all neighbours of jmol   (:320-343) */
            for (int d = 0; d < 3; ++d) t[d] = -t[d];
            const int nnj = NN(s, jmol, ils);
            for (int ln2 = 0; ln2 < nnj; ++ln2) {
                const int kmol = JN(s, ln2, jmol, ils) - 1;
                const int ki   = VN(s, ln2, jmol, ils) - 1;
                double t2[3];
                for (int d = 0; d < 3; ++d)
                    t2[d] = ((LJR(s,d,kmol,ils) + IV(s,d,ki,ils)) + j_ivect[d]) - jlj[d];
                sqlist[vl] = t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2];
                cthetalist[vl] = (t[0] * t2[0] + t[1] * t2[1] + t[2] * t2[2]) * ir1_ij;
                ++vl;
            }
            /* :354-386.  Out-of-range entries carry preflist = 0 in the reference
             * (0 * exp(stale scratch)); restated as "skip". */
            for (int ki = 0; ki < vl; ++ki) {
                if (sqlist[ki] < rcsq) {
                    const double vinv = 1.0 / sqrt(sqlist[ki]);
                    double vexp = vinv * sqlist[ki] - MW_SIGMA * sw_a;
                    vexp = sw_gamma * MW_SIGMA / vexp;
                    const double ct = cthetalist[ki] * vinv;
                    double pref;
                    if (ct < 0.99) pref = (ct - cos0) * (ct - cos0);
                    else pref = 0.0;
                    iEtb = iEtb + pref * exp(vexp);
                }
            }
            iEtb = iEtb * exp3;
        }
        Etb = Etb + iEtb;
    }
    Evdw = Evdw + mw_lambda * MW_EPSILON * Etb;
    return Evdw;
}

/* molint.F90:91-153 */
void orc_energy_init(orc_system *s)
{
    for (int ils = 0; ils < s->nlat; ++ils)
        s->volume[ils] = fabs(orc_determinant(s->h + 9 * ils));
    for (int ils = 0; ils < s->nlat; ++ils) orc_compute_ivects(s, ils);
    for (int ils = 0; ils < s->nlat; ++ils) {
        orc_compute_neighbours(s, ils);
        orc_compute_model_energy(s, ils);
    }
}

/* ------------------------------------------------------------------ */
/* mc_moves.F90                                                        */
/* ------------------------------------------------------------------ */
/* userparams.f90:14-79 */
void orc_params_default(orc_params *p)
{
    memset(p, 0, sizeof(*p));
    p->temperature = 240.0;
    p->pressure = 1.0 / aup_to_atm;
    p->npt = 1;
    p->mc_max_trans = 0.6;
    p->mc_dv_max = 0.1;
    p->mc_target_ratio = 0.5;
    p->wl_factor = (double)0.05f;        /* default-real literal, userparams.f90:32 */
    p->wl_swetnam = 0;
    p->wl_alpha = 1.0;
    p->eta_interp = 1;
    p->samplerun = 0;
    p->leshift = 0;
    p->nbins = 201;
    p->mu_min = -8000.0; p->mu_max = 8000.0;
    p->allow_switch = p->allow_vol = p->allow_trans = 1;
    p->mc_trans_prob = 0.5; p->mc_vol_prob = 0.01; p->mc_switch_prob = 0.0;
    p->mc_always_switch = 1;
    p->list_update_int = 50;
    p->eq_mc_cycles = 25000;
    p->max_mc_cycles = 1000;
    p->eq_adjust_mc = 0;
    p->monitor_int = 1000;
    p->dd = 0;
    p->window_overlap = 2;
    p->ls = 1;
}

/* order parameter, left-to-right association: mc_moves.F90:859-861, :2255-2257, main.f90:172-174 */
static double mu_flat(const orc_system *s)
{
    const double beta = 1.0 / (kB * s->p.temperature);
    double mu = s->model_energy[0] + s->p.pressure * s->volume[0] - s->model_energy[1] - s->p.pressure * s->volume[1];
    if (s->p.leshift) mu = mu - s->ref_enthalpy[0] + s->ref_enthalpy[1];
    mu = mu * beta - (double)s->nwater * log(s->volume[0] / s->volume[1]);
    return mu;
}

/* order parameter, parenthesised association: mc_moves.F90:1370-1372, :1525-1527, :1583-1585 */
static double mu_paren(const orc_system *s)
{
    const double beta = 1.0 / (kB * s->p.temperature);
    double mu = (s->model_energy[0] + s->p.pressure * s->volume[0]) - (s->model_energy[1] + s->p.pressure * s->volume[1]);
    if (s->p.leshift) mu = mu - s->ref_enthalpy[0] + s->ref_enthalpy[1];
    mu = mu * beta - (double)s->nwater * log(s->volume[0] / s->volume[1]);
    return mu;
}

static double gp_ratio(double a, double ssum, int Ns)   /* mc_moves.F90:584-594 / :604-613 */
{
    double r = 1.1, r_new;
    int k = 0;
    for (;;) {
        ++k;
        const double tmpsum = a * (1.0 - powi(r, Ns)) / (1.0 - r);
        r_new = r * pow(ssum / tmpsum, 1.0 / (double)Ns);
        if (fabs(r_new - r) <= 2.0 * DBL_EPSILON) break;
        if (k > 1000000) break;
        r = r_new;
    }
    return r;
}

/* The part of the start-up sequence that defines hot-path state:
 * main.f90:146-150 (ref_enthalpy), mc_moves.F90:557-656 (bin grid), :659-722
 * (windows), :734-776 (weights), :781-814 (log_unbiased_norm), :857-872.
 * Expects orc_energy_init() to have run.  file_weights == NULL means "no
 * eta_weights.dat in the run directory". */
int orc_mc_init(orc_system *s, const orc_params *pin, int rank, int size,
                const double *file_weights, int n_file_weights, double file_wl_factor)
{
    s->p = *pin;
    orc_params *p = &s->p;
    s->rank = rank; s->size = size;
    s->ls = p->ls;
    const int N = s->nwater;

    /* main.f90:146-150 */
    for (int ils = 0; ils < s->nlat; ++ils) {
        s->ref_enthalpy[ils] = s->model_energy[ils];
        if (p->npt) s->ref_enthalpy[ils] = s->ref_enthalpy[ils] + p->pressure * s->volume[ils];
    }
    if (fabs(p->input_ref_enthalpy[0]) > F_TINY || fabs(p->input_ref_enthalpy[1]) > F_TINY) {
        s->ref_enthalpy[0] = p->input_ref_enthalpy[0];
        s->ref_enthalpy[1] = p->input_ref_enthalpy[1];
    }

    memset(s->mc_translations, 0, sizeof(int) * N);
    if (p->nbins % 2 == 0) p->nbins = p->nbins + 1;            /* :557 */
    const int nb = p->nbins;
    free(s->mu_bin); free(s->binwidth); free(s->weight); free(s->histogram); free(s->unbiased_hist);
    free(s->eta_last_sync); free(s->hist_last_sync); free(s->uhist_last_sync);
    s->mu_bin = (double *)calloc(nb, sizeof(double));
    s->binwidth = (double *)calloc(nb, sizeof(double));
    s->weight = (double *)calloc(nb, sizeof(double));
    s->histogram = (double *)calloc(nb, sizeof(double));
    s->unbiased_hist = (double *)calloc(nb, sizeof(double));
    s->eta_last_sync = (double *)calloc(nb, sizeof(double));   /* comms_mpi.f90:73-104 */
    s->hist_last_sync = (double *)calloc(nb, sizeof(double));
    s->uhist_last_sync = (double *)calloc(nb, sizeof(double));

    s->s_pos = fabs(p->mu_max) - 0.5;
    s->s_neg = fabs(p->mu_min) - 0.5;
    s->a_pos = 1.0; s->a_neg = 1.0;
    const int Ns = nb / 2;
    s->r_pos = gp_ratio(s->a_pos, s->s_pos, Ns);
    s->r_neg = gp_ratio(s->a_neg, s->s_neg, Ns);

    /* :622-648, bins 1-based in the comments */
    double mu_u = -0.5, mu_l;
    int k = 0;
    for (int ibin = nb / 2; ibin >= 1; --ibin) {
        mu_l = mu_u - s->a_neg * powi(s->r_neg, k);
        s->mu_bin[ibin - 1] = 0.5 * (mu_u + mu_l);
        s->binwidth[ibin - 1] = mu_u - mu_l;
        mu_u = mu_l;
        ++k;
    }
    s->mu_bin[nb / 2] = 0.0;
    s->binwidth[nb / 2] = 1.0;
    mu_l = 0.5;
    k = 0;
    for (int ibin = nb / 2 + 2; ibin <= nb; ++ibin) {
        mu_u = mu_l + s->a_pos * powi(s->r_pos, k);
        s->mu_bin[ibin - 1] = 0.5 * (mu_u + mu_l);
        s->binwidth[ibin - 1] = mu_u - mu_l;
        mu_l = mu_u;
        ++k;
    }
    s->av_binwidth = 0.0;
    for (int ibin = 0; ibin < nb; ++ibin) s->av_binwidth = s->av_binwidth + s->binwidth[ibin];
    s->av_binwidth = s->av_binwidth / (double)nb;

    /* :659-722 */
    if (p->dd) {
        const int bpw = nb / size;
        const int ov = (size == 1) ? 0 : p->window_overlap;      /* io.f90:249 */
        if (rank == 0) {
            s->my_start_bin = 1;
            s->my_end_bin = bpw + ov;
            s->my_mu_min = p->mu_min;
            double sum = 0.0;
            for (int i = 0; i < s->my_end_bin; ++i) sum += s->binwidth[i];
            s->my_mu_max = p->mu_min + sum;
        }
        if (size > 1) {
            if (rank >= 1 && rank <= size - 2) {
                s->my_start_bin = rank * bpw - ov;
                s->my_end_bin = (rank + 1) * bpw + ov;
                double sum = 0.0;
                for (int i = 0; i < s->my_start_bin - 1; ++i) sum += s->binwidth[i];
                s->my_mu_min = p->mu_min + sum;
                sum = 0.0;
                for (int i = 0; i < s->my_end_bin; ++i) sum += s->binwidth[i];
                s->my_mu_max = p->mu_min + sum;
            }
            if (rank == size - 1) {
                s->my_start_bin = rank * bpw - ov;
                s->my_end_bin = nb;
                double sum = 0.0;
                for (int i = 0; i < s->my_start_bin - 1; ++i) sum += s->binwidth[i];
                s->my_mu_min = p->mu_min + sum;
                s->my_mu_max = p->mu_max;
            }
        }
        if (s->my_mu_max < 0.0) s->ls = 1;
        if (s->my_mu_min > 0.0) s->ls = 2;
    } else {
        s->my_start_bin = 1; s->my_end_bin = nb;
        s->my_mu_min = p->mu_min; s->my_mu_max = p->mu_max;
    }

    /* :734-776 */
    s->orig_wl_factor = p->wl_factor;
    s->wl_factor = p->wl_factor;
    if (s->nlat == 2) {
        if (file_weights) {
            if (file_wl_factor > 1e-10f) {       /* default-real literal 1e-10, :757 */
                s->wl_factor = (s->wl_factor < file_wl_factor) ? s->wl_factor : file_wl_factor;
                if (p->samplerun) s->wl_factor = 0.0;
            }
            for (int i = 0; i < n_file_weights && i < nb; ++i) s->weight[i] = file_weights[i];
        }
        /* comms_allreduce_eta at :776 leaves every rank with rank 0's weights and
         * eta_last_sync = weight */
        memcpy(s->eta_last_sync, s->weight, sizeof(double) * nb);

        /* :781-806 */
        double hits = (double)p->max_mc_cycles - (double)p->eq_mc_cycles;
        hits = hits * (double)(size * N) / (double)nb;
        double incr = hits * s->av_binwidth;
        s->log_unbiased_norm = log(incr) + s->weight[0];
        for (int kk = 1; kk < nb; ++kk) {
            incr = hits * s->av_binwidth;
            if (s->log_unbiased_norm > s->weight[kk] + log(incr)) {
                s->log_unbiased_norm = s->log_unbiased_norm +
                    log(1.0 + incr * exp(s->weight[kk] - s->log_unbiased_norm));
            } else {
                s->log_unbiased_norm = log(incr) + s->weight[kk] +
                    log(1.0 + exp(s->log_unbiased_norm - s->weight[kk]) / incr);
            }
        }
        if (p->dd) {                                            /* :808-814 */
            for (int i = 0; i < s->my_start_bin - 1; ++i) s->weight[i] = 0.0;
            for (int i = s->my_end_bin; i < nb; ++i) s->weight[i] = 0.0;
        }
    }

    /* :857-872, main.f90:170-175 */
    if (s->nlat == 2) s->ls_mu = mu_flat(s);
    else s->ls_mu = 0.0;      /* never assigned for a single box (save variable) */
    s->walker_in_window = p->dd ? 0 : 1;

    s->mc_cycle_num = 0;
    s->acc_r = s->acc_v = s->acc_s = s->att_r = s->att_v = s->att_s = 0;
    s->average_energy[0] = s->average_energy[1] = 0.0;
    s->max_dmu = 0.0; s->min_dmu = F_HUGE;
    s->wl_invt_active = 0;
    /* firstcycle (:85) is cleared when a smaller increment came from eta_weights.dat (:817-821) */
    s->firstcycle = !(s->nlat == 2 && s->wl_factor < s->orig_wl_factor);
    s->histogram_reset = 0;
    s->sumhist = 0.0;
    s->firstpass = 1;
    s->error = 0;
    return 0;
}

/* mc_moves.F90:2187-2215 ; returns the 1-based bin number */
int orc_mu_to_bin(const orc_system *s, double mu)
{
    const int nb = s->p.nbins;
    if (fabs(mu) <= 0.5) return nb / 2 + 1;
    if (mu > 0.0) {
        const double arg = 1.0 - (mu - 0.5) * (1.0 - s->r_pos) / s->a_pos;
        return nb / 2 + 2 + (int)(log(arg) / log(s->r_pos));
    } else {
        const double arg = 1.0 - (fabs(mu) - 0.5) * (1.0 - s->r_neg) / s->a_neg;
        return nb / 2 - (int)(log(arg) / log(s->r_neg));
    }
}

/* mc_moves.F90:893-964.  The reference leaves the result undefined when the
 * walker has not reached its window (:913); the oracle defines it as 0. */
double orc_eta_weight(orc_system *s, double mu)
{
    if (!s->walker_in_window) return 0.0;
    if (mu < s->my_mu_min) return F_HUGE;
    if (mu > s->my_mu_max) return F_HUGE;
    const int k = orc_mu_to_bin(s, mu);          /* 1-based */
    const double *w = s->weight - 1, *bw = s->binwidth - 1, *mb = s->mu_bin - 1;
    if (s->p.eta_interp) {
        double g;
        if (k == s->my_start_bin) {
            g = 2.0 * (w[k + 1] - w[k]) / (bw[k] + bw[k + 1]);
            return w[k] + (mu - mb[k]) * g;
        } else if (k == s->my_end_bin) {
            g = 2.0 * (w[k] - w[k - 1]) / (bw[k] + bw[k - 1]);
            return w[k] + (mu - mb[k]) * g;
        } else if (mu > mb[k]) {
            g = 2.0 * (w[k + 1] - w[k]) / (bw[k] + bw[k + 1]);
            return w[k] + (mu - mb[k]) * g;
        } else {
            g = 2.0 * (w[k] - w[k - 1]) / (bw[k] + bw[k - 1]);
            return w[k - 1] + (mu - mb[k - 1]) * g;
        }
    }
    return w[k];
}

static inline int partner(const orc_system *s, int ls) { return (s->nlat == 2) ? (3 - ls) : 1; }

/* mc_moves.F90:966-1213 */
void orc_mc_water_translation(orc_system *s)
{
    const int N = s->nwater, nlat = s->nlat;
    const int ls = s->ls, lsn = partner(s, ls);
    const double beta = 1.0 / (kB * s->p.temperature);
    double old_energy[2] = {0, 0}, new_energy[2] = {0, 0}, deltaE[2] = {0, 0}, backup[2] = {0, 0};
    double transvec[2][3];

    double x = orc_rng_draw(&s->rng);
    int imol = (int)(x * (double)N) + 1;
    if (imol > N) imol = N;
    imol -= 1;
    s->mc_translations[imol] += 1;

    for (int ils = 0; ils < nlat; ++ils) {
        old_energy[ils] = orc_compute_local_real_energy(s, imol, ils);
        backup[ils] = s->model_energy[ils];
        s->model_energy[ils] = s->model_energy[ils] - old_energy[ils];
    }

    x = orc_rng_draw(&s->rng);
    double y = orc_rng_draw(&s->rng);
    double z = orc_rng_draw(&s->rng);
    x = 2.0 * x - 1.0; y = 2.0 * y - 1.0; z = 2.0 * z - 1.0;
    const double norm = 1.0 / sqrt(x * x + y * y + z * z);
    x = x * norm; y = y * norm; z = z * norm;
    const double r = orc_rng_draw(&s->rng) * 2.0 - 1.0;
    x = x * s->p.mc_max_trans * r;
    y = y * s->p.mc_max_trans * r;
    z = z * s->p.mc_max_trans * r;

    const double *rm = s->recip + 9 * (ls - 1);
    double sx = H(rm,1,1) * x + H(rm,2,1) * y + H(rm,3,1) * z;
    double sy = H(rm,1,2) * x + H(rm,2,2) * y + H(rm,3,2) * z;
    double sz = H(rm,1,3) * x + H(rm,2,3) * y + H(rm,3,3) * z;
    sx = sx * 0.5 * invPi; sy = sy * 0.5 * invPi; sz = sz * 0.5 * invPi;

    transvec[ls - 1][0] = x; transvec[ls - 1][1] = y; transvec[ls - 1][2] = z;
    if (nlat == 2) {
        const double *hm = s->h + 9 * (lsn - 1);
        for (int d = 1; d <= 3; ++d)
            transvec[lsn - 1][d - 1] = H(hm,d,1) * sx + H(hm,d,2) * sy + H(hm,d,3) * sz;
    }

    for (int ils = 0; ils < nlat; ++ils) {
        for (int d = 0; d < 3; ++d) LJR(s,d,imol,ils) = LJR(s,d,imol,ils) + transvec[ils][d];
        new_energy[ils] = orc_compute_local_real_energy(s, imol, ils);
        s->model_energy[ils] = s->model_energy[ils] + new_energy[ils];
        deltaE[ils] = new_energy[ils] - old_energy[ils];
    }

    double diffkT;
    if (nlat == 1) {
        diffkT = beta * deltaE[0];
    } else {
        const double eta_old = orc_eta_weight(s, s->ls_mu);
        s->ls_mu = s->ls_mu + (deltaE[0] - deltaE[1]) * beta;
        const double eta_new = orc_eta_weight(s, s->ls_mu);
        diffkT = deltaE[ls - 1] * beta + eta_new - eta_old;
    }

    const double zeta = orc_rng_draw(&s->rng);
    const double e = exp(-diffkT);
    if (zeta < ((e < 1.0) ? e : 1.0)) {
        s->acc_r += 1;
        const double dmu = fabs(deltaE[0] - deltaE[1]) * beta;
        if (dmu < s->min_dmu) s->min_dmu = dmu;
        if (dmu > s->max_dmu) s->max_dmu = dmu;
    } else {
        for (int ils = 0; ils < nlat; ++ils) {
            for (int d = 0; d < 3; ++d) LJR(s,d,imol,ils) = LJR(s,d,imol,ils) - transvec[ils][d];
            s->model_energy[ils] = backup[ils];
        }
        if (nlat == 2) s->ls_mu = s->ls_mu - (deltaE[0] - deltaE[1]) * beta;
    }
}

/* fractional rescale of one position: s = recip^T r /(2 pi); r += (h s - r)
 * mc_moves.F90:1290-1315 (and the three identical blocks :1322-1347, :1443-1468, :1475-1500) */
static void rescale_pos(double *pos, const double *rm, const double *hm)
{
    const double o0 = pos[0], o1 = pos[1], o2 = pos[2];
    double n0 = H(rm,1,1) * o0 + H(rm,2,1) * o1 + H(rm,3,1) * o2;
    double n1 = H(rm,1,2) * o0 + H(rm,2,2) * o1 + H(rm,3,2) * o2;
    double n2 = H(rm,1,3) * o0 + H(rm,2,3) * o1 + H(rm,3,3) * o2;
    n0 = n0 * 0.5 * invPi; n1 = n1 * 0.5 * invPi; n2 = n2 * 0.5 * invPi;
    double tv[3];
    for (int d = 1; d <= 3; ++d) tv[d - 1] = H(hm,d,1) * n0 + H(hm,d,2) * n1 + H(hm,d,3) * n2;
    tv[0] = tv[0] - o0; tv[1] = tv[1] - o1; tv[2] = tv[2] - o2;
    pos[0] = pos[0] + tv[0]; pos[1] = pos[1] + tv[1]; pos[2] = pos[2] + tv[2];
}

/* mc_moves.F90:1216-1534 */
void orc_mc_volume(orc_system *s)
{
    const int N = s->nwater, nlat = s->nlat, ls = s->ls;
    const double beta = 1.0 / (kB * s->p.temperature);
    double backup[2], old_energy[2], new_energy[2] = {0, 0}, deltaE[2] = {0, 0}, old_volume[2];
    double old_h[18], old_recip[18], dh[9];
    double old_eta = 0.0, new_eta = 0.0, old_mu = 0.0;

    for (int ils = 0; ils < nlat; ++ils) {
        backup[ils] = s->model_energy[ils];
        old_energy[ils] = s->model_energy[ils];
        orc_recipmatrix(s->h + 9 * ils, s->recip + 9 * ils);
    }
    memcpy(old_h, s->h, sizeof(double) * 9 * nlat);
    memcpy(old_recip, s->recip, sizeof(double) * 9 * nlat);
    for (int ils = 0; ils < nlat; ++ils) old_volume[ils] = s->volume[ils];

    double x = orc_rng_draw(&s->rng);
    const int idim = (int)(x * 3.0) + 1;
    x = orc_rng_draw(&s->rng);
    const int jdim = (int)(x * 3.0) + 1;
    x = orc_rng_draw(&s->rng);

    memset(dh, 0, sizeof(dh));
    H(dh, idim, jdim) = (2.0 * x - 1.0) * s->p.mc_dv_max;
    H(dh, jdim, idim) = H(dh, idim, jdim);
    for (int ils = 0; ils < nlat; ++ils)
        for (int k = 0; k < 9; ++k) s->h[9 * ils + k] = s->h[9 * ils + k] + dh[k];

    for (int ils = 0; ils < nlat; ++ils) {
        const double *rm = s->recip + 9 * ils, *hm = s->h + 9 * ils;
        for (int imol = 0; imol < N; ++imol) rescale_pos(&LJR(s,0,imol,ils), rm, hm);
        for (int imol = 0; imol < N; ++imol) rescale_pos(&REF(s,0,imol,ils), rm, hm);
        s->volume[ils] = fabs(orc_determinant(hm));
        orc_recipmatrix(hm, s->recip + 9 * ils);
        orc_compute_ivects(s, ils);
        orc_compute_model_energy(s, ils);
        new_energy[ils] = s->model_energy[ils];
    }
    for (int ils = 0; ils < nlat; ++ils) deltaE[ils] = new_energy[ils] - old_energy[ils];

    if (nlat == 2) {
        old_eta = orc_eta_weight(s, s->ls_mu);
        old_mu = s->ls_mu;
        s->ls_mu = mu_paren(s);
        new_eta = orc_eta_weight(s, s->ls_mu);
    }

    x = orc_rng_draw(&s->rng);

    const double diffkT = beta * deltaE[ls - 1] + new_eta - old_eta
        + beta * s->p.pressure * (s->volume[ls - 1] - old_volume[ls - 1])
        - (double)N * log(s->volume[ls - 1] / old_volume[ls - 1]);

    double compare = exp(-diffkT);
    compare = (compare < 1.0) ? compare : 1.0;

    if (x < compare) {
        s->acc_v += 1;
        if (nlat == 2) {
            const double dmu = fabs(old_mu - s->ls_mu);
            if (dmu < s->min_dmu) s->min_dmu = dmu;
            if (dmu > s->max_dmu) s->max_dmu = dmu;
        }
    } else {
        for (int ils = 0; ils < nlat; ++ils) s->volume[ils] = old_volume[ils];
        memcpy(s->h, old_h, sizeof(double) * 9 * nlat);
        for (int ils = 0; ils < nlat; ++ils) {
            /* :1443-1500: recip is still the NEW cell's, h is the OLD one again */
            const double *rm = s->recip + 9 * ils, *hm = s->h + 9 * ils;
            for (int imol = 0; imol < N; ++imol) rescale_pos(&LJR(s,0,imol,ils), rm, hm);
            for (int imol = 0; imol < N; ++imol) rescale_pos(&REF(s,0,imol,ils), rm, hm);
        }
        memcpy(s->recip, old_recip, sizeof(double) * 9 * nlat);
        for (int ils = 0; ils < nlat; ++ils) orc_compute_ivects(s, ils);
        for (int ils = 0; ils < nlat; ++ils) s->model_energy[ils] = backup[ils];
        if (nlat == 2) s->ls_mu = mu_paren(s);
    }
}

/* mc_moves.F90:1536-1594 */
void orc_mc_lattice_switch(orc_system *s)
{
    if (s->nlat == 1) { s->error = 20; return; }
    const double beta = 1.0 / (kB * s->p.temperature);
    const int ls = s->ls, lsn = partner(s, ls);
    const double old_eta = orc_eta_weight(s, s->ls_mu);
    const double new_eta = orc_eta_weight(s, s->ls_mu);
    double diffkT;
    if (s->p.npt) {
        diffkT = beta * s->model_energy[lsn - 1] - beta * s->model_energy[ls - 1]
            + beta * s->p.pressure * (s->volume[lsn - 1] - s->volume[ls - 1])
            - (double)s->nwater * log(s->volume[lsn - 1] / s->volume[ls - 1]) + new_eta - old_eta;
        if (s->p.leshift) diffkT = diffkT - beta * s->ref_enthalpy[lsn - 1] + beta * s->ref_enthalpy[ls - 1];
    } else {
        diffkT = beta * s->model_energy[lsn - 1] - beta * s->model_energy[ls - 1] + new_eta - old_eta;
        if (s->p.leshift) diffkT = diffkT - beta * s->ref_enthalpy[lsn - 1] + beta * s->ref_enthalpy[ls - 1];
    }
    const double e = exp(-diffkT);
    const double compare = (e < 1.0) ? e : 1.0;
    const double x = orc_rng_draw(&s->rng);
    if (x < compare) {
        s->acc_s += 1;
        s->ls_mu = mu_paren(s);
        s->ls = lsn;
    }
}

/* mc_moves.F90:1597-1689 */
void orc_mc_update_wl_bins(orc_system *s)
{
    const orc_params *p = &s->p;
    const int nb = p->nbins;
    if (s->mc_cycle_num < p->eq_mc_cycles) return;
    const int k = orc_mu_to_bin(s, s->ls_mu);
    if (k < 1 || k > nb) return;
    s->histogram[k - 1] = s->histogram[k - 1] + s->av_binwidth / s->binwidth[k - 1];
    if (p->samplerun) {
        const double incr = s->av_binwidth / s->binwidth[k - 1];
        s->unbiased_hist[k - 1] = s->unbiased_hist[k - 1] +
            incr * exp(orc_eta_weight(s, s->ls_mu) - s->log_unbiased_norm);
        return;
    }
    if (p->wl_swetnam) {
        s->sumhist = s->sumhist + 1.0;
        double f = 0.0;
        for (int i = 0; i < nb; ++i) {
            const double binfrac = s->binwidth[i] / (p->mu_max - p->mu_min - 1.0);
            const double d = s->histogram[i] * s->binwidth[i] / s->sumhist - binfrac;
            f = f + d * d;
        }
        f = sqrt(f / (double)nb);
        f = log(f);
        f = f * p->wl_alpha * (double)nb;
        s->wl_factor = (f < s->orig_wl_factor) ? f : s->orig_wl_factor;
    } else if (s->wl_invt_active) {
        const double t = (double)nb / (double)(s->mc_cycle_num * s->nwater);
        s->wl_factor = (s->wl_factor < t) ? s->wl_factor : t;
    }
    const double incr = s->wl_factor;
    s->weight[k - 1] = s->weight[k - 1] + s->av_binwidth * incr / s->binwidth[k - 1];
    double minbin = s->weight[s->my_start_bin - 1];
    for (int i = s->my_start_bin; i <= s->my_end_bin; ++i)
        if (s->weight[i - 1] < minbin) minbin = s->weight[i - 1];
    for (int i = s->my_start_bin; i <= s->my_end_bin; ++i) s->weight[i - 1] = s->weight[i - 1] - minbin;
}

/* mc_moves.F90:117-255: the hot part of one MC cycle (move-type probabilities,
 * dd window sanity, list refresh, nwater trial moves, average-energy
 * accumulation).  The periodic host bookkeeping of :257-316 is NOT part of
 * the hot path; its state effects are orc_mc_monitor / orc_allreduce_bins /
 * orc_mc_chain_sync, called by the test harness at the reference's intervals. */
int orc_mc_cycle(orc_system *s)
{
    orc_params *p = &s->p;
    s->mc_cycle_num += 1;
    if (s->firstpass) {
        s->firstpass = 0;
        if (p->mc_always_switch) p->mc_switch_prob = 0.0;
        if (!p->allow_switch)    p->mc_switch_prob = 0.0;
        if (!p->npt)             p->mc_vol_prob = 0.0;
        if (!p->allow_vol)       p->mc_vol_prob = 0.0;
        if (!p->allow_trans)     p->mc_trans_prob = 0.0;
        const double sum_prob = p->mc_trans_prob + p->mc_vol_prob + p->mc_switch_prob;
        s->transP = p->mc_trans_prob / sum_prob;
        s->volP = p->mc_vol_prob / sum_prob;
        s->swP = p->mc_switch_prob / sum_prob;
        s->volP = s->volP + s->transP;
        s->swP = s->swP + s->volP;
        if (s->swP < 0.999) { s->error = 30; return s->error; }
    }
    if (p->dd) {
        if (s->mc_cycle_num < p->eq_mc_cycles) {
            s->walker_in_window = (s->ls_mu > s->my_mu_min) && (s->ls_mu < s->my_mu_max);
        } else if (s->mc_cycle_num == p->eq_mc_cycles) {
            if (!s->walker_in_window) { s->error = 31; return s->error; }
        } else {
            s->walker_in_window = 1;
        }
    }
    if (s->mc_cycle_num % p->list_update_int == 0)
        for (int ils = 0; ils < s->nlat; ++ils) orc_compute_neighbours(s, ils);

    const int dd_eq = p->dd && (s->mc_cycle_num < p->eq_mc_cycles);
    for (int imove = 0; imove < s->nwater; ++imove) {
        const double xi = orc_rng_draw(&s->rng);
        if (xi < s->transP) {
            orc_mc_water_translation(s);
            orc_mc_update_wl_bins(s);
            s->att_r += 1;
        } else if (xi < s->volP) {
            orc_mc_volume(s);
            orc_mc_update_wl_bins(s);
            s->att_v += 1;
        } else if (xi < s->swP) {
            if (!dd_eq) { orc_mc_lattice_switch(s); s->att_s += 1; }
        }
        if (p->mc_always_switch) {
            if (!dd_eq) { orc_mc_lattice_switch(s); s->att_s += 1; }
        }
    }
    for (int ils = 0; ils < s->nlat; ++ils) {
        s->average_energy[ils] = s->average_energy[ils] + s->model_energy[ils];
        if (p->npt) s->average_energy[ils] = s->average_energy[ils] + p->pressure * s->volume[ils];
    }
    if (s->rng.underrun) s->error = 40;
    return s->error;
}

int orc_mc_run(orc_system *s, int ncycles)
{
    for (int c = 0; c < ncycles; ++c)
        if (orc_mc_cycle(s)) return s->error;
    return 0;
}

/* State effects of mc_monitor_stats: step-size adjustment during equilibration
 * (mc_moves.F90:1722-1732), energy re-synchronisation (:1786-1792) and the
 * counter reset (:1797-1810).  File output is out of scope. */
void orc_mc_monitor(orc_system *s)
{
    orc_params *p = &s->p;
    const double atr = (double)s->acc_r / (double)s->att_r;
    const double avr = (double)s->acc_v / (double)s->att_v;
    if (p->eq_adjust_mc && s->mc_cycle_num < p->eq_mc_cycles) {
        /* Fortran max(x, c) with x = NaN (0/0) is processor dependent; the oracle
         * keeps fmax semantics (NaN -> the floor). */
        p->mc_max_trans = fmax(p->mc_max_trans * atr / p->mc_target_ratio, 0.1);
        p->mc_dv_max = fmax(p->mc_dv_max * avr / p->mc_target_ratio, 0.0001);
    }
    for (int ils = 0; ils < s->nlat; ++ils) orc_compute_model_energy(s, ils);
    s->acc_r = s->acc_v = s->acc_s = s->att_r = s->att_v = s->att_s = 0;
    memset(s->mc_translations, 0, sizeof(int) * s->nwater);
    s->average_energy[0] = s->average_energy[1] = 0.0;
    s->max_dmu = 0.0; s->min_dmu = F_HUGE;
}

/* mc_moves.F90:2217-2416 */
void orc_mc_chain_sync(orc_system *s)
{
    if (s->nlat != 2) return;
    const int N = s->nwater;
    orc_compute_model_energy(s, 0);
    orc_compute_model_energy(s, 1);
    s->ls_mu = mu_flat(s);
    double hdiff[9];
    for (int k = 0; k < 9; ++k) hdiff[k] = s->h[k] - s->ref_h[k];
    for (int k = 0; k < 9; ++k) s->h[9 + k] = s->ref_h[9 + k] + hdiff[k];
    orc_recipmatrix(s->h, s->recip);
    orc_recipmatrix(s->h + 9, s->recip + 9);
    for (int iw = 0; iw < N; ++iw) {
        double sv[2][3], rsv[2][3], sd[2][3];
        for (int ils = 0; ils < 2; ++ils) {
            const double *rm = s->recip + 9 * ils;
            const double *a = &LJR(s,0,iw,ils), *b = &REF(s,0,iw,ils);
            sv[ils][0] = H(rm,1,1) * a[0] + H(rm,2,1) * a[1] + H(rm,3,1) * a[2];
            sv[ils][1] = H(rm,1,2) * a[0] + H(rm,2,2) * a[1] + H(rm,3,2) * a[2];
            sv[ils][2] = H(rm,1,3) * a[0] + H(rm,2,3) * a[1] + H(rm,3,3) * a[2];
            rsv[ils][0] = H(rm,1,1) * b[0] + H(rm,2,1) * b[1] + H(rm,3,1) * b[2];
            rsv[ils][1] = H(rm,1,2) * b[0] + H(rm,2,2) * b[1] + H(rm,3,2) * b[2];
            rsv[ils][2] = H(rm,1,3) * b[0] + H(rm,2,3) * b[1] + H(rm,3,3) * b[2];
            for (int d = 0; d < 3; ++d) {
                sv[ils][d] = sv[ils][d] * 0.5 * invPi;
                rsv[ils][d] = rsv[ils][d] * 0.5 * invPi;
                sd[ils][d] = sv[ils][d] - rsv[ils][d];
            }
        }
        for (int d = 0; d < 3; ++d) sv[1][d] = rsv[1][d] + sd[0][d];
        const double *hm = s->h + 9;
        /* matmul(hmatrix(:,:,2), svect(:,2)): column-ordered accumulation */
        for (int d = 1; d <= 3; ++d)
            LJR(s, d - 1, iw, 1) = H(hm,d,1) * sv[1][0] + H(hm,d,2) * sv[1][1] + H(hm,d,3) * sv[1][2];
    }
    for (int ils = 0; ils < 2; ++ils) {
        s->volume[ils] = fabs(orc_determinant(s->h + 9 * ils));
        orc_compute_ivects(s, ils);
    }
    orc_compute_model_energy(s, 0);
    orc_compute_model_energy(s, 1);
    s->ls_mu = mu_flat(s);
}

/* comms_mpi.f90:244-277 (eta), :461-493 (hist), :495-530 (uhist): delta-since-
 * last-sync all-reduce-sum over the walkers (= MPI ranks), re-base.  The sum
 * is taken in rank order 0..n-1 (MPI leaves the order unspecified). */
static void allreduce_one(orc_system **w, int n, int nb, size_t off_arr, size_t off_base)
{
    double *buff = (double *)calloc(nb, sizeof(double));
    for (int r = 0; r < n; ++r) {
        double *arr = *(double **)((char *)w[r] + off_arr);
        double *base = *(double **)((char *)w[r] + off_base);
        for (int k = 0; k < nb; ++k) {
            arr[k] = arr[k] - base[k];
            buff[k] = buff[k] + arr[k];
        }
    }
    for (int r = 0; r < n; ++r) {
        double *arr = *(double **)((char *)w[r] + off_arr);
        double *base = *(double **)((char *)w[r] + off_base);
        for (int k = 0; k < nb; ++k) {
            arr[k] = buff[k] + base[k];
            base[k] = arr[k];
        }
    }
    free(buff);
}

void orc_allreduce_bins(orc_system **w, int n)
{
    if (n < 1) return;
    const int nb = w[0]->p.nbins;
    allreduce_one(w, n, nb, offsetof(orc_system, weight), offsetof(orc_system, eta_last_sync));
    allreduce_one(w, n, nb, offsetof(orc_system, histogram), offsetof(orc_system, hist_last_sync));
    if (w[0]->p.samplerun)
        allreduce_one(w, n, nb, offsetof(orc_system, unbiased_hist), offsetof(orc_system, uhist_last_sync));
}

/* mc_moves.F90:403-501 (mc_checkpoint_load) followed by :842-862 (restart branch of mc_init) */
void orc_mc_restore(orc_system *s, int mc_cycle_num, double mc_max_trans, double mc_dv_max, double wl_factor,
                    int wl_invt_active, int ls, const double *histogram, const double *weight,
                    const double *unbiased_hist, const double *hmatrix, const double *ref_ljr, const double *ljr)
{
    const int nb = s->p.nbins, n3 = 3 * s->nwater * s->nlat;
    s->mc_cycle_num = mc_cycle_num;
    s->p.mc_max_trans = mc_max_trans; s->p.mc_dv_max = mc_dv_max;
    s->wl_factor = wl_factor;
    memcpy(s->histogram, histogram, sizeof(double) * nb);
    memcpy(s->weight, weight, sizeof(double) * nb);
    s->wl_invt_active = wl_invt_active ? 1 : 0;
    if (s->p.samplerun) memcpy(s->unbiased_hist, unbiased_hist, sizeof(double) * nb);
    if (!s->p.dd) {                                   /* comms_set_histogram / comms_set_uhistogram */
        memcpy(s->hist_last_sync, s->histogram, sizeof(double) * nb);
        if (s->p.samplerun) memcpy(s->uhist_last_sync, s->unbiased_hist, sizeof(double) * nb);
    }
    double sum = 0.0;
    for (int k = 0; k < nb; ++k) sum = sum + s->histogram[k];
    s->sumhist = sum;
    if (s->wl_factor < s->orig_wl_factor) s->firstcycle = 0;
    memcpy(s->h, hmatrix, sizeof(double) * 9 * s->nlat);
    memcpy(s->ref_ljr, ref_ljr, sizeof(double) * n3);
    memcpy(s->ljr, ljr, sizeof(double) * n3);
    s->ls = ls;
    /* :842-856 */
    for (int ils = 0; ils < s->nlat; ++ils) {
        s->volume[ils] = fabs(orc_determinant(s->h + 9 * ils));
        orc_recipmatrix(s->h + 9 * ils, s->recip + 9 * ils);
        orc_compute_ivects(s, ils);
    }
    if (s->nlat == 2) orc_mc_chain_sync(s);
    for (int ils = 0; ils < s->nlat; ++ils) orc_compute_model_energy(s, ils);
    if (s->nlat == 2) s->ls_mu = mu_flat(s);          /* :857-862 */
}

/* ------------------------------------------------------------------ */
/* periodic bookkeeping on the reduced arrays                          */
/* ------------------------------------------------------------------ */
static long nint_d(double x) { return (long)(x < 0.0 ? x - 0.5 : x + 0.5); }    /* Fortran nint */

/* mc_moves.F90:1936-2185 */
int orc_mc_check_flatness(orc_system **w, int n, const orc_flat_params *fp, orc_flat_report *rep)
{
    orc_flat_report r0; memset(&r0, 0, sizeof(r0));
    if (n < 1) { if (rep) *rep = r0; return 0; }
    const int nb = w[0]->p.nbins;
    const int dd = w[0]->p.dd;
    int rc = 0;
    /* :1961 guard, evaluated per rank on its own histogram (all ranks agree in a valid run) */
    int *skip = (int *)calloc(n, sizeof(int));
    for (int r = 0; r < n; ++r) {
        double sum = 0.0;
        for (int k = 0; k < nb; ++k) sum = sum + w[r]->histogram[k];
        skip[r] = w[r]->p.samplerun || (sum < F_TINY);
    }
    /* :1964-1966 */
    if (!dd && !skip[0]) allreduce_one(w, n, nb, offsetof(orc_system, histogram), offsetof(orc_system, hist_last_sync));
    int flat0 = 0;
    for (int r = 0; r < n; ++r) {
        orc_system *s = w[r];
        orc_flat_report rr; memset(&rr, 0, sizeof(rr));
        rr.wl_factor = s->wl_factor;
        if (skip[r]) { if (r == 0) r0 = rr; continue; }
        rr.checked = 1;
        double mn = s->histogram[0], mx = s->histogram[0];
        for (int k = 1; k < nb; ++k) { if (s->histogram[k] < mn) mn = s->histogram[k]; if (s->histogram[k] > mx) mx = s->histogram[k]; }
        long mini = nint_d(mn);
        if (s->firstcycle && !s->histogram_reset && mini > fp->wl_minhist) {          /* :1973-1980 */
            s->histogram_reset = 1;
            for (int k = 0; k < nb; ++k) { s->histogram[k] = 0.0; s->hist_last_sync[k] = 0.0; }
            rr.hist_reset = 1;
            if (r == 0) r0 = rr;
            continue;
        }
        double av = 0.0; int count = 0;                                               /* :1983-1989 */
        for (int k = s->my_start_bin; k <= s->my_end_bin; ++k) { av = av + s->histogram[k - 1]; count += 1; }
        av = av / (double)count;
        rr.mean = av; rr.max_pct = 100.0 * mx / av; rr.min_pct = 100.0 * mn / av;
        if (!(s->wl_invt_active || s->p.wl_swetnam)) {
            int flat = 1;
            if (fp->wl_schedule == 0) {
                for (int k = s->my_start_bin; k <= s->my_end_bin; ++k)
                    if (fabs(s->histogram[k - 1] - av) / av > fp->wl_flattol) flat = 0;
            } else if (fp->wl_schedule == 1) {
                double m2 = s->histogram[s->my_start_bin - 1];
                for (int k = s->my_start_bin; k <= s->my_end_bin; ++k) if (s->histogram[k - 1] < m2) m2 = s->histogram[k - 1];
                if (nint_d(m2) < fp->wl_minhist) flat = 0;
            } else if (fp->wl_schedule == 2) {
                for (int k = s->my_start_bin; k <= s->my_end_bin; ++k)
                    if (s->histogram[k - 1] < (1.0 - fp->wl_flattol) * av) flat = 0;
            } else { rc = 40; flat = 0; }                                             /* stop 'unknown wl_schedule' */
            if (!dd) { if (r == 0) flat0 = flat; else flat = flat0; }                 /* comms_bcastlog(flat): rank 0 decides */
            if (flat) {
                if (!dd) {
                    const double mid = s->weight[nb / 2];                             /* weight(nbins/2+1) */
                    for (int k = 0; k < nb; ++k) s->weight[k] = s->weight[k] - mid;
                    for (int k = 0; k < nb; ++k) { s->histogram[k] = 0.0; s->hist_last_sync[k] = 0.0; }
                } else {
                    for (int k = 0; k < nb; ++k) s->histogram[k] = 0.0;
                }
                s->wl_factor = s->wl_factor * 0.5;
                s->firstcycle = 0;
            }
            rr.flat = flat;
            const double wl_invt = (double)nb / (double)(s->mc_cycle_num * s->nwater); /* :2134-2142 */
            if (s->wl_factor < wl_invt && s->wl_factor > F_TINY) {
                if (fp->wl_useinvt) { s->wl_invt_active = 1; s->wl_factor = wl_invt; rr.invt_switched = 1; }
            }
        }
        rr.wl_factor = s->wl_factor;
        if (r == 0) r0 = rr;
    }
    free(skip);
    if (rep) *rep = r0;
    return rc;
}

/* comms_mpi.f90:299-375: rank 0 stitches the windows in rank order, scaling each new window so that
 * the mean log of the 2*overlap+1 bins around the seam agrees */
void orc_join_uhist(orc_system **w, int n, int overlap, double *joined)
{
    const int nb = w[0]->p.nbins;
    const int bpw = nb / n;
    for (int k = 0; k < nb; ++k) joined[k] = w[0]->unbiased_hist[k];
    for (int ir = 1; ir < n; ++ir) {
        const double *recv = w[ir]->unbiased_hist;
        const int my_end = ir * bpw;
        double myave = 0.0, nextav = 0.0;
        for (int k = my_end - overlap; k <= my_end + overlap; ++k) myave = myave + log(joined[k - 1]);
        myave = myave / (double)(2 * overlap + 1);
        for (int k = my_end - overlap; k <= my_end + overlap; ++k) nextav = nextav + log(recv[k - 1]);
        nextav = nextav / (double)(2 * overlap + 1);
        double shift = myave - nextav;
        if (isnan(shift)) shift = 0.0;
        for (int k = my_end + 1; k <= nb; ++k) joined[k - 1] = recv[k - 1] * exp(shift);
    }
}

/* comms_mpi.f90:377-459 */
void orc_join_eta(orc_system **w, int n, int overlap, double *joined)
{
    const int nb = w[0]->p.nbins;
    const int bpw = nb / n;
    for (int k = 0; k < nb; ++k) joined[k] = w[0]->weight[k];
    for (int ir = 1; ir < n; ++ir) {
        const double *recv = w[ir]->weight;
        const int my_end = ir * bpw;
        double myave = 0.0, nextav = 0.0;
        for (int k = my_end - overlap; k <= my_end + overlap; ++k) myave = myave + joined[k - 1];
        myave = myave / (double)(2 * overlap + 1);
        for (int k = my_end - overlap; k <= my_end + overlap; ++k) nextav = nextav + recv[k - 1];
        nextav = nextav / (double)(2 * overlap + 1);
        const double shift = myave - nextav;
        for (int k = my_end + 1; k <= nb; ++k) joined[k - 1] = recv[k - 1] + shift;
    }
    const double mid = joined[nb / 2];
    for (int k = 0; k < nb; ++k) joined[k] = joined[k] - mid;
}

/* mc_moves.F90:2498-2621 (called for sample runs only, :305) */
double orc_mc_deltaG_from_hist(orc_system **w, int n, double *normP)
{
    const int nb = w[0]->p.nbins;
    orc_system *s0 = w[0];
    double *joined = (double *)calloc(nb, sizeof(double));
    if (!s0->p.dd) {
        allreduce_one(w, n, nb, offsetof(orc_system, unbiased_hist), offsetof(orc_system, uhist_last_sync));
        for (int k = 0; k < nb; ++k) joined[k] = s0->unbiased_hist[k];
    } else {
        orc_join_uhist(w, n, s0->p.window_overlap, joined);
    }
    double Pnorm = 0.0;
    for (int i = 0; i < nb; ++i) Pnorm = Pnorm + joined[i] * s0->binwidth[i];
    double *np_ = normP ? normP : joined;
    for (int i = 0; i < nb; ++i) np_[i] = joined[i] / Pnorm;
    double pA = 0.0, pB = 0.0;
    for (int i = 0; i < nb / 2; ++i) pA = pA + np_[i] * s0->binwidth[i];
    for (int i = nb / 2; i < nb; ++i) pB = pB + np_[i] * s0->binwidth[i];
    double deltaG = log(pA / pB);
    const double beta = 1.0 / (kB * s0->p.temperature);
    if (s0->p.leshift) deltaG = deltaG + beta * s0->ref_enthalpy[1] - beta * s0->ref_enthalpy[0];
    free(joined);
    return deltaG;
}

/* ------------------------------------------------------------------ */
/* batch helpers: independent walkers over host threads (CPU baseline)  */
/* ------------------------------------------------------------------ */
#include <pthread.h>
#include <unistd.h>

typedef struct {
    orc_system **w; int n; int ncycles; double *e_out;
    int next; int err; int what;
    pthread_mutex_t mu;
} orc_job;

static void *orc_worker(void *arg)
{
    orc_job *j = (orc_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        if (j->what == 0) {
            const int e = orc_mc_run(j->w[i], j->ncycles);
            if (e) { pthread_mutex_lock(&j->mu); j->err |= e; pthread_mutex_unlock(&j->mu); }
        } else {
            for (int ils = 0; ils < j->w[i]->nlat; ++ils) {
                orc_compute_model_energy(j->w[i], ils);
                j->e_out[i * 2 + ils] = j->w[i]->model_energy[ils];
            }
        }
    }
    return NULL;
}

int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return (n < 1) ? 1 : (int)n;
}

static int orc_run_job(orc_job *j, int nthreads)
{
    if (nthreads < 1) nthreads = orc_max_threads();
    if (nthreads > j->n) nthreads = j->n;
    if (nthreads < 1) nthreads = 1;
    pthread_mutex_init(&j->mu, NULL);
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, orc_worker, j);
    orc_worker(j);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&j->mu);
    return j->err;
}

int orc_mc_run_many(orc_system **w, int n, int ncycles, int nthreads)
{
    orc_job j; memset(&j, 0, sizeof(j));
    j.w = w; j.n = n; j.ncycles = ncycles; j.what = 0;
    return orc_run_job(&j, nthreads);
}

void orc_model_energy_many(orc_system **w, int n, int nthreads, double *e_out)
{
    orc_job j; memset(&j, 0, sizeof(j));
    j.w = w; j.n = n; j.e_out = e_out; j.what = 1;
    orc_run_job(&j, nthreads);
}

/* ------------------------------------------------------------------ */
/* name-based accessors for the ctypes test harness                    */
/* ------------------------------------------------------------------ */
double *orc_ptr_d(orc_system *s, const char *name)
{
    if (!strcmp(name, "ljr")) return s->ljr;
    if (!strcmp(name, "ref_ljr")) return s->ref_ljr;
    if (!strcmp(name, "h")) return s->h;
    if (!strcmp(name, "ref_h")) return s->ref_h;
    if (!strcmp(name, "recip")) return s->recip;
    if (!strcmp(name, "volume")) return s->volume;
    if (!strcmp(name, "model_energy")) return s->model_energy;
    if (!strcmp(name, "ivect")) return s->ivect;
    if (!strcmp(name, "ref_enthalpy")) return s->ref_enthalpy;
    if (!strcmp(name, "average_energy")) return s->average_energy;
    if (!strcmp(name, "histogram")) return s->histogram;
    if (!strcmp(name, "weight")) return s->weight;
    if (!strcmp(name, "unbiased_hist")) return s->unbiased_hist;
    if (!strcmp(name, "mu_bin")) return s->mu_bin;
    if (!strcmp(name, "binwidth")) return s->binwidth;
    if (!strcmp(name, "eta_last_sync")) return s->eta_last_sync;
    if (!strcmp(name, "hist_last_sync")) return s->hist_last_sync;
    if (!strcmp(name, "uhist_last_sync")) return s->uhist_last_sync;
    return NULL;
}

int *orc_ptr_i(orc_system *s, const char *name)
{
    if (!strcmp(name, "nn")) return s->nn;
    if (!strcmp(name, "jn")) return s->jn;
    if (!strcmp(name, "vn")) return s->vn;
    if (!strcmp(name, "nivect")) return s->nivect;
    if (!strcmp(name, "mc_translations")) return s->mc_translations;
    return NULL;
}

double orc_get_d(const orc_system *s, const char *name)
{
    if (!strcmp(name, "ls_mu")) return s->ls_mu;
    if (!strcmp(name, "av_binwidth")) return s->av_binwidth;
    if (!strcmp(name, "log_unbiased_norm")) return s->log_unbiased_norm;
    if (!strcmp(name, "r_pos")) return s->r_pos;
    if (!strcmp(name, "r_neg")) return s->r_neg;
    if (!strcmp(name, "a_pos")) return s->a_pos;
    if (!strcmp(name, "a_neg")) return s->a_neg;
    if (!strcmp(name, "wl_factor")) return s->wl_factor;
    if (!strcmp(name, "orig_wl_factor")) return s->orig_wl_factor;
    if (!strcmp(name, "my_mu_min")) return s->my_mu_min;
    if (!strcmp(name, "my_mu_max")) return s->my_mu_max;
    if (!strcmp(name, "min_dmu")) return s->min_dmu;
    if (!strcmp(name, "max_dmu")) return s->max_dmu;
    if (!strcmp(name, "mc_max_trans")) return s->p.mc_max_trans;
    if (!strcmp(name, "mc_dv_max")) return s->p.mc_dv_max;
    if (!strcmp(name, "transP")) return s->transP;
    if (!strcmp(name, "volP")) return s->volP;
    if (!strcmp(name, "swP")) return s->swP;
    return NAN;
}

int64_t orc_get_i(const orc_system *s, const char *name)
{
    if (!strcmp(name, "ls")) return s->ls;
    if (!strcmp(name, "mc_cycle_num")) return s->mc_cycle_num;
    if (!strcmp(name, "acc_r")) return s->acc_r;
    if (!strcmp(name, "acc_v")) return s->acc_v;
    if (!strcmp(name, "acc_s")) return s->acc_s;
    if (!strcmp(name, "att_r")) return s->att_r;
    if (!strcmp(name, "att_v")) return s->att_v;
    if (!strcmp(name, "att_s")) return s->att_s;
    if (!strcmp(name, "my_start_bin")) return s->my_start_bin;
    if (!strcmp(name, "my_end_bin")) return s->my_end_bin;
    if (!strcmp(name, "walker_in_window")) return s->walker_in_window;
    if (!strcmp(name, "nbins")) return s->p.nbins;
    if (!strcmp(name, "error")) return s->error;
    if (!strcmp(name, "nn_warnings")) return s->nn_warnings;
    if (!strcmp(name, "rng_index")) return (s->rng.mode == 1) ? s->rng.fifo_pos : (int64_t)s->rng.index;
    if (!strcmp(name, "rng_fifo_pos")) return s->rng.fifo_pos;
    if (!strcmp(name, "wl_invt_active")) return s->wl_invt_active;
    if (!strcmp(name, "firstcycle")) return s->firstcycle;
    if (!strcmp(name, "histogram_reset")) return s->histogram_reset;
    return -1;
}

void orc_set_d(orc_system *s, const char *name, double v)
{
    if (!strcmp(name, "ls_mu")) s->ls_mu = v;
    else if (!strcmp(name, "wl_factor")) s->wl_factor = v;
    else if (!strcmp(name, "mc_max_trans")) s->p.mc_max_trans = v;
    else if (!strcmp(name, "mc_dv_max")) s->p.mc_dv_max = v;
}

void orc_set_i(orc_system *s, const char *name, int64_t v)
{
    if (!strcmp(name, "ls")) s->ls = (int)v;
    else if (!strcmp(name, "mc_cycle_num")) s->mc_cycle_num = (int)v;
    else if (!strcmp(name, "walker_in_window")) s->walker_in_window = (int)v;
    else if (!strcmp(name, "wl_invt_active")) s->wl_invt_active = (int)v;
    else if (!strcmp(name, "firstcycle")) s->firstcycle = (int)v;
    else if (!strcmp(name, "histogram_reset")) s->histogram_reset = (int)v;
}

void orc_set_rng_philox(orc_system *s, uint64_t seed, uint32_t stream, uint64_t start_index)
{
    orc_rng_philox(&s->rng, seed, stream, start_index);
}

void orc_set_rng_fifo(orc_system *s, const double *u, int64_t n)
{
    orc_rng_fifo(&s->rng, u, n);
}
