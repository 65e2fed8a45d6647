"""Second, independent restatement of the hot path -- plain Python / numpy written DIRECTLY from the Fortran of
keb721/mc_water_ls_mw, not from oracle/mw_oracle.c -- used to pin the C oracle (tests/test_oracle_numpy.py).

    compute_ivects              molint.F90:174-217
    compute_neighbours          molint.F90:501-559
    compute_model_energy        molint.F90:407-499
    compute_local_real_energy   molint.F90:220-404
    util_determinant / util_recipmatrix    util.f90:16-77
    bin grid, log_unbiased_norm, initial ls_mu     mc_moves.F90:557-656, :781-806, :857-862
    mu_to_bin / eta_weight      mc_moves.F90:2187-2215, :893-964
    mc_water_translation        mc_moves.F90:966-1213
    mc_volume                   mc_moves.F90:1216-1534
    mc_lattice_switch           mc_moves.F90:1536-1594
    mc_update_wl_bins           mc_moves.F90:1597-1689
    move loop of mc_cycle       mc_moves.F90:145-255 (incl. the 'dd' window check :181-208 and the switch ban :237,:244)
    'dd' windows of mc_init     mc_moves.F90:660-703, :808-812
    mc_monitor_stats (state effects)        mc_moves.F90:1722-1732, :1786-1810
    mc_check_chain_synchronisation          mc_moves.F90:2217-2416

Everything is IEEE double arithmetic in the reference's operation order (Python floats never contract into FMAs);
1-based indices are kept in the list arrays (jn, vn) as the reference stores them.  Random numbers come from a
caller-supplied FIFO in the reference's draw order.  Test infrastructure: no product code imports this.
"""
from __future__ import annotations

import math
import sys

import numpy as np

# ---- constants (constants.f90:23-24,39,43 ; molint.F90:64-74) ------------------------------------------------
PI = 3.141592653589793238462643383279502884197
INVPI = 1.0 / PI
KB = 1.0 / 3.1577465e5
ANG_TO_BOHR = 1.0 / 0.5291772108
MW_SIGMA = 2.3925 * ANG_TO_BOHR
MW_EPSILON = 6.189 / 627.509469
MW_LAMBDA = 23.15
SW_BIGA = 7.049556277
SW_B = 0.6022245584
SW_GAMMA = 1.2
SW_A = 1.8
COS0 = float(np.float32(-0.33331324756))      # a default-real literal in the reference (molint.F90:74)
MAXNEIGH = 50


def determinant(m):            # util.f90:16-41; m[i][j] = matrix(i+1, j+1)
    det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1])
    det = det - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
    det = det + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0])
    return det


def recipmatrix(h):            # util.f90:43-77
    r = [[0.0] * 3 for _ in range(3)]
    r[0][0] = h[1][1] * h[2][2] - h[1][2] * h[2][1]
    r[0][1] = h[1][2] * h[2][0] - h[1][0] * h[2][2]
    r[0][2] = h[1][0] * h[2][1] - h[1][1] * h[2][0]
    r[1][0] = h[0][2] * h[2][1] - h[0][1] * h[2][2]
    r[1][1] = h[0][0] * h[2][2] - h[0][2] * h[2][0]
    r[1][2] = h[0][1] * h[2][0] - h[0][0] * h[2][1]
    r[2][0] = h[0][1] * h[1][2] - h[0][2] * h[1][1]
    r[2][1] = h[0][2] * h[1][0] - h[0][0] * h[1][2]
    r[2][2] = h[0][0] * h[1][1] - h[0][1] * h[1][0]
    vol = h[0][0] * r[0][0] + h[0][1] * r[0][1] + h[0][2] * r[0][2]
    return [[r[i][j] * 2.0 * PI / vol for j in range(3)] for i in range(3)]


class Box:
    """The module-level state of `model`, `energy` and `mc_moves` for one rank."""

    def __init__(self, up, hflat, ljr, weights=None, file_wl_factor=0.0, rank=0, size=1):
        """up: decks.UserParams (internal units); hflat[nlat][9] column-major hmatrix as the xmol reader stores it;
        ljr[nlat][N][3]."""
        self.up = up
        self.N = up.nwater
        self.nlat = up.num_lattices
        # hmatrix(i,j,ils) -> self.h[ils][i][j]; the flat xmol order is h(1,1),h(2,1),h(3,1),h(1,2)... (init.f90:80-106)
        self.h = [[[float(hflat[l][j * 3 + i]) for j in range(3)] for i in range(3)] for l in range(self.nlat)]
        self.r = [[[float(x) for x in ljr[l][i]] for i in range(self.N)] for l in range(self.nlat)]
        self.ref = [[list(p) for p in lat] for lat in self.r]                    # init.f90:103
        self.ref_h = [[row[:] for row in self.h[l]] for l in range(self.nlat)]   # init.f90:90
        self.recip = [recipmatrix(self.h[l]) for l in range(self.nlat)]          # init.f90:90
        self.volume = [0.0] * self.nlat
        self.model_energy = [0.0] * self.nlat
        self.nivect = [0] * self.nlat
        self.ivect = [None] * self.nlat
        self.nn = [[0] * self.N for _ in range(self.nlat)]
        self.jn = [[[0] * MAXNEIGH for _ in range(self.N)] for _ in range(self.nlat)]
        self.vn = [[[0] * MAXNEIGH for _ in range(self.N)] for _ in range(self.nlat)]
        self.fifo = None
        self.fpos = 0
        self.energy_init()
        self.mc_init(weights, file_wl_factor, rank, size)

    # ---- random numbers -------------------------------------------------------------------------------------
    def set_fifo(self, u):
        self.fifo = [float(x) for x in u]
        self.fpos = 0

    def rnd(self):
        x = self.fifo[self.fpos]
        self.fpos += 1
        return x

    # ---- module energy --------------------------------------------------------------------------------------
    def energy_init(self):                                       # molint.F90:91-153
        for l in range(self.nlat):
            self.volume[l] = abs(determinant(self.h[l]))
            self.compute_neighbours(l)
            self.compute_model_energy(l)

    def compute_ivects(self, l):                                 # molint.F90:174-217
        h = self.h[l]
        col = lambda j: [h[0][j], h[1][j], h[2][j]]
        dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
        im = math.floor(SW_A * MW_SIGMA / math.sqrt(dot(col(0), col(0)))) + 1
        jm = math.floor(SW_A * MW_SIGMA / math.sqrt(dot(col(1), col(1)))) + 1
        km = math.floor(SW_A * MW_SIGMA / math.sqrt(dot(col(2), col(2)))) + 1
        self.nivect[l] = (2 * im + 1) * (2 * jm + 1) * (2 * km + 1)
        iv = [[0.0, 0.0, 0.0]]
        for ic in range(-im, im + 1):
            sx = [float(ic) * c for c in col(0)]
            for jc in range(-jm, jm + 1):
                sy = [float(jc) * c for c in col(1)]
                for kc in range(-km, km + 1):
                    sz = [float(kc) * c for c in col(2)]
                    if abs(ic) + abs(jc) + abs(kc) == 0:
                        continue
                    iv.append([(sx[d] + sy[d]) + sz[d] for d in range(3)])
        self.ivect[l] = iv

    def compute_neighbours(self, l):                             # molint.F90:501-559
        rn = SW_A * MW_SIGMA * 1.18
        self.compute_ivects(l)
        iv = np.array(self.ivect[l])                             # [nivect][3]
        R = self.r[l]
        for i in range(self.N):
            ilj = R[i]
            n = 0
            for j in range(self.N):
                v = [R[j][d] - ilj[d] for d in range(3)]
                # tmpvect = v_ij + ivect(:,k); r2 = dot_product: ((t1*t1 + t2*t2) + t3*t3), elementwise over k
                t0 = v[0] + iv[:, 0]; t1 = v[1] + iv[:, 1]; t2 = v[2] + iv[:, 2]
                r2 = (t0 * t0 + t1 * t1) + t2 * t2
                for k in np.nonzero(r2 < rn * rn)[0]:
                    if k == 0 and j == i:
                        continue
                    n += 1
                    self.jn[l][i][n - 1] = j + 1
                    self.vn[l][i][n - 1] = int(k) + 1
            self.nn[l][i] = n

    def compute_model_energy(self, l):                           # molint.F90:407-499
        R, iv = self.r[l], self.ivect[l]
        rcsq = MW_SIGMA * SW_A * MW_SIGMA * SW_A
        E = 0.0
        for i in range(self.N):
            ilj = R[i]
            for ln in range(self.nn[l][i]):
                j = self.jn[l][i][ln] - 1; ji = self.vn[l][i][ln] - 1
                t = [(R[j][d] + iv[ji][d]) - ilj[d] for d in range(3)]
                r2 = t[0] ** 2 + t[1] ** 2 + t[2] ** 2
                if r2 < rcsq:
                    r1 = math.sqrt(r2)
                    e2 = math.exp(MW_SIGMA / (r1 - MW_SIGMA * SW_A))
                    tmpE = SW_BIGA * MW_EPSILON * (SW_B * (MW_SIGMA * MW_SIGMA / r2) ** 2 - 1.0)
                    tmpE = tmpE * e2
                    e2 = math.exp(SW_GAMMA * MW_SIGMA / (r1 - MW_SIGMA * SW_A))
                    E = E + 0.5 * tmpE
                    for ln2 in range(ln + 1, self.nn[l][i]):
                        k = self.jn[l][i][ln2] - 1; ki = self.vn[l][i][ln2] - 1
                        t2 = [(R[k][d] + iv[ki][d]) - ilj[d] for d in range(3)]
                        r2k = t2[0] ** 2 + t2[1] ** 2 + t2[2] ** 2
                        if r2k < rcsq:
                            r1k = math.sqrt(r2k)
                            ct = (t[0] * t2[0] + t[1] * t2[1] + t[2] * t2[2]) / (r1k * r1)
                            csq = (ct - COS0) ** 2
                            e1 = math.exp(SW_GAMMA * MW_SIGMA / (r1k - MW_SIGMA * SW_A))
                            E = E + MW_LAMBDA * MW_EPSILON * e1 * e2 * csq
        self.model_energy[l] = E

    def compute_local_real_energy(self, imol, l):                # molint.F90:220-404 (imol 0-based here)
        R, iv = self.r[l], self.ivect[l]
        rcsq = MW_SIGMA * SW_A * MW_SIGMA * SW_A
        Evdw = 0.0
        Etb = 0.0
        ilj = R[imol]
        nni = self.nn[l][imol]
        for ln in range(nni):
            iEtb = 0.0
            j = self.jn[l][imol][ln] - 1; ji = self.vn[l][imol][ln] - 1
            jv = iv[ji]
            jlj = [R[j][d] + jv[d] for d in range(3)]
            t = [jlj[d] - ilj[d] for d in range(3)]
            r2 = t[0] ** 2 + t[1] ** 2 + t[2] ** 2
            if r2 < rcsq:
                ir1 = 1.0 / math.sqrt(r2)
                r1 = ir1 * r2
                isr1 = 1.0 / (r1 - MW_SIGMA * SW_A)
                exp2 = math.exp(MW_SIGMA * isr1)
                exp3 = math.exp(SW_GAMMA * MW_SIGMA * isr1)
                tmpE = SW_BIGA * MW_EPSILON * (SW_B * (MW_SIGMA * MW_SIGMA * ir1 * ir1) ** 2 - 1.0)
                tmpE = tmpE * exp2
                Evdw = Evdw + tmpE
                sq, ctl = [], []
                for ln2 in range(ln + 1, nni):                   # j--i--k
                    k = self.jn[l][imol][ln2] - 1; ki = self.vn[l][imol][ln2] - 1
                    klj = [R[k][d] + iv[ki][d] for d in range(3)]
                    t2 = [klj[d] - ilj[d] for d in range(3)]
                    sq.append(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2])
                    ctl.append((t[0] * t2[0] + t[1] * t2[1] + t[2] * t2[2]) * ir1)
                tm = [-x for x in t]                              # now vector from j to i
                for ln2 in range(self.nn[l][j]):                  # i--j--k
                    k = self.jn[l][j][ln2] - 1; ki = self.vn[l][j][ln2] - 1
                    klj = [(R[k][d] + iv[ki][d]) + jv[d] for d in range(3)]
                    t2 = [klj[d] - jlj[d] for d in range(3)]
                    sq.append(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2])
                    ctl.append((tm[0] * t2[0] + tm[1] * t2[1] + tm[2] * t2[2]) * ir1)
                for q, c in zip(sq, ctl):
                    if q < rcsq:
                        vinv = 1.0 / math.sqrt(q)
                        vexp = vinv * q - MW_SIGMA * SW_A
                        vexp = SW_GAMMA * MW_SIGMA / vexp
                        c = c * vinv
                        pref = (c - COS0) ** 2 if c < 0.99 else 0.0
                        iEtb = iEtb + pref * math.exp(vexp)
                    # out-of-range entries carry preflist = 0 (:382): nothing to add
                iEtb = iEtb * exp3
            Etb = Etb + iEtb
        return Evdw + MW_LAMBDA * MW_EPSILON * Etb

    # ---- mc_init: grid, weights, norm, ls_mu ----------------------------------------------------------------
    def mc_init(self, weights, file_wl_factor, rank, size):
        up = self.up
        nb = up.nbins + (1 if up.nbins % 2 == 0 else 0)          # :557
        self.nbins = nb
        s_pos = abs(up.mu_max) - 0.5; s_neg = abs(up.mu_min) - 0.5
        a_pos = a_neg = 1.0
        Ns = nb // 2

        def ratio(a, s):                                          # :584-594
            r = 1.1; k = 0
            while True:
                k += 1
                p = 1.0
                for _ in range(Ns):                               # r**Ns, integer power by repeated multiplication
                    p = p * r
                tmpsum = a * (1.0 - p) / (1.0 - r)
                r_new = r * (s / tmpsum) ** (1.0 / float(Ns))
                if abs(r_new - r) <= 2.0 * sys.float_info.epsilon or k > 1000000:
                    return r
                r = r_new

        self.r_pos, self.r_neg = ratio(a_pos, s_pos), ratio(a_neg, s_neg)
        self.a_pos, self.a_neg = a_pos, a_neg
        mu_bin = [0.0] * nb; bw = [0.0] * nb

        def ipow(x, n):
            p = 1.0
            for _ in range(n):
                p = p * x
            return p

        mu_u = -0.5; k = 0
        for ibin in range(nb // 2, 0, -1):                        # :604-613
            mu_l = mu_u - a_neg * ipow(self.r_neg, k)
            mu_bin[ibin - 1] = 0.5 * (mu_u + mu_l); bw[ibin - 1] = mu_u - mu_l
            mu_u = mu_l; k += 1
        mu_bin[nb // 2] = 0.0; bw[nb // 2] = 1.0
        mu_l = 0.5; k = 0
        for ibin in range(nb // 2 + 2, nb + 1):                   # :623-632
            mu_u = mu_l + a_pos * ipow(self.r_pos, k)
            mu_bin[ibin - 1] = 0.5 * (mu_u + mu_l); bw[ibin - 1] = mu_u - mu_l
            mu_l = mu_u; k += 1
        av = 0.0
        for x in bw:
            av = av + x
        self.av_binwidth = av / float(nb)
        self.mu_bin, self.binwidth = mu_bin, bw
        self.weight = [0.0] * nb; self.histogram = [0.0] * nb; self.unbiased_hist = [0.0] * nb
        self.wl_factor = up.wl_factor; self.orig_wl_factor = up.wl_factor
        self.log_unbiased_norm = 0.0
        if self.nlat == 2:
            if weights is not None:                               # :751-766
                if file_wl_factor > float(np.float32(1e-10)):
                    self.wl_factor = min(self.wl_factor, file_wl_factor)
                    if up.samplerun:
                        self.wl_factor = 0.0
                for i in range(min(nb, len(weights))):
                    self.weight[i] = float(weights[i])
            hits = float(up.max_mc_cycles) - float(up.eq_mc_cycles)   # :781-806
            hits = hits * float(size * self.N) / float(nb)
            incr = hits * self.av_binwidth
            lun = math.log(incr) + self.weight[0]
            for k in range(1, nb):
                incr = hits * self.av_binwidth
                if lun > self.weight[k] + math.log(incr):
                    lun = lun + math.log(1.0 + incr * math.exp(self.weight[k] - lun))
                else:
                    lun = math.log(incr) + self.weight[k] + math.log(1.0 + math.exp(lun - self.weight[k]) / incr)
            self.log_unbiased_norm = lun
        self.ls = up.ls
        self.dd = (getattr(up, "parallel_strategy", "mw") == "dd")
        if self.dd:                                               # :660-703 (1-based bins, sums left to right)
            def bsum(n):
                t = 0.0
                for i in range(n):
                    t = t + bw[i]
                return t
            ov = up.window_overlap
            bpw = nb // size
            if rank == 0:
                self.my_start_bin = 1
                self.my_end_bin = bpw + ov
                self.my_mu_min = up.mu_min
                self.my_mu_max = up.mu_min + bsum(self.my_end_bin)
            if size > 1:
                if 1 <= rank <= size - 2:
                    self.my_start_bin = rank * bpw - ov
                    self.my_end_bin = (rank + 1) * bpw + ov
                    self.my_mu_min = up.mu_min + bsum(self.my_start_bin - 1)
                    self.my_mu_max = up.mu_min + bsum(self.my_end_bin)
                if rank == size - 1:
                    self.my_start_bin = rank * bpw - ov
                    self.my_end_bin = nb
                    self.my_mu_min = up.mu_min + bsum(self.my_start_bin - 1)
                    self.my_mu_max = up.mu_max
            if self.my_mu_max < 0.0: self.ls = 1
            if self.my_mu_min > 0.0: self.ls = 2
            if self.nlat == 2:                                    # :808-812
                for i in range(0, self.my_start_bin - 1):
                    self.weight[i] = 0.0
                for i in range(self.my_end_bin, nb):
                    self.weight[i] = 0.0
            self.walker_in_window = False                         # :112, set by the first cycles (:181-208)
        else:
            # 'mw' strategy: the whole range is this rank's window (:712-718, :872)
            self.my_start_bin, self.my_end_bin = 1, nb
            self.my_mu_min, self.my_mu_max = up.mu_min, up.mu_max
            self.walker_in_window = True
        beta = 1.0 / (KB * up.temperature)
        self.ref_enthalpy = [self.model_energy[l] + (up.pressure * self.volume[l] if up.mc_ensemble == "npt" else 0.0)
                             for l in range(self.nlat)]            # main.f90:146-150
        self.ls_mu = 0.0
        if self.nlat == 2:                                        # :857-862
            mu = self.model_energy[0] + up.pressure * self.volume[0] - self.model_energy[1] - up.pressure * self.volume[1]
            if up.leshift:
                mu = mu - self.ref_enthalpy[0] + self.ref_enthalpy[1]
            self.ls_mu = mu * beta - float(self.N) * math.log(self.volume[0] / self.volume[1])
        self.mc_cycle_num = 0
        self.mc_max_trans, self.mc_dv_max = up.mc_max_trans, up.mc_dv_max
        self.acc = [0, 0, 0]; self.att = [0, 0, 0]
        self.mc_translations = [0] * self.N
        self.average_energy = [0.0, 0.0]
        self.min_dmu = sys.float_info.max; self.max_dmu = 0.0
        self.sumhist = 0.0; self.wl_invt_active = False
        self.trace = []                                           # (move type, accepted) per trial move
        # cumulative move probabilities (:153-176)
        sw, vp, tp = up.mc_switch_prob, up.mc_vol_prob, up.mc_trans_prob
        if up.mc_always_switch: sw = 0.0
        if not up.allow_switch: sw = 0.0
        if up.mc_ensemble == "nvt": vp = 0.0
        if not up.allow_vol: vp = 0.0
        if not up.allow_trans: tp = 0.0
        sp = tp + vp + sw
        self.transP = tp / sp; self.volP = vp / sp; self.swP = sw / sp
        self.volP = self.volP + self.transP; self.swP = self.swP + self.volP

    # ---- order parameter / weights --------------------------------------------------------------------------
    def mu_to_bin(self, mu):                                      # :2187-2215 (1-based)
        nb = self.nbins
        if abs(mu) <= 0.5:
            return nb // 2 + 1
        if mu > 0.0:
            arg = 1.0 - (mu - 0.5) * (1.0 - self.r_pos) / self.a_pos
            return nb // 2 + 2 + int(math.log(arg) / math.log(self.r_pos))
        arg = 1.0 - (abs(mu) - 0.5) * (1.0 - self.r_neg) / self.a_neg
        return nb // 2 - int(math.log(arg) / math.log(self.r_neg))

    def eta_weight(self, mu):                                     # :893-964
        if not self.walker_in_window:
            return 0.0
        if mu < self.my_mu_min or mu > self.my_mu_max:
            return sys.float_info.max
        k = self.mu_to_bin(mu)
        w = lambda q: self.weight[q - 1]; bw = lambda q: self.binwidth[q - 1]; mb = lambda q: self.mu_bin[q - 1]
        if not self.up.eta_interp:
            return w(k)
        if k == self.my_start_bin:
            g = 2.0 * (w(k + 1) - w(k)) / (bw(k) + bw(k + 1))
            return w(k) + (mu - mb(k)) * g
        if k == self.my_end_bin:
            g = 2.0 * (w(k) - w(k - 1)) / (bw(k) + bw(k - 1))
            return w(k) + (mu - mb(k)) * g
        if mu > mb(k):
            g = 2.0 * (w(k + 1) - w(k)) / (bw(k) + bw(k + 1))
            return w(k) + (mu - mb(k)) * g
        g = 2.0 * (w(k) - w(k - 1)) / (bw(k) + bw(k - 1))
        return w(k - 1) + (mu - mb(k - 1)) * g

    def _mu_from_scratch(self, beta):                             # :1370-1372 / :1525-1527 / :1583-1585
        up = self.up
        mu = (self.model_energy[0] + up.pressure * self.volume[0]) - (self.model_energy[1] + up.pressure * self.volume[1])
        if up.leshift:
            mu = mu - self.ref_enthalpy[0] + self.ref_enthalpy[1]
        return mu * beta - float(self.N) * math.log(self.volume[0] / self.volume[1])

    # ---- moves ----------------------------------------------------------------------------------------------
    def mc_water_translation(self):                               # :966-1213
        up = self.up
        ls = self.ls - 1; lsn = 1 - ls
        beta = 1.0 / (KB * up.temperature)
        x = self.rnd()
        imol = min(int(x * float(self.N)) + 1, self.N) - 1
        self.mc_translations[imol] += 1
        old = [0.0, 0.0]; backup = [0.0, 0.0]
        for l in range(self.nlat):
            old[l] = self.compute_local_real_energy(imol, l)
            backup[l] = self.model_energy[l]
            self.model_energy[l] = self.model_energy[l] - old[l]
        x = self.rnd(); y = self.rnd(); z = self.rnd()
        x = 2.0 * x - 1.0; y = 2.0 * y - 1.0; z = 2.0 * z - 1.0
        norm = 1.0 / math.sqrt(x * x + y * y + z * z)
        x = x * norm; y = y * norm; z = z * norm
        r = self.rnd() * 2.0 - 1.0
        x = x * self.mc_max_trans * r; y = y * self.mc_max_trans * r; z = z * self.mc_max_trans * r
        rm = self.recip[ls]
        sx = rm[0][0] * x + rm[1][0] * y + rm[2][0] * z
        sy = rm[0][1] * x + rm[1][1] * y + rm[2][1] * z
        sz = rm[0][2] * x + rm[1][2] * y + rm[2][2] * z
        sx = sx * 0.5 * INVPI; sy = sy * 0.5 * INVPI; sz = sz * 0.5 * INVPI
        tv = [[0.0] * 3, [0.0] * 3]
        tv[ls] = [x, y, z]
        if self.nlat == 2:
            hm = self.h[lsn]
            tv[lsn] = [hm[d][0] * sx + hm[d][1] * sy + hm[d][2] * sz for d in range(3)]
        dE = [0.0, 0.0]
        for l in range(self.nlat):
            for d in range(3):
                self.r[l][imol][d] = self.r[l][imol][d] + tv[l][d]
            new = self.compute_local_real_energy(imol, l)
            self.model_energy[l] = self.model_energy[l] + new
            dE[l] = new - old[l]
        if self.nlat == 1:
            diffkT = beta * dE[0]
        else:
            eta_old = self.eta_weight(self.ls_mu)
            self.ls_mu = self.ls_mu + (dE[0] - dE[1]) * beta
            eta_new = self.eta_weight(self.ls_mu)
            diffkT = dE[ls] * beta + eta_new - eta_old
        zeta = self.rnd()
        if zeta < min(1.0, _exp(-diffkT)):
            self.acc[0] += 1
            dmu = abs(dE[0] - dE[1]) * beta
            if dmu < self.min_dmu: self.min_dmu = dmu
            if dmu > self.max_dmu: self.max_dmu = dmu
            return True
        for l in range(self.nlat):
            for d in range(3):
                self.r[l][imol][d] = self.r[l][imol][d] - tv[l][d]
            self.model_energy[l] = backup[l]
        if self.nlat == 2:
            self.ls_mu = self.ls_mu - (dE[0] - dE[1]) * beta
        return False

    def _rescale(self, arr, rm, hm):                              # :1290-1315 and its copies
        for p in arr:
            o = list(p)
            n0 = rm[0][0] * o[0] + rm[1][0] * o[1] + rm[2][0] * o[2]
            n1 = rm[0][1] * o[0] + rm[1][1] * o[1] + rm[2][1] * o[2]
            n2 = rm[0][2] * o[0] + rm[1][2] * o[1] + rm[2][2] * o[2]
            n0 = n0 * 0.5 * INVPI; n1 = n1 * 0.5 * INVPI; n2 = n2 * 0.5 * INVPI
            t = [hm[d][0] * n0 + hm[d][1] * n1 + hm[d][2] * n2 for d in range(3)]
            t = [t[d] - o[d] for d in range(3)]
            for d in range(3):
                p[d] = p[d] + t[d]

    def mc_volume(self):                                          # :1216-1534
        up = self.up
        beta = 1.0 / (KB * up.temperature)
        ls = self.ls - 1
        nl = self.nlat
        backup = list(self.model_energy); oldE = list(self.model_energy)
        for l in range(nl):
            self.recip[l] = recipmatrix(self.h[l])
        old_h = [[row[:] for row in self.h[l]] for l in range(nl)]
        old_recip = [[row[:] for row in self.recip[l]] for l in range(nl)]
        old_vol = list(self.volume)
        x = self.rnd(); idim = int(x * 3.0)
        x = self.rnd(); jdim = int(x * 3.0)
        x = self.rnd()
        dh = (2.0 * x - 1.0) * self.mc_dv_max
        for l in range(nl):
            delta = [[0.0] * 3 for _ in range(3)]
            delta[idim][jdim] = dh; delta[jdim][idim] = delta[idim][jdim]
            self.h[l] = [[self.h[l][i][j] + delta[i][j] for j in range(3)] for i in range(3)]
        newE = [0.0, 0.0]
        for l in range(nl):
            self._rescale(self.r[l], self.recip[l], self.h[l])
            self._rescale(self.ref[l], self.recip[l], self.h[l])
            self.volume[l] = abs(determinant(self.h[l]))
            self.recip[l] = recipmatrix(self.h[l])
            self.compute_ivects(l)
            self.compute_model_energy(l)
            newE[l] = self.model_energy[l]
        dE = [newE[l] - oldE[l] for l in range(nl)] + [0.0]
        old_eta = new_eta = 0.0; old_mu = 0.0
        if nl == 2:
            old_eta = self.eta_weight(self.ls_mu)
            old_mu = self.ls_mu
            self.ls_mu = self._mu_from_scratch(beta)
            new_eta = self.eta_weight(self.ls_mu)
        x = self.rnd()
        diffkT = beta * dE[ls] + new_eta - old_eta + beta * up.pressure * (self.volume[ls] - old_vol[ls]) \
            - float(self.N) * math.log(self.volume[ls] / old_vol[ls])
        compare = min(1.0, _exp(-diffkT))
        if x < compare:
            self.acc[1] += 1
            if nl == 2:
                dmu = abs(old_mu - self.ls_mu)
                if dmu < self.min_dmu: self.min_dmu = dmu
                if dmu > self.max_dmu: self.max_dmu = dmu
            return True
        for l in range(nl):
            self.volume[l] = old_vol[l]
            self.h[l] = [row[:] for row in old_h[l]]
        for l in range(nl):                                       # recip is still the TRIAL cell's here (:1443-1507)
            self._rescale(self.r[l], self.recip[l], self.h[l])
            self._rescale(self.ref[l], self.recip[l], self.h[l])
        for l in range(nl):
            self.recip[l] = [row[:] for row in old_recip[l]]
        for l in range(nl):
            self.compute_ivects(l)
        for l in range(nl):
            self.model_energy[l] = backup[l]
        if nl == 2:
            self.ls_mu = self._mu_from_scratch(beta)
        return False

    def mc_lattice_switch(self):                                  # :1536-1594
        up = self.up
        beta = 1.0 / (KB * up.temperature)
        ls = self.ls - 1; lsn = 1 - ls
        old_eta = self.eta_weight(self.ls_mu); new_eta = self.eta_weight(self.ls_mu)
        E, V = self.model_energy, self.volume
        if up.mc_ensemble == "npt":
            diffkT = beta * E[lsn] - beta * E[ls] + beta * up.pressure * (V[lsn] - V[ls]) \
                - float(self.N) * math.log(V[lsn] / V[ls]) + new_eta - old_eta
        else:
            diffkT = beta * E[lsn] - beta * E[ls] + new_eta - old_eta
        if up.leshift:
            diffkT = diffkT - beta * self.ref_enthalpy[lsn] + beta * self.ref_enthalpy[ls]
        compare = min(1.0, _exp(-diffkT))
        x = self.rnd()
        if x < compare:
            self.acc[2] += 1
            self.ls_mu = self._mu_from_scratch(beta)
            self.ls = lsn + 1
            return True
        return False

    def mc_update_wl_bins(self):                                  # :1597-1689
        up = self.up
        if self.mc_cycle_num < up.eq_mc_cycles:
            return
        k = self.mu_to_bin(self.ls_mu)
        if k < 1 or k > self.nbins:
            return
        self.histogram[k - 1] = self.histogram[k - 1] + self.av_binwidth / self.binwidth[k - 1]
        if up.samplerun:
            incr = self.av_binwidth / self.binwidth[k - 1]
            self.unbiased_hist[k - 1] = self.unbiased_hist[k - 1] + incr * math.exp(self.eta_weight(self.ls_mu) - self.log_unbiased_norm)
            return
        if up.wl_swetnam:
            self.sumhist = self.sumhist + 1.0
            f = 0.0
            for i in range(self.nbins):
                binfrac = self.binwidth[i] / (up.mu_max - up.mu_min - 1.0)
                f = f + (self.histogram[i] * self.binwidth[i] / self.sumhist - binfrac) ** 2
            f = math.sqrt(f / float(self.nbins))
            f = math.log(f)
            f = f * up.wl_alpha * float(self.nbins)
            self.wl_factor = min(f, self.orig_wl_factor)
        elif self.wl_invt_active:
            self.wl_factor = min(self.wl_factor, float(self.nbins) / float(self.mc_cycle_num * self.N))
        incr = self.wl_factor
        self.weight[k - 1] = self.weight[k - 1] + self.av_binwidth * incr / self.binwidth[k - 1]
        minbin = min(self.weight[self.my_start_bin - 1:self.my_end_bin])
        for i in range(self.my_start_bin - 1, self.my_end_bin):
            self.weight[i] = self.weight[i] - minbin

    def mc_cycle(self):                                           # :145-255
        up = self.up
        self.mc_cycle_num += 1
        if self.dd:                                               # :181-208
            if self.mc_cycle_num < up.eq_mc_cycles:
                self.walker_in_window = (self.ls_mu > self.my_mu_min) and (self.ls_mu < self.my_mu_max)
            elif self.mc_cycle_num == up.eq_mc_cycles:
                if not self.walker_in_window:
                    raise RuntimeError("Error : Not all walkers have reached their designated window")
            else:
                self.walker_in_window = True
        no_switch = self.dd and self.mc_cycle_num < up.eq_mc_cycles   # :237, :244
        if self.mc_cycle_num % up.list_update_int == 0:
            for l in range(self.nlat):
                self.compute_neighbours(l)
        for _ in range(self.N):
            xi = self.rnd()
            if xi < self.transP:
                ok = self.mc_water_translation()
                self.mc_update_wl_bins()
                self.att[0] += 1
                self.trace.append((0, int(ok)))
            elif xi < self.volP:
                ok = self.mc_volume()
                self.mc_update_wl_bins()
                self.att[1] += 1
                self.trace.append((1, int(ok)))
            elif xi < self.swP:
                if not no_switch:
                    ok = self.mc_lattice_switch()
                    self.att[2] += 1
                    self.trace.append((2, int(ok)))
            if up.mc_always_switch and self.nlat == 2:
                if not no_switch:
                    ok = self.mc_lattice_switch()
                    self.att[2] += 1
                    self.trace.append((3, int(ok)))
        for l in range(self.nlat):
            self.average_energy[l] = self.average_energy[l] + self.model_energy[l]
        if up.mc_ensemble == "npt":
            for l in range(self.nlat):
                self.average_energy[l] = self.average_energy[l] + up.pressure * self.volume[l]

    # ---- periodic bookkeeping that changes the walker's state -------------------------------------------------
    def mc_monitor_stats(self):                                   # :1722-1732, :1786-1810 (state effects only)
        up = self.up
        if up.eq_adjust_mc and self.mc_cycle_num < up.eq_mc_cycles:
            # a ratio of 0 / 0 attempts is NaN in the reference; max(NaN, x) is compiler-defined -- not exercised here
            atr = float(self.acc[0]) / float(self.att[0])
            avr = float(self.acc[1]) / float(self.att[1])
            self.mc_max_trans = max(self.mc_max_trans * atr / up.mc_target_ratio, 0.1)
            self.mc_dv_max = max(self.mc_dv_max * avr / up.mc_target_ratio, 0.0001)
        for l in range(self.nlat):                                # :1786-1792: the stored energies are replaced
            self.compute_model_energy(l)
        self.acc = [0, 0, 0]; self.att = [0, 0, 0]
        self.mc_translations = [0] * self.N
        self.average_energy = [0.0, 0.0]
        self.max_dmu = 0.0; self.min_dmu = sys.float_info.max

    def mc_check_chain_synchronisation(self):                     # :2217-2416
        up = self.up
        beta = 1.0 / (KB * up.temperature)
        self.compute_model_energy(0); self.compute_model_energy(1)
        hd = [[self.h[0][i][j] - self.ref_h[0][i][j] for j in range(3)] for i in range(3)]       # :2262
        self.h[1] = [[self.ref_h[1][i][j] + hd[i][j] for j in range(3)] for i in range(3)]       # :2277
        self.recip[0] = recipmatrix(self.h[0]); self.recip[1] = recipmatrix(self.h[1])

        def scaled(rm, v):                                        # :2294-2306
            t = [rm[0][c] * v[0] + rm[1][c] * v[1] + rm[2][c] * v[2] for c in range(3)]
            return [t[c] * 0.5 * INVPI for c in range(3)]

        for i in range(self.N):
            sv = [scaled(self.recip[l], self.r[l][i]) for l in range(2)]
            rv = [scaled(self.recip[l], self.ref[l][i]) for l in range(2)]
            d1 = [sv[0][c] - rv[0][c] for c in range(3)]
            s2 = [rv[1][c] + d1[c] for c in range(3)]             # :2331
            hm = self.h[1]
            self.r[1][i] = [hm[d][0] * s2[0] + hm[d][1] * s2[1] + hm[d][2] * s2[2] for d in range(3)]   # matmul, :2332
        for l in range(2):
            self.volume[l] = abs(determinant(self.h[l]))
            self.compute_ivects(l)
        self.compute_model_energy(0); self.compute_model_energy(1)
        # :2409-2411 -- written without the parentheses of :1370-1372: left to right
        mu = self.model_energy[0] + up.pressure * self.volume[0] - self.model_energy[1] - up.pressure * self.volume[1]
        if up.leshift:
            mu = mu - self.ref_enthalpy[0] + self.ref_enthalpy[1]
        self.ls_mu = mu * beta - float(self.N) * math.log(self.volume[0] / self.volume[1])

    # ---- views for the fixtures -------------------------------------------------------------------------------
    def hflat(self):
        return np.array([[self.h[l][i][j] for j in range(3) for i in range(3)] for l in range(self.nlat)])


def _exp(x):
    """exp() that overflows to +inf like the Fortran intrinsic instead of raising."""
    try:
        return math.exp(x)
    except OverflowError:
        return float("inf")
