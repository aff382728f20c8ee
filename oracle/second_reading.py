"""A SECOND, independently written reading of the reference's column step
(TEST INFRASTRUCTURE ONLY, like everything under oracle/).

Why: oracle/mckpp_oracle.c is a literal C transcription of the Fortran and the
reference cannot be compiled here, so nothing but the three EOS check values pins it to the
reference.  This module restates the path from the Fortran a second time, in a
deliberately different structure (numpy, whole profiles at a time):

* ``bldepth``   src/mckpp_physics_verticalmixing_bldepth_mod.F90:32-203
    per-level quantities for ALL levels at once (the Fortran and the C oracle walk level by
    level with a two-slot ka/ku rotation), then one running maximum for Rib and a first-hit search;
* ``wscale``    src/mckpp_physics_verticalmixing_wscale_mod.F90:12-97   (array-at-a-time)
* ``ocnstep``   src/mckpp_physics_ocnstep_mod.F90:43-357
    whole-profile expressions and an explicit decision function instead of DO/GOTO 45;
* ``physics_driver``  src/mckpp_physics_driver_mod.F90:15-73 (column loop + bottomtemp)
* ``ocnint2`` + ``tridcof``/``tridrhs``/``tridmat``/``rhsmod_salt``
    src/mckpp_physics_ocnint_mod.F90:19-221, src/mckpp_physics_solvers.F90:14-335
    coefficient and right-hand-side arrays at a time, the recurrences as scalar loops;
* ``vmix_head2`` = ``abk80`` (Sig80, Bet80, Alf80), ``cpsw``, ntflux/``swdk``, surface fluxes
    src/mckpp_physics_verticalmixing_mod.F90:47-100, src/mckpp_physics_state_equations.F90, src/mckpp_fluxes_mod.F90:93-137
* ``vmix_tail2`` = reference integral, ``rimix`` + ``z121`` (one vector expression over the original
    neighbours), ``ddmix``, ``blmix_enhance``, the merge of kppmix, the bottom limits
    src/mckpp_physics_verticalmixing_mod.F90:102-159 and the ..._kppmix/rimix/z121/ddmix/blmix/enhance modules
* ``check_profile2``  src/mckpp_physics_overrides.F90:42-125

tests/test_second_reading.py runs the readings against each other -- the second in place of the first
(ocnstep, ocnint, check_profile) and as a probe after every vmix of the first -- on all five BASELINE
configurations (scaled) and on cases for every switch and exit, and requires bit-identical results;
oracle/AUDIT.md maps the Fortran statements of bldepth and ocnstep to both.
Evaluation order follows the Fortran (left to right, a*b/c = (a*b)/c; x**2 = x*x, x**4 = (x*x)*(x*x));
``exp`` is libm's through math.exp (numpy's SIMD exp may differ in the last bit).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

import oracle_lib

_SETUP = False


def _lib():
    global _SETUP
    L = oracle_lib.lib()
    if not _SETUP:
        L.orc_col_new.restype = C.c_void_p
        L.orc_col_new.argtypes = [C.POINTER(oracle_lib.OrcConst)]
        L.orc_col_free.argtypes = [C.c_void_p]
        L.orc_col_load.argtypes = [C.c_void_p, C.POINTER(oracle_lib.OrcConst), C.POINTER(oracle_lib.Orc3d), C.c_int, C.c_int]
        L.orc_col_store.argtypes = [C.c_void_p, C.POINTER(oracle_lib.OrcConst), C.POINTER(oracle_lib.Orc3d), C.c_int,
                                    C.c_int, C.c_int]
        L.orc_col_vmix.argtypes = [C.c_void_p, C.POINTER(oracle_lib.OrcConst), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.orc_col_ocnint.argtypes = [C.c_void_p, C.POINTER(oracle_lib.OrcConst), C.c_int, C.c_void_p, C.c_void_p]
        L.orc_col_check_profile.argtypes = [C.c_void_p, C.POINTER(oracle_lib.OrcConst)]
        L.orc_col_array.restype = C.POINTER(C.c_double)
        L.orc_col_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_col_get.restype = C.c_double
        L.orc_col_get.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_col_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_col_modeadv.restype = C.c_int
        L.orc_col_modeadv.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_wscale.argtypes = [C.POINTER(oracle_lib.OrcConst), C.c_double, C.c_double, C.c_double, C.c_double,
                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _SETUP = True
    return L


# --------------------------------------------------------------------------- wscale
def wscale(vonk, wmt, wst, sigma, hbl, ustar, bfsfc):
    """wscale_mod.F90:12-97 for arrays of (sigma, hbl, bfsfc) and scalar ustar.
    wmt, wst: (892, 50) arrays indexed [iz, ju]."""
    sigma = np.asarray(sigma, dtype=np.float64)
    hbl = np.broadcast_to(np.asarray(hbl, dtype=np.float64), sigma.shape)
    bfsfc = np.broadcast_to(np.asarray(bfsfc, dtype=np.float64), sigma.shape)
    ni, nj = 890, 48
    zmin, zmax, umin, umax, c1 = -4.e-7, 0.0, 0.0, 0.04, 5.0
    deltaz = (zmax - zmin) / (ni + 1)                   # :56
    deltau = (umax - umin) / (nj + 1)                   # :57
    zehat = vonk * sigma * hbl * bfsfc                  # :60
    tab = zehat <= zmax                                 # :62
    zdiff = zehat - zmin                                # :63
    qz = zdiff / deltaz
    # int() truncates toward zero; the clamp to [0, ni] makes out-of-int32-range values harmless
    iz = np.clip(np.trunc(np.clip(qz, -2.0e9, 2.0e9)).astype(np.int64), 0, ni)        # :64-66
    udiff = ustar - umin                                # :69
    qu = udiff / deltau
    ju = int(min(max(math.trunc(min(max(qu, -2.0e9), 2.0e9)), 0), nj))                # :70-72
    zfrac = qz - iz.astype(np.float64)                  # :75  (fractions are NOT clamped)
    ufrac = qu - float(ju)                              # :76
    fzfrac = 1. - zfrac                                 # :78
    wam = fzfrac * wmt[iz, ju + 1] + zfrac * wmt[iz + 1, ju + 1]     # :79-80
    wbm = fzfrac * wmt[iz, ju] + zfrac * wmt[iz + 1, ju]             # :81-82
    wm_t = (1. - ufrac) * wbm + ufrac * wam                          # :83
    was = fzfrac * wst[iz, ju + 1] + zfrac * wst[iz + 1, ju + 1]     # :85-86
    wbs = fzfrac * wst[iz, ju] + zfrac * wst[iz + 1, ju]             # :87-88
    ws_t = (1. - ufrac) * wbs + ufrac * was                          # :89
    ucube = ustar * ustar * ustar                       # :91  ustar**3
    with np.errstate(all="ignore"):
        wm_s = vonk * ustar * ucube / (ucube + c1 * zehat)           # :92
    return np.where(tab, wm_t, wm_s), np.where(tab, ws_t, wm_s)


def swfrac(fact, z, jwtype):
    """MCKPP_PHYSICS_SWFRAC, swfrac_mod.F90:49-79 (scalar)."""
    rfac = (0.58, 0.62, 0.67, 0.77, 0.78)
    a1 = (0.35, 0.6, 1.0, 1.5, 1.4)
    a2 = (23.0, 20.0, 17.0, 14.0, 7.9)
    j = jwtype - 1
    r1 = max(z * fact / a1[j], -80.)
    r2 = max(z * fact / a2[j], -80.)
    return rfac[j] * math.exp(r1) + (1. - rfac[j]) * math.exp(r2)


# --------------------------------------------------------------------------- bldepth
def bldepth(zm, hm, vonk, wmt, wst, dVsq, Ritop, dbloc, swfrac_tab, ustar, Bo, Bosol, f, ocdepth, jerlov, l_initflag):
    """bldepth_mod.F90:32-203, array-at-a-time over the levels kl = 2..km.

    zm, hm: 1-based views (index 0 unused, zm[1..kmp1]); dVsq, Ritop, dbloc, swfrac_tab
    likewise 1-based.  Returns (hbl, kbl, bfsfc, stable, caseA)."""
    kmp1 = len(zm) - 1
    km = kmp1 - 1
    epsln, Ricr, epsilon, cekman, cmonob, cs, cv = 1.e-16, 0.30, 0.1, 0.7, 1.0, 98.96, 1.6
    Vtc = cv * math.sqrt(0.2 / cs / epsilon) / (vonk * vonk) / Ricr         # :91  vonk**2
    hek = cekman * ustar / (abs(f) + epsln)                                 # :103
    kl = np.arange(2, km + 1)
    # while kbl is still km every level evaluates :119-126 afresh, so these are pure maps
    hcase = -zm[kl]                                                         # caseA used as hbl, :119
    bfs = Bo + Bosol * (1. - swfrac_tab[kl])                                # :122
    stab = 0.5 + np.copysign(0.5, bfs + epsln)                              # :123
    sig = stab * 1. + (1. - stab) * epsilon                                 # :124
    wm, ws = wscale(vonk, wmt, wst, sig, hcase, ustar, bfs)                 # :128
    dzu = zm[kl - 1] - zm[kl]
    dzl = zm[kl] - zm[kl + 1]
    bvsq = 0.5 * (dbloc[kl - 1] / dzu + dbloc[kl] / dzl)                    # :132-133
    Vtsq = -zm[kl] * ws * np.sqrt(np.abs(bvsq)) * Vtc                       # :134
    with np.errstate(all="ignore"):
        ribq = Ritop[kl] / (dVsq[kl] + Vtsq + epsln)                        # :136
    fmonob = stab * 1.0                                                     # :144
    dmo = cmonob * ustar * ustar * ustar / vonk / (np.abs(bfs) + epsln)     # :145-146
    dmo = fmonob * dmo - (1. - fmonob) * zm[kmp1]                           # :147
    hekman = stab * 1.0 * hek - (1. - stab * 1.0) * zm[kmp1]                # :157-158

    # the carried pair: Rib(ka) is the clamped Rib of the previous level (0 before kl = 2),
    # dmo(ka) the previous level's dmo (-zm(kmp1) before kl = 2)                :99-100, 188-190
    rib_prev = 0.0
    dmo_prev = -zm[kmp1]
    hbl, kbl = -zm[km], km                                                  # :101-102
    for i, k in enumerate(kl):
        rib_u = max(ribq[i], rib_prev + epsln)                              # :137
        with np.errstate(all="ignore"):
            hri = -zm[k - 1] + dzu[i] * (Ricr - rib_prev) / (rib_u - rib_prev)      # :139-140
        if dmo[i] <= -zm[k]:                                                # :148
            hm_ = (dmo[i] - dmo_prev) / dzu[i]                              # :149
            hmonob = (dmo[i] + hm_ * zm[k]) / (1. - hm_)                    # :150
        else:
            hmonob = -zm[kmp1]                                              # :152
        hmin = min(min(min(hri, hmonob), hekman[i]), -ocdepth)              # :161
        if hmin < -zm[k]:                                                   # :162
            if not l_initflag and hmin < -zm[k - 1]:                        # :173-174
                hmin2 = min(min(hri, hmonob), -ocdepth)                     # :175
                if hmin2 < -zm[k]:
                    hmin = hmin2
            hbl, kbl = hmin, int(k)                                         # :182-183
            break                    # IF(kbl.ge.km) is false from here on unless k == km, the last level
        rib_prev, dmo_prev = rib_u, dmo[i]
    sw = swfrac(-1.0, hbl, jerlov)                                          # :193
    bfsfc = Bo + Bosol * (1. - sw)                                          # :195
    stable = 0.5 + math.copysign(0.5, bfsfc)                                # :196
    bfsfc = bfsfc + stable * epsln                                          # :197
    caseA = 0.5 + math.copysign(0.5, -zm[kbl] - 0.5 * hm[kbl] - hbl)        # :201
    return hbl, kbl, bfsfc, stable, caseA


# --------------------------------------------------------------------------- ocnstep
class Column:
    """One kpp_1d_fields column held by the C oracle, with numpy views on its arrays."""

    def __init__(self, orc: "oracle_lib.Oracle"):
        self.L = _lib()
        self.orc = orc
        self.c = orc.c
        self.h = self.L.orc_col_new(C.byref(self.c))
        self.nzp1 = int(self.c.nzp1)
        self.nz = int(self.c.nz)
        n1 = self.nzp1 + 1
        self.U = self._arr("U").reshape(2, n1)[:, 1:]            # [l, k-1]
        self.X = self._arr("X").reshape(2, n1)[:, 1:]
        self.Us = self._arr("Us").reshape(2, 2, n1)[:, :, 1:]    # [time level, l, k-1]
        self.Xs = self._arr("Xs").reshape(2, 2, n1)[:, :, 1:]
        self.hmixd = self._arr("hmixd")
        self.difm, self.difs, self.dift, self.ghat = (self._arr(n) for n in ("difm", "difs", "dift", "ghat"))
        nt = int(self.c.nztmax) + 1
        self.wU = self._arr("wU").reshape(3, nt)                 # [component, interface 0:nztmax]
        self.wX = self._arr("wX").reshape(3, nt)
        self.talpha, self.sbeta = self._arr("talpha"), self._arr("sbeta")

    def _arr(self, name):
        lb, n = C.c_int(0), C.c_int(0)
        p = self.L.orc_col_array(self.h, name.encode(), C.byref(lb), C.byref(n))
        assert n.value > 0, name
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def get(self, name):
        return self.L.orc_col_get(self.h, name.encode())

    def set(self, name, v):
        self.L.orc_col_set(self.h, name.encode(), float(v))

    def vmix(self):
        h, k = C.c_double(0.0), C.c_int(0)
        self.L.orc_col_vmix(self.h, C.byref(self.c), C.byref(h), C.byref(k))
        return h.value, k.value

    def ocnint(self, kmix, Uo, Xo):
        # column-major (nzp1, 2): component l occupies a contiguous run of nzp1 values
        uo = np.ascontiguousarray(Uo, dtype=np.float64)
        xo = np.ascontiguousarray(Xo, dtype=np.float64)
        self.L.orc_col_ocnint(self.h, C.byref(self.c), int(kmix), uo.ctypes.data_as(C.c_void_p), xo.ctypes.data_as(C.c_void_p))

    def close(self):
        if self.h:
            self.L.orc_col_free(self.h)
            self.h = None


# --------------------------------------------------------------------------- first half of vmix: equation of state, surface fluxes
def abk80(S, T1, P):
    """MCKPP_ABK80 with alpha and beta requested, kappa not, for P != 0 (state_equations.F90:133-190 -> Sig80 :371-476,
    Bet80 :206-250, Alf80 :271-317), whole profiles at a time.  Returns alpha, beta, sig0."""
    S, T1, P = (np.asarray(v, dtype=np.float64) for v in (S, T1, P))
    assert (P != 0).all()
    T = np.where(T1 < -2., -2., T1)                                               # :143-144
    # ---- Sig80
    P0 = P / 10.0
    SR = np.sqrt(np.abs(S))
    R1 = ((((6.536332E-9 * T - 1.120083E-6) * T + 1.001685E-4) * T - 9.095290E-3) * T + 6.793952E-2) * T - .157406
    R2 = (((5.3875E-9 * T - 8.2467E-7) * T + 7.6438E-5) * T - 4.0899E-3) * T + 8.24493E-1
    R3 = (-1.6546E-6 * T + 1.0227E-4) * T - 5.72466E-3
    R4 = 4.8314E-4
    Sig0 = (R4 * S + R3 * SR + R2) * S + R1
    Rho0 = 1000.0 + Sig0
    B1 = (-5.3009E-4 * T + 1.6483E-2) * T + 7.944E-2                              # BlkMod :438-
    A1 = ((-6.1670E-5 * T + 1.09987E-2) * T - 0.603459) * T + 54.6746
    KW = (((-5.155288E-5 * T + 1.360477E-2) * T - 2.327105) * T + 148.4206) * T + 19652.21
    K0 = (B1 * SR + A1) * S + KW
    E = (9.1697E-10 * T + 2.0816E-8) * T - 9.9348E-7
    BW = (5.2787E-8 * T - 6.12293E-6) * T + 8.50935E-5
    B = BW + E * S
    D = 1.91075E-4
    C_ = (-1.6078E-6 * T - 1.0981E-5) * T + 2.2838E-3
    AW = ((-5.77905E-7 * T + 1.16092E-4) * T + 1.43713E-3) * T + 3.239908
    A = (D * SR + C_) * S + AW
    K = (B * P0 + A) * P0 + K0
    PK = P0 / K
    Sig = (1000.0 * PK + Sig0) / (1.0 - PK)
    Rho = 1000.0 + Sig
    # ---- Bet80
    SR5 = SR * 1.5
    DRho = R2 + SR5 * R3 + (S + S) * R4
    DK0 = A1 + SR5 * B1
    DA = C_ + SR5 * D
    DK = (E * P0 + DA) * P0 + DK0
    ABFac = Rho0 * P0 / ((K - P0) * (K - P0))
    Beta = DRho / (1. - PK) - ABFac * DK
    Beta = Beta / Rho
    # ---- Alf80 (its own, differentiated, power series)
    R1 = (((.3268166E-7 * T - .4480332e-5) * T + .3005055e-3) * T - .1819058E-1) * T + 6.793952E-2
    R2 = ((.215500E-7 * T - .247401E-5) * T + .152876E-3) * T - 4.0899E-3
    R3 = -.33092E-5 * T + 1.0227E-4
    Alph0 = (R3 * SR + R2) * S + R1
    B1 = -.106018E-2 * T + 1.6483E-2
    A1 = (-.18501E-3 * T + .219974E-1) * T - 0.603459
    KW = ((-.2062115E-3 * T + .4081431E-1) * T - .4654210E+1) * T + 148.4206
    K0 = (B1 * SR + A1) * S + KW
    E = .183394E-8 * T + 2.0816E-8
    BW = .105574E-6 * T - 6.12293E-6
    AlphB = BW + E * S
    C_ = -.32156E-5 * T - 1.0981E-5
    AW = (-.1733715E-5 * T + .232184E-3) * T + 1.43713E-3
    AlphaA = C_ * S + AW
    AlphK = (AlphB * P0 + AlphaA) * P0 + K0
    Alpha = Alph0 / (1. - PK) - ABFac * AlphK
    Alpha = -Alpha / Rho
    return Alpha, Beta, Sig0


def cpsw(S, T1, P0):
    """MCKPP_CPSW (state_equations.F90:7-58), whole profiles at a time."""
    S, T1, P0 = (np.asarray(v, dtype=np.float64) for v in (S, T1, P0))
    T = np.where(T1 < -2., -2., T1)
    P = P0 / 10.
    SR = np.sqrt(np.abs(S))
    A = (-1.38385E-3 * T + 0.1072763) * T - 7.643575
    B = (5.148E-5 * T - 4.07718E-3) * T + 0.1770383
    C_ = (((2.093236E-5 * T - 2.654387E-3) * T + 0.1412855) * T - 3.720283) * T + 4217.4
    CP0 = (B * SR + A) * S + C_
    A = (((1.7168E-8 * T + 2.0357E-6) * T - 3.13885E-4) * T + 1.45747E-2) * T - 0.49592
    B = (((2.2956E-11 * T - 4.0027E-9) * T + 2.87533E-7) * T - 1.08645E-5) * T + 2.4931E-4
    C_ = ((6.136E-13 * T - 6.5637E-11) * T + 2.6380E-9) * T - 5.422E-8
    CP1 = ((C_ * P + B) * P + A) * P
    A = (((-2.9179E-10 * T + 2.5941E-8) * T + 9.802E-7) * T - 1.28315E-4) * T + 4.9247E-3
    B = (3.122E-8 * T - 1.517E-6) * T - 1.2331E-4
    A = (A + B * SR) * S
    B = ((1.8448E-11 * T - 2.3905E-9) * T + 1.17054E-7) * T - 2.9558E-6
    B = (B + 9.971E-8 * SR) * S
    C_ = (3.513E-13 * T - 1.7682E-11) * T + 5.540E-10
    C_ = (C_ - 1.4300E-12 * T * SR) * S
    CP2 = ((C_ * P + B) * P + A) * P
    return CP0 + CP1 + CP2


def swdk(z, j):
    """mckpp_fluxes_swdk (fluxes_mod.F90:120-137), scalar (libm exp, as the oracle)."""
    rfac = (0.58, 0.62, 0.67, 0.77, 0.78)
    a1 = (0.35, 0.6, 1.0, 1.5, 1.4)
    a2 = (23.0, 20.0, 17.0, 14.0, 7.9)
    return rfac[j - 1] * math.exp(z / a1[j - 1]) + (1.0 - rfac[j - 1]) * math.exp(z / a2[j - 1])


def vmix_head2(col, cf, swdk_before):
    """verticalmixing_mod.F90:47-100 and mckpp_fluxes_ntflux (fluxes_mod.F90:93-116): EOS at every level, the
    non-turbulent flux profile, the kinematic surface fluxes, ustar, B0, B0sol.  swdk_before = the column's swdk_opt
    as it was before the call (it is only recomputed when ntime <= 1)."""
    k_ = cf.consts
    nz, nzp1 = col.nz, col.nzp1
    zm, dmv = cf.zm, cf.dm
    X = col.X
    g = col.get
    Sref, Ssurf, ntime, jerlov = g("Sref"), g("Ssurf"), int(g("ntime")), int(g("jerlov"))
    out = {}
    _, _, s0 = abk80(np.array([0.0]), X[0, 0:1], -zm[0:1])                        # :47-48
    rhoh2o = 1000. + float(s0[0])
    _, _, s0 = abk80(np.array([k_.sice]), X[0, 0:1], -zm[0:1])                    # :49-50
    rhob = 1000. + float(s0[0])
    alpha, beta, sig0 = abk80(X[1] + Sref, X[0], -zm[:nzp1])                      # :54-63
    rho, cp, ta, sb, buoy = (np.zeros(nzp1 + 1) for _ in range(5))
    rho[1:] = 1000. + sig0
    cp[1:] = cpsw(X[1] + Sref, X[0], -zm[:nzp1])
    ta[1:], sb[1:] = alpha, beta
    buoy[1:] = -k_.grav * sig0 / 1000.
    rho[0], cp[0], ta[0], sb[0] = rho[1], cp[1], ta[1], sb[1]                     # :65-68
    sf = col._arr("sflux")
    ns = int(col.c.nsflxs)
    sflux = lambda i: float(sf[4 * ns + i - 1])                                   # sflux(i,5,0)
    swdk_opt = np.array(swdk_before[:nz + 1])
    if ntime <= 1:                                                                # fluxes_mod.F90:103-108
        swdk_opt = np.array([swdk(-float(dmv[k]), jerlov) for k in range(nz + 1)])
    wXNT = None
    if ntime >= 1:                                                                # :110-116
        wXNT = -sflux(3) * swdk_opt / (rho[0] * cp[0])
    wU0 = (-sflux(1) / rho[0], -sflux(2) / rho[0])                                # :76-77
    tau = math.sqrt(sflux(1) * sflux(1) + sflux(2) * sflux(2)) + 1.e-16           # :78
    ustar = math.sqrt(tau / rho[0])                                               # :80
    wX01 = -sflux(4) / rho[0] / cp[0]                                             # :83
    wX02 = Ssurf * sflux(6) / rhoh2o + (Ssurf - k_.sice) * sflux(5) / rhob        # :86-88
    B0 = -k_.grav * (ta[0] * wX01 - sb[0] * wX02)                                 # :91-92
    B0sol = k_.grav * ta[0] * sflux(3) / (rho[0] * cp[0])                         # :94-95
    out.update(rho=rho, cp=cp, talpha=ta, sbeta=sb, buoy=buoy[1:], swdk_opt=swdk_opt, wXNT=wXNT, rhoh2o=rhoh2o,
               wU0=np.array(wU0), wX0=np.array([wX01, wX02, -B0]), ustar=ustar, Bo=B0, Bosol=B0sol)
    return out


def vmix_head_probe(cf, log):
    """probe for ocnstep(): after every vmix of the C oracle, recompute the first half of vmix with the second reading
    and append (name, ours, theirs).  (swdk_opt only changes while ntime <= 1, where it is recomputed from scratch;
    later the column's stored profile is the input.)"""
    def probe(col: Column):
        ours = vmix_head2(col, cf, np.array(col._arr("swdk_opt")))
        nz, nzp1 = col.nz, col.nzp1
        nt = int(col.c.nztmax) + 1
        A = col._arr
        theirs = dict(rho=A("rho")[:nzp1 + 1], cp=A("cp")[:nzp1 + 1], talpha=A("talpha")[:nzp1 + 1], sbeta=A("sbeta")[:nzp1 + 1],
                      buoy=A("buoy")[1:nzp1 + 1], swdk_opt=A("swdk_opt")[:nz + 1], wXNT=A("wXNT").reshape(2, nt)[0, :nz + 1],
                      rhoh2o=col.get("rhoh2o"), wU0=A("wU").reshape(3, nt)[0:2, 0], wX0=A("wX").reshape(3, nt)[:, 0],
                      ustar=col.get("dbg_ustar"), Bo=col.get("dbg_Bo"), Bosol=col.get("dbg_Bosol"))
        if int(col.get("ntime")) <= 1:      # the table bldepth builds while ntime <= 1 (MCKPP_PHYSICS_SWFRAC_OPT, swfrac_mod.F90:14-47, called with fact = hbf, bldepth_mod.F90:114)
            ours["swfrac"] = np.array([swfrac(1.0, float(cf.zm[l]), int(col.get("jerlov"))) for l in range(nzp1)])
            theirs["swfrac"] = A("swfrac")[1:nzp1 + 1]
        for name, t in theirs.items():
            if ours[name] is None:
                continue
            log.append((name, np.atleast_1d(np.array(ours[name], dtype=np.float64)), np.atleast_1d(np.array(t, dtype=np.float64))))
    return probe


# --------------------------------------------------------------------------- second half of vmix: kppmix and below
def z121(V, vlo, vhi):
    """MCKPP_PHYSICS_VERTICALMIXING_Z121 (z121_mod.F90:7-45), all levels at once.  V[0..kmp1]; returns the smoothed
    array and the weights.  The reference walks down with the previous level's original value parked in V(0):
    every output is a combination of ORIGINAL neighbours, which is what is written here."""
    kmp1 = len(V) - 1
    km = kmp1 - 1
    o = np.array(V, dtype=np.float64)
    o[0] = 0.0; o[kmp1] = 0.0                                                     # :24-27
    w = np.zeros(kmp1 + 1)
    w[1:km + 1] = np.where((o[1:km + 1] < vlo) | (o[1:km + 1] > vhi), 0.0, 1.0)   # :29-36
    out = o.copy()
    out[1:km + 1] = (w[0:km] * o[0:km] + 2. * o[1:km + 1] + w[2:km + 2] * o[2:km + 2]) / (w[0:km] + 2.0 + w[2:km + 2])   # :38-44
    out[0] = o[km] if km >= 1 else 0.0                                            # V(0) = tmp of the last level
    return out, w


def rimix(zm, dbloc, Shsq, km):
    """MCKPP_PHYSICS_VERTICALMIXING_RIMIX (rimix_mod.F90:13-106).  zm[k-1] = zm(k); dbloc, Shsq by Fortran index.
    Returns Rig[0..km] (slot 0 unused) and difm, difs, dift[0..km+1]."""
    epsln, Riinfty, Ricon = 1.e-16, 0.8, -0.2
    difm0, difs0, difmiw, difsiw, difmcon, difscon, c1, c0 = 0.005, 0.005, 0.0001, 0.00001, 0.0, 0.0, 1.0, 0.0
    kmp1 = km + 1
    Rig = np.zeros(km + 1)
    Rig[1:] = dbloc[1:km + 1] * (zm[0:km] - zm[1:km + 1]) / (Shsq[1:km + 1] + epsln)   # :47-48
    V = np.zeros(kmp1 + 1)
    V[1:km + 1] = Rig[1:]
    sm, _w = z121(V, c0, Riinfty)                                                 # :56-58 (mRi = 1)
    Rigg = np.maximum(Rig[1:], Ricon)                                             # :65 (unsmoothed)
    ratio = np.minimum((Ricon - Rigg) / Ricon, c1)
    fcon = (c1 - ratio * ratio)
    fcon = fcon * fcon * fcon
    Rigg = np.maximum(sm[1:km + 1], c0)                                           # :70 (smoothed)
    ratio = np.minimum(Rigg / Riinfty, c1)
    fri = (c1 - ratio * ratio)
    fri = fri * fri * fri
    difm, difs = np.zeros(kmp1 + 1), np.zeros(kmp1 + 1)
    difm[1:km + 1] = (difmiw + fcon * difmcon + fri * difm0)                      # :93
    difs[1:km + 1] = (difsiw + fcon * difscon + fri * difs0)                      # :94
    dift = difs.copy()                                                            # :95
    return Rig, difm, difs, dift


def ddmix(alphaDT, betaDS, difs, dift, km):
    """MCKPP_PHYSICS_VERTICALMIXING_DDMIX (ddmix_mod.F90:12-52): adds to difs, dift in place."""
    Rrho0, dsfmax = 1.9, 1.0e-4
    for ki in range(1, km + 1):
        a, b = float(alphaDT[ki]), float(betaDS[ki])
        if a > b and b > 0.:                                                      # salt fingering :31-36
            Rrho = min(a / b, Rrho0)
            t = (Rrho - 1) / (Rrho0 - 1)
            diffdd = 1.0 - t * t
            diffdd = dsfmax * diffdd * diffdd * diffdd
            dift[ki] = dift[ki] + diffdd * 0.8 / Rrho
            difs[ki] = difs[ki] + diffdd
        elif a < 0.0 and b < 0.0 and a < b:                                       # diffusive convection :39-46
            Rrho = a / b
            diffdd = 1.5e-6 * 9.0 * 0.101 * math.exp(4.6 * math.exp(-0.54 * (1 / Rrho - 1)))
            prandtl = 0.15 * Rrho
            if Rrho > 0.5:
                prandtl = (1.85 - 0.85 / Rrho) * Rrho
            dift[ki] = dift[ki] + diffdd
            difs[ki] = difs[ki] + prandtl * diffdd


def blmix_enhance(vonk, wmt, wst, zm, hm, km, ustar, bfsfc, hbl, stable, caseA, kbl, difm, difs, dift):
    """mckpp_physics_verticalmixing_blmix (blmix_mod.F90:13-151) followed by ..._ENHANCE (enhance_mod.F90:10-51).
    zm[k-1] = zm(k), hm[k-1] = hm(k); difm, difs, dift[0..km+1] are the interior values.
    Returns blmc[3][0..km] (slot 0 unused) and ghat[0..km] as the two routines leave them."""
    epsln, epsilon, c1, cs, cstar = 1.e-20, 0.1, 5.0, 98.96, 5.0
    Z = lambda k: zm[k - 1]
    H = lambda k: hm[k - 1]
    cg = cstar * vonk * (cs * vonk * epsilon) ** (1. / 3.)                        # :62
    sigma = stable * 1.0 + (1. - stable) * epsilon                                # :65
    wm, ws = (float(v) for v in wscale(vonk, wmt, wst, sigma, hbl, ustar, bfsfc))
    ic = int(caseA + epsln)
    kn = ic * (kbl - 1) + (1 - ic) * kbl                                          # :68
    delhat = 0.5 * H(kn) - Z(kn) - hbl                                            # :71
    R = 1.0 - delhat / H(kn)

    def at_hbl(d):                                                                # :73-87, one diffusivity at a time
        dvdzup = (d[kn - 1] - d[kn]) / H(kn)
        dvdzdn = (d[kn] - d[kn + 1]) / H(kn + 1)
        dp = 0.5 * ((1. - R) * (dvdzup + abs(dvdzup)) + R * (dvdzdn + abs(dvdzdn)))
        return dp, d[kn] + dp * delhat

    viscp, visch = at_hbl(difm)
    difsp, difsh = at_hbl(difs)
    diftp, difth = at_hbl(dift)
    f1 = stable * c1 * bfsfc / ((ustar * ustar) * (ustar * ustar) + epsln)        # :89  ustar**4
    gat1, dat1 = [0.0] * 3, [0.0] * 3
    for m, (dh, dp, w) in enumerate(((visch, viscp, wm), (difsh, difsp, ws), (difth, diftp, ws))):
        gat1[m] = dh / hbl / (w + epsln)                                          # :90-100
        dat1[m] = min(-dp / (w + epsln) + f1 * dh, 0.)
    ki = np.arange(1, km + 1)
    sig = (-zm[0:km] + 0.5 * hm[0:km]) / hbl                                      # :113
    sigma_k = stable * sig + (1. - stable) * np.minimum(sig, epsilon)
    wm_k, ws_k = wscale(vonk, wmt, wst, sigma_k, hbl, ustar, bfsfc)
    a1, a2, a3 = sig - 2., 3. - 2. * sig, sig - 1.
    blmc = np.zeros((3, km + 1))
    for m, w in enumerate((wm_k, ws_k, ws_k)):
        G = a1 + a2 * gat1[m] + a3 * dat1[m]                                      # :123-125
        blmc[m, 1:] = hbl * w * sig * (1. + sig * G)                              # :128-130
    ghat = np.zeros(km + 1)
    ghat[1:] = (1. - stable) * cg / (ws_k * hbl + epsln)                          # :133
    # diffusivities at the kbl-1 grid level :137-150
    sg = -Z(kbl - 1) / hbl
    sigma = stable * sg + (1. - stable) * min(sg, epsilon)
    wm, ws = (float(v) for v in wscale(vonk, wmt, wst, sigma, hbl, ustar, bfsfc))
    b1, b2, b3 = sg - 2., 3. - 2. * sg, sg - 1.
    dkm1 = [hbl * w * sg * (1. + sg * (b1 + b2 * gat1[m] + b3 * dat1[m])) for m, w in enumerate((wm, ws, ws))]
    # enhance :33-48
    k1 = kbl - 1
    if 1 <= k1 <= km - 1:
        delta = (hbl + Z(k1)) / (Z(k1) - Z(k1 + 1))
        for m, d in enumerate((difm, difs, dift)):
            dkmp5 = caseA * d[k1] + (1. - caseA) * blmc[m, k1]
            dstar = ((1. - delta) * (1. - delta)) * dkm1[m] + (delta * delta) * dkmp5
            blmc[m, k1] = (1. - delta) * d[k1] + delta * dstar
        ghat[k1] = (1. - caseA) * ghat[k1]
    return blmc, ghat


def vmix_tail2(col, cf):
    """verticalmixing_mod.F90:102-159 and everything it calls except bldepth (whose inputs and results are taken
    from the C oracle's last call; bldepth has its own second reading above): alphaDT/betaDS, the surface-layer
    reference integral, Ritop, dVsq, dbloc, Shsq, then kppmix = rimix (+z121) + ddmix + blmix + enhance + the merge,
    and the bottom limits.  Reads the iterate and the EOS results of the column; returns the arrays by name."""
    k_ = cf.consts
    nz, nzp1 = col.nz, col.nzp1
    zm, hm = cf.zm, cf.hm
    U, X = col.U, col.X
    buoy, talpha, sbeta = col._arr("buoy"), col._arr("talpha"), col._arr("sbeta")
    epsilon = 0.1
    out = {}
    alphaDT, betaDS = np.zeros(nzp1 + 1), np.zeros(nzp1 + 1)
    alphaDT[1:nz + 1] = 0.5 * (talpha[1:nz + 1] + talpha[2:nz + 2]) * (X[0, 0:nz] - X[0, 1:nz + 1])     # :103-104
    betaDS[1:nz + 1] = 0.5 * (sbeta[1:nz + 1] + sbeta[2:nz + 2]) * (X[1, 0:nz] - X[1, 1:nz + 1])       # :105-106
    Ritop, dVsq = np.zeros(nzp1 + 1), np.zeros(nzp1 + 1)
    dz = zm[0:nz] - zm[1:nz + 1]                                  # zm(kl) - zm(kl+1)
    for n in range(1, nz + 1):                                                    # :110-137
        zref = epsilon * zm[n - 1]
        wz = max(zm[0], zref)
        uref = U[0, 0] * wz / zref
        vref = U[1, 0] * wz / zref
        bref = buoy[1] * wz / zref
        # the levels the integral visits: kl = 1.. while zref < zm(kl)
        m = int(np.argmax(zref >= zm[0:nz])) if (zref >= zm[0:nz]).any() else nz
        wzs = np.minimum(dz[:m], zm[:m] - zref)                                   # :119
        dels = 0.5 * wzs / dz[:m]                                                 # :120
        tu = wzs * (U[0, :m] + dels * (U[0, 1:m + 1] - U[0, :m])) / zref          # :121-122, term by term
        tv = wzs * (U[1, :m] + dels * (U[1, 1:m + 1] - U[1, :m])) / zref
        tb = wzs * (buoy[1:m + 1] + dels * (buoy[2:m + 2] - buoy[1:m + 1])) / zref
        for j in range(m):                                        # subtracted in level order
            uref = uref - tu[j]; vref = vref - tv[j]; bref = bref - tb[j]
        Ritop[n] = (zref - zm[n - 1]) * (bref - buoy[n])                          # :130
        du, dv = uref - U[0, n - 1], vref - U[1, n - 1]
        dVsq[n] = du * du + dv * dv                                               # :134
    dbloc = np.zeros(nz + 1)
    dbloc[1:] = buoy[1:nz + 1] - buoy[2:nz + 2]                                   # :133
    Shsq = np.zeros(nzp1 + 1)
    d1, d2 = U[0, 0:nz] - U[0, 1:nz + 1], U[1, 0:nz] - U[1, 1:nz + 1]
    Shsq[1:nz + 1] = d1 * d1 + d2 * d2                                            # :135-136
    out.update(alphaDT=alphaDT, betaDS=betaDS, Ritop=Ritop, dVsq=dVsq, dbloc=dbloc, Shsq=Shsq)
    # ---- kppmix (kppmix_mod.F90:25-126)
    km, kmp1 = nz, nzp1
    difm, difs, dift = np.zeros(kmp1 + 1), np.zeros(kmp1 + 1), np.zeros(kmp1 + 1)
    Rig = None
    if k_.LRI:
        Rig, difm, difs, dift = rimix(zm, dbloc, Shsq, km)
    if k_.LDD:
        ddmix(alphaDT, betaDS, difs, dift, km)
    difm[kmp1], difs[kmp1], dift[kmp1] = difm[km], difs[km], dift[km]             # :79-81
    ghat = None
    if k_.LKPP:
        g = col.get
        hbl, kbl = g("dbg_hbl"), int(g("dbg_kbl"))
        wmt = np.asarray(cf.wmt).reshape(50, 892).T if np.asarray(cf.wmt).ndim == 1 else np.asarray(cf.wmt)
        wst = np.asarray(cf.wst).reshape(50, 892).T if np.asarray(cf.wst).ndim == 1 else np.asarray(cf.wst)
        blmc, ghat = blmix_enhance(k_.vonk, wmt, wst, zm, hm, km, g("dbg_ustar"), g("dbg_bfsfc"), hbl,
                                   g("dbg_stable"), g("dbg_caseA"), kbl, difm, difs, dift)
        inside = np.arange(km + 1) < kbl                                          # :103-111
        inside[0] = False
        difm[:km + 1] = np.where(inside, blmc[0], difm[:km + 1])
        difs[:km + 1] = np.where(inside, blmc[1], difs[:km + 1])
        dift[:km + 1] = np.where(inside, blmc[2], dift[:km + 1])
        ghat = np.where(inside, ghat, 0.0)
    # ---- bottom limits (verticalmixing_mod.F90:151-159)
    difm[nz:nzp1 + 1] = 0.0001
    difs[nz:nzp1 + 1] = 0.00001
    dift[nz:nzp1 + 1] = 0.00001
    if ghat is not None:
        ghat[nz] = 0.0
    out.update(Rig=Rig, difm=difm, difs=difs, dift=dift, ghat=ghat)
    return out


def vmix_tail_probe(cf, log):
    """probe for ocnstep(): after every vmix of the C oracle, recompute the second half of vmix with the second
    reading and append (name, ours, theirs) triples to `log`."""
    def probe(col: Column):
        ours = vmix_tail2(col, cf)
        nz, nzp1 = col.nz, col.nzp1
        spans = dict(alphaDT=(1, nz), betaDS=(1, nz), Ritop=(1, nz), dVsq=(1, nz), dbloc=(1, nz), Shsq=(1, nz),
                     Rig=(1, nz), difm=(0, nzp1), difs=(0, nzp1), dift=(0, nzp1), ghat=(1, nz))
        for name, (a, b) in spans.items():
            if ours[name] is None:
                continue
            log.append((name, np.array(ours[name][a:b + 1]), np.array(col._arr(name)[a:b + 1])))
    return probe


# --------------------------------------------------------------------------- ocnint and the solvers
def tridcof(tri, diff, nz):
    """mckpp_physics_solvers_tridcof (solvers.F90:14-46), array-at-a-time.  tri[k, j] = tri(k,j,1); diff[0..nz].
    Returns cu, cc, cl addressed by the Fortran index (slot 0 unused)."""
    cu, cc, cl = np.zeros(nz + 1), np.zeros(nz + 1), np.zeros(nz + 1)
    cc[1] = 1. + tri[1, 1] * diff[1]                                              # :32
    cl[1] = -(tri[1, 1] * diff[1])                                                # :33
    t0, t1 = tri[2:nz + 1, 0], tri[2:nz + 1, 1]
    cu[2:] = -(t0 * diff[1:nz])                                                   # :37
    cc[2:] = 1. + t1 * diff[2:nz + 1] + t0 * diff[1:nz]                           # :38
    cl[2:] = -(t1 * diff[2:nz + 1])                                               # :39
    cl[nz] = 0.                                                                   # :43
    return cu, cc, cl


def tridrhs(tri, hm, yo, ntflux, diff, ghat, sturflux, ghatflux, dto, nz):
    """mckpp_physics_solvers_tridrhs (solvers.F90:56-109) for npd = 1 (its only call).  hm[k-1] = hm(k),
    yo[k-1] = yo(k) (nz+1 values), ntflux[0..nz], diff[0..nz], ghat[1..nz] by Fortran index."""
    divflx = 1.0 / float(1)
    rhs = np.zeros(nz + 1)
    h = hm[:nz]                                                                   # h(1..nz)
    rhs[1] = yo[0] + dto / h[0] * (ghatflux * diff[1] * ghat[1] - sturflux * divflx + ntflux[1] - ntflux[0])   # :86-87
    i = np.arange(2, nz + 1)                                                      # :101-104 and the bottom layer :107-111
    rhs[2:] = yo[1:nz] + dto / h[1:] * (ghatflux * (diff[i] * ghat[i] - diff[i - 1] * ghat[i - 1]) + ntflux[i] - ntflux[i - 1])
    if nz > 1:
        rhs[nz] = rhs[nz] + yo[nz] * tri[nz, 1] * diff[nz]
    return rhs


def tridmat(cu, cc, cl, rhs, yo_below, nz):
    """mckpp_physics_solvers_tridmat (solvers.F90:114-161).  Returns (yn[k-1] for k = 1..nz+1, zero pivot met)."""
    yn = [0.0] * (nz + 2)
    gam = [0.0] * (nz + 2)
    pivot = False
    bet = float(cc[1])
    yn[1] = float(rhs[1]) / bet                                                   # :136
    for i in range(2, nz + 1):
        gam[i] = float(cl[i - 1]) / bet                                           # :138
        bet = float(cc[i]) - float(cu[i]) * gam[i]                                # :139
        if bet == 0.:                                                             # :140-150: the reference aborts;
            pivot = True                                                          # oracle and GPU flag and go on
            bet = 1.E-12
        yn[i] = (float(rhs[i]) - float(cu[i]) * yn[i - 1]) / bet                  # :153
    for i in range(nz - 1, 0, -1):
        yn[i] = yn[i] - gam[i + 1] * yn[i + 1]                                    # :156-158
    yn[nz + 1] = float(yo_below)                                                  # :159
    return np.array(yn[1:]), pivot


def rhsmod_salt(mode, A, dto, km, dmk, nz, rhs, zm, hm):
    """mckpp_physics_solvers_rhsmod (solvers.F90:176-335) for jsclr = 2 (its only call).  rhs by Fortran index."""
    if mode <= 0:
        return
    fact = dto * A * 0.033

    def total(n1, n2):                       # delta = delta + hm(n), n = n1..n2, in that order
        d = 0.0
        for n in range(n1, n2 + 1):
            d = d + float(hm[n - 1])
        return d

    if mode == 1:
        rhs[1] = rhs[1] + fact / hm[0]                                            # :229-233
        return
    if mode == 2:
        n1, n2 = 1, km - 1                                                        # :235-245
        delta = total(n1, n2)
    elif mode == 3:
        n1, n2 = 1, nz                                                            # :247-257
        delta = total(n1, n2)
    elif mode == 4:
        n1 = 1                                                                    # :259-273
        while zm[n1 - 1] >= -100.:
            n1 += 1
        n2 = nz - 1
        delta = total(n1, n2)
    elif mode == 5:
        rhs[nz] = rhs[nz] + fact / hm[nz - 1]                                     # :275-279
        return
    elif mode in (6, 7):
        if mode == 6:
            n1 = 1                                                                # :297-309
            depth = float(hm[0])
            dmax = dmk - 0.5 * (hm[km - 1] + hm[km - 2])
        else:
            n1 = km - 1                                                           # :311-322
            depth = dmk - 0.5 * hm[km - 1]
            dmax = 100.
        delta, n2 = 0.0, n1
        for n in range(n1, nz + 1):
            n2 = n
            delta = delta + float(hm[n - 1])
            depth = depth + float(hm[n])
            if depth >= dmax:
                break
    else:
        raise ValueError("mode out of range")                                     # :324-327 (validated upstream)
    if n2 >= n1:                                                                  # (an empty DO loop divides nothing)
        rhs[n1:n2 + 1] = rhs[n1:n2 + 1] + fact / delta


def ocnint2(col, cf, kmixe, Uo, Xo):
    """mckpp_physics_ocnint (ocnint_mod.F90:19-221) on the column's current iterate, written against the Fortran
    independently of the C oracle's ocnint(): coefficient and right-hand-side arrays at a time, the recurrences of
    tridmat as scalar loops.  Writes U, X, fcorr, tinc_fcorr, ocnTcorr, sinc_fcorr, scorr of the column, as
    ocnint does.  Returns True if a zero pivot was met."""
    k_ = cf.consts
    NZ, NZP1 = col.nz, col.nzp1
    nt = int(col.c.nztmax) + 1
    tri = np.asarray(cf.tri)[:, :, 0]
    hm, zm, dmv = cf.hm, cf.zm, cf.dm
    dto = k_.dto
    U, X = col.U, col.X
    difm, dift, difs, ghat = (col._arr(n) for n in ("difm", "dift", "difs", "ghat"))
    wU = col._arr("wU").reshape(3, nt)
    wX = col._arr("wX").reshape(3, nt)
    wXNT = col._arr("wXNT").reshape(2, nt)
    rho, cp = col._arr("rho"), col._arr("cp")
    ftemp = col.get("f")                                                          # :43
    pivot = False
    lev = slice(1, NZ - 1)                                                        # 0-based positions of levels 2..NZ-1

    # ---- U (:45-60); U(:,2) is still the iterate's V
    cu, cc, cl = tridcof(tri, difm, NZ)
    rhs = np.zeros(NZ + 1)
    rhs[1] = Uo[0, 0] + dto * (ftemp * .5 * (Uo[1, 0] + U[1, 0]) - wU[0, 0] / hm[0])
    rhs[2:NZ] = Uo[0, lev] + dto * ftemp * .5 * (Uo[1, lev] + U[1, lev])
    rhs[NZ] = Uo[0, NZ - 1] + dto * ftemp * .5 * (Uo[1, NZ - 1] + U[1, NZ - 1]) + tri[NZ, 1] * difm[NZ] * Uo[0, NZ]
    unew, pz = tridmat(cu, cc, cl, rhs, Uo[0, NZ], NZ)
    pivot |= pz
    U[0, :] = unew
    # ---- V, with the new U (:62-72)
    rhs[1] = Uo[1, 0] - dto * (ftemp * .5 * (Uo[0, 0] + U[0, 0]) + wU[1, 0] / hm[0])
    rhs[2:NZ] = Uo[1, lev] - dto * ftemp * .5 * (Uo[0, lev] + U[0, lev])
    rhs[NZ] = Uo[1, NZ - 1] - dto * ftemp * .5 * (Uo[0, NZ - 1] + U[0, NZ - 1]) + tri[NZ, 1] * difm[NZ] * Uo[1, NZ]
    vnew, pz = tridmat(cu, cc, cl, rhs, Uo[1, NZ], NZ)
    pivot |= pz
    U[1, :] = vnew

    # ---- temperature (:81-163)
    ghatflux = sturflux = wX[0, 0]
    cu, cc, cl = tridcof(tri, dift, NZ)
    rhs = tridrhs(tri, hm, Xo[0], wXNT[0], dift, ghat, sturflux, ghatflux, dto, NZ)
    if k_.L_RELAX_SST and not k_.L_FCORR_WITHZ and not k_.L_FCORR:                # :98-112
        relax_sst = col.get("relax_sst")
        if relax_sst > 1.e-10:
            sst0 = col.get("SST0")
            if not k_.L_RELAX_CALCONLY:
                rhs[1] = rhs[1] + dto * relax_sst * (sst0 - Xo[0, 0]) * dmv[kmixe] / hm[0]
            col.set("fcorr", relax_sst * (sst0 - Xo[0, 0]) * dmv[kmixe] * rho[1] * cp[1])
        else:
            col.set("fcorr", 0.0)
    if k_.L_FCORR and not k_.L_RELAX_SST and not k_.L_FCORR_WITHZ:                # :119-123
        rhs[1] = rhs[1] + dto * col.get("fcorr_twod") / (rho[1] * cp[1] * hm[0])
    kk = slice(1, NZP1 + 1)
    tinc = np.zeros(NZP1 + 1)                                                     # :132
    if k_.L_FCORR_WITHZ and not k_.L_FCORR:                                       # :133-138
        tinc[kk] = dto * col._arr("fcorr_withz")[kk] / (rho[kk] * cp[kk])
    if k_.L_RELAX_OCNT:                                                           # :143-151
        tinc[kk] = tinc[kk] + dto * col.get("relax_ocnT") * (col._arr("ocnT_clim")[kk] - Xo[0])
    rhs[1:] = rhs[1:] + tinc[1:NZ + 1]                                            # :152-153 (levels the solver reads)
    col._arr("tinc_fcorr")[kk] = tinc[kk]
    col._arr("ocnTcorr")[kk] = tinc[kk] * rho[kk] * cp[kk] / dto                  # :157-158
    tnew, pz = tridmat(cu, cc, cl, rhs, Xo[0, NZ], NZ)
    pivot |= pz
    X[0, :] = tnew

    # ---- salinity (:165-218)
    cu, cc, cl = tridcof(tri, difs, NZ)
    ghatflux = sturflux = wX[1, 0]
    rhs = tridrhs(tri, hm, Xo[1], wXNT[1], difs, ghat, sturflux, ghatflux, dto, NZ)
    adv = col._arr("advection").reshape(2, -1)
    for imode in range(1, int(col.get("nmodeadv2")) + 1):                         # :178-184
        rhsmod_salt(int(col.L.orc_col_modeadv(col.h, imode, 2)), float(adv[1, imode - 1]), dto, kmixe, dmv[kmixe], NZ,
                    rhs, zm, hm)
    sinc = np.zeros(NZP1 + 1)                                                     # :188
    if k_.L_SFCORR_WITHZ and not k_.L_SFCORR:                                     # :190-194
        sinc[kk] = dto * col._arr("sfcorr_withz")[kk]
    if k_.L_RELAX_SAL:                                                            # :200-207
        sinc[kk] = sinc[kk] + dto * col.get("relax_sal") * (col._arr("sal_clim")[kk] - Xo[1])
    rhs[1:] = rhs[1:] + sinc[1:NZ + 1]                                            # :208-209
    col._arr("sinc_fcorr")[kk] = sinc[kk]
    col._arr("scorr")[kk] = sinc[kk] / dto                                        # :213
    snew, pz = tridmat(cu, cc, cl, rhs, Xo[1, NZ], NZ)
    pivot |= pz
    X[1, :] = snew
    return pivot


# --------------------------------------------------------------------------- check_profile
def check_profile2(col, cf):
    """mckpp_physics_overrides_check_profile (overrides.F90:42-125) on the column, array-at-a-time."""
    k_ = cf.consts
    nzp1 = col.nzp1
    n1 = nzp1 + 1
    U, X = col.U, col.X
    status = int(col.get("status"))
    comp_flag = bool(col.get("comp_flag"))
    tclim, sclim = col._arr("ocnT_clim")[1:n1], col._arr("sal_clim")[1:n1]
    if comp_flag:                                                                 # :57-78
        if k_.have_ocnT_file and k_.have_sal_file:
            X[0, :] = tclim
            X[1, :] = sclim
        U[:, :] = col._arr("U_init").reshape(2, n1)[:, 1:]
        col.set("reset_flag", 999)
        status |= 4
    l_ocean = bool(col.get("l_ocean"))
    if l_ocean and k_.L_NO_FREEZE:                                                # :85-94
        cold = X[0] < -1.8
        tinc = col._arr("tinc_fcorr")
        tinc[1:n1] = np.where(cold, tinc[1:n1] + (-1.8 - X[0]), tinc[1:n1])
        X[0, :] = np.where(cold, -1.8, X[0])
        ff = col.get("freeze_flag")
        for _ in range(int(cold.sum())):                       # added once per level, in level order
            ff = ff + 1.0 / float(nzp1)
        col.set("freeze_flag", ff)
    if l_ocean and k_.L_NO_ISOTHERM:                                              # :102-121
        zm = cf.zm
        nb = int(k_.iso_bot)
        dz = zm[1:nb] - zm[0:nb - 1]                           # j = 2..iso_bot
        terms = np.abs(X[0, 1:nb] - X[0, 0:nb - 1]) * dz
        dtdz_total = dz_total = 0.
        for t, d in zip(terms, dz):
            dtdz_total = dtdz_total + float(t)
            dz_total = dz_total + float(d)
        dtdz_total = dtdz_total / dz_total
        if abs(dtdz_total) < k_.iso_thresh:
            X[0, :] = tclim
            X[1, :] = sclim
            col.set("reset_flag", (-1.) * col.get("reset_flag"))
            status |= 32
    else:
        col.set("reset_flag", 0)                                                  # :123
    col.set("status", status)


def _another_pass(iter_, iconv, hmixn, hmixe, itermax, cap):
    """ocnstep_mod.F90:170-183: after a convergence pass, is there a `goto 45`?
    Returns (again, hit_safety_cap)."""
    if iconv < 3:
        if iter_ < itermax:
            return True, False
        if hmixn > hmixe:           # 'use shallower hmix' branch: keeps iterating past itermax
            if iter_ >= itermax + cap:
                return False, True  # oracle/GPU extension, never reached by a healthy column
            return True, False
    return False, False


def ocnstep(col: Column, cf, probe=None, second_ocnint=False):
    """mckpp_physics_ocnstep (ocnstep_mod.F90:43-357) on `col`.  probe(col) is called after every vmix;
    second_ocnint: integrate with ocnint2 (the second reading of ocnint and the solvers) instead of the C oracle's.
    Returns (iter, nreint)."""
    k_ = cf.consts
    zm, hm, dm = cf.zm, cf.hm, cf.dm
    NZ, NZP1 = col.nz, col.nzp1
    lam = 0.5
    comp_iter_max = 10
    U, X, Us, Xs = col.U, col.X, col.Us, col.Xs
    Uo, Xo = U.copy(), X.copy()                                   # :82-83
    comp_flag, nint = True, 0                                     # :84-85 (reset_flag counts integrations)
    col.set("dampu_flag", 0.0); col.set("dampv_flag", 0.0)        # :86-87
    old, new = int(col.get("old")), int(col.get("new"))
    f = col.get("f")
    status = int(col.get("status"))
    it = 0
    hmixe = hmixn = 0.0
    kmixe = kmixn = 0
    while comp_flag and nint <= comp_iter_max:                    # :89
        if old < 0 or old > 1:                                    # :93-97
            status |= 64; old = new
        if new < 0 or new > 1:                                    # :98-102
            status |= 64; new = old
        col.set("old", old); col.set("new", new)
        U[:] = 2. * Us[new] - Us[old]                             # :103-105
        X[:] = 2. * Xs[new] - Xs[old]                             # :108-110
        Ux, Xx = U.copy(), X.copy()
        it, iconv = 0, 0                                          # :116-117
        while True:
            U[:] = lam * Ux + (1 - lam) * U; Ux[:] = U            # :125-126 / :144-145
            X[:] = lam * Xx + (1 - lam) * X; Xx[:] = X            # :129-130 / :148-149
            h, kk = col.vmix()                                    # :133 / :152
            if probe is not None:
                probe(col)
            if second_ocnint:
                if ocnint2(col, cf, kk, Uo, Xo):
                    status |= 8
            else:
                col.ocnint(kk, Uo, Xo)                            # :134 / :153
                status |= int(col.get("status")) & 8
            if it < 3:                                            # the DO iter=0,2 passes; iter is 3 after them
                hmixe, kmixe = h, kk
                it += 1
                continue
            hmixn, kmixn = h, kk
            it += 1                                               # :154
            tol = k_.hmixtolfrac * hm[kmixn - 1]                  # :157
            if kmixn == NZP1:
                tol = k_.hmixtolfrac * hm[NZ - 1]                 # :158
            iconv = 0 if abs(hmixn - hmixe) > tol else iconv + 1  # :159-169
            again, capped = _another_pass(it, iconv, hmixn, hmixe, k_.itermax, 1000)
            if capped:
                status |= 16
            if not again:
                break
            hmixe, kmixe = hmixn, kmixn                           # :172-173 / :178-179
        if it > k_.itermax + 1:                                   # :184
            status |= 1
        # ---- instability trap :200-227
        bad = (np.abs(U[0, :NZ]) >= 10) | (np.abs(U[1, :NZ]) >= 10) | (np.abs(X[0, :NZ] - X[0, 1:NZP1]) >= 10)
        comp_flag = bool(bad.any())
        for _ in range(int(bad.sum())):                           # f*1.01 once per offending level :205
            f = f * 1.01
        if not comp_flag:
            sq = [(U[0] - Uo[0]) * (U[0] - Uo[0]) * hm[:NZP1] / dm[NZ], (U[1] - Uo[1]) * (U[1] - Uo[1]) * hm[:NZP1] / dm[NZ],
                  (X[0] - Xo[0]) * (X[0] - Xo[0]) * hm[:NZP1] / dm[NZ], (X[1] - Xo[1]) * (X[1] - Xo[1]) * hm[:NZP1] / dm[NZ]]
            for terms in sq:                                      # :210-218, summed in level order
                acc = 0.0
                for t in terms:
                    acc = acc + float(t)
                if math.sqrt(acc) >= 1:                           # :221-225
                    comp_flag = True
                    f = f * 1.01
        col.set("f", f)
        nint += 1                                                 # :228
        if nint > comp_iter_max:
            status |= 2                                           # :229
    col.set("comp_flag", 1 if comp_flag else 0)
    col.set("reset_flag", nint)
    # ---- diagnostic fluxes :242-256
    deltaz = 0.5 * (hm[:NZ] + hm[1:NZP1])                         # :243
    ks = slice(1, NZ + 1)
    for n in (0, 1):
        col.wX[n, ks] = -col.difs[ks] * ((X[n, :NZ] - X[n, 1:NZP1]) / deltaz - col.ghat[ks] * col.wX[n, 0])    # :245-246
    if k_.LDD:
        col.wX[0, ks] = -col.dift[ks] * ((X[0, :NZ] - X[0, 1:NZP1]) / deltaz - col.ghat[ks] * col.wX[0, 0])    # :248-250
    col.wX[2, ks] = k_.grav * (col.talpha[ks] * col.wX[0, ks] - col.sbeta[ks] * col.wX[1, ks])                 # :251-252
    for n in (0, 1):
        col.wU[n, ks] = -col.difm[ks] * (U[n, :NZ] - U[n, 1:NZP1]) / deltaz                                    # :254
    # ---- results :305-315
    col.set("hmix", hmixn); col.set("kmix", kmixn)
    col.set("uref", U[0, 0]); col.set("vref", U[1, 0]); col.set("Tref", X[0, 0])
    col.set("Ssurf", col.get("SSref") if k_.L_SSref else X[1, 0] + col.get("Sref"))
    # ---- current damping :317-340
    if k_.L_DAMP_CURR:
        damp = [0.0, 0.0]
        r = k_.dt_uvdamp * (86400. / k_.dto)
        for k in range(NZP1):
            for l in (0, 1):
                a = 0.99 * abs(U[l, k])
                b = U[l, k] * U[l, k] / r
                Ui = min(a, b)
                if b < a:
                    damp[l] = damp[l] + 1.0 / float(NZP1)
                U[l, k] = U[l, k] - math.copysign(abs(Ui), U[l, k])
        col.set("dampu_flag", damp[0]); col.set("dampv_flag", damp[1])
    # ---- rotate the time levels :343-353
    old = new
    new = 1 - old
    col.set("old", old); col.set("new", new)
    col.hmixd[new] = hmixn
    Us[new] = U
    Xs[new] = X
    col.set("status", status)
    return it, nint


def physics_driver(orc: "oracle_lib.Oracle", cf, fields, ntime, probe=None, second_ocnint=False):
    """mckpp_physics_driver (physics_driver_mod.F90:15-73) with the second reading of ocnstep."""
    L = _lib()
    col = Column(orc)
    try:
        for ipt in range(1, int(orc.c.npts) + 1):
            if not fields["run_physics"][ipt - 1]:
                continue
            L.orc_col_load(col.h, C.byref(orc.c), C.byref(orc.s), ipt, int(ntime))        # :49
            it, nreint = ocnstep(col, cf, probe, second_ocnint)                           # :53
            if second_ocnint:
                check_profile2(col, cf)                                                   # :56, second reading
            else:
                L.orc_col_check_profile(col.h, C.byref(orc.c))                            # :56
            L.orc_col_store(col.h, C.byref(orc.c), C.byref(orc.s), ipt, it, nreint)       # :59
    finally:
        col.close()
    if cf.consts.L_VARY_BOTTOM_TEMP:                                                      # :68-70, overrides.F90:12-24
        nzp1 = cf.dims.nzp1
        tinc = fields["bottom_temp"] - fields["X"][:, nzp1 - 1, 0]
        fields["tinc_fcorr"][:, nzp1 - 1] = tinc
        fields["ocnTcorr"][:, nzp1 - 1] = tinc * fields["rho"][:, nzp1] * fields["cp"][:, nzp1] / cf.consts.dto
        fields["X"][:, nzp1 - 1, 0] = fields["bottom_temp"]


def initialize_ocean_model2(orc: "oracle_lib.Oracle", cf, fields, probe=None):
    """The per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL (initialize_ocean.F90:54-104): initial vmix with
    L_INITFLAG, hmix/kmix/Tref, the initial diagnostic fluxes, the two saved time levels.  vmix is the C oracle's
    (probe(col) is called after it, as in ocnstep)."""
    L = _lib()
    k_ = cf.consts
    col = Column(orc)
    try:
        NZ, NZP1 = col.nz, col.nzp1
        hm = cf.hm
        for ipt in range(1, int(orc.c.npts) + 1):
            if not fields["run_physics"][ipt - 1]:
                continue
            L.orc_col_load(col.h, C.byref(orc.c), C.byref(orc.s), ipt, 0)       # ntime = 0 (time_control.F90:31)
            U, X = col.U, col.X
            col.set("l_initflag", 1)                                              # :58
            h, kk = col.vmix()                                                    # :59
            if probe is not None:
                probe(col)
            col.set("l_initflag", 0)                                              # :60
            col.set("hmix", h); col.set("kmix", kk); col.set("Tref", X[0, 0])     # :61-63
            deltaz = 0.5 * (hm[:NZ] + hm[1:NZP1])                                 # :66
            ks = slice(1, NZ + 1)
            for n in (0, 1):                                                      # :67-71
                col.wX[n, ks] = -col.difs[ks] * ((X[n, :NZ] - X[n, 1:NZP1]) / deltaz - col.ghat[ks] * col.wX[n, 0])
            if k_.LDD:                                                            # :72-73
                col.wX[0, ks] = -col.dift[ks] * ((X[0, :NZ] - X[0, 1:NZP1]) / deltaz - col.ghat[ks] * col.wX[0, 0])
            col.wX[2, ks] = k_.grav * (col.talpha[ks] * col.wX[0, ks] - col.sbeta[ks] * col.wX[1, ks])   # :74-75
            for n in (0, 1):                                                      # :76-79
                col.wU[n, ks] = -col.difm[ks] * (U[n, :NZ] - U[n, 1:NZP1]) / deltaz
            col.set("old", 0); col.set("new", 1)                                  # :85-86
            col.hmixd[0] = h; col.hmixd[1] = h                                    # :88-89
            col.Us[0] = U; col.Us[1] = U                                          # :90-99
            col.Xs[0] = X; col.Xs[1] = X
            # kpp_1d_fields is INTENT(OUT) in 3dto1d: the flags ocnstep would set are undefined here; like the
            # oracle, keep the 3-D values
            for nm in ("reset_flag", "dampu_flag", "dampv_flag"):
                col.set(nm, fields[nm][ipt - 1])
            L.orc_col_store(col.h, C.byref(orc.c), C.byref(orc.s), ipt, 0, 0)     # :101
    finally:
        col.close()


def bldepth_probe(cf, log):
    """probe for ocnstep(): replays the bldepth call the C oracle just made through the second reading and
    appends (ours, theirs) to `log`."""
    n1 = cf.dims.nzp1 + 1
    zm = np.concatenate([[0.0], cf.zm])
    hm = np.concatenate([[0.0], cf.hm])
    wmt = np.asarray(cf.wmt).reshape(50, 892).T if np.asarray(cf.wmt).ndim == 1 else np.asarray(cf.wmt)
    wst = np.asarray(cf.wst).reshape(50, 892).T if np.asarray(cf.wst).ndim == 1 else np.asarray(cf.wst)

    def probe(col: Column):
        g = col.get
        ours = bldepth(zm, hm, cf.consts.vonk, wmt, wst, col._arr("dVsq")[:n1], col._arr("Ritop")[:n1],
                       col._arr("dbloc"), col._arr("swfrac")[:n1], g("dbg_ustar"), g("dbg_Bo"), g("dbg_Bosol"),
                       g("f"), g("ocdepth"), int(g("jerlov")), bool(g("l_initflag")))
        theirs = (g("dbg_hbl"), int(g("dbg_kbl")), g("dbg_bfsfc"), g("dbg_stable"), g("dbg_caseA"))
        log.append((ours, theirs))
    return probe
