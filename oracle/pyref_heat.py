"""Second, independent transliteration of the reference's heat / ablation step -- pure Python, tiny grids only.

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/heat_oracle.c (dict-of-tuples arrays with Fortran
index triples instead of flat buffers, one function per Fortran routine) so that a slip in one transliteration
shows up as a mismatch against the other; both use IEEE doubles and the same libm, so they must agree bit
for bit.  Follows /root/reference/src: thermalConst_mod.f90:1-88, 3dFD.f90:21-230 (numproc = 1), :233-309,
:312-361, :365-421, :424-466, mcpolar.f90:65-71,123-140,174.  Not pinned by a compiled reference; oracle/heat_oracle.c, which
this file must equal bit for bit, reproduces what oracle/f90interp.py computes from the reference's own text
(tests/test_oracle_reference_heat.py).
"""
from __future__ import annotations

import math

WATER0 = 0.75
PROTEIN = 1.0 - WATER0
AIR_CP = 1.006e3
LW = 2256.0e3


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:          # C's exp returns +inf (the explicit scheme diverges once voxels turn to air)
        return math.inf


def air_thermal_cond(T):
    return -0.188521 * _exp(-0.000367259 * (T - 273.15)) + 0.212453


def air_density(T):
    return 101.325e3 / (287.058 * T)


def skin_density(w):
    return 1000.0 / (w + 0.649 * PROTEIN)


def skin_heat_cap(w):
    return 1000.0 * (4.2 * w + 1.09 * PROTEIN)


def skin_thermal_cond(w, rho):
    return rho * (6.28e-4 * w + 1.17e-4 * PROTEIN)


class Heat:
    """Module Heat (3dFD.f90) + the driver lines that feed it, one MPI rank."""

    def __init__(self, n, xmax, ymax, zmax, power=70.0, energy=400.0, total_time=2.0, loops=1, rep_rate=1e7,
                 pulses_to_do=1, pulsetype="gaussian", kappa0=680.0):
        self.n, self.xmax, self.ymax, self.zmax = n, xmax, ymax, zmax
        self.power, self.loops, self.rep_rate, self.pulses_to_do, self.pulsetype = power, loops, rep_rate, pulses_to_do, pulsetype
        halo = range(0, n + 2)
        inner = range(1, n + 1)
        self.all_h = [(i, j, k) for k in halo for j in halo for i in halo]
        self.all_i = [(i, j, k) for k in inner for j in inner for i in inner]       # Fortran k,j,i sweep order
        # mcpolar.f90:65-71
        self.time = self.pulse_count = self.rep_count = 0.0
        self.laser_on, self.laser_flag, self.pulse_flag, self.pulses_done = 1.0, True, False, 0
        # gridset.f90:33-45
        self.rhokap = {v: 0.0 for v in self.all_h}
        for v in self.all_i:
            self.rhokap[v] = kappa0
        # mcpolar.f90:123-129
        self.temp = {}
        for (i, j, k) in self.all_h:
            self.temp[(i, j, k)] = 25.0 + 273.0 if k in (0, n + 1) else 5.0 + 273.0
        # initThermalCoeff, 3dFD.f90:249-293
        self.dx = (2.0 * xmax * 1.0e-2) / (n + 2.0)
        self.dy = (2.0 * ymax * 1.0e-2) / (n + 2.0)
        self.dz = (2.0 * zmax * 1.0e-2) / (n + 2.0)
        cp0, rho0 = skin_heat_cap(WATER0), skin_density(WATER0)
        kap0 = skin_thermal_cond(WATER0, rho0)
        alpha0 = kap0 / (rho0 * skin_heat_cap(WATER0))
        self.Q = {v: 0.0 for v in self.all_i}
        self.water = {v: WATER0 for v in self.all_i}
        self.tissue = {v: 0.0 for v in self.all_i}
        self.thres = {(v, m): 0.0 for v in self.all_i for m in (1, 2, 3)}
        self.alpha = {v: alpha0 for v in self.all_h}
        for j in halo:
            for i in halo:
                self.alpha[(i, j, n + 1)] = air_thermal_cond(25.0 + 273.0) / (air_density(25.0 + 273.0) * AIR_CP)
        self.kappa = {v: air_thermal_cond(25.0 + 273.0) for v in self.all_h}
        for v in self.all_i:
            self.kappa[v] = skin_thermal_cond(WATER0, rho0)
        self.density = {v: rho0 for v in self.all_h}
        self.heatcap = {v: cp0 for v in self.all_h}
        constd = (1.0 / (self.dx * self.dx)) + (1.0 / (self.dy * self.dy)) + (1.0 / (self.dz * self.dz))
        self.delt = 1.0 / (1.0 * alpha0 * constd)
        self.coeff = {v: 0.0 for v in self.all_h}
        for v in self.all_i:
            self.coeff[v] = alpha0 * self.delt / kap0
        self.pulselength = (energy * 1.0e-3 * float(9 * 9)) / power
        self.vol = (2.0 * xmax * 1.0e-2 / n) * (2.0 * ymax * 1.0e-2 / n) * (2.0 * zmax * 1.0e-2 / n)
        self.mass = rho0 * self.vol
        self.qvapor = LW * self.mass
        self.real_pulse = {"tophat": self.pulselength, "gaussian": 20000.0 * self.pulselength,
                           "triangular": 2.0 * self.pulselength}[pulsetype]
        self.total_time = total_time
        if pulsetype == "gaussian":                                           # mcpolar.f90:134-137
            self.total_time = 2.0 * self.pulselength * (2.0 * math.sqrt(2.0 * math.log(2.0)))
            self.real_pulse = self.total_time
        elif int(self.total_time / self.delt) <= int(self.real_pulse / self.delt):
            self.total_time = self.delt * (self.real_pulse / self.delt + 2000.0)

    # 3dFD.f90:365-421
    def get_pwr(self):
        if self.pulsetype == "gaussian":
            fact = 2.0 * math.sqrt(2.0 * math.log(2.0))
            mu, sig = fact * self.pulselength, self.pulselength / fact
            return self.power * math.exp(-((self.time - mu) * (self.time - mu)) / (2.0 * (sig * sig)))
        if self.pulsetype == "tophat":
            return self.power if self.laser_flag else 0.0
        m, c = self.power / self.pulselength, 2.0 * self.power
        if not self.laser_flag:
            return 0.0
        if self.pulse_flag or self.time >= self.pulselength:
            self.pulse_flag = True
            p = -m * self.time + c
            return p if p >= 0.0 else 0.0
        return m * self.time

    # mcpolar.f90:174
    def scale(self, jmean, nphotons_total):
        n = self.n
        f = (self.get_pwr() / 81.0) / (nphotons_total * (2.0 * self.xmax * 1.0e-2 / n) * (2.0 * self.ymax * 1.0e-2 / n)
                                       * (2.0 * self.zmax * 1.0e-2 / n))
        return {v: jmean[v] * f for v in self.all_i}

    # 3dFD.f90:21-230 with numproc = 1
    def sim_3d(self, jmean):
        t0 = dict(self.temp)
        tn = dict(t0)
        if self.pulselength < self.delt:
            self.delt = self.pulselength / 100.0
        for _ in range(self.loops):
            for (i, j, k) in self.all_i:
                c = (i, j, k)
                u = []
                for (p, m, h) in (((i, j, k + 1), (i, j, k - 1), self.dz), ((i, j + 1, k), (i, j - 1, k), self.dy),
                                  ((i + 1, j, k), (i - 1, j, k), self.dx)):
                    kp, km = 0.5 * (self.kappa[c] + self.kappa[p]), 0.5 * (self.kappa[c] + self.kappa[m])
                    dp, dm = 0.5 * (self.density[c] + self.density[p]), 0.5 * (self.density[c] + self.density[m])
                    hp, hm = 0.5 * (self.heatcap[c] + self.heatcap[p]), 0.5 * (self.heatcap[c] + self.heatcap[m])
                    a = 0.5 * (km / (dm * hm)) * (1.0 / (h * h))
                    d = 0.5 * (kp / (dp * hp)) * (1.0 / (h * h))
                    b = 0.5 * (a + d)
                    u.append(a * t0[m] - 2.0 * b * t0[c] + d * t0[p])
                u_zz, u_yy, u_xx = u
                dT = self.delt * (u_xx + u_yy + u_zz)
                dE = self.laser_on * jmean[c] * self.delt * self.vol + self.heatcap[c] * self.mass * dT
                if tn[c] >= 100.0 + 273.0 and self.Q[c] < self.qvapor:
                    if dE > 0.0:
                        self.Q[c] = min(self.Q[c] + dE, self.qvapor)
                        tn[c] = 100.0 + 273.0
                    else:
                        tn[c] = tn[c] + dT + self.laser_on * self.coeff[c] * jmean[c]
                else:
                    tn[c] = tn[c] + dT + self.laser_on * self.coeff[c] * jmean[c]
            t0 = dict(tn)
            if self.pulse_count >= self.real_pulse and self.laser_flag:
                self.laser_flag, self.laser_on, self.pulse_count, self.rep_count = False, 0.0, 0.0, 0.0
                self.pulses_done += 1
            elif self.rep_count >= self.rep_rate and not self.laser_flag and self.pulses_done < self.pulses_to_do:
                self.laser_flag, self.laser_on, self.pulse_count, self.rep_count = True, 1.0, 0.0, 0.0
            self.pulse_count += self.delt
            self.rep_count += self.delt
            self.time += self.delt
        n = self.n
        for (i, j, k) in self.all_h:
            if 1 <= k <= n:
                self.temp[(i, j, k)] = t0[(i, j, k)]

    # 3dFD.f90:424-466
    def arrhenius(self):
        A, dE, R = 3.1e98, 6.3e5, 8.314
        for v in self.all_i:
            T = self.temp[v]
            if 43.0 + 273.0 <= T < 100.0 + 273.0 and self.rhokap[v] >= 0.0:
                self.tissue[v] = self.tissue[v] + self.delt * A * _exp(-dE / (R * T))
            if self.thres[(v, 1)] == 0.0 and self.tissue[v] >= 0.53:
                self.thres[(v, 1)] = self.time
            elif self.thres[(v, 2)] == 0.0 and self.tissue[v] >= 1.0:
                self.thres[(v, 2)] = self.time
            elif self.thres[(v, 3)] == 0.0 and self.tissue[v] >= 10000.0:
                self.thres[(v, 3)] = self.time

    # 3dFD.f90:312-361
    def setup_thermal_coeff(self, ablate_temp):
        for v in self.all_i:
            w = WATER0 - WATER0 * (self.Q[v] / self.qvapor)
            self.water[v] = max(min(min(w, WATER0), self.water[v]), 0.0)
        for (i, j, k) in self.all_i:                                          # in-place, sweep order matters
            c = (i, j, k)
            if self.temp[c] >= ablate_temp + 273.0:
                self.rhokap[c] = 0.0
            elif self.rhokap[c] > 0.0:
                self.density[c] = skin_density(self.water[c])
                self.rhokap[c] = self.water[c] * 510.0 + 170.0
                self.heatcap[c] = skin_heat_cap(self.water[c])
                self.kappa[c] = skin_thermal_cond(self.water[c], self.density[c])
                self.coeff[c] = self.delt / (self.density[c] * self.heatcap[c])
            summ = (self.rhokap[(i, j, k + 1)] + self.rhokap[(i, j + 1, k)] + self.rhokap[(i + 1, j, k)]
                    + self.rhokap[(i, j, k - 1)] + self.rhokap[(i, j - 1, k)] + self.rhokap[(i - 1, j, k)])
            if summ == 0.0:
                self.rhokap[c] = 0.0
            if self.rhokap[c] <= 0.01:
                self.density[c] = air_density(self.temp[c])
                self.heatcap[c] = 1.006e3
                self.rhokap[c] = 0.0
                self.kappa[c] = air_thermal_cond(self.temp[c])
                self.alpha[c] = self.kappa[c] / (self.density[c] * self.heatcap[c])
                self.coeff[c] = self.delt / (air_density(self.temp[c]) * self.heatcap[c])
