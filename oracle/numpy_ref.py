"""Independent numpy float32 restatement of the lattice step — TEST INFRASTRUCTURE ONLY.

A second, separately written restatement of the reference's CPU algorithm (dense adjacency,
vectorised over postsynaptic neurons, explicit ascending-presynaptic accumulation order) used to
(a) pin the C oracle and (b) generate the golden fixtures under tests/golden/ (see
tests/golden/make_golden.py).  Only small cases: the gather loops are pure Python.

Every elementwise op is IEEE f32 (numpy float32 arrays), so +,-,*,/ results are bit-identical to
the C oracle; np.exp / np.power may differ from glibc expf/powf in the last ulp, so models that
use them (HH, NMDA, AdEx, Destexhe, STDP) are compared with a tolerance.

Reference citations are relative to /root/reference/backend/src/neuron/.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
LIF, QIF, ADLIF, ADEX, IZH, LEAKY_IZH, SIMPLE_LIF, HH, BCM_IZH = range(9)


def _f(x):
    return np.asarray(x, dtype=f32)


class DenseLattice:
    """One lattice, dense graph conn[pre, post] / w[pre, post], fields as f32 vectors."""

    def __init__(self, model, n, conn, w, fields):
        self.model = model
        self.n = n
        self.conn = np.asarray(conn, dtype=bool)
        self.w = _f(w).copy()
        self.f = {k: (_f(v).copy() if np.asarray(v).dtype.kind == "f" else np.asarray(v).copy()) for k, v in fields.items()}
        self.electrical, self.chemical = True, False
        self.do_plasticity = False
        self.use_bcm = False   # plasticity rule BCM (plasticity/mod.rs:80-112) instead of STDP; BCM_IZH lattices only
        self.bcm = dict(decay=f32(0.1), average_scalar=f32(0.1), dt=f32(0.1))
        self.stdp = dict(a_plus=f32(2), a_minus=f32(2), tau_plus=f32(4.5), tau_minus=f32(4.5), dt=f32(0.1))
        self.clock = 0
        self.ntk = 0  # 0 approximate, 1 destexhe
        self.rck = 0
        # chemical state: dicts of (n,3) arrays
        self.nt_flags = np.zeros((n, 3), bool)
        self.nt = dict(t=np.zeros((n, 3), f32), t_max=np.ones((n, 3), f32), clearance_constant=np.full((n, 3), 0.01, f32),
                       v_p=np.full((n, 3), 2, f32), k_p=np.full((n, 3), 5, f32))
        self.rc_flags = np.zeros((n, 3), bool)
        self.rc = dict(r=np.zeros((n, 3), f32), alpha=np.ones((n, 3), f32), beta=np.ones((n, 3), f32),
                       g=np.tile(_f([1.0, 0.6, 1.2]), (n, 1)), e=np.tile(_f([0.0, 0.0, -80.0]), (n, 1)),
                       mg=np.full((n, 3), 0.3, f32), current=np.zeros((n, 3), f32))
        self.is_spiking = np.zeros(n, bool)
        self.lft = np.full(n, -1, np.int64)
        self.was_increasing = np.zeros(n, bool)
        self.v_hist, self.s_hist = [], []

    # ---- inputs (mod.rs:702-754) --------------------------------------------------------------
    def electrical_inputs(self):
        v, gap = self.f["current_voltage"], self.f["gap_conductance"]
        acc = np.zeros(self.n, f32)
        for pre in range(self.n):
            m = self.conn[pre]
            if m.any():
                acc[m] = acc[m] + (gap[m] * (v[pre] - v[m])) * self.w[pre][m]
        cnt = self.conn.sum(axis=0)
        return acc / np.maximum(cnt, 1).astype(f32)

    def chemical_inputs(self):
        acc = np.zeros((self.n, 3), f32)
        cnt = np.zeros((self.n, 3), np.int64)
        for pre in range(self.n):
            m = self.conn[pre]
            for ty in range(3):
                if self.nt_flags[pre, ty] and m.any():
                    acc[m, ty] = acc[m, ty] + self.nt["t"][pre, ty] * self.w[pre][m]
                    cnt[m, ty] += 1
        has = cnt > 0
        out = np.zeros((self.n, 3), f32)
        out[has] = acc[has] / cnt[has].astype(f32)
        return out, has

    # ---- kinetics (iterate_and_spike/mod.rs:147-150, 192-196, 403-406, 434-437, 1101-1137, 1286-1304)
    def _receptors(self, t_in, has, v, dt, c_m):
        upd = has & self.rc_flags
        if self.rck == 0:
            self.rc["r"][upd] = t_in[upd]
        else:
            r, a, b = self.rc["r"], self.rc["alpha"], self.rc["beta"]
            new = r + ((a * t_in) * (f32(1) - r) - b * r) * dt[:, None]
            r[upd] = new[upd]
        g, e, r, mg = self.rc["g"], self.rc["e"], self.rc["r"], self.rc["mg"]
        cur = self.rc["current"]
        vv = v[:, None]
        plain = (g * r) * (vv - e)
        nm = (((f32(1) / (f32(1) + ((np.exp(f32(-0.062) * vv) * mg) / f32(3.75)))) * g) * r) * (vv - e)
        for ty in range(3):
            m = self.rc_flags[:, ty]
            cur[m, ty] = (nm if ty == 1 else plain)[m, ty]
        total = np.zeros(self.n, f32)
        for ty in range(3):
            m = self.rc_flags[:, ty]
            total[m] = total[m] + cur[m, ty]
        return total * (dt / c_m)

    def _release(self, v, spiking_prev, dt):
        t, tm = self.nt["t"], self.nt["t_max"]
        if self.ntk == 0:
            new = t + (((dt[:, None] * -self.nt["clearance_constant"]) * t) + spiking_prev[:, None].astype(f32) * tm)
            new = np.minimum(tm, np.maximum(new, f32(0)))
        else:
            new = tm / (f32(1) + np.exp(-(v[:, None] - self.nt["v_p"]) / self.nt["k_p"]))
        t[self.nt_flags] = new[self.nt_flags]

    # ---- one step (mod.rs:884-982 with deferred STDP, see oracle/snn_oracle.h) -------------------
    def step(self):
        F, m = self.f, self.model
        n = self.n
        I = self.electrical_inputs() if self.electrical else np.zeros(n, f32)
        if self.chemical:
            t_in, has = self.chemical_inputs()
        v, dt = F["current_voltage"], F["dt"]
        c_m = F["c_m"]
        rc_dv = np.zeros(n, f32)
        if self.chemical:
            rc_dv = self._receptors(t_in, has, v, dt, c_m)
        elif m == HH:
            total = np.zeros(n, f32)
            for ty in range(3):
                mm = self.rc_flags[:, ty]
                total[mm] = total[mm] + self.rc["current"][mm, ty]
            rc_dv = total * (dt / c_m)
        spiking_prev = self.is_spiking.copy()
        spike = np.zeros(n, bool)
        if m == BCM_IZH:
            # BCMIzhikevichNeuron bookkeeping (integrate_and_fire/mod.rs:1458-1467, 1485-1494): before the update, with the
            # previous step's spike flag; the chemical variant divides by the window only
            F["num_spikes"] = F["num_spikes"] + spiking_prev.astype(F["num_spikes"].dtype)
            clk = F["firing_rate_clock"] + dt
            roll = clk >= F["firing_rate_window"]
            nsp = F["num_spikes"].astype(f32)
            cur = nsp / F["firing_rate_window"] if self.chemical else nsp / (F["firing_rate_window"] * dt)
            per = F["period"].astype(f32)
            avg = F["average_activity"]
            avg2 = avg - avg / per
            avg2 = avg2 + cur / per
            F["current_activity"] = np.where(roll, cur, F["current_activity"]).astype(f32)
            F["average_activity"] = np.where(roll, avg2, avg).astype(f32)
            F["firing_rate_clock"] = np.where(roll, f32(0), clk).astype(f32)
        if m in (IZH, LEAKY_IZH, BCM_IZH):
            w = F["w_value"]
            if m != LEAKY_IZH:
                dv = (((((f32(0.04) * (v * v)) + (f32(5) * v)) + f32(140)) - w) + I) * (dt / c_m)
            else:
                dv = (((((f32(0.04) * (v * v)) + (f32(5) * v)) + f32(140)) - (w * (v - F["e_l"]))) + I) * (dt / c_m)
            dw = (F["a"] * (F["b"] * v - w)) * (dt / F["tau_m"])
            v = v + (dv + (-rc_dv)) if self.chemical else v + dv
            w = w + dw
            self._release(v, spiking_prev, dt)
            spike = v >= F["v_th"]
            v = np.where(spike, F["c"], v)
            w = np.where(spike, w + F["d"], w)
            F["w_value"] = w.astype(f32)
        elif m in (LIF, QIF, ADLIF, ADEX):
            adapt = m in (ADLIF, ADEX)
            if m == LIF:
                dv = ((F["leak_constant"] * (v - F["e_l"])) + (F["integration_constant"] * (I / F["g_l"]))) * (dt / F["tau_m"])
            elif m == QIF:
                dv = (((F["alpha"] * (v - F["v_reset"])) * (v - F["v_c"])) + F["integration_constant"] * I) * (dt / F["tau_m"])
            else:
                w = F["w_value"]
                base = F["leak_constant"] * (v - F["e_l"])
                if m == ADEX:
                    base = base + (F["slope_factor"] * np.exp((v - F["v_th"]) / F["slope_factor"]))
                dv = ((base + (F["integration_constant"] * (I / F["g_l"]))) - (w / F["g_l"])) * (dt / c_m)
                dw = (F["alpha"] * (v - F["e_l"]) - w) * (dt / F["tau_m"])
            v = v + (dv + (-rc_dv)) if self.chemical else v + dv
            if adapt:
                w = w + dw
            self._release(v, spiking_prev, dt)
            refr = F["refractory_count"]
            in_ref = refr > 0
            spike = (~in_ref) & (v >= F["v_th"])
            v = np.where(in_ref | spike, F["v_reset"], v)
            refr = np.where(in_ref, refr - f32(1), np.where(spike, F["tref"] / dt, refr))
            F["refractory_count"] = refr.astype(f32)
            if adapt:
                F["w_value"] = np.where(spike, w + F["beta"], w).astype(f32)
        elif m == SIMPLE_LIF:
            dv = (F["g"] * (v - F["e"]) + I) * dt
            v = v + (dv + (-rc_dv)) if self.chemical else v + dv
            self._release(v, spiking_prev, dt)
            spike = v >= F["v_th"]
            v = np.where(spike, F["v_reset"], v)
        else:  # HH (hodgkin_huxley/mod.rs:156-241, ion_channels/mod.rs:40-44, 219-235, 268-281, 310-312)
            last = v.copy()
            ms, hs, ns = F["m"], F["h"], F["n"]
            ma = f32(0.1) * ((v + f32(40)) / (f32(1) - np.exp(-(v + f32(40)) / f32(10))))
            mb = f32(4) * np.exp(-(v + f32(65)) / f32(18))
            ha = f32(0.07) * np.exp(-(v + f32(65)) / f32(20))
            hb = f32(1) / (np.exp(-(v + f32(35)) / f32(10)) + f32(1))
            ms = ms + dt * (ma * (f32(1) - ms) - mb * ms)
            hs = hs + dt * (ha * (f32(1) - hs) - hb * hs)
            i_na = ((np.power(ms, f32(3)) * hs) * F["g_na"]) * (v - F["e_na"])
            na = (f32(0.01) * (v + f32(55))) / (f32(1) - np.exp(-(v + f32(55)) / f32(10)))
            nb = f32(0.125) * np.exp(-(v + f32(65)) / f32(80))
            ns = ns + dt * (na * (f32(1) - ns) - nb * ns)
            i_k = (np.power(ns, f32(4)) * F["g_k"]) * (v - F["e_k"])
            i_kl = F["g_k_leak"] * (v - F["e_k_leak"])
            i_sum = I - ((i_na + i_k) + i_kl)
            v = v + ((dt * i_sum) / c_m - rc_dv)
            self._release(v, spiking_prev, dt)
            inc = last < v
            spike = (v > F["v_th"]) & self.was_increasing & ~inc
            self.was_increasing = inc
            F["m"], F["h"], F["n"] = ms.astype(f32), hs.astype(f32), ns.astype(f32)
        F["current_voltage"] = v.astype(f32)
        self.is_spiking = spike
        self.lft = np.where(spike, self.clock, self.lft)
        self.v_hist.append(F["current_voltage"].copy())
        self.s_hist.append(spike.copy())
        if self.do_plasticity:
            self._stdp(spike)
        self.clock += 1

    def _stdp_dw(self, tp, tq):
        s = self.stdp
        if tp < 0 or tq < 0:
            return f32(0)
        tp, tq = f32(tp), f32(tq)
        if tp < tq:
            return s["a_plus"] * np.exp((f32(-1) * np.abs((tp - tq) * s["dt"])) / s["tau_plus"])
        if tp > tq:
            return (f32(-1) * s["a_minus"]) * np.exp((f32(-1) * np.abs((tq - tp) * s["dt"])) / s["tau_minus"])
        return f32(0)

    def _bcm_w(self, w, pre, post):
        b, F = self.bcm, self.f
        sliding_threshold = F["average_activity"][post] / b["average_scalar"]
        activity_term = F["current_activity"][post] * (F["current_activity"][post] - sliding_threshold)
        return f32(w + (activity_term * F["current_activity"][pre] - b["decay"] * w) * b["dt"])

    def _stdp(self, spike):
        # deferred (LatticeNetwork::iterate, mod.rs:2573-2576): in-edges then out-edges of every spiking neuron
        if self.use_bcm:
            for p in np.nonzero(spike)[0]:
                for i in np.nonzero(self.conn[:, p])[0]:
                    self.w[i, p] = self._bcm_w(self.w[i, p], i, p)
                for j in np.nonzero(self.conn[p, :])[0]:
                    self.w[p, j] = self._bcm_w(self.w[p, j], p, j)
            return
        for p in np.nonzero(spike)[0]:
            for i in np.nonzero(self.conn[:, p])[0]:
                self.w[i, p] = self.w[i, p] + self._stdp_dw(self.lft[i], self.lft[p])
            for j in np.nonzero(self.conn[p, :])[0]:
                self.w[p, j] = self.w[p, j] + self._stdp_dw(self.lft[p], self.lft[j])

    def run(self, iterations):
        if not (self.electrical or self.chemical):
            return
        for _ in range(iterations):
            self.step()


class RewardDenseLattice(DenseLattice):
    """RewardModulatedLattice<TraceRSTDP, ..., RewardModulatedSTDP> (mod.rs:2717-3416, plasticity/mod.rs:114-234), written
    independently of the C oracle: dense per-edge TraceRSTDP members, the modulator applied literally inside the node loop in
    ascending node order (in-edges, then out-edges of the node that was just stepped)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        n = self.n
        self.counter = np.zeros((n, n), np.int64)
        self.dw = np.zeros((n, n), f32)
        self.c = np.zeros((n, n), f32)
        self.do_modulation = True
        self.mod = dict(dopamine=f32(0), tau_d=f32(20), tau_c=f32(0.0001), a_plus=f32(2), a_minus=f32(2), tau_plus=f32(4.5),
                        tau_minus=f32(4.5), dt=f32(0.1))

    def _update_weight(self, a, b, t_pre, t_post):
        m = self.mod
        self.stdp = m  # same STDP curve, the modulator's own constants
        self.dw[a, b] = self.dw[a, b] + self._stdp_dw(t_pre, t_post)
        if self.counter[a, b] == 0:
            self.counter[a, b] = 1
        else:
            self.c[a, b] = self.c[a, b] * np.exp(-m["dt"] / m["tau_c"]) + m["tau_c"] * self.dw[a, b]
            self.counter[a, b] = 0
            self.dw[a, b] = f32(0)
        self.w[a, b] = self.w[a, b] + self.c[a, b] * m["dopamine"]

    def step(self, reward=None):
        if reward is not None:  # RewardModulatedSTDP::update, after the inputs and before iterate (mod.rs:3160-3172)
            m = self.mod
            m["dopamine"] = f32(m["dopamine"] * np.exp(-m["dt"] / m["tau_d"]) + m["tau_d"] * f32(reward))
        old = self.lft.copy()
        self.do_plasticity = False
        super().step()   # neuron updates do not read weights, so stepping all of them first changes nothing
        if not self.do_modulation:
            return
        new = self.lft
        cur = old.copy()
        for p in range(self.n):
            cur[p] = new[p]
            for i in np.nonzero(self.conn[:, p])[0]:
                self._update_weight(i, p, cur[i], cur[p])
            for j in np.nonzero(self.conn[p, :])[0]:
                self._update_weight(p, j, cur[p], cur[j])

    def run_with_rewards(self, rewards):
        for r in rewards:
            self.step(r)
