"""ctypes wrapper of the CPU oracle (oracle/libsnn_oracle.so) — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  It exposes the same back-end protocol as snn_b200.backend so one scenario description can
be stepped by the CUDA library and by the oracle and then compared.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsnn_oracle.so")

F32, U32, I32 = 0, 1, 2
_NP = {F32: np.float32, U32: np.uint32, I32: np.int32}


class Rstdp(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("dopamine", "tau_d", "tau_c", "a_plus", "a_minus", "tau_plus", "tau_minus", "dt")]


class Bcm(C.Structure):
    _fields_ = [("decay", C.c_float), ("average_scalar", C.c_float), ("dt", C.c_float)]


class Stdp(C.Structure):
    _fields_ = [("a_plus", C.c_float), ("a_minus", C.c_float), ("tau_plus", C.c_float), ("tau_minus", C.c_float),
                ("dt", C.c_float)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "snn_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    P, u64, u32, i32, f = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_float
    sig = {
        "orc_network_create": ([i32] * 5, P),
        "orc_network_destroy": ([P], None),
        "orc_add_lattice": ([P, u64, u32, u32], i32),
        "orc_add_train_lattice": ([P, u64, u32, u32], i32),
        "orc_lattice_size": ([P, u64], u64),
        "orc_set_field": ([P, u64, C.c_char_p, P, u64, i32], i32),
        "orc_get_field": ([P, u64, C.c_char_p, P, u64, i32], i32),
        "orc_fill_field_f32": ([P, u64, C.c_char_p, f], i32),
        "orc_fill_field_u32": ([P, u64, C.c_char_p, u32], i32),
        "orc_fill_field_i32": ([P, u64, C.c_char_p, i32], i32),
        "orc_set_preset_firing_times": ([P, u64, P, P, u64, u64], i32),
        "orc_connect_dense": ([P, u64, u64, P, P, u64, u64], i32),
        "orc_connect_csr": ([P, u64, u64, P, P, P, u64, u64], i32),
        "orc_connect_grid": ([P, u64, u32, f], i32),
        "orc_get_connection_dense": ([P, u64, u64, P, P, u64, u64], i32),
        "orc_connection_nnz": ([P, u64, u64], u64),
        "orc_get_connection_csr": ([P, u64, u64, P, P, P], i32),
        "orc_set_synapses": ([P, i32, i32], None),
        "orc_set_parallel": ([P, i32], None),
        "orc_set_clock": ([P, u64], None),
        "orc_get_clock": ([P], u64),
        "orc_set_lattice_flags": ([P, u64, i32, i32, i32], i32),
        "orc_set_plasticity": ([P, u64, C.POINTER(Stdp)], i32),
        "orc_set_dt": ([P, f], None),
        "orc_reset_timing": ([P], None),
        "orc_seed": ([P, u64], None),
        "orc_run": ([P, u64], i32),
        "orc_set_bcm_plasticity": ([P, u64, i32, C.POINTER(Bcm)], i32),
        "orc_set_reward_modulator": ([P, i32, i32, C.POINTER(Rstdp)], i32),
        "orc_get_dopamine": ([P], f),
        "orc_run_with_reward": ([P, f], i32),
        "orc_add_reward_lattice": ([P, u64, u32, u32], i32),
        "orc_set_lattice_reward_modulator": ([P, u64, i32, C.POINTER(Rstdp)], i32),
        "orc_get_lattice_reward_modulator": ([P, u64, C.POINTER(i32), C.POINTER(Rstdp)], i32),
        "orc_mark_connection_reward": ([P, u64, u64, i32], i32),
        "orc_run_network_with_reward": ([P, f], i32),
        "orc_get_connection_traces": ([P, u64, u64, P, P, P], i32),
        "orc_set_connection_traces": ([P, u64, u64, P, P, P, P], i32),
        "orc_history_len": ([P, u64], u64),
        "orc_set_reduced_history": ([P, u64, i32, i32, f, f, f], i32),
        "orc_get_reduced_history": ([P, u64, i32, P, u64], i32),
        "orc_get_grid_history": ([P, u64, P, u64], i32),
        "orc_get_spike_history": ([P, u64, P, u64], i32),
        "orc_reset_history": ([P], None),
        "orc_chemical_inputs_dense": ([P, P, P, P, u32, u32, P, P], None),
        "orc_stdp_update": ([C.POINTER(Stdp), f, i32, i32], f),
        "orc_refractoriness_effect": ([i32, f, u64, u64, f, f, f], f),
        "orc_chance_from_firing_rate": ([f, f], f),
        "orc_adjmat_create": ([], P),
        "orc_adjmat_destroy": ([P], None),
        "orc_adjmat_add_node": ([P, u32, u32], None),
        "orc_adjmat_edit_weight": ([P, u32, u32, u32, u32, i32, f], i32),
        "orc_adjmat_lookup_weight": ([P, u32, u32, u32, u32, C.POINTER(i32), C.POINTER(f)], i32),
        "orc_adjmat_incoming": ([P, u32, u32, P, i32], i32),
        "orc_adjmat_outgoing": ([P, u32, u32, P, i32], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


class OracleError(RuntimeError):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _as(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# field dtypes by name (everything else is f32)
_U32_FIELDS = {"is_spiking", "was_increasing", "counter", "neurotransmitters$flags", "receptors$flags", "period", "num_spikes"}
_I32_FIELDS = {"last_firing_time"}
_PER3_PREFIX = "neurotransmitters$"


def _meta(name):
    dt = U32 if name in _U32_FIELDS else (I32 if name in _I32_FIELDS else F32)
    per = 3 if (name.startswith(_PER3_PREFIX) or name == "receptors$flags") else 1
    return dt, per


class OracleBackend:
    """Same protocol as snn_b200.backend.CudaNetworkBackend / CudaLatticeBackend."""

    def __init__(self, model, ntk=0, rck=0, train_kind=0, refract=0, rows=None, cols=None):
        self.L = lib()
        self.h = self.L.orc_network_create(model, ntk, rck, train_kind, refract)
        self._flags = {}
        self._red = {}
        if rows is not None:
            self.add_lattice(0, rows, cols)

    def close(self):
        if self.h:
            self.L.orc_network_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r:
            raise OracleError(r)

    def add_lattice(self, id, rows, cols):
        self._ck(self.L.orc_add_lattice(self.h, id, rows, cols))
        self._flags[id] = [0, 0, 0]

    def add_reward_lattice(self, id, rows, cols):
        self._ck(self.L.orc_add_reward_lattice(self.h, id, rows, cols))
        self._flags[id] = [0, 0, 0]

    def set_lattice_reward_modulator(self, id, do_modulation, **m):
        s = Rstdp(*[float(m[k]) for k in self._RSTDP])
        self._ck(self.L.orc_set_lattice_reward_modulator(self.h, id, int(do_modulation), C.byref(s)))

    def get_lattice_reward_modulator(self, id):
        s, dm = Rstdp(), C.c_int32(0)
        self._ck(self.L.orc_get_lattice_reward_modulator(self.h, id, C.byref(dm), C.byref(s)))
        return bool(dm.value), {k: getattr(s, k) for k in self._RSTDP}

    def mark_connection_reward(self, pre_id, post_id, reward_modulated):
        self._ck(self.L.orc_mark_connection_reward(self.h, pre_id, post_id, int(reward_modulated)))

    def run_network_with_rewards(self, rewards):
        for r in np.asarray(rewards, np.float32).reshape(-1):
            self._ck(self.L.orc_run_network_with_reward(self.h, float(r)))

    def add_train_lattice(self, id, rows, cols):
        self._ck(self.L.orc_add_train_lattice(self.h, id, rows, cols))
        self._flags[id] = [0, 0, 0]

    def size(self, id=0):
        return self.L.orc_lattice_size(self.h, id)

    def set_field(self, id, name, arr):
        dt, _ = _meta(name)
        a = _as(np.asarray(arr).reshape(-1), _NP[dt])
        self._ck(self.L.orc_set_field(self.h, id, name.encode(), _ptr(a), a.size, dt))

    def get_field(self, id, name):
        dt, per = _meta(name)
        out = np.empty(self.size(id) * per, dtype=_NP[dt])
        self._ck(self.L.orc_get_field(self.h, id, name.encode(), _ptr(out), out.size, dt))
        return out

    def fill_field(self, id, name, value):
        dt, _ = _meta(name)
        fn = {F32: self.L.orc_fill_field_f32, U32: self.L.orc_fill_field_u32, I32: self.L.orc_fill_field_i32}[dt]
        self._ck(fn(self.h, id, name.encode(), float(value) if dt == F32 else int(value)))

    def set_preset_firing_times(self, id, offsets, times):
        off, t = _as(offsets, np.uint64), _as(times, np.float32)
        self._ck(self.L.orc_set_preset_firing_times(self.h, id, _ptr(off), _ptr(t), off.size - 1, t.size))

    def connect_dense(self, pre_id, post_id, connections, weights, index_to_position=None):
        c, w = _as(connections, np.uint32), _as(weights, np.float32)
        if index_to_position is not None:
            itp = np.asarray(index_to_position)
            n = itp.size
            c2, w2 = np.zeros((n, n), np.uint32), np.zeros((n, n), np.float32)
            c2[np.ix_(itp, itp)] = c.reshape(n, n)
            w2[np.ix_(itp, itp)] = w.reshape(n, n)
            c, w = _as(c2, np.uint32), _as(w2, np.float32)
        self._ck(self.L.orc_connect_dense(self.h, pre_id, post_id, _ptr(c), _ptr(w), self.size(pre_id), self.size(post_id)))

    def connect_csr(self, pre_id, post_id, row_ptr, pre, weights):
        rp, pr, w = _as(row_ptr, np.uint64), _as(pre, np.uint32), _as(weights, np.float32)
        self._ck(self.L.orc_connect_csr(self.h, pre_id, post_id, _ptr(rp), _ptr(pr), _ptr(w), rp.size - 1, pr.size))

    def connect_grid(self, id, radius, weight):
        self._ck(self.L.orc_connect_grid(self.h, id, int(radius), float(weight)))

    def connection_nnz(self, pre_id=0, post_id=0):
        return self.L.orc_connection_nnz(self.h, pre_id, post_id)

    def get_connection_csr(self, pre_id=0, post_id=0):
        nnz = self.connection_nnz(pre_id, post_id)
        rp = np.zeros(self.size(post_id) + 1, np.uint64)
        pr = np.zeros(max(nnz, 1), np.uint32)
        w = np.zeros(max(nnz, 1), np.float32)
        self._ck(self.L.orc_get_connection_csr(self.h, pre_id, post_id, _ptr(rp), _ptr(pr), _ptr(w)))
        return rp, pr[:nnz], w[:nnz]

    def get_connection_dense(self, pre_id=0, post_id=0):
        n_pre, n_post = self.size(pre_id), self.size(post_id)
        c = np.zeros(n_pre * n_post, np.uint32)
        w = np.zeros(n_pre * n_post, np.float32)
        self._ck(self.L.orc_get_connection_dense(self.h, pre_id, post_id, _ptr(c), _ptr(w), n_pre, n_post))
        return c.reshape(n_pre, n_post), w.reshape(n_pre, n_post)

    # Graph::lookup_weight / edit_weight (graph/mod.rs:196-226) on the oracle's adjacency; called as (pre, post[, weight]) in
    # the lattice role and as (pre_id, post_id, pre, post[, weight]) in the network role, like the two CUDA back ends
    def lookup_weight(self, *a):
        pre_id, post_id, pre, post = (0, 0) + tuple(a) if len(a) == 2 else a
        rp, pr, w = self.get_connection_csr(pre_id, post_id)
        s, t = int(rp[post]), int(rp[post + 1])
        hit = np.nonzero(pr[s:t] == pre)[0]
        return float(w[s + hit[0]]) if hit.size else None

    def edit_weight(self, *a):
        pre_id, post_id, pre, post, weight = (0, 0) + tuple(a) if len(a) == 3 else a
        rp, pr, w = self.get_connection_csr(pre_id, post_id)
        s, t = int(rp[post]), int(rp[post + 1])
        hit = np.nonzero(pr[s:t] == pre)[0]
        if hit.size and weight is not None:
            w = w.copy(); w[s + hit[0]] = weight
        elif hit.size:
            pr, w = np.delete(pr, s + hit[0]), np.delete(w, s + hit[0])
            rp = rp.copy(); rp[post + 1:] -= 1
        elif weight is not None:
            at = s + int(np.searchsorted(pr[s:t], pre))
            pr, w = np.insert(pr, at, pre), np.insert(w, at, np.float32(weight))
            rp = rp.copy(); rp[post + 1:] += 1
        else:
            return
        self.connect_csr(pre_id, post_id, rp, pr, w)

    def spike_aggregate(self, id=0):
        """SpikeHistory::aggregate, neuron/mod.rs:335-359: literal sum of the boolean raster over its steps."""
        h = self.spike_history(id)
        return h.astype(np.int64).sum(axis=0) if h.shape[0] else np.zeros(self.size(id), np.int64)

    # options use the product's option numbering (include/snn_b200.h snn_option_t)
    def set_option(self, option, value, id=None):
        value = int(value)
        if option == 0:
            self._el = value
            self.L.orc_set_synapses(self.h, value, getattr(self, "_ch", 0))
        elif option == 1:
            self._ch = value
            self.L.orc_set_synapses(self.h, getattr(self, "_el", 1), value)
        elif option in (2, 3, 4):
            lid = 0 if id is None else id
            self._flags[lid][option - 2] = value
            fl = self._flags[lid]
            self._ck(self.L.orc_set_lattice_flags(self.h, lid, fl[0], fl[1], fl[2]))
        elif option == 5:
            self.L.orc_set_clock(self.h, value)
        elif option == 6:
            self.L.orc_set_parallel(self.h, value)
        elif option == 7:
            self.L.orc_seed(self.h, value)
        elif option in (8, 10):
            lid = 0 if id is None else id
            r = self._red.setdefault(lid, [0, 0, 0.007, 0.8, 251.0])
            r[0 if option == 8 else 1] = value
            self._ck(self.L.orc_set_reduced_history(self.h, lid, *r))
        elif option == 9:
            pass
        else:
            raise OracleError(64)

    def get_option(self, option, id=None):
        if option == 5:
            return self.L.orc_get_clock(self.h)
        raise OracleError(64)

    def set_plasticity(self, id, a_plus, a_minus, tau_plus, tau_minus, dt):
        s = Stdp(a_plus, a_minus, tau_plus, tau_minus, dt)
        self._ck(self.L.orc_set_plasticity(self.h, id, C.byref(s)))

    _RSTDP = ("dopamine", "tau_d", "tau_c", "a_plus", "a_minus", "tau_plus", "tau_minus", "dt")

    def set_reward_modulator(self, enable, do_modulation, **m):
        # the dopamine level lives on the back end between calls (the front end pushes its last read-back value)
        s = Rstdp(*[float(m[k]) for k in self._RSTDP])
        self._ck(self.L.orc_set_reward_modulator(self.h, int(enable), int(do_modulation), C.byref(s)))
        self._rstdp = dict(m)

    def get_reward_modulator(self):
        d = dict(self._rstdp)
        d["dopamine"] = self.L.orc_get_dopamine(self.h)
        return d

    def run_with_rewards(self, rewards):
        for r in np.asarray(rewards, np.float32).reshape(-1):
            self._ck(self.L.orc_run_with_reward(self.h, float(r)))

    def connection_traces(self, pre_id=0, post_id=0):
        nnz = self.connection_nnz(pre_id, post_id)
        cnt, dw, c = np.zeros(max(nnz, 1), np.uint32), np.zeros(max(nnz, 1), np.float32), np.zeros(max(nnz, 1), np.float32)
        self._ck(self.L.orc_get_connection_traces(self.h, pre_id, post_id, _ptr(cnt), _ptr(dw), _ptr(c)))
        return cnt[:nnz], dw[:nnz], c[:nnz]

    def set_bcm_plasticity(self, id, enable, decay, average_scalar, dt):
        s = Bcm(decay, average_scalar, dt)
        self._ck(self.L.orc_set_bcm_plasticity(self.h, 0 if id is None else id, int(enable), C.byref(s)))

    def set_connection_traces(self, weight=None, counter=None, dw=None, c=None, pre_id=0, post_id=0):
        arrs = [None if x is None else _as(np.asarray(x).reshape(-1), t)
                for x, t in ((weight, np.float32), (counter, np.uint32), (dw, np.float32), (c, np.float32))]
        self._ck(self.L.orc_set_connection_traces(self.h, pre_id, post_id, *[None if a is None else _ptr(a) for a in arrs]))

    def set_dt(self, dt):
        self.L.orc_set_dt(self.h, float(dt))

    def reset_timing(self):
        self.L.orc_reset_timing(self.h)

    def run(self, iterations):
        self._ck(self.L.orc_run(self.h, int(iterations)))

    def history_len(self, id=0):
        return self.L.orc_history_len(self.h, id)

    def grid_history(self, id=0):
        steps, n = self.history_len(id), self.size(id)
        out = np.zeros(steps * n, np.float32)
        self._ck(self.L.orc_get_grid_history(self.h, id, _ptr(out), out.size))
        return out.reshape(steps, n)

    def spike_history(self, id=0):
        steps, n = self.history_len(id), self.size(id)
        out = np.zeros(steps * n, np.uint8)
        self._ck(self.L.orc_get_spike_history(self.h, id, _ptr(out), out.size))
        return out.reshape(steps, n)

    def set_eeg_parameters(self, id, reference_voltage, distance, conductivity):
        lid = 0 if id is None else id
        r = self._red.setdefault(lid, [0, 0, 0.007, 0.8, 251.0])
        r[2:] = [float(reference_voltage), float(distance), float(conductivity)]
        self._ck(self.L.orc_set_reduced_history(self.h, lid, *r))

    def _reduced(self, id, eeg):
        out = np.zeros(self.history_len(id), np.float32)
        self._ck(self.L.orc_get_reduced_history(self.h, id, eeg, _ptr(out), out.size))
        return out

    def average_history(self, id=0):
        return self._reduced(id, 0)

    def eeg_history(self, id=0):
        return self._reduced(id, 1)

    def reset_history(self):
        self.L.orc_reset_history(self.h)
