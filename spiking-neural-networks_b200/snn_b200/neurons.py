"""Host-side neuron / synapse parameter objects mirroring the reference's Rust structs.

Each class carries the same public fields and defaults as its Rust counterpart; a lattice flattens
them into the named SoA fields of the C ABI exactly as `IterateAndSpikeGPU::convert_to_gpu` does
(reference: iterate_and_spike/mod.rs:3156-3189, integrate_and_fire/mod.rs:729-773).
"""
from __future__ import annotations

import copy
from enum import IntEnum

from . import _capi as K


class IonotropicNeurotransmitterType(IntEnum):
    """iterate_and_spike/mod.rs:1068-1073; numeric order :1322-1330."""
    AMPA = 0
    NMDA = 1
    GABA = 2


class _Params:
    _defaults: dict = {}

    def __init__(self, **kw):
        for k, v in self._defaults.items():
            setattr(self, k, v)
        for k, v in kw.items():
            if k not in self._defaults:
                raise AttributeError(f"{type(self).__name__} has no field {k!r}")
            setattr(self, k, v)

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={getattr(self, k)!r}' for k in self._defaults)})"

    def __eq__(self, other):
        return type(self) is type(other) and all(getattr(self, k) == getattr(other, k) for k in self._defaults)


# ---- neurotransmitter kinetics (iterate_and_spike/mod.rs:122-366) -------------------------------
class ApproximateNeurotransmitter(_Params):
    kind = K.NT_APPROXIMATE
    _defaults = dict(t_max=1.0, t=0.0, clearance_constant=0.01)


class DestexheNeurotransmitter(_Params):
    kind = K.NT_DESTEXHE
    _defaults = dict(t_max=1.0, t=0.0, v_p=2.0, k_p=5.0)


class DiscreteSpikeNeurotransmitter(_Params):
    kind = K.NT_DISCRETE_SPIKE
    _defaults = dict(t_max=1.0, t=0.0)


class ExponentialDecayNeurotransmitter(_Params):
    kind = K.NT_EXPONENTIAL_DECAY
    _defaults = dict(t_max=1.0, t=0.0, decay_constant=2.0)


NT_KINETICS = {c.kind: c for c in (ApproximateNeurotransmitter, DestexheNeurotransmitter, DiscreteSpikeNeurotransmitter,
                                   ExponentialDecayNeurotransmitter)}


# ---- receptor kinetics (iterate_and_spike/mod.rs:391-533) ---------------------------------------
class ApproximateReceptor(_Params):
    kind = K.RC_APPROXIMATE
    _defaults = dict(r=0.0)


class DestexheReceptor(_Params):
    kind = K.RC_DESTEXHE
    _defaults = dict(r=0.0, alpha=1.0, beta=1.0)


class ExponentialDecayReceptor(_Params):
    kind = K.RC_EXPONENTIAL_DECAY
    _defaults = dict(r_max=1.0, r=0.0, decay_constant=2.0)


RC_KINETICS = {c.kind: c for c in (ApproximateReceptor, DestexheReceptor, ExponentialDecayReceptor)}


# ---- ionotropic receptors (iterate_and_spike/mod.rs:1077-1167) ----------------------------------
class _Receptor(_Params):
    def __init__(self, r=None, **kw):
        super().__init__(**kw)
        self.r = r if r is not None else ApproximateReceptor()

    def __eq__(self, other):
        return super().__eq__(other) and self.r == other.r


class AMPAReceptor(_Receptor):
    type = IonotropicNeurotransmitterType.AMPA
    _defaults = dict(current=0.0, g=1.0, e=0.0)


class NMDAReceptor(_Receptor):
    type = IonotropicNeurotransmitterType.NMDA
    _defaults = dict(current=0.0, g=0.6, mg=0.3, e=0.0)


class GABAReceptor(_Receptor):
    type = IonotropicNeurotransmitterType.GABA
    _defaults = dict(current=0.0, g=1.2, e=-80.0)


RECEPTORS = {c.type: c for c in (AMPAReceptor, NMDAReceptor, GABAReceptor)}

# ---- neuron models ----------------------------------------------------------------------------
_COMMON_TAIL = dict(is_spiking=False, last_firing_time=None)

_MODEL_DEFAULTS = {
    # integrate_and_fire/mod.rs:149-172
    K.MODEL_LIF: dict(current_voltage=-75.0, refractory_count=0.0, leak_constant=-1.0, integration_constant=1.0,
                      gap_conductance=7.0, v_th=-55.0, v_reset=-75.0, tau_m=10.0, c_m=100.0, g_l=10.0, v_init=-75.0,
                      e_l=-75.0, tref=10.0, dt=0.1),
    # :298-320
    K.MODEL_QIF: dict(current_voltage=-75.0, refractory_count=0.0, integration_constant=1.0, gap_conductance=7.0, alpha=1.0,
                      v_th=-55.0, v_reset=-75.0, v_c=-60.0, tau_m=100.0, c_m=100.0, v_init=-75.0, tref=10.0, dt=0.1),
    # :970-997
    K.MODEL_ADLIF: dict(current_voltage=-75.0, refractory_count=0.0, leak_constant=-1.0, integration_constant=1.0,
                        gap_conductance=7.0, w_value=0.0, alpha=6.0, beta=10.0, v_th=-55.0, v_reset=-75.0, tau_m=10.0,
                        c_m=100.0, g_l=10.0, v_init=-75.0, e_l=-75.0, tref=10.0, w_init=0.0, dt=0.1),
    # :1106-1134
    K.MODEL_ADEX: dict(current_voltage=-75.0, refractory_count=0.0, leak_constant=-1.0, integration_constant=1.0,
                       gap_conductance=7.0, w_value=0.0, alpha=6.0, beta=10.0, slope_factor=1.0, v_th=-55.0, v_reset=-75.0,
                       tau_m=10.0, c_m=100.0, g_l=10.0, v_init=-75.0, e_l=-75.0, tref=10.0, w_init=0.0, dt=0.1),
    # :1198-1220
    K.MODEL_IZH: dict(current_voltage=-65.0, gap_conductance=7.0, w_value=30.0, a=0.02, b=0.2, c=-55.0, d=8.0, v_th=30.0,
                      tau_m=1.0, c_m=100.0, v_init=-65.0, w_init=30.0, dt=0.1),
    # :1313-1336
    K.MODEL_LEAKY_IZH: dict(current_voltage=-65.0, gap_conductance=7.0, w_value=30.0, a=0.02, b=0.2, c=-55.0, d=8.0,
                            v_th=30.0, tau_m=10.0, c_m=100.0, v_init=-65.0, e_l=-65.0, w_init=30.0, dt=0.1),
    # :1552-1570
    K.MODEL_SIMPLE_LIF: dict(current_voltage=-75.0, gap_conductance=10.0, v_th=-55.0, v_reset=-75.0, c_m=100.0, g=-0.1,
                             v_init=-75.0, e=0.0, dt=0.1),
    # :1411-1438 (BCMIzhikevichNeuron: Izhikevich + BCMActivity bookkeeping)
    K.MODEL_BCM_IZH: dict(current_voltage=-65.0, gap_conductance=7.0, w_value=30.0, a=0.02, b=0.2, c=-55.0, d=8.0, v_th=30.0,
                          tau_m=1.0, c_m=100.0, v_init=-65.0, w_init=30.0, dt=0.1, average_activity=0.0, current_activity=0.0,
                          period=3, num_spikes=0, firing_rate_clock=0.0, firing_rate_window=500.0),
    # hodgkin_huxley/mod.rs:80-99; ion_channels/mod.rs:23-31, 205-215, 255-264, 299-307
    K.MODEL_HH: {"current_voltage": -65.0, "gap_conductance": 7.0, "dt": 0.01, "c_m": 1.0, "v_th": 0.0,
                 "na_channel$g_na": 120.0, "na_channel$e_na": 50.0, "na_channel$current": 0.0,
                 "na_channel$m$alpha": 0.0, "na_channel$m$beta": 0.0, "na_channel$m$state": 0.0,
                 "na_channel$h$alpha": 0.0, "na_channel$h$beta": 0.0, "na_channel$h$state": 0.0,
                 "k_channel$g_k": 36.0, "k_channel$e_k": -77.0, "k_channel$current": 0.0,
                 "k_channel$n$alpha": 0.0, "k_channel$n$beta": 0.0, "k_channel$n$state": 0.0,
                 "k_leak_channel$g_k_leak": 0.3, "k_leak_channel$e_k_leak": -55.0, "k_leak_channel$current": 0.0,
                 "was_increasing": False},
}


class Neuron:
    """Base of the IterateAndSpike mirrors.  Scalar fields are plain attributes; HH's nested channel
    members are reachable both as `n["na_channel$m$state"]` and `n.na_channel_m_state`."""
    model = None
    default_nt = ApproximateNeurotransmitter
    default_rc = ApproximateReceptor

    def __init__(self, **kw):
        object.__setattr__(self, "_v", dict(_MODEL_DEFAULTS[self.model]))
        self._v.update(_COMMON_TAIL)
        object.__setattr__(self, "synaptic_neurotransmitters", {})  # Neurotransmitters::default is empty (:2169-2175)
        object.__setattr__(self, "receptors", {})                   # Ionotropic::default is empty (:1307-1313)
        for k, v in kw.items():
            setattr(self, k, v)

    @staticmethod
    def _key(name):
        return name

    def __getattr__(self, name):
        v = object.__getattribute__(self, "_v")
        if name in v:
            return v[name]
        alt = name.replace("_channel_", "_channel$").replace("$m_", "$m$").replace("$h_", "$h$").replace("$n_", "$n$")
        if alt in v:
            return v[alt]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in ("synaptic_neurotransmitters", "receptors"):
            object.__setattr__(self, name, value)
            return
        v = self._v
        if name in v:
            v[name] = value
            return
        alt = name.replace("_channel_", "_channel$").replace("$m_", "$m$").replace("$h_", "$h$").replace("$n_", "$n$")
        if alt in v:
            v[alt] = value
            return
        raise AttributeError(f"{type(self).__name__} has no field {name!r}")

    def __getitem__(self, k):
        return self._v[k]

    def __setitem__(self, k, val):
        if k not in self._v:
            raise KeyError(k)
        self._v[k] = val

    def scalar_fields(self):
        return self._v

    def clone(self):
        return copy.deepcopy(self)

    @classmethod
    def default_impl(cls):
        return cls()

    def __repr__(self):
        return f"{type(self).__name__}({self._v}, nt={self.synaptic_neurotransmitters}, receptors={self.receptors})"


class LeakyIntegrateAndFireNeuron(Neuron):
    model = K.MODEL_LIF


class QuadraticIntegrateAndFireNeuron(Neuron):
    model = K.MODEL_QIF


class AdaptiveLeakyIntegrateAndFireNeuron(Neuron):
    model = K.MODEL_ADLIF


class AdaptiveExpLeakyIntegrateAndFireNeuron(Neuron):
    model = K.MODEL_ADEX


class IzhikevichNeuron(Neuron):
    model = K.MODEL_IZH


class LeakyIzhikevichNeuron(Neuron):
    model = K.MODEL_LEAKY_IZH


class BCMIzhikevichNeuron(Neuron):
    """integrate_and_fire/mod.rs:1358-1520; get_activity / get_averaged_activity mirror BCMActivity (:1510-1519)."""
    model = K.MODEL_BCM_IZH

    def get_activity(self):
        return self.current_activity

    def get_averaged_activity(self):
        return self.average_activity


class SimpleLeakyIntegrateAndFire(Neuron):
    model = K.MODEL_SIMPLE_LIF


class HodgkinHuxleyNeuron(Neuron):
    """default_impl is HodgkinHuxleyNeuron<DestexheNeurotransmitter, DestexheReceptor> (hodgkin_huxley/mod.rs:101-106)."""
    model = K.MODEL_HH
    default_nt = DestexheNeurotransmitter
    default_rc = DestexheReceptor


NEURON_CLASSES = {c.model: c for c in (LeakyIntegrateAndFireNeuron, QuadraticIntegrateAndFireNeuron,
                                       AdaptiveLeakyIntegrateAndFireNeuron, AdaptiveExpLeakyIntegrateAndFireNeuron,
                                       IzhikevichNeuron, LeakyIzhikevichNeuron, SimpleLeakyIntegrateAndFire,
                                       HodgkinHuxleyNeuron, BCMIzhikevichNeuron)}


# ---- spike trains (spike_train/mod.rs) ----------------------------------------------------------
class DeltaDiracRefractoriness(_Params):
    kind = K.REFRACT_DELTA_DIRAC
    _defaults = dict(k=10000.0)


class ExponentialDecayRefractoriness(_Params):
    kind = K.REFRACT_EXPONENTIAL_DECAY
    _defaults = dict(k=10000.0)


class SpikeTrain:
    kind = None
    _extra: dict = {}

    def __init__(self, **kw):
        object.__setattr__(self, "_v", dict(current_voltage=0.0, v_th=30.0, v_resting=0.0, dt=0.1, is_spiking=False,
                                            last_firing_time=None, **self._extra))
        object.__setattr__(self, "synaptic_neurotransmitters", {})
        object.__setattr__(self, "neural_refractoriness", DeltaDiracRefractoriness())
        object.__setattr__(self, "firing_times", [])
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        v = object.__getattribute__(self, "_v")
        if name in v:
            return v[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in ("synaptic_neurotransmitters", "neural_refractoriness", "firing_times"):
            object.__setattr__(self, name, value)
        elif name in self._v:
            self._v[name] = value
        else:
            raise AttributeError(f"{type(self).__name__} has no field {name!r}")

    def scalar_fields(self):
        return self._v

    def clone(self):
        return copy.deepcopy(self)

    @classmethod
    def default_impl(cls):
        return cls()


class PoissonNeuron(SpikeTrain):
    kind = K.TRAIN_POISSON
    _extra = dict(chance_of_firing=0.0)

    @classmethod
    def from_firing_rate(cls, hertz: float, dt: float):
        """spike_train/mod.rs:330-337: chance = 1 / ((1000 / dt) / hertz), evaluated in f32."""
        import numpy as np
        f = np.float32
        p = cls()
        p.dt = float(f(dt))
        with np.errstate(divide="ignore"):
            p.chance_of_firing = float(f(1.0) / ((f(1000.0) / f(dt)) / f(hertz)))
        return p

    default_impl_from_firing_rate = from_firing_rate


class RateSpikeTrain(SpikeTrain):
    kind = K.TRAIN_RATE
    _extra = dict(rate=0.0, step=0.0)


class PresetSpikeTrain(SpikeTrain):
    kind = K.TRAIN_PRESET
    _extra = dict(internal_clock=0.0, counter=0)


class STDP(_Params):
    """plasticity/mod.rs:14-39."""
    _defaults = dict(a_plus=2.0, a_minus=2.0, tau_plus=4.5, tau_minus=4.5, dt=0.1)


class BCM(_Params):
    """plasticity/mod.rs:80-112: Bienenstock-Cooper-Munro rule over the BCMActivity of BCMIzhikevichNeuron lattices."""
    _defaults = dict(decay=0.1, average_scalar=0.1, dt=0.1)


class RewardModulatedSTDP(_Params):
    """plasticity/mod.rs:155-189; `update(reward)` (:193-195) runs on the back end in run_lattice_with_reward."""
    _defaults = dict(dopamine=0.0, tau_d=20.0, tau_c=0.0001, a_plus=2.0, a_minus=2.0, tau_plus=4.5, tau_minus=4.5, dt=0.1)
