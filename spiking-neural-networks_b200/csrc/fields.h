// fields.h — the named structure-of-arrays field directory of the boundary.
//
// Names and dtypes are the reference's buffer-map keys (IterateAndSpikeGPU::convert_to_gpu,
// integrate_and_fire/mod.rs:729-773; Neurotransmitters::convert_to_gpu, iterate_and_spike/mod.rs:2566-2656;
// Ionotropic ReceptorsGPU, iterate_and_spike/mod.rs:1366-1766; spike trains, spike_train/mod.rs:560-625,
// 1212-1240).  Models the reference never gave a GPU map (Izhikevich, HH, ...) use the Rust struct
// field names, nested members joined with '$' as the reference does for its own nested structs.
#pragma once
#include <cstring>

#include "common.h"

namespace snn {

enum FieldKind : int {
    FK_NEURON_DEV = 0,   // f32 per neuron, device array f[slot]
    FK_NEURON_COLD,      // f32 per neuron, never read by the step loop: host-side storage only
    FK_V,                // current_voltage (node array, ping-pong)
    FK_SPIKING,          // is_spiking u32 <-> 1 bit per node
    FK_LFT,              // last_firing_time i32 (node array)
    FK_WAS_INC,          // HH was_increasing u32 <-> 1 bit per neuron
    FK_NT_FLAGS,         // u32 [n*3]
    FK_NT_T,             // f32 [n*3] (ping-pong)
    FK_NT,               // f32 [n*3] nt[slot]
    FK_RC_FLAGS,         // u32 [n*3]
    FK_RC,               // f32 per neuron, rc[slot] of type `aux`
    FK_TRAIN_DEV,        // f32 per train, tf[slot]
    FK_TRAIN_COUNTER     // u32 per train stored as bits in tf[TF_COUNTER]
};

struct FieldDef {
    const char *name;
    int dtype;
    int per;        // elements per cell in the flat host array (1 or 3)
    int kind;
    int slot;
    int aux;        // receptor type for FK_RC; kinetics mask for FK_NT/FK_RC (bit = kinetics enum), 0 = all
    uint32_t models;  // bit per snn_model_t (neurons) or per snn_spike_train_t (trains)
};

#define SNN_M(x) (1u << (x))
constexpr uint32_t kAllModels = 0x1FFu;
constexpr uint32_t kIF4 = SNN_M(SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE) | SNN_M(SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE) |
                          SNN_M(SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE) | SNN_M(SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE);
constexpr uint32_t kLeaky3 = SNN_M(SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE) | SNN_M(SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE) |
                             SNN_M(SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE);
constexpr uint32_t kAdapt2 = SNN_M(SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE) | SNN_M(SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE);
constexpr uint32_t kBcm = SNN_M(SNN_MODEL_BCM_IZHIKEVICH);
constexpr uint32_t kIzh2 = SNN_M(SNN_MODEL_IZHIKEVICH) | SNN_M(SNN_MODEL_LEAKY_IZHIKEVICH) | kBcm;   // every Izhikevich variant
constexpr uint32_t kHH = SNN_M(SNN_MODEL_HODGKIN_HUXLEY);
constexpr uint32_t kSimple = SNN_M(SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE);

static const FieldDef kNeuronFields[] = {
    {"current_voltage", SNN_F32, 1, FK_V, 0, 0, kAllModels},
    {"gap_conductance", SNN_F32, 1, FK_NEURON_DEV, F_GAP, 0, kAllModels},
    {"dt", SNN_F32, 1, FK_NEURON_DEV, F_DT, 0, kAllModels},
    {"c_m", SNN_F32, 1, FK_NEURON_DEV, F_CM, 0, kAllModels},
    {"v_th", SNN_F32, 1, FK_NEURON_DEV, F_VTH, 0, kAllModels},
    {"is_spiking", SNN_U32, 1, FK_SPIKING, 0, 0, kAllModels},
    {"last_firing_time", SNN_I32, 1, FK_LFT, 0, 0, kAllModels},
    {"v_reset", SNN_F32, 1, FK_NEURON_DEV, F_VRESET, 0, kIF4 | kSimple},
    {"v_init", SNN_F32, 1, FK_NEURON_COLD, 0, 0, kAllModels & ~kHH},
    {"refractory_count", SNN_F32, 1, FK_NEURON_DEV, F_REFR, 0, kIF4},
    {"tref", SNN_F32, 1, FK_NEURON_DEV, F_TREF, 0, kIF4},
    {"leak_constant", SNN_F32, 1, FK_NEURON_DEV, F_LEAK, 0, kLeaky3},
    {"integration_constant", SNN_F32, 1, FK_NEURON_DEV, F_INTEG, 0, kIF4},
    {"e_l", SNN_F32, 1, FK_NEURON_DEV, F_EL, 0, kLeaky3 | SNN_M(SNN_MODEL_LEAKY_IZHIKEVICH)},
    {"g_l", SNN_F32, 1, FK_NEURON_DEV, F_GL, 0, kLeaky3},
    {"tau_m", SNN_F32, 1, FK_NEURON_DEV, F_TAUM, 0, kIF4 | kIzh2},
    {"alpha", SNN_F32, 1, FK_NEURON_DEV, F_ALPHA, 0, SNN_M(SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE) | kAdapt2},
    {"beta", SNN_F32, 1, FK_NEURON_DEV, F_BETA, 0, kAdapt2},
    {"slope_factor", SNN_F32, 1, FK_NEURON_DEV, F_SLOPE, 0, SNN_M(SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE)},
    {"w_value", SNN_F32, 1, FK_NEURON_DEV, F_W, 0, kAdapt2 | kIzh2},
    {"w_init", SNN_F32, 1, FK_NEURON_COLD, 1, 0, kAdapt2 | kIzh2},
    {"v_c", SNN_F32, 1, FK_NEURON_DEV, F_VC, 0, SNN_M(SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE)},
    {"a", SNN_F32, 1, FK_NEURON_DEV, F_A, 0, kIzh2},
    {"b", SNN_F32, 1, FK_NEURON_DEV, F_B, 0, kIzh2},
    {"c", SNN_F32, 1, FK_NEURON_DEV, F_C, 0, kIzh2},
    {"d", SNN_F32, 1, FK_NEURON_DEV, F_D, 0, kIzh2},
    // BCMIzhikevichNeuron (integrate_and_fire/mod.rs:1393-1404); usize fields travel as u32 like the reference's bools
    {"average_activity", SNN_F32, 1, FK_NEURON_DEV, F_AVG_ACT, 0, kBcm},
    {"current_activity", SNN_F32, 1, FK_NEURON_DEV, F_CUR_ACT, 0, kBcm},
    {"period", SNN_U32, 1, FK_NEURON_DEV, F_PERIOD, 0, kBcm},
    {"num_spikes", SNN_U32, 1, FK_NEURON_DEV, F_NSPK, 0, kBcm},
    {"firing_rate_clock", SNN_F32, 1, FK_NEURON_DEV, F_FRCLK, 0, kBcm},
    {"firing_rate_window", SNN_F32, 1, FK_NEURON_DEV, F_FRWIN, 0, kBcm},
    {"g", SNN_F32, 1, FK_NEURON_DEV, F_G, 0, kSimple},
    {"e", SNN_F32, 1, FK_NEURON_DEV, F_E, 0, kSimple},
    {"na_channel$g_na", SNN_F32, 1, FK_NEURON_DEV, F_GNA, 0, kHH},
    {"na_channel$e_na", SNN_F32, 1, FK_NEURON_DEV, F_ENA, 0, kHH},
    {"na_channel$current", SNN_F32, 1, FK_NEURON_DEV, F_NA_CUR, 0, kHH},
    {"na_channel$m$alpha", SNN_F32, 1, FK_NEURON_DEV, F_M_ALPHA, 0, kHH},
    {"na_channel$m$beta", SNN_F32, 1, FK_NEURON_DEV, F_M_BETA, 0, kHH},
    {"na_channel$m$state", SNN_F32, 1, FK_NEURON_DEV, F_M, 0, kHH},
    {"na_channel$h$alpha", SNN_F32, 1, FK_NEURON_DEV, F_H_ALPHA, 0, kHH},
    {"na_channel$h$beta", SNN_F32, 1, FK_NEURON_DEV, F_H_BETA, 0, kHH},
    {"na_channel$h$state", SNN_F32, 1, FK_NEURON_DEV, F_H, 0, kHH},
    {"k_channel$g_k", SNN_F32, 1, FK_NEURON_DEV, F_GK, 0, kHH},
    {"k_channel$e_k", SNN_F32, 1, FK_NEURON_DEV, F_EK, 0, kHH},
    {"k_channel$current", SNN_F32, 1, FK_NEURON_DEV, F_K_CUR, 0, kHH},
    {"k_channel$n$alpha", SNN_F32, 1, FK_NEURON_DEV, F_N_ALPHA, 0, kHH},
    {"k_channel$n$beta", SNN_F32, 1, FK_NEURON_DEV, F_N_BETA, 0, kHH},
    {"k_channel$n$state", SNN_F32, 1, FK_NEURON_DEV, F_N, 0, kHH},
    {"k_leak_channel$g_k_leak", SNN_F32, 1, FK_NEURON_DEV, F_GKL, 0, kHH},
    {"k_leak_channel$e_k_leak", SNN_F32, 1, FK_NEURON_DEV, F_EKL, 0, kHH},
    {"k_leak_channel$current", SNN_F32, 1, FK_NEURON_DEV, F_KL_CUR, 0, kHH},
    {"was_increasing", SNN_U32, 1, FK_WAS_INC, 0, 0, kHH},
    // chemical, per type
    {"neurotransmitters$flags", SNN_U32, 3, FK_NT_FLAGS, 0, 0, kAllModels},
    {"neurotransmitters$t", SNN_F32, 3, FK_NT_T, 0, 0, kAllModels},
    {"neurotransmitters$t_max", SNN_F32, 3, FK_NT, NTF_TMAX, 0, kAllModels},
    {"neurotransmitters$clearance_constant", SNN_F32, 3, FK_NT, NTF_P1, SNN_M(SNN_NT_APPROXIMATE), kAllModels},
    {"neurotransmitters$v_p", SNN_F32, 3, FK_NT, NTF_P1, SNN_M(SNN_NT_DESTEXHE), kAllModels},
    {"neurotransmitters$k_p", SNN_F32, 3, FK_NT, NTF_P2, SNN_M(SNN_NT_DESTEXHE), kAllModels},
    {"neurotransmitters$decay_constant", SNN_F32, 3, FK_NT, NTF_P1, SNN_M(SNN_NT_EXPONENTIAL_DECAY), kAllModels},
    {"receptors$flags", SNN_U32, 3, FK_RC_FLAGS, 0, 0, kAllModels},
};
constexpr int kNumNeuronFields = sizeof(kNeuronFields) / sizeof(kNeuronFields[0]);

// receptor fields are generated per type: receptors$<TYPE><suffix>
struct RcFieldDef { const char *suffix; int slot; int kinetics_mask; int nmda_only; };
static const RcFieldDef kRcFields[] = {
    {"_current", RCF_CUR, 0, 0},
    {"_g", RCF_G, 0, 0},
    {"_e", RCF_E, 0, 0},
    {"_mg", RCF_MG, 0, 1},
    {"$r$kinetics$r", RCF_R, 0, 0},
    {"$r$kinetics$alpha", RCF_K1, SNN_M(SNN_RC_DESTEXHE), 0},
    {"$r$kinetics$beta", RCF_K2, SNN_M(SNN_RC_DESTEXHE), 0},
    {"$r$kinetics$r_max", RCF_K1, SNN_M(SNN_RC_EXPONENTIAL_DECAY), 0},
    {"$r$kinetics$decay_constant", RCF_K2, SNN_M(SNN_RC_EXPONENTIAL_DECAY), 0},
};
constexpr int kNumRcFields = sizeof(kRcFields) / sizeof(kRcFields[0]);
static const char *const kRcTypeNames[kNT] = {"AMPA", "NMDA", "GABA"};

static const FieldDef kTrainFields[] = {
    {"current_voltage", SNN_F32, 1, FK_V, 0, 0, 0x7u},
    {"v_th", SNN_F32, 1, FK_TRAIN_DEV, TF_VTH, 0, 0x7u},
    {"v_resting", SNN_F32, 1, FK_TRAIN_DEV, TF_VREST, 0, 0x7u},
    {"dt", SNN_F32, 1, FK_TRAIN_DEV, TF_DT, 0, 0x7u},
    {"is_spiking", SNN_U32, 1, FK_SPIKING, 0, 0, 0x7u},
    {"last_firing_time", SNN_I32, 1, FK_LFT, 0, 0, 0x7u},
    {"neural_refractoriness$k", SNN_F32, 1, FK_TRAIN_DEV, TF_K, 0, 0x7u},
    {"chance_of_firing", SNN_F32, 1, FK_TRAIN_DEV, TF_CHANCE, 0, SNN_M(SNN_TRAIN_POISSON)},
    {"rate", SNN_F32, 1, FK_TRAIN_DEV, TF_RATE, 0, SNN_M(SNN_TRAIN_RATE)},
    {"step", SNN_F32, 1, FK_TRAIN_DEV, TF_STEP, 0, SNN_M(SNN_TRAIN_RATE)},
    {"internal_clock", SNN_F32, 1, FK_TRAIN_DEV, TF_ICLOCK, 0, SNN_M(SNN_TRAIN_PRESET)},
    {"counter", SNN_U32, 1, FK_TRAIN_COUNTER, TF_COUNTER, 0, SNN_M(SNN_TRAIN_PRESET)},
    {"neurotransmitters$flags", SNN_U32, 3, FK_NT_FLAGS, 0, 0, 0x7u},
    {"neurotransmitters$t", SNN_F32, 3, FK_NT_T, 0, 0, 0x7u},
    {"neurotransmitters$t_max", SNN_F32, 3, FK_NT, NTF_TMAX, 0, 0x7u},
    {"neurotransmitters$clearance_constant", SNN_F32, 3, FK_NT, NTF_P1, SNN_M(SNN_NT_APPROXIMATE), 0x7u},
    {"neurotransmitters$v_p", SNN_F32, 3, FK_NT, NTF_P1, SNN_M(SNN_NT_DESTEXHE), 0x7u},
    {"neurotransmitters$k_p", SNN_F32, 3, FK_NT, NTF_P2, SNN_M(SNN_NT_DESTEXHE), 0x7u},
    {"neurotransmitters$decay_constant", SNN_F32, 3, FK_NT, NTF_P1, SNN_M(SNN_NT_EXPONENTIAL_DECAY), 0x7u},
};
constexpr int kNumTrainFields = sizeof(kTrainFields) / sizeof(kTrainFields[0]);

// Default values: <Model>::default() of the reference (integrate_and_fire/mod.rs:149-172, 298-320, 970-997,
// 1106-1134, 1198-1220, 1313-1336, 1552-1570; hodgkin_huxley/mod.rs:80-99; ion_channels/mod.rs:205-215,
// 255-264, 299-307)
inline float neuron_default(int model, int kind, int slot) {
    const bool izh = model == SNN_MODEL_IZHIKEVICH || model == SNN_MODEL_LEAKY_IZHIKEVICH || model == SNN_MODEL_BCM_IZHIKEVICH;
    const bool hh = model == SNN_MODEL_HODGKIN_HUXLEY;
    if (kind == FK_V) return (izh || hh) ? -65.f : -75.f;
    if (kind == FK_NEURON_COLD) {
        if (slot == 0) return izh ? -65.f : -75.f;  // v_init
        return izh ? 30.f : 0.f;                    // w_init
    }
    switch (slot) {
    case F_GAP: return model == SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE ? 10.f : 7.f;
    case F_DT: return hh ? 0.01f : 0.1f;
    case F_CM: return hh ? 1.f : 100.f;
    case F_VTH: return izh ? 30.f : (hh ? 0.f : -55.f);
    case F_VRESET: return -75.f;
    case F_REFR: return 0.f;
    case F_TREF: return 10.f;
    case F_LEAK: return -1.f;
    case F_INTEG: return 1.f;
    case F_EL: return model == SNN_MODEL_LEAKY_IZHIKEVICH ? -65.f : -75.f;
    case F_GL: return 10.f;
    case F_TAUM:
        if (model == SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE) return 100.f;
        if (model == SNN_MODEL_IZHIKEVICH || model == SNN_MODEL_BCM_IZHIKEVICH) return 1.f;
        return 10.f;
    case F_ALPHA: return model == SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE ? 1.f : 6.f;
    case F_BETA: return 10.f;
    case F_SLOPE: return 1.f;
    case F_W: return izh ? 30.f : 0.f;
    case F_VC: return -60.f;
    case F_A: return 0.02f;
    case F_B: return 0.2f;
    case F_C: return -55.f;
    case F_D: return 8.f;
    case F_G: return -0.1f;
    case F_E: return 0.f;
    case F_GNA: return 120.f;
    case F_ENA: return 50.f;
    case F_GK: return 36.f;
    case F_EK: return -77.f;
    case F_GKL: return 0.3f;
    case F_EKL: return -55.f;
    case F_PERIOD: { const uint32_t three = 3u; float bits; memcpy(&bits, &three, 4); return bits; }   // period: usize = 3
    case F_FRWIN: return 500.f;
    default: return 0.f;  // gate states, currents, rates
    }
}

// spike_train/mod.rs:297-311, 778-794, 997-1012; refractoriness k: :50-57
inline float train_default(int slot) {
    switch (slot) {
    case TF_VTH: return 30.f;
    case TF_DT: return 0.1f;
    case TF_K: return 10000.f;
    default: return 0.f;
    }
}

// iterate_and_spike/mod.rs:136-145, 174-182, 315-322, 338-346
inline float nt_default(int ntk, int slot) {
    if (slot == NTF_TMAX) return 1.f;
    if (slot == NTF_P1) return ntk == SNN_NT_APPROXIMATE ? 0.01f : (ntk == SNN_NT_DESTEXHE ? 2.f : 2.0f);
    return 5.f;  // k_p
}

// iterate_and_spike/mod.rs:417-425, 491-495, 525-533, 1085-1094, 1115-1125, 1148-1157
inline float rc_default(int rck, int type, int slot) {
    switch (slot) {
    case RCF_R: return 0.f;
    case RCF_K1: return 1.f;                                  // alpha | r_max
    case RCF_K2: return rck == SNN_RC_DESTEXHE ? 1.f : 2.f;   // beta | decay_constant
    case RCF_G: return type == SNN_NT_AMPA ? 1.f : (type == SNN_NT_NMDA ? 0.6f : 1.2f);
    case RCF_E: return type == SNN_NT_GABA ? -80.f : 0.f;
    case RCF_MG: return 0.3f;
    default: return 0.f;
    }
}

}  // namespace snn
