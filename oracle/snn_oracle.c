/*
 * snn_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see snn_oracle.h for the rules and the
 * parity-pin status).  Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math (oracle/Makefile).
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/backend/src/).  Arithmetic is written operation by operation in the order the
 * Rust source evaluates it; Rust never contracts a*b+c, hence -ffp-contract=off.
 */
#define _GNU_SOURCE
#include "snn_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ cell structs (AoS, like the
 * reference's Vec<Vec<T>> cell grids) */

typedef struct {
    /* Neurotransmitters<N, T> entry: neurotransmitters HashMap<N, T> (iterate_and_spike/mod.rs:2162) */
    uint32_t present;
    float t, t_max;
    float clearance_constant; /* ApproximateNeurotransmitter :165-172 */
    float v_p, k_p;           /* DestexheNeurotransmitter    :125-134 */
    float decay_constant;     /* ExponentialDecayNeurotransmitter :329-336 */
} o_nt;

typedef struct {
    /* Ionotropic<T> entry (iterate_and_spike/mod.rs:1170-1179): AMPA/NMDA/GABA receptor + kinetics */
    uint32_t present;
    float current, g, e, mg;
    float r;                  /* all kinetics */
    float alpha, beta;        /* DestexheReceptor :394-401 */
    float r_max, decay_constant; /* ExponentialDecayReceptor :501-508 */
} o_rc;

typedef struct {
    float alpha, beta, state; /* BasicGatingVariable ion_channels/mod.rs:13-21 */
} o_gate;

typedef struct {
    /* union of the fields of the eight neuron structs */
    float current_voltage, v_th, v_reset, v_init, refractory_count, tref;
    float leak_constant, integration_constant, gap_conductance, e_l, g_l, tau_m, c_m, dt;
    float alpha, beta, slope_factor, w_value, w_init, v_c;
    float a, b, c, d;
    float g, e;
    /* Hodgkin-Huxley */
    float g_na, e_na, na_current;
    o_gate m, h;
    float g_k, e_k, k_current;
    o_gate n;
    float g_k_leak, e_k_leak, k_leak_current;
    /* BCMIzhikevichNeuron integrate_and_fire/mod.rs:1393-1404 */
    float average_activity, current_activity, firing_rate_clock, firing_rate_window;
    uint32_t period, num_spikes;
    uint32_t was_increasing;
    uint32_t is_spiking;
    int32_t last_firing_time; /* Option<usize>, -1 = None */
    o_nt nt[ORC_NT];
    o_rc rc[ORC_NT];
} o_neuron;

typedef struct {
    float current_voltage, v_th, v_resting, dt;
    float chance_of_firing;     /* PoissonNeuron */
    float rate, step;           /* RateSpikeTrain */
    float internal_clock;       /* PresetSpikeTrain */
    uint32_t counter;
    float *firing_times;
    uint32_t n_firing_times;
    float k;                    /* neural_refractoriness.k */
    uint32_t is_spiking;
    int32_t last_firing_time;
    o_nt nt[ORC_NT];
} o_train;

typedef struct {
    uint32_t pre; /* canonical global node index */
    float w;
    /* TraceRSTDP plasticity/mod.rs:121-136 (w is its `weight`); all zero for plain f32 weights */
    uint32_t counter;
    float dw, c;
    /* RewardModulatedConnection neuron/mod.rs:3429-3443 for edges of a network's connecting graph: 0 = Weight(f32),
     * 1 = RewardModulatedWeight(TraceRSTDP) */
    uint32_t rm;
} o_edge;

typedef struct {
    uint64_t id;
    uint32_t rows, cols;
    uint64_t n;
    uint64_t base; /* canonical global index of cell 0 */
    int is_train;
    o_neuron *cells;
    o_train *trains;
    int do_plasticity, update_grid_history, update_spike_history;
    orc_stdp plasticity;
    /* member of RewardModulatedLatticeNetwork::reward_modulated_lattices (neuron/mod.rs:3470-3471): its own graph holds
     * TraceRSTDP weights and its plasticity is a RewardModulatedSTDP (neuron/mod.rs:2719-2751) */
    int is_reward, do_modulation;
    orc_rstdp rstdp;
    int use_bcm;          /* the lattice's Plasticity is BCM (plasticity/mod.rs:80-112) instead of STDP */
    orc_bcm bcm;
    uint64_t internal_clock;
    float *grid_history;
    /* AverageVoltageHistory / EEGHistory neuron/mod.rs:231-322 */
    int update_average_history, update_eeg_history;
    float eeg_reference_voltage, eeg_distance, eeg_conductivity;
    float *average_history, *eeg_history;
    uint64_t red_cap;
    uint8_t *spike_history;
    uint64_t hist_len, hist_cap;
} o_lattice;

struct orc_network {
    int model, ntk, rck, train_kind, refract_kind;
    o_lattice **lat; /* neuron lattices then train lattices, each group ascending id (canonical) */
    int n_lat;       /* total entries */
    int n_neuron_lat;
    uint64_t n_neurons, n_nodes;
    /* in-edges per neuron node (canonical index < n_neurons), sorted by pre ascending */
    o_edge **in;
    uint32_t *in_len, *in_cap;
    int electrical, chemical, parallel;
    /* RewardModulatedLattice neuron/mod.rs:2719-2751: the lattice's plasticity is a RewardModulatedSTDP over TraceRSTDP weights */
    int reward_mode, do_modulation;
    orc_rstdp rstdp;
    uint64_t internal_clock;
    uint64_t rng;
    /* scratch */
    float *inp_e;
    float *inp_t;       /* n_neurons * ORC_NT */
    uint8_t *inp_has;   /* n_neurons * ORC_NT */
    /* transposed index built lazily for STDP out-edges */
    uint64_t *out_ptr;
    uint32_t *out_post, *out_pos;
    int out_valid;
};

/* ------------------------------------------------------------------ defaults */

static void nt_default(o_nt *x) {
    /* ApproximateNeurotransmitter::default :174-182; Destexhe :136-145; ExponentialDecay :338-346;
     * DiscreteSpike :315-322 */
    x->present = 0;
    x->t = 0.f; x->t_max = 1.f; x->clearance_constant = 0.01f; x->v_p = 2.f; x->k_p = 5.f; x->decay_constant = 2.0f;
}

static void rc_default(o_rc *x, int type) {
    /* AMPAReceptor::default :1085-1094, NMDAReceptor::default :1115-1125, GABAReceptor::default :1148-1157,
     * DestexheReceptor::default :417-425, ApproximateReceptor::default :491-495, ExponentialDecayReceptor :525-533 */
    x->present = 0; x->current = 0.f; x->mg = 0.3f;
    if (type == 0) { x->g = 1.f; x->e = 0.f; }
    else if (type == 1) { x->g = 0.6f; x->e = 0.f; }
    else { x->g = 1.2f; x->e = -80.f; }
    x->r = 0.f; x->alpha = 1.f; x->beta = 1.f; x->r_max = 1.f; x->decay_constant = 2.f;
}

static void neuron_default(o_neuron *c, int model) {
    memset(c, 0, sizeof *c);
    c->last_firing_time = -1;
    for (int k = 0; k < ORC_NT; k++) { nt_default(&c->nt[k]); rc_default(&c->rc[k], k); }
    switch (model) {
    case ORC_LIF: /* integrate_and_fire/mod.rs:149-172 */
        c->current_voltage = -75.f; c->refractory_count = 0.f; c->leak_constant = -1.f; c->integration_constant = 1.f;
        c->gap_conductance = 7.f; c->v_th = -55.f; c->v_reset = -75.f; c->tau_m = 10.f; c->c_m = 100.f; c->g_l = 10.f;
        c->v_init = -75.f; c->e_l = -75.f; c->tref = 10.f; c->dt = 0.1f; break;
    case ORC_QIF: /* :298-320 */
        c->current_voltage = -75.f; c->refractory_count = 0.f; c->integration_constant = 1.f; c->gap_conductance = 7.f;
        c->alpha = 1.f; c->v_th = -55.f; c->v_reset = -75.f; c->v_c = -60.f; c->tau_m = 100.f; c->c_m = 100.f;
        c->v_init = -75.f; c->tref = 10.f; c->dt = 0.1f; break;
    case ORC_ADLIF: /* :970-997 */
    case ORC_ADEX:  /* :1106-1134 */
        c->current_voltage = -75.f; c->refractory_count = 0.f; c->leak_constant = -1.f; c->integration_constant = 1.f;
        c->gap_conductance = 7.f; c->w_value = 0.f; c->alpha = 6.0f; c->beta = 10.0f; c->slope_factor = 1.f;
        c->v_th = -55.f; c->v_reset = -75.f; c->tau_m = 10.f; c->c_m = 100.f; c->g_l = 10.f; c->v_init = -75.f;
        c->e_l = -75.f; c->tref = 10.f; c->w_init = 0.f; c->dt = 0.1f; break;
    case ORC_IZH: /* :1198-1220 */
        c->current_voltage = -65.f; c->gap_conductance = 7.f; c->w_value = 30.f; c->a = 0.02f; c->b = 0.2f;
        c->c = -55.0f; c->d = 8.0f; c->v_th = 30.f; c->tau_m = 1.f; c->c_m = 100.f; c->v_init = -65.f;
        c->w_init = 30.f; c->dt = 0.1f; break;
    case ORC_BCM_IZH: /* :1411-1438 */
        c->current_voltage = -65.f; c->gap_conductance = 7.f; c->w_value = 30.f; c->a = 0.02f; c->b = 0.2f;
        c->c = -55.0f; c->d = 8.0f; c->v_th = 30.f; c->tau_m = 1.f; c->c_m = 100.f; c->v_init = -65.f;
        c->w_init = 30.f; c->dt = 0.1f; c->average_activity = 0.f; c->current_activity = 0.f; c->period = 3;
        c->num_spikes = 0; c->firing_rate_clock = 0.f; c->firing_rate_window = 500.f; break;
    case ORC_LEAKY_IZH: /* :1313-1336 */
        c->current_voltage = -65.f; c->gap_conductance = 7.f; c->w_value = 30.f; c->a = 0.02f; c->b = 0.2f;
        c->c = -55.0f; c->d = 8.0f; c->v_th = 30.f; c->tau_m = 10.f; c->c_m = 100.f; c->v_init = -65.f;
        c->e_l = -65.f; c->w_init = 30.f; c->dt = 0.1f; break;
    case ORC_SIMPLE_LIF: /* :1552-1570 */
        c->current_voltage = -75.f; c->gap_conductance = 10.f; c->v_th = -55.f; c->v_reset = -75.f; c->c_m = 100.f;
        c->g = -0.1f; c->v_init = -75.f; c->e = 0.f; c->dt = 0.1f; break;
    case ORC_HH: /* hodgkin_huxley/mod.rs:80-99; ion_channels/mod.rs:205-215, 255-264, 299-307 */
        c->current_voltage = -65.f; c->gap_conductance = 7.f; c->dt = 0.01f; c->c_m = 1.f; c->v_th = 0.f;
        c->g_na = 120.f; c->e_na = 50.f; c->g_k = 36.f; c->e_k = -77.f; c->g_k_leak = 0.3f; c->e_k_leak = -55.f;
        break;
    }
}

static void train_default(o_train *s) {
    /* PoissonNeuron::default spike_train/mod.rs:297-311; RateSpikeTrain::default :997-1012;
     * PresetSpikeTrain::default :778-794; refractoriness default k = 10000 (:50-57) */
    memset(s, 0, sizeof *s);
    s->current_voltage = 0.f; s->v_th = 30.f; s->v_resting = 0.f; s->dt = 0.1f; s->k = 10000.f;
    s->last_firing_time = -1;
    for (int k = 0; k < ORC_NT; k++) nt_default(&s->nt[k]);
}

/* ------------------------------------------------------------------ field directory */

typedef struct { const char *name; int dtype; size_t off; int models; } o_fielddef;
#define M(x) (1 << (x))
#define ALLM 0x1FF
#define OFF(f) offsetof(o_neuron, f)
static const o_fielddef neuron_fields[] = {
    {"current_voltage", ORC_F32, OFF(current_voltage), ALLM},
    {"gap_conductance", ORC_F32, OFF(gap_conductance), ALLM},
    {"dt", ORC_F32, OFF(dt), ALLM},
    {"c_m", ORC_F32, OFF(c_m), ALLM},
    {"v_th", ORC_F32, OFF(v_th), ALLM},
    {"is_spiking", ORC_U32, OFF(is_spiking), ALLM},
    {"last_firing_time", ORC_I32, OFF(last_firing_time), ALLM},
    {"v_reset", ORC_F32, OFF(v_reset), M(ORC_LIF) | M(ORC_QIF) | M(ORC_ADLIF) | M(ORC_ADEX) | M(ORC_SIMPLE_LIF)},
    {"v_init", ORC_F32, OFF(v_init), ALLM & ~M(ORC_HH)},
    {"refractory_count", ORC_F32, OFF(refractory_count), M(ORC_LIF) | M(ORC_QIF) | M(ORC_ADLIF) | M(ORC_ADEX)},
    {"tref", ORC_F32, OFF(tref), M(ORC_LIF) | M(ORC_QIF) | M(ORC_ADLIF) | M(ORC_ADEX)},
    {"leak_constant", ORC_F32, OFF(leak_constant), M(ORC_LIF) | M(ORC_ADLIF) | M(ORC_ADEX)},
    {"integration_constant", ORC_F32, OFF(integration_constant), M(ORC_LIF) | M(ORC_QIF) | M(ORC_ADLIF) | M(ORC_ADEX)},
    {"e_l", ORC_F32, OFF(e_l), M(ORC_LIF) | M(ORC_ADLIF) | M(ORC_ADEX) | M(ORC_LEAKY_IZH)},
    {"g_l", ORC_F32, OFF(g_l), M(ORC_LIF) | M(ORC_ADLIF) | M(ORC_ADEX)},
    {"tau_m", ORC_F32, OFF(tau_m), ALLM & ~M(ORC_HH) & ~M(ORC_SIMPLE_LIF)},
    {"alpha", ORC_F32, OFF(alpha), M(ORC_QIF) | M(ORC_ADLIF) | M(ORC_ADEX)},
    {"beta", ORC_F32, OFF(beta), M(ORC_ADLIF) | M(ORC_ADEX)},
    {"slope_factor", ORC_F32, OFF(slope_factor), M(ORC_ADEX)},
    {"w_value", ORC_F32, OFF(w_value), M(ORC_ADLIF) | M(ORC_ADEX) | M(ORC_IZH) | M(ORC_LEAKY_IZH) | M(ORC_BCM_IZH)},
    {"w_init", ORC_F32, OFF(w_init), M(ORC_ADLIF) | M(ORC_ADEX) | M(ORC_IZH) | M(ORC_LEAKY_IZH) | M(ORC_BCM_IZH)},
    {"v_c", ORC_F32, OFF(v_c), M(ORC_QIF)},
    {"average_activity", ORC_F32, OFF(average_activity), M(ORC_BCM_IZH)},
    {"current_activity", ORC_F32, OFF(current_activity), M(ORC_BCM_IZH)},
    {"period", ORC_U32, OFF(period), M(ORC_BCM_IZH)},
    {"num_spikes", ORC_U32, OFF(num_spikes), M(ORC_BCM_IZH)},
    {"firing_rate_clock", ORC_F32, OFF(firing_rate_clock), M(ORC_BCM_IZH)},
    {"firing_rate_window", ORC_F32, OFF(firing_rate_window), M(ORC_BCM_IZH)},
    {"a", ORC_F32, OFF(a), M(ORC_IZH) | M(ORC_LEAKY_IZH) | M(ORC_BCM_IZH)},
    {"b", ORC_F32, OFF(b), M(ORC_IZH) | M(ORC_LEAKY_IZH) | M(ORC_BCM_IZH)},
    {"c", ORC_F32, OFF(c), M(ORC_IZH) | M(ORC_LEAKY_IZH) | M(ORC_BCM_IZH)},
    {"d", ORC_F32, OFF(d), M(ORC_IZH) | M(ORC_LEAKY_IZH) | M(ORC_BCM_IZH)},
    {"g", ORC_F32, OFF(g), M(ORC_SIMPLE_LIF)},
    {"e", ORC_F32, OFF(e), M(ORC_SIMPLE_LIF)},
    {"na_channel$g_na", ORC_F32, OFF(g_na), M(ORC_HH)},
    {"na_channel$e_na", ORC_F32, OFF(e_na), M(ORC_HH)},
    {"na_channel$current", ORC_F32, OFF(na_current), M(ORC_HH)},
    {"na_channel$m$alpha", ORC_F32, OFF(m.alpha), M(ORC_HH)},
    {"na_channel$m$beta", ORC_F32, OFF(m.beta), M(ORC_HH)},
    {"na_channel$m$state", ORC_F32, OFF(m.state), M(ORC_HH)},
    {"na_channel$h$alpha", ORC_F32, OFF(h.alpha), M(ORC_HH)},
    {"na_channel$h$beta", ORC_F32, OFF(h.beta), M(ORC_HH)},
    {"na_channel$h$state", ORC_F32, OFF(h.state), M(ORC_HH)},
    {"k_channel$g_k", ORC_F32, OFF(g_k), M(ORC_HH)},
    {"k_channel$e_k", ORC_F32, OFF(e_k), M(ORC_HH)},
    {"k_channel$current", ORC_F32, OFF(k_current), M(ORC_HH)},
    {"k_channel$n$alpha", ORC_F32, OFF(n.alpha), M(ORC_HH)},
    {"k_channel$n$beta", ORC_F32, OFF(n.beta), M(ORC_HH)},
    {"k_channel$n$state", ORC_F32, OFF(n.state), M(ORC_HH)},
    {"k_leak_channel$g_k_leak", ORC_F32, OFF(g_k_leak), M(ORC_HH)},
    {"k_leak_channel$e_k_leak", ORC_F32, OFF(e_k_leak), M(ORC_HH)},
    {"k_leak_channel$current", ORC_F32, OFF(k_leak_current), M(ORC_HH)},
    {"was_increasing", ORC_U32, OFF(was_increasing), M(ORC_HH)},
};
#define N_NEURON_FIELDS (sizeof neuron_fields / sizeof neuron_fields[0])

#define TOFF(f) offsetof(o_train, f)
static const o_fielddef train_fields[] = {
    {"current_voltage", ORC_F32, TOFF(current_voltage), ALLM},
    {"v_th", ORC_F32, TOFF(v_th), ALLM},
    {"v_resting", ORC_F32, TOFF(v_resting), ALLM},
    {"dt", ORC_F32, TOFF(dt), ALLM},
    {"is_spiking", ORC_U32, TOFF(is_spiking), ALLM},
    {"last_firing_time", ORC_I32, TOFF(last_firing_time), ALLM},
    {"neural_refractoriness$k", ORC_F32, TOFF(k), ALLM},
    {"chance_of_firing", ORC_F32, TOFF(chance_of_firing), M(ORC_TRAIN_POISSON)},
    {"rate", ORC_F32, TOFF(rate), M(ORC_TRAIN_RATE)},
    {"step", ORC_F32, TOFF(step), M(ORC_TRAIN_RATE)},
    {"internal_clock", ORC_F32, TOFF(internal_clock), M(ORC_TRAIN_PRESET)},
    {"counter", ORC_U32, TOFF(counter), M(ORC_TRAIN_PRESET)},
};
#define N_TRAIN_FIELDS (sizeof train_fields / sizeof train_fields[0])

/* per-type (n*3, neuron-major) neurotransmitter fields */
typedef struct { const char *name; int dtype; size_t off; } o_ntfielddef;
static const o_ntfielddef nt_fields[] = {
    {"neurotransmitters$flags", ORC_U32, offsetof(o_nt, present)},
    {"neurotransmitters$t", ORC_F32, offsetof(o_nt, t)},
    {"neurotransmitters$t_max", ORC_F32, offsetof(o_nt, t_max)},
    {"neurotransmitters$clearance_constant", ORC_F32, offsetof(o_nt, clearance_constant)},
    {"neurotransmitters$v_p", ORC_F32, offsetof(o_nt, v_p)},
    {"neurotransmitters$k_p", ORC_F32, offsetof(o_nt, k_p)},
    {"neurotransmitters$decay_constant", ORC_F32, offsetof(o_nt, decay_constant)},
};
#define N_NT_FIELDS (sizeof nt_fields / sizeof nt_fields[0])

static const char *rc_type_names[ORC_NT] = {"AMPA", "NMDA", "GABA"};
typedef struct { const char *suffix; size_t off; int nmda_only; } o_rcfielddef;
static const o_rcfielddef rc_fields[] = {
    {"_current", offsetof(o_rc, current), 0},
    {"_g", offsetof(o_rc, g), 0},
    {"_e", offsetof(o_rc, e), 0},
    {"_mg", offsetof(o_rc, mg), 1},
    {"$r$kinetics$r", offsetof(o_rc, r), 0},
    {"$r$kinetics$alpha", offsetof(o_rc, alpha), 0},
    {"$r$kinetics$beta", offsetof(o_rc, beta), 0},
    {"$r$kinetics$r_max", offsetof(o_rc, r_max), 0},
    {"$r$kinetics$decay_constant", offsetof(o_rc, decay_constant), 0},
};
#define N_RC_FIELDS (sizeof rc_fields / sizeof rc_fields[0])

/* ------------------------------------------------------------------ network bookkeeping */

orc_network *orc_network_create(int model, int ntk, int rck, int train_kind, int refract_kind) {
    orc_network *net = calloc(1, sizeof *net);
    net->model = model; net->ntk = ntk; net->rck = rck; net->train_kind = train_kind; net->refract_kind = refract_kind;
    net->electrical = 1; net->chemical = 0; /* Lattice::default neuron/mod.rs:589-606; LatticeNetwork::default :1577-1588 */
    net->rng = 0x9E3779B97F4A7C15ull;
    return net;
}

static void free_graph(orc_network *net) {
    if (net->in) {
        for (uint64_t i = 0; i < net->n_neurons; i++) free(net->in[i]);
        free(net->in); free(net->in_len); free(net->in_cap);
        net->in = NULL; net->in_len = net->in_cap = NULL;
    }
    free(net->out_ptr); free(net->out_post); free(net->out_pos);
    net->out_ptr = NULL; net->out_post = net->out_pos = NULL; net->out_valid = 0;
    free(net->inp_e); free(net->inp_t); free(net->inp_has);
    net->inp_e = net->inp_t = NULL; net->inp_has = NULL;
}

void orc_network_destroy(orc_network *net) {
    if (!net) return;
    free_graph(net);
    for (int i = 0; i < net->n_lat; i++) {
        o_lattice *L = net->lat[i];
        if (L->trains) for (uint64_t j = 0; j < L->n; j++) free(L->trains[j].firing_times);
        free(L->cells); free(L->trains); free(L->grid_history); free(L->spike_history); free(L->average_history); free(L->eeg_history); free(L);
    }
    free(net->lat); free(net);
}

static o_lattice *find_lat(orc_network *net, uint64_t id) {
    for (int i = 0; i < net->n_lat; i++) if (net->lat[i]->id == id) return net->lat[i];
    return NULL;
}

static void ensure_graph(orc_network *net);

/* Insert a lattice keeping the canonical order and re-index any edges that already exist (a lattice that
 * was connected internally before joining the network keeps its graph: neuron/mod.rs:1663-1678 moves the whole
 * Lattice, graph included). */
static int add_lat(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols, int is_train) {
    if (find_lat(net, id)) return 32; /* GraphIDAlreadyPresent neuron/mod.rs:1669-1671 */
    /* remember the old index space */
    int old_n_lat = net->n_lat;
    uint64_t old_n_neurons = net->n_neurons;
    uint64_t *old_base = malloc(sizeof(uint64_t) * (old_n_lat + 1));
    for (int i = 0; i < old_n_lat; i++) old_base[i] = net->lat[i]->base;
    o_lattice **old_order = malloc(sizeof(o_lattice *) * (old_n_lat + 1));
    for (int i = 0; i < old_n_lat; i++) old_order[i] = net->lat[i];
    o_edge **old_in = net->in; uint32_t *old_len = net->in_len, *old_cap = net->in_cap;
    net->in = NULL; net->in_len = net->in_cap = NULL;
    free(net->out_ptr); free(net->out_post); free(net->out_pos);
    net->out_ptr = NULL; net->out_post = net->out_pos = NULL; net->out_valid = 0;
    free(net->inp_e); free(net->inp_t); free(net->inp_has);
    net->inp_e = net->inp_t = NULL; net->inp_has = NULL;

    o_lattice *L = calloc(1, sizeof *L);
    L->id = id; L->rows = rows; L->cols = cols; L->n = (uint64_t)rows * cols; L->is_train = is_train;
    L->plasticity = (orc_stdp){2.f, 2.f, 4.5f, 4.5f, 0.1f}; /* STDP::default plasticity/mod.rs:29-39 */
    if (is_train) {
        L->trains = malloc(sizeof(o_train) * (L->n ? L->n : 1));
        for (uint64_t j = 0; j < L->n; j++) train_default(&L->trains[j]);
    } else {
        L->cells = malloc(sizeof(o_neuron) * (L->n ? L->n : 1));
        for (uint64_t j = 0; j < L->n; j++) neuron_default(&L->cells[j], net->model);
    }
    net->lat = realloc(net->lat, sizeof(o_lattice *) * (net->n_lat + 1));
    /* insert keeping [neuron lattices asc id][train lattices asc id] */
    int pos = net->n_lat;
    for (int i = 0; i < net->n_lat; i++) {
        o_lattice *X = net->lat[i];
        int before = (!is_train && X->is_train) || (is_train == X->is_train && id < X->id);
        if (before) { pos = i; break; }
    }
    memmove(&net->lat[pos + 1], &net->lat[pos], sizeof(o_lattice *) * (net->n_lat - pos));
    net->lat[pos] = L; net->n_lat++;
    if (!is_train) net->n_neuron_lat++;
    uint64_t base = 0; net->n_neurons = 0;
    for (int i = 0; i < net->n_lat; i++) {
        net->lat[i]->base = base; base += net->lat[i]->n;
        if (!net->lat[i]->is_train) net->n_neurons = base;
    }
    net->n_nodes = base;
    /* carry the existing edges over to the new index space */
    if (old_in) {
        ensure_graph(net);
        for (int i = 0; i < old_n_lat; i++) {
            o_lattice *B = old_order[i];
            if (B->is_train) continue;
            for (uint64_t q = 0; q < B->n; q++) {
                uint64_t orow = old_base[i] + q, nrow = B->base + q;
                net->in[nrow] = old_in[orow]; net->in_len[nrow] = old_len[orow]; net->in_cap[nrow] = old_cap[orow];
                for (uint32_t k = 0; k < old_len[orow]; k++) {
                    uint32_t p = net->in[nrow][k].pre;
                    for (int a = 0; a < old_n_lat; a++)
                        if (p >= old_base[a] && p < old_base[a] + old_order[a]->n) {
                            net->in[nrow][k].pre = (uint32_t)(p - old_base[a] + old_order[a]->base);
                            break;
                        }
                }
            }
        }
        (void)old_n_neurons;
        free(old_in); free(old_len); free(old_cap);
    }
    free(old_base); free(old_order);
    return 0;
}

int orc_add_lattice(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols) { return add_lat(net, id, rows, cols, 0); }
int orc_add_train_lattice(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols) { return add_lat(net, id, rows, cols, 1); }
uint64_t orc_lattice_size(orc_network *net, uint64_t id) { o_lattice *L = find_lat(net, id); return L ? L->n : 0; }

static void ensure_graph(orc_network *net) {
    if (net->in) return;
    uint64_t n = net->n_neurons ? net->n_neurons : 1;
    net->in = calloc(n, sizeof(o_edge *));
    net->in_len = calloc(n, sizeof(uint32_t));
    net->in_cap = calloc(n, sizeof(uint32_t));
}

/* ------------------------------------------------------------------ field access */

static int copy_field(void *cell_base, size_t stride, uint64_t n, size_t off, int fdtype, int dtype, void *data,
                      uint64_t count, uint64_t per, uint64_t sub, int set) {
    /* per = elements per cell in the flat array (1 or 3), sub = which of them this offset is */
    if (dtype != fdtype) return 66;
    if (count != n * per) return 67;
    for (uint64_t i = 0; i < n; i++) {
        char *p = (char *)cell_base + i * stride + off;
        char *q = (char *)data + (i * per + sub) * 4;
        if (set) memcpy(p, q, 4); else memcpy(q, p, 4);
    }
    return 0;
}

static int field_io(orc_network *net, uint64_t id, const char *name, void *data, uint64_t count, int dtype, int set) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    void *base = L->is_train ? (void *)L->trains : (void *)L->cells;
    size_t stride = L->is_train ? sizeof(o_train) : sizeof(o_neuron);
    size_t nt_off0 = L->is_train ? offsetof(o_train, nt) : offsetof(o_neuron, nt);
    const o_fielddef *fd = L->is_train ? train_fields : neuron_fields;
    size_t nfd = L->is_train ? N_TRAIN_FIELDS : N_NEURON_FIELDS;
    int kind = L->is_train ? net->train_kind : net->model;
    for (size_t i = 0; i < nfd; i++)
        if (!strcmp(fd[i].name, name) && (fd[i].models & M(kind)))
            return copy_field(base, stride, L->n, fd[i].off, fd[i].dtype, dtype, data, count, 1, 0, set);
    for (size_t i = 0; i < N_NT_FIELDS; i++)
        if (!strcmp(nt_fields[i].name, name)) {
            for (int k = 0; k < ORC_NT; k++) {
                int r = copy_field(base, stride, L->n, nt_off0 + k * sizeof(o_nt) + nt_fields[i].off, nt_fields[i].dtype,
                                   dtype, data, count, ORC_NT, k, set);
                if (r) return r;
            }
            return 0;
        }
    if (!L->is_train) {
        if (!strcmp(name, "receptors$flags")) {
            for (int k = 0; k < ORC_NT; k++) {
                int r = copy_field(base, stride, L->n, offsetof(o_neuron, rc) + k * sizeof(o_rc) + offsetof(o_rc, present),
                                   ORC_U32, dtype, data, count, ORC_NT, k, set);
                if (r) return r;
            }
            return 0;
        }
        for (int k = 0; k < ORC_NT; k++)
            for (size_t i = 0; i < N_RC_FIELDS; i++) {
                if (rc_fields[i].nmda_only && k != 1) continue;
                char buf[96];
                snprintf(buf, sizeof buf, "receptors$%s%s", rc_type_names[k], rc_fields[i].suffix);
                if (!strcmp(buf, name))
                    return copy_field(base, stride, L->n, offsetof(o_neuron, rc) + k * sizeof(o_rc) + rc_fields[i].off,
                                      ORC_F32, dtype, data, count, 1, 0, set);
            }
    }
    return 65;
}

int orc_set_field(orc_network *net, uint64_t id, const char *name, const void *data, uint64_t count, int dtype) {
    return field_io(net, id, name, (void *)data, count, dtype, 1);
}
int orc_get_field(orc_network *net, uint64_t id, const char *name, void *out, uint64_t count, int dtype) {
    return field_io(net, id, name, out, count, dtype, 0);
}

static int fill_any(orc_network *net, uint64_t id, const char *name, uint32_t bits, int dtype) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    /* try per-cell then per-type sizes */
    uint64_t sizes[2] = {L->n, L->n * ORC_NT};
    for (int s = 0; s < 2; s++) {
        uint32_t *tmp = malloc(4 * (sizes[s] ? sizes[s] : 1));
        for (uint64_t i = 0; i < sizes[s]; i++) tmp[i] = bits;
        int r = field_io(net, id, name, tmp, sizes[s], dtype, 1);
        free(tmp);
        if (r != 67) return r;
    }
    return 67;
}
int orc_fill_field_f32(orc_network *net, uint64_t id, const char *name, float v) { uint32_t b; memcpy(&b, &v, 4); return fill_any(net, id, name, b, ORC_F32); }
int orc_fill_field_u32(orc_network *net, uint64_t id, const char *name, uint32_t v) { return fill_any(net, id, name, v, ORC_U32); }
int orc_fill_field_i32(orc_network *net, uint64_t id, const char *name, int32_t v) { uint32_t b; memcpy(&b, &v, 4); return fill_any(net, id, name, b, ORC_I32); }

int orc_set_preset_firing_times(orc_network *net, uint64_t id, const uint64_t *offsets, const float *times,
                                uint64_t n_trains, uint64_t n_times) {
    o_lattice *L = find_lat(net, id);
    if (!L || !L->is_train) return 35;
    if (n_trains != L->n || offsets[n_trains] != n_times) return 67;
    for (uint64_t i = 0; i < L->n; i++) {
        o_train *s = &L->trains[i];
        free(s->firing_times);
        s->n_firing_times = (uint32_t)(offsets[i + 1] - offsets[i]);
        s->firing_times = malloc(sizeof(float) * (s->n_firing_times ? s->n_firing_times : 1));
        memcpy(s->firing_times, times + offsets[i], sizeof(float) * s->n_firing_times);
    }
    return 0;
}

/* ------------------------------------------------------------------ graph editing
 * Semantics of Graph::edit_weight over AdjacencyMatrix / AdjacencyList (graph/mod.rs:208-226,
 * 1046-1079): Some(w) connects (Some(0.0) is still a connection and counts in the averaging
 * denominator), None disconnects. */

static void edge_set(orc_network *net, uint64_t post, uint32_t pre, int has, float w) {
    o_edge *e = net->in[post];
    uint32_t len = net->in_len[post];
    uint32_t lo = 0, hi = len;
    while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (e[mid].pre < pre) lo = mid + 1; else hi = mid; }
    int found = lo < len && e[lo].pre == pre;
    if (has) {
        if (found) { e[lo].w = w; return; }
        if (len == net->in_cap[post]) {
            net->in_cap[post] = net->in_cap[post] ? net->in_cap[post] * 2 : 8;
            e = net->in[post] = realloc(e, sizeof(o_edge) * net->in_cap[post]);
        }
        memmove(&e[lo + 1], &e[lo], sizeof(o_edge) * (len - lo));
        e[lo].pre = pre; e[lo].w = w; e[lo].counter = 0; e[lo].dw = 0.f; e[lo].c = 0.f; e[lo].rm = 0; net->in_len[post] = len + 1;
    } else if (found) {
        memmove(&e[lo], &e[lo + 1], sizeof(o_edge) * (len - lo - 1));
        net->in_len[post] = len - 1;
    }
    net->out_valid = 0;
}

static int check_connect(orc_network *net, uint64_t pre_id, uint64_t post_id, o_lattice **A, o_lattice **B) {
    /* LatticeNetwork::connect preconditions, neuron/mod.rs:1852-1862 */
    o_lattice *pre = find_lat(net, pre_id), *post = find_lat(net, post_id);
    if (post && post->is_train) return 36;
    if (!pre) return 34;
    if (!post) return 33;
    *A = pre; *B = post;
    ensure_graph(net);
    return 0;
}

static void clear_block(orc_network *net, o_lattice *A, o_lattice *B) {
    for (uint64_t q = 0; q < B->n; q++) {
        uint64_t post = B->base + q;
        o_edge *e = net->in[post];
        uint32_t len = net->in_len[post], w = 0;
        for (uint32_t k = 0; k < len; k++)
            if (!(e[k].pre >= A->base && e[k].pre < A->base + A->n)) e[w++] = e[k];
        net->in_len[post] = w;
    }
    net->out_valid = 0;
}

int orc_connect_dense(orc_network *net, uint64_t pre_id, uint64_t post_id, const uint32_t *connections,
                      const float *weights, uint64_t n_pre, uint64_t n_post) {
    o_lattice *A, *B;
    int r = check_connect(net, pre_id, post_id, &A, &B);
    if (r) return r;
    if (n_pre != A->n || n_post != B->n) return 67;
    clear_block(net, A, B);
    for (uint64_t q = 0; q < n_post; q++)
        for (uint64_t p = 0; p < n_pre; p++)
            if (connections[p * n_post + q]) edge_set(net, B->base + q, (uint32_t)(A->base + p), 1, weights[p * n_post + q]);
    return 0;
}

int orc_connect_csr(orc_network *net, uint64_t pre_id, uint64_t post_id, const uint64_t *row_ptr,
                    const uint32_t *pre, const float *weights, uint64_t n_post, uint64_t nnz) {
    o_lattice *A, *B;
    int r = check_connect(net, pre_id, post_id, &A, &B);
    if (r) return r;
    if (n_post != B->n || row_ptr[n_post] != nnz) return 67;
    for (uint64_t k = 0; k < nnz; k++) if (pre[k] >= A->n) return 16;
    clear_block(net, A, B);
    for (uint64_t q = 0; q < n_post; q++)
        for (uint64_t k = row_ptr[q]; k < row_ptr[q + 1]; k++)
            edge_set(net, B->base + q, (uint32_t)(A->base + pre[k]), 1, weights[k]);
    return 0;
}

int orc_connect_grid(orc_network *net, uint64_t id, uint32_t radius, float weight) {
    o_lattice *A, *B;
    int r = check_connect(net, id, id, &A, &B);
    if (r) return r;
    clear_block(net, A, A);
    long R = (long)radius, rows = A->rows, cols = A->cols;
    for (long i = 0; i < rows; i++)
        for (long j = 0; j < cols; j++) {
            uint64_t post = A->base + (uint64_t)(i * cols + j);
            for (long di = -R; di <= R; di++)
                for (long dj = -R; dj <= R; dj++) {
                    long a = i + di, b = j + dj;
                    if ((di == 0 && dj == 0) || a < 0 || b < 0 || a >= rows || b >= cols) continue;
                    edge_set(net, post, (uint32_t)(A->base + (uint64_t)(a * cols + b)), 1, weight);
                }
        }
    return 0;
}

uint64_t orc_connection_nnz(orc_network *net, uint64_t pre_id, uint64_t post_id) {
    o_lattice *A = find_lat(net, pre_id), *B = find_lat(net, post_id);
    if (!A || !B || B->is_train || !net->in) return 0;
    uint64_t c = 0;
    for (uint64_t q = 0; q < B->n; q++) {
        uint64_t post = B->base + q;
        for (uint32_t k = 0; k < net->in_len[post]; k++) {
            uint32_t p = net->in[post][k].pre;
            if (p >= A->base && p < A->base + A->n) c++;
        }
    }
    return c;
}

int orc_get_connection_csr(orc_network *net, uint64_t pre_id, uint64_t post_id, uint64_t *row_ptr, uint32_t *pre,
                           float *weights) {
    o_lattice *A = find_lat(net, pre_id), *B = find_lat(net, post_id);
    if (!A || !B || B->is_train) return 35;
    ensure_graph(net);
    uint64_t c = 0;
    for (uint64_t q = 0; q < B->n; q++) {
        uint64_t post = B->base + q;
        if (row_ptr) row_ptr[q] = c;
        for (uint32_t k = 0; k < net->in_len[post]; k++) {
            uint32_t p = net->in[post][k].pre;
            if (p >= A->base && p < A->base + A->n) {
                if (pre) pre[c] = (uint32_t)(p - A->base);
                if (weights) weights[c] = net->in[post][k].w;
                c++;
            }
        }
    }
    if (row_ptr) row_ptr[B->n] = c;
    return 0;
}

int orc_get_connection_dense(orc_network *net, uint64_t pre_id, uint64_t post_id, uint32_t *connections,
                             float *weights, uint64_t n_pre, uint64_t n_post) {
    o_lattice *A = find_lat(net, pre_id), *B = find_lat(net, post_id);
    if (!A || !B || B->is_train) return 35;
    if (n_pre != A->n || n_post != B->n) return 67;
    ensure_graph(net);
    for (uint64_t i = 0; i < n_pre * n_post; i++) { if (connections) connections[i] = 0; if (weights) weights[i] = 0.f; }
    for (uint64_t q = 0; q < n_post; q++) {
        uint64_t post = B->base + q;
        for (uint32_t k = 0; k < net->in_len[post]; k++) {
            uint32_t p = net->in[post][k].pre;
            if (p >= A->base && p < A->base + A->n) {
                uint64_t idx = (uint64_t)(p - A->base) * n_post + q;
                if (connections) connections[idx] = 1;
                if (weights) weights[idx] = net->in[post][k].w;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ options */

void orc_set_synapses(orc_network *net, int electrical, int chemical) { net->electrical = electrical; net->chemical = chemical; }
void orc_set_parallel(orc_network *net, int parallel) { net->parallel = parallel; }
void orc_set_clock(orc_network *net, uint64_t clock) {
    net->internal_clock = clock;
    for (int i = 0; i < net->n_lat; i++) net->lat[i]->internal_clock = clock;
}
uint64_t orc_get_clock(orc_network *net) { return net->internal_clock; }
void orc_seed(orc_network *net, uint64_t seed) { net->rng = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull; if (!net->rng) net->rng = 1; }

int orc_set_lattice_flags(orc_network *net, uint64_t id, int do_plasticity, int update_grid_history,
                          int update_spike_history) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    L->do_plasticity = do_plasticity; L->update_grid_history = update_grid_history;
    L->update_spike_history = update_spike_history;
    return 0;
}

int orc_set_reduced_history(orc_network *net, uint64_t id, int average, int eeg, float reference_voltage, float distance,
                            float conductivity) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    L->update_average_history = average; L->update_eeg_history = eeg;
    L->eeg_reference_voltage = reference_voltage; L->eeg_distance = distance; L->eeg_conductivity = conductivity;
    return 0;
}

int orc_get_reduced_history(orc_network *net, uint64_t id, int eeg, float *out, uint64_t capacity) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    if (capacity < L->hist_len) return 67;
    const float *h = eeg ? L->eeg_history : L->average_history;
    if (h) memcpy(out, h, sizeof(float) * L->hist_len);
    return 0;
}

int orc_set_bcm_plasticity(orc_network *net, uint64_t id, int enable, const orc_bcm *bcm) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    L->use_bcm = enable;
    if (bcm) L->bcm = *bcm;
    return 0;
}

/* BCM::update_weight plasticity/mod.rs:101-106 (BCMActivity of BCMIzhikevichNeuron: current / average activity) */
static float bcm_update(const orc_bcm *b, float weight, const o_neuron *pre, const o_neuron *post) {
    float sliding_threshold = post->average_activity / b->average_scalar;
    float activity_term = post->current_activity * (post->current_activity - sliding_threshold);
    float weight_decay = b->decay * weight;
    return weight + (activity_term * pre->current_activity - weight_decay) * b->dt;
}

int orc_set_plasticity(orc_network *net, uint64_t id, const orc_stdp *stdp) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    L->plasticity = *stdp;
    return 0;
}

/* LatticeNetwork::set_dt neuron/mod.rs:1655-1660 -> Lattice::set_dt :649-652 (neurons + plasticity),
 * SpikeTrainLattice::set_dt :1355-1357; PoissonNeuron::set_dt rescales chance_of_firing
 * (spike_train/mod.rs:345-349) */
void orc_set_dt(orc_network *net, float dt) {
    for (int i = 0; i < net->n_lat; i++) {
        o_lattice *L = net->lat[i];
        if (L->is_train) {
            for (uint64_t j = 0; j < L->n; j++) {
                o_train *s = &L->trains[j];
                if (net->train_kind == ORC_TRAIN_POISSON) {
                    float scalar = dt / s->dt;
                    s->chance_of_firing *= scalar;
                }
                s->dt = dt;
            }
        } else {
            for (uint64_t j = 0; j < L->n; j++) L->cells[j].dt = dt;
            L->plasticity.dt = dt; L->bcm.dt = dt;
            if (L->is_reward) L->rstdp.dt = dt; /* RewardModulatedLattice::set_dt neuron/mod.rs:2867-2870 */
        }
    }
    if (net->reward_mode) net->rstdp.dt = dt;
}

/* reset_timing neuron/mod.rs:405-420, 1710-1717 */
void orc_reset_timing(orc_network *net) {
    net->internal_clock = 0;
    for (int i = 0; i < net->n_lat; i++) {
        o_lattice *L = net->lat[i];
        L->internal_clock = 0;
        for (uint64_t j = 0; j < L->n; j++) {
            if (L->is_train) L->trains[j].last_firing_time = -1; else L->cells[j].last_firing_time = -1;
        }
    }
}

/* ------------------------------------------------------------------ kinetics */

/* NeurotransmitterKinetics::apply_t_change given NeurotransmittersIntermediate{current_voltage,
 * is_spiking, dt} (intermediate_delegate/mod.rs:10-24) */
static void nt_apply_t_change(o_nt *x, int kind, float voltage, int is_spiking, float dt) {
    float flag = is_spiking ? 1.f : 0.f; /* bool_to_float :184-190 */
    switch (kind) {
    case ORC_NTK_APPROX: { /* iterate_and_spike/mod.rs:193-196 */
        float a = dt * -x->clearance_constant;
        float b = a * x->t;
        float c = flag * x->t_max;
        x->t = x->t + (b + c);
        x->t = fminf(x->t_max, fmaxf(x->t, 0.f));
        break;
    }
    case ORC_NTK_DESTEXHE: /* :148-150 */
        x->t = x->t_max / (1.f + expf(-(voltage - x->v_p) / x->k_p));
        break;
    case ORC_NTK_DISCRETE: /* :302-304 */
        x->t = x->t_max * flag;
        break;
    case ORC_NTK_EXPDECAY: { /* :348-357 */
        float t_change = -x->t * expf(dt / -x->decay_constant);
        x->t = x->t + (t_change + flag * x->t_max);
        x->t = fminf(x->t_max, fmaxf(x->t, 0.f));
        break;
    }
    }
}

/* Neurotransmitters::apply_t_changes :2245-2248 */
static void nts_apply(o_nt *nt, int kind, float voltage, int is_spiking, float dt) {
    for (int k = 0; k < ORC_NT; k++) if (nt[k].present) nt_apply_t_change(&nt[k], kind, voltage, is_spiking, dt);
}

/* ReceptorKinetics::apply_r_change :403-406, 434-437, 510-514 */
static void rc_apply_r_change(o_rc *x, int kind, float t, float dt) {
    switch (kind) {
    case ORC_RCK_APPROX: x->r = t; break;
    case ORC_RCK_DESTEXHE: {
        float a = x->alpha * t;
        float b = a * (1.f - x->r);
        float c = x->beta * x->r;
        x->r = x->r + (b - c) * dt;
        break;
    }
    case ORC_RCK_EXPDECAY: {
        float dec = -x->r * expf(dt / -x->decay_constant);
        x->r = x->r + (dec + t);
        x->r = fminf(x->r_max, fmaxf(x->r, 0.f));
        break;
    }
    }
}

/* Ionotropic::update_receptor_kinetics :1186-1206 — only types present in both the input map and the
 * neuron's receptor map are touched */
static void receptors_update_kinetics(o_neuron *c, int rck, const float *t, const uint8_t *has, float dt) {
    for (int k = 0; k < ORC_NT; k++)
        if (has[k] && c->rc[k].present) rc_apply_r_change(&c->rc[k], rck, t[k], dt);
}

/* Ionotropic::set_receptor_currents :1259-1284 with AMPA/NMDA/GABA::iterate :1101-1103, 1132-1137, 1164-1166 */
static void receptors_set_currents(o_neuron *c, float v) {
    if (c->rc[0].present) c->rc[0].current = (c->rc[0].g * c->rc[0].r) * (v - c->rc[0].e);
    if (c->rc[1].present) {
        o_rc *x = &c->rc[1];
        float ex = expf(-0.062f * v);
        float den = 1.0f + ((ex * x->mg) / 3.75f);
        x->current = (((1.0f / den) * x->g) * x->r) * (v - x->e);
    }
    if (c->rc[2].present) c->rc[2].current = (c->rc[2].g * c->rc[2].r) * (v - c->rc[2].e);
}

/* Ionotropic::get_receptor_currents :1286-1304 */
static float receptors_get_currents(const o_neuron *c, float dt, float c_m) {
    float total = 0.f;
    if (c->rc[0].present) total += c->rc[0].current;
    if (c->rc[1].present) total += c->rc[1].current;
    if (c->rc[2].present) total += c->rc[2].current;
    return total * (dt / c_m);
}

/* ------------------------------------------------------------------ neuron models */

/* impl_default_handle_spiking! integrate_and_fire/mod.rs:83-104 */
static int default_handle_spiking(o_neuron *c) {
    int is_spiking = 0;
    if (c->refractory_count > 0.f) {
        c->current_voltage = c->v_reset;
        c->refractory_count -= 1.f;
    } else if (c->current_voltage >= c->v_th) {
        is_spiking = 1;
        c->current_voltage = c->v_reset;
        c->refractory_count = c->tref / c->dt;
    }
    c->is_spiking = (uint32_t)is_spiking;
    return is_spiking;
}

/* adaptive_handle_spiking :1013-1029 */
static int adaptive_handle_spiking(o_neuron *c) {
    int is_spiking = 0;
    if (c->refractory_count > 0.f) {
        c->current_voltage = c->v_reset;
        c->refractory_count -= 1.f;
    } else if (c->current_voltage >= c->v_th) {
        is_spiking = 1;
        c->current_voltage = c->v_reset;
        c->w_value += c->beta;
        c->refractory_count = c->tref / c->dt;
    }
    c->is_spiking = (uint32_t)is_spiking;
    return is_spiking;
}

/* izhikevich_handle_spiking :1235-1247 */
static int izhikevich_handle_spiking(o_neuron *c) {
    int is_spiking = 0;
    if (c->current_voltage >= c->v_th) {
        is_spiking = 1;
        c->current_voltage = c->c;
        c->w_value += c->d;
    }
    c->is_spiking = (uint32_t)is_spiking;
    return is_spiking;
}

/* SimpleLeakyIntegrateAndFire::handle_spiking :1579-1590 */
static int simple_handle_spiking(o_neuron *c) {
    int is_spiking = 0;
    if (c->current_voltage >= c->v_th) { is_spiking = 1; c->current_voltage = c->v_reset; }
    c->is_spiking = (uint32_t)is_spiking;
    return is_spiking;
}

static float get_dv(const o_neuron *c, int model, float i) {
    float v = c->current_voltage;
    switch (model) {
    case ORC_LIF: /* leaky_get_dv_change :176-181 */
        return ((c->leak_constant * (v - c->e_l)) + (c->integration_constant * (i / c->g_l))) * (c->dt / c->tau_m);
    case ORC_QIF: /* quadratic_get_dv_change :324-327 */
        return (((c->alpha * (v - c->v_reset)) * (v - c->v_c)) + c->integration_constant * i) * (c->dt / c->tau_m);
    case ORC_ADLIF: /* adaptive_get_dv_change :1035-1041 */
        return (((c->leak_constant * (v - c->e_l)) + (c->integration_constant * (i / c->g_l))) - (c->w_value / c->g_l)) *
               (c->dt / c->c_m);
    case ORC_ADEX: /* exp_adaptive_get_dv_change :1138-1145 */
        return ((((c->leak_constant * (v - c->e_l)) + (c->slope_factor * expf((v - c->v_th) / c->slope_factor))) +
                 (c->integration_constant * (i / c->g_l))) - (c->w_value / c->g_l)) * (c->dt / c->c_m);
    case ORC_IZH: case ORC_BCM_IZH: /* izhikevich_get_dv_change :1255-1260, :1446-1451 (powf(2.0) == v*v) */
        return (((((0.04f * (v * v)) + (5.f * v)) + 140.f) - c->w_value) + i) * (c->dt / c->c_m);
    case ORC_LEAKY_IZH: /* izhikevich_leaky_get_dv_change :1342-1348 */
        return (((((0.04f * (v * v)) + (5.f * v)) + 140.f) - (c->w_value * (v - c->e_l))) + i) * (c->dt / c->c_m);
    case ORC_SIMPLE_LIF: /* get_dv_change :1592-1594 */
        return (c->g * (v - c->e) + i) * c->dt;
    }
    return 0.f;
}

static float get_dw(const o_neuron *c, int model) {
    switch (model) {
    case ORC_ADLIF: case ORC_ADEX: /* adaptive_get_dw_change :1002-1009 */
        return (c->alpha * (c->current_voltage - c->e_l) - c->w_value) * (c->dt / c->tau_m);
    case ORC_IZH: case ORC_LEAKY_IZH: case ORC_BCM_IZH: /* izhikevich_get_dw_change :1225-1231 */
        return (c->a * (c->b * c->current_voltage - c->w_value)) * (c->dt / c->tau_m);
    }
    return 0.f;
}

static int handle_spiking(o_neuron *c, int model) {
    switch (model) {
    case ORC_LIF: case ORC_QIF: return default_handle_spiking(c);
    case ORC_ADLIF: case ORC_ADEX: return adaptive_handle_spiking(c);
    case ORC_IZH: case ORC_LEAKY_IZH: case ORC_BCM_IZH: return izhikevich_handle_spiking(c);
    case ORC_SIMPLE_LIF: return simple_handle_spiking(c);
    }
    return 0;
}

/* BasicGatingVariable::update ion_channels/mod.rs:40-44 */
static void gate_update(o_gate *g, float dt) {
    float alpha_state = g->alpha * (1.f - g->state);
    float beta_state = g->beta * g->state;
    g->state += dt * (alpha_state - beta_state);
}

/* HodgkinHuxleyNeuron::update_gates hodgkin_huxley/mod.rs:182-186; NaIonChannel ion_channels/mod.rs:219-235;
 * KIonChannel :268-281; KLeakChannel :310-312 */
static void hh_update_gates(o_neuron *c) {
    float v = c->current_voltage, dt = c->dt;
    c->m.alpha = 0.1f * ((v + 40.f) / (1.f - expf(-(v + 40.f) / 10.f)));
    c->m.beta = 4.f * expf(-(v + 65.f) / 18.f);
    c->h.alpha = 0.07f * expf(-(v + 65.f) / 20.f);
    c->h.beta = 1.f / (expf(-(v + 35.f) / 10.f) + 1.f);
    gate_update(&c->m, dt);
    gate_update(&c->h, dt);
    c->na_current = ((powf(c->m.state, 3.f) * c->h.state) * c->g_na) * (v - c->e_na);
    c->n.alpha = (0.01f * (v + 55.f)) / (1.f - expf(-(v + 55.f) / 10.f));
    c->n.beta = 0.125f * expf(-(v + 65.f) / 80.f);
    gate_update(&c->n, dt);
    c->k_current = (powf(c->n.state, 4.f) * c->g_k) * (v - c->e_k);
    c->k_leak_current = c->g_k_leak * (v - c->e_k_leak);
}

/* HodgkinHuxleyNeuron::update_cell_voltage hodgkin_huxley/mod.rs:156-165 */
static void hh_update_cell_voltage(o_neuron *c, float input_current) {
    float i_ligand_gates = receptors_get_currents(c, c->dt, c->c_m);
    float i_sum = input_current - ((c->na_current + c->k_current) + c->k_leak_current);
    c->current_voltage += (c->dt * i_sum) / c->c_m - i_ligand_gates;
}

/* IterateAndSpike::iterate_and_spike / iterate_with_neurotransmitter_and_spike for every model.
 * chem != 0 selects the *_with_neurotransmitter_* variant. */
static int neuron_iterate(o_neuron *c, int model, int ntk, int rck, float input, int chem, const float *t,
                          const uint8_t *has) {
    if (model == ORC_HH) { /* hodgkin_huxley/mod.rs:188-241 */
        float last_voltage = c->current_voltage;
        if (chem) { /* update_receptors :173-179 */
            receptors_update_kinetics(c, rck, t, has, c->dt);
            receptors_set_currents(c, c->current_voltage);
        }
        hh_update_gates(c);
        hh_update_cell_voltage(c, input);
        nts_apply(c->nt, ntk, c->current_voltage, (int)c->is_spiking, c->dt); /* update_neurotransmitters :168-170 */
        int increasing_right_now = last_voltage < c->current_voltage;
        int threshold_crossed = c->current_voltage > c->v_th;
        int is_spiking = threshold_crossed && c->was_increasing && !increasing_right_now;
        c->is_spiking = (uint32_t)is_spiking;
        c->was_increasing = (uint32_t)increasing_right_now;
        return is_spiking;
    }
    /* integrate_and_fire/mod.rs:189-214 (LIF), :222-252 (macro: AdLIF, AdEx, Izh, LeakyIzh), :339-364 (QIF),
     * :1604-1629 (SimpleLIF) */
    int adaptive = model == ORC_ADLIF || model == ORC_ADEX || model == ORC_IZH || model == ORC_LEAKY_IZH || model == ORC_BCM_IZH;
    if (model == ORC_BCM_IZH) { /* BCMIzhikevichNeuron :1458-1467 (iterate_and_spike), :1485-1494 (with neurotransmitter) */
        if (c->is_spiking) c->num_spikes += 1;
        c->firing_rate_clock += c->dt;
        if (c->firing_rate_clock >= c->firing_rate_window) {
            c->firing_rate_clock = 0.f;
            if (chem) c->current_activity = (float)c->num_spikes / c->firing_rate_window;
            else c->current_activity = (float)c->num_spikes / (c->firing_rate_window * c->dt);
            c->average_activity -= c->average_activity / (float)c->period;
            c->average_activity += c->current_activity / (float)c->period;
        }
    }
    if (chem) {
        receptors_update_kinetics(c, rck, t, has, c->dt);
        receptors_set_currents(c, c->current_voltage);
    }
    float dv = get_dv(c, model, input);
    float dw = adaptive ? get_dw(c, model) : 0.f;
    if (chem) {
        float neurotransmitter_dv = -receptors_get_currents(c, c->dt, c->c_m);
        c->current_voltage += dv + neurotransmitter_dv;
    } else {
        c->current_voltage += dv;
    }
    if (adaptive) c->w_value += dw;
    nts_apply(c->nt, ntk, c->current_voltage, (int)c->is_spiking, c->dt);
    return handle_spiking(c, model);
}

/* ------------------------------------------------------------------ spike trains */

static float rng_uniform(orc_network *net) {
    /* stand-in for rand::thread_rng().gen_range(0.0..=1.0) (spike_train/mod.rs:354): statistics only */
    uint64_t x = net->rng;
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    net->rng = x;
    return (float)((x >> 40) * (1.0 / 16777216.0));
}

static int train_iterate(orc_network *net, o_train *s) {
    int is_spiking = 0;
    switch (net->train_kind) {
    case ORC_TRAIN_POISSON: /* spike_train/mod.rs:352-368 */
        if (rng_uniform(net) <= s->chance_of_firing) { s->current_voltage = s->v_th; is_spiking = 1; }
        else { s->current_voltage = s->v_resting; }
        s->is_spiking = (uint32_t)is_spiking;
        break;
    case ORC_TRAIN_RATE: /* :1015-1030 */
        s->step += s->dt;
        if (s->rate != 0.f && s->step >= s->rate) { s->step = 0.f; s->current_voltage = s->v_th; s->is_spiking = 1; }
        else { s->current_voltage = s->v_resting; s->is_spiking = 0; }
        is_spiking = (int)s->is_spiking;
        break;
    case ORC_TRAIN_PRESET: /* :803-828 */
        s->internal_clock += s->dt;
        if (s->n_firing_times && s->internal_clock > s->firing_times[s->counter]) {
            s->current_voltage = s->v_th;
            s->internal_clock = 0.f;
            s->counter += 1;
            if (s->counter == s->n_firing_times) s->counter = 0;
            is_spiking = 1;
        } else {
            s->current_voltage = s->v_resting;
        }
        s->is_spiking = (uint32_t)is_spiking;
        break;
    }
    /* spike trains release with the CURRENT step's flag (is_spiking is assigned before apply_t_changes) */
    nts_apply(s->nt, net->ntk, s->current_voltage, is_spiking, s->dt);
    return is_spiking;
}

float orc_refractoriness_effect(int kind, float k, uint64_t timestep, uint64_t last_firing_time, float v_max,
                                float v_resting, float dt) {
    /* impl_default_neural_refractoriness! get_effect spike_train/mod.rs:68-73; delta_dirac_effect :84-86;
     * exponential_decay_effect :174-176 */
    float a = v_max - v_resting;
    float time_difference = (float)(timestep - last_firing_time);
    if (kind == ORC_REFRACT_DELTA_DIRAC)
        return a * expf((-1.f / (k / dt)) * (time_difference * time_difference)) + v_resting;
    return a * expf((-1.f / (k / dt)) * time_difference) + v_resting;
}

float orc_chance_from_firing_rate(float hertz, float dt) {
    /* PoissonNeuron::from_firing_rate spike_train/mod.rs:330-337 */
    return 1.f / ((1000.f / dt) / hertz);
}

/* ------------------------------------------------------------------ plasticity */

float orc_stdp_update(const orc_stdp *p, float weight, int32_t t_pre_i, int32_t t_post_i) {
    /* STDP::update_weight plasticity/mod.rs:46-65 */
    float delta_w = 0.f;
    if (t_pre_i >= 0 && t_post_i >= 0) {
        float t_pre = (float)t_pre_i, t_post = (float)t_post_i;
        if (t_pre < t_post) {
            delta_w = p->a_plus * expf((-1.f * fabsf((t_pre - t_post) * p->dt)) / p->tau_plus);
        } else if (t_pre > t_post) {
            delta_w = (-1.f * p->a_minus) * expf((-1.f * fabsf((t_post - t_pre) * p->dt)) / p->tau_minus);
        }
    }
    return weight + delta_w;
}

/* ------------------------------------------------------------------ stepping */

static o_lattice *lat_of_node(orc_network *net, uint64_t node) {
    for (int i = 0; i < net->n_lat; i++)
        if (node >= net->lat[i]->base && node < net->lat[i]->base + net->lat[i]->n) return net->lat[i];
    return NULL;
}

typedef struct { o_lattice *L; uint64_t idx; } o_ref;

static inline o_ref node_ref(orc_network *net, uint64_t node) {
    o_ref r; r.L = lat_of_node(net, node); r.idx = node - r.L->base; return r;
}

/* Lattice::calculate_internal_electrical_input_from_positions neuron/mod.rs:702-730 and
 * LatticeNetwork::calculate_electrical_input_from_positions :2115-2167 (union of internal and
 * connecting in-edges, one common denominator; gap_junction :54-60; spike_train_gap_junction :119-137) */
static float electrical_input(orc_network *net, o_lattice *B, uint64_t q) {
    uint64_t post = B->base + q;
    const o_neuron *pn = &B->cells[q];
    uint32_t len = net->in ? net->in_len[post] : 0;
    float input_val = 0.f;
    for (uint32_t k = 0; k < len; k++) {
        o_edge e = net->in[post][k];
        o_ref r = (e.pre >= B->base && e.pre < B->base + B->n) ? (o_ref){B, e.pre - B->base} : node_ref(net, e.pre);
        float final_input;
        if (!r.L->is_train) {
            const o_neuron *in = &r.L->cells[r.idx];
            final_input = pn->gap_conductance * (in->current_voltage - pn->current_voltage);
        } else {
            const o_train *s = &r.L->trains[r.idx];
            if (s->last_firing_time < 0) final_input = s->v_resting;
            else final_input = pn->gap_conductance * orc_refractoriness_effect(net->refract_kind, s->k, net->internal_clock,
                                                                              (uint64_t)s->last_firing_time, s->v_th,
                                                                              s->v_resting, s->dt);
        }
        input_val = input_val + final_input * e.w;
    }
    float averager = len == 0 ? 1.f : (float)len;
    return input_val / averager;
}

/* calculate_internal_neurotransmitter_input_from_positions neuron/mod.rs:733-754 /
 * calculate_neurotransmitter_input_from_positions :2169-2212 with weight_neurotransmitter_concentration and
 * aggregate_neurotransmitter_concentrations iterate_and_spike/mod.rs:2837-2866 */
static void chemical_input(orc_network *net, o_lattice *B, uint64_t q, float *t_out, uint8_t *has_out) {
    uint64_t post = B->base + q;
    uint32_t len = net->in ? net->in_len[post] : 0;
    float cum[ORC_NT] = {0.f, 0.f, 0.f};
    uint32_t cnt[ORC_NT] = {0, 0, 0};
    for (uint32_t k = 0; k < len; k++) {
        o_edge e = net->in[post][k];
        o_ref r = (e.pre >= B->base && e.pre < B->base + B->n) ? (o_ref){B, e.pre - B->base} : node_ref(net, e.pre);
        const o_nt *nt = r.L->is_train ? r.L->trains[r.idx].nt : r.L->cells[r.idx].nt;
        for (int ty = 0; ty < ORC_NT; ty++)
            if (nt[ty].present) {
                float value = nt[ty].t * e.w;
                cum[ty] = cum[ty] + value;
                cnt[ty]++;
            }
    }
    for (int ty = 0; ty < ORC_NT; ty++) {
        has_out[ty] = cnt[ty] > 0;
        t_out[ty] = cnt[ty] ? cum[ty] / (float)cnt[ty] : 0.f;
    }
}

void orc_chemical_inputs_dense(const uint32_t *connections, const float *weights, const uint32_t *flags,
                               const float *t, uint32_t n, uint32_t num_types, float *counts, float *res) {
    /* same aggregation over the reference's dense GraphGPU layout [pre*n+post]
     * (gpu_lattices/mod.rs:94-139): res[post*T+ty] = sum_pre w*t_pre[ty] / count, count = #pres having ty */
    for (uint32_t post = 0; post < n; post++)
        for (uint32_t ty = 0; ty < num_types; ty++) {
            float sum = 0.f, count = 0.f;
            for (uint32_t pre = 0; pre < n; pre++)
                if (connections[pre * n + post] == 1 && flags[pre * num_types + ty] == 1) {
                    sum = sum + weights[pre * n + post] * t[pre * num_types + ty];
                    count += 1.f;
                }
            counts[post * num_types + ty] = count;
            res[post * num_types + ty] = count > 0.f ? sum / count : 0.f;
        }
}

static void build_out_index(orc_network *net) {
    if (net->out_valid) return;
    free(net->out_ptr); free(net->out_post); free(net->out_pos);
    uint64_t nn = net->n_nodes, E = 0;
    for (uint64_t p = 0; p < net->n_neurons; p++) E += net->in_len[p];
    net->out_ptr = calloc(nn + 2, sizeof(uint64_t));
    net->out_post = malloc(sizeof(uint32_t) * (E ? E : 1));
    net->out_pos = malloc(sizeof(uint32_t) * (E ? E : 1));
    for (uint64_t p = 0; p < net->n_neurons; p++)
        for (uint32_t k = 0; k < net->in_len[p]; k++) net->out_ptr[net->in[p][k].pre + 2]++;
    for (uint64_t i = 0; i < nn; i++) net->out_ptr[i + 2] += net->out_ptr[i + 1];
    /* posts ascending because p loop ascends */
    for (uint64_t p = 0; p < net->n_neurons; p++)
        for (uint32_t k = 0; k < net->in_len[p]; k++) {
            uint64_t slot = net->out_ptr[net->in[p][k].pre + 1]++;
            net->out_post[slot] = (uint32_t)p; net->out_pos[slot] = k;
        }
    net->out_valid = 1;
}

static int32_t node_lft(orc_network *net, uint64_t node) {
    o_ref r = node_ref(net, node);
    return r.L->is_train ? r.L->trains[r.idx].last_firing_time : r.L->cells[r.idx].last_firing_time;
}

/* Lattice::update_weights_from_neurons neuron/mod.rs:849-881;
 * LatticeNetwork::update_weights_from_neurons_across_lattices :2308-2366 + _within_lattices :2368-2417.
 * In-edges use the spiking neuron's own lattice plasticity; out-edges the TARGET lattice's plasticity. */
static void update_weights_from_neuron(orc_network *net, o_lattice *L, uint64_t q) {
    uint64_t p = L->base + q;
    int32_t lft_p = L->cells[q].last_firing_time;
    for (uint32_t k = 0; k < net->in_len[p]; k++) {
        o_edge *e = &net->in[p][k];
        if (L->use_bcm) { /* single lattice of BCM neurons: the presynaptic node is a neuron of the same lattice */
            o_ref rp = node_ref(net, e->pre);
            e->w = bcm_update(&L->bcm, e->w, &rp.L->cells[rp.idx], &L->cells[q]);
            continue;
        }
        e->w = orc_stdp_update(&L->plasticity, e->w, node_lft(net, e->pre), lft_p);
    }
    for (uint64_t s = net->out_ptr[p]; s < net->out_ptr[p + 1]; s++) {
        uint64_t post = net->out_post[s];
        o_ref r = node_ref(net, post);
        o_edge *e = &net->in[post][net->out_pos[s]];
        if (r.L->use_bcm) { e->w = bcm_update(&r.L->bcm, e->w, &L->cells[q], &r.L->cells[r.idx]); continue; }
        e->w = orc_stdp_update(&r.L->plasticity, e->w, lft_p, r.L->cells[r.idx].last_firing_time);
    }
}

/* RewardModulatedSTDP::update_weight plasticity/mod.rs:197-229 */
static void rstdp_update_weight(const orc_rstdp *m, o_edge *e, int32_t t_pre_i, int32_t t_post_i) {
    float delta_w = 0.f;
    if (t_pre_i >= 0 && t_post_i >= 0) {
        float t_pre = (float)t_pre_i, t_post = (float)t_post_i;
        if (t_pre < t_post) {
            delta_w = m->a_plus * expf((-1.f * fabsf((t_pre - t_post) * m->dt)) / m->tau_plus);
        } else if (t_pre > t_post) {
            delta_w = (-1.f * m->a_minus) * expf((-1.f * fabsf((t_post - t_pre) * m->dt)) / m->tau_minus);
        }
    }
    e->dw += delta_w;
    if (e->counter == 0) {
        e->counter = 1;
    } else {
        e->c = e->c * expf(-m->dt / m->tau_c) + m->tau_c * e->dw; /* TraceRSTDP::update_trace :140-142 */
        e->counter = 0;
        e->dw = 0.f;
    }
    e->w += e->c * m->dopamine;
}

/* RewardModulatedLattice::update_weights_from_neurons neuron/mod.rs:3022-3054: every in-edge, then every out-edge of the
 * neuron that was just stepped (do_update is always true, plasticity/mod.rs:231-233) */
static void rstdp_update_from_neuron(orc_network *net, o_lattice *L, uint64_t q) {
    uint64_t p = L->base + q;
    for (uint32_t k = 0; k < net->in_len[p]; k++) {
        o_edge *e = &net->in[p][k];
        rstdp_update_weight(&net->rstdp, e, node_lft(net, e->pre), L->cells[q].last_firing_time);
    }
    for (uint64_t s = net->out_ptr[p]; s < net->out_ptr[p + 1]; s++) {
        uint64_t post = net->out_post[s];
        o_ref r = node_ref(net, post);
        rstdp_update_weight(&net->rstdp, &net->in[post][net->out_pos[s]], L->cells[q].last_firing_time,
                            r.L->cells[r.idx].last_firing_time);
    }
}

/* RewardModulatedLatticeNetwork::post_neuron_update_step neuron/mod.rs:5030-5062, the half for the network's
 * reward_modulated_lattices: every neuron of a lattice with do_modulation (RewardModulatedSTDP::do_update is always true,
 * plasticity/mod.rs:231-233) runs update_weights_from_neurons_across_reward_lattices (:4855-4977) and
 * _within_reward_lattices (:4979-5028) AFTER all neurons have stepped, so both ends carry this step's last_firing_time.
 *  - an edge of the lattice's own graph is visited as the in-edge of its post end and as the out-edge of its pre end: two
 *    RewardModulatedSTDP::update_weight calls a step with the lattice's modulator;
 *  - a connecting edge INTO the lattice: RewardModulatedWeight -> one call with the post lattice's modulator whatever the
 *    input is (:4885-4925); Weight -> STDP::update_weight with the INPUT lattice's plasticity and only when the input is a
 *    plain Lattice (:4868-4884; a spike-train input is left alone);
 *  - connecting edges OUT of such a lattice are looked up with the end points swapped (:4931-4934) and the reference
 *    panics unless the reverse edge exists; orc_run refuses that configuration (error 90) instead of restating it. */
static void reward_network_update(orc_network *net) {
    for (int li = 0; li < net->n_neuron_lat; li++) {
        o_lattice *B = net->lat[li];
        if (!B->is_reward || !B->do_modulation) continue;
        for (uint64_t q = 0; q < B->n; q++) {
            uint64_t post = B->base + q;
            int32_t t_post = B->cells[q].last_firing_time;
            for (uint32_t k = 0; k < net->in_len[post]; k++) {
                o_edge *e = &net->in[post][k];
                o_lattice *A = lat_of_node(net, e->pre);
                int32_t t_pre = node_lft(net, e->pre);
                if (A == B) {
                    rstdp_update_weight(&B->rstdp, e, t_pre, t_post);
                    rstdp_update_weight(&B->rstdp, e, t_pre, t_post);
                } else if (e->rm) {
                    rstdp_update_weight(&B->rstdp, e, t_pre, t_post);
                } else if (!A->is_train && !A->is_reward) {
                    e->w = orc_stdp_update(&A->plasticity, e->w, t_pre, t_post);
                }
            }
        }
    }
}

/* the configurations of a RewardModulatedLatticeNetwork the reference itself cannot run (see reward_network_update) */
static int reward_network_check(orc_network *net) {
    int any = 0;
    for (int li = 0; li < net->n_neuron_lat; li++) any |= net->lat[li]->is_reward;
    if (!any || !net->in) return 0;
    for (int li = 0; li < net->n_neuron_lat; li++) {
        o_lattice *B = net->lat[li];
        for (uint64_t q = 0; q < B->n; q++) {
            uint64_t post = B->base + q;
            for (uint32_t k = 0; k < net->in_len[post]; k++) {
                const o_edge *e = &net->in[post][k];
                o_lattice *A = lat_of_node(net, e->pre);
                if (A == B) continue;
                if (A->is_reward && A->do_modulation) return 90;            /* :4931-4934 */
                if (!A->is_train && !A->is_reward && A->do_plasticity) return 90; /* :4760-4763 */
                if (!B->is_reward && B->do_plasticity && (e->rm || A->is_reward)) return 90; /* :4727-4731, 4741-4747 */
                if (!B->is_reward && e->rm) return 90; /* a TraceRSTDP nobody advances */
            }
        }
    }
    return 0;
}

static void history_push(o_lattice *L) {
    if (!L->is_train && (L->update_average_history || L->update_eeg_history)) {
        if (L->hist_len >= L->red_cap) {
            L->red_cap = L->red_cap ? L->red_cap * 2 : 64;
            while (L->red_cap <= L->hist_len) L->red_cap *= 2;
            L->average_history = realloc(L->average_history, sizeof(float) * L->red_cap);
            L->eeg_history = realloc(L->eeg_history, sizeof(float) * L->red_cap);
        }
        /* AverageVoltageHistory::update :310-316: voltages.into_iter().sum::<f32>() / length (sequential f32 sum) */
        float sum = 0.f, total_current = 0.f;
        for (uint64_t j = 0; j < L->n; j++) {
            sum += L->cells[j].current_voltage;
            total_current += L->cells[j].current_voltage - L->eeg_reference_voltage; /* EEGHistory::update :266-279 */
        }
        L->average_history[L->hist_len] = sum / (float)L->n;
        L->eeg_history[L->hist_len] = (1.f / (4.f * 3.14159274101257324f * L->eeg_conductivity * L->eeg_distance)) * total_current;
        if (!L->update_grid_history && !L->update_spike_history) { L->hist_len++; return; }
    }
    if (!L->update_grid_history && !L->update_spike_history) return;
    if (L->hist_len == L->hist_cap) {
        L->hist_cap = L->hist_cap ? L->hist_cap * 2 : 64;
        if (L->update_grid_history) L->grid_history = realloc(L->grid_history, sizeof(float) * L->hist_cap * (L->n ? L->n : 1));
        if (L->update_spike_history) L->spike_history = realloc(L->spike_history, L->hist_cap * (L->n ? L->n : 1));
    }
    for (uint64_t j = 0; j < L->n; j++) {
        if (L->update_grid_history)
            L->grid_history[L->hist_len * L->n + j] = L->is_train ? L->trains[j].current_voltage : L->cells[j].current_voltage;
        if (L->update_spike_history)
            L->spike_history[L->hist_len * L->n + j] = (uint8_t)(L->is_train ? L->trains[j].is_spiking : L->cells[j].is_spiking);
    }
    L->hist_len++;
}

/* one timestep: LatticeNetwork::iterate / iterate_with_neurotransmission / iterate_chemical_synapses_only
 * (neuron/mod.rs:2419-2594) preceded by get_all_*_inputs (:2228-2306); for a single lattice this is
 * Lattice::run_lattice_* (neuron/mod.rs:1035-1088) with canonicalisation (3). */
static void step(orc_network *net) {
    int el = net->electrical, ch = net->chemical;
    /* phase 1: inputs from pre-step state (Jacobi) */
    for (int li = 0; li < net->n_neuron_lat; li++) {
        o_lattice *B = net->lat[li];
        long n = (long)B->n;
#pragma omp parallel for schedule(static) if (net->parallel)
        for (long q = 0; q < n; q++) {
            uint64_t post = B->base + (uint64_t)q;
            if (el) net->inp_e[post] = electrical_input(net, B, (uint64_t)q);
            if (ch) chemical_input(net, B, (uint64_t)q, &net->inp_t[post * ORC_NT], &net->inp_has[post * ORC_NT]);
        }
    }
    /* phase 2: neuron update, ascending canonical order */
    if (net->reward_mode && net->do_modulation && net->in) build_out_index(net);
    for (int li = 0; li < net->n_neuron_lat; li++) {
        o_lattice *B = net->lat[li];
        for (uint64_t q = 0; q < B->n; q++) {
            uint64_t post = B->base + q;
            o_neuron *c = &B->cells[q];
            float input = el ? net->inp_e[post] : 0.f; /* chemical-only feeds 0. (neuron/mod.rs:929-931) */
            int sp = neuron_iterate(c, net->model, net->ntk, net->rck, input, ch, &net->inp_t[post * ORC_NT],
                                    &net->inp_has[post * ORC_NT]);
            if (sp) c->last_firing_time = (int32_t)net->internal_clock; /* :964-966 */
            /* RewardModulatedLattice::iterate neuron/mod.rs:3127-3156: the modulator runs inside the node loop, so an edge sees
             * the new last_firing_time of its lower-indexed end and still the old one of the other (canonical order (1)) */
            if (net->reward_mode && net->do_modulation && net->in) rstdp_update_from_neuron(net, B, q);
        }
        history_push(B); /* before this step's STDP (:2565-2576) */
    }
    /* phase 3: deferred STDP for spiking neurons of lattices with do_plasticity (:2559-2562, 2573-2576) */
    int any_plastic = 0;
    for (int li = 0; li < net->n_neuron_lat; li++) if (net->lat[li]->do_plasticity) any_plastic = 1;
    if (any_plastic && net->in && !net->reward_mode) {
        build_out_index(net);
        for (int li = 0; li < net->n_neuron_lat; li++) {
            o_lattice *B = net->lat[li];
            if (!B->do_plasticity) continue;
            for (uint64_t q = 0; q < B->n; q++)
                if (B->cells[q].is_spiking) update_weights_from_neuron(net, B, q); /* STDP::do_update plasticity/mod.rs:67-69 */
        }
    }
    if (net->in) reward_network_update(net);
    net->internal_clock += 1;
    for (int li = 0; li < net->n_neuron_lat; li++) net->lat[li]->internal_clock = net->internal_clock;
    /* phase 4: spike trains step after the neurons with their own clock (:2588-2591, 1377-1393) */
    for (int li = net->n_neuron_lat; li < net->n_lat; li++) {
        o_lattice *S = net->lat[li];
        for (uint64_t j = 0; j < S->n; j++) {
            int sp = train_iterate(net, &S->trains[j]);
            if (sp) S->trains[j].last_firing_time = (int32_t)S->internal_clock;
        }
        history_push(S);
        S->internal_clock += 1;
    }
}

int orc_run(orc_network *net, uint64_t iterations) {
    /* run_lattice / run_lattices dispatch neuron/mod.rs:1213-1218, 2668-2673: (false,false) is a no-op */
    if (!net->electrical && !net->chemical) return 0;
    ensure_graph(net);
    { int r = reward_network_check(net); if (r) return r; }
    if (!net->inp_e) {
        uint64_t n = net->n_neurons ? net->n_neurons : 1;
        net->inp_e = calloc(n, sizeof(float));
        net->inp_t = calloc(n * ORC_NT, sizeof(float));
        net->inp_has = calloc(n * ORC_NT, 1);
    }
    for (uint64_t it = 0; it < iterations; it++) step(net);
    return 0;
}

/* RewardModulatedLattice: switch the lattice's plasticity to RewardModulatedSTDP over TraceRSTDP weights */
int orc_set_reward_modulator(orc_network *net, int enable, int do_modulation, const orc_rstdp *m) {
    net->reward_mode = enable; net->do_modulation = do_modulation;
    if (m) net->rstdp = *m;
    return 0;
}
float orc_get_dopamine(orc_network *net) { return net->rstdp.dopamine; }

/* RewardModulatedLattice::run_lattice_with_reward neuron/mod.rs:3250-3257 (one timestep): inputs, then
 * RewardModulatedSTDP::update(reward) plasticity/mod.rs:193-195, then iterate */
int orc_run_with_reward(orc_network *net, float reward) {
    if (!net->electrical && !net->chemical) return 0;
    net->rstdp.dopamine = net->rstdp.dopamine * expf(-net->rstdp.dt / net->rstdp.tau_d) + net->rstdp.tau_d * reward;
    return orc_run(net, 1);
}

/* RewardModulatedLatticeNetwork::add_reward_modulated_lattice neuron/mod.rs:3615-3634 (do_modulation defaults to true,
 * :2772; RewardModulatedSTDP::default plasticity/mod.rs:167-181) */
int orc_add_reward_lattice(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols) {
    int r = add_lat(net, id, rows, cols, 0);
    if (r) return r;
    o_lattice *L = find_lat(net, id);
    L->is_reward = 1; L->do_modulation = 1;
    L->rstdp = (orc_rstdp){0.f, 20.f, 0.0001f, 2.f, 2.f, 4.5f, 4.5f, 0.1f};
    return 0;
}
int orc_set_lattice_reward_modulator(orc_network *net, uint64_t id, int do_modulation, const orc_rstdp *m) {
    o_lattice *L = find_lat(net, id);
    if (!L || !L->is_reward) return 35;
    L->do_modulation = do_modulation;
    if (m) L->rstdp = *m;
    return 0;
}
int orc_get_lattice_reward_modulator(orc_network *net, uint64_t id, int *do_modulation, orc_rstdp *m) {
    o_lattice *L = find_lat(net, id);
    if (!L || !L->is_reward) return 35;
    if (do_modulation) *do_modulation = L->do_modulation;
    if (m) *m = L->rstdp;
    return 0;
}
/* connect_with_reward_modulation neuron/mod.rs:4076-4209: the block's edges hold RewardModulatedWeight (rm = 1) or
 * Weight (rm = 0) values */
int orc_mark_connection_reward(orc_network *net, uint64_t pre_id, uint64_t post_id, int rm) {
    o_lattice *A = find_lat(net, pre_id), *B = find_lat(net, post_id);
    if (!A || !B || B->is_train) return 35;
    ensure_graph(net);
    for (uint64_t q = 0; q < B->n; q++) {
        uint64_t post = B->base + q;
        for (uint32_t k = 0; k < net->in_len[post]; k++) {
            o_edge *e = &net->in[post][k];
            if (e->pre >= A->base && e->pre < A->base + A->n) e->rm = (uint32_t)(rm != 0);
        }
    }
    return 0;
}
/* RewardModulatedLatticeNetwork::run_lattices_with_reward neuron/mod.rs:5385-5392 -> run_lattices_*(reward) :5280-5297:
 * inputs from the pre-step state, RewardModulatedSTDP::update(reward) (plasticity/mod.rs:193-195) on every
 * reward-modulated lattice, then iterate. The inputs do not read the dopamine, so updating it first is the same. */
int orc_run_network_with_reward(orc_network *net, float reward) {
    if (!net->electrical && !net->chemical) return 0;
    for (int li = 0; li < net->n_neuron_lat; li++) {
        o_lattice *L = net->lat[li];
        if (!L->is_reward) continue;
        L->rstdp.dopamine = L->rstdp.dopamine * expf(-L->rstdp.dt / L->rstdp.tau_d) + L->rstdp.tau_d * reward;
    }
    return orc_run(net, 1);
}

/* per-edge TraceRSTDP members in the order of orc_get_connection_csr */
int orc_get_connection_traces(orc_network *net, uint64_t pre_id, uint64_t post_id, uint32_t *counter, float *dw, float *c) {
    o_lattice *A = find_lat(net, pre_id), *B = find_lat(net, post_id);
    if (!A || !B || B->is_train) return 35;
    ensure_graph(net);
    uint64_t n = 0;
    for (uint64_t q = 0; q < B->n; q++) {
        uint64_t post = B->base + q;
        for (uint32_t k = 0; k < net->in_len[post]; k++) {
            const o_edge *e = &net->in[post][k];
            if (e->pre >= A->base && e->pre < A->base + A->n) {
                if (counter) counter[n] = e->counter;
                if (dw) dw[n] = e->dw;
                if (c) c[n] = e->c;
                n++;
            }
        }
    }
    return 0;
}

/* Graph::edit_weight with whole TraceRSTDP values, order of orc_get_connection_csr; NULL = keep */
int orc_set_connection_traces(orc_network *net, uint64_t pre_id, uint64_t post_id, const float *weight, const uint32_t *counter,
                              const float *dw, const float *c) {
    o_lattice *A = find_lat(net, pre_id), *B = find_lat(net, post_id);
    if (!A || !B || B->is_train) return 35;
    ensure_graph(net);
    uint64_t n = 0;
    for (uint64_t q = 0; q < B->n; q++) {
        uint64_t post = B->base + q;
        for (uint32_t k = 0; k < net->in_len[post]; k++) {
            o_edge *e = &net->in[post][k];
            if (e->pre >= A->base && e->pre < A->base + A->n) {
                if (weight) e->w = weight[n];
                if (counter) e->counter = counter[n];
                if (dw) e->dw = dw[n];
                if (c) e->c = c[n];
                n++;
            }
        }
    }
    return 0;
}

uint64_t orc_history_len(orc_network *net, uint64_t id) { o_lattice *L = find_lat(net, id); return L ? L->hist_len : 0; }

int orc_get_grid_history(orc_network *net, uint64_t id, float *out, uint64_t capacity) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    if (capacity < L->hist_len * L->n) return 67;
    if (L->grid_history) memcpy(out, L->grid_history, sizeof(float) * L->hist_len * L->n);
    return 0;
}

int orc_get_spike_history(orc_network *net, uint64_t id, uint8_t *out, uint64_t capacity) {
    o_lattice *L = find_lat(net, id);
    if (!L) return 35;
    if (capacity < L->hist_len * L->n) return 67;
    if (L->spike_history) memcpy(out, L->spike_history, L->hist_len * L->n);
    return 0;
}

void orc_reset_history(orc_network *net) {
    for (int i = 0; i < net->n_lat; i++) net->lat[i]->hist_len = 0;
}

/* ------------------------------------------------------------------ AdjacencyMatrix doc-test semantics
 * graph/mod.rs:138-297 */
struct orc_adjmat {
    uint32_t *px, *py; int n;
    int *has; float *w; /* n x n, [pre*n+post] */
};

orc_adjmat *orc_adjmat_create(void) { return calloc(1, sizeof(orc_adjmat)); }
void orc_adjmat_destroy(orc_adjmat *g) { if (g) { free(g->px); free(g->py); free(g->has); free(g->w); free(g); } }
static int adj_find(orc_adjmat *g, uint32_t x, uint32_t y) {
    for (int i = 0; i < g->n; i++) if (g->px[i] == x && g->py[i] == y) return i;
    return -1;
}
void orc_adjmat_add_node(orc_adjmat *g, uint32_t x, uint32_t y) {
    if (adj_find(g, x, y) >= 0) return; /* :182-184 */
    int n = g->n, m = n + 1;
    int *has = calloc((size_t)m * m, sizeof(int)); float *w = calloc((size_t)m * m, sizeof(float));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { has[i * m + j] = g->has[i * n + j]; w[i * m + j] = g->w[i * n + j]; }
    free(g->has); free(g->w); g->has = has; g->w = w;
    g->px = realloc(g->px, sizeof(uint32_t) * m); g->py = realloc(g->py, sizeof(uint32_t) * m);
    g->px[n] = x; g->py[n] = y; g->n = m;
}
int orc_adjmat_edit_weight(orc_adjmat *g, uint32_t px, uint32_t py, uint32_t qx, uint32_t qy, int has, float w) {
    int q = adj_find(g, qx, qy); if (q < 0) return 2; /* postsynaptic checked first :209-211 */
    int p = adj_find(g, px, py); if (p < 0) return 1;
    g->has[p * g->n + q] = has; g->w[p * g->n + q] = has ? w : 0.f;
    return 0;
}
int orc_adjmat_lookup_weight(orc_adjmat *g, uint32_t px, uint32_t py, uint32_t qx, uint32_t qy, int *has, float *w) {
    int q = adj_find(g, qx, qy); if (q < 0) return 2;
    int p = adj_find(g, px, py); if (p < 0) return 1;
    *has = g->has[p * g->n + q]; *w = g->w[p * g->n + q];
    return 0;
}
int orc_adjmat_incoming(orc_adjmat *g, uint32_t x, uint32_t y, uint32_t *out, int cap) {
    int q = adj_find(g, x, y); if (q < 0) return -3;
    int c = 0;
    for (int p = 0; p < g->n; p++) if (g->has[p * g->n + q]) { if (c < cap) { out[2 * c] = g->px[p]; out[2 * c + 1] = g->py[p]; } c++; }
    return c;
}
int orc_adjmat_outgoing(orc_adjmat *g, uint32_t x, uint32_t y, uint32_t *out, int cap) {
    int p = adj_find(g, x, y); if (p < 0) return -3;
    int c = 0;
    for (int q = 0; q < g->n; q++) if (g->has[p * g->n + q]) { if (c < cap) { out[2 * c] = g->px[q]; out[2 * c + 1] = g->py[q]; } c++; }
    return c;
}
