/*
 * snn_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the reference's CPU algorithm for the lattice stepping path
 * (NikhilMukraj/spiking-neural-networks, backend/src/neuron/...).  It exists so that the CUDA
 * path can be checked against it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libsnn_b200.so) never links,
 * loads or calls anything in this directory.
 *
 * PARITY PIN STATUS: the Rust reference cannot be built in this image (no rustc/cargo, crates
 * not vendored), so the oracle is pinned against the reference's own known-answer vectors and
 * behavioural tests (tests/test_oracle_kats.py lists each with its file:line), plus an
 * independent numpy float32 restatement (oracle/numpy_ref.py).  Poisson spike trains are
 * "parity unpinned" at the RNG boundary: the reference draws from rand 0.8.5 thread_rng
 * (ChaCha12, OS-seeded; spike_train/mod.rs:354) and fixes no seed anywhere; only rate statistics
 * can be compared.
 *
 * Canonicalisation where the reference is nondeterministic (HashSet iteration order):
 *  (1) nodes are visited in ascending canonical index: lattices by ascending id, row-major inside
 *      a lattice, then spike-train lattices by ascending id;
 *  (2) in-edge sums accumulate in ascending canonical presynaptic index, plain f32 adds;
 *  (3) STDP is applied after all neurons of the step have been updated (LatticeNetwork::iterate,
 *      neuron/mod.rs:2573-2576).  Lattice::iterate (neuron/mod.rs:954-982) applies it inside the
 *      node loop, which differs only when two connected neurons spike in the same step — a case
 *      in which the reference's own result depends on hash order;
 *  (4) RewardModulatedLattice::iterate (neuron/mod.rs:3127-3156) runs the reward modulator inside the node loop and
 *      do_update is always true, so the order matters for every edge in every step: kept inside the loop, in the
 *      ascending node order of (1) (an edge first sees the new last_firing_time of its lower-indexed end only).
 */
#ifndef SNN_ORACLE_H
#define SNN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NT 3 /* AMPA, NMDA, GABA : iterate_and_spike/mod.rs:1068-1073, 1322-1330 */

enum { ORC_LIF = 0, ORC_QIF = 1, ORC_ADLIF = 2, ORC_ADEX = 3, ORC_IZH = 4, ORC_LEAKY_IZH = 5, ORC_SIMPLE_LIF = 6, ORC_HH = 7, ORC_BCM_IZH = 8 };
enum { ORC_NTK_APPROX = 0, ORC_NTK_DESTEXHE = 1, ORC_NTK_DISCRETE = 2, ORC_NTK_EXPDECAY = 3 };
enum { ORC_RCK_APPROX = 0, ORC_RCK_DESTEXHE = 1, ORC_RCK_EXPDECAY = 2 };
enum { ORC_TRAIN_POISSON = 0, ORC_TRAIN_RATE = 1, ORC_TRAIN_PRESET = 2 };
enum { ORC_REFRACT_DELTA_DIRAC = 0, ORC_REFRACT_EXP_DECAY = 1 };
enum { ORC_F32 = 0, ORC_U32 = 1, ORC_I32 = 2 };

typedef struct orc_network orc_network;

typedef struct orc_stdp { float a_plus, a_minus, tau_plus, tau_minus, dt; } orc_stdp;
/* RewardModulatedSTDP plasticity/mod.rs:155-189 (defaults 0, 20, 0.0001, 2, 2, 4.5, 4.5, 0.1) */
/* BCM plasticity/mod.rs:80-97 (defaults 0.1, 0.1, 0.1) */
typedef struct orc_bcm { float decay, average_scalar, dt; } orc_bcm;
typedef struct orc_rstdp { float dopamine, tau_d, tau_c, a_plus, a_minus, tau_plus, tau_minus, dt; } orc_rstdp;

orc_network *orc_network_create(int model, int nt_kinetics, int rc_kinetics, int train_kind, int refract_kind);
void orc_network_destroy(orc_network *net);

/* returns 0 on success, nonzero on error (duplicate id etc.) */
int orc_add_lattice(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols);
int orc_add_train_lattice(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols);
uint64_t orc_lattice_size(orc_network *net, uint64_t id);

/* named SoA views over the AoS cell grids; names identical to the product's field names */
int orc_set_field(orc_network *net, uint64_t id, const char *name, const void *data, uint64_t count, int dtype);
int orc_get_field(orc_network *net, uint64_t id, const char *name, void *out, uint64_t count, int dtype);
int orc_fill_field_f32(orc_network *net, uint64_t id, const char *name, float v);
int orc_fill_field_u32(orc_network *net, uint64_t id, const char *name, uint32_t v);
int orc_fill_field_i32(orc_network *net, uint64_t id, const char *name, int32_t v);
int orc_set_preset_firing_times(orc_network *net, uint64_t id, const uint64_t *offsets, const float *times,
                                uint64_t n_trains, uint64_t n_times);

/* connect(pre_id, post_id): dense [pre*n_post+post] or CSR by post (pre = flat index in pre lattice) */
int orc_connect_dense(orc_network *net, uint64_t pre_id, uint64_t post_id, const uint32_t *connections,
                      const float *weights, uint64_t n_pre, uint64_t n_post);
int orc_connect_csr(orc_network *net, uint64_t pre_id, uint64_t post_id, const uint64_t *row_ptr,
                    const uint32_t *pre, const float *weights, uint64_t n_post, uint64_t nnz);
/* Moore-neighbourhood internal graph: max(|dr|,|dc|) <= radius && x != y, constant weight */
int orc_connect_grid(orc_network *net, uint64_t id, uint32_t radius, float weight);
int orc_get_connection_dense(orc_network *net, uint64_t pre_id, uint64_t post_id, uint32_t *connections,
                             float *weights, uint64_t n_pre, uint64_t n_post);
uint64_t orc_connection_nnz(orc_network *net, uint64_t pre_id, uint64_t post_id);
/* CSR of a block in canonical order (pre ascending per post row) */
int orc_get_connection_csr(orc_network *net, uint64_t pre_id, uint64_t post_id, uint64_t *row_ptr,
                           uint32_t *pre, float *weights);

void orc_set_synapses(orc_network *net, int electrical, int chemical);
void orc_set_parallel(orc_network *net, int parallel); /* OpenMP over the input-gather phase only */
void orc_set_clock(orc_network *net, uint64_t clock);
uint64_t orc_get_clock(orc_network *net);
int orc_set_lattice_flags(orc_network *net, uint64_t id, int do_plasticity, int update_grid_history,
                          int update_spike_history);
int orc_set_plasticity(orc_network *net, uint64_t id, const orc_stdp *stdp);
void orc_set_dt(orc_network *net, float dt);
void orc_reset_timing(orc_network *net);
void orc_seed(orc_network *net, uint64_t seed); /* oracle-local xorshift for Poisson; statistics only */

/* RunNetwork::run_lattices / RunLattice::run_lattice */
int orc_run(orc_network *net, uint64_t iterations);

/* AverageVoltageHistory / EEGHistory (neuron/mod.rs:231-322): one f32 per step, sequential f32 sums like the reference */
int orc_set_reduced_history(orc_network *net, uint64_t id, int average, int eeg, float reference_voltage, float distance,
                            float conductivity);
int orc_get_reduced_history(orc_network *net, uint64_t id, int eeg, float *out, uint64_t capacity);
/* RewardModulatedLattice (neuron/mod.rs:2717-3416) with RewardModulatedSTDP over TraceRSTDP weights (plasticity/mod.rs:114-234) */
/* the lattice's plasticity rule becomes BCM (needs the BCM Izhikevich model); do_plasticity keeps switching it */
int orc_set_bcm_plasticity(orc_network *net, uint64_t id, int enable, const orc_bcm *bcm);
int orc_set_reward_modulator(orc_network *net, int enable, int do_modulation, const orc_rstdp *m);
float orc_get_dopamine(orc_network *net);
int orc_run_with_reward(orc_network *net, float reward);
/* RewardModulatedLatticeNetwork neuron/mod.rs:3455-5455: reward-modulated lattices inside a network */
int orc_add_reward_lattice(orc_network *net, uint64_t id, uint32_t rows, uint32_t cols);
int orc_set_lattice_reward_modulator(orc_network *net, uint64_t id, int do_modulation, const orc_rstdp *m);
int orc_get_lattice_reward_modulator(orc_network *net, uint64_t id, int *do_modulation, orc_rstdp *m);
int orc_mark_connection_reward(orc_network *net, uint64_t pre_id, uint64_t post_id, int rm);
int orc_run_network_with_reward(orc_network *net, float reward);
int orc_get_connection_traces(orc_network *net, uint64_t pre_id, uint64_t post_id, uint32_t *counter, float *dw, float *c);
int orc_set_connection_traces(orc_network *net, uint64_t pre_id, uint64_t post_id, const float *weight, const uint32_t *counter,
                              const float *dw, const float *c);
uint64_t orc_history_len(orc_network *net, uint64_t id);
int orc_get_grid_history(orc_network *net, uint64_t id, float *out, uint64_t capacity);
int orc_get_spike_history(orc_network *net, uint64_t id, uint8_t *out, uint64_t capacity);
void orc_reset_history(orc_network *net);

/* ---- stand-alone pieces pinned by the reference's KATs ---- */
/* chemical input aggregation for one postsynaptic node over an n-node dense graph with per-node
 * type flags, the arithmetic of get_neurotransmitter_inputs / calculate_network_*_chemical_inputs
 * (gpu_lattices/mod.rs:94-139, 1325-1384; CPU: iterate_and_spike/mod.rs:2837-2866) */
void orc_chemical_inputs_dense(const uint32_t *connections, const float *weights, const uint32_t *flags,
                               const float *t, uint32_t n, uint32_t num_types, float *counts, float *res);
/* STDP::update_weight, plasticity/mod.rs:46-65 (t = -1 for None) */
float orc_stdp_update(const orc_stdp *p, float weight, int32_t t_pre, int32_t t_post);
/* spike_train_gap_junction effect, neuron/mod.rs:119-137 + spike_train/mod.rs:84-88,174-176 */
float orc_refractoriness_effect(int kind, float k, uint64_t timestep, uint64_t last_firing_time, float v_max,
                                float v_resting, float dt);
/* PoissonNeuron::from_firing_rate, spike_train/mod.rs:330-337 */
float orc_chance_from_firing_rate(float hertz, float dt);

/* AdjacencyMatrix doc-test semantics (graph/mod.rs:112-137): tiny position-keyed dense graph */
typedef struct orc_adjmat orc_adjmat;
orc_adjmat *orc_adjmat_create(void);
void orc_adjmat_destroy(orc_adjmat *g);
void orc_adjmat_add_node(orc_adjmat *g, uint32_t x, uint32_t y);
/* return 0 ok, 1 presynaptic not found, 2 postsynaptic not found, 3 position not found */
int orc_adjmat_edit_weight(orc_adjmat *g, uint32_t px, uint32_t py, uint32_t qx, uint32_t qy, int has, float w);
int orc_adjmat_lookup_weight(orc_adjmat *g, uint32_t px, uint32_t py, uint32_t qx, uint32_t qy, int *has, float *w);
/* writes up to cap (x,y) pairs into out, returns count or -3 */
int orc_adjmat_incoming(orc_adjmat *g, uint32_t x, uint32_t y, uint32_t *out, int cap);
int orc_adjmat_outgoing(orc_adjmat *g, uint32_t x, uint32_t y, uint32_t *out, int cap);

#ifdef __cplusplus
}
#endif
#endif
