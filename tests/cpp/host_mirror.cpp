// host_mirror.cpp — exercises include/snn_b200.hpp (the C++ mirror of the reference's Rust API above the C ABI) the way the
// reference's own integration tests drive `Lattice` / `LatticeGPU`:
//   backend/tests/gpu_connection_behavior.rs:51-95  (3x3 QIF lattice, connect(x != y, weight 2), graph look-ups, 1000 steps)
//   backend/tests/rate_spike_train_lattices.rs:62-90 (a network holding only a spike-train lattice still steps it)
//   backend/src/neuron/mod.rs:1852-1862              (LatticeNetwork::connect error order)
// Built by tests/test_cpp_host_mirror.py; prints HOST_MIRROR_OK on success.  With --no-gpu it only checks what works without a
// device (error behaviour of create()).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "snn_b200.hpp"

using namespace snn_b200;

#define REQUIRE(cond)                                                             \
    do {                                                                          \
        if (!(cond)) { std::printf("REQUIRE failed: %s (line %d)\n", #cond, __LINE__); return 1; } \
    } while (0)

static int without_device() {
    int32_t n = 0;
    if (snn_device_count(&n) == SNN_OK && n > 0) return 0;   // a device is present: nothing to check here
    Lattice lattice;
    try {
        lattice.populate(BaseNeuron(SNN_MODEL_IZHIKEVICH), 2, 2);
    } catch (const SpikingNeuralNetworksError &e) {
        // no CPU fallback: GPUError::GetDeviceFailure (error/mod.rs:221-238)
        REQUIRE(e.status == SNN_GPU_GET_DEVICE_FAILURE);
        std::printf("HOST_MIRROR_OK (no device: %s)\n", e.what());
        return 0;
    }
    std::printf("populate() succeeded without a device\n");
    return 1;
}

int main(int argc, char **argv) {
    if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) return without_device();

    // ---- gpu_connection_behavior.rs:51-95 --------------------------------------------------------------------------------
    const BaseNeuron base_neuron = BaseNeuron(SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE).with("gap_conductance", 10.f);
    Lattice lattice;
    lattice.populate(base_neuron, 3, 3);
    lattice.update_grid_history = true;
    lattice.connect([](Position x, Position y) { return x != y; }, [](Position, Position) { return 2.f; });
    lattice.electrical_synapse = true;
    lattice.chemical_synapse = false;
    for (std::size_t a = 0; a < 9; ++a)
        for (std::size_t b = 0; b < 9; ++b) {
            float w = 0.f;
            const bool some = lattice.lookup_weight({a / 3, a % 3}, {b / 3, b % 3}, &w);
            REQUIRE(some == (a != b));
            if (some) REQUIRE(w == 2.f);
        }
    // identical neurons, all-to-all: zero gap current, so every cell follows the isolated trajectory
    Lattice isolated;
    isolated.populate(base_neuron, 1, 1);
    isolated.update_grid_history = true;
    lattice.run_lattice(1000);
    isolated.run_lattice(1000);
    REQUIRE(lattice.internal_clock() == 1000 && lattice.history_len() == 1000);
    const std::vector<float> hist = lattice.grid_history(), iso = isolated.grid_history();
    REQUIRE(hist.size() == 9000 && iso.size() == 1000);
    for (std::size_t s = 0; s < 1000; ++s)
        for (std::size_t c = 0; c < 9; ++c) REQUIRE(hist[s * 9 + c] == iso[s]);

    // ---- heterogeneous Izhikevich lattice: the grid generator equals the predicate it stands for -----------------------------
    Lattice a, b;
    for (Lattice *l : {&a, &b}) {
        l->populate(BaseNeuron(SNN_MODEL_IZHIKEVICH).with("gap_conductance", 10.f).with("c_m", 2.f), 6, 7);
        l->apply_given_position("current_voltage", [](Position p, float) { return -65.f + 9.f * (float)p.first + 5.f * (float)p.second; });
        l->apply_given_position("b", [](Position p, float) { return 0.25f + 0.01f * (float)((p.first + 2 * p.second) % 10); });
        l->update_grid_history = l->update_spike_history = true;
        l->do_plasticity = true;
        l->plasticity.a_plus = l->plasticity.a_minus = 0.02f;
    }
    a.connect_grid(1, 1.f);
    b.connect([](Position x, Position y) {
        const long dr = (long)x.first - (long)y.first, dc = (long)x.second - (long)y.second;
        return x != y && std::labs(dr) <= 1 && std::labs(dc) <= 1;
    });
    a.run_lattice(300);
    b.run_lattice(300);
    REQUIRE(a.grid_history() == b.grid_history() && a.spike_history() == b.spike_history());
    std::size_t spikes = 0;
    for (uint8_t s : a.spike_history()) spikes += s;
    REQUIRE(spikes > 10);
    REQUIRE(a.get_last_firing_times() == b.get_last_firing_times());
    a.reset_timing();
    REQUIRE(a.internal_clock() == 0);
    for (int32_t t : a.get_last_firing_times()) REQUIRE(t == -1);

    // ---- RewardModulatedLattice: run_lattice_with_reward updates the dopamine level and moves the weights -----------------
    RewardModulatedLattice r;
    r.populate(BaseNeuron(SNN_MODEL_IZHIKEVICH).with("gap_conductance", 10.f).with("c_m", 2.f).with("b", 0.3f), 5, 5);
    r.apply_given_position("current_voltage", [](Position p, float) { return -60.f + 7.f * (float)p.first + 11.f * (float)p.second; });
    r.connect_grid(1, 1.f);
    r.reward_modulator.tau_c = 0.05f;
    for (int s = 0; s < 80; ++s) r.run_lattice_with_reward(s % 3 == 0 ? 0.5f : -0.1f);
    REQUIRE(r.internal_clock() == 80 && r.reward_modulator.dopamine != 0.f);
    float w01 = 0.f;
    REQUIRE(r.lookup_weight({0, 0}, {0, 1}, &w01));

    // ---- LatticeNetwork: error order of connect, spike trains alone still step (rate_spike_train_lattices.rs:62-90) -------
    LatticeNetwork net(SNN_MODEL_IZHIKEVICH, SNN_TRAIN_RATE);
    net.add_spike_train_lattice(0, {{"rate", 10.f}, {"dt", 1.f}}, 2, 2);
    net.set_lattice_option(0, SNN_OPT_UPDATE_SPIKE_HISTORY, 1);
    net.run_lattices(100);
    std::size_t train_spikes = 0;
    for (uint8_t s : net.spike_history(0)) train_spikes += s;
    REQUIRE(train_spikes == 4 * 10);   // one spike every 10 steps per train
    net.add_lattice(1, BaseNeuron(SNN_MODEL_IZHIKEVICH), 2, 2);
    try {
        net.add_lattice(1, BaseNeuron(SNN_MODEL_IZHIKEVICH), 3, 3);
        REQUIRE(false);
    } catch (const SpikingNeuralNetworksError &e) { REQUIRE(e.status == SNN_NET_GRAPH_ID_ALREADY_PRESENT); }
    try {
        net.connect(1, 0, [](Position, Position) { return true; });
        REQUIRE(false);
    } catch (const SpikingNeuralNetworksError &e) { REQUIRE(e.status == SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN); }
    try {
        net.connect(7, 1, [](Position, Position) { return true; });
        REQUIRE(false);
    } catch (const SpikingNeuralNetworksError &e) { REQUIRE(e.status == SNN_NET_PRESYNAPTIC_ID_NOT_FOUND); }
    net.connect(0, 1, [](Position, Position) { return true; }, [](Position, Position) { return 0.5f; });
    net.run_lattices(50);

    // ---- RewardModulatedLatticeNetwork in the shape of examples/lsm_architecture: trains -> liquid -> reward-modulated read-out ----
    RewardModulatedLatticeNetwork lsm(SNN_MODEL_IZHIKEVICH, SNN_TRAIN_RATE);
    const BaseNeuron cell = BaseNeuron(SNN_MODEL_IZHIKEVICH).with("gap_conductance", 10.f).with("c_m", 2.f).with("b", 0.3f);
    lsm.add_spike_train_lattice(0, {{"rate", 7.f}, {"dt", 1.f}}, 2, 3);
    lsm.add_lattice(1, cell, 4, 4);
    RewardModulatedSTDP mod;
    mod.tau_c = 0.1f; mod.a_plus = 0.002f; mod.a_minus = 0.002f;
    lsm.add_reward_modulated_lattice(2, cell, 3, 3, mod);
    const auto always = [](Position, Position) { return true; };
    const auto apart = [](Position x, Position y) { return x != y; };
    lsm.connect(1, 1, apart, [](Position, Position) { return 0.4f; });
    lsm.connect_reward_modulated_lattice_interally(2, apart, [](Position, Position) { return 0.3f; });
    lsm.connect(0, 1, always, [](Position, Position) { return 0.8f; });
    lsm.connect_with_reward_modulation(0, 2, always, [](Position, Position) { return RewardModulatedConnection::RewardModulatedWeight(0.6f); });
    lsm.connect_with_reward_modulation(1, 2, always, [](Position, Position) { return RewardModulatedConnection::Weight(0.5f); });
    const struct { std::size_t pre, post; bool with_reward; int status; } refused[] = {
        {1, 2, false, SNN_NET_CONNECT_FUNCTION_MUST_HAVE_NON_REWARD_MODULATED_LATTICE},
        {2, 1, false, SNN_NET_CONNECT_FUNCTION_MUST_HAVE_NON_REWARD_MODULATED_LATTICE},
        {0, 1, true, SNN_NET_CANNOT_CONNECT_WITH_REWARD_MODULATED_CONNECTION},
        {2, 2, true, SNN_NET_REWARD_MODULATED_CONNECTION_NOT_COMPATIBLE_INTERNALLY},
        {2, 0, true, SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN},
        {9, 2, true, SNN_NET_PRESYNAPTIC_ID_NOT_FOUND}};
    for (const auto &c : refused) {
        try {
            if (c.with_reward) lsm.connect_with_reward_modulation(c.pre, c.post, always, [](Position, Position) { return RewardModulatedConnection::RewardModulatedWeight(1.f); });
            else lsm.connect(c.pre, c.post, always);
            REQUIRE(false);
        } catch (const SpikingNeuralNetworksError &e) { REQUIRE(e.status == c.status); }
    }
    const std::vector<float> w02_before = lsm.connection_weights(0, 2), w22_before = lsm.connection_weights(2, 2);
    for (int s = 0; s < 200; ++s) lsm.run_lattices_with_reward(s % 4 == 0 ? 0.004f : -0.001f);
    lsm.run_lattices(20);
    const std::vector<float> w02 = lsm.connection_weights(0, 2), w22 = lsm.connection_weights(2, 2), w01n = lsm.connection_weights(0, 1);
    REQUIRE(w02.size() == 6 * 9 && w22.size() == 9 * 8);
    std::size_t moved = 0;
    for (std::size_t k = 0; k < w02.size(); ++k) moved += w02[k] != w02_before[k];
    for (std::size_t k = 0; k < w22.size(); ++k) moved += w22[k] != w22_before[k];
    REQUIRE(moved > 20 && lsm.reward_modulator(2).dopamine != 0.f);
    for (float w : w01n) REQUIRE(w == 0.8f);   // plain lattice without do_plasticity: its in-edges stay
    lsm.connect_with_reward_modulation(2, 1, [](Position x, Position y) { return x == y; }, [](Position, Position) { return RewardModulatedConnection::Weight(0.2f); });
    try {
        lsm.run_lattices(1);   // connecting edges out of a reward-modulated lattice with do_modulation: the reference panics
        REQUIRE(false);
    } catch (const SpikingNeuralNetworksError &e) { REQUIRE(e.status == SNN_UNSUPPORTED); }

    // ---- SpikeTrainLattice on its own: tests/rate_spike_train.rs:54-72 (rate 100, dt 1: a spike every 100th step) -----------
    SpikeTrainLattice trains(SNN_TRAIN_RATE, 4);
    trains.populate({{"rate", 100.f}, {"dt", 1.f}}, 2, 3);
    trains.update_spike_history = true;
    trains.run_lattice(1001);
    const std::vector<uint8_t> ts = trains.spike_history();
    REQUIRE(ts.size() == 1001 * 6);
    for (std::size_t s = 0; s < 1001; ++s)
        for (std::size_t c = 0; c < 6; ++c) REQUIRE((ts[s * 6 + c] != 0) == (s != 0 && (s + 1) % 100 == 0));

    std::printf("HOST_MIRROR_OK (%zu lattice spikes, dopamine %.4f, w(0,0)->(0,1) %.5f)\n", spikes, r.reward_modulator.dopamine, w01);
    return 0;
}
