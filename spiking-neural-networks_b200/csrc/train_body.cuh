// train_body.cuh — one timestep of 32 spike trains (shared by train_kernel, kernels.cu, and the multi-step kernel, step_multi.cu).
#pragma once
#include "step_body.cuh"

namespace snn {

// ------------------------------------------------------------------------------------------------
// spike trains (SpikeTrainLattice::iterate, neuron/mod.rs:1377-1393)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t *hi) {
    const uint64_t prod = (uint64_t)a * b;
    *hi = (uint32_t)(prod >> 32);
    return (uint32_t)prod;
}

// Philox4x32-10 counter-based generator keyed by (seed), counter (train index, per-handle draw number)
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint32_t idx, uint64_t draw) {
    uint32_t c0 = idx, c1 = (uint32_t)draw, c2 = 0x5EEDu, c3 = (uint32_t)(draw >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, hi1;
        const uint32_t lo0 = mulhilo32(0xD2511F53u, c0, &hi0);
        const uint32_t lo1 = mulhilo32(0xCD9E8D57u, c2, &hi1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// one warp steps 32 spike trains (warp_global * 32 < p.n_trains)
__device__ __forceinline__ void train_step(const TrainParams &p, uint32_t warp_global, uint32_t lane) {
    const uint32_t tn = warp_global * 32u + lane;
    const bool valid = tn < p.n_trains;
    const uint32_t tc = valid ? tn : p.n_trains - 1;
    const uint32_t i = p.train0 + tc;
    int tl = 0;
    for (int k = 1; k < p.n_tl; ++k) if (tc >= p.tl_base[k]) tl = k;
    const float v_th = p.tf[TF_VTH][tc], v_rest = p.tf[TF_VREST][tc], dt = p.tf[TF_DT][tc];
    bool spike = false;
    if (p.kind == SNN_TRAIN_POISSON) {
        // PoissonNeuron::iterate, spike_train/mod.rs:352-368; uniform in [0,1] from Philox (the reference's
        // thread_rng is unseeded, so only the firing statistics are comparable)
        const float u = (float)(philox_u32(p.seed, tc, p.draw) >> 8) * (1.0f / 16777215.0f);
        spike = u <= p.tf[TF_CHANCE][tc];
    } else if (p.kind == SNN_TRAIN_RATE) {
        // RateSpikeTrain::iterate, spike_train/mod.rs:1015-1030
        float step = p.tf[TF_STEP][tc];
        const float rate = p.tf[TF_RATE][tc];
        step += dt;
        if (rate != 0.f && step >= rate) { step = 0.f; spike = true; }
        if (valid) p.tf[TF_STEP][tc] = step;
    } else {
        // PresetSpikeTrain::iterate, spike_train/mod.rs:803-828
        float clk = p.tf[TF_ICLOCK][tc];
        uint32_t counter = __float_as_uint(p.tf[TF_COUNTER][tc]);
        const uint64_t f0 = p.ft_off[tc], f1 = p.ft_off[tc + 1];
        clk += dt;
        if (f1 > f0 && clk > p.ft[f0 + counter]) {
            spike = true;
            clk = 0.f;
            counter += 1;
            if (counter == (uint32_t)(f1 - f0)) counter = 0;
        }
        if (valid) { p.tf[TF_ICLOCK][tc] = clk; p.tf[TF_COUNTER][tc] = __uint_as_float(counter); }
    }
    const float v = spike ? v_th : v_rest;
    // spike trains release with the flag of THIS step (is_spiking is assigned before apply_t_changes)
    const uint32_t ntm = p.node_flags[i] & 0xFu;
    for (int ty = 0; ty < kNT; ++ty) {
        if (!(ntm & (1u << ty))) continue;
        const size_t o = (size_t)ty * p.nt_stride + i;
        const float t_old = p.t_in[(size_t)ty * p.t_stride + i];
        const float t_new = nt_apply(p.ntk, t_old, p.nt[NTF_TMAX][o], p.nt[NTF_P1][o], p.nt[NTF_P2][o], v, spike, dt);
        if (valid) p.t_out[(size_t)ty * p.t_stride + i] = t_new;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, spike && valid);
    if (lane == 0) {
        p.spk_out[(p.train0 >> 5) + warp_global] = word;
        // per-lattice spike history is assembled on the host from the full train raster
        if (p.spike_hist) p.spike_hist[warp_global] = word;
    }
    if (valid) {
        p.v_out[i] = v;
        const int lft_old = p.lft_in[i];
        if (p.lft_pp) p.lft_out[i] = spike ? (int)p.tl_clock[tl] : lft_old;
        else if (spike) p.lft_out[i] = (int)p.tl_clock[tl];
        if (p.grid_hist) p.grid_hist[tc] = v;
    }
}


}  // namespace snn
