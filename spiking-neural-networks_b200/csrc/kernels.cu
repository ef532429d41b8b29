// kernels.cu — hand-written sm_100a kernels of the lattice stepping path.
//
// One fused kernel per timestep replaces the reference's per-step sequence
//   get_internal_electrical_inputs / get_internal_neurotransmitter_inputs   (neuron/mod.rs:757-826)
//   iterate / iterate_with_neurotransmission / iterate_chemical_synapses_only (neuron/mod.rs:884-982)
//   update_weights_from_neurons (STDP)                                      (neuron/mod.rs:849-881)
// and the OpenCL twin's four launches per step (gpu_lattices/mod.rs:809-881).
//
// Compiled with -fmad=false: Rust never contracts a*b+c, and the deterministic models
// (Izhikevich, LIF, QIF, ...) must reproduce the oracle's spike raster bit for bit.  Every
// expression below is written in the reference's evaluation order.
//
// Layout: structure-of-arrays in HBM, one thread per postsynaptic neuron, one warp per 32-row
// slice of the sliced-ELL in-edge table so that every per-edge load (col, weight) is one coalesced
// 128-byte line per warp; neighbour V / t / last_firing_time are gathered through L1/L2.
// The path is HBM-bound elementwise + sparse gather: no tensor cores.
#include "common.h"

#include <cfloat>

namespace snn {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lat_index(const StepParams &p, uint32_t ln) {
    if (p.n_lat <= 1) return 0;
    int l = 0;
#pragma unroll 1
    for (int k = 1; k < p.n_lat; ++k)
        if (ln >= p.lat[k].base) l = k;
    return l;
}

// STDP::update_weight, plasticity/mod.rs:46-65
__device__ __forceinline__ float stdp_delta(const LatInfo &L, int t_pre_i, int t_post_i) {
    float delta_w = 0.f;
    if (t_pre_i >= 0 && t_post_i >= 0) {
        const float t_pre = (float)t_pre_i, t_post = (float)t_post_i;
        if (t_pre < t_post) {
            delta_w = L.a_plus * expf((-1.f * fabsf((t_pre - t_post) * L.dt)) / L.tau_plus);
        } else if (t_pre > t_post) {
            delta_w = (-1.f * L.a_minus) * expf((-1.f * fabsf((t_post - t_pre) * L.dt)) / L.tau_minus);
        }
    }
    return delta_w;
}

// NeurotransmitterKinetics::apply_t_change, iterate_and_spike/mod.rs:147-150, 192-196, 301-304, 352-357
__device__ __forceinline__ float nt_apply(int kind, float t, float t_max, float p1, float p2, float voltage,
                                          bool is_spiking, float dt) {
    const float flag = is_spiking ? 1.f : 0.f;
    switch (kind) {
    case SNN_NT_APPROXIMATE: {
        const float a = dt * -p1;
        const float b = a * t;
        const float c = flag * t_max;
        t = t + (b + c);
        return fminf(t_max, fmaxf(t, 0.f));
    }
    case SNN_NT_DESTEXHE:
        return t_max / (1.f + expf(-(voltage - p1) / p2));
    case SNN_NT_DISCRETE_SPIKE:
        return t_max * flag;
    default: {  // SNN_NT_EXPONENTIAL_DECAY
        const float t_change = -t * expf(dt / -p1);
        t = t + (t_change + flag * t_max);
        return fminf(t_max, fmaxf(t, 0.f));
    }
    }
}

// ReceptorKinetics::apply_r_change, iterate_and_spike/mod.rs:403-406, 434-437, 510-514
__device__ __forceinline__ float rc_apply(int kind, float r, float k1, float k2, float t, float dt) {
    switch (kind) {
    case SNN_RC_APPROXIMATE:
        return t;
    case SNN_RC_DESTEXHE: {
        const float a = k1 * t;
        const float b = a * (1.f - r);
        const float c = k2 * r;
        return r + (b - c) * dt;
    }
    default: {  // SNN_RC_EXPONENTIAL_DECAY: k1 = r_max, k2 = decay_constant
        const float dec = -r * expf(dt / -k2);
        r = r + (dec + t);
        return fminf(k1, fmaxf(r, 0.f));
    }
    }
}

// AMPA / NMDA / GABA ::iterate, iterate_and_spike/mod.rs:1101-1103, 1132-1137, 1164-1166
__device__ __forceinline__ float receptor_current(int type, float g, float e, float mg, float r, float v) {
    if (type == SNN_NT_NMDA) {
        const float ex = expf(-0.062f * v);
        const float den = 1.0f + ((ex * mg) / 3.75f);
        return (((1.0f / den) * g) * r) * (v - e);
    }
    return (g * r) * (v - e);
}

// NeuralRefractoriness::get_effect, spike_train/mod.rs:68-73, 84-86, 174-176
__device__ __forceinline__ float refract_effect(int kind, float k, uint32_t timestep, uint32_t last, float v_max,
                                                float v_resting, float dt) {
    const float a = v_max - v_resting;
    const float td = (float)(timestep - last);
    if (kind == SNN_REFRACT_DELTA_DIRAC) return a * expf((-1.f / (k / dt)) * (td * td)) + v_resting;
    return a * expf((-1.f / (k / dt)) * td) + v_resting;
}

__device__ __forceinline__ float ldf(const float *p, size_t i) { return __ldg(p + i); }

// ------------------------------------------------------------------------------------------------
// halo synchronisation over NVLink peer memory (multi-GPU row strips)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Spin until the neighbour's arrival counter reaches `epoch`.  Bounded (~4e9 SM cycles, about two seconds): a
// neighbour that died must not hang this GPU; the host turns the error flag into SNN_GPU_WAIT_ERROR.
__device__ __forceinline__ void halo_wait(const unsigned long long *flag, unsigned long long epoch, unsigned int *err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000ll) { atomicExch(err, 1u); break; }
    }
}

// ------------------------------------------------------------------------------------------------
// in-edge gather
// ------------------------------------------------------------------------------------------------
struct EdgeAcc {
    float acc_e;
    uint32_t n_in;
    float acc_t[kNT];
    uint32_t cnt[kNT];
};

// CHEMG: 0 = no chemical gather, 1 = exactly one neurotransmitter type in the whole node array (type index ty0),
//        3 = general per-edge type masks.  NET: the node array holds several lattices and/or spike trains.
//
// Edges are consumed in chunks of 8 (slice widths are padded to a multiple of 4): first all coalesced col/weight
// loads of the chunk, then all neighbour gathers, then the strictly ordered accumulation.  Eight independent
// requests per thread keep enough bytes in flight to cover HBM latency; the arithmetic order is the canonical one
// (ascending presynaptic index).  Padding slots carry weight 0 and gather the neuron itself, so they add an exact
// +0 and need no branch.
template <int CHEMG, bool STDP, bool NET>
__device__ __forceinline__ void gather_edges(const StepParams &p, uint32_t warp_global, uint32_t lane, uint32_t i, float v,
                                             float gap, int lft_me, bool post_trig, int li, uint32_t ty0, EdgeAcc &A) {
    constexpr int U = 8;
    // uniform slice width (stencil graphs): no dependent load for the slice bounds
    const uint32_t k0 = p.uniform_width ? warp_global * p.uniform_width : __ldg(p.slice_off + warp_global);
    const uint32_t k1 = p.uniform_width ? k0 + p.uniform_width : __ldg(p.slice_off + warp_global + 1);
    const bool pending = STDP && p.apply_pending;
    const bool do_e = p.electrical != 0;
    const int prev = (int)p.clock - 1;
    const float *t0 = (CHEMG == 1) ? p.t_in + (size_t)ty0 * p.t_stride : nullptr;
    for (uint32_t k = k0; k < k1; k += U) {
        const bool full = k + U <= k1;
        const uint32_t *cp = p.col + (size_t)k * 32u + lane;
        float *wp = p.wgt + (size_t)k * 32u + lane;
        uint32_t c[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u < 4 || full) { c[u] = __ldg(cp + u * 32); w[u] = wp[u * 32]; }
            else { c[u] = kColPad; w[u] = 0.f; }
        }
        uint32_t j[U];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = (c[u] == kColPad) ? i : (c[u] & kColIdxMask);
        float vj[U];
        if (do_e) {
#pragma unroll
            for (int u = 0; u < U; ++u) vj[u] = p.v_in[j[u]];
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) vj[u] = v;
        }
        int lj[U];
        if (pending) {
            // a spike train steps AFTER the neurons and their STDP (neuron/mod.rs:2573-2591): the rule of step s must
            // see the train's last_firing_time from before its step-s iterate, which still sits in the other
            // ping-pong buffer (the train kernel of this step has not run yet)
#pragma unroll
            for (int u = 0; u < U; ++u) lj[u] = (NET && (c[u] & kColTrainBit) && c[u] != kColPad) ? p.lft_out[j[u]] : p.lft_in[j[u]];
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) lj[u] = -1;
        }
        float tj[U][CHEMG == 3 ? kNT : 1];
        if (CHEMG == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) tj[u][0] = t0[j[u]];
        } else if (CHEMG == 3) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t m = (c[u] == kColPad) ? 0u : (c[u] >> kColNtShift);
#pragma unroll
                for (int ty = 0; ty < kNT; ++ty) tj[u][ty] = (m & (1u << ty)) ? p.t_in[(size_t)ty * p.t_stride + j[u]] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = c[u] != kColPad;
            float wu = w[u];
            if (pending) {
                // lazy application of the previous step's STDP while the edge streams by: in-edge rule if the post
                // neuron spiked last step, out-edge rule if the pre neuron did (update_weights_from_neurons,
                // neuron/mod.rs:849-881, 2308-2417); both use the post lattice's rule.  No edge can get two non-zero
                // updates in one step.
                bool pre_trig = ok && lj[u] == prev;
                if (NET) {
                    if (pre_trig) pre_trig = !(c[u] & kColTrainBit) && p.lat[lat_index(p, j[u] - p.own0)].do_plasticity != 0;
                } else {
                    pre_trig = pre_trig && p.lat[0].do_plasticity != 0;
                }
                if ((post_trig && ok) || pre_trig) {
                    const float d = stdp_delta(p.lat[li], lj[u], lft_me);
                    wu = wu + d;
                    if (post_trig && pre_trig) wu = wu + d;
                    wp[u * 32] = wu;
                }
            }
            if (do_e) {
                float final_input = gap * (vj[u] - v);  // gap_junction, neuron/mod.rs:54-60
                if (NET) {
                    if (ok && (c[u] & kColTrainBit)) {
                        // spike_train_gap_junction, neuron/mod.rs:119-137
                        const uint32_t tjx = j[u] - p.train0;
                        const int lt = p.lft_in[j[u]];
                        const float v_rest = ldf(p.tf[TF_VREST], tjx);
                        if (lt < 0) final_input = v_rest;
                        else
                            final_input = gap * refract_effect(p.refract, ldf(p.tf[TF_K], tjx), p.clock, (uint32_t)lt,
                                                               ldf(p.tf[TF_VTH], tjx), v_rest, ldf(p.tf[TF_DT], tjx));
                    }
                }
                A.acc_e = A.acc_e + final_input * wu;
            }
            if (CHEMG == 1) {
                // weight_neurotransmitter_concentration + aggregate, iterate_and_spike/mod.rs:2837-2866
                const bool has = ok && ((c[u] >> (kColNtShift + ty0)) & 1u);
                const float term = tj[u][0] * wu;
                A.acc_t[0] = A.acc_t[0] + (has ? term : 0.f);
                A.cnt[0] += has ? 1u : 0u;
            } else if (CHEMG == 3) {
                const uint32_t m = ok ? (c[u] >> kColNtShift) : 0u;
#pragma unroll
                for (int ty = 0; ty < kNT; ++ty) {
                    const bool has = (m >> ty) & 1u;
                    const float term = tj[u][ty] * wu;
                    A.acc_t[ty] = A.acc_t[ty] + (has ? term : 0.f);
                    A.cnt[ty] += has ? 1u : 0u;
                }
            }
            A.n_in += ok ? 1u : 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// the fused step kernel
// ------------------------------------------------------------------------------------------------
template <int MODEL, int CHEMG, bool NTREL, bool STDP, bool NET>
__global__ void __launch_bounds__(256) step_kernel(const __grid_constant__ StepParams p) {
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (warp_global * 32u >= p.n_neurons) return;
    const uint32_t ln = warp_global * 32u + lane;  // local neuron number
    const bool valid = ln < p.n_neurons;
    const uint32_t lnc = valid ? ln : p.n_neurons - 1;  // clamped for loads
    const uint32_t i = p.own0 + lnc;                    // node index

    // ---- multi-GPU: warps whose in-edges can reach ghost nodes wait for the neighbour's boundary
    // values of the previous step (pushed over NVLink by the neighbour's own step kernel)
    bool export_lo = false, export_hi = false;
    if (!NET && (p.halo[0].active | p.halo[1].active)) {
        const uint32_t w0 = warp_global * 32u, w1 = min(w0 + 32u, p.n_neurons);
        const bool near_lo = p.halo[0].active && w0 < p.halo[0].first + p.halo[0].count;
        const bool near_hi = p.halo[1].active && w1 > p.halo[1].first;
        if (near_lo) {
            if (lane == 0) halo_wait(p.halo[0].my_flag, p.halo_epoch, p.halo_done + 2);
            __syncwarp();
        }
        if (near_hi) {
            if (lane == 0) halo_wait(p.halo[1].my_flag, p.halo_epoch, p.halo_done + 2);
            __syncwarp();
        }
        export_lo = near_lo && valid && ln >= p.halo[0].first && ln < p.halo[0].first + p.halo[0].count;
        export_hi = near_hi && valid && ln >= p.halo[1].first && ln < p.halo[1].first + p.halo[1].count;
    }

    // ---- own state and parameters: every load is issued before the edge loop so that it overlaps with it
    float v = p.v_in[i];
    const float gap = ldf(p.f[F_GAP], lnc);
    const float dt = ldf(p.f[F_DT], lnc);
    const float v_th = ldf(p.f[F_VTH], lnc);
    constexpr bool NEEDS_CM = NTREL || MODEL == SNN_MODEL_HODGKIN_HUXLEY || MODEL == SNN_MODEL_IZHIKEVICH ||
                              MODEL == SNN_MODEL_LEAKY_IZHIKEVICH || MODEL == SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE ||
                              MODEL == SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE;
    const float c_m = NEEDS_CM ? ldf(p.f[F_CM], lnc) : 1.f;
    const int lft_me = (STDP || p.lft_pp) ? p.lft_in[i] : 0;
    const uint32_t spk_word_in = __ldg(p.spk_in + (p.own0 >> 5) + warp_global);
    const bool spiking_prev = (spk_word_in >> lane) & 1u;
    const uint32_t flags = NTREL ? p.node_flags[i] : 0u;
    constexpr bool IZH = MODEL == SNN_MODEL_IZHIKEVICH || MODEL == SNN_MODEL_LEAKY_IZHIKEVICH;
    constexpr bool IF4 = MODEL == SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE || MODEL == SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE ||
                         MODEL == SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE || MODEL == SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE;
    constexpr bool ADAPT = MODEL == SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE || MODEL == SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE;
    constexpr bool LEAKY = MODEL == SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE || ADAPT;
    // model parameters (only the ones this model reads are loaded; the rest fold away)
    float w_adapt = (IZH || ADAPT) ? p.f[F_W][lnc] : 0.f;
    const float pa = IZH ? ldf(p.f[F_A], lnc) : 0.f, pb = IZH ? ldf(p.f[F_B], lnc) : 0.f;
    const float pc = IZH ? ldf(p.f[F_C], lnc) : 0.f, pd = IZH ? ldf(p.f[F_D], lnc) : 0.f;
    const float tau_m = (IZH || IF4) ? ldf(p.f[F_TAUM], lnc) : 1.f;
    const float e_l = (LEAKY || MODEL == SNN_MODEL_LEAKY_IZHIKEVICH) ? ldf(p.f[F_EL], lnc) : 0.f;
    const float v_reset = (IF4 || MODEL == SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE) ? ldf(p.f[F_VRESET], lnc) : 0.f;
    const float integ = IF4 ? ldf(p.f[F_INTEG], lnc) : 0.f;
    float refr = IF4 ? p.f[F_REFR][lnc] : 0.f;
    const float tref = IF4 ? ldf(p.f[F_TREF], lnc) : 0.f;
    const float g_l = LEAKY ? ldf(p.f[F_GL], lnc) : 1.f;
    const float leak = LEAKY ? ldf(p.f[F_LEAK], lnc) : 0.f;
    const float alpha = (ADAPT || MODEL == SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE) ? ldf(p.f[F_ALPHA], lnc) : 0.f;
    const float beta = ADAPT ? ldf(p.f[F_BETA], lnc) : 0.f;

    // the single neurotransmitter type of the CHEMG == 1 fast path
    const uint32_t ty0 = (CHEMG == 1) ? (uint32_t)(__ffs((int)p.nt_used) - 1) : 0u;

    int li = 0;
    bool post_trig = false;
    if (STDP) {
        li = NET ? lat_index(p, lnc) : 0;
        post_trig = p.apply_pending && p.lat[li].do_plasticity && lft_me == (int)p.clock - 1;
    }

    // ---- gather over in-edges (ascending presynaptic index = canonical summation order) ----------
    EdgeAcc A;
    A.acc_e = 0.f; A.n_in = 0;
#pragma unroll
    for (int ty = 0; ty < kNT; ++ty) { A.acc_t[ty] = 0.f; A.cnt[ty] = 0; }
    const bool do_e = p.electrical != 0;
    const bool do_c = NTREL && p.chemical != 0;
    gather_edges<CHEMG, STDP, NET>(p, warp_global, lane, i, v, gap, lft_me, post_trig, li, ty0, A);
    // neuron/mod.rs:722-729: divide by the number of incoming edges (1 if none)
    const float input = do_e ? A.acc_e / (A.n_in == 0 ? 1.f : (float)A.n_in) : 0.f;

    // ---- receptors (iterate_with_neurotransmitter_and_spike: kinetics then currents from pre-update V)
    float rc_total = 0.f;
    if (NTREL) {
        const uint32_t rcm = flags >> 4;
        if (do_c) {
#pragma unroll
            for (int ty = 0; ty < kNT; ++ty) {
                if (!(p.rc_used & (1u << ty))) continue;
                if (!(rcm & (1u << ty))) continue;
                const size_t o = (size_t)ty * p.rc_stride + lnc;
                float r = p.rc[RCF_R][o];
                const uint32_t cnt = (CHEMG == 1) ? (ty == (int)ty0 ? A.cnt[0] : 0u) : A.cnt[ty];
                if (cnt > 0) {
                    const float acc = (CHEMG == 1) ? A.acc_t[0] : A.acc_t[ty];
                    const float tin = acc / (float)cnt;
                    float k1 = 0.f, k2 = 0.f;
                    if (p.rck != SNN_RC_APPROXIMATE) { k1 = ldf(p.rc[RCF_K1], o); k2 = ldf(p.rc[RCF_K2], o); }
                    r = rc_apply(p.rck, r, k1, k2, tin, dt);
                    if (valid) p.rc[RCF_R][o] = r;
                }
                const float mg = (ty == SNN_NT_NMDA) ? ldf(p.rc[RCF_MG], o) : 0.f;
                rc_total += receptor_current(ty, ldf(p.rc[RCF_G], o), ldf(p.rc[RCF_E], o), mg, r, v);
            }
        } else if (MODEL == SNN_MODEL_HODGKIN_HUXLEY) {
            // hodgkin_huxley/mod.rs:161-164: the electrical-only path still subtracts the stored ligand currents
#pragma unroll
            for (int ty = 0; ty < kNT; ++ty)
                if ((p.rc_used & (1u << ty)) && (rcm & (1u << ty))) rc_total += p.rc[RCF_CUR][(size_t)ty * p.rc_stride + lnc];
        }
    }
    const float rc_dv = rc_total * (dt / c_m);  // Ionotropic::get_receptor_currents, iterate_and_spike/mod.rs:1286-1304

    // ---- neuron update --------------------------------------------------------------------------
    bool spike = false;
    float v_release;  // membrane voltage seen by the neurotransmitter kinetics (post-update, pre-reset)
    if constexpr (IZH) {
        float dv;
        if constexpr (MODEL == SNN_MODEL_IZHIKEVICH)  // integrate_and_fire/mod.rs:1255-1260
            dv = (((((0.04f * (v * v)) + (5.f * v)) + 140.f) - w_adapt) + input) * (dt / c_m);
        else                                            // :1342-1348
            dv = (((((0.04f * (v * v)) + (5.f * v)) + 140.f) - (w_adapt * (v - e_l))) + input) * (dt / c_m);
        const float dw = (pa * (pb * v - w_adapt)) * (dt / tau_m);  // :1225-1231
        if (do_c) v += dv + (-rc_dv); else v += dv;               // :226, :246
        w_adapt += dw;
        v_release = v;
        if (v >= v_th) {                                            // izhikevich_handle_spiking :1235-1247
            spike = true;
            v = pc;
            w_adapt += pd;
        }
        if (valid) p.f[F_W][lnc] = w_adapt;
    } else if constexpr (IF4) {
        float dw = 0.f, dv;
        if constexpr (MODEL == SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE) {  // :176-181
            dv = ((leak * (v - e_l)) + (integ * (input / g_l))) * (dt / tau_m);
        } else if constexpr (MODEL == SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE) {  // :324-327
            dv = (((alpha * (v - v_reset)) * (v - ldf(p.f[F_VC], lnc))) + integ * input) * (dt / tau_m);
        } else {
            if constexpr (MODEL == SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE)  // :1035-1041
                dv = (((leak * (v - e_l)) + (integ * (input / g_l))) - (w_adapt / g_l)) * (dt / c_m);
            else {                                                                  // :1138-1145
                const float sf = ldf(p.f[F_SLOPE], lnc);
                dv = ((((leak * (v - e_l)) + (sf * expf((v - v_th) / sf))) + (integ * (input / g_l))) - (w_adapt / g_l)) * (dt / c_m);
            }
            dw = (alpha * (v - e_l) - w_adapt) * (dt / tau_m);  // :1002-1009
        }
        if (do_c) v += dv + (-rc_dv); else v += dv;
        if (ADAPT) w_adapt += dw;
        v_release = v;
        // impl_default_handle_spiking :87-102 / adaptive_handle_spiking :1013-1029
        if (refr > 0.f) {
            v = v_reset;
            refr -= 1.f;
        } else if (v >= v_th) {
            spike = true;
            v = v_reset;
            if (ADAPT) w_adapt += beta;
            refr = tref / dt;
        }
        if (valid) {
            p.f[F_REFR][lnc] = refr;
            if (ADAPT) p.f[F_W][lnc] = w_adapt;
        }
    } else if constexpr (MODEL == SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE) {
        const float dv = (ldf(p.f[F_G], lnc) * (v - ldf(p.f[F_E], lnc)) + input) * dt;  // :1592-1594
        if (do_c) v += dv + (-rc_dv); else v += dv;
        v_release = v;
        if (v >= v_th) { spike = true; v = v_reset; }  // :1579-1590
    } else {  // Hodgkin-Huxley, hodgkin_huxley/mod.rs:156-241; ion_channels/mod.rs:40-44, 219-235, 268-281, 310-312
        const float last_voltage = v;
        float m = p.f[F_M][lnc], h = p.f[F_H][lnc], n = p.f[F_N][lnc];
        const float m_alpha = 0.1f * ((v + 40.f) / (1.f - expf(-(v + 40.f) / 10.f)));
        const float m_beta = 4.f * expf(-(v + 65.f) / 18.f);
        const float h_alpha = 0.07f * expf(-(v + 65.f) / 20.f);
        const float h_beta = 1.f / (expf(-(v + 35.f) / 10.f) + 1.f);
        m += dt * (m_alpha * (1.f - m) - m_beta * m);
        h += dt * (h_alpha * (1.f - h) - h_beta * h);
        const float i_na = ((powf(m, 3.f) * h) * ldf(p.f[F_GNA], lnc)) * (v - ldf(p.f[F_ENA], lnc));
        const float n_alpha = (0.01f * (v + 55.f)) / (1.f - expf(-(v + 55.f) / 10.f));
        const float n_beta = 0.125f * expf(-(v + 65.f) / 80.f);
        n += dt * (n_alpha * (1.f - n) - n_beta * n);
        const float i_k = (powf(n, 4.f) * ldf(p.f[F_GK], lnc)) * (v - ldf(p.f[F_EK], lnc));
        const float i_kl = ldf(p.f[F_GKL], lnc) * (v - ldf(p.f[F_EKL], lnc));
        const float i_sum = input - ((i_na + i_k) + i_kl);
        v += (dt * i_sum) / c_m - rc_dv;
        v_release = v;
        const uint32_t wi_word = p.was_inc[warp_global];
        const bool was_increasing = (wi_word >> lane) & 1u;
        const bool increasing_right_now = last_voltage < v;
        spike = (v > v_th) && was_increasing && !increasing_right_now;
        const uint32_t wi_new = __ballot_sync(0xffffffffu, increasing_right_now && valid);
        if (lane == 0) p.was_inc[warp_global] = wi_new;
        if (valid) { p.f[F_M][lnc] = m; p.f[F_H][lnc] = h; p.f[F_N][lnc] = n; }
    }

    // ---- neurotransmitter release: post-update voltage, previous step's spike flag
    //      (intermediate_delegate/mod.rs:18-24, integrate_and_fire/mod.rs:229)
    if (NTREL) {
        const uint32_t ntm = flags & 0xFu;
#pragma unroll
        for (int ty = 0; ty < kNT; ++ty) {
            if (!(p.nt_used & (1u << ty))) continue;
            if (!(ntm & (1u << ty))) continue;
            const size_t o = (size_t)ty * p.nt_stride + i;
            const float t_old = p.t_in[(size_t)ty * p.t_stride + i];
            float p1 = 0.f, p2 = 0.f;
            if (p.ntk != SNN_NT_DISCRETE_SPIKE) p1 = ldf(p.nt[NTF_P1], o);
            if (p.ntk == SNN_NT_DESTEXHE) p2 = ldf(p.nt[NTF_P2], o);
            const float t_new = nt_apply(p.ntk, t_old, ldf(p.nt[NTF_TMAX], o), p1, p2, v_release, spiking_prev, dt);
            if (valid) {
                p.t_out[(size_t)ty * p.t_stride + i] = t_new;
                if (export_lo) p.halo[0].peer_t[p.out_par][(size_t)ty * p.halo[0].peer_t_stride + p.halo[0].peer_node0 + (ln - p.halo[0].first)] = t_new;
                if (export_hi) p.halo[1].peer_t[p.out_par][(size_t)ty * p.halo[1].peer_t_stride + p.halo[1].peer_node0 + (ln - p.halo[1].first)] = t_new;
            }
        }
    }

    // ---- spike compaction: one bit per neuron, one word per warp --------------------------------
    const uint32_t spk_word = __ballot_sync(0xffffffffu, spike && valid);
    if (lane == 0) {
        p.spk_out[(p.own0 >> 5) + warp_global] = spk_word;
        if (p.spike_hist) p.spike_hist[warp_global] = spk_word;
    }
    if (valid) {
        p.v_out[i] = v;
        // set_last_firing_time(Some(internal_clock)), neuron/mod.rs:964-966
        int lft_new = lft_me;
        if (p.lft_pp) { lft_new = spike ? (int)p.clock : lft_me; p.lft_out[i] = lft_new; }
        else if (spike) { lft_new = (int)p.clock; p.lft_out[i] = lft_new; }
        if (p.grid_hist) p.grid_hist[ln] = v;  // GridVoltageHistory::update, neuron/mod.rs:293-296
        if (export_lo) {
            const HaloDir &H = p.halo[0];
            const uint32_t dst = H.peer_node0 + (ln - H.first);
            H.peer_v[p.out_par][dst] = v;
            if (p.lft_pp) H.peer_lft[p.out_par][dst] = lft_new;
        }
        if (export_hi) {
            const HaloDir &H = p.halo[1];
            const uint32_t dst = H.peer_node0 + (ln - H.first);
            H.peer_v[p.out_par][dst] = v;
            if (p.lft_pp) H.peer_lft[p.out_par][dst] = lft_new;
        }
    }

    // ---- multi-GPU: publish "my boundary values of this step have landed" to each neighbour.
    // Every exporting warp fences its remote stores, then bumps a local counter; the last one raises
    // the neighbour's arrival flag to halo_epoch + 1.
    if (!NET && (p.halo[0].active | p.halo[1].active)) {
        const uint32_t w0 = warp_global * 32u, w1 = min(w0 + 32u, p.n_neurons);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const HaloDir &H = p.halo[d];
            if (!H.active) continue;
            const bool mine = w0 < H.first + H.count && w1 > H.first;
            if (!mine) continue;
            __syncwarp();
            if (lane == 0) {
                __threadfence_system();
                const uint32_t first_w = H.first >> 5, last_w = (H.first + H.count - 1) >> 5;
                const uint32_t n_warps = last_w - first_w + 1;
                const unsigned int done = atomicAdd(&p.halo_done[d], 1u) + 1u;
                if (done == n_warps) {
                    p.halo_done[d] = 0u;
                    __threadfence_system();
                    st_release_sys(H.peer_flag, p.halo_epoch + 1ull);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: push the current boundary state into the neighbours' ghost slots (start of every run)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ StepParams p) {
    // here *_in are the CURRENT state buffers and out_par their parity
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const HaloDir &H = p.halo[d];
        if (!H.active) continue;
        if (t < H.count) {
            const uint32_t i = p.own0 + H.first + t, dst = H.peer_node0 + t;
            H.peer_v[p.out_par][dst] = p.v_in[i];
            H.peer_lft[p.out_par][dst] = p.lft_in[i];
            for (int ty = 0; ty < kNT; ++ty)
                if (p.nt_used & (1u << ty))
                    H.peer_t[p.out_par][(size_t)ty * H.peer_t_stride + dst] = p.t_in[(size_t)ty * p.t_stride + i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(&p.halo_done[0], 1u) + 1u;
        if (done == gridDim.x) {
            p.halo_done[0] = 0u;
            __threadfence_system();
            for (int d = 0; d < 2; ++d)
                if (p.halo[d].active) st_release_sys(p.halo[d].peer_flag, p.halo_epoch + 1ull);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// end-of-run flush of the lazily applied STDP (the last step's updates are still pending)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flush_stdp_kernel(const __grid_constant__ StepParams p) {
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (warp_global * 32u >= p.n_neurons) return;
    const uint32_t ln = warp_global * 32u + lane;
    if (ln >= p.n_neurons) return;
    const uint32_t i = p.own0 + ln;
    const int lft_me = p.lft_in[i];
    const int li = lat_index(p, ln);
    const bool post_trig = p.lat[li].do_plasticity && lft_me == (int)p.clock - 1;
    const uint32_t k0 = p.slice_off[warp_global], k1 = p.slice_off[warp_global + 1];
    for (uint32_t k = k0; k < k1; ++k) {
        const size_t e = (size_t)k * 32u + lane;
        const uint32_t c = __ldg(p.col + e);
        if (c == kColPad) continue;
        const uint32_t j = c & kColIdxMask;
        // spike trains: last_firing_time from before their last iterate (see gather_edges); lft_out is that buffer here
        const int lft_pre = (c & kColTrainBit) ? p.lft_out[j] : p.lft_in[j];
        bool pre_trig = !(c & kColTrainBit) && lft_pre == (int)p.clock - 1;
        if (pre_trig) pre_trig = p.lat[p.n_lat > 1 ? lat_index(p, j - p.own0) : 0].do_plasticity != 0;
        if (post_trig | pre_trig) {
            float w = p.wgt[e];
            const float d = stdp_delta(p.lat[li], lft_pre, lft_me);
            w = w + d;
            if (post_trig & pre_trig) w = w + d;
            p.wgt[e] = w;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// derived fields (currents, gate rates): recomputed once after a run from the retained pre-update V
// ------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ StepParams p, const float *v_prev) {
    const uint32_t ln = blockIdx.x * blockDim.x + threadIdx.x;
    if (ln >= p.n_neurons) return;
    const uint32_t i = p.own0 + ln;
    const float v = v_prev[i];
    if (p.chemical) {
        const uint32_t rcm = p.node_flags[i] >> 4;
        for (int ty = 0; ty < kNT; ++ty) {
            if (!(rcm & (1u << ty))) continue;
            const size_t o = (size_t)ty * p.rc_stride + ln;
            const float mg = (ty == SNN_NT_NMDA) ? p.rc[RCF_MG][o] : 0.f;
            p.rc[RCF_CUR][o] = receptor_current(ty, p.rc[RCF_G][o], p.rc[RCF_E][o], mg, p.rc[RCF_R][o], v);
        }
    }
    if constexpr (MODEL == SNN_MODEL_HODGKIN_HUXLEY) {
        const float m = p.f[F_M][ln], h = p.f[F_H][ln], n = p.f[F_N][ln];
        p.f[F_M_ALPHA][ln] = 0.1f * ((v + 40.f) / (1.f - expf(-(v + 40.f) / 10.f)));
        p.f[F_M_BETA][ln] = 4.f * expf(-(v + 65.f) / 18.f);
        p.f[F_H_ALPHA][ln] = 0.07f * expf(-(v + 65.f) / 20.f);
        p.f[F_H_BETA][ln] = 1.f / (expf(-(v + 35.f) / 10.f) + 1.f);
        p.f[F_NA_CUR][ln] = ((powf(m, 3.f) * h) * p.f[F_GNA][ln]) * (v - p.f[F_ENA][ln]);
        p.f[F_N_ALPHA][ln] = (0.01f * (v + 55.f)) / (1.f - expf(-(v + 55.f) / 10.f));
        p.f[F_N_BETA][ln] = 0.125f * expf(-(v + 65.f) / 80.f);
        p.f[F_K_CUR][ln] = (powf(n, 4.f) * p.f[F_GK][ln]) * (v - p.f[F_EK][ln]);
        p.f[F_KL_CUR][ln] = p.f[F_GKL][ln] * (v - p.f[F_EKL][ln]);
    }
}

// ------------------------------------------------------------------------------------------------
// spike trains (SpikeTrainLattice::iterate, neuron/mod.rs:1377-1393)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t *hi) {
    const uint64_t prod = (uint64_t)a * b;
    *hi = (uint32_t)(prod >> 32);
    return (uint32_t)prod;
}

// Philox4x32-10 counter-based generator keyed by (seed), counter (train index, clock)
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint32_t idx, uint32_t clock) {
    uint32_t c0 = idx, c1 = clock, c2 = 0x5EEDu, c3 = 0u;
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

__global__ void __launch_bounds__(256) train_kernel(const __grid_constant__ TrainParams p) {
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (warp_global * 32u >= p.n_trains) return;
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
        const float u = (float)(philox_u32(p.seed, tc, p.tl_clock[tl]) >> 8) * (1.0f / 16777215.0f);
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

// ------------------------------------------------------------------------------------------------
// graph construction on the device
// ------------------------------------------------------------------------------------------------
// CSR (by post, pre ascending) -> sliced ELL, one warp per slice
__global__ void sell_from_csr_kernel(const uint64_t *row_ptr, const uint32_t *pre, const float *w,
                                     const uint8_t *node_flags, uint32_t train0, uint32_t n_rows,
                                     const uint32_t *slice_off, uint32_t *col, float *wgt) {
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (warp_global * 32u >= n_rows) return;
    const uint32_t r = warp_global * 32u + lane;
    uint64_t b = 0, len = 0;
    if (r < n_rows) { b = row_ptr[r]; len = row_ptr[r + 1] - b; }
    const uint32_t k0 = slice_off[warp_global], k1 = slice_off[warp_global + 1];
    for (uint32_t k = k0; k < k1; ++k) {
        const size_t e = (size_t)k * 32u + lane;
        const uint32_t kk = k - k0;
        if (kk < len) {
            const uint32_t j = pre[b + kk];
            uint32_t c = j | ((uint32_t)(node_flags[j] & 0x7u) << kColNtShift);
            if (j >= train0) c |= kColTrainBit;
            col[e] = c;
            wgt[e] = w[b + kk];
        } else {
            col[e] = kColPad;
            wgt[e] = 0.f;
        }
    }
}

// Moore-neighbourhood stencil written straight into sliced-ELL form (uniform slice width)
__global__ void sell_grid_kernel(uint32_t rows_local, uint32_t cols, uint32_t row0_global, uint32_t rows_global,
                                 uint32_t radius, float weight, uint32_t own0, const uint8_t *node_flags,
                                 uint32_t width, uint32_t *slice_off, uint32_t *col, float *wgt) {
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t n = (uint64_t)rows_local * cols;
    if ((uint64_t)warp_global * 32u >= n) return;
    if (lane == 0) {
        slice_off[warp_global] = warp_global * width;
        if ((uint64_t)(warp_global + 1) * 32u >= n) slice_off[warp_global + 1] = (warp_global + 1) * width;
    }
    const uint64_t ln = (uint64_t)warp_global * 32u + lane;
    const bool valid = ln < n;
    const int64_t r = valid ? (int64_t)(ln / cols) + row0_global : 0, c0 = valid ? (int64_t)(ln % cols) : 0;
    const int R = (int)radius;
    uint32_t k = 0;
    const size_t base = (size_t)warp_global * width * 32u + lane;
    if (valid) {
        for (int dr = -R; dr <= R; ++dr)
            for (int dc = -R; dc <= R; ++dc) {
                const int64_t a = r + dr, b = c0 + dc;
                if ((dr == 0 && dc == 0) || a < 0 || b < 0 || a >= (int64_t)rows_global || b >= (int64_t)cols) continue;
                // node index of (a, b): the first owned neuron sits at own0; rows above/below the strip are ghosts
                const int64_t j = (int64_t)own0 + (a - (int64_t)row0_global) * (int64_t)cols + b;
                col[base + (size_t)k * 32u] = (uint32_t)j | ((uint32_t)(node_flags[j] & 0x7u) << kColNtShift);
                wgt[base + (size_t)k * 32u] = weight;
                ++k;
            }
    }
    for (; k < width; ++k) {
        col[base + (size_t)k * 32u] = kColPad;
        wgt[base + (size_t)k * 32u] = 0.f;
    }
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t v, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// u32 bools (reference buffer dtype for is_spiking) <-> 1 bit per node
__global__ void bits_from_u32_kernel(const uint32_t *src, uint32_t *words, uint64_t n, uint64_t bit0) {
    // arbitrary bit offset (a lattice segment need not start on a word boundary): one atomic per bit,
    // used only by set_field
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t b = bit0 + t;
    if (src[t] != 0u) atomicOr(&words[b >> 5], 1u << (b & 31u));
    else atomicAnd(&words[b >> 5], ~(1u << (b & 31u)));
}
__global__ void u32_from_bits_kernel(const uint32_t *words, uint32_t *dst, uint64_t n, uint64_t bit0) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = (words[(bit0 + t) >> 5] >> ((bit0 + t) & 31u)) & 1u;
}

// neuron-major [n][3] (reference layout) <-> type-major [3][stride] (device layout)
__global__ void transpose_in_kernel(const float *src_nm, float *dst_tm, uint64_t n, uint64_t stride) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * kNT) dst_tm[(t % kNT) * stride + t / kNT] = src_nm[t];
}
__global__ void transpose_out_kernel(const float *src_tm, float *dst_nm, uint64_t n, uint64_t stride) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * kNT) dst_nm[t] = src_tm[(t % kNT) * stride + t / kNT];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

template <int MODEL, int CHEMG, bool NTREL, bool NET>
static cudaError_t launch_step_3(const StepParams &p, bool stdp, cudaStream_t s) {
    const unsigned grid = blocks_for((uint64_t)((p.n_neurons + 31u) / 32u) * 32u, 256);
    if (stdp) step_kernel<MODEL, CHEMG, NTREL, true, NET><<<grid, 256, 0, s>>>(p);
    else step_kernel<MODEL, CHEMG, NTREL, false, NET><<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_step_model(const StepParams &p, int chemg, bool ntrel, bool stdp, bool net, cudaStream_t s) {
    if (net) {
        // several lattices and/or spike trains: general per-edge type masks
        if (chemg) return launch_step_3<MODEL, 3, true, true>(p, stdp, s);
        if (ntrel) return launch_step_3<MODEL, 0, true, true>(p, stdp, s);
        return launch_step_3<MODEL, 0, false, true>(p, stdp, s);
    }
    if (chemg == 1) return launch_step_3<MODEL, 1, true, false>(p, stdp, s);
    if (chemg == 3) return launch_step_3<MODEL, 3, true, false>(p, stdp, s);
    if (ntrel) return launch_step_3<MODEL, 0, true, false>(p, stdp, s);
    return launch_step_3<MODEL, 0, false, false>(p, stdp, s);
}

cudaError_t launch_step(const StepParams &p, int model, int chemg, bool ntrel, bool stdp, bool net, cudaStream_t s) {
    if (p.n_neurons == 0) return cudaSuccess;
    switch (model) {
    case SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE: return launch_step_model<SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE: return launch_step_model<SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE: return launch_step_model<SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE: return launch_step_model<SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_IZHIKEVICH: return launch_step_model<SNN_MODEL_IZHIKEVICH>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_LEAKY_IZHIKEVICH: return launch_step_model<SNN_MODEL_LEAKY_IZHIKEVICH>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE: return launch_step_model<SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_HODGKIN_HUXLEY: return launch_step_model<SNN_MODEL_HODGKIN_HUXLEY>(p, chemg, ntrel, stdp, net, s);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_trains(const TrainParams &p, cudaStream_t s) {
    if (p.n_trains == 0) return cudaSuccess;
    train_kernel<<<blocks_for((uint64_t)((p.n_trains + 31u) / 32u) * 32u, 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_halo_push(const StepParams &p, cudaStream_t s) {
    const uint32_t cnt = max(p.halo[0].active ? p.halo[0].count : 0u, p.halo[1].active ? p.halo[1].count : 0u);
    if (cnt == 0) return cudaSuccess;
    halo_push_kernel<<<blocks_for(cnt, 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_flush_stdp(const StepParams &p, cudaStream_t s) {
    if (p.n_neurons == 0) return cudaSuccess;
    flush_stdp_kernel<<<blocks_for((uint64_t)((p.n_neurons + 31u) / 32u) * 32u, 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const StepParams &p, int model, const float *v_prev, cudaStream_t s) {
    if (p.n_neurons == 0) return cudaSuccess;
    const unsigned grid = blocks_for(p.n_neurons, 256);
    if (model == SNN_MODEL_HODGKIN_HUXLEY) finalize_kernel<SNN_MODEL_HODGKIN_HUXLEY><<<grid, 256, 0, s>>>(p, v_prev);
    else finalize_kernel<SNN_MODEL_IZHIKEVICH><<<grid, 256, 0, s>>>(p, v_prev);
    return cudaGetLastError();
}

cudaError_t launch_sell_from_csr(const uint64_t *row_ptr, const uint32_t *pre, const float *w, const uint8_t *node_flags,
                                 uint32_t train0, uint32_t n_rows, const uint32_t *slice_off, uint32_t *col, float *wgt,
                                 cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    sell_from_csr_kernel<<<blocks_for((uint64_t)((n_rows + 31u) / 32u) * 32u, 256), 256, 0, s>>>(
        row_ptr, pre, w, node_flags, train0, n_rows, slice_off, col, wgt);
    return cudaGetLastError();
}

cudaError_t launch_sell_grid(uint32_t rows_local, uint32_t cols, uint32_t row0_global, uint32_t rows_global,
                             uint32_t radius, float weight, uint32_t own0, const uint8_t *node_flags, uint32_t width,
                             uint32_t *slice_off, uint32_t *col, float *wgt, cudaStream_t s) {
    const uint64_t n = (uint64_t)rows_local * cols;
    if (n == 0) return cudaSuccess;
    sell_grid_kernel<<<blocks_for(((n + 31u) / 32u) * 32u, 256), 256, 0, s>>>(
        rows_local, cols, row0_global, rows_global, radius, weight, own0, node_flags, width, slice_off, col, wgt);
    return cudaGetLastError();
}

cudaError_t launch_fill_u32(uint32_t *p, uint32_t v, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    fill_u32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, v, n);
    return cudaGetLastError();
}
cudaError_t launch_bits_from_u32(const uint32_t *src, uint32_t *words, uint64_t n, uint64_t bit0, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    bits_from_u32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, words, n, bit0);
    return cudaGetLastError();
}
cudaError_t launch_u32_from_bits(const uint32_t *words, uint32_t *dst, uint64_t n, uint64_t bit0, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    u32_from_bits_kernel<<<blocks_for(n, 256), 256, 0, s>>>(words, dst, n, bit0);
    return cudaGetLastError();
}
cudaError_t launch_transpose_in(const float *src_nm, float *dst_tm, uint64_t n, uint64_t stride, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    transpose_in_kernel<<<blocks_for(n * kNT, 256), 256, 0, s>>>(src_nm, dst_tm, n, stride);
    return cudaGetLastError();
}
cudaError_t launch_transpose_out(const float *src_tm, float *dst_nm, uint64_t n, uint64_t stride, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    transpose_out_kernel<<<blocks_for(n * kNT, 256), 256, 0, s>>>(src_tm, dst_nm, n, stride);
    return cudaGetLastError();
}

}  // namespace snn
