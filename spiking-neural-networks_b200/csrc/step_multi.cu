// step_multi.cu — K timesteps per launch for lattices and networks that live in L2 (BASELINE.json configs[0], [2], [3]).
//
// The reference's accelerator loop costs 4 launches + 4 host waits per timestep (gpu_lattices/mod.rs:809-881); the one-launch-
// per-timestep kernels of this library (kernels.cu, step_wide.cu) still sit on the launch-to-launch floor for small lattices:
// a 100 x 100 Izhikevich lattice steps in 4.1 us of which ~1 us is work.  Here the whole chunk of timesteps runs inside ONE
// cooperative launch: every CTA keeps a private copy of the step parameters in shared memory, rewrites the few fields that
// change from step to step (clock, ping-pong buffers, history record) and meets the other CTAs at a grid-wide barrier between
// timesteps — and, in networks whose neurons read the spike trains' previous last_firing_time (lazy STDP), once more between the
// neurons and the spike trains of a step.  The per-neuron step is the shared neuron_step (step_body.cuh) and the spike-train
// step the shared train_step (train_body.cuh): same arithmetic in the same order, bit-identical to the per-launch path.
//
// Coherence: data written in step s is read in step s + 1 by other SMs.  The barrier is release (fence + atomic) on arrival and
// acquire (ld.acquire.gpu, which also drops the SM's stale L1 lines) on departure; nothing that changes during the launch is read
// through the non-coherent path (__ldg).
#include "step_body.cuh"
#include "train_body.cuh"

namespace snn {

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid-wide barrier on a monotonically increasing counter (zeroed by the host before the launch); every CTA of the cooperative
// launch is resident, so spinning is safe
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_gpu_u32(counter) < target) {}
    }
    __syncthreads();
}

template <int MODEL, int CHEMG, bool NTREL, bool STDP, bool NET, bool WIDE>
__global__ void __launch_bounds__(WIDE ? kWideWarps * 32 : 256)
step_multi_kernel(const __grid_constant__ StepParams p0, const __grid_constant__ TrainParams t0, const __grid_constant__ MultiParams m) {
    extern __shared__ __align__(16) unsigned char multi_sm[];   // WIDE: 2 x wide_buf_bytes(CHEMG)
    __shared__ StepParams p;
    __shared__ TrainParams tp;
    {   // cooperative copy of the launch parameters into shared memory
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&p0);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&p);
        for (uint32_t k = threadIdx.x; k < sizeof(StepParams) / 4; k += blockDim.x) dst[k] = src[k];
        const uint32_t *src2 = reinterpret_cast<const uint32_t *>(&t0);
        uint32_t *dst2 = reinterpret_cast<uint32_t *>(&tp);
        for (uint32_t k = threadIdx.x; k < sizeof(TrainParams) / 4; k += blockDim.x) dst2[k] = src2[k];
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t warps_per_cta = blockDim.x >> 5;
    const uint32_t n_slices = (p0.n_neurons + 31u) >> 5, n_twarps = (t0.n_trains + 31u) >> 5;
    unsigned int arrivals = 0;
    for (uint32_t s = 0; s < m.steps; ++s) {
        __syncthreads();   // every warp is done with the previous step's parameters
        if (threadIdx.x == 0) {
            const uint32_t in = (m.cur ^ s) & 1u, out = in ^ 1u;
            p.clock = p0.clock + s;
            p.apply_pending = (STDP && (s > 0u || m.first_pending)) ? 1u : 0u;
            p.v_in = m.v[in]; p.v_out = m.v[out];
            p.spk_in = m.spk[in]; p.spk_out = m.spk[out];
            p.t_in = m.t[in]; p.t_out = m.t[out];
            const uint32_t li = m.lft_pp ? ((m.lft_loc ^ s) & 1u) : m.lft_loc;
            p.lft_in = m.lft[li]; p.lft_out = m.lft[m.lft_pp ? li ^ 1u : li];
            p.grid_hist = m.grid_hist ? m.grid_hist + (size_t)s * m.n_neurons : nullptr;
            p.spike_hist = m.spike_hist ? m.spike_hist + (size_t)s * m.n_words : nullptr;
            p.out_par = out;
            if (t0.n_trains) {
                tp.v_in = p.v_in; tp.v_out = p.v_out; tp.spk_in = p.spk_in; tp.spk_out = p.spk_out;
                tp.t_in = p.t_in; tp.t_out = p.t_out; tp.lft_in = p.lft_in; tp.lft_out = p.lft_out;
                for (int l = 0; l < t0.n_tl; ++l) tp.tl_clock[l] = t0.tl_clock[l] + s;
                tp.draw = t0.draw + s;
                tp.grid_hist = m.tgrid_hist ? m.tgrid_hist + (size_t)s * m.n_trains : nullptr;
                tp.spike_hist = m.tspike_hist ? m.tspike_hist + (size_t)s * m.t_words : nullptr;
            }
        }
        __syncthreads();
        // ---- the neurons of this timestep
        if constexpr (WIDE) {
            for (uint32_t slice = blockIdx.x; slice < n_slices; slice += gridDim.x) {   // uniform per CTA: gather_edges_wide syncs the CTA
                const uint32_t ln = slice * 32u + lane;
                const bool valid = ln < p.n_neurons;
                const uint32_t lnc = valid ? ln : p.n_neurons - 1;
                const uint32_t k0 = p.uniform_width ? slice * p.uniform_width : __ldg(p.slice_off + slice);
                const uint32_t k1 = p.uniform_width ? k0 + p.uniform_width : __ldg(p.slice_off + slice + 1);
                const float *tt0 = nullptr;
                if (CHEMG == 1) tt0 = p.t_in + (size_t)(__ffs((int)p.nt_used) - 1) * p.t_stride;
                const WideSrc src{{p, lnc, p.own0 + lnc, lane, k0, k1, tt0}, warp, (uint32_t)kWideWarps, multi_sm};
                neuron_step<MODEL, CHEMG, NTREL, STDP, NET>(p, src, slice, lane, ln, lnc, valid, false, false);
                __syncthreads();   // the chunk buffers are reused by the CTA's next slice
            }
        } else {
            for (uint32_t slice = blockIdx.x * warps_per_cta + warp; slice < n_slices; slice += gridDim.x * warps_per_cta) {
                const uint32_t ln = slice * 32u + lane;
                const bool valid = ln < p.n_neurons;
                const uint32_t lnc = valid ? ln : p.n_neurons - 1;
                const uint32_t k0 = p.uniform_width ? slice * p.uniform_width : __ldg(p.slice_off + slice);
                const uint32_t k1 = p.uniform_width ? k0 + p.uniform_width : __ldg(p.slice_off + slice + 1);
                const float *tt0 = nullptr;
                if (CHEMG == 1) tt0 = p.t_in + (size_t)(__ffs((int)p.nt_used) - 1) * p.t_stride;
                const GlobalSrc src{p, lnc, p.own0 + lnc, lane, k0, k1, tt0};
                neuron_step<MODEL, CHEMG, NTREL, STDP, NET>(p, src, slice, lane, ln, lnc, valid, false, false);
            }
        }
        // ---- the spike trains of this timestep (they step after the neurons, neuron/mod.rs:2582-2591).  The lazy STDP of the neurons
        // reads the trains' last_firing_time from before this step out of lft_out, which the trains overwrite now: wait for every
        // neuron first
        if (NET && t0.n_trains) {
            if (m.train_sync) { arrivals += gridDim.x; grid_barrier(m.barrier, arrivals); }
            for (uint32_t tw = blockIdx.x * warps_per_cta + warp; tw < n_twarps; tw += gridDim.x * warps_per_cta) train_step(tp, tw, lane);
        }
        if (s + 1u < m.steps) { arrivals += gridDim.x; grid_barrier(m.barrier, arrivals); }
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template <int MODEL, int CHEMG, bool NTREL, bool STDP, bool NET, bool WIDE>
static cudaError_t launch_multi_5(const StepParams &p, const TrainParams &t, const MultiParams &m, int device, bool dry, cudaStream_t s) {
    auto k = step_multi_kernel<MODEL, CHEMG, NTREL, STDP, NET, WIDE>;
    const int threads = WIDE ? kWideWarps * 32 : 256;
    const size_t smem = WIDE ? 2u * wide_buf_bytes(CHEMG) : 0u;
    cudaError_t e;
    if (smem > 40u * 1024u) { e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    const uint32_t n_slices = (p.n_neurons + 31u) / 32u, n_twarps = (t.n_trains + 31u) / 32u;
    const uint32_t want_n = WIDE ? n_slices : (n_slices + 7u) / 8u, want_t = WIDE ? (n_twarps + kWideWarps - 1u) / kWideWarps : (n_twarps + 7u) / 8u;
    const uint32_t want = want_n > want_t ? want_n : want_t;
    const uint32_t cap = (uint32_t)per_sm * (uint32_t)sms;
    if (cap == 0) return cudaErrorLaunchOutOfResources;
    // more than a few slices per warp and timestep: the launch-per-step kernels fill the machine better
    if (want > 4u * cap) return cudaErrorLaunchOutOfResources;
    if (dry) return cudaSuccess;
    const unsigned grid = want < cap ? (want ? want : 1u) : cap;
    void *args[] = {(void *)&p, (void *)&t, (void *)&m};
    return cudaLaunchCooperativeKernel((const void *)k, dim3(grid), dim3(threads), args, smem, s);
}

template <int MODEL, int CHEMG, bool NTREL, bool NET>
static cudaError_t launch_multi_3(const StepParams &p, const TrainParams &t, const MultiParams &m, bool stdp, bool wide, int device, bool dry,
                                  cudaStream_t s) {
    if (stdp) return wide ? launch_multi_5<MODEL, CHEMG, NTREL, true, NET, true>(p, t, m, device, dry, s)
                          : launch_multi_5<MODEL, CHEMG, NTREL, true, NET, false>(p, t, m, device, dry, s);
    return wide ? launch_multi_5<MODEL, CHEMG, NTREL, false, NET, true>(p, t, m, device, dry, s)
                : launch_multi_5<MODEL, CHEMG, NTREL, false, NET, false>(p, t, m, device, dry, s);
}

template <int MODEL>
static cudaError_t launch_multi_model(const StepParams &p, const TrainParams &t, const MultiParams &m, int chemg, bool ntrel, bool stdp, bool net,
                                      bool wide, int device, bool dry, cudaStream_t s) {
    if (net) {
        if (chemg) return launch_multi_3<MODEL, 3, true, true>(p, t, m, stdp, wide, device, dry, s);
        if (ntrel) return launch_multi_3<MODEL, 0, true, true>(p, t, m, stdp, wide, device, dry, s);
        return launch_multi_3<MODEL, 0, false, true>(p, t, m, stdp, wide, device, dry, s);
    }
    if (chemg == 1) return launch_multi_3<MODEL, 1, true, false>(p, t, m, stdp, wide, device, dry, s);
    if (chemg == 3) return launch_multi_3<MODEL, 3, true, false>(p, t, m, stdp, wide, device, dry, s);
    if (ntrel) return launch_multi_3<MODEL, 0, true, false>(p, t, m, stdp, wide, device, dry, s);
    return launch_multi_3<MODEL, 0, false, false>(p, t, m, stdp, wide, device, dry, s);
}

cudaError_t launch_step_multi(const StepParams &p, const TrainParams &t, const MultiParams &m, int model, int chemg, bool ntrel, bool stdp,
                              bool net, bool wide, int device, bool dry, cudaStream_t s) {
    switch (model) {
    case SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE: return launch_multi_model<SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE: return launch_multi_model<SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE: return launch_multi_model<SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE: return launch_multi_model<SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_IZHIKEVICH: return launch_multi_model<SNN_MODEL_IZHIKEVICH>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_LEAKY_IZHIKEVICH: return launch_multi_model<SNN_MODEL_LEAKY_IZHIKEVICH>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE: return launch_multi_model<SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_HODGKIN_HUXLEY: return launch_multi_model<SNN_MODEL_HODGKIN_HUXLEY>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    case SNN_MODEL_BCM_IZHIKEVICH: return launch_multi_model<SNN_MODEL_BCM_IZHIKEVICH>(p, t, m, chemg, ntrel, stdp, net, wide, device, dry, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace snn
