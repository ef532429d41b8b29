// step_wide.cu — the general step kernel for graphs with wide rows (hundreds of in-edges per neuron): one CTA per 32-row
// slice, all of its warps share the per-edge work of the slice (gather_edges_wide, step_body.cuh), the leader warp does the
// ordered accumulation and the neuron update.  Same arithmetic, bit-identical results; used for single-GPU handles whose
// mean slice width is >= kWideMinWidth (BASELINE.json configs[3]: 784 spike trains -> 400 excitatory <-> 400 inhibitory).
#include "step_body.cuh"

namespace snn {

template <int MODEL, int CHEMG, bool NTREL, bool STDP, bool NET>
__global__ void __launch_bounds__(kWideWarps * 32) step_wide_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ __align__(16) unsigned char wide_sm[];   // 2 x wide_buf_bytes(CHEMG): double-buffered chunk terms
    const uint32_t warp_global = blockIdx.x;   // one CTA per slice
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t ln = warp_global * 32u + lane;
    const bool valid = ln < p.n_neurons;
    const uint32_t lnc = valid ? ln : p.n_neurons - 1;
    const uint32_t k0 = p.uniform_width ? warp_global * p.uniform_width : __ldg(p.slice_off + warp_global);
    const uint32_t k1 = p.uniform_width ? k0 + p.uniform_width : __ldg(p.slice_off + warp_global + 1);
    const float *t0 = nullptr;
    if (CHEMG == 1) t0 = p.t_in + (size_t)(__ffs((int)p.nt_used) - 1) * p.t_stride;
    const WideSrc src{{p, lnc, p.own0 + lnc, lane, k0, k1, t0}, warp, (uint32_t)kWideWarps, wide_sm};
    neuron_step<MODEL, CHEMG, NTREL, STDP, NET>(p, src, warp_global, lane, ln, lnc, valid, false, false);
}

template <int MODEL, int CHEMG, bool NTREL, bool NET>
static cudaError_t launch_wide_3(const StepParams &p, bool stdp, cudaStream_t s) {
    const unsigned grid = (p.n_neurons + 31u) / 32u;
    constexpr size_t smem = 2u * wide_buf_bytes(CHEMG);
    if (stdp) {
        auto k = step_wide_kernel<MODEL, CHEMG, NTREL, true, NET>;
        if (smem > 48u * 1024u) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        k<<<grid, kWideWarps * 32, smem, s>>>(p);
    } else {
        auto k = step_wide_kernel<MODEL, CHEMG, NTREL, false, NET>;
        if (smem > 48u * 1024u) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        k<<<grid, kWideWarps * 32, smem, s>>>(p);
    }
    return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_wide_model(const StepParams &p, int chemg, bool ntrel, bool stdp, bool net, cudaStream_t s) {
    if (net) {
        if (chemg) return launch_wide_3<MODEL, 3, true, true>(p, stdp, s);
        if (ntrel) return launch_wide_3<MODEL, 0, true, true>(p, stdp, s);
        return launch_wide_3<MODEL, 0, false, true>(p, stdp, s);
    }
    if (chemg == 1) return launch_wide_3<MODEL, 1, true, false>(p, stdp, s);
    if (chemg == 3) return launch_wide_3<MODEL, 3, true, false>(p, stdp, s);
    if (ntrel) return launch_wide_3<MODEL, 0, true, false>(p, stdp, s);
    return launch_wide_3<MODEL, 0, false, false>(p, stdp, s);
}

cudaError_t launch_step_wide(const StepParams &p, int model, int chemg, bool ntrel, bool stdp, bool net, cudaStream_t s) {
    if (p.n_neurons == 0) return cudaSuccess;
    switch (model) {
    case SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE: return launch_wide_model<SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE: return launch_wide_model<SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE: return launch_wide_model<SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE: return launch_wide_model<SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_IZHIKEVICH: return launch_wide_model<SNN_MODEL_IZHIKEVICH>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_LEAKY_IZHIKEVICH: return launch_wide_model<SNN_MODEL_LEAKY_IZHIKEVICH>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE: return launch_wide_model<SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_HODGKIN_HUXLEY: return launch_wide_model<SNN_MODEL_HODGKIN_HUXLEY>(p, chemg, ntrel, stdp, net, s);
    case SNN_MODEL_BCM_IZHIKEVICH: return launch_wide_model<SNN_MODEL_BCM_IZHIKEVICH>(p, chemg, ntrel, stdp, net, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace snn
