// step_win.cu — the step kernel for 8-neighbour (radius-1) stencil lattices: one persistent CTA per SM, every operand
// staged by TMA, including the neighbour rows.
//
// Why: the L1-gather kernels (kernels.cu, step_tma.cu) spend their time on ~24 dependent neighbour gathers per neuron
// and on ~870 issued instructions per 32 neurons (ncu, profiles/r1_step_tma_full.txt: 53 % issue slots, 32 % of stall
// samples on the long scoreboard, 40 % of DRAM peak).  For a row-major lattice the presynaptic neurons of a tile of
// 256 consecutive neurons lie in three contiguous index ranges of the node arrays (the tile's own range and the same
// range one row up / one row down, each widened by one neuron), so here the producer lane copies those three windows
// of V, last_firing_time and t with TMA bulk copies next to the tile's own parameters and col/weight rows.  Consumer
// warps then run the shared neuron_step (step_body.cuh — same arithmetic, bit-identical results) on shared memory
// only: no global load is left on their critical path, operand offsets are compile-time immediates (WinLayout,
// common.h), and HBM latency is hidden by the depth of the stage ring instead of by warp occupancy.
//
// CTA = G groups of 8 consumer warps + kWinProducers producer warps; tile n of the CTA goes to group n % G and stage n % S.
// The col words are still read and used (the window is selected by comparing the presynaptic index against the
// tile's bounds), so any per-edge weight / type content of the sliced-ELL table keeps working; only the index
// pattern must be the radius-1 stencil that set_graph_grid generates.
#include "step_body.cuh"
#include "tma_util.cuh"

namespace snn {

// shared-memory loads with a compile-time immediate offset.  `volatile` keeps them between the mbarrier wait and the
// mbarrier arrive of their stage (both volatile asm); ptxas still schedules them freely inside that span.
template <uint32_t OFF> __device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <uint32_t OFF> __device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <uint32_t OFF> __device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ float lds_f32_rt(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <int MODEL, int CHEMG, bool NTREL, bool STDP>
struct WinSrc {
    static constexpr bool kEarlyLoads = false;   // operands sit in shared memory: read them where they are used
    static constexpr bool kCheapEdges = false;
    static constexpr bool kWide = false;
    static constexpr uint32_t kWidth = kWinWidth;
    static __host__ __device__ constexpr WinLayout lay() { return win_layout(MODEL, CHEMG, NTREL, STDP); }

    const StepParams &p;
    const WinParams &wp;
    uint32_t me;                // shared address of this stage + 4 * (neuron within the tile): per-neuron operand rows
    uint32_t own;               // shared address of this neuron inside a (row-relative) window row
    uint32_t edge;              // shared address of this lane's first col word relative to o_col (weights: + o_wgt - o_col)
    uint32_t lo_thr, hi_thr;    // presynaptic node index < lo_thr: upper window, > hi_thr: lower window
    uint32_t d0, d1, d2;        // shared address of V[node 0] as seen through the upper / middle / lower window (mod 2^32)
    uint32_t k0g, lane;         // first k-row of this warp's slice in the global edge arrays
    uint32_t tid;               // neuron within the tile

    template <int SLOT> __device__ __forceinline__ float f() const {
        constexpr uint32_t o = lay().o_f[SLOT];
        static_assert(o != 0xFFFFFFFFu, "field is not staged for this model (win_field_read)");
        return lds_f32<o>(me);
    }
    template <int SLOT> __device__ __forceinline__ float state() const { return f<SLOT>(); }
    __device__ __forceinline__ float v() const { return lds_f32<lay().o_vwin + kWinRowBytes>(own); }
    __device__ __forceinline__ int lft() const {
        return (int)lds_u32<lay().o_lwin + (lay().lrows == 3 ? kWinRowBytes : 0u)>(own);
    }
    // flags: one byte per neuron; `me` advances 4 bytes per neuron, the byte row needs 1
    __device__ __forceinline__ uint32_t flags() const { return lds_u8<lay().o_flags>(me - 3u * tid); }
    // one word per warp, used late (neurotransmitter release): a plain read-only load, issued early by the compiler
    __device__ __forceinline__ uint32_t spk_prev_word(uint32_t warp_global) const { return __ldg(p.spk_in + (p.own0 >> 5) + warp_global); }
    __device__ __forceinline__ float t_own(int ty) const {
        constexpr uint32_t o = lay().o_twin, tr = lay().trows;
        const uint32_t q = (CHEMG == 1) ? 0u : (uint32_t)ty;
        return lds_f32_rt(own + o + (q * tr + (tr == 3 ? 1u : 0u)) * kWinRowBytes);
    }
    __device__ __forceinline__ float nt(int slot, int ty) const { return lds_f32_rt(me + wp.o_nt[slot][ty]); }
    __device__ __forceinline__ float rc(int slot, int ty) const { return lds_f32_rt(me + wp.o_rc[slot][ty]); }
    __device__ __forceinline__ float rc_state(int slot, int ty) const { return rc(slot, ty); }
    __device__ __forceinline__ uint32_t width() const { return kWinWidth; }
    __device__ __forceinline__ uint32_t col(uint32_t kk) const { return lds_u32<lay().o_col>(edge + kk * 128u); }
    __device__ __forceinline__ float wgt(uint32_t kk) const { return lds_f32<lay().o_wgt>(edge + kk * 128u); }
    __device__ __forceinline__ float *wgt_ptr(uint32_t kk) const { return p.wgt + (size_t)(k0g + kk) * 32u + lane; }
    // STDP: the new weight goes to the stage (this step's gather reads it from there) and to HBM
    __device__ __forceinline__ void wgt_update(uint32_t kk, float x) const {
        asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(edge + kk * 128u), "n"(lay().o_wgt), "f"(x) : "memory");
        *wgt_ptr(kk) = x;
    }
    // generic-proxy stores into a stage that TMA (async proxy) overwrites after the empty barrier
    __device__ __forceinline__ void wgt_updates_done() const { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

    // neighbour gathers: the handle is the shared address of V[j] inside whichever window holds node j
    typedef uint32_t Handle;
    __device__ __forceinline__ Handle gh(uint32_t j) const {
        uint32_t d = d1;
        if (j < lo_thr) d = d0;
        if (j > hi_thr) d = d2;
        return d + j * 4u;
    }
    __device__ __forceinline__ Handle gh_row(int row, uint32_t j) const { return (row == 0 ? d0 : (row == 1 ? d1 : d2)) + j * 4u; }
    __device__ __forceinline__ float gv(Handle h) const { return lds_f32<lay().o_vwin>(h); }
    __device__ __forceinline__ int glft(Handle h) const {
        static_assert(!STDP || lay().lrows == 3, "gathered last_firing_time needs all three window rows");
        return (int)lds_u32<lay().o_lwin>(h);
    }
    __device__ __forceinline__ float gt0(Handle h) const { return lds_f32<lay().o_twin>(h); }
    __device__ __forceinline__ float gt(Handle h, int ty) const { return lds_f32_rt(h + lay().o_twin + (uint32_t)ty * 3u * kWinRowBytes); }
};

// Order in which the CTAs take the tiles.  Whole lattices: ascending, or descending on odd timesteps (the tail of the arrays that
// the previous step left in L2 is what this step reads first).  Row strips: the tiles that import ghosts / export boundary rows
// go FIRST, so a strip's exports of step s leave at the start of step s and its neighbour — which needs them at the start of its
// step s + 1 — finds them long arrived: the NVLink store / fence / flag latency (~10 us) is off the critical path of every step.
__device__ __forceinline__ uint32_t win_tile_of(const StepParams &p, const WinParams &wp, uint32_t seq) {
    const uint32_t nb = wp.first_lo + wp.first_hi;
    if (seq < wp.first_lo) return seq;
    if (seq < nb) return wp.n_tiles - 1u - (seq - wp.first_lo);
    const uint32_t k = seq - nb, n_int = wp.n_tiles - nb;
    return wp.first_lo + (p.reverse ? n_int - 1u - k : k);
}

#ifndef SNN_WIN_PRODUCERS
#define SNN_WIN_PRODUCERS 4
#endif
constexpr int kWinProducers = SNN_WIN_PRODUCERS;   // producer warps per CTA; together they issue the copies of every tile
template <int G> constexpr int win_threads() { return (G * 8 + kWinProducers) * 32; }

template <int MODEL, int CHEMG, bool NTREL, bool STDP, int G>
__global__ void __launch_bounds__(win_threads<G>(), 1)
step_win_kernel(const __grid_constant__ StepParams p, const __grid_constant__ WinParams wp) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr WinLayout L = win_layout(MODEL, CHEMG, NTREL, STDP);
    const uint32_t S = wp.stages;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * wp.stage_bytes);
    uint64_t *empty = full + S;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], kWinProducers);   // every producer warp announces its share of the bytes
            mbar_init(&empty[s], 8);   // the eight warps of the group that consumed the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();   // everything above ran while the previous timestep's kernel was still draining (launch_pdl)
    if (halo_failed(p)) return;
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        p.dbg[0] = (unsigned long long)clock64(); p.dbg[1] = gt;
    }

    if (warp >= G * 8) {
        // ---- producers: kWinProducers warps fill every stage TOGETHER.  A cp.async.bulk costs its warp ~32 ns of issue
        // time whatever its size or the number of active lanes (tools/tma_bench.cu: 1 warp sustains 4.3 TB/s chip-wide with
        // 1 KB copies, 2 warps 7.0 TB/s), and a tile needs ~45 of them: one warp per tile put 1.4 us of serial issue in
        // front of every stage.  Copy k of a tile is issued by warp k % P, lane k / P; each warp announces its own bytes
        // on the stage's full barrier (initialised to P arrivals).
        const uint32_t q = warp - G * 8;
        bool waited_lo = !p.halo[0].active, waited_hi = !p.halo[1].active;
        const uint32_t ty0 = (CHEMG == 1) ? (uint32_t)(__ffs((int)p.nt_used) - 1) : 0u;
        const bool lft_on = wp.lft_copy != 0;
        uint32_t n_t = 0, t_types[kNT] = {0, 0, 0};
        if constexpr (NTREL) {
            for (uint32_t k = 0; k < L.tslots; ++k) {
                const uint32_t ty = (CHEMG == 1) ? ty0 : k;
                if (p.nt_used & (1u << ty)) t_types[n_t++] = k;
            }
        }
        // this lane's copy: a per-tile contiguous operand stream, or window row r of node array arr (0: V, 1:
        // last_firing_time, 2..: t slots)
        const uint32_t idx = q + kWinProducers * lane;
        const bool is_stream = idx < wp.n_streams;
        const uint32_t wv = idx - wp.n_streams, arr = wv / 3u, r = wv - arr * 3u;
        bool is_win = !is_stream && arr < 2u + n_t;
        if (is_win && arr == 1u) is_win = lft_on && (L.lrows == 3 || r == 1u);
        if (is_win && arr >= 2u) is_win = (L.trows == 3 || r == 1u);
        TmaStream my{nullptr, 0u, 0u};
        if (is_stream) my = wp.st[idx];
        uint32_t s = 0, ph = 0;
        for (uint32_t tile_seq = blockIdx.x; tile_seq < wp.n_tiles; tile_seq += gridDim.x) {
            const uint32_t tile = win_tile_of(p, wp, tile_seq);
            mbar_wait_backoff(&empty[s], ph ^ 1u);   // the whole warp probes: no divergent lane-0 loop with 31 lanes parked at a reconvergence barrier
            const uint32_t ts = tile * kWinTile;   // first local neuron of the tile
            // multi-GPU: ghost rows are written by the neighbouring GPU over NVLink; they must have landed before TMA reads
            // them.  Tiles whose upper window reaches below own0 read lower ghosts, tiles whose lower window reaches past the
            // last owned neuron read upper ghosts.
            bool ghosts = false;
            if (!waited_lo && ts <= wp.cols) {
                if (lane == 0) halo_wait(p.halo[0].my_flag, p.halo_epoch, p.halo_done + 2, p.halo_timeout_ns);
                waited_lo = ghosts = true;
            }
            if (!waited_hi && ts + kWinTile + 1u + wp.cols > p.n_neurons) {
                if (lane == 0) halo_wait(p.halo[1].my_flag, p.halo_epoch, p.halo_done + 2, p.halo_timeout_ns);
                waited_hi = ghosts = true;
            }
            unsigned char *dst = smem + (size_t)s * wp.stage_bytes;
            const void *src = nullptr;
            uint32_t bytes = 0;
            if (is_stream) {
                src = my.src + (size_t)tile * my.bytes_per_tile;
                dst += my.smem_off;
                bytes = my.bytes_per_tile;
            } else if (is_win) {
                // window row r: node indices [a4, a4 + kWinRowElems) with a4 = floor4(own0 + ts + (r - 1) * cols - 1), clamped
                // to the allocation
                const int64_t a = (int64_t)p.own0 + ts + ((int64_t)r - 1) * wp.cols - 1;
                const int64_t a4 = a & ~(int64_t)3;
                const int64_t b = a4 < 0 ? 0 : a4;
                const int64_t e = (a4 + kWinRowElems) > (int64_t)wp.node_cap ? (int64_t)wp.node_cap : (a4 + kWinRowElems);
                if (e > b) {
                    bytes = (uint32_t)(e - b) * 4u;
                    const uint32_t so = (uint32_t)(b - a4) * 4u;   // bytes skipped at the front of the window row
                    if (arr == 0u) {
                        dst += L.o_vwin + r * kWinRowBytes + so;
                        src = p.v_in + b;
                    } else if (arr == 1u) {
                        dst += L.o_lwin + (L.lrows == 3 ? r : 0u) * kWinRowBytes + so;
                        src = p.lft_in + b;
                    } else {
                        const uint32_t k = t_types[arr - 2u];
                        const uint32_t ty = (CHEMG == 1) ? ty0 : k;
                        dst += L.o_twin + (k * L.trows + (L.trows == 3 ? r : 0u)) * kWinRowBytes + so;
                        src = p.t_in + (size_t)ty * p.t_stride + b;
                    }
                }
            }
            const uint32_t tx = __reduce_add_sync(0xffffffffu, bytes);
            if (lane == 0) mbar_arrive_expect_tx(&full[s], tx);
            __syncwarp();
            if (ghosts) fence_proxy_async();   // lane 0's acquire of the arrival flag, then every lane's TMA reads of the ghosts
            if (bytes) tma_bulk_g2s(dst, src, bytes, &full[s]);
            if (++s == S) { s = 0; ph ^= 1u; }
        }
        return;
    }

    // ---- consumers: group g takes the CTA's tiles g, g + G, ...; 8 warps x 32 neurons per tile --------------------
    const uint32_t g = warp >> 3, wg = warp & 7u, tid = threadIdx.x & 255u;
    const uint32_t smem_s = smem_u32(smem);
    uint32_t s = g % S, ph = (g / S) & 1u;
    for (uint32_t tile_seq = blockIdx.x + g * gridDim.x; tile_seq < wp.n_tiles; tile_seq += G * gridDim.x) {
        const uint32_t tile = win_tile_of(p, wp, tile_seq);
        const uint32_t warp_global = tile * 8u + wg;
        const uint32_t ln = warp_global * 32u + lane;
        const bool active = warp_global * 32u < p.n_neurons;
        const bool valid = ln < p.n_neurons;
        const uint32_t lnc = valid ? ln : p.n_neurons - 1;
        bool export_lo = false, export_hi = false;
        // only the few tiles at the two ends of a strip touch a halo (CTA-uniform test): everything else runs the plain step
        const bool boundary = tile < wp.bnd_lo || tile + wp.bnd_hi >= wp.n_tiles;
        if (boundary) halo_export_flags(p, ln, valid, export_lo, export_hi);
        // A group returns to a stage only every lcm(S, G) / S rounds, and a parity wait cannot tell "round k landed" from
        // "round k - 1 is still landing" (two phases apart).  Bulk copies of different stages complete out of order, so
        // first make sure that round k - 1 of this stage has been consumed (the empty barrier; a producer that issued this
        // warp's previous tile has already seen round k - 2 consumed, so this wait is unambiguous), then wait for the data.
        if (S % G != 0u) mbar_wait(&empty[s], ph ^ 1u);
        mbar_wait(&full[s], ph);
        if (active) {
            const uint32_t ts_node = p.own0 + tile * kWinTile;
            const uint32_t st = smem_s + s * wp.stage_bytes;
            // shared address of V[node 0] through window row r: st + r * row_bytes - 4 * floor4(ts_node + (r - 1) * cols - 1)
            const int32_t a0 = (int32_t)ts_node - (int32_t)wp.cols - 1, a1 = (int32_t)ts_node - 1, a2 = (int32_t)ts_node + (int32_t)wp.cols - 1;
            const uint32_t d0 = st + 0u * kWinRowBytes - (uint32_t)(a0 & ~3) * 4u;
            const uint32_t d1 = st + 1u * kWinRowBytes - (uint32_t)(a1 & ~3) * 4u;
            const uint32_t d2 = st + 2u * kWinRowBytes - (uint32_t)(a2 & ~3) * 4u;
            // this neuron inside a row-relative window: node (ts_node + tid) sits (ts_node + tid - floor4(a1)) elements in
            const uint32_t own = st + ((uint32_t)((int32_t)ts_node - (a1 & ~3)) + tid) * 4u;
            const WinSrc<MODEL, CHEMG, NTREL, STDP> src{p, wp, st + tid * 4u, own, st + (wg * kWinWidth * 32u + lane) * 4u,
                                                         ts_node ? ts_node - 1u : 0u, ts_node + kWinTile, d0, d1, d2,
                                                         warp_global * kWinWidth, lane, tid};
            neuron_step<MODEL, CHEMG, NTREL, STDP, false>(p, src, warp_global, lane, ln, lnc, valid, export_lo, export_hi);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (boundary && active) halo_publish(p, warp_global, lane);
        s += G;
        if (s >= S) { s -= S; ph ^= 1u; }
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        p.dbg[2] = (unsigned long long)clock64(); p.dbg[3] = gt;
    }
}

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template <int MODEL, int CHEMG, bool NTREL>
static cudaError_t launch_win_3(const StepParams &p, const WinParams &wp, bool stdp, unsigned grid, size_t smem, cudaStream_t s) {
    constexpr int G = win_groups(MODEL, CHEMG);
    if (stdp) {
        auto k = step_win_kernel<MODEL, CHEMG, NTREL, true, G>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_pdl<PDL_STEP>(pdl_ok(p), k, dim3(grid), dim3(win_threads<G>()), smem, s, p, wp);
    } else {
        auto k = step_win_kernel<MODEL, CHEMG, NTREL, false, G>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_pdl<PDL_STEP>(pdl_ok(p), k, dim3(grid), dim3(win_threads<G>()), smem, s, p, wp);
    }
    return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_win_model(const StepParams &p, const WinParams &wp, int chemg, bool ntrel, bool stdp, unsigned grid,
                                    size_t smem, cudaStream_t s) {
    if (chemg == 1) return launch_win_3<MODEL, 1, true>(p, wp, stdp, grid, smem, s);
    if (chemg == 3) return launch_win_3<MODEL, 3, true>(p, wp, stdp, grid, smem, s);
    if (ntrel) return launch_win_3<MODEL, 0, true>(p, wp, stdp, grid, smem, s);
    return launch_win_3<MODEL, 0, false>(p, wp, stdp, grid, smem, s);
}

cudaError_t launch_step_win(const StepParams &p, const WinParams &wp, int model, int chemg, bool ntrel, bool stdp, unsigned grid,
                            cudaStream_t s) {
    if (p.n_neurons == 0) return cudaSuccess;
    const size_t smem = (size_t)wp.stages * wp.stage_bytes + 2 * wp.stages * sizeof(uint64_t);
    switch (model) {
    case SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE: return launch_win_model<SNN_MODEL_LEAKY_INTEGRATE_AND_FIRE>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE: return launch_win_model<SNN_MODEL_QUADRATIC_INTEGRATE_AND_FIRE>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE: return launch_win_model<SNN_MODEL_ADAPTIVE_LEAKY_INTEGRATE_AND_FIRE>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE: return launch_win_model<SNN_MODEL_ADAPTIVE_EXP_LEAKY_INTEGRATE_AND_FIRE>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_IZHIKEVICH: return launch_win_model<SNN_MODEL_IZHIKEVICH>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_LEAKY_IZHIKEVICH: return launch_win_model<SNN_MODEL_LEAKY_IZHIKEVICH>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE: return launch_win_model<SNN_MODEL_SIMPLE_LEAKY_INTEGRATE_AND_FIRE>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_HODGKIN_HUXLEY: return launch_win_model<SNN_MODEL_HODGKIN_HUXLEY>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    case SNN_MODEL_BCM_IZHIKEVICH: return launch_win_model<SNN_MODEL_BCM_IZHIKEVICH>(p, wp, chemg, ntrel, stdp, grid, smem, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace snn
