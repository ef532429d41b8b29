// engine.cu — host runtime: device memory layout, named-field I/O, graph ingestion, the step loop.
#include "engine.h"
#include <chrono>
#include <cmath>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cstdio>
#include <cstring>
#include <functional>

namespace snn {

#define CK(call, status)                                                          \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return cuda_fail(_e, (status), #call);             \
    } while (0)

static inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// Host -> device copy that has LANDED when it returns.  A plain cudaMemcpy from pageable memory may return once the data sits in
// the driver's staging buffer; the engine's kernels run on a non-blocking stream that does not wait for the legacy stream's DMA,
// so a kernel launched right after could read an array whose tail has not arrived yet (seen as garbage in the last slices of a
// rebuilt graph).  Copy on the engine's stream and wait for it.
static cudaError_t h2d_sync(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    return e != cudaSuccess ? e : cudaStreamSynchronize(s);
}

template <class T>
static cudaError_t dev_alloc(T **p, size_t n) {
    *p = nullptr;
    return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
}

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

Engine::Engine(int model_, int ntk_, int rck_, int train_kind_, int refract_, int device_)
    : model(model_), ntk(ntk_), rck(rck_), train_kind(train_kind_), refract(refract_), device(device_) {
    // handles of one process get distinct (but reproducible) default keys: two networks built the same way must not replay the
    // same Poisson noise; SNN_OPT_RNG_SEED pins the key
    static std::atomic<uint64_t> serial{0};
    seed = splitmix64(0x5EED5EEDull + serial.fetch_add(1));
}

Engine::~Engine() {
    if (device >= 0) cudaSetDevice(device);
    for (int d = 0; d < 2; ++d) {
        if (peer_slab_[d]) cudaIpcCloseMemHandle(peer_slab_[d]);
        if (peer_flags_[d]) cudaIpcCloseMemHandle(peer_flags_[d]);
    }
    for (auto &G : gpeers_)
        if (G.ipc) { if (G.slab) cudaIpcCloseMemHandle(G.slab); if (G.flags) cudaIpcCloseMemHandle(G.flags); }
    if (d_gpeers_) cudaFree(d_gpeers_);
    if (d_gexp_off_) cudaFree(d_gexp_off_);
    if (d_gexp_ent_) cudaFree(d_gexp_ent_);
    if (d_gslice_) cudaFree(d_gslice_);
    free_device();
    if (flags_) cudaFree(flags_);
    if (halo_done_) cudaFree(halo_done_);
    if (multi_barrier_) cudaFree(multi_barrier_);
    if (rs_tab_) cudaFree(rs_tab_);
    if (rnet_tab_) cudaFree(rnet_tab_);
    if (wide_scratch_) cudaFree(wide_scratch_);
    if (scratch_) cudaFree(scratch_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (stream_) cudaStreamDestroy(stream_);
}

int Engine::cuda_fail(cudaError_t e, int status, const char *what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    last_error = buf;
    cudaGetLastError();
    return status;
}

int Engine::init() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SNN_GPU_GET_DEVICE_FAILURE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) return fail(SNN_GPU_GET_DEVICE_FAILURE, "cudaGetDevice failed");
    }
    if (device >= count) return fail(SNN_GPU_GET_DEVICE_FAILURE, "device ordinal out of range");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    CK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), SNN_GPU_QUEUE_FAILURE);
    CK(cudaEventCreate(&ev0_), SNN_GPU_QUEUE_FAILURE);
    CK(cudaEventCreate(&ev1_), SNN_GPU_QUEUE_FAILURE);
    // [0],[1] step arrivals from rank-1 / rank+1, [2],[3] edge-kernel arrivals, [8 + r] general-graph partition: rank r completed a step
    CK(dev_alloc(&flags_, 8 + kMaxRanks), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(cudaMemsetAsync(flags_, 0, (8 + kMaxRanks) * sizeof(unsigned long long), stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(dev_alloc(&multi_barrier_, 1), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&halo_done_, 8), SNN_GPU_BUFFER_CREATE_ERROR);  // [4] ns / 16 spent in halo waits, [5] waits that spun (diagnostics); [0],[1] completion counters, [2] halo time-out flag, [3] per-edge kernel CTAs done
    CK(cudaMemsetAsync(halo_done_, 0, 8 * sizeof(unsigned int), stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

int Engine::ensure_scratch(size_t bytes) {
    if (bytes <= scratch_bytes_) return SNN_OK;
    if (scratch_) cudaFree(scratch_);
    scratch_ = nullptr; scratch_bytes_ = 0;
    CK(cudaMalloc(&scratch_, bytes), SNN_GPU_BUFFER_CREATE_ERROR);
    scratch_bytes_ = bytes;
    return SNN_OK;
}

Lat *Engine::find(uint64_t id) {
    for (auto &L : lats_) if (L.id == id) return &L;
    return nullptr;
}
const Lat *Engine::find(uint64_t id) const {
    for (auto &L : lats_) if (L.id == id) return &L;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
void Engine::compute_layout() {
    // canonical order: neuron lattices by ascending id, then spike-train lattices by ascending id
    std::stable_sort(lats_.begin(), lats_.end(), [](const Lat &a, const Lat &b) {
        if (a.is_train != b.is_train) return !a.is_train;
        return a.id < b.id;
    });
    n_neurons = n_trains = 0;
    for (auto &L : lats_) {
        if (L.is_train) { L.off = n_trains; n_trains += L.n; }
        else { L.off = n_neurons; n_neurons += L.n; }
    }
    // row strips: `halo_` ghost rows on each side that has a neighbour; general-graph partition: the ghost lists themselves
    const uint32_t halo_lo = gpart_ ? (uint32_t)g_lo_.size() : ((part_world > 1 && part_rank > 0) ? halo_ : 0);
    const uint32_t halo_hi = gpart_ ? (uint32_t)g_hi_.size() : ((part_world > 1 && part_rank < part_world - 1) ? halo_ : 0);
    own0_ = (uint32_t)round_up(halo_lo, 32);
    train0_ = (uint32_t)round_up(own0_ + n_neurons, 32);
    // upper ghosts follow the owned neurons without a gap: the stencil generator addresses the row below the strip as
    // own0 + rows_local * cols + c
    ghost_hi0_ = (part_world > 1) ? own0_ + (uint32_t)n_neurons : train0_;
    n_nodes_ = (part_world > 1) ? ghost_hi0_ + halo_hi : train0_ + (uint32_t)n_trains;
    // capacities cover whole 256-neuron tiles so that the TMA-staged kernel may copy full tiles past the last neuron
    node_cap_ = round_up(std::max<uint64_t>(n_nodes_, 1), 32) + kTmaTile + 32;
    neuron_cap_ = round_up(std::max<uint64_t>(n_neurons, 1), kTmaTile);
    train_cap_ = round_up(std::max<uint64_t>(n_trains, 1), 32);
}

void Engine::free_device() {
    auto fr = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
    if (slab_) cudaFree(slab_);
    slab_ = nullptr;
    V_[0] = V_[1] = nullptr; LFT_[0] = LFT_[1] = nullptr; T_[0] = T_[1] = nullptr;
    fr(SPK_[0]); fr(SPK_[1]); fr(was_inc_);
    node_flags_ = nullptr;   // lives in the slab
    for (auto &p : NT_) fr(p);
    for (auto &p : F_) fr(p);
    for (auto &p : RC_) fr(p);
    for (auto &p : TF_) fr(p);
    fr(ft_off_); fr(ft_); fr(d_lat_);
    fr(slice_off_); fr(col_); fr(wgt_);
    free_reward_arrays();
    chem_alloc_ = false;
    graph_dirty_ = true;
    grid_fast_ = false;
}

static cudaError_t fill_f32(float *p, float v, uint64_t n, cudaStream_t s) {
    uint32_t b; memcpy(&b, &v, 4);
    return launch_fill_u32((uint32_t *)p, b, n, s);
}

int Engine::alloc_device() {
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    // slab: V[2], LFT[2] (+ T[2] once chemistry is touched) in one allocation so that a neighbouring rank can
    // map all halo-visible arrays with one CUDA IPC handle
    const size_t vb = round_up(node_cap_ * 4, 256);
    slab_off_v_[0] = 0; slab_off_v_[1] = vb; slab_off_lft_[0] = 2 * vb; slab_off_lft_[1] = 3 * vb;
    // node flags ride in the slab too: a neighbouring strip reads the neurotransmitter type sets of my boundary rows at attach time
    const size_t fbytes = round_up(node_cap_, 256);
    slab_off_flags_ = 4 * vb;
    slab_off_t_[0] = 4 * vb + fbytes; slab_off_t_[1] = slab_off_t_[0] + round_up(node_cap_ * 4 * kNT, 256);
    slab_bytes_ = 4 * vb + fbytes;  // T appended by ensure_chem (which reallocates the slab)
    CK(cudaMalloc(&slab_, slab_bytes_), SNN_GPU_BUFFER_CREATE_ERROR);
    node_flags_ = (uint8_t *)slab_ + slab_off_flags_;
    for (int k = 0; k < 2; ++k) {
        V_[k] = (float *)((char *)slab_ + slab_off_v_[k]);
        LFT_[k] = (int *)((char *)slab_ + slab_off_lft_[k]);
        CK(fill_f32(V_[k], 0.f, node_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        CK(launch_fill_u32((uint32_t *)LFT_[k], 0xFFFFFFFFu, node_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        CK(dev_alloc(&SPK_[k], node_cap_ / 32 + 1), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(launch_fill_u32(SPK_[k], 0u, node_cap_ / 32 + 1, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    }
    CK(cudaMemsetAsync(node_flags_, 0, node_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    h_node_flags_.assign(node_cap_, 0);
    flags_cache_valid_ = false;
    // neuron fields used by this model
    for (int i = 0; i < kNumNeuronFields; ++i) {
        const FieldDef &fd = kNeuronFields[i];
        if (fd.kind != FK_NEURON_DEV || !(fd.models & SNN_M(model))) continue;
        CK(dev_alloc(&F_[fd.slot], neuron_cap_), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(fill_f32(F_[fd.slot], neuron_default(model, fd.kind, fd.slot), neuron_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    }
    if (model == SNN_MODEL_HODGKIN_HUXLEY) {
        CK(dev_alloc(&was_inc_, neuron_cap_ / 32 + 1), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(launch_fill_u32(was_inc_, 0u, neuron_cap_ / 32 + 1, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    }
    // neuron voltage default
    if (n_neurons)
        CK(fill_f32(V_[0] + own0_, neuron_default(model, FK_V, 0), n_neurons, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    if (n_neurons) CK(fill_f32(V_[1] + own0_, neuron_default(model, FK_V, 0), n_neurons, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    // trains
    if (n_trains) {
        for (int s = 0; s < TF_COUNT; ++s) {
            CK(dev_alloc(&TF_[s], train_cap_), SNN_GPU_BUFFER_CREATE_ERROR);
            CK(fill_f32(TF_[s], train_default(s), train_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        }
        std::vector<uint64_t> zero(n_trains + 1, 0);
        CK(dev_alloc(&ft_off_, n_trains + 1), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(cudaMemcpyAsync(ft_off_, zero.data(), (n_trains + 1) * 8, cudaMemcpyHostToDevice, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        CK(dev_alloc(&ft_, 1), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    }
    CK(dev_alloc(&d_lat_, kMaxLattices), SNN_GPU_BUFFER_CREATE_ERROR);
    for (auto &L : lats_) {
        if (L.is_train) continue;
        if (L.cold[0].size() != L.n) L.cold[0].assign(L.n, neuron_default(model, FK_NEURON_COLD, 0));
        if (L.cold[1].size() != L.n) L.cold[1].assign(L.n, neuron_default(model, FK_NEURON_COLD, 1));
    }
    cur_ = 0; lft_loc_ = 0;
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

int Engine::ensure_chem() {
    if (chem_alloc_) return SNN_OK;
    if (peer_slab_[0] || peer_slab_[1] || layout_frozen_) return fail(SNN_UNSUPPORTED, "chemical fields must be set before the halo slab is exported");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    // grow the slab to hold T[2]
    const size_t tb = round_up(node_cap_ * 4 * kNT, 256);
    const size_t new_bytes = slab_off_t_[0] + 2 * tb;
    void *ns = nullptr;
    CK(cudaMalloc(&ns, new_bytes), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(cudaMemcpyAsync(ns, slab_, slab_bytes_, cudaMemcpyDeviceToDevice, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(cudaMemsetAsync((char *)ns + slab_off_t_[0], 0, 2 * tb, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    cudaFree(slab_);
    slab_ = ns; slab_bytes_ = new_bytes;
    for (int k = 0; k < 2; ++k) {
        V_[k] = (float *)((char *)slab_ + slab_off_v_[k]);
        LFT_[k] = (int *)((char *)slab_ + slab_off_lft_[k]);
        T_[k] = (float *)((char *)slab_ + slab_off_t_[k]);
    }
    node_flags_ = (uint8_t *)slab_ + slab_off_flags_;
    for (int s = 0; s < NTF_COUNT; ++s) {
        CK(dev_alloc(&NT_[s], node_cap_ * kNT), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(fill_f32(NT_[s], nt_default(ntk, s), node_cap_ * kNT, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    }
    for (int s = 0; s < RCF_COUNT; ++s) {
        CK(dev_alloc(&RC_[s], neuron_cap_ * kNT), SNN_GPU_BUFFER_CREATE_ERROR);
        for (int ty = 0; ty < kNT; ++ty)
            CK(fill_f32(RC_[s] + (size_t)ty * neuron_cap_, rc_default(rck, ty, s), neuron_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    }
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    chem_alloc_ = true;
    return SNN_OK;
}

int Engine::set_partition(uint32_t rows_g, uint32_t cols, int rank, int world) {
    if (!lats_.empty()) return fail(SNN_INVALID_ARGUMENT, "set_partition must precede add_lattice");
    if (world < 1 || rank < 0 || rank >= world) return fail(SNN_INVALID_ARGUMENT, "bad partition rank/world");
    part_rank = rank; part_world = world; rows_global = rows_g;
    row0_global = snn_partition_begin(rows_g, world, rank);
    halo_ = world > 1 ? cols : 0;  // one grid row; grows with the stencil radius (connect_grid)
    return SNN_OK;
}

struct FieldSnap { std::string name; int dtype; std::vector<uint8_t> data; uint64_t count; };
struct LatSnap { uint64_t id; std::vector<FieldSnap> fields; };

int Engine::add_lattice(uint64_t id, uint32_t rows, uint32_t cols, bool is_train, bool is_reward) {
    // LatticeNetwork::add_lattice / add_spike_train_lattice, neuron/mod.rs:1663-1698
    if (find(id)) return fail(SNN_NET_GRAPH_ID_ALREADY_PRESENT, "Graph id already present in network, id: " + std::to_string(id));
    int n_neuron_lat = 0;
    for (auto &L : lats_) if (!L.is_train) n_neuron_lat++;
    if (!is_train && n_neuron_lat >= kMaxLattices) return fail(SNN_UNSUPPORTED, "too many lattices");
    if (part_world > 1 && (!lats_.empty() || is_train)) return fail(SNN_UNSUPPORTED, "a partitioned handle holds exactly one neuron lattice");
    if (is_reward && (is_train || part_world > 1 || reward_mode || bcm_mode))
        return fail(SNN_UNSUPPORTED, "reward-modulated lattices of a network: single-GPU network handles only");
    Lat nl;
    nl.is_reward = is_reward;   // RewardModulatedLatticeNetwork::add_reward_modulated_lattice, neuron/mod.rs:3615-3634
    nl.id = id; nl.rows = rows; nl.cols = cols; nl.n = (uint64_t)rows * cols; nl.is_train = is_train;
    if ((is_train ? n_trains : n_neurons) + nl.n + 64ull * (part_world > 1 ? cols : 0) > kMaxNodes)
        return fail(SNN_UNSUPPORTED, "lattice too large for one device (node index is 28 bits)");
    return relayout_add(nl);
}

int Engine::relayout_add(const Lat &nl) {
    // snapshot every existing field through the public get path, rebuild the device layout, restore.
    // Lattices are added at set-up time; the single-lattice (large) case never has anything to snapshot.
    if (dev_weights_newer_) { int r = sync_weights_to_host(); if (r) return r; }
    std::vector<LatSnap> snaps;
    const bool had_chem = chem_alloc_;
    for (auto &L : lats_) {
        LatSnap s; s.id = L.id;
        uint32_t cnt = 0; field_count(L.id, &cnt);
        for (uint32_t i = 0; i < cnt; ++i) {
            const char *name; int32_t dt; uint32_t per;
            field_info(L.id, i, &name, &dt, &per);
            FieldDef fd;
            if (lookup_field(L, name, &fd)) continue;
            const bool is_chem = fd.kind == FK_NT_FLAGS || fd.kind == FK_NT_T || fd.kind == FK_NT || fd.kind == FK_RC_FLAGS || fd.kind == FK_RC;
            if (is_chem && !had_chem) continue;
            FieldSnap f; f.name = name; f.dtype = dt; f.count = L.n * per; f.data.resize(std::max<uint64_t>(f.count, 1) * 4);
            int r = get_field(L.id, name, f.data.data(), f.count, dt);
            if (r) return r;
            s.fields.push_back(std::move(f));
        }
        snaps.push_back(std::move(s));
    }
    const uint64_t saved_clock = internal_clock;
    free_device();
    lats_.push_back(nl);
    compute_layout();
    int r = alloc_device();
    if (r) return r;
    if (had_chem) { r = ensure_chem(); if (r) return r; }
    for (auto &s : snaps)
        for (auto &f : s.fields) {
            r = set_field(s.id, f.name.c_str(), f.data.data(), f.count, f.dtype);
            if (r) return r;
        }
    // re-upload preset firing times
    for (auto &L : lats_)
        if (L.is_train && !L.ft_off.empty()) {
            std::vector<uint64_t> off = L.ft_off; std::vector<float> t = L.ft;
            r = set_preset_firing_times(L.id, off.data(), t.data(), L.n, t.size());
            if (r) return r;
        }
    internal_clock = saved_clock;
    graph_dirty_ = true;
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// field directory
// ------------------------------------------------------------------------------------------------
static bool field_visible(const FieldDef &fd, int kind_bit, int ntk_, int rck_) {
    if (!(fd.models & SNN_M(kind_bit))) return false;
    if (fd.kind == FK_NT && fd.aux && !(fd.aux & SNN_M(ntk_))) return false;
    (void)rck_;
    return true;
}

int Engine::field_count(uint64_t id, uint32_t *count) const {
    const Lat *L = find(id);
    if (!L) return SNN_NET_ID_NOT_FOUND_IN_LATTICES;
    uint32_t c = 0;
    if (L->is_train) {
        for (int i = 0; i < kNumTrainFields; ++i) if (field_visible(kTrainFields[i], train_kind, ntk, rck)) c++;
    } else {
        for (int i = 0; i < kNumNeuronFields; ++i) if (field_visible(kNeuronFields[i], model, ntk, rck)) c++;
        for (int ty = 0; ty < kNT; ++ty)
            for (int i = 0; i < kNumRcFields; ++i) {
                if (kRcFields[i].nmda_only && ty != SNN_NT_NMDA) continue;
                if (kRcFields[i].kinetics_mask && !(kRcFields[i].kinetics_mask & SNN_M(rck))) continue;
                c++;
            }
    }
    *count = c;
    return SNN_OK;
}

static thread_local char g_name_buf[128];

int Engine::field_info(uint64_t id, uint32_t index, const char **name, int32_t *dtype, uint32_t *per) const {
    const Lat *L = find(id);
    if (!L) return SNN_NET_ID_NOT_FOUND_IN_LATTICES;
    uint32_t c = 0;
    if (L->is_train) {
        for (int i = 0; i < kNumTrainFields; ++i)
            if (field_visible(kTrainFields[i], train_kind, ntk, rck)) {
                if (c == index) { *name = kTrainFields[i].name; *dtype = kTrainFields[i].dtype; *per = kTrainFields[i].per; return SNN_OK; }
                c++;
            }
    } else {
        for (int i = 0; i < kNumNeuronFields; ++i)
            if (field_visible(kNeuronFields[i], model, ntk, rck)) {
                if (c == index) { *name = kNeuronFields[i].name; *dtype = kNeuronFields[i].dtype; *per = kNeuronFields[i].per; return SNN_OK; }
                c++;
            }
        for (int ty = 0; ty < kNT; ++ty)
            for (int i = 0; i < kNumRcFields; ++i) {
                if (kRcFields[i].nmda_only && ty != SNN_NT_NMDA) continue;
                if (kRcFields[i].kinetics_mask && !(kRcFields[i].kinetics_mask & SNN_M(rck))) continue;
                if (c == index) {
                    snprintf(g_name_buf, sizeof g_name_buf, "receptors$%s%s", kRcTypeNames[ty], kRcFields[i].suffix);
                    *name = g_name_buf; *dtype = SNN_F32; *per = 1;
                    return SNN_OK;
                }
                c++;
            }
    }
    return SNN_INVALID_ARGUMENT;
}

int Engine::lookup_field(const Lat &L, const char *name, FieldDef *out) const {
    if (L.is_train) {
        for (int i = 0; i < kNumTrainFields; ++i)
            if (field_visible(kTrainFields[i], train_kind, ntk, rck) && !strcmp(kTrainFields[i].name, name)) { *out = kTrainFields[i]; return SNN_OK; }
        return SNN_UNKNOWN_FIELD;
    }
    for (int i = 0; i < kNumNeuronFields; ++i)
        if (field_visible(kNeuronFields[i], model, ntk, rck) && !strcmp(kNeuronFields[i].name, name)) { *out = kNeuronFields[i]; return SNN_OK; }
    if (!strncmp(name, "receptors$", 10)) {
        for (int ty = 0; ty < kNT; ++ty) {
            const size_t tl = strlen(kRcTypeNames[ty]);
            if (strncmp(name + 10, kRcTypeNames[ty], tl)) continue;
            for (int i = 0; i < kNumRcFields; ++i) {
                if (kRcFields[i].nmda_only && ty != SNN_NT_NMDA) continue;
                if (kRcFields[i].kinetics_mask && !(kRcFields[i].kinetics_mask & SNN_M(rck))) continue;
                if (!strcmp(name + 10 + tl, kRcFields[i].suffix)) {
                    *out = FieldDef{nullptr, SNN_F32, 1, FK_RC, kRcFields[i].slot, ty, kAllModels};
                    return SNN_OK;
                }
            }
        }
    }
    return SNN_UNKNOWN_FIELD;
}

// ------------------------------------------------------------------------------------------------
// field I/O
// ------------------------------------------------------------------------------------------------
int Engine::set_bits(uint32_t *words, const uint32_t *host_u32, uint64_t n, uint64_t bit0) {
    int r = ensure_scratch(n * 4);
    if (r) return r;
    CK(cudaMemcpyAsync(scratch_, host_u32, n * 4, cudaMemcpyHostToDevice, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(launch_bits_from_u32((const uint32_t *)scratch_, words, n, bit0, stream_), SNN_GPU_QUEUE_FAILURE);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

int Engine::get_bits(const uint32_t *words, uint32_t *host_u32, uint64_t n, uint64_t bit0) {
    int r = ensure_scratch(n * 4);
    if (r) return r;
    CK(launch_u32_from_bits(words, (uint32_t *)scratch_, n, bit0, stream_), SNN_GPU_QUEUE_FAILURE);
    CK(cudaMemcpyAsync(host_u32, scratch_, n * 4, cudaMemcpyDeviceToHost, stream_), SNN_GPU_BUFFER_READ_ERROR);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

int Engine::field_io(Lat &L, const FieldDef &fd, void *data, uint64_t count, bool set) {
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    const uint64_t n = L.n;
    if (n == 0) return SNN_OK;
    const uint32_t no = node_off(L);
    const cudaMemcpyKind dir = set ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    const int wr = set ? SNN_GPU_BUFFER_WRITE_ERROR : SNN_GPU_BUFFER_READ_ERROR;
    auto plain = [&](void *dev, uint64_t elems) -> int {
        if (set) CK(cudaMemcpyAsync(dev, data, elems * 4, dir, stream_), wr);
        else CK(cudaMemcpyAsync(data, dev, elems * 4, dir, stream_), wr);
        CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
        return SNN_OK;
    };
    // neuron-major [n][3] host <-> type-major device [3][stride] at element offset `o`
    auto typed = [&](float *dev_tm, uint64_t stride, uint64_t o) -> int {
        int r = ensure_scratch(n * kNT * 4);
        if (r) return r;
        if (set) {
            CK(cudaMemcpyAsync(scratch_, data, n * kNT * 4, cudaMemcpyHostToDevice, stream_), wr);
            CK(launch_transpose_in((const float *)scratch_, dev_tm + o, n, stride, stream_), SNN_GPU_QUEUE_FAILURE);
        } else {
            CK(launch_transpose_out(dev_tm + o, (float *)scratch_, n, stride, stream_), SNN_GPU_QUEUE_FAILURE);
            CK(cudaMemcpyAsync(data, scratch_, n * kNT * 4, cudaMemcpyDeviceToHost, stream_), wr);
        }
        CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
        return SNN_OK;
    };
    switch (fd.kind) {
    case FK_NEURON_DEV:
        return plain(F_[fd.slot] + L.off, n);
    case FK_NEURON_COLD:
        if (set) memcpy(L.cold[fd.slot].data(), data, n * 4); else memcpy(data, L.cold[fd.slot].data(), n * 4);
        return SNN_OK;
    case FK_V:
        return plain(V_[cur_] + no, n);
    case FK_LFT: {
        int r = plain(LFT_[lft_loc_] + no, n);
        if (r) return r;
        if (set && part_world > 1)  // keep both parities coherent for the halo protocol
            CK(cudaMemcpyAsync(LFT_[lft_loc_ ^ 1] + no, LFT_[lft_loc_] + no, n * 4, cudaMemcpyDeviceToDevice, stream_), wr);
            CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
        return SNN_OK;
    }
    case FK_SPIKING:
        return set ? set_bits(SPK_[cur_], (const uint32_t *)data, n, no) : get_bits(SPK_[cur_], (uint32_t *)data, n, no);
    case FK_WAS_INC:
        return set ? set_bits(was_inc_, (const uint32_t *)data, n, L.off) : get_bits(was_inc_, (uint32_t *)data, n, L.off);
    case FK_NT_FLAGS:
    case FK_RC_FLAGS: {
        const int shift = fd.kind == FK_NT_FLAGS ? 0 : 4;
        uint32_t *h = (uint32_t *)data;
        if (set) {
            int r = ensure_chem();
            if (r) return r;
            bool changed = false;
            if (n >= (1u << 16)) {
                // large lattices: pack on the device (a host loop over 3 n words costs more than the copy), then refresh
                // the host mirror only if something changed
                r = ensure_scratch(n * kNT * 4 + 16);
                if (r) return r;
                unsigned int *d_changed = (unsigned int *)((char *)scratch_ + n * kNT * 4);
                CK(cudaMemsetAsync(d_changed, 0, 4, stream_), wr);
                CK(cudaMemcpyAsync(scratch_, h, n * kNT * 4, cudaMemcpyHostToDevice, stream_), wr);
                CK(launch_pack_flags((const uint32_t *)scratch_, node_flags_ + no, n, shift, d_changed, stream_), SNN_GPU_QUEUE_FAILURE);
                unsigned int hc = 0;
                CK(cudaMemcpyAsync(&hc, d_changed, 4, cudaMemcpyDeviceToHost, stream_), wr);
                CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
                changed = hc != 0;
                if (changed) CK(cudaMemcpy(h_node_flags_.data() + no, node_flags_ + no, n, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
            } else {
                for (uint64_t i = 0; i < n; ++i) {
                    uint8_t m = 0;
                    for (int ty = 0; ty < kNT; ++ty) if (h[i * kNT + ty]) m |= (uint8_t)(1u << ty);
                    uint8_t &f = h_node_flags_[no + i];
                    const uint8_t nf = (uint8_t)((f & ~(0xFu << shift)) | (m << shift));
                    changed |= nf != f;
                    f = nf;
                }
                if (changed) {
                    CK(cudaMemcpyAsync(node_flags_ + no, h_node_flags_.data() + no, n, cudaMemcpyHostToDevice, stream_), wr);
                    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
                }
            }
            if (!changed) return SNN_OK;
            flags_cache_valid_ = false;
            if (shift == 0) {
                // the presynaptic type bits are baked into the edge words: re-encode at the next run
                if (dev_weights_newer_) { r = sync_weights_to_host(); if (r) return r; }
                graph_dirty_ = true;
            }
        } else {
            for (uint64_t i = 0; i < n; ++i)
                for (int ty = 0; ty < kNT; ++ty) h[i * kNT + ty] = (h_node_flags_[no + i] >> (shift + ty)) & 1u;
        }
        return SNN_OK;
    }
    case FK_NT_T: {
        if (set) { int r = ensure_chem(); if (r) return r; }
        if (!chem_alloc_) { memset(data, 0, n * kNT * 4); return SNN_OK; }
        return typed(T_[cur_], node_cap_, no);
    }
    case FK_NT: {
        if (set) { int r = ensure_chem(); if (r) return r; }
        if (!chem_alloc_) {
            float *o = (float *)data;
            for (uint64_t i = 0; i < n * kNT; ++i) o[i] = nt_default(ntk, fd.slot);
            return SNN_OK;
        }
        return typed(NT_[fd.slot], node_cap_, no);
    }
    case FK_RC: {
        if (set) { int r = ensure_chem(); if (r) return r; }
        if (!chem_alloc_) {
            float *o = (float *)data;
            for (uint64_t i = 0; i < n; ++i) o[i] = rc_default(rck, fd.aux, fd.slot);
            return SNN_OK;
        }
        return plain(RC_[fd.slot] + (size_t)fd.aux * neuron_cap_ + L.off, n);
    }
    case FK_TRAIN_DEV:
    case FK_TRAIN_COUNTER:
        return plain(TF_[fd.slot] + L.off, n);
    }
    (void)count;
    return SNN_UNKNOWN_FIELD;
}

int Engine::set_field(uint64_t id, const char *name, const void *data, uint64_t count, int dtype) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices, id: " + std::to_string(id));
    if (!name || (!data && L->n)) return fail(SNN_INVALID_ARGUMENT, "null argument");
    FieldDef fd;
    if (lookup_field(*L, name, &fd)) return fail(SNN_UNKNOWN_FIELD, std::string("unknown field: ") + name);
    if (dtype != fd.dtype) return fail(SNN_DTYPE_MISMATCH, std::string("dtype mismatch for field: ") + name);
    if (count != L->n * (uint64_t)fd.per) return fail(SNN_SIZE_MISMATCH, std::string("size mismatch for field: ") + name);
    return field_io(*L, fd, const_cast<void *>(data), count, true);
}

int Engine::get_field(uint64_t id, const char *name, void *out, uint64_t count, int dtype) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices, id: " + std::to_string(id));
    if (!name || (!out && L->n)) return fail(SNN_INVALID_ARGUMENT, "null argument");
    FieldDef fd;
    if (lookup_field(*L, name, &fd)) return fail(SNN_UNKNOWN_FIELD, std::string("unknown field: ") + name);
    if (dtype != fd.dtype) return fail(SNN_DTYPE_MISMATCH, std::string("dtype mismatch for field: ") + name);
    if (count != L->n * (uint64_t)fd.per) return fail(SNN_SIZE_MISMATCH, std::string("size mismatch for field: ") + name);
    return field_io(*L, fd, out, count, false);
}

int Engine::fill_field(uint64_t id, const char *name, uint32_t bits, int dtype) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices, id: " + std::to_string(id));
    FieldDef fd;
    if (!name || lookup_field(*L, name, &fd)) return fail(SNN_UNKNOWN_FIELD, std::string("unknown field: ") + (name ? name : "(null)"));
    if (dtype != fd.dtype) return fail(SNN_DTYPE_MISMATCH, std::string("dtype mismatch for field: ") + name);
    if (L->n == 0) return SNN_OK;
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    // fast device-side fills for the plain arrays; everything else goes through the host path
    uint32_t *dev = nullptr;
    switch (fd.kind) {
    case FK_NEURON_DEV: dev = (uint32_t *)(F_[fd.slot] + L->off); break;
    case FK_V: dev = (uint32_t *)(V_[cur_] + node_off(*L)); break;
    case FK_TRAIN_DEV: case FK_TRAIN_COUNTER: dev = (uint32_t *)(TF_[fd.slot] + L->off); break;
    case FK_RC: { int r = ensure_chem(); if (r) return r; dev = (uint32_t *)(RC_[fd.slot] + (size_t)fd.aux * neuron_cap_ + L->off); break; }
    default: break;
    }
    if (dev) {
        CK(launch_fill_u32(dev, bits, L->n, stream_), SNN_GPU_QUEUE_FAILURE);
        CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
        return SNN_OK;
    }
    std::vector<uint32_t> tmp(L->n * fd.per, bits);
    return field_io(*L, fd, tmp.data(), tmp.size(), true);
}

int Engine::set_preset_firing_times(uint64_t id, const uint64_t *offsets, const float *times, uint64_t nt, uint64_t n_times) {
    Lat *L = find(id);
    if (!L || !L->is_train) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in spike train lattices");
    if (train_kind != SNN_TRAIN_PRESET) return fail(SNN_UNSUPPORTED, "spike train type has no firing_times");
    if (nt != L->n || !offsets || offsets[nt] != n_times) return fail(SNN_SIZE_MISMATCH, "firing_times size mismatch");
    L->ft_off.assign(offsets, offsets + nt + 1);
    L->ft.assign(times, times + n_times);
    // rebuild the concatenated CSR over all train lattices
    std::vector<uint64_t> off(n_trains + 1, 0);
    std::vector<float> all;
    for (auto &X : lats_) {
        if (!X.is_train) continue;
        for (uint64_t i = 0; i < X.n; ++i) {
            off[X.off + i] = all.size();
            if (!X.ft_off.empty()) all.insert(all.end(), X.ft.begin() + X.ft_off[i], X.ft.begin() + X.ft_off[i + 1]);
        }
    }
    off[n_trains] = all.size();
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    if (ft_) cudaFree(ft_);
    CK(dev_alloc(&ft_, all.size()), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(h2d_sync(ft_off_, off.data(), off.size() * 8, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    if (!all.empty()) CK(h2d_sync(ft_, all.data(), all.size() * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// general-graph partition (second half): contiguous node ranges per rank, in-edges from any node of any rank
// ------------------------------------------------------------------------------------------------
// Node order on a rank: [ghosts below my range, ascending global index | my neurons | ghosts above], so local order == global
// order and the canonical (ascending presynaptic index) summation order of every row is the one of the unpartitioned lattice.
int Engine::relayout_keep_fields(const std::function<void()> &change) {
    // snapshot every field through the public get path, change the layout parameters, rebuild, restore
    std::vector<LatSnap> snaps;
    const bool had_chem = chem_alloc_;
    for (auto &L : lats_) {
        LatSnap sn; sn.id = L.id;
        uint32_t cnt = 0; field_count(L.id, &cnt);
        for (uint32_t i = 0; i < cnt; ++i) {
            const char *name; int32_t dt; uint32_t per;
            field_info(L.id, i, &name, &dt, &per);
            FieldDef fd;
            if (lookup_field(L, name, &fd)) continue;
            const bool is_chem = fd.kind == FK_NT_FLAGS || fd.kind == FK_NT_T || fd.kind == FK_NT || fd.kind == FK_RC_FLAGS || fd.kind == FK_RC;
            if (is_chem && !had_chem) continue;
            FieldSnap f; f.name = name; f.dtype = dt; f.count = L.n * per; f.data.resize(std::max<uint64_t>(f.count, 1) * 4);
            int r = get_field(L.id, name, f.data.data(), f.count, dt);
            if (r) return r;
            sn.fields.push_back(std::move(f));
        }
        snaps.push_back(std::move(sn));
    }
    const uint64_t saved_clock = internal_clock;
    free_device();
    change();
    compute_layout();
    int r = alloc_device();
    if (r) return r;
    if (had_chem) { r = ensure_chem(); if (r) return r; }
    for (auto &sn : snaps)
        for (auto &f : sn.fields) { r = set_field(sn.id, f.name.c_str(), f.data.data(), f.count, f.dtype); if (r) return r; }
    internal_clock = saved_clock;
    graph_dirty_ = true;
    return SNN_OK;
}

int Engine::gpart_enable(const std::vector<uint32_t> &remote) {
    if (layout_frozen_) return fail(SNN_UNSUPPORTED, "the graph of a partitioned handle cannot change its reach after the slab was exported");
    const uint64_t cols = lats_[0].cols, g0 = (uint64_t)row0_global * cols, g1 = g0 + n_neurons;
    std::vector<uint32_t> lo, hi;
    for (uint32_t g : remote) (g < g0 ? lo : hi).push_back(g);
    (void)g1;
    if (gpart_ && lo == g_lo_ && hi == g_hi_) return SNN_OK;
    for (int d = 0; d < 2; ++d) if (halo_dir_[d].active) return fail(SNN_UNSUPPORTED, "handle already attached as a row strip");
    int r = relayout_keep_fields([&]() { gpart_ = true; g_lo_ = lo; g_hi_ = hi; });
    if (r) return r;
    gpeers_.clear();
    gpart_dirty_ = true;
    return SNN_OK;
}

int64_t Engine::global_to_node(uint64_t g) const {
    const uint64_t cols = lats_.empty() ? 0 : lats_[0].cols, g0 = (uint64_t)row0_global * cols;
    if (g >= g0 && g < g0 + n_neurons) return (int64_t)own0_ + (int64_t)(g - g0);
    if (!gpart_) {
        const int64_t node = (int64_t)g + (int64_t)own0_ - (int64_t)g0;
        return (node < 0 || node >= (int64_t)n_nodes_) ? -1 : node;
    }
    const std::vector<uint32_t> &v = g < g0 ? g_lo_ : g_hi_;
    auto it = std::lower_bound(v.begin(), v.end(), (uint32_t)g);
    if (it == v.end() || *it != g) return -1;
    const uint64_t k = (uint64_t)(it - v.begin());
    return g < g0 ? (int64_t)own0_ - (int64_t)g_lo_.size() + (int64_t)k : (int64_t)ghost_hi0_ + (int64_t)k;
}

uint64_t Engine::node_to_global(uint32_t node) const {
    const uint64_t cols = lats_.empty() ? 0 : lats_[0].cols, g0 = (uint64_t)row0_global * cols;
    if (node >= own0_ && node < own0_ + n_neurons) return g0 + (node - own0_);
    if (!gpart_) return (uint64_t)((int64_t)node + (int64_t)g0 - (int64_t)own0_);
    if (node < own0_) return g_lo_[node - (own0_ - (uint32_t)g_lo_.size())];
    return g_hi_[node - ghost_hi0_];
}

Engine::GPeerHost *Engine::gpeer(int rank, bool create) {
    for (auto &g : gpeers_) if (g.rank == rank) return &g;
    if (!create) return nullptr;
    GPeerHost g; g.rank = rank;
    auto it = gpeers_.begin();
    while (it != gpeers_.end() && it->rank < rank) ++it;
    return &*gpeers_.insert(it, std::move(g));
}

int Engine::gpart_wants(int peer, uint32_t *global_idx, uint64_t capacity, uint64_t *n, uint32_t *first_slot) {
    if (part_world <= 1 || peer < 0 || peer >= part_world || peer == part_rank) return fail(SNN_INVALID_ARGUMENT, "bad peer rank");
    if (n) *n = 0;
    if (first_slot) *first_slot = 0;
    if (!gpart_) return SNN_OK;   // a row strip has no gather lists
    const uint64_t cols = lats_[0].cols;
    const uint64_t q0 = (uint64_t)snn_partition_begin(rows_global, part_world, peer) * cols, q1 = (uint64_t)snn_partition_begin(rows_global, part_world, peer + 1) * cols;
    const std::vector<uint32_t> &v = peer < part_rank ? g_lo_ : g_hi_;
    auto a = std::lower_bound(v.begin(), v.end(), (uint32_t)q0), b = std::lower_bound(v.begin(), v.end(), (uint32_t)std::min<uint64_t>(q1, 0xFFFFFFFFull));
    const uint64_t cnt = (uint64_t)(b - a);
    if (n) *n = cnt;
    const uint32_t base = peer < part_rank ? own0_ - (uint32_t)g_lo_.size() : ghost_hi0_;
    if (first_slot) *first_slot = base + (uint32_t)(a - v.begin());
    if (global_idx) {
        if (capacity < cnt) return fail(SNN_SIZE_MISMATCH, "index buffer too small");
        std::copy(a, b, global_idx);
    }
    return SNN_OK;
}

int Engine::gpart_set_exports(int peer, const uint32_t *global_idx, uint64_t n, uint32_t first_slot_at_peer) {
    if (part_world <= 1 || peer < 0 || peer >= part_world || peer == part_rank) return fail(SNN_INVALID_ARGUMENT, "bad peer rank");
    if (n && !global_idx) return fail(SNN_INVALID_ARGUMENT, "null argument");
    if (!gpart_) return fail(SNN_INVALID_ARGUMENT, "not a general-graph partition (set SNN_OPT_GENERAL_PARTITION before the graph on every rank)");
    if (n == 0 && !gpeer(peer, false)) return SNN_OK;
    const uint64_t cols = lats_[0].cols, g0 = (uint64_t)row0_global * cols;
    GPeerHost *G = gpeer(peer, true);
    G->exp_local.clear();
    for (uint64_t k = 0; k < n; ++k) {
        if (global_idx[k] < g0 || global_idx[k] >= g0 + n_neurons) return fail(SNN_GRAPH_POSITION_NOT_FOUND, "Position not found, position: " + std::to_string(global_idx[k]));
        if (k && global_idx[k] <= global_idx[k - 1]) return fail(SNN_INVALID_ARGUMENT, "export list must be ascending");
        G->exp_local.push_back((uint32_t)(global_idx[k] - g0));
    }
    G->peer_slot0 = first_slot_at_peer;
    gpart_dirty_ = true;
    return SNN_OK;
}

int Engine::gpart_attach(int peer, const IpcBlob *blob_in, Engine *local_peer) {
    if (part_world <= 1 || peer < 0 || peer >= part_world || peer == part_rank) return fail(SNN_INVALID_ARGUMENT, "bad peer rank");
    if (!gpart_) return fail(SNN_INVALID_ARGUMENT, "not a general-graph partition (set SNN_OPT_GENERAL_PARTITION before the graph on every rank)");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    GPeerHost *G = gpeer(peer, true);
    if (G->attached) return fail(SNN_INVALID_ARGUMENT, "peer already attached");
    IpcBlob blob;
    if (local_peer) {
        if (local_peer == this) return fail(SNN_INVALID_ARGUMENT, "bad peer handle");
        if (local_peer->device != device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, device, local_peer->device), SNN_GPU_GET_DEVICE_FAILURE);
            if (!can) return fail(SNN_UNSUPPORTED, "no peer access between the two devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(local_peer->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, SNN_GPU_GET_DEVICE_FAILURE, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        int r = local_peer->ipc_export_layout(&blob);
        if (r) return fail(r, local_peer->last_error);
        G->slab = local_peer->slab_; G->flags = local_peer->flags_;
        local_peer->layout_frozen_ = true;
        local_peers_.push_back(local_peer);
    } else {
        blob = *blob_in;
        if (blob.magic != 0x534E4E42u || blob.version != SNN_B200_ABI_VERSION) return fail(SNN_INVALID_ARGUMENT, "bad ipc blob");
        CK(cudaIpcOpenMemHandle(&G->slab, blob.slab, cudaIpcMemLazyEnablePeerAccess), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(cudaIpcOpenMemHandle(&G->flags, blob.flags, cudaIpcMemLazyEnablePeerAccess), SNN_GPU_BUFFER_CREATE_ERROR);
        G->ipc = true;
    }
    if (blob.rank != peer || blob.world != part_world) return fail(SNN_INVALID_ARGUMENT, "ipc blob is not from that rank");
    if ((blob.chem != 0) != chem_alloc_) return fail(SNN_INVALID_ARGUMENT, "the ranks disagree on chemistry");
    G->view = blob;
    G->attached = true;
    layout_frozen_ = true;
    // neurotransmitter / receptor type flags of the ghosts that this peer owns (baked into the col words of edges from ghosts)
    uint64_t cnt = 0; uint32_t slot0 = 0;
    int r = gpart_wants(peer, nullptr, 0, &cnt, &slot0);
    if (r) return r;
    if (cnt) {
        std::vector<uint8_t> pf(blob.n_neurons);
        CK(cudaMemcpy(pf.data(), (const uint8_t *)G->slab + blob.off_flags + blob.own0, blob.n_neurons, cudaMemcpyDefault), SNN_GPU_BUFFER_READ_ERROR);
        const uint64_t cols = lats_[0].cols, q0 = (uint64_t)snn_partition_begin(rows_global, part_world, peer) * cols;
        bool changed = false;
        for (uint64_t k = 0; k < cnt; ++k) {
            const uint32_t node = slot0 + (uint32_t)k;
            const uint8_t f = pf[node_to_global(node) - q0];
            changed |= h_node_flags_[node] != f;
            h_node_flags_[node] = f;
        }
        if (changed) {
            CK(h2d_sync(node_flags_ + slot0, h_node_flags_.data() + slot0, cnt, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
            flags_cache_valid_ = false;
            if (dev_weights_newer_) { r = sync_weights_to_host(); if (r) return r; }
            graph_dirty_ = true;
        }
    }
    gpart_dirty_ = true;
    return SNN_OK;
}

int Engine::gpart_build_device() {
    if (!gpart_dirty_) return SNN_OK;
    auto fr = [](auto *&p) { if (p) cudaFree(p); p = nullptr; };
    fr(d_gpeers_); fr(d_gexp_off_); fr(d_gexp_ent_); fr(d_gslice_);
    if (gpeers_.size() > 15) return fail(SNN_UNSUPPORTED, "too many peers");
    std::vector<GPeer> hp(std::max<size_t>(gpeers_.size(), 1));
    std::vector<std::vector<uint32_t>> per(n_neurons);
    for (size_t q = 0; q < gpeers_.size(); ++q) {
        GPeerHost &G = gpeers_[q];
        if (!G.attached) return fail(SNN_INVALID_ARGUMENT, "general-graph partition: rank " + std::to_string(G.rank) + " is not attached");
        GPeer &D = hp[q];
        for (int k = 0; k < 2; ++k) {
            D.v[k] = (float *)((char *)G.slab + G.view.off_v[k]);
            D.lft[k] = (int *)((char *)G.slab + G.view.off_lft[k]);
            D.t[k] = (float *)((char *)G.slab + G.view.off_t[k]);
        }
        D.t_stride = G.view.t_stride;
        D.peer_flag = (unsigned long long *)G.flags + 8 + part_rank;
        D.my_flag = flags_ + 8 + G.rank;
        for (size_t k = 0; k < G.exp_local.size(); ++k) per[G.exp_local[k]].push_back(((uint32_t)q << 28) | (G.peer_slot0 + (uint32_t)k));
    }
    std::vector<uint32_t> off(n_neurons + 1, 0), ent;
    std::vector<uint8_t> gs(std::max<size_t>(n_slices_, 1), 0);
    for (uint64_t i = 0; i < n_neurons; ++i) {
        off[i] = (uint32_t)ent.size();
        ent.insert(ent.end(), per[i].begin(), per[i].end());
        if (!per[i].empty()) gs[i / 32] |= 2u;
    }
    off[n_neurons] = (uint32_t)ent.size();
    for (size_t sl = 0; sl < h_gslice_.size() && sl < gs.size(); ++sl) gs[sl] |= h_gslice_[sl] & 1u;
    CK(dev_alloc(&d_gpeers_, hp.size()), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&d_gexp_off_, off.size()), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&d_gexp_ent_, ent.size()), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&d_gslice_, gs.size()), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(h2d_sync(d_gpeers_, hp.data(), hp.size() * sizeof(GPeer), stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(h2d_sync(d_gexp_off_, off.data(), off.size() * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(h2d_sync(d_gexp_ent_, ent.data(), ent.size() * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(h2d_sync(d_gslice_, gs.data(), gs.size(), stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    gpart_dirty_ = false;
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// graph ingestion
// ------------------------------------------------------------------------------------------------
static int check_connect(Engine &E, uint64_t pre_id, uint64_t post_id, Lat **A, Lat **B) {
    // LatticeNetwork::connect preconditions, neuron/mod.rs:1852-1862
    Lat *pre = E.find(pre_id), *post = E.find(post_id);
    if (post && post->is_train)
        return E.fail(SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN,
                      "Postsynaptic lattice cannot be a spike train lattice because spike trains cannot take inputs");
    if (!pre) return E.fail(SNN_NET_PRESYNAPTIC_ID_NOT_FOUND, "Presynaptic id not present in network, id: " + std::to_string(pre_id));
    if (!post) return E.fail(SNN_NET_POSTSYNAPTIC_ID_NOT_FOUND, "Postsynaptic id not present in network, id: " + std::to_string(post_id));
    *A = pre; *B = post;
    return SNN_OK;
}

static void sort_rows(Block &b) {
    const uint64_t n = b.row_ptr.size() - 1;
    std::vector<std::pair<uint32_t, float>> tmp;
    for (uint64_t q = 0; q < n; ++q) {
        const uint64_t s = b.row_ptr[q], e = b.row_ptr[q + 1];
        bool sorted = true;
        for (uint64_t k = s + 1; k < e; ++k) if (b.pre[k - 1] > b.pre[k]) { sorted = false; break; }
        if (sorted) continue;
        tmp.clear();
        for (uint64_t k = s; k < e; ++k) tmp.emplace_back(b.pre[k], b.w[k]);
        std::stable_sort(tmp.begin(), tmp.end(), [](auto &x, auto &y) { return x.first < y.first; });
        for (uint64_t k = s; k < e; ++k) { b.pre[k] = tmp[k - s].first; b.w[k] = tmp[k - s].second; }
    }
}

int Engine::connect_dense(uint64_t pre_id, uint64_t post_id, const uint32_t *connections, const float *weights,
                          const uint32_t *itp, uint64_t n_pre, uint64_t n_post) {
    Lat *A, *B;
    int r = check_connect(*this, pre_id, post_id, &A, &B);
    if (r) return r;
    if (n_pre != A->n || n_post != B->n) return fail(SNN_GRAPH_DIMENSIONS_DO_NOT_MATCH, "Dimensions do not match");
    if (n_pre != 0 && n_post != 0 && (!connections || !weights)) return fail(SNN_INVALID_ARGUMENT, "null graph pointers");
    if (part_world > 1) return fail(SNN_UNSUPPORTED, "dense graphs are not supported on partitioned handles");
    if (itp) {
        if (pre_id != post_id) return fail(SNN_INVALID_ARGUMENT, "index_to_position only applies to an internal graph");
        for (uint64_t i = 0; i < n_pre; ++i)
            if (itp[i] >= n_pre) return fail(SNN_GRAPH_POSITION_NOT_FOUND, "Position not found, position: " + std::to_string(itp[i]));
    }
    if (dev_weights_newer_) { r = sync_weights_to_host(); if (r) return r; }
    Block b;
    b.kind = Block::CSR;
    b.row_ptr.assign(n_post + 1, 0);
    // rows are indexed by cell position; graph index g maps to position itp[g]
    std::vector<uint64_t> cnt(n_post + 1, 0);
    for (uint64_t q = 0; q < n_post; ++q) {
        const uint64_t qp = itp ? itp[q] : q;
        uint64_t c = 0;
        for (uint64_t p = 0; p < n_pre; ++p) if (connections[p * n_post + q]) c++;
        cnt[qp] = c;
    }
    for (uint64_t q = 0; q < n_post; ++q) b.row_ptr[q + 1] = b.row_ptr[q] + cnt[q];
    b.pre.resize(b.row_ptr[n_post]); b.w.resize(b.row_ptr[n_post]);
    for (uint64_t q = 0; q < n_post; ++q) {
        const uint64_t qp = itp ? itp[q] : q;
        uint64_t o = b.row_ptr[qp];
        for (uint64_t p = 0; p < n_pre; ++p)
            if (connections[p * n_post + q]) { b.pre[o] = (uint32_t)(itp ? itp[p] : p); b.w[o] = weights[p * n_post + q]; ++o; }
    }
    if (itp) sort_rows(b);
    blocks_[{pre_id, post_id}] = std::move(b);
    graph_dirty_ = true;
    return SNN_OK;
}

int Engine::connect_csr(uint64_t pre_id, uint64_t post_id, const uint64_t *row_ptr, const uint32_t *pre, const float *weights,
                        uint64_t n_post, uint64_t nnz) {
    Lat *A, *B;
    int r = check_connect(*this, pre_id, post_id, &A, &B);
    if (r) return r;
    if (n_post != B->n) return fail(SNN_GRAPH_DIMENSIONS_DO_NOT_MATCH, "Dimensions do not match");
    if (!row_ptr || (nnz && (!pre || !weights))) return fail(SNN_INVALID_ARGUMENT, "null graph pointers");
    if (row_ptr[0] != 0 || row_ptr[n_post] != nnz) return fail(SNN_SIZE_MISMATCH, "row_ptr does not match nnz");
    const uint64_t pre_limit = part_world > 1 ? (uint64_t)rows_global * A->cols : A->n;
    for (uint64_t k = 0; k < nnz; ++k)
        if (pre[k] >= pre_limit) return fail(SNN_GRAPH_PRESYNAPTIC_NOT_FOUND, "Presynaptic position not found, position: " + std::to_string(pre[k]));
    for (uint64_t q = 0; q < n_post; ++q)
        if (row_ptr[q + 1] < row_ptr[q]) return fail(SNN_INVALID_ARGUMENT, "row_ptr must be non-decreasing");
    if (dev_weights_newer_) { r = sync_weights_to_host(); if (r) return r; }
    Block b;
    b.kind = Block::CSR;
    b.row_ptr.assign(row_ptr, row_ptr + n_post + 1);
    b.pre.assign(pre, pre + nnz);
    b.w.assign(weights, weights + nnz);
    sort_rows(b);
    if (part_world > 1) {
        // edges that reach beyond the strip's halo rows (or any edge, once the handle is a general-graph partition): the ghosts
        // become gather lists of exactly the remote nodes this rank's rows read
        const int64_t g0 = (int64_t)row0_global * A->cols, g1 = g0 + (int64_t)A->n;
        bool beyond = gpart_ || force_gpart;
        std::vector<uint32_t> remote;
        for (uint64_t k = 0; k < nnz; ++k) {
            const int64_t g = b.pre[k];
            if (g >= g0 && g < g1) continue;
            remote.push_back((uint32_t)g);
            if (g < g0 - (int64_t)halo_ || g >= g1 + (int64_t)halo_) beyond = true;
        }
        if (beyond) {
            std::sort(remote.begin(), remote.end());
            remote.erase(std::unique(remote.begin(), remote.end()), remote.end());
            blocks_.clear();
            r = gpart_enable(remote);
            if (r) return r;
            A = B = find(pre_id);
        }
    }
    blocks_[{pre_id, post_id}] = std::move(b);
    graph_dirty_ = true;
    return SNN_OK;
}

int Engine::connect_grid(uint64_t id, uint32_t radius, float weight) {
    Lat *A, *B;
    int r = check_connect(*this, id, id, &A, &B);
    if (r) return r;
    if (radius > 7) return fail(SNN_UNSUPPORTED, "stencil radius too large");
    if (part_world > 1) {
        const uint32_t need = radius * A->cols;
        if (need > A->n && A->n) return fail(SNN_UNSUPPORTED, "strip is thinner than the stencil radius");
        if (need != halo_) {
            // re-layout with a halo of `radius` rows (snapshot/restore through the field path)
            if (peer_slab_[0] || peer_slab_[1] || layout_frozen_) return fail(SNN_UNSUPPORTED, "cannot change the stencil after the halo slab was exported");
            Lat copy = *A;
            std::vector<LatSnap> unused;
            // temporarily remove the lattice and add it back with the new halo
            std::vector<FieldSnap> fields;
            uint32_t cnt = 0; field_count(id, &cnt);
            for (uint32_t i = 0; i < cnt; ++i) {
                const char *name; int32_t dt; uint32_t per;
                field_info(id, i, &name, &dt, &per);
                FieldDef fd;
                if (lookup_field(*A, name, &fd)) continue;
                const bool is_chem = fd.kind == FK_NT_FLAGS || fd.kind == FK_NT_T || fd.kind == FK_NT || fd.kind == FK_RC_FLAGS || fd.kind == FK_RC;
                if (is_chem && !chem_alloc_) continue;
                FieldSnap f; f.name = name; f.dtype = dt; f.count = A->n * per; f.data.resize(std::max<uint64_t>(f.count, 1) * 4);
                r = get_field(id, name, f.data.data(), f.count, dt);
                if (r) return r;
                fields.push_back(std::move(f));
            }
            const bool had_chem = chem_alloc_;
            free_device();
            halo_ = need;
            compute_layout();
            r = alloc_device();
            if (r) return r;
            if (had_chem) { r = ensure_chem(); if (r) return r; }
            for (auto &f : fields) { r = set_field(id, f.name.c_str(), f.data.data(), f.count, f.dtype); if (r) return r; }
            A = B = find(id);
        }
    }
    // the learned weights of every OTHER block live on the device until synced: replacing this block must not discard them
    if (dev_weights_newer_) {
        bool others = false;
        for (auto &kv : blocks_) others |= kv.first != std::make_pair(id, id);
        if (others) { r = sync_weights_to_host(); if (r) return r; }
    }
    Block b;
    b.kind = Block::GRID;
    b.radius = radius; b.weight = weight;
    blocks_[{id, id}] = std::move(b);
    dev_weights_newer_ = false;
    graph_dirty_ = true;
    return SNN_OK;
}

int Engine::materialize_grid(Block &b, const Lat &L) {
    // host CSR of the Moore stencil (same enumeration order as sell_grid_kernel: ascending flat index)
    const int64_t R = b.radius, rows_g = part_world > 1 ? rows_global : L.rows, cols = L.cols;
    const int64_t r0 = part_world > 1 ? row0_global : 0;
    b.row_ptr.assign(L.n + 1, 0);
    b.pre.clear(); b.w.clear();
    for (int64_t i = 0; i < (int64_t)L.rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            for (int64_t di = -R; di <= R; ++di)
                for (int64_t dj = -R; dj <= R; ++dj) {
                    const int64_t a = r0 + i + di, c = j + dj;
                    if ((di == 0 && dj == 0) || a < 0 || c < 0 || a >= rows_g || c >= cols) continue;
                    // pre index: global flat index on partitioned handles, local otherwise
                    b.pre.push_back((uint32_t)(a * cols + c));
                    b.w.push_back(b.weight);
                }
            b.row_ptr[i * cols + j + 1] = b.pre.size();
        }
    b.kind = Block::CSR;
    b.from_grid_radius = b.radius;
    return SNN_OK;
}

void Engine::refresh_flag_cache() const {
    if (flags_cache_valid_) return;
    uint32_t a = 0;
    for (uint32_t i = 0; i < n_nodes_; ++i) a |= h_node_flags_[i];
    nt_used_ = a & 0x7u; rc_used_ = (a >> 4) & 0x7u;
    flags_cache_valid_ = true;
}
uint32_t Engine::nt_used() const { refresh_flag_cache(); return nt_used_; }
uint32_t Engine::rc_used() const { refresh_flag_cache(); return rc_used_; }

int Engine::finalize_graph() {
    if (!graph_dirty_) return SNN_OK;
    free_reward_arrays();   // a rebuilt graph starts from TraceRSTDP::default traces
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    // the stencil fast path re-generates the table in place when its size is unchanged (set_graph_grid called again,
    // e.g. once per run by a caller that mirrors LatticeGPU's upload-everything-per-run behaviour)
    const uint32_t new_slices = (uint32_t)((n_neurons + 31) / 32);
    bool reuse = false;
    if (blocks_.size() == 1 && blocks_.begin()->second.kind == Block::GRID && col_ && wgt_ && slice_off_ && n_slices_ == new_slices) {
        const Block &b0 = blocks_.begin()->second;
        const uint32_t width0 = (2 * b0.radius + 1) * (2 * b0.radius + 1) - 1;
        reuse = sell_alloc_krows_ == round_up(new_slices, kTmaConsumerWarps) * width0;
    }
    if (!reuse) {
        if (slice_off_) cudaFree(slice_off_);
        if (col_) cudaFree(col_);
        if (wgt_) cudaFree(wgt_);
        slice_off_ = col_ = nullptr; wgt_ = nullptr;
        sell_alloc_krows_ = 0;
    }
    grid_fast_ = false;
    uniform_width_ = 0;
    n_slices_ = new_slices;
    if (!reuse) CK(dev_alloc(&slice_off_, (size_t)n_slices_ + 1), SNN_GPU_BUFFER_CREATE_ERROR);
    // partitioned handles: ghost rows inherit the neurotransmitter type sets of the adjacent owned rows
    if (part_world > 1 && n_neurons && !gpart_) {
        // (until the neighbour is attached: ipc_attach replaces them with the neighbour's real boundary flags)
        for (uint32_t g = 0; g < halo_; ++g) {
            if (part_rank > 0 && !ghost_flags_from_peer_[0]) h_node_flags_[own0_ - halo_ + g] = h_node_flags_[own0_ + (g % std::max<uint64_t>(n_neurons, 1))];
            if (part_rank < part_world - 1 && !ghost_flags_from_peer_[1]) h_node_flags_[ghost_hi0_ + g] = h_node_flags_[own0_ + n_neurons - halo_ + g];
        }
        CK(h2d_sync(node_flags_, h_node_flags_.data(), node_cap_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        flags_cache_valid_ = false;
    }
    int n_neuron_lat = 0;
    const Lat *only = nullptr;
    for (auto &L : lats_) if (!L.is_train) { n_neuron_lat++; only = &L; }
    // fast path: one lattice whose only block is a grid stencil -> built directly in sliced-ELL form on the device
    if (n_neuron_lat == 1 && blocks_.size() == 1 && blocks_.begin()->second.kind == Block::GRID &&
        blocks_.begin()->first == std::make_pair(only->id, only->id)) {
        const Block &b = blocks_.begin()->second;
        const uint32_t width = (2 * b.radius + 1) * (2 * b.radius + 1) - 1;
        sell_krows_ = (uint64_t)n_slices_ * width;
        const uint64_t alloc_krows = round_up(n_slices_, kTmaConsumerWarps) * width;  // whole tiles for the TMA kernel
        if (!reuse) {
            CK(dev_alloc(&col_, alloc_krows * 32), SNN_GPU_BUFFER_CREATE_ERROR);
            CK(dev_alloc(&wgt_, alloc_krows * 32), SNN_GPU_BUFFER_CREATE_ERROR);
            sell_alloc_krows_ = alloc_krows;
            // only the slices beyond the last neuron are not written by the generator
            const uint64_t tail0 = sell_krows_ * 32, tail_n = (alloc_krows - sell_krows_) * 32;
            if (tail_n) {
                CK(cudaMemsetAsync(col_ + tail0, 0xFF, tail_n * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
                CK(cudaMemsetAsync(wgt_ + tail0, 0, tail_n * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
            }
        }
        if (n_neurons == 0) CK(cudaMemsetAsync(slice_off_, 0, 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        CK(launch_sell_grid(only->rows, only->cols, part_world > 1 ? row0_global : 0, part_world > 1 ? rows_global : only->rows,
                            b.radius, b.weight, own0_, node_flags_, width, slice_off_, col_, wgt_, stream_), SNN_GPU_QUEUE_FAILURE);
        CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
        grid_fast_ = true;
        uniform_width_ = width;
        graph_dirty_ = false;
        dev_weights_newer_ = false;
        return SNN_OK;
    }
    // general path: merge every block into one CSR over canonical node indices
    for (auto &kv : blocks_)
        if (kv.second.kind == Block::GRID) materialize_grid(kv.second, *find(kv.first.first));
    std::vector<uint64_t> row_ptr(n_neurons + 1, 0);
    // blocks into each post lattice ordered by the pre lattice's node offset => rows stay sorted by node index
    std::vector<std::vector<std::pair<const Lat *, const Block *>>> into(lats_.size());
    for (size_t li = 0; li < lats_.size(); ++li) {
        if (lats_[li].is_train) continue;
        for (auto &kv : blocks_)
            if (kv.first.second == lats_[li].id) {
                const Lat *A = find(kv.first.first);
                if (A) into[li].emplace_back(A, &kv.second);
            }
        std::sort(into[li].begin(), into[li].end(), [&](auto &x, auto &y) { return node_off(*x.first) < node_off(*y.first); });
        for (uint64_t q = 0; q < lats_[li].n; ++q) {
            uint64_t c = 0;
            for (auto &ab : into[li]) c += ab.second->row_ptr[q + 1] - ab.second->row_ptr[q];
            row_ptr[lats_[li].off + q + 1] = c;
        }
    }
    for (uint64_t i = 0; i < n_neurons; ++i) row_ptr[i + 1] += row_ptr[i];
    const uint64_t nnz = row_ptr[n_neurons];
    std::vector<uint32_t> pre(std::max<uint64_t>(nnz, 1));
    std::vector<float> w(std::max<uint64_t>(nnz, 1));
    const int64_t node_shift = part_world > 1 ? (int64_t)own0_ - (int64_t)row0_global * (only ? only->cols : 0) : 0;
    std::vector<uint64_t> ghost_rows;   // general-graph partition: rows with an in-edge from a ghost
    for (size_t li = 0; li < lats_.size(); ++li) {
        if (lats_[li].is_train) continue;
        for (uint64_t q = 0; q < lats_[li].n; ++q) {
            uint64_t o = row_ptr[lats_[li].off + q];
            for (auto &ab : into[li]) {
                const Block &b = *ab.second;
                const uint32_t base = node_off(*ab.first);
                for (uint64_t k = b.row_ptr[q]; k < b.row_ptr[q + 1]; ++k) {
                    int64_t node = gpart_ ? global_to_node(b.pre[k]) : (part_world > 1 ? (int64_t)b.pre[k] + node_shift : (int64_t)base + b.pre[k]);
                    if (node < 0 || node >= (int64_t)n_nodes_)
                        return fail(SNN_UNSUPPORTED, "edge reaches beyond the halo of this strip");
                    if (gpart_ && (node < (int64_t)own0_ || node >= (int64_t)own0_ + (int64_t)n_neurons)) ghost_rows.push_back(lats_[li].off + q);
                    pre[o] = (uint32_t)node; w[o] = b.w[k]; ++o;
                }
            }
        }
    }
    std::vector<uint32_t> slice_off(n_slices_ + 1, 0);
    for (uint32_t s = 0; s < n_slices_; ++s) {
        uint64_t mx = 0;
        for (uint64_t r = (uint64_t)s * 32; r < std::min<uint64_t>((uint64_t)s * 32 + 32, n_neurons); ++r)
            mx = std::max(mx, row_ptr[r + 1] - row_ptr[r]);
        mx = round_up(mx, 4);  // the step kernel consumes edges in chunks of 4/8 without tail checks
        const uint64_t next = (uint64_t)slice_off[s] + mx;
        if (next > 0xFFFFFFF0ull / 32) return fail(SNN_UNSUPPORTED, "graph too large");
        slice_off[s + 1] = (uint32_t)next;
    }
    sell_krows_ = slice_off[n_slices_];
    // a materialised stencil (weights came back from the device: STDP, edit_weight) keeps the uniform-width layout of the
    // generator, so the window / TMA step kernels stay eligible after a rebuild
    uint32_t stencil_width = 0;
    if (n_neuron_lat == 1 && blocks_.size() == 1 && blocks_.begin()->second.from_grid_radius &&
        blocks_.begin()->first == std::make_pair(only->id, only->id)) {
        const uint32_t R = blocks_.begin()->second.from_grid_radius;
        stencil_width = (2 * R + 1) * (2 * R + 1) - 1;
        for (uint32_t s = 0; s < n_slices_; ++s) {
            if (slice_off[s + 1] - slice_off[s] > stencil_width) { stencil_width = 0; break; }
        }
        if (stencil_width) {
            for (uint32_t s = 0; s <= n_slices_; ++s) slice_off[s] = s * stencil_width;
            sell_krows_ = (uint64_t)n_slices_ * stencil_width;
        }
    }
    const uint64_t alloc_krows = stencil_width ? round_up(n_slices_, kTmaConsumerWarps) * stencil_width : sell_krows_;
    CK(dev_alloc(&col_, alloc_krows * 32), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&wgt_, alloc_krows * 32), SNN_GPU_BUFFER_CREATE_ERROR);
    if (stencil_width) {
        sell_alloc_krows_ = alloc_krows;
        if (alloc_krows > sell_krows_) {
            CK(cudaMemsetAsync(col_ + sell_krows_ * 32, 0xFF, (alloc_krows - sell_krows_) * 32 * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
            CK(cudaMemsetAsync(wgt_ + sell_krows_ * 32, 0, (alloc_krows - sell_krows_) * 32 * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        }
    }
    CK(h2d_sync(slice_off_, slice_off.data(), ((size_t)n_slices_ + 1) * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    uint64_t *d_rp = nullptr; uint32_t *d_pre = nullptr; float *d_w = nullptr;
    CK(dev_alloc(&d_rp, n_neurons + 1), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&d_pre, nnz), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&d_w, nnz), SNN_GPU_BUFFER_CREATE_ERROR);
    h2d_sync(d_rp, row_ptr.data(), (n_neurons + 1) * 8, stream_);
    if (nnz) { h2d_sync(d_pre, pre.data(), nnz * 4, stream_); h2d_sync(d_w, w.data(), nnz * 4, stream_); }
    cudaError_t e = launch_sell_from_csr(d_rp, d_pre, d_w, node_flags_, part_world > 1 ? 0xFFFFFFFFu : train0_, (uint32_t)n_neurons,
                                         slice_off_, col_, wgt_, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d_rp); cudaFree(d_pre); cudaFree(d_w);
    if (e != cudaSuccess) return cuda_fail(e, SNN_GPU_QUEUE_FAILURE, "sell_from_csr");
    if (stencil_width) { grid_fast_ = true; uniform_width_ = stencil_width; }
    if (gpart_) {
        h_gslice_.assign(n_slices_, 0);
        for (uint64_t row : ghost_rows) h_gslice_[row / 32] |= 1u;
        gpart_dirty_ = true;
    }
    graph_dirty_ = false;
    dev_weights_newer_ = false;
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// RewardModulatedLattice
// ------------------------------------------------------------------------------------------------
int Engine::set_bcm_plasticity(bool enable, const snn_bcm_t *b) {
    if (enable && model != SNN_MODEL_BCM_IZHIKEVICH) return fail(SNN_INVALID_ARGUMENT, "BCM plasticity needs neurons with BCMActivity (SNN_MODEL_BCM_IZHIKEVICH)");
    int n_neuron_lat = 0;
    for (auto &L : lats_) if (!L.is_train) n_neuron_lat++;
    if (enable && (part_world > 1 || n_neuron_lat > 1 || n_trains > 0)) return fail(SNN_UNSUPPORTED, "BCM plasticity: single-GPU Lattice handles only");
    bcm_mode = enable;
    if (b) bcm = *b;
    return SNN_OK;
}

int Engine::set_reward_modulator(bool enable, bool modulate, const snn_rstdp_t *m) {
    int n_neuron_lat = 0;
    for (auto &L : lats_) if (!L.is_train) n_neuron_lat++;
    if (enable && (n_neuron_lat > 1 || n_trains > 0)) return fail(SNN_UNSUPPORTED, "snn_lattice_set_reward_modulator is for single-lattice handles; networks add reward-modulated lattices with snn_network_add_reward_modulated_lattice");
    reward_mode = enable; do_modulation = modulate;
    if (m) rstdp = *m;
    if (!enable) free_reward_arrays();
    return SNN_OK;
}

// RewardModulatedLatticeNetwork: per-lattice modulators and the kind of the connecting blocks
int Engine::set_lattice_reward_modulator(uint64_t id, bool modulate, const snn_rstdp_t *m) {
    Lat *L = find(id);
    if (!L || !L->is_reward) return fail(SNN_INVALID_ARGUMENT, "not a reward-modulated lattice of this network, id: " + std::to_string(id));
    L->do_modulation = modulate;
    if (m) L->rstdp = *m;
    return SNN_OK;
}

int Engine::get_lattice_reward_modulator(uint64_t id, int32_t *modulate, snn_rstdp_t *m) {
    Lat *L = find(id);
    if (!L || !L->is_reward) return fail(SNN_INVALID_ARGUMENT, "not a reward-modulated lattice of this network, id: " + std::to_string(id));
    if (modulate) *modulate = L->do_modulation ? 1 : 0;
    if (m) *m = L->rstdp;
    return SNN_OK;
}

int Engine::set_connection_reward(uint64_t pre_id, uint64_t post_id, bool reward_modulated) {
    // connect_with_reward_modulation neuron/mod.rs:4076-4209: the block's values are RewardModulatedConnection::RewardModulatedWeight
    // (TraceRSTDP) instead of ::Weight(f32)
    const Lat *A = find(pre_id), *B = find(post_id);
    if (B && B->is_train)
        return fail(SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN, "Postsynaptic lattice cannot be a spike train lattice because spike trains cannot take inputs");
    if (!A) return fail(SNN_NET_PRESYNAPTIC_ID_NOT_FOUND, "Presynaptic id not present in network, id: " + std::to_string(pre_id));
    if (!(B && B->is_reward) && !A->is_reward)
        return fail(SNN_NET_CANNOT_CONNECT_WITH_REWARD_MODULATED_CONNECTION, "When connecting reward modulated network, at least one lattice has to be reward modulated");
    if (!B) return fail(SNN_NET_POSTSYNAPTIC_ID_NOT_FOUND, "Postsynaptic id not present in network, id: " + std::to_string(post_id));
    if (pre_id == post_id)
        return fail(SNN_NET_REWARD_MODULATED_CONNECTION_NOT_COMPATIBLE_INTERNALLY,
                    "When connecting reward modulated lattice, RewardModulatedConnection cannot be used to connect a reward modulated lattice internally");
    auto it = blocks_.find({pre_id, post_id});
    if (it == blocks_.end()) return fail(SNN_INVALID_ARGUMENT, "no such connecting block: connect it first");
    it->second.reward_conn = reward_modulated;
    return SNN_OK;
}

void Engine::free_reward_arrays() {
    if (rs_counter_) cudaFree(rs_counter_);
    if (rs_dw_) cudaFree(rs_dw_);
    if (rs_c_) cudaFree(rs_c_);
    rs_counter_ = nullptr; rs_dw_ = rs_c_ = nullptr; rs_elems_ = 0;
}

int Engine::ensure_reward_arrays() {
    const uint64_t elems = std::max<uint64_t>(std::max(sell_alloc_krows_, sell_krows_) * 32, 1);
    if (rs_counter_ && rs_elems_ == elems) return SNN_OK;
    free_reward_arrays();
    CK(dev_alloc(&rs_counter_, elems), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&rs_dw_, elems), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(dev_alloc(&rs_c_, elems), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(cudaMemsetAsync(rs_counter_, 0, elems, stream_), SNN_GPU_BUFFER_WRITE_ERROR);   // TraceRSTDP::default: counter 0, dw 0, c 0
    CK(cudaMemsetAsync(rs_dw_, 0, elems * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(cudaMemsetAsync(rs_c_, 0, elems * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    rs_elems_ = elems;
    rs_canonical_ = true;
    return SNN_OK;
}

bool Engine::has_reward_lattices() const {
    for (auto &L : lats_) if (L.is_reward) return true;
    return false;
}

// element index (k * 32 + lane) of every edge of block pre_id -> post_id, in the block's CSR order; rows are laid out as
// finalize_graph wrote them (blocks into a lattice ordered by the presynaptic lattice's node offset)
int Engine::block_elements(uint64_t pre_id, uint64_t post_id, std::vector<size_t> *elems, Block **blk) {
    Lat *B = find(post_id);
    if (!find(pre_id) || !B || B->is_train) return fail(SNN_INVALID_ARGUMENT, "no such connection block");
    *blk = nullptr;
    elems->clear();
    auto it = blocks_.find({pre_id, post_id});
    if (it == blocks_.end()) return SNN_OK;
    *blk = &it->second;
    std::vector<uint32_t> so((size_t)n_slices_ + 1);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    CK(cudaMemcpy(so.data(), slice_off_, so.size() * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    Block tmp;
    const Block *target = &it->second;
    if (target->kind == Block::GRID) { tmp = *target; materialize_grid(tmp, *B); target = &tmp; }   // single stencil lattice: keep the device fast path
    std::vector<std::pair<const Lat *, const Block *>> into;
    for (auto &kv : blocks_)
        if (kv.first.second == post_id) { const Lat *A = find(kv.first.first); if (A) into.emplace_back(A, &kv.second == &it->second ? target : &kv.second); }
    std::sort(into.begin(), into.end(), [&](auto &x, auto &y) { return node_off(*x.first) < node_off(*y.first); });
    elems->reserve(target->pre.size());
    for (uint64_t q = 0; q < B->n; ++q) {
        const uint64_t row = B->off + q;
        const uint32_t sl = (uint32_t)(row / 32), lane = (uint32_t)(row % 32);
        uint32_t k = so[sl];
        for (auto &ab : into) {
            const uint64_t len = ab.second->row_ptr[q + 1] - ab.second->row_ptr[q];
            if (ab.second == target)
                for (uint64_t e = 0; e < len; ++e) elems->push_back((size_t)(k + e) * 32 + lane);
            k += (uint32_t)len;
        }
    }
    return SNN_OK;
}

int Engine::get_block_traces(uint64_t pre_id, uint64_t post_id, uint32_t *counter, float *dw, float *c, uint64_t nnz) {
    if (!reward_mode && !has_reward_lattices()) return fail(SNN_INVALID_ARGUMENT, "handle holds no reward-modulated lattice");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    int r = finalize_graph();
    if (r) return r;
    r = ensure_reward_arrays();
    if (r) return r;
    std::vector<size_t> el;
    Block *b = nullptr;
    r = block_elements(pre_id, post_id, &el, &b);
    if (r) return r;
    if (el.size() != nnz) return fail(SNN_SIZE_MISMATCH, "nnz mismatch");
    if (nnz == 0) return SNN_OK;
    std::vector<uint8_t> hc(rs_elems_);
    std::vector<float> hd(rs_elems_), hcc(rs_elems_);
    CK(cudaMemcpy(hc.data(), rs_counter_, rs_elems_, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    CK(cudaMemcpy(hd.data(), rs_dw_, rs_elems_ * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    CK(cudaMemcpy(hcc.data(), rs_c_, rs_elems_ * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    for (uint64_t e = 0; e < nnz; ++e) {
        if (counter) counter[e] = hc[el[e]];
        if (dw) dw[e] = hd[el[e]];
        if (c) c[e] = hcc[el[e]];
    }
    return SNN_OK;
}

int Engine::set_block_traces(uint64_t pre_id, uint64_t post_id, const float *weight, const uint32_t *counter, const float *dw, const float *c,
                             uint64_t nnz) {
    if (!reward_mode && !has_reward_lattices()) return fail(SNN_INVALID_ARGUMENT, "handle holds no reward-modulated lattice");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    int r = finalize_graph();
    if (r) return r;
    r = ensure_reward_arrays();
    if (r) return r;
    r = sync_weights_to_host();   // the host copy of the weights must be current before parts of it are overwritten
    if (r) return r;
    std::vector<size_t> el;
    Block *b = nullptr;
    r = block_elements(pre_id, post_id, &el, &b);
    if (r) return r;
    if (el.size() != nnz) return fail(SNN_SIZE_MISMATCH, "nnz mismatch");
    if (nnz == 0) return SNN_OK;
    const uint64_t wel = std::max(sell_alloc_krows_, sell_krows_) * 32;
    std::vector<uint8_t> hc(rs_elems_);
    std::vector<float> hd(rs_elems_), hcc(rs_elems_), hw(rs_elems_);
    CK(cudaMemcpy(hc.data(), rs_counter_, rs_elems_, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    CK(cudaMemcpy(hd.data(), rs_dw_, rs_elems_ * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    CK(cudaMemcpy(hcc.data(), rs_c_, rs_elems_ * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    CK(cudaMemcpy(hw.data(), wgt_, std::min<uint64_t>(rs_elems_, wel) * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    for (uint64_t e = 0; e < nnz; ++e) {
        const size_t o = el[e];
        if (weight) { hw[o] = weight[e]; if (b->kind == Block::CSR) b->w[e] = weight[e]; }
        // the canonical state concerns edges that get two calls per timestep: those of a reward-modulated lattice's own graph
        if (counter) { hc[o] = (uint8_t)counter[e]; if (counter[e] != 0u && pre_id == post_id) rs_canonical_ = false; }
        if (dw) { hd[o] = dw[e]; if (dw[e] != 0.f && pre_id == post_id) rs_canonical_ = false; }
        if (c) hcc[o] = c[e];
    }
    if (weight) CK(h2d_sync(wgt_, hw.data(), wel * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    if (counter) CK(h2d_sync(rs_counter_, hc.data(), rs_elems_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    if (dw) CK(h2d_sync(rs_dw_, hd.data(), rs_elems_ * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    if (c) CK(h2d_sync(rs_c_, hcc.data(), rs_elems_ * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    if (weight && b->kind == Block::GRID) dev_weights_newer_ = true;   // a stencil block's weights live on the device
    return SNN_OK;
}

int Engine::get_connection_traces(uint32_t *counter, float *dw, float *c, uint64_t nnz) {
    const Lat *only = nullptr;
    for (auto &L : lats_) if (!L.is_train) only = &L;
    if (!only || !reward_mode) return fail(SNN_INVALID_ARGUMENT, "handle is not a reward-modulated lattice");
    return get_block_traces(only->id, only->id, counter, dw, c, nnz);
}

int Engine::set_connection_traces(const float *weight, const uint32_t *counter, const float *dw, const float *c, uint64_t nnz) {
    const Lat *only = nullptr;
    for (auto &L : lats_) if (!L.is_train) only = &L;
    if (!only || !reward_mode) return fail(SNN_INVALID_ARGUMENT, "handle is not a reward-modulated lattice");
    return set_block_traces(only->id, only->id, weight, counter, dw, c, nnz);
}

int Engine::sync_weights_to_host() {
    if (!dev_weights_newer_) return SNN_OK;
    if (graph_dirty_ || !col_) { dev_weights_newer_ = false; return SNN_OK; }
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    std::vector<uint32_t> so((size_t)n_slices_ + 1), col(std::max<uint64_t>(sell_krows_ * 32, 1));
    std::vector<float> wg(std::max<uint64_t>(sell_krows_ * 32, 1));
    CK(cudaMemcpy(so.data(), slice_off_, so.size() * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    if (sell_krows_) {
        CK(cudaMemcpy(col.data(), col_, sell_krows_ * 32 * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
        CK(cudaMemcpy(wg.data(), wgt_, sell_krows_ * 32 * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    }
    for (auto &kv : blocks_)
        if (kv.second.kind == Block::GRID) materialize_grid(kv.second, *find(kv.first.first));
    // walk every row in the same order finalize_graph wrote it
    for (auto &L : lats_) {
        if (L.is_train) continue;
        std::vector<std::pair<const Lat *, Block *>> into;
        for (auto &kv : blocks_)
            if (kv.first.second == L.id) { const Lat *A = find(kv.first.first); if (A) into.emplace_back(A, &kv.second); }
        std::sort(into.begin(), into.end(), [&](auto &x, auto &y) { return node_off(*x.first) < node_off(*y.first); });
        for (uint64_t q = 0; q < L.n; ++q) {
            const uint64_t row = L.off + q;
            const uint32_t s = (uint32_t)(row / 32), lane = (uint32_t)(row % 32);
            uint32_t k = so[s];
            for (auto &ab : into) {
                Block &b = *ab.second;
                for (uint64_t e = b.row_ptr[q]; e < b.row_ptr[q + 1]; ++e, ++k) b.w[e] = wg[(size_t)k * 32 + lane];
            }
        }
    }
    dev_weights_newer_ = false;
    return SNN_OK;
}

int Engine::connection_nnz(uint64_t pre_id, uint64_t post_id, uint64_t *nnz) {
    Lat *A = find(pre_id), *B = find(post_id);
    if (!A || !B) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    auto it = blocks_.find({pre_id, post_id});
    if (it == blocks_.end()) { *nnz = 0; return SNN_OK; }
    if (it->second.kind == Block::GRID) {
        int r = sync_weights_to_host();
        if (r) return r;
        if (it->second.kind == Block::GRID) materialize_grid(it->second, *A);
    }
    *nnz = it->second.pre.size();
    return SNN_OK;
}

int Engine::get_connection_csr(uint64_t pre_id, uint64_t post_id, uint64_t *row_ptr, uint32_t *pre, float *weights,
                               uint64_t n_post, uint64_t nnz) {
    Lat *A = find(pre_id), *B = find(post_id);
    if (!A || !B || B->is_train) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    if (n_post != B->n) return fail(SNN_GRAPH_DIMENSIONS_DO_NOT_MATCH, "Dimensions do not match");
    int r = sync_weights_to_host();
    if (r) return r;
    auto it = blocks_.find({pre_id, post_id});
    if (it == blocks_.end()) {
        if (nnz) return fail(SNN_SIZE_MISMATCH, "nnz mismatch");
        if (row_ptr) std::fill(row_ptr, row_ptr + n_post + 1, 0);
        return SNN_OK;
    }
    Block &b = it->second;
    if (b.kind == Block::GRID) {
        // materialise on a copy so that the device fast path stays available
        Block tmp = b;
        materialize_grid(tmp, *A);
        if (tmp.pre.size() != nnz) return fail(SNN_SIZE_MISMATCH, "nnz mismatch");
        if (row_ptr) memcpy(row_ptr, tmp.row_ptr.data(), (n_post + 1) * 8);
        if (pre) memcpy(pre, tmp.pre.data(), nnz * 4);
        if (weights) memcpy(weights, tmp.w.data(), nnz * 4);
        return SNN_OK;
    }
    if (b.pre.size() != nnz) return fail(SNN_SIZE_MISMATCH, "nnz mismatch");
    if (row_ptr) memcpy(row_ptr, b.row_ptr.data(), (n_post + 1) * 8);
    if (pre && nnz) memcpy(pre, b.pre.data(), nnz * 4);
    if (weights && nnz) memcpy(weights, b.w.data(), nnz * 4);
    return SNN_OK;
}

int Engine::get_connection_dense(uint64_t pre_id, uint64_t post_id, uint32_t *connections, float *weights, uint64_t n_pre,
                                 uint64_t n_post) {
    Lat *A = find(pre_id), *B = find(post_id);
    if (!A || !B || B->is_train) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    if (n_pre != A->n || n_post != B->n) return fail(SNN_GRAPH_DIMENSIONS_DO_NOT_MATCH, "Dimensions do not match");
    if (part_world > 1) return fail(SNN_UNSUPPORTED, "dense graphs are not supported on partitioned handles");
    int r = sync_weights_to_host();
    if (r) return r;
    if (connections) std::fill(connections, connections + n_pre * n_post, 0u);
    if (weights) std::fill(weights, weights + n_pre * n_post, 0.f);
    auto it = blocks_.find({pre_id, post_id});
    if (it == blocks_.end()) return SNN_OK;
    Block tmp;
    const Block *b = &it->second;
    if (b->kind == Block::GRID) { tmp = *b; materialize_grid(tmp, *A); b = &tmp; }
    for (uint64_t q = 0; q < n_post; ++q)
        for (uint64_t k = b->row_ptr[q]; k < b->row_ptr[q + 1]; ++k) {
            const uint64_t idx = (uint64_t)b->pre[k] * n_post + q;
            if (connections) connections[idx] = 1;
            if (weights) weights[idx] = b->w[k];
        }
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// per-row graph access on the device table (Graph::get_incoming_connections / lookup_weight / edit_weight)
// ------------------------------------------------------------------------------------------------
int Engine::get_connection_rows(uint64_t pre_id, uint64_t post_id, uint64_t row_begin, uint64_t row_end, uint64_t *row_ptr, uint32_t *pre,
                                float *weights, uint64_t capacity, uint64_t *nnz_out) {
    Lat *A = find(pre_id), *B = find(post_id);
    if (!A || !B || B->is_train) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    if (row_begin > row_end || row_end > B->n) return fail(SNN_GRAPH_POSITION_NOT_FOUND, "Position not found, position: " + std::to_string(row_end));
    if (nnz_out) *nnz_out = 0;
    if (row_ptr) row_ptr[0] = 0;
    if (row_begin == row_end) return SNN_OK;
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    int r = finalize_graph();
    if (r) return r;
    const uint64_t g0 = B->off + row_begin, g1 = B->off + row_end;   // local neuron numbers
    const uint32_t s0 = (uint32_t)(g0 / 32), s1 = (uint32_t)((g1 - 1) / 32);
    std::vector<uint32_t> so(s1 - s0 + 2);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    CK(cudaMemcpy(so.data(), slice_off_ + s0, so.size() * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    const uint64_t k0 = so.front(), k1 = so.back();
    std::vector<uint32_t> col(std::max<uint64_t>((k1 - k0) * 32, 1));
    std::vector<float> wg(col.size());
    if (k1 > k0) {
        CK(cudaMemcpy(col.data(), col_ + k0 * 32, (k1 - k0) * 32 * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
        CK(cudaMemcpy(wg.data(), wgt_ + k0 * 32, (k1 - k0) * 32 * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    }
    // node range of the presynaptic lattice; partitioned handles report GLOBAL flat indices (ghost nodes included)
    const int64_t a0 = part_world > 1 ? 0 : (int64_t)node_off(*A), a1 = part_world > 1 ? (int64_t)n_nodes_ : a0 + (int64_t)A->n;
    const int64_t shift = part_world > 1 ? (int64_t)row0_global * A->cols - (int64_t)own0_ : -a0;
    uint64_t o = 0;
    for (uint64_t g = g0; g < g1; ++g) {
        const uint32_t s = (uint32_t)(g / 32), lane = (uint32_t)(g % 32);
        for (uint64_t k = so[s - s0]; k < so[s - s0 + 1]; ++k) {
            const size_t e = (size_t)(k - k0) * 32 + lane;
            if (col[e] == kColPad) break;   // valid entries first
            const int64_t j = (int64_t)(col[e] & kColIdxMask);
            if (j < a0 || j >= a1) continue;
            if (o < capacity) { if (pre) pre[o] = gpart_ ? (uint32_t)node_to_global((uint32_t)j) : (uint32_t)(j + shift); if (weights) weights[o] = wg[e]; }
            ++o;
        }
        if (row_ptr) row_ptr[g - g0 + 1] = o;
    }
    if (nnz_out) *nnz_out = o;
    if (o > capacity && (pre || weights)) return fail(SNN_SIZE_MISMATCH, "edge buffers too small: " + std::to_string(o) + " edges in the row range");
    return SNN_OK;
}

int Engine::check_edge_endpoints(uint64_t pre_id, uint64_t post_id, uint64_t pre, uint64_t post, Lat **A, Lat **B) {
    *A = find(pre_id); *B = find(post_id);
    if (!*A || !*B || (*B)->is_train) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    // lookup_weight / edit_weight check the postsynaptic key first (graph/mod.rs:197-202, 209-214)
    if (post >= (*B)->n) return fail(SNN_GRAPH_POSTSYNAPTIC_NOT_FOUND, "Postsynaptic position not found, position: " + std::to_string(post));
    const uint64_t pre_limit = part_world > 1 ? (uint64_t)rows_global * (*A)->cols : (*A)->n;
    if (pre >= pre_limit) return fail(SNN_GRAPH_PRESYNAPTIC_NOT_FOUND, "Presynaptic position not found, position: " + std::to_string(pre));
    return SNN_OK;
}

int Engine::find_edge(uint64_t row, uint32_t j, int64_t *elem) {
    *elem = -1;
    const uint32_t s = (uint32_t)(row / 32), lane = (uint32_t)(row % 32);
    uint32_t so[2];
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    CK(cudaMemcpy(so, slice_off_ + s, 8, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    if (so[1] == so[0]) return SNN_OK;
    std::vector<uint32_t> col((size_t)(so[1] - so[0]) * 32);
    CK(cudaMemcpy(col.data(), col_ + (size_t)so[0] * 32, col.size() * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    for (uint32_t k = 0; k < so[1] - so[0]; ++k) {
        const uint32_t c = col[(size_t)k * 32 + lane];
        if (c == kColPad) break;
        if ((c & kColIdxMask) == j) { *elem = ((int64_t)so[0] + k) * 32 + lane; break; }
    }
    return SNN_OK;
}

int Engine::lookup_weight(uint64_t pre_id, uint64_t post_id, uint64_t pre, uint64_t post, float *weight, int32_t *connected) {
    Lat *A, *B;
    int r = check_edge_endpoints(pre_id, post_id, pre, post, &A, &B);
    if (r) return r;
    *connected = 0; *weight = 0.f;
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    r = finalize_graph();
    if (r) return r;
    const int64_t node = part_world > 1 ? global_to_node(pre) : (int64_t)node_off(*A) + (int64_t)pre;
    if (node < 0 || node >= (int64_t)n_nodes_) return SNN_OK;   // partitioned: not among this rank's ghosts, cannot be connected
    int64_t e = -1;
    r = find_edge(B->off + post, (uint32_t)node, &e);
    if (r || e < 0) return r;
    CK(cudaMemcpy(weight, wgt_ + e, 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
    *connected = 1;
    return SNN_OK;
}

int Engine::edit_weight(uint64_t pre_id, uint64_t post_id, uint64_t pre, uint64_t post, bool has, float weight) {
    Lat *A, *B;
    int r = check_edge_endpoints(pre_id, post_id, pre, post, &A, &B);
    if (r) return r;
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    r = finalize_graph();
    if (r) return r;
    const int64_t node = part_world > 1 ? global_to_node(pre) : (int64_t)node_off(*A) + (int64_t)pre;
    int64_t e = -1;
    if (node >= 0 && node < (int64_t)n_nodes_) { r = find_edge(B->off + post, (uint32_t)node, &e); if (r) return r; }
    if (e >= 0 && has) {
        // Some(w) over Some(_): the adjacency is unchanged, one word on the device
        CK(h2d_sync(wgt_ + e, &weight, 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
        dev_weights_newer_ = true;
        return SNN_OK;
    }
    if (e < 0 && !has) return SNN_OK;   // None over None
    // the adjacency changes (None -> Some, Some -> None): edit the host CSR of the block and rebuild the device table lazily
    if (part_world > 1) return fail(SNN_UNSUPPORTED, "adding or removing edges is not supported on partitioned handles (weights of existing edges are)");
    dev_weights_newer_ = true;   // the device holds the live weights (STDP, earlier in-place edits)
    r = sync_weights_to_host();
    if (r) return r;
    Block &b = blocks_[{pre_id, post_id}];
    if (b.kind == Block::GRID) materialize_grid(b, *A);
    b.from_grid_radius = 0;   // no longer the generator's adjacency
    if (b.row_ptr.size() != B->n + 1) b.row_ptr.assign(B->n + 1, 0);
    const uint64_t rs = b.row_ptr[post], re = b.row_ptr[post + 1];
    uint64_t pos = rs;
    while (pos < re && b.pre[pos] < pre) ++pos;
    if (has) {
        b.pre.insert(b.pre.begin() + pos, (uint32_t)pre);
        b.w.insert(b.w.begin() + pos, weight);
        for (uint64_t q = post + 1; q <= B->n; ++q) b.row_ptr[q] += 1;
    } else {
        b.pre.erase(b.pre.begin() + pos);
        b.w.erase(b.w.begin() + pos);
        for (uint64_t q = post + 1; q <= B->n; ++q) b.row_ptr[q] -= 1;
    }
    graph_dirty_ = true;
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// options
// ------------------------------------------------------------------------------------------------
int Engine::set_dt(float dt) {
    // LatticeNetwork::set_dt neuron/mod.rs:1655-1660; Lattice::set_dt :649-652; PoissonNeuron::set_dt
    // rescales chance_of_firing (spike_train/mod.rs:345-349)
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    for (auto &L : lats_) {
        if (L.n == 0) { if (!L.is_train) { L.stdp.dt = dt; if (L.is_reward) L.rstdp.dt = dt; } continue; }
        if (L.is_train) {
            std::vector<float> old(L.n), ch(L.n);
            CK(cudaMemcpy(old.data(), TF_[TF_DT] + L.off, L.n * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
            if (train_kind == SNN_TRAIN_POISSON) {
                CK(cudaMemcpy(ch.data(), TF_[TF_CHANCE] + L.off, L.n * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
                for (uint64_t i = 0; i < L.n; ++i) { const float scalar = dt / old[i]; ch[i] *= scalar; }
                CK(h2d_sync(TF_[TF_CHANCE] + L.off, ch.data(), L.n * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
            }
            CK(fill_f32(TF_[TF_DT] + L.off, dt, L.n, stream_), SNN_GPU_QUEUE_FAILURE);
        } else {
            CK(fill_f32(F_[F_DT] + L.off, dt, L.n, stream_), SNN_GPU_QUEUE_FAILURE);
            L.stdp.dt = dt;
            if (L.is_reward) L.rstdp.dt = dt;   // RewardModulatedLattice::set_dt, neuron/mod.rs:2867-2870
        }
    }
    if (reward_mode) rstdp.dt = dt;
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

int Engine::reset_timing() {
    // neuron/mod.rs:405-420, 1710-1717
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    internal_clock = 0;
    for (auto &L : lats_) L.clock = 0;
    for (int k = 0; k < 2; ++k)
        if (LFT_[k]) CK(launch_fill_u32((uint32_t *)LFT_[k], 0xFFFFFFFFu, node_cap_, stream_), SNN_GPU_QUEUE_FAILURE);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

int Engine::reset_history() {
    for (auto &L : lats_) { L.grid_history.clear(); L.spike_history.clear(); L.spike_agg.clear(); L.average_history.clear(); L.eeg_history.clear(); L.hist_len = 0; }
    return SNN_OK;
}

int Engine::history_len(uint64_t id, uint64_t *steps) const {
    const Lat *L = find(id);
    if (!L) return SNN_NET_ID_NOT_FOUND_IN_LATTICES;
    *steps = L->hist_len;
    return SNN_OK;
}

int Engine::get_grid_history(uint64_t id, float *out, uint64_t capacity) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    if (capacity < L->grid_history.size()) return fail(SNN_SIZE_MISMATCH, "history buffer too small");
    if (!L->grid_history.empty()) memcpy(out, L->grid_history.data(), L->grid_history.size() * 4);
    return SNN_OK;
}

int Engine::get_reduced_history(uint64_t id, bool eeg, float *out, uint64_t capacity) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    const std::vector<float> &h = eeg ? L->eeg_history : L->average_history;
    if (capacity < h.size()) return fail(SNN_SIZE_MISMATCH, "history buffer too small");
    if (!h.empty()) memcpy(out, h.data(), h.size() * 4);
    return SNN_OK;
}

int Engine::set_eeg_parameters(uint64_t id, float reference_voltage, float distance, float conductivity) {
    Lat *L = find(id);
    if (!L || L->is_train) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    L->eeg_ref = reference_voltage; L->eeg_dist = distance; L->eeg_cond = conductivity;
    return SNN_OK;
}

int Engine::get_spike_aggregate(uint64_t id, int64_t *out, uint64_t capacity) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    // an empty history aggregates to an empty vector in the reference; here: zeros (the shape is known)
    if (capacity < L->n) return fail(SNN_SIZE_MISMATCH, "aggregate buffer too small");
    if (!out) return L->n ? fail(SNN_INVALID_ARGUMENT, "null argument") : SNN_OK;
    if (L->spike_agg.size() == L->n) memcpy(out, L->spike_agg.data(), L->n * 8);
    else std::fill(out, out + L->n, (int64_t)0);
    return SNN_OK;
}

int Engine::get_spike_history(uint64_t id, uint8_t *out, uint64_t capacity) {
    Lat *L = find(id);
    if (!L) return fail(SNN_NET_ID_NOT_FOUND_IN_LATTICES, "Id not present in lattices");
    if (capacity < L->spike_history.size()) return fail(SNN_SIZE_MISMATCH, "history buffer too small");
    if (!L->spike_history.empty()) memcpy(out, L->spike_history.data(), L->spike_history.size());
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// the step loop
// ------------------------------------------------------------------------------------------------
int Engine::upload_lat_table() {
    LatInfo tab[kMaxLattices];
    memset(tab, 0, sizeof tab);
    int k = 0;
    for (auto &L : lats_) {
        if (L.is_train) continue;
        tab[k].base = (uint32_t)L.off; tab[k].n = (uint32_t)L.n;
        tab[k].a_plus = L.stdp.a_plus; tab[k].a_minus = L.stdp.a_minus; tab[k].tau_plus = L.stdp.tau_plus;
        tab[k].tau_minus = L.stdp.tau_minus; tab[k].dt = L.stdp.dt;
        tab[k].do_plasticity = L.do_plasticity && !L.is_reward; tab[k].grid_hist = L.grid_hist; tab[k].spike_hist = L.spike_hist;
        ++k;
    }
    if (k == 0) k = 1;
    CK(cudaMemcpyAsync(d_lat_, tab, sizeof(LatInfo) * kMaxLattices, cudaMemcpyHostToDevice, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    return SNN_OK;
}

void Engine::fill_step_params(StepParams &p) {
    memset(&p, 0, sizeof p);
    p.own0 = own0_; p.n_neurons = (uint32_t)n_neurons; p.n_nodes = n_nodes_;
    p.electrical = electrical; p.chemical = chemical;
    p.nt_used = nt_used(); p.rc_used = rc_used();
    p.ntk = ntk; p.rck = rck; p.refract = refract;
    p.slice_off = slice_off_; p.col = col_; p.wgt = wgt_;
    p.uniform_width = uniform_width_;
    p.t_stride = node_cap_; p.node_flags = node_flags_;
    for (int s = 0; s < F_COUNT; ++s) p.f[s] = F_[s];
    p.was_inc = was_inc_;
    for (int s = 0; s < NTF_COUNT; ++s) p.nt[s] = NT_[s];
    p.nt_stride = node_cap_;
    for (int s = 0; s < RCF_COUNT; ++s) p.rc[s] = RC_[s];
    p.rc_stride = neuron_cap_;
    p.train0 = train0_; p.n_trains = (uint32_t)n_trains;
    for (int s = 0; s < TF_COUNT; ++s) p.tf[s] = TF_[s];
    p.lat = d_lat_;
    int nl = 0;
    for (auto &L : lats_) if (!L.is_train) nl++;
    p.n_lat = std::max(nl, 1);
    p.halo[0] = halo_dir_[0]; p.halo[1] = halo_dir_[1];
    p.halo_done = halo_done_;
    p.halo_timeout_ns = halo_timeout_ms * 1000000ull;
    if (gpart_ && !gpeers_.empty()) {
        p.gpeers = d_gpeers_; p.n_gpeers = (uint32_t)gpeers_.size();
        p.gexp_off = d_gexp_off_; p.gexp_ent = d_gexp_ent_; p.gslice = d_gslice_; p.n_gslices = n_slices_;
    }
}

// Operand streams of the TMA-staged kernel for the current configuration; false = not eligible (use the general kernel).
bool Engine::build_tma_params(TmaParams &tp, bool ntrel, bool stdp, bool lft_pp, unsigned *grid) {
    memset(&tp, 0xFF, sizeof tp);
    if (!grid_fast_ || !uniform_width_ || n_neurons == 0) return false;
    int mode = use_tma;
    if (mode < 0) {
        const char *e = getenv("SNN_B200_TMA");
        mode = e ? atoi(e) : 1;
    }
    if (mode == 0) return false;
    // tiny lattices are launch-latency bound: the persistent kernel only pays off with at least a few tiles per SM
    if (mode < 2 && n_neurons < 64 * 1024) return false;
    uint32_t off = 0, n = 0;
    auto add = [&](const void *src, uint32_t bytes_per_tile) -> uint32_t {
        if (n >= (uint32_t)kMaxTmaStreams) return 0xFFFFFFFFu;
        tp.st[n].src = (const unsigned char *)src;
        tp.st[n].bytes_per_tile = bytes_per_tile;
        tp.st[n].smem_off = off;
        const uint32_t o = off;
        off += (uint32_t)round_up(bytes_per_tile, 128);
        ++n;
        return o;
    };
    const uint32_t fb = kTmaTile * 4;
    tp.o_v = add(V_[0] + own0_, fb);            // src patched per step (ping-pong)
    tp.o_lft = (stdp || lft_pp) ? add(LFT_[0] + own0_, fb) : 0u;
    tp.o_flags = ntrel ? add(node_flags_ + own0_, kTmaTile) : 0u;
    const uint32_t eb = kTmaConsumerWarps * uniform_width_ * 32u * 4u;
    tp.o_col = add(col_, eb);
    tp.o_wgt = add(wgt_, eb);
    for (int i = 0; i < kNumNeuronFields; ++i) {
        const FieldDef &fd = kNeuronFields[i];
        if (fd.kind != FK_NEURON_DEV || !(fd.models & SNN_M(model)) || fd.slot >= F_NA_CUR) continue;
        tp.o_f[fd.slot] = add(F_[fd.slot], fb);
    }
    if (ntrel) {
        const uint32_t ntu = nt_used(), rcu = rc_used();
        for (int ty = 0; ty < kNT; ++ty) {
            if (ntu & (1u << ty)) {
                tp.o_t[ty] = add(T_[0] + (size_t)ty * node_cap_ + own0_, fb);   // src patched per step
                tp.o_nt[NTF_TMAX][ty] = add(NT_[NTF_TMAX] + (size_t)ty * node_cap_ + own0_, fb);
                if (ntk != SNN_NT_DISCRETE_SPIKE) tp.o_nt[NTF_P1][ty] = add(NT_[NTF_P1] + (size_t)ty * node_cap_ + own0_, fb);
                if (ntk == SNN_NT_DESTEXHE) tp.o_nt[NTF_P2][ty] = add(NT_[NTF_P2] + (size_t)ty * node_cap_ + own0_, fb);
            }
            if (rcu & (1u << ty)) {
                tp.o_rc[RCF_R][ty] = add(RC_[RCF_R] + (size_t)ty * neuron_cap_, fb);
                tp.o_rc[RCF_G][ty] = add(RC_[RCF_G] + (size_t)ty * neuron_cap_, fb);
                tp.o_rc[RCF_E][ty] = add(RC_[RCF_E] + (size_t)ty * neuron_cap_, fb);
                if (rck != SNN_RC_APPROXIMATE) {
                    tp.o_rc[RCF_K1][ty] = add(RC_[RCF_K1] + (size_t)ty * neuron_cap_, fb);
                    tp.o_rc[RCF_K2][ty] = add(RC_[RCF_K2] + (size_t)ty * neuron_cap_, fb);
                }
                if (ty == SNN_NT_NMDA) tp.o_rc[RCF_MG][ty] = add(RC_[RCF_MG] + (size_t)ty * neuron_cap_, fb);
            }
        }
    }
    if (n >= (uint32_t)kMaxTmaStreams) return false;
    if (const char *ek = getenv("SNN_B200_TMA_EXPERIMENT_KEEP")) n = std::min<uint32_t>(n, (uint32_t)atoi(ek));  // timing experiment only: results are wrong
    tp.n_streams = n;
    tp.stage_bytes = off;
    tp.tx_bytes = 0;
    for (uint32_t k = 0; k < n; ++k) tp.tx_bytes += tp.st[k].bytes_per_tile;
    tp.n_tiles = (uint32_t)((n_neurons + kTmaTile - 1) / kTmaTile);
    // shared-memory budget: as many resident CTAs per SM as the kernel's register budget allows (3, or 2 for
    // Hodgkin-Huxley / general neurotransmitter masks: tma_min_ctas in step_tma.cu) with at least two stages each
    const uint32_t budget = 220 * 1024;
    uint32_t ctas = 3, stages = (budget / ctas) / off;
    if (stages < 2) { ctas = 2; stages = (budget / ctas) / off; }
    const char *es = getenv("SNN_B200_TMA_STAGES"), *ec = getenv("SNN_B200_TMA_CTAS");
    if (ec) { ctas = (uint32_t)std::max(1, atoi(ec)); stages = (budget / ctas) / off; }
    if (stages < 2) { ctas = 1; stages = budget / off; }
    if (stages < 2) return false;
    stages = std::min<uint32_t>(stages, 4);
    if (es) stages = (uint32_t)std::max(2, atoi(es));
    if ((uint64_t)stages * off + 64 > 227 * 1024) return false;
    tp.stages = stages;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    *grid = (unsigned)std::min<uint64_t>(tp.n_tiles, (uint64_t)sms * ctas);
    return true;
}

// Operand streams of the window-staged kernel (step_win.cu): radius-1 stencil lattices.  false = not eligible.
bool Engine::build_win_params(WinParams &wp, int chemg, bool ntrel, bool stdp, bool lft_pp, unsigned *grid) {
    memset(&wp, 0xFF, sizeof wp);
    if (!grid_fast_ || uniform_width_ != kWinWidth || n_neurons == 0) return false;
    int mode = use_tma;
    if (mode < 0) {
        const char *e = getenv("SNN_B200_TMA");
        mode = e ? atoi(e) : 1;
    }
    if (mode == 0) return false;
    if (const char *e = getenv("SNN_B200_WIN")) if (atoi(e) == 0) return false;
    // tiny lattices are launch-latency bound: the persistent kernel only pays off with at least a few tiles per SM
    if (mode < 2 && n_neurons < 64 * 1024) return false;
    uint32_t cols = 0;
    for (auto &L : lats_) if (!L.is_train) cols = L.cols;
    if (cols == 0) return false;
    const WinLayout L = win_layout(model, chemg, ntrel, stdp);
    uint32_t off = L.fixed_end, n = 0, tx = 0;
    auto add_at = [&](const void *src, uint32_t bytes_per_tile, uint32_t smem_off) {
        if (n >= (uint32_t)kMaxTmaStreams) { n = kMaxTmaStreams + 1; return; }
        wp.st[n].src = (const unsigned char *)src;
        wp.st[n].bytes_per_tile = bytes_per_tile;
        wp.st[n].smem_off = smem_off;
        tx += bytes_per_tile;
        ++n;
    };
    auto add = [&](const void *src, uint32_t bytes_per_tile) -> uint32_t {
        const uint32_t o = off;
        add_at(src, bytes_per_tile, o);
        off += (uint32_t)round_up(bytes_per_tile, 128);
        return o;
    };
    const uint32_t fb = kWinTile * 4;
    if (ntrel) add_at(node_flags_ + own0_, kWinTile, L.o_flags);
    add_at(col_, kWinEdgeBytes, L.o_col);
    add_at(wgt_, kWinEdgeBytes, L.o_wgt);
    for (int sl = 0; sl < F_NA_CUR; ++sl)
        if (L.o_f[sl] != 0xFFFFFFFFu) {
            if (!F_[sl]) return false;
            add_at(F_[sl], fb, L.o_f[sl]);
        }
    if (ntrel) {
        const uint32_t ntu = nt_used(), rcu = rc_used();
        for (int ty = 0; ty < kNT; ++ty) {
            if (ntu & (1u << ty)) {
                wp.o_nt[NTF_TMAX][ty] = add(NT_[NTF_TMAX] + (size_t)ty * node_cap_ + own0_, fb);
                if (ntk != SNN_NT_DISCRETE_SPIKE) wp.o_nt[NTF_P1][ty] = add(NT_[NTF_P1] + (size_t)ty * node_cap_ + own0_, fb);
                if (ntk == SNN_NT_DESTEXHE) wp.o_nt[NTF_P2][ty] = add(NT_[NTF_P2] + (size_t)ty * node_cap_ + own0_, fb);
            }
            if (rcu & (1u << ty)) {
                wp.o_rc[RCF_R][ty] = add(RC_[RCF_R] + (size_t)ty * neuron_cap_, fb);
                wp.o_rc[RCF_G][ty] = add(RC_[RCF_G] + (size_t)ty * neuron_cap_, fb);
                wp.o_rc[RCF_E][ty] = add(RC_[RCF_E] + (size_t)ty * neuron_cap_, fb);
                if (rck != SNN_RC_APPROXIMATE) {
                    wp.o_rc[RCF_K1][ty] = add(RC_[RCF_K1] + (size_t)ty * neuron_cap_, fb);
                    wp.o_rc[RCF_K2][ty] = add(RC_[RCF_K2] + (size_t)ty * neuron_cap_, fb);
                }
                if (ty == SNN_NT_NMDA) wp.o_rc[RCF_MG][ty] = add(RC_[RCF_MG] + (size_t)ty * neuron_cap_, fb);
            }
        }
    }
    if (n > (uint32_t)kMaxTmaStreams) return false;
    wp.n_streams = n;
    wp.stage_bytes = off;
    wp.fixed_tx_bytes = tx;
    wp.n_tiles = (uint32_t)((n_neurons + kWinTile - 1) / kWinTile);
    wp.node_cap = (uint32_t)node_cap_;
    wp.cols = cols;
    wp.lft_copy = (stdp || lft_pp) ? 1u : 0u;
    // tiles whose windows reach ghost rows (they are also the ones that export): same predicates as the producer's waits
    wp.first_lo = wp.first_hi = wp.bnd_lo = wp.bnd_hi = 0;
    if (part_world > 1) {
        for (uint32_t t = 0; t < wp.n_tiles; ++t) {
            const uint64_t ts = (uint64_t)t * kWinTile;
            if (halo_dir_[0].active && ts <= cols) wp.first_lo++;
            if (halo_dir_[1].active && ts + kWinTile + 1u + cols > n_neurons) wp.first_hi++;
        }
        wp.bnd_lo = wp.first_lo; wp.bnd_hi = wp.first_hi;
        if (wp.first_lo + wp.first_hi > wp.n_tiles) wp.first_lo = wp.first_hi = 0;   // a strip of a few rows: every tile is a boundary tile
        if (const char *e = getenv("SNN_B200_BOUNDARY_FIRST")) if (atoi(e) == 0) wp.first_lo = wp.first_hi = 0;
    }
    const uint32_t G = (uint32_t)win_groups(model, chemg);
    uint32_t stages = (224u * 1024u) / (off + 16u);
    if (const char *es = getenv("SNN_B200_WIN_STAGES")) stages = std::min<uint32_t>(stages, (uint32_t)std::max(1, atoi(es)));
    stages = std::min<uint32_t>(stages, 8);
    if (stages < G) return false;   // the consumer groups need a stage each
    wp.stages = stages;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    *grid = (unsigned)std::min<uint64_t>((wp.n_tiles + G - 1) / G, (uint64_t)sms);
    if (*grid == 0) *grid = 1;
    return true;
}

// Partition handles of ONE process (attach_local / gpart_attach_local) are stepped from one host thread each, and their kernels
// wait for each other on the device.  CUDA only guarantees progress if every kernel a kernel waits for was launched EARLIER:
// streams that happen to share a hardware work queue execute in launch order, so a waiting kernel queued ahead of the kernel it
// waits for would never see it start (observed as time-outs with 4 handles on one device).  Hence: before launching a kernel
// that waits for the counter value `need`, wait on the host until every local peer has launched the kernel that raises it.
void Engine::wait_local_peers_launched(unsigned long long need, bool second) {
    for (Engine *q : local_peers_) {
        const std::atomic<unsigned long long> &c = second ? q->launched_pub2_ : q->launched_pub_;
        const auto t0 = std::chrono::steady_clock::now();
        while (c.load(std::memory_order_acquire) < need) {
            std::this_thread::sleep_for(std::chrono::microseconds(5));
            // a peer that never runs: fall through after the halo time-out, the device-side wait then reports the error
            if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() > (double)halo_timeout_ms) return;
        }
    }
}

int Engine::run(uint64_t iterations, float *elapsed_ms, uint64_t *launches, const float *rewards) {
    if (elapsed_ms) *elapsed_ms = 0.f;
    if (launches) *launches = 0;
    // gpu_lattices/mod.rs:1089-1091: empty lattice or zero iterations -> Ok(()); neuron/mod.rs:1217: both flags off -> Ok(())
    if (iterations == 0 || (n_neurons == 0 && n_trains == 0)) return SNN_OK;
    if (!electrical && !chemical) return SNN_OK;
    if (internal_clock + iterations > 0x7FFFFFF0ull) return fail(SNN_UNSUPPORTED, "internal clock would overflow the i32 last_firing_time");
    if (part_world > 1 && !gpart_) {
        if ((part_rank > 0 && !halo_dir_[0].active) || (part_rank < part_world - 1 && !halo_dir_[1].active))
            return fail(SNN_INVALID_ARGUMENT, "partitioned handle: neighbouring strips are not attached");
    }
    if (gpart_ && (reward_mode || bcm_mode)) return fail(SNN_UNSUPPORTED, "general-graph partitions step plain and STDP lattices");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    if (chemical) { int r = ensure_chem(); if (r) return r; }
    int r = finalize_graph();
    if (r) return r;
    r = upload_lat_table();
    if (r) return r;
    if (gpart_) { r = gpart_build_device(); if (r) return r; }

    bool stdp = false, want_grid = false, want_spk = false, want_tgrid = false, want_tspk = false, want_red = false;
    for (auto &L : lats_) {
        if (!L.is_train) { stdp |= L.do_plasticity && !L.is_reward; want_grid |= L.grid_hist; want_spk |= L.spike_hist; want_red |= L.avg_hist || L.eeg_hist; }
        else { want_tgrid |= L.grid_hist; want_tspk |= L.spike_hist; }
    }
    // kernel specialisation: ntrel = neurotransmitter / receptor state must be stepped; chemg = how the chemical gather
    // reads presynaptic types (0 none, 1 a single type in the whole node array, 3 per-edge masks); net = several
    // lattices and/or spike trains share the node array
    const bool ntrel = chem_alloc_ && (chemical || nt_used() != 0 || (model == SNN_MODEL_HODGKIN_HUXLEY && rc_used() != 0));
    int n_neuron_lat = 0;
    for (auto &L : lats_) if (!L.is_train) n_neuron_lat++;
    const bool net = n_neuron_lat > 1 || n_trains > 0;
    int chemg = 0;
    if (chemical && ntrel && nt_used() != 0) chemg = (!net && __builtin_popcount(nt_used()) == 1) ? 1 : 3;
    // reward-modulated lattice: the modulator replaces the plasticity rule and needs last_firing_time before and after a step
    const bool rmod = reward_mode && do_modulation && n_neurons > 0;
    // BCM rule instead of STDP: applied by its own per-edge kernel right after each step (it reads the activities the step wrote)
    const bool bcm_on = bcm_mode && stdp && !reward_mode && n_neurons > 0;
    if (reward_mode || bcm_mode) stdp = false;
    if (rmod) { int rr = ensure_reward_arrays(); if (rr) return rr; }
    RstdpParams rsp{rstdp.dopamine, rstdp.tau_c, rstdp.a_plus, rstdp.a_minus, rstdp.tau_plus, rstdp.tau_minus, rstdp.dt,
                    rs_counter_, rs_dw_, rs_c_, rs_canonical_ ? 1u : 0u, nullptr, 0u};
    if (rmod && !getenv("SNN_B200_RSTDP_NOTAB")) {
        // difference table of the STDP term: long enough for the exponential to have underflowed to zero at its end
        const double reach = 110.0 * std::max((double)rstdp.tau_plus, (double)rstdp.tau_minus) / std::max((double)rstdp.dt, 1e-30);
        if (reach > 0 && reach < 60000.0 && rstdp.tau_plus > 0.f && rstdp.tau_minus > 0.f && rstdp.dt > 0.f) {
            const uint32_t tab_n = (uint32_t)reach + 8u;
            if (rs_tab_n_ != tab_n) {
                if (rs_tab_) cudaFree(rs_tab_);
                rs_tab_ = nullptr; rs_tab_n_ = 0;
                CK(dev_alloc(&rs_tab_, (size_t)tab_n * 2), SNN_GPU_BUFFER_CREATE_ERROR);
                rs_tab_n_ = tab_n;
            }
            CK(launch_rstdp_table(rsp, rs_tab_, tab_n, stream_), SNN_GPU_QUEUE_FAILURE);   // the parameters may have changed since the last run
            rsp.tab = rs_tab_; rsp.tab_n = tab_n;
        }
    }
    // RewardModulatedLatticeNetwork: reward-modulated lattices next to plain lattices and spike trains
    const bool net_reward = has_reward_lattices();
    bool net_rmod = false;
    RnetParams rnp;
    memset(&rnp, 0, sizeof rnp);
    if (net_reward) {
        if (part_world > 1 || reward_mode || bcm_mode) return fail(SNN_UNSUPPORTED, "reward-modulated lattice networks run on single-GPU network handles");
        // the configurations the reference cannot run itself: update_weights_from_neurons_across_(reward_)lattices looks the OUTGOING
        // connecting edges up with the end points swapped and unwraps the result (neuron/mod.rs:4760-4763, 4931-4934), and its
        // incoming arms unwrap lattice kinds that need not be there (:4727-4731, 4741-4747)
        for (auto &kv : blocks_) {
            if (kv.first.first == kv.first.second) continue;
            if (kv.second.kind != Block::GRID && kv.second.pre.empty()) continue;
            const Lat *A = find(kv.first.first), *B = find(kv.first.second);
            if (!A || !B) continue;
            if (A->is_reward && A->do_modulation)
                return fail(SNN_UNSUPPORTED, "connecting edges out of a reward-modulated lattice with do_modulation: the reference panics on them (neuron/mod.rs:4931-4934)");
            if (!A->is_train && !A->is_reward && A->do_plasticity)
                return fail(SNN_UNSUPPORTED, "connecting edges out of a plastic lattice of a reward-modulated network: the reference panics on them (neuron/mod.rs:4760-4763)");
            if (!B->is_reward && kv.second.reward_conn)
                return fail(SNN_UNSUPPORTED, "RewardModulatedWeight connections must end in a reward-modulated lattice");
            if (!B->is_reward && B->do_plasticity && A->is_reward)
                return fail(SNN_UNSUPPORTED, "a plastic lattice fed by a reward-modulated lattice: the reference panics on it (neuron/mod.rs:4727-4731)");
        }
        std::map<uint64_t, int> cls;   // lattice id -> presynaptic class of RnetParams::kinds
        std::vector<const Lat *> nl;
        int kn = 0, kt = 0;
        for (auto &L : lats_) {
            if (L.is_train) { rnp.tl_base[kt] = (uint32_t)L.off; cls[L.id] = kMaxLattices + kt; ++kt; continue; }
            RnetLat &R = rnp.lat[kn];
            R.dopamine = L.rstdp.dopamine; R.tau_c = L.rstdp.tau_c; R.a_plus = L.rstdp.a_plus; R.a_minus = L.rstdp.a_minus;
            R.tau_plus = L.rstdp.tau_plus; R.tau_minus = L.rstdp.tau_minus; R.dt = L.rstdp.dt;
            R.flags = (L.is_reward ? 1u : 0u) | (L.is_reward && L.do_modulation ? 2u : 0u);
            net_rmod |= L.is_reward && L.do_modulation && L.n > 0;
            rnp.nbase[kn] = (uint32_t)L.off;
            nl.push_back(&L);
            cls[L.id] = kn++;
        }
        rnp.n_lat = (uint32_t)kn; rnp.nbase[kn] = (uint32_t)n_neurons;
        rnp.n_tl = (uint32_t)kt; rnp.tl_base[kt] = (uint32_t)n_trains; rnp.train0 = train0_;
        for (int post = 0; post < kn; ++post) {
            if (!nl[post]->is_reward) continue;
            uint64_t k = 1ull << (2 * post);                                 // own graph
            for (int pre = 0; pre < kn; ++pre)
                if (pre != post && !nl[pre]->is_reward) k |= 3ull << (2 * pre);   // Weight block fed by a plain lattice (the default kind)
            rnp.kinds[post] = k;
        }
        for (auto &kv : blocks_)
            if (kv.second.reward_conn && kv.first.first != kv.first.second && cls.count(kv.first.first) && cls.count(kv.first.second)) {
                uint64_t &k = rnp.kinds[cls[kv.first.second]];
                const int q = cls[kv.first.first];
                k = (k & ~(3ull << (2 * q))) | (2ull << (2 * q));
            }
        if (net_rmod) { int rr = ensure_reward_arrays(); if (rr) return rr; }
        rnp.counter = rs_counter_; rnp.dw = rs_dw_; rnp.c = rs_c_;
        rnp.canonical = (rs_canonical_ && !getenv("SNN_B200_RNET_NOCANON")) ? 1u : 0u;
        if (net_rmod && !getenv("SNN_B200_RSTDP_NOTAB")) {
            // one difference table per neuron lattice: the modulator's STDP term for a reward-modulated lattice, the lattice's own
            // STDP for a plain one (what Weight edges out of it use); long enough for every exponential to have underflowed
            double reach = 0.0;
            bool ok = true;
            for (const Lat *L : nl) {
                const float tp_ = L->is_reward ? L->rstdp.tau_plus : L->stdp.tau_plus, tm_ = L->is_reward ? L->rstdp.tau_minus : L->stdp.tau_minus;
                const float dt_ = L->is_reward ? L->rstdp.dt : L->stdp.dt;
                ok = ok && tp_ > 0.f && tm_ > 0.f && dt_ > 0.f;
                reach = std::max(reach, 110.0 * std::max((double)tp_, (double)tm_) / std::max((double)dt_, 1e-30));
            }
            if (ok && reach > 0 && reach < 60000.0) {
                const uint32_t tab_n = (uint32_t)reach + 8u;
                const size_t need = (size_t)kn * 2u * tab_n;
                if (rnet_tab_elems_ != need) {
                    if (rnet_tab_) cudaFree(rnet_tab_);
                    rnet_tab_ = nullptr; rnet_tab_elems_ = 0;
                    CK(dev_alloc(&rnet_tab_, need), SNN_GPU_BUFFER_CREATE_ERROR);
                    rnet_tab_elems_ = need;
                }
                for (int l = 0; l < kn; ++l) {   // the parameters may have changed since the last run
                    const Lat &L = *nl[l];
                    RstdpParams tpar{};
                    if (L.is_reward) { tpar.a_plus = L.rstdp.a_plus; tpar.a_minus = L.rstdp.a_minus; tpar.tau_plus = L.rstdp.tau_plus; tpar.tau_minus = L.rstdp.tau_minus; tpar.dt = L.rstdp.dt; }
                    else { tpar.a_plus = L.stdp.a_plus; tpar.a_minus = L.stdp.a_minus; tpar.tau_plus = L.stdp.tau_plus; tpar.tau_minus = L.stdp.tau_minus; tpar.dt = L.stdp.dt; }
                    CK(launch_rstdp_table(tpar, rnet_tab_ + (size_t)l * 2u * tab_n, tab_n, stream_), SNN_GPU_QUEUE_FAILURE);
                }
                rnp.tab = rnet_tab_; rnp.tab_n = tab_n;
            }
        }
    }
    const bool lft_pp = stdp || rmod || net_rmod || (n_trains && electrical) || part_world > 1;
    const bool rmod_part = rmod && part_world > 1;

    StepParams sp;
    fill_step_params(sp);
    sp.lft_pp = lft_pp;
    if (rmod_part)   // step s + 1 overwrites ghost slots that the neighbour's per-edge kernel of step s still reads: wait for that one
        for (int d = 0; d < 2; ++d) sp.halo[d].my_flag = halo_dir_[d].my_flag2;
    TmaParams tma;
    unsigned tma_grid = 0;
    WinParams win;
    unsigned win_grid = 0;
    const bool win_ok = !net && build_win_params(win, chemg, ntrel, stdp, lft_pp, &win_grid);
    const bool tma_ok = !win_ok && !net && build_tma_params(tma, ntrel, stdp, lft_pp, &tma_grid);
    // wide rows (hundreds of in-edges per neuron, e.g. all-to-all spike-train input): one CTA per slice
    bool wide = !win_ok && !tma_ok && part_world == 1 && n_slices_ > 0 && sell_krows_ / n_slices_ >= kWideMinWidth;
    if (const char *e = getenv("SNN_B200_WIDE")) wide = wide && atoi(e) != 0;
    // very wide slices (hundreds of k-rows): one SM per slice cannot issue the per-edge instructions fast enough — two passes, the
    // per-edge terms of every (slice, chunk) on its own CTA, then the ordered sums (step_wide.cu)
    bool wide_split = false;
    uint32_t wide_chunks_cap = 0;
    if (wide) {
        const uint64_t mean = sell_krows_ / n_slices_;
        wide_split = mean >= 2 * wide_chunk_krows();
        if (const char *e = getenv("SNN_B200_WIDE_SPLIT")) wide_split = atoi(e) != 0;
        if (wide_split) {
            std::vector<uint32_t> so((size_t)n_slices_ + 1);
            CK(cudaMemcpy(so.data(), slice_off_, so.size() * 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
            uint32_t mx = 0;
            for (uint32_t sl = 0; sl < n_slices_; ++sl) mx = std::max(mx, so[sl + 1] - so[sl]);
            wide_chunks_cap = (mx + wide_chunk_krows() - 1) / wide_chunk_krows();
            const size_t need = (size_t)n_slices_ * wide_chunks_cap * wide_chunk_bytes(chemg) + (size_t)n_slices_ * wide_part_bytes();
            if (need != wide_scratch_bytes_) {
                if (wide_scratch_) cudaFree(wide_scratch_);
                wide_scratch_ = nullptr; wide_scratch_bytes_ = 0;
                CK(cudaMalloc((void **)&wide_scratch_, need), SNN_GPU_BUFFER_CREATE_ERROR);
                wide_scratch_bytes_ = need;
            }
            // the arrival counters of the sum pass start at zero (they reset themselves after every timestep)
            CK(cudaMemsetAsync(wide_scratch_ + (size_t)n_slices_ * wide_chunks_cap * wide_chunk_bytes(chemg), 0, (size_t)n_slices_ * wide_part_bytes(), stream_),
               SNN_GPU_BUFFER_WRITE_ERROR);
        }
    }
    // indices of the ping-ponged streams, patched every step
    int tma_iv = -1, tma_il = -1, tma_it[kNT] = {-1, -1, -1};
    if (tma_ok) {
        for (uint32_t k = 0; k < tma.n_streams; ++k) {
            if (tma.st[k].smem_off == tma.o_v) tma_iv = (int)k;
            if ((stdp || lft_pp) && tma.st[k].smem_off == tma.o_lft) tma_il = (int)k;
            for (int ty = 0; ty < kNT; ++ty) if (tma.o_t[ty] != 0xFFFFFFFFu && tma.st[k].smem_off == tma.o_t[ty]) tma_it[ty] = (int)k;
        }
    }
    TrainParams tp;
    memset(&tp, 0, sizeof tp);
    if (n_trains) {
        tp.train0 = train0_; tp.n_trains = (uint32_t)n_trains; tp.kind = train_kind; tp.ntk = ntk; tp.seed = seed;
        tp.t_stride = node_cap_; tp.node_flags = node_flags_; tp.nt_stride = node_cap_;
        for (int s = 0; s < NTF_COUNT; ++s) tp.nt[s] = NT_[s];
        for (int s = 0; s < TF_COUNT; ++s) tp.tf[s] = TF_[s];
        tp.ft_off = ft_off_; tp.ft = ft_; tp.lft_pp = lft_pp;
        if (!chem_alloc_) {  // train kernel reads node_flags only; nt pointers unused when no types are present
        }
    }

    // history staging: chunks of steps recorded on the device, drained to the host between chunks
    const uint64_t n_words = (n_neurons + 31) / 32, t_words = (n_trains + 31) / 32;
    uint64_t per_step_bytes = 0;
    // AverageVoltageHistory / EEGHistory are reductions of the staged grid record (it stays on the device for them)
    const bool copy_grid = want_grid;
    want_grid = want_grid || want_red;
    if (want_grid) per_step_bytes += n_neurons * 4;
    if (want_spk) per_step_bytes += n_words * 4;
    if (want_tgrid) per_step_bytes += n_trains * 4;
    if (want_tspk) per_step_bytes += t_words * 4;
    uint64_t chunk = iterations;
    if (per_step_bytes) chunk = std::max<uint64_t>(1, std::min<uint64_t>(iterations, (256ull << 20) / per_step_bytes));
    float *d_grid = nullptr, *d_tgrid = nullptr; uint32_t *d_spk = nullptr, *d_tspk = nullptr;
    if (want_grid) CK(dev_alloc(&d_grid, chunk * n_neurons), SNN_GPU_BUFFER_CREATE_ERROR);
    if (want_spk) CK(dev_alloc(&d_spk, chunk * n_words), SNN_GPU_BUFFER_CREATE_ERROR);
    if (want_tgrid) CK(dev_alloc(&d_tgrid, chunk * n_trains), SNN_GPU_BUFFER_CREATE_ERROR);
    if (want_tspk) CK(dev_alloc(&d_tspk, chunk * t_words), SNN_GPU_BUFFER_CREATE_ERROR);
    double *d_red = nullptr; uint32_t *d_red_lat = nullptr;
    std::vector<Lat *> red_lats;
    if (want_red) {
        std::vector<uint32_t> meta;
        for (auto &L : lats_) if (!L.is_train && (L.avg_hist || L.eeg_hist)) red_lats.push_back(&L);   // lats_ is not resized inside run()
        const size_t nl = red_lats.size();
        meta.resize(3 * nl);
        for (size_t k = 0; k < nl; ++k) {
            meta[k] = (uint32_t)red_lats[k]->off; meta[nl + k] = (uint32_t)red_lats[k]->n;
            memcpy(&meta[2 * nl + k], &red_lats[k]->eeg_ref, 4);
        }
        CK(dev_alloc(&d_red, chunk * nl * 2), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(dev_alloc(&d_red_lat, 3 * nl), SNN_GPU_BUFFER_CREATE_ERROR);
        CK(h2d_sync(d_red_lat, meta.data(), meta.size() * 4, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
    }
    auto free_hist = [&]() { cudaFree(d_grid); cudaFree(d_spk); cudaFree(d_tgrid); cudaFree(d_tspk); cudaFree(d_red); cudaFree(d_red_lat); };

    uint64_t n_launch = 0;
    float total_ms = 0.f;
    int status = SNN_OK;
    auto bail = [&](cudaError_t e, int st, const char *what) { status = cuda_fail(e, st, what); };

    // multi-GPU: push the current boundary state into the neighbours' ghost slots
    if (part_world > 1) {
        StepParams hp = sp;
        hp.v_in = V_[cur_]; hp.lft_in = LFT_[lft_loc_]; hp.t_in = T_[cur_]; hp.out_par = (uint32_t)cur_;
        hp.halo_epoch = halo_epoch_;
        if (lft_loc_ != cur_) {  // keep lft parity aligned with the V parity on partitioned handles
            cudaMemcpyAsync(LFT_[cur_], LFT_[lft_loc_], node_cap_ * 4, cudaMemcpyDeviceToDevice, stream_);
            lft_loc_ = cur_; hp.lft_in = LFT_[lft_loc_];
        }
        cudaError_t e = gpart_ ? launch_gpart_push(hp, stream_) : launch_halo_push(hp, stream_);
        if (e != cudaSuccess) { free_hist(); return cuda_fail(e, SNN_GPU_QUEUE_FAILURE, "halo_push"); }
        halo_epoch_ += 1;
        launched_pub_.store(halo_epoch_, std::memory_order_release);
        launched_pub2_.store(halo_epoch_, std::memory_order_release);
        n_launch++;
        // rendez-vous before any step kernel is enqueued: the in-kernel waits are bounded, so ordinary host-side skew between
        // the ranks (Python work, history drains, allocation) must be absorbed here, where nothing has been computed yet
        e = cudaStreamSynchronize(stream_);
        if (e != cudaSuccess) { free_hist(); return cuda_fail(e, SNN_GPU_WAIT_ERROR, "halo_push"); }
        const auto t_wait0 = std::chrono::steady_clock::now();
        for (;;) {
            unsigned long long f[8 + kMaxRanks];
            e = cudaMemcpy(f, flags_, sizeof f, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { free_hist(); return cuda_fail(e, SNN_GPU_BUFFER_READ_ERROR, "halo rendez-vous"); }
            bool ready = true;
            for (int d = 0; d < 2; ++d) if (halo_dir_[d].active && f[d] < halo_epoch_) ready = false;
            for (auto &G : gpeers_) if (f[8 + G.rank] < halo_epoch_) ready = false;
            if (ready) break;
            if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wait0).count() > 4.0 * (double)halo_timeout_ms) {
                free_hist();
                return fail(SNN_GPU_WAIT_ERROR, "timed out waiting for the neighbouring strips to enter run() (is every rank calling run with the same number of steps?)");
            }
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    }

    static const bool halo_nowait = getenv("SNN_B200_HALO_NOWAIT") != nullptr;
    static const bool win_reverse = !(getenv("SNN_B200_WIN_REVERSE") && atoi(getenv("SNN_B200_WIN_REVERSE")) == 0);
    // small lattices / networks (everything the staged kernels do not take): a whole chunk of timesteps per cooperative launch
    // (step_multi.cu).  SNN_OPT_STEPS_PER_GRAPH: 0 = as many as the history chunk holds, 1 = one launch per timestep, k = at most k.
    bool multi = part_world == 1 && n_neurons > 0 && !win_ok && !tma_ok && !bcm_on && !reward_mode && !net_reward && steps_per_graph != 1 &&
                 !getenv("SNN_DEBUG_TIMING");
    if (wide && wide_split) multi = false;   // the two-pass wide kernels are one-launch-per-step kernels
    if (const char *e = getenv("SNN_B200_MULTI")) multi = multi && atoi(e) != 0;
    if (multi) {
        StepParams probe = sp;
        TrainParams tprobe = tp;
        MultiParams mprobe{};
        mprobe.wide_stage = (wide && n_nodes_ <= 4096) ? 1u : 0u;
        multi = launch_step_multi(probe, tprobe, mprobe, model, chemg, ntrel, stdp, net, wide, device, true, stream_) == cudaSuccess;
        cudaGetLastError();
    }
    uint64_t done = 0;
    bool first_step = true;
    while (done < iterations && status == SNN_OK) {
        const uint64_t steps = std::min(chunk, iterations - done);
        cudaEventRecord(ev0_, stream_);
        const auto host_t0 = std::chrono::steady_clock::now();
        static const bool dbg_timing = getenv("SNN_DEBUG_TIMING") != nullptr;
        std::vector<cudaEvent_t> dbg_ev;
        unsigned long long *dbg_clk = nullptr;
        if (dbg_timing && steps <= 32) { cudaMalloc(&dbg_clk, 32 * 4 * sizeof(unsigned long long)); cudaMemsetAsync(dbg_clk, 0, 32 * 4 * sizeof(unsigned long long), stream_); }
        for (uint64_t s0 = 0; multi && s0 < steps;) {
            const uint64_t k = steps_per_graph ? std::min<uint64_t>(steps_per_graph, steps - s0) : steps - s0;
            sp.clock = (uint32_t)internal_clock;
            sp.dbg = nullptr; sp.reverse = 0u;
            if (n_trains) {
                tp.n_tl = 0;
                for (auto &L : lats_)
                    if (L.is_train) { tp.tl_base[tp.n_tl] = (uint32_t)L.off; tp.tl_clock[tp.n_tl] = (uint32_t)L.clock; tp.n_tl++; }
                tp.tl_base[tp.n_tl] = (uint32_t)n_trains;
                tp.draw = train_draws;
            }
            MultiParams mp{};
            mp.steps = (uint32_t)k; mp.first_pending = first_step ? 0u : 1u;
            mp.cur = (uint32_t)cur_; mp.lft_loc = (uint32_t)lft_loc_; mp.lft_pp = lft_pp ? 1u : 0u;
            mp.train_sync = (stdp && n_trains) ? 1u : 0u;
            for (int q = 0; q < 2; ++q) { mp.v[q] = V_[q]; mp.lft[q] = LFT_[q]; mp.spk[q] = SPK_[q]; mp.t[q] = T_[q]; }
            mp.grid_hist = want_grid ? d_grid + s0 * n_neurons : nullptr;
            mp.spike_hist = want_spk ? d_spk + s0 * n_words : nullptr;
            mp.tgrid_hist = want_tgrid ? d_tgrid + s0 * n_trains : nullptr;
            mp.tspike_hist = want_tspk ? d_tspk + s0 * t_words : nullptr;
            mp.n_neurons = n_neurons; mp.n_words = n_words; mp.n_trains = n_trains; mp.t_words = t_words;
            mp.barrier = multi_barrier_;
            mp.cache_rows = (!stdp && !getenv("SNN_B200_MULTI_NOCACHE")) ? 1u : 0u;
            mp.wide_stage = (wide && n_nodes_ <= 4096 && !(getenv("SNN_B200_WIDE_STAGE") && atoi(getenv("SNN_B200_WIDE_STAGE")) == 0)) ? 1u : 0u;
            cudaError_t e = cudaMemsetAsync(multi_barrier_, 0, sizeof(unsigned int), stream_);
            if (e == cudaSuccess) e = launch_step_multi(sp, tp, mp, model, chemg, ntrel, stdp, net, wide, device, false, stream_);
            if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_step_multi"); break; }
            n_launch++;
            // the host mirrors what the launch did to the ping-pong state and the clocks
            internal_clock += k;
            if (k & 1ull) { cur_ ^= 1; if (lft_pp) lft_loc_ ^= 1; }
            if (n_trains) { train_draws += k; for (auto &L : lats_) if (L.is_train) L.clock += k; }
            first_step = false;
            s0 += k;
        }
        for (uint64_t s = 0; !multi && s < steps; ++s) {
            if (dbg_timing && steps <= 32) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, stream_); dbg_ev.push_back(ev); }
            const int in = cur_, out = cur_ ^ 1;
            sp.clock = (uint32_t)internal_clock;
            sp.dbg = dbg_clk ? dbg_clk + 4 * s : nullptr;
            sp.reverse = (win_reverse && (internal_clock & 1ull)) ? 1u : 0u;
            sp.apply_pending = (stdp && !first_step) ? 1u : 0u;
            sp.v_in = V_[in]; sp.v_out = V_[out];
            sp.spk_in = SPK_[in]; sp.spk_out = SPK_[out];
            sp.t_in = T_[in]; sp.t_out = T_[out];
            if (lft_pp) { sp.lft_in = LFT_[lft_loc_]; sp.lft_out = LFT_[lft_loc_ ^ 1]; }
            else { sp.lft_in = LFT_[lft_loc_]; sp.lft_out = LFT_[lft_loc_]; }
            sp.grid_hist = want_grid ? d_grid + s * n_neurons : nullptr;
            sp.spike_hist = want_spk ? d_spk + s * n_words : nullptr;
            sp.out_par = (uint32_t)out;
            sp.halo_epoch = halo_nowait ? 0ull : halo_epoch_;   // timing experiment only (SNN_B200_HALO_NOWAIT): results are wrong
            if (!local_peers_.empty()) wait_local_peers_launched(halo_epoch_, rmod_part);
            // spike-train parameters of this timestep (the trains step after the neurons, with their own clocks)
            auto fill_train_step = [&](uint64_t step_in_chunk) {
                tp.v_in = V_[in]; tp.v_out = V_[out]; tp.spk_in = SPK_[in]; tp.spk_out = SPK_[out];
                tp.t_in = T_[in]; tp.t_out = T_[out];
                tp.lft_in = sp.lft_in; tp.lft_out = sp.lft_out;
                tp.n_tl = 0;
                for (auto &L : lats_)
                    if (L.is_train) { tp.tl_base[tp.n_tl] = (uint32_t)L.off; tp.tl_clock[tp.n_tl] = (uint32_t)L.clock; tp.n_tl++; }
                tp.tl_base[tp.n_tl] = (uint32_t)n_trains;
                tp.draw = train_draws++;
                tp.grid_hist = want_tgrid ? d_tgrid + step_in_chunk * n_trains : nullptr;
                tp.spike_hist = want_tspk ? d_tspk + step_in_chunk * t_words : nullptr;
            };
            // two-pass wide kernels: the spike trains ride in extra CTAs of the sum pass (nothing in that pass reads train state; the
            // terms pass, which does, has completed) — one launch and one launch gap less per timestep
            const bool trains_fused = n_neurons && !win_ok && !tma_ok && wide && wide_split && n_trains > 0;
            if (trains_fused) fill_train_step(s);
            if (n_neurons && win_ok) {
                cudaError_t e = launch_step_win(sp, win, model, chemg, ntrel, stdp, win_grid, stream_);
                if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_step_win"); break; }
                n_launch++;
            } else if (n_neurons && tma_ok) {
                tma.st[tma_iv].src = (const unsigned char *)(sp.v_in + own0_);
                if (tma_il >= 0) tma.st[tma_il].src = (const unsigned char *)(sp.lft_in + own0_);
                for (int ty = 0; ty < kNT; ++ty)
                    if (tma_it[ty] >= 0) tma.st[tma_it[ty]].src = (const unsigned char *)(sp.t_in + (size_t)ty * node_cap_ + own0_);
                cudaError_t e = launch_step_tma(sp, tma, model, chemg, ntrel, stdp, tma_grid, stream_);
                if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_step_tma"); break; }
                n_launch++;
            } else if (n_neurons) {
                cudaError_t e = wide ? launch_step_wide(sp, model, chemg, ntrel, stdp, net, wide_split ? wide_scratch_ : nullptr, wide_chunks_cap, trains_fused ? &tp : nullptr, stream_)
                                     : launch_step(sp, model, chemg, ntrel, stdp, net, stream_);
                if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_step"); break; }
                n_launch += (wide && wide_split) ? 2 : 1;   // terms pass + sum pass
            }
            if (reward_mode && rewards) {
                // RewardModulatedSTDP::update, plasticity/mod.rs:193-195 (run_lattice_with_reward, neuron/mod.rs:3160-3172)
                rstdp.dopamine = rstdp.dopamine * expf(-rstdp.dt / rstdp.tau_d) + rstdp.tau_d * rewards[done + s];
            }
            if (net_reward) {
                // RewardModulatedLatticeNetwork::run_lattices_with_reward (neuron/mod.rs:5280-5297): every reward-modulated lattice's
                // modulator takes the reward before iterate; then post_neuron_update_step's pass over their rows
                int kn = 0;
                for (auto &L : lats_) {
                    if (L.is_train) continue;
                    if (L.is_reward && rewards) L.rstdp.dopamine = L.rstdp.dopamine * expf(-L.rstdp.dt / L.rstdp.tau_d) + L.rstdp.tau_d * rewards[done + s];
                    rnp.lat[kn++].dopamine = L.rstdp.dopamine;
                }
                if (net_rmod) {
                    cudaError_t e = launch_rstdp_net_edges(sp, rnp, stream_);
                    if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_rstdp_net_edges"); break; }
                    n_launch++;
                }
            }
            if (bcm_on) {
                cudaError_t e = launch_bcm_edges(sp, BcmParams{bcm.decay, bcm.average_scalar, bcm.dt}, stream_);
                if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_bcm_edges"); break; }
                n_launch++;
            }
            if (rmod) {
                rsp.dopamine = rstdp.dopamine;
                StepParams ep = sp;
                if (rmod_part) {
                    // the kernel needs the neighbours' step-s exports (their last_firing_time of this step in my ghost slots)
                    for (int d = 0; d < 2; ++d) ep.halo[d].my_flag = halo_dir_[d].my_flag;
                    ep.halo_epoch = halo_epoch_ + 1;
                }
                if (rmod_part && !local_peers_.empty()) {
                    launched_pub_.store(halo_epoch_ + 1, std::memory_order_release);   // my step kernel of this timestep is in the queue
                    wait_local_peers_launched(halo_epoch_ + 1, false);
                }
                cudaError_t e = launch_rstdp_edges(ep, rsp, stream_);
                if (rmod_part) launched_pub2_.store(halo_epoch_ + 1, std::memory_order_release);
                if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_rstdp_edges"); break; }
                n_launch++;
            }
            if (part_world > 1) { halo_epoch_ += 1; launched_pub_.store(halo_epoch_, std::memory_order_release); }
            // LatticeNetwork::iterate: clock += 1, then the spike trains step with their own clocks
            // (neuron/mod.rs:2582-2591)
            if (n_trains) {
                if (!trains_fused) {
                    fill_train_step(s);
                    cudaError_t e = launch_trains(tp, stream_);
                    if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "launch_trains"); break; }
                    n_launch++;
                }
                for (auto &L : lats_) if (L.is_train) L.clock += 1;
            }
            internal_clock += 1;
            cur_ = out;
            if (lft_pp) lft_loc_ ^= 1;
            first_step = false;
        }
        if (status != SNN_OK) break;
        done += steps;
        if (done == iterations && stdp) {
            // the last step's STDP is still pending: apply it now so that weights are final at the API boundary
            StepParams fp = sp;
            fp.clock = (uint32_t)internal_clock;
            fp.lft_in = LFT_[lft_loc_];
            fp.lft_out = LFT_[lft_loc_ ^ 1];  // spike trains: last_firing_time from before their last iterate
            fp.halo_epoch = halo_epoch_;
            if (!local_peers_.empty()) wait_local_peers_launched(halo_epoch_, false);
            cudaError_t e = launch_flush_stdp(fp, stream_);
            if (e != cudaSuccess) { bail(e, SNN_GPU_QUEUE_FAILURE, "flush_stdp"); break; }
            n_launch++;
        }
        cudaEventRecord(ev1_, stream_);
        const auto host_t1 = std::chrono::steady_clock::now();
        cudaError_t e = cudaStreamSynchronize(stream_);
        if (e != cudaSuccess) { bail(e, SNN_GPU_WAIT_ERROR, "cudaStreamSynchronize(step loop)"); break; }
        if (dbg_timing && !dbg_ev.empty()) {
            fprintf(stderr, "[snn] per-step device us:");
            for (size_t k = 0; k < dbg_ev.size(); ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, dbg_ev[k], k + 1 < dbg_ev.size() ? dbg_ev[k + 1] : ev1_);
                fprintf(stderr, " %.0f", ms * 1e3f);
                cudaEventDestroy(dbg_ev[k]);
            }
            fprintf(stderr, "\n");
            if (dbg_clk) {
                unsigned long long h[32 * 4];
                cudaMemcpy(h, dbg_clk, sizeof h, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[snn] effective SM MHz of CTA 0 per step:");
                for (uint64_t k = 0; k < steps; ++k)
                    if (h[4 * k + 3] > h[4 * k + 1]) fprintf(stderr, " %.0f", (double)(h[4 * k + 2] - h[4 * k]) * 1e3 / (double)(h[4 * k + 3] - h[4 * k + 1]));
                fprintf(stderr, "\n");
                cudaFree(dbg_clk);
            }
        }
        sp.dbg = nullptr;
        if (dbg_timing) {
            const auto host_t2 = std::chrono::steady_clock::now();
            fprintf(stderr, "[snn] %llu steps: enqueue %.1f us/step on the host, then waited %.1f us/step\n", (unsigned long long)steps,
                    std::chrono::duration<double, std::micro>(host_t1 - host_t0).count() / steps,
                    std::chrono::duration<double, std::micro>(host_t2 - host_t1).count() / steps);
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0_, ev1_);
        total_ms += ms;
        // drain histories
        if (per_step_bytes) {
            std::vector<float> hg, htg; std::vector<uint32_t> hs, hts;
            cudaError_t he = cudaSuccess;
            if (copy_grid) { hg.resize(steps * n_neurons); he = cudaMemcpy(hg.data(), d_grid, hg.size() * 4, cudaMemcpyDeviceToHost); }
            if (he != cudaSuccess) { bail(he, SNN_GPU_BUFFER_READ_ERROR, "grid history drain"); break; }
            if (want_red) {
                const size_t nl = red_lats.size();
                cudaError_t re = launch_history_reduce(d_grid, n_neurons, (uint32_t)steps, d_red_lat, d_red_lat + nl,
                                                       (const float *)(d_red_lat + 2 * nl), (int)nl, d_red, stream_);
                std::vector<double> hr(steps * nl * 2);
                // stream_ is non-blocking: a copy on the legacy stream would not wait for the reduction
                if (re == cudaSuccess) re = cudaMemcpyAsync(hr.data(), d_red, hr.size() * 8, cudaMemcpyDeviceToHost, stream_);
                if (re == cudaSuccess) re = cudaStreamSynchronize(stream_);
                if (re != cudaSuccess) { bail(re, SNN_GPU_BUFFER_READ_ERROR, "history reduce"); break; }
                for (size_t k = 0; k < nl; ++k) {
                    Lat &L = *red_lats[k];
                    for (uint64_t s = 0; s < steps; ++s) {
                        const double sum_v = hr[(s * nl + k) * 2], sum_dv = hr[(s * nl + k) * 2 + 1];
                        // AverageVoltageHistory::update, neuron/mod.rs:310-316: sum / len as f32
                        if (L.avg_hist) L.average_history.push_back((float)sum_v / (float)L.n);
                        // EEGHistory::update, neuron/mod.rs:266-279: (1 / (4 pi sigma d)) * total_current
                        if (L.eeg_hist) L.eeg_history.push_back((1.f / (4.f * 3.14159274101257324f * L.eeg_cond * L.eeg_dist)) * (float)sum_dv);
                    }
                    if (!L.grid_hist && !L.spike_hist) L.hist_len += steps;
                }
            }
            if (want_spk) { hs.resize(steps * n_words); he = cudaMemcpy(hs.data(), d_spk, hs.size() * 4, cudaMemcpyDeviceToHost); }
            if (he == cudaSuccess && want_tgrid) { htg.resize(steps * n_trains); he = cudaMemcpy(htg.data(), d_tgrid, htg.size() * 4, cudaMemcpyDeviceToHost); }
            if (he == cudaSuccess && want_tspk) { hts.resize(steps * t_words); he = cudaMemcpy(hts.data(), d_tspk, hts.size() * 4, cudaMemcpyDeviceToHost); }
            if (he != cudaSuccess) { bail(he, SNN_GPU_BUFFER_READ_ERROR, "history drain"); break; }
            // SpikeHistory::aggregate (neuron/mod.rs:335-359): the sum over the chunk's steps is taken on the device from the
            // staged raster; the host only adds one count per neuron and chunk
            std::vector<uint32_t> cnt_n, cnt_t;
            for (int dom = 0; dom < 2 && he == cudaSuccess; ++dom) {
                const bool on = dom ? want_tspk : want_spk;
                const uint64_t dn = dom ? n_trains : n_neurons, dw = dom ? t_words : n_words;
                if (!on || dn == 0) continue;
                std::vector<uint32_t> &cnt = dom ? cnt_t : cnt_n;
                cnt.resize(dn);
                if (ensure_scratch(dn * 4)) { he = cudaErrorMemoryAllocation; break; }
                he = launch_spike_count(dom ? d_tspk : d_spk, (uint32_t)steps, dw, dn, (uint32_t *)scratch_, stream_);
                if (he == cudaSuccess) he = cudaMemcpyAsync(cnt.data(), scratch_, dn * 4, cudaMemcpyDeviceToHost, stream_);
                if (he == cudaSuccess) he = cudaStreamSynchronize(stream_);
            }
            if (he != cudaSuccess) { bail(he, SNN_GPU_BUFFER_READ_ERROR, "spike aggregate"); break; }
            for (auto &L : lats_) {
                if (!L.grid_hist && !L.spike_hist) continue;
                const uint64_t dom_n = L.is_train ? n_trains : n_neurons, dom_w = L.is_train ? t_words : n_words;
                const std::vector<float> &G = L.is_train ? htg : hg;
                const std::vector<uint32_t> &S = L.is_train ? hts : hs;
                if (L.spike_hist) {
                    const std::vector<uint32_t> &cnt = L.is_train ? cnt_t : cnt_n;
                    if (L.spike_agg.size() != L.n) L.spike_agg.assign(L.n, 0);
                    for (uint64_t j = 0; j < L.n; ++j) L.spike_agg[j] += cnt[L.off + j];
                }
                for (uint64_t s = 0; s < steps; ++s) {
                    if (L.grid_hist) L.grid_history.insert(L.grid_history.end(), G.begin() + s * dom_n + L.off, G.begin() + s * dom_n + L.off + L.n);
                    if (L.spike_hist) {
                        const size_t o = L.spike_history.size();
                        L.spike_history.resize(o + L.n);
                        for (uint64_t j = 0; j < L.n; ++j) {
                            const uint64_t b = L.off + j;
                            L.spike_history[o + j] = (uint8_t)((S[s * dom_w + (b >> 5)] >> (b & 31)) & 1u);
                        }
                    }
                }
                L.hist_len += steps;
            }
        }
    }
    free_hist();
    if (status != SNN_OK) return status;
    if (part_world > 1) {
        if (getenv("SNN_DEBUG_HALO")) {
            unsigned int d[8];
            cudaMemcpy(d, halo_done_, sizeof d, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[snn rank %d] %llu steps: %u halo waits spun, %.1f us waited in total (%.2f us per step)\n", part_rank,
                    (unsigned long long)iterations, d[5], d[4] * 16e-3, d[4] * 16e-3 / (double)iterations);
            cudaMemsetAsync(halo_done_ + 4, 0, 8, stream_);
            cudaStreamSynchronize(stream_);
        }
        unsigned int err = 0;
        CK(cudaMemcpy(&err, halo_done_ + 2, 4, cudaMemcpyDeviceToHost), SNN_GPU_BUFFER_READ_ERROR);
        if (err) {
            cudaMemsetAsync(halo_done_, 0, 4 * sizeof(unsigned int), stream_);
            cudaStreamSynchronize(stream_);
            return fail(SNN_GPU_WAIT_ERROR, "timed out waiting for a neighbouring strip's halo (is every rank running the same number of steps?)");
        }
    }
    if (stdp || rmod || net_rmod || bcm_on) dev_weights_newer_ = true;
    // derived fields (receptor currents, HH gate rates / channel currents) from the retained pre-update V
    if ((chemical && chem_alloc_) || model == SNN_MODEL_HODGKIN_HUXLEY) {
        StepParams fp = sp;
        fp.chemical = chemical && chem_alloc_;
        CK(launch_finalize(fp, model, V_[cur_ ^ 1], stream_), SNN_GPU_QUEUE_FAILURE);
        CK(cudaStreamSynchronize(stream_), SNN_GPU_WAIT_ERROR);
    }
    if (elapsed_ms) *elapsed_ms = total_ms;
    if (launches) *launches = n_launch;
    return SNN_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU halo plumbing (CUDA IPC; the caller moves the blobs between ranks)
// ------------------------------------------------------------------------------------------------
int Engine::ipc_export(IpcBlob *blob) {
    int r = ipc_export_layout(blob);
    if (r) return r;
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    CK(cudaIpcGetMemHandle(&blob->slab, slab_), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(cudaIpcGetMemHandle(&blob->flags, flags_), SNN_GPU_BUFFER_CREATE_ERROR);
    layout_frozen_ = true;   // a neighbour will hold pointers into the slab
    return SNN_OK;
}

int Engine::ipc_export_layout(IpcBlob *blob) {
    if (part_world <= 1) return fail(SNN_INVALID_ARGUMENT, "handle is not partitioned");
    memset(blob, 0, sizeof *blob);
    blob->magic = 0x534E4E42u; blob->version = SNN_B200_ABI_VERSION;
    for (int k = 0; k < 2; ++k) { blob->off_v[k] = slab_off_v_[k]; blob->off_lft[k] = slab_off_lft_[k]; blob->off_t[k] = slab_off_t_[k]; }
    blob->off_flags = slab_off_flags_;
    blob->t_stride = node_cap_;
    blob->own0 = own0_; blob->n_neurons = (uint32_t)n_neurons; blob->ghost_hi0 = ghost_hi0_; blob->halo = halo_;
    blob->cols = lats_.empty() ? 0 : lats_[0].cols; blob->chem = chem_alloc_;
    blob->rank = part_rank; blob->world = part_world;
    return SNN_OK;
}

int Engine::ipc_attach(int direction, const IpcBlob *blob) {
    if (part_world <= 1) return fail(SNN_INVALID_ARGUMENT, "handle is not partitioned");
    if (direction != -1 && direction != 1) return fail(SNN_INVALID_ARGUMENT, "direction must be -1 or +1");
    if (blob->magic != 0x534E4E42u || blob->version != SNN_B200_ABI_VERSION) return fail(SNN_INVALID_ARGUMENT, "bad ipc blob");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    const int d = direction < 0 ? 0 : 1;
    if (peer_slab_[d] || halo_dir_[d].active) return fail(SNN_INVALID_ARGUMENT, "direction already attached");
    void *slab = nullptr, *flags = nullptr;
    CK(cudaIpcOpenMemHandle(&slab, blob->slab, cudaIpcMemLazyEnablePeerAccess), SNN_GPU_BUFFER_CREATE_ERROR);
    CK(cudaIpcOpenMemHandle(&flags, blob->flags, cudaIpcMemLazyEnablePeerAccess), SNN_GPU_BUFFER_CREATE_ERROR);
    peer_slab_[d] = slab; peer_flags_[d] = flags;
    int r = attach_view(direction, blob, slab, flags);
    if (r) {
        cudaIpcCloseMemHandle(slab); cudaIpcCloseMemHandle(flags);
        peer_slab_[d] = peer_flags_[d] = nullptr;
    }
    return r;
}

// The neighbouring strip lives in THIS process (another handle on the same device, or on a device with peer access): no CUDA
// IPC, the neighbour's slab and arrival counters are addressed directly.  Both handles must attach each other.  On one device
// the caller steps the handles from separate host threads and keeps the strips small enough for their step kernels to be
// resident together (a strip's boundary warps spin on the neighbour's progress).
int Engine::attach_local(int direction, Engine *peer) {
    if (part_world <= 1) return fail(SNN_INVALID_ARGUMENT, "handle is not partitioned");
    if (direction != -1 && direction != 1) return fail(SNN_INVALID_ARGUMENT, "direction must be -1 or +1");
    if (!peer || peer == this) return fail(SNN_INVALID_ARGUMENT, "bad neighbour handle");
    CK(cudaSetDevice(device), SNN_GPU_GET_DEVICE_FAILURE);
    const int d = direction < 0 ? 0 : 1;
    if (peer_slab_[d] || halo_dir_[d].active) return fail(SNN_INVALID_ARGUMENT, "direction already attached");
    if (peer->device != device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, device, peer->device), SNN_GPU_GET_DEVICE_FAILURE);
        if (!can) return fail(SNN_UNSUPPORTED, "no peer access between the two devices");
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, SNN_GPU_GET_DEVICE_FAILURE, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    IpcBlob blob;
    int r = peer->ipc_export_layout(&blob);
    if (r) return fail(r, peer->last_error);
    r = attach_view(direction, &blob, peer->slab_, peer->flags_);
    if (!r) {
        peer->layout_frozen_ = layout_frozen_ = true;   // raw pointers into each other's slabs from now on
        local_peers_.push_back(peer);
    }
    return r;
}

int Engine::attach_view(int direction, const IpcBlob *blob, void *slab, void *flags) {
    if (blob->rank != part_rank + direction || blob->world != part_world) return fail(SNN_INVALID_ARGUMENT, "ipc blob is not from the neighbouring rank");
    if (blob->halo != halo_ || (blob->chem != 0) != chem_alloc_) return fail(SNN_INVALID_ARGUMENT, "neighbouring strips disagree on halo width / chemistry");
    if (halo_ > n_neurons || halo_ > blob->n_neurons) return fail(SNN_UNSUPPORTED, "strip thinner than the halo");
    const int d = direction < 0 ? 0 : 1;
    HaloDir &H = halo_dir_[d];
    memset(&H, 0, sizeof H);
    H.count = halo_;
    // d == 0: my first `halo` neurons go to the upper ghost slots of rank-1; d == 1: my last `halo` neurons go to
    // the lower ghost slots of rank+1
    H.first = d == 0 ? 0 : (uint32_t)n_neurons - halo_;
    H.peer_node0 = d == 0 ? blob->ghost_hi0 : blob->own0 - blob->halo;
    H.my_ghost0 = d == 0 ? own0_ - halo_ : ghost_hi0_;
    for (int k = 0; k < 2; ++k) {
        H.peer_v[k] = (float *)((char *)slab + blob->off_v[k]);
        H.peer_lft[k] = (int *)((char *)slab + blob->off_lft[k]);
        H.peer_t[k] = (float *)((char *)slab + blob->off_t[k]);
    }
    H.peer_t_stride = blob->t_stride;
    // the neighbour's arrival counter for data coming from me: I am its rank+1 when d == 0 (slot 1), its rank-1 when d == 1 (slot 0)
    H.peer_flag = (unsigned long long *)flags + (d == 0 ? 1 : 0);
    H.my_flag = flags_ + d;
    H.peer_flag2 = H.peer_flag + 2;
    H.my_flag2 = H.my_flag + 2;
    H.active = 1;
    // my ghost rows take the neurotransmitter / receptor type flags of the neighbour's boundary rows (they are baked into the
    // col words of edges from ghosts): d == 0 reads the LAST halo_ neurons of rank - 1, d == 1 the FIRST halo_ of rank + 1
    {
        const uint8_t *peer_f = (const uint8_t *)slab + blob->off_flags + (d == 0 ? blob->own0 + blob->n_neurons - halo_ : blob->own0);
        std::vector<uint8_t> tmp(halo_);
        CK(cudaMemcpy(tmp.data(), peer_f, halo_, cudaMemcpyDefault), SNN_GPU_BUFFER_READ_ERROR);
        bool changed = !ghost_flags_from_peer_[d];
        for (uint32_t g = 0; g < halo_; ++g) { changed |= h_node_flags_[H.my_ghost0 + g] != tmp[g]; h_node_flags_[H.my_ghost0 + g] = tmp[g]; }
        ghost_flags_from_peer_[d] = true;
        if (changed) {
            CK(h2d_sync(node_flags_ + H.my_ghost0, tmp.data(), halo_, stream_), SNN_GPU_BUFFER_WRITE_ERROR);
            flags_cache_valid_ = false;
            if (dev_weights_newer_) { int r = sync_weights_to_host(); if (r) return r; }
            graph_dirty_ = true;
        }
    }
    return SNN_OK;
}

}  // namespace snn
