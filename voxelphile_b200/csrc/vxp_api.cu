// vxp_api.cu — the C ABI of libvxp.so (include/vxp.h).  Thin: argument checks, context-owned
// arenas for the host-level calls, launches.  No CPU fallback anywhere: every compute entry
// point ends in a kernel launch or an error.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/vxp.h"
#include "vxp_launch.h"

// A few host threads that copy a caller's PAGEABLE buffer into pinned staging memory in parallel (one memcpy thread moves
// ~10 GB/s, a PCIe 5 x16 link 55): created on the first host-level call that is handed pageable memory.
struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    const unsigned char* src = nullptr;
    unsigned char* dst = nullptr;
    size_t bytes = 0;
    uint64_t generation = 0;
    unsigned pending = 0;
    bool stop = false;

    explicit CopyPool(unsigned threads) {
        for (unsigned t = 0; t < threads; ++t) workers.emplace_back([this, t, threads] { run(t, threads); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
        }
        cv_job.notify_all();
        for (auto& w : workers) w.join();
    }
    void run(unsigned t, unsigned threads) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_job.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            const size_t piece = ((bytes + threads - 1) / threads + 4095) & ~(size_t)4095;
            const size_t lo = (size_t)t * piece, hi = lo + piece < bytes ? lo + piece : bytes;
            const unsigned char* s_ = src;
            unsigned char* d_ = dst;
            lk.unlock();
            if (lo < hi) std::memcpy(d_ + lo, s_ + lo, hi - lo);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    void copy(void* d, const void* s_, size_t n) {   // returns when the bytes are in place
        std::unique_lock<std::mutex> lk(mu);
        src = static_cast<const unsigned char*>(s_);
        dst = static_cast<unsigned char*>(d);
        bytes = n;
        pending = (unsigned)workers.size();
        ++generation;
        cv_job.notify_all();
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

struct vxp_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    std::string last_error;

    // arenas of the host-level calls (grown on demand, never shrunk)
    uint16_t* d_voxels = nullptr;   size_t voxels_cap = 0;   // chunks
    int32_t* d_nbr6 = nullptr;      size_t nbr_cap = 0;      // chunks
    vxp_coord* d_coords = nullptr;  size_t coords_cap = 0;   // chunks
    vxp_chunk_meta* d_meta = nullptr; size_t meta_cap = 0;   // chunks
    uint64_t* d_verts = nullptr;    uint32_t* d_indices = nullptr; size_t quads_cap = 0;
    uint64_t* d_total = nullptr;
    // boundary-summary scratch of the mesh kernel (vxp_launch.h: BoundaryScratch)
    // Two launch slots used alternately, so that consecutive launches share nothing (launch chaining, vxp_mesh_common.cuh).
    uint32_t* d_tags[2] = {};       unsigned long long* d_planes[2] = {};    size_t bnd_cap = 0;   // chunks
    uint32_t epoch = 0;
    unsigned long long* d_counters = nullptr;   // one zero-initialised launch counter per slot, [2]: the pack kernel's byte cursor
    int slot = 0;
    int overlap = 0;                            // vxp_ctx_set_overlap
    int host_indices = 1;                       // vxp_ctx_set_host_indices
    bool launched = false;                      // a mesh launch has been enqueued ...
    cudaStream_t launch_stream = nullptr;       // ... on this stream
    bool chain_open = false;                    // ... with overlap enabled, and nothing else of ours since
    // incremental re-mesh: per-chunk "on the dirty list" stamps, and the host-level call's resident world / arenas
    uint32_t* d_dirty_flags = nullptr;  size_t dirty_cap = 0;  uint32_t dirty_epoch = 0;
    uint32_t world_n = 0, world_stored = 0;  bool world_has_nbr = false, world_valid = false;
    vxp_edit* d_edits = nullptr;        size_t edits_cap = 0;
    uint32_t* d_dirty = nullptr;        size_t dirty_list_cap = 0;
    uint32_t* d_dirty_count = nullptr;
    uint32_t* h_dirty = nullptr;        size_t h_dirty_cap = 0;
    uint32_t* h_dirty_count = nullptr;
    // wire format: unpack status word (device + pinned copy) and the host-level call's upload arenas
    uint32_t* d_status = nullptr;       uint32_t* h_status = nullptr;   bool status_pending = false;
    unsigned char* d_packed = nullptr;  size_t packed_cap = 0;          // bytes
    unsigned long long* d_offsets = nullptr;  size_t offsets_cap = 0;   // chunks
    cudaEvent_t ev_chain = nullptr;
    // column scratch of the fused terrain+mesh path (vxp_launch.h: ColumnScratch)
    unsigned char* d_columns = nullptr;  size_t columns_cap = 0;   // chunks
    vxp::ColumnScratch columns{};
    vxp_chunk_meta* h_meta = nullptr; size_t h_meta_cap = 0;
    uint64_t* h_verts = nullptr;    uint32_t* h_indices = nullptr; size_t h_quads_cap = 0;
    uint64_t* h_total = nullptr;
    // host-level pipeline (vxp_mesh_chunks_host): copy streams, per-slice events and running totals
    static constexpr int kMaxSlices = 16;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_start = nullptr, ev_h2d[kMaxSlices] = {}, ev_k[kMaxSlices] = {};
    uint64_t* h_totals = nullptr;   // pinned, kMaxSlices entries
    // uploads from pageable caller memory: two pinned staging buffers used alternately, filled by the copy pool
    static constexpr size_t kStageBytes = (size_t)32 << 20;
    unsigned char* h_stage[2] = {};
    cudaEvent_t ev_stage[2] = {};
    int stage_next = 0;
    CopyPool* pool = nullptr;
};

namespace {

vxp_status cuda_fail(vxp_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    if (e == cudaErrorMemoryAllocation) return VXP_ERR_OOM;
    return VXP_ERR_CUDA;
}
#define VXP_CUDA(ctx, expr)                                        \
    do {                                                           \
        cudaError_t e__ = (expr);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #expr); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Host -> device copy of a caller buffer on `stream`.  Pinned (or otherwise CUDA-registered) memory goes straight to
// cudaMemcpyAsync.  Pageable memory (a plain malloc / Vec<u16>) would make the driver stage the copy through its own
// small bounce buffer at ~11 GB/s, a fifth of the link: instead the context's copy pool moves it, 32 MiB at a time, into
// two pinned staging buffers used alternately, each followed by its own asynchronous copy - the memcpy of piece k+1
// overlaps the transfer of piece k.  The source may be re-used by the caller when the call returns.
vxp_status upload(vxp_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return VXP_OK;
    cudaPointerAttributes attr{};
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, h_src) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    if (!pageable || bytes < ((size_t)1 << 20)) {
        VXP_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
        return VXP_OK;
    }
    if (!ctx->pool) {
        unsigned hw = std::thread::hardware_concurrency();
        unsigned threads = hw >= 16 ? 8u : hw >= 4 ? hw / 2 : 1u;
        ctx->pool = new (std::nothrow) CopyPool(threads);
        if (!ctx->pool) return VXP_ERR_OOM;
        for (int k = 0; k < 2; ++k) {
            VXP_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_stage[k]), vxp_ctx::kStageBytes));
            VXP_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stage[k], cudaEventDisableTiming));
        }
    }
    const unsigned char* src = static_cast<const unsigned char*>(h_src);
    unsigned char* dst = static_cast<unsigned char*>(d_dst);
    for (size_t off = 0; off < bytes; off += vxp_ctx::kStageBytes) {
        const size_t n = bytes - off < vxp_ctx::kStageBytes ? bytes - off : vxp_ctx::kStageBytes;
        const int b = ctx->stage_next;
        ctx->stage_next ^= 1;
        VXP_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage[b]));       // the transfer that last used this buffer is done
        ctx->pool->copy(ctx->h_stage[b], src + off, n);
        VXP_CUDA(ctx, cudaMemcpyAsync(dst + off, ctx->h_stage[b], n, cudaMemcpyHostToDevice, stream));
        VXP_CUDA(ctx, cudaEventRecord(ctx->ev_stage[b], stream));
    }
    return VXP_OK;
}

template <class T>
vxp_status grow_device(vxp_ctx* ctx, T** p, size_t* cap, size_t need, size_t elems_per_unit) {
    if (need <= *cap) return VXP_OK;
    size_t want = need + need / 4;
    if (*p) VXP_CUDA(ctx, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(p), want * elems_per_unit * sizeof(T)));
    *cap = want;
    return VXP_OK;
}
template <class T>
vxp_status grow_pinned(vxp_ctx* ctx, T** p, size_t* cap, size_t need, size_t elems_per_unit) {
    if (need <= *cap) return VXP_OK;
    size_t want = need + need / 4;
    if (*p) VXP_CUDA(ctx, cudaFreeHost(*p));
    *p = nullptr;
    *cap = 0;
    VXP_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(p), want * elems_per_unit * sizeof(T)));
    *cap = want;
    return VXP_OK;
}

vxp_status grow_quads(vxp_ctx* ctx, size_t need) {
    if (need <= ctx->quads_cap) return VXP_OK;
    size_t want = need + need / 4;
    if (ctx->d_verts) VXP_CUDA(ctx, cudaFree(ctx->d_verts));
    if (ctx->d_indices) VXP_CUDA(ctx, cudaFree(ctx->d_indices));
    ctx->d_verts = nullptr;
    ctx->d_indices = nullptr;
    ctx->quads_cap = 0;
    VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_verts), want * 4 * sizeof(uint64_t)));
    VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_indices), want * 6 * sizeof(uint32_t)));
    ctx->quads_cap = want;
    return VXP_OK;
}
vxp_status grow_host_quads(vxp_ctx* ctx, size_t need) {
    if (need <= ctx->h_quads_cap) return VXP_OK;
    size_t want = need + need / 4;
    if (ctx->h_verts) VXP_CUDA(ctx, cudaFreeHost(ctx->h_verts));
    if (ctx->h_indices) VXP_CUDA(ctx, cudaFreeHost(ctx->h_indices));
    ctx->h_verts = nullptr;
    ctx->h_indices = nullptr;
    ctx->h_quads_cap = 0;
    VXP_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_verts), want * 4 * sizeof(uint64_t)));
    VXP_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&ctx->h_indices), want * 6 * sizeof(uint32_t)));
    ctx->h_quads_cap = want;
    return VXP_OK;
}

// Scratch for one mesh launch over n chunks, in the slot the launch will use: grown on demand (cudaFree
// synchronises with any launch still using the old area), zeroed on allocation and when the 30-bit epoch wraps.
vxp_status next_boundary_scratch(vxp_ctx* ctx, uint32_t n, vxp::BoundaryScratch* out) {
    if (n > ctx->bnd_cap) {
        const size_t want = (size_t)n + n / 4;
        for (int k = 0; k < 2; ++k) {
            if (ctx->d_tags[k]) VXP_CUDA(ctx, cudaFree(ctx->d_tags[k]));
            if (ctx->d_planes[k]) VXP_CUDA(ctx, cudaFree(ctx->d_planes[k]));
            ctx->d_tags[k] = nullptr;
            ctx->d_planes[k] = nullptr;
        }
        ctx->bnd_cap = 0;
        for (int k = 0; k < 2; ++k) {
            VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_tags[k]), vxp::boundary_tag_bytes((uint32_t)want)));
            VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_planes[k]), vxp::boundary_plane_bytes((uint32_t)want)));
            VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_tags[k], 0, vxp::boundary_tag_bytes((uint32_t)want), ctx->stream));
            VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_planes[k], 0, vxp::boundary_plane_bytes((uint32_t)want), ctx->stream));
        }
        ctx->bnd_cap = want;
        ctx->epoch = 0;
    }
    if (++ctx->epoch >= (1u << 30)) {
        for (int k = 0; k < 2; ++k) {
            VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_tags[k], 0, vxp::boundary_tag_bytes((uint32_t)ctx->bnd_cap), ctx->stream));
            VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_planes[k], 0, vxp::boundary_plane_bytes((uint32_t)ctx->bnd_cap), ctx->stream));
        }
        ctx->epoch = 1;
    }
    *out = vxp::BoundaryScratch{ctx->d_tags[ctx->slot], ctx->d_planes[ctx->slot], ctx->epoch};
    return VXP_OK;
}

// One whole-batch mesh launch in the context's next slot.  Launches of one context are chained on one stream
// (the chain's guarantees hold per stream): if the caller has switched streams since the last launch, the new
// stream first waits for the old one.
vxp_status chained_mesh_launch(vxp_ctx* ctx, const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, int greedy,
                               const vxp_mesh_dev& dev) {
    vxp::BoundaryScratch scratch{nullptr, nullptr, 0};
    if (n) {
        const vxp_status s = next_boundary_scratch(ctx, n, &scratch);
        if (s != VXP_OK) return s;
        if (ctx->launched && ctx->launch_stream != ctx->stream) {
            VXP_CUDA(ctx, cudaEventRecord(ctx->ev_chain, ctx->launch_stream));
            VXP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chain, 0));
        }
    }
    // the launch may start early only behind another overlapping mesh launch of this context on this stream: after
    // anything else of ours (fills, fused and host-level calls) it serialises as usual
    const int early = ctx->overlap && ctx->chain_open && ctx->launch_stream == ctx->stream;
    VXP_CUDA(ctx, vxp::launch_mesh(d_voxels, d_nbr6, n, greedy, dev, scratch, ctx->d_counters + ctx->slot, early, ctx->stream));
    if (n) {
        ctx->launches += 1;
        ctx->slot ^= 1;
        ctx->launched = true;
        ctx->launch_stream = ctx->stream;
        ctx->chain_open = ctx->overlap != 0;
    }
    return VXP_OK;
}

// Column scratch for a fused launch over n chunks: one allocation carved into the ColumnScratch arrays.  The hash
// table and the launch counters are zeroed when allocated; every call gets the next key epoch (the table is
// re-zeroed when the 10-bit epoch wraps), so calls start without a memset.
vxp_status column_scratch(vxp_ctx* ctx, uint32_t n, vxp::ColumnScratch* out) {
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    if (n > ctx->columns_cap) {
        const size_t want = (size_t)n + n / 4;
        const size_t slots = vxp::column_table_slots((uint32_t)want);
        const size_t b_keys = up(slots * 8), b_ids = up(slots * 4), b_slot = up(want * 4), b_work = up(want * 16), b_coord = up(want * 8),
                     b_h = up(want * vxp::column_record_bytes()), b_cnt = 256;
        if (ctx->d_columns) VXP_CUDA(ctx, cudaFree(ctx->d_columns));
        ctx->d_columns = nullptr;
        ctx->columns_cap = 0;
        VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_columns), b_keys + b_ids + b_slot + b_work + b_coord + b_h + b_cnt));
        unsigned char* p = ctx->d_columns;
        ctx->columns.keys = reinterpret_cast<unsigned long long*>(p); p += b_keys;
        ctx->columns.id_of_slot = reinterpret_cast<uint32_t*>(p); p += b_ids;
        ctx->columns.slot_of_chunk = reinterpret_cast<uint32_t*>(p); p += b_slot;
        ctx->columns.work = reinterpret_cast<uint4*>(p); p += b_work;
        ctx->columns.coord_of_id = reinterpret_cast<int2*>(p); p += b_coord;
        ctx->columns.heights = reinterpret_cast<short*>(p); p += b_h;
        ctx->columns.count = reinterpret_cast<uint32_t*>(p);
        ctx->columns.mask = (uint32_t)(slots - 1);
        ctx->columns.epoch = 0;
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->columns.keys, 0, b_keys, ctx->stream));
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->columns.id_of_slot, 0, b_ids, ctx->stream));
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->columns.count, 0, b_cnt, ctx->stream));
        ctx->columns_cap = want;
    }
    if (++ctx->columns.epoch >= 1024u) {
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->columns.keys, 0, ((size_t)ctx->columns.mask + 1) * 8, ctx->stream));
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->columns.id_of_slot, 0, ((size_t)ctx->columns.mask + 1) * 4, ctx->stream));   // epoch-stamped ready words (shared fill)
        ctx->columns.epoch = 1;
    }
    *out = ctx->columns;
    return VXP_OK;
}

// terrain fill: batches of at least this many chunks go through the shared column heights.  Measured on the 16^3 block
// (4096 chunks, 16 per column): 60.1 us direct, 65.6 us through the columns - the fill kernel itself drops to ~50 us but
// two memsets and two small kernels in front of it cost ~15 us, which only a larger batch amortises.
#ifndef VXP_COLUMN_FILL_MIN
#define VXP_COLUMN_FILL_MIN 16384
#endif
constexpr uint32_t kColumnFillMin = VXP_COLUMN_FILL_MIN;
// ... and from this many chunks up to there, through the one-launch variant of the same idea (terrain_fill_shared_kernel)
#ifndef VXP_SHARED_FILL_MIN
#define VXP_SHARED_FILL_MIN 256
#endif
constexpr uint32_t kSharedFillMin = VXP_SHARED_FILL_MIN;

bool valid_mode(vxp_mode m) { return m == VXP_MODE_CULLED || m == VXP_MODE_GREEDY; }

bool valid_out(const vxp_mesh_dev* o, uint32_t n) {
    if (!o || !o->total_quads) return false;
    if (n && !o->meta) return false;
    if (o->capacity_quads && !o->verts) return false;    // indices may be NULL: not written
    if ((reinterpret_cast<uintptr_t>(o->verts) & 31u) || (reinterpret_cast<uintptr_t>(o->indices) & 7u)) return false;
    if (o->capacity_quads > 0xFFFFFFFEull) return false;  // first_quad is a u32; 0xFFFFFFFF is the overflow mark
    return true;
}

// Run one mesh launch into the context's arena, growing it until the result fits; download.
template <class Launch>
vxp_status mesh_into_host(vxp_ctx* ctx, uint32_t n, Launch launch, vxp_mesh_host* out) {
    vxp_status s;
    if ((s = grow_device(ctx, &ctx->d_meta, &ctx->meta_cap, n, 1)) != VXP_OK) return s;
    if ((s = grow_pinned(ctx, &ctx->h_meta, &ctx->h_meta_cap, n, 1)) != VXP_OK) return s;
    if ((s = grow_quads(ctx, (size_t)n * 64 + 4096)) != VXP_OK) return s;  // first guess; grows below
    uint64_t total = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        vxp_mesh_dev dev{ctx->d_verts, ctx->host_indices ? ctx->d_indices : nullptr, ctx->d_meta, ctx->d_total, ctx->quads_cap};
        VXP_CUDA(ctx, launch(dev));   // counts itself in ctx->launches
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_total, ctx->d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        total = *ctx->h_total;
        if (total <= ctx->quads_cap) break;
        if (attempt == 2) return VXP_ERR_CAPACITY;
        if ((s = grow_quads(ctx, total)) != VXP_OK) return s;
    }
    if ((s = grow_host_quads(ctx, total)) != VXP_OK) return s;
    if (n)
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_meta, ctx->d_meta, (size_t)n * sizeof(vxp_chunk_meta),
                                      cudaMemcpyDeviceToHost, ctx->stream));
    if (total) {
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_verts, ctx->d_verts, total * 4 * sizeof(uint64_t),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->host_indices)
            VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_indices, ctx->d_indices, total * 6 * sizeof(uint32_t),
                                          cudaMemcpyDeviceToHost, ctx->stream));
    }
    VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->chain_open = false;
    out->verts = ctx->h_verts;
    out->indices = ctx->host_indices ? ctx->h_indices : nullptr;
    out->meta = ctx->h_meta;
    out->total_quads = total;
    return VXP_OK;
}

}  // namespace

extern "C" {

int vxp_spec_version(void) { return VXP_SPEC_VERSION; }

const char* vxp_status_str(vxp_status s) {
    switch (s) {
        case VXP_OK: return "ok";
        case VXP_ERR_BAD_ARG: return "bad argument";
        case VXP_ERR_NO_DEVICE: return "no CUDA device (libvxp has no CPU fallback)";
        case VXP_ERR_OOM: return "out of memory";
        case VXP_ERR_CAPACITY: return "output capacity exceeded";
        case VXP_ERR_CUDA: return "CUDA error";
    }
    return "unknown status";
}

vxp_status vxp_ctx_create(int device, vxp_ctx** out) {
    if (!out) return VXP_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return VXP_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return VXP_ERR_BAD_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VXP_ERR_CUDA;
    if (prop.major != 10) return VXP_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    vxp_ctx* ctx = new (std::nothrow) vxp_ctx();
    if (!ctx) return VXP_ERR_OOM;
    ctx->device = device;
    DeviceGuard g(device);
    cudaError_t e = g.ok ? cudaSuccess : cudaErrorInvalidDevice;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = vxp::configure_device();
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_total), sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_counters), 3 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_counters, 0, 3 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_dirty_count), sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_dirty_count), sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_status), sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_status), sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_total), sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_totals), vxp_ctx::kMaxSlices * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming);
    for (int k = 0; k < vxp_ctx::kMaxSlices && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_k[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        vxp_status s = cuda_fail(nullptr, e, "ctx_create");
        vxp_ctx_destroy(ctx);
        return s;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return VXP_OK;
}

void vxp_ctx_destroy(vxp_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    cudaFree(ctx->d_voxels);
    cudaFree(ctx->d_nbr6);
    cudaFree(ctx->d_coords);
    cudaFree(ctx->d_meta);
    cudaFree(ctx->d_verts);
    cudaFree(ctx->d_indices);
    cudaFree(ctx->d_total);
    cudaFree(ctx->d_columns);
    for (int k = 0; k < 2; ++k) {
        cudaFree(ctx->d_tags[k]);
        cudaFree(ctx->d_planes[k]);
    }
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_dirty_flags);
    cudaFree(ctx->d_edits);
    cudaFree(ctx->d_dirty);
    cudaFree(ctx->d_dirty_count);
    cudaFreeHost(ctx->h_dirty);
    cudaFreeHost(ctx->h_dirty_count);
    cudaFree(ctx->d_status);
    cudaFreeHost(ctx->h_status);
    cudaFree(ctx->d_packed);
    cudaFree(ctx->d_offsets);
    if (ctx->ev_chain) cudaEventDestroy(ctx->ev_chain);
    cudaFreeHost(ctx->h_meta);
    cudaFreeHost(ctx->h_verts);
    cudaFreeHost(ctx->h_indices);
    cudaFreeHost(ctx->h_total);
    cudaFreeHost(ctx->h_totals);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    for (int k = 0; k < vxp_ctx::kMaxSlices; ++k) {
        if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
        if (ctx->ev_k[k]) cudaEventDestroy(ctx->ev_k[k]);
    }
    delete ctx->pool;
    for (int k = 0; k < 2; ++k) {
        cudaFreeHost(ctx->h_stage[k]);
        if (ctx->ev_stage[k]) cudaEventDestroy(ctx->ev_stage[k]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

vxp_status vxp_ctx_set_stream(vxp_ctx* ctx, void* cuda_stream) {
    if (!ctx) return VXP_ERR_BAD_ARG;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return VXP_OK;
}

vxp_status vxp_ctx_reset_stream(vxp_ctx* ctx) {
    if (!ctx) return VXP_ERR_BAD_ARG;
    ctx->stream = ctx->own_stream;
    return VXP_OK;
}

vxp_status vxp_ctx_set_overlap(vxp_ctx* ctx, int enabled) {
    if (!ctx) return VXP_ERR_BAD_ARG;
    ctx->overlap = enabled ? 1 : 0;
    ctx->chain_open = false;
    return VXP_OK;
}

vxp_status vxp_ctx_set_host_indices(vxp_ctx* ctx, int enabled) {
    if (!ctx) return VXP_ERR_BAD_ARG;
    ctx->host_indices = enabled ? 1 : 0;
    return VXP_OK;
}

// The stream is idle: did an unpack since the last look meet a malformed record?  The device word is sticky (set by the
// kernels, never cleared by a launch), so any number of unpack calls may be queued without a synchronisation between them.
static vxp_status unpack_verdict(vxp_ctx* ctx) {
    if (!ctx->status_pending) return VXP_OK;
    ctx->status_pending = false;
    VXP_CUDA(ctx, cudaMemcpy(ctx->h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (*ctx->h_status == 0u) return VXP_OK;
    VXP_CUDA(ctx, cudaMemset(ctx->d_status, 0, sizeof(uint32_t)));
    ctx->last_error = "vxp_unpack_chunks: malformed record";
    return VXP_ERR_BAD_ARG;
}

vxp_status vxp_ctx_synchronize(vxp_ctx* ctx) {
    if (!ctx) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->chain_open = false;
    return unpack_verdict(ctx);
}

const char* vxp_last_error(const vxp_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t vxp_launch_count(const vxp_ctx* ctx) { return ctx ? ctx->launches : 0; }

vxp_status vxp_nbr6_from_coords(const vxp_coord* coords, uint32_t n, int32_t* nbr6) {
    if (n && (!coords || !nbr6)) return VXP_ERR_BAD_ARG;
    auto key = [](int32_t x, int32_t y, int32_t z) {
        return (uint64_t)((uint32_t)x & 0x1FFFFFu) | (uint64_t)((uint32_t)y & 0x1FFFFFu) << 21 |
               (uint64_t)((uint32_t)z & 0x1FFFFFu) << 42;
    };
    const int32_t lim = 1 << 20;
    std::unordered_map<uint64_t, int32_t> index;
    try {
        index.reserve((size_t)n * 2);
        for (uint32_t i = 0; i < n; ++i) {
            const vxp_coord c = coords[i];
            if (c.x < -lim + 1 || c.x >= lim - 1 || c.y < -lim + 1 || c.y >= lim - 1 || c.z < -lim + 1 ||
                c.z >= lim - 1)
                return VXP_ERR_BAD_ARG;
            if (!index.emplace(key(c.x, c.y, c.z), (int32_t)i).second) return VXP_ERR_BAD_ARG;  // duplicate
        }
    } catch (const std::bad_alloc&) {
        return VXP_ERR_OOM;
    }
    static const int d[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};
    for (uint32_t i = 0; i < n; ++i)
        for (int f = 0; f < 6; ++f) {
            auto it = index.find(key(coords[i].x + d[f][0], coords[i].y + d[f][1], coords[i].z + d[f][2]));
            nbr6[6 * (size_t)i + f] = it == index.end() ? -1 : it->second;
        }
    return VXP_OK;
}

vxp_status vxp_host_alloc(size_t bytes, void** out) {
    if (!out) return VXP_ERR_BAD_ARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? VXP_ERR_OOM : VXP_ERR_NO_DEVICE;
    }
    return VXP_OK;
}
void vxp_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ device level

vxp_status vxp_terrain_fill_batch(vxp_ctx* ctx, const vxp_coord* d_coords, uint32_t n, uint64_t seed,
                                  uint16_t* d_voxels) {
    if (!ctx || (n && (!d_coords || !d_voxels)) || !aligned16(d_voxels)) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    ctx->chain_open = false;
    if (n >= kColumnFillMin) {   // large batch: the chunks of a column share its heights
        vxp::ColumnScratch cols{};
        const vxp_status s = column_scratch(ctx, n, &cols);
        if (s != VXP_OK) return s;
        VXP_CUDA(ctx, vxp::launch_terrain_fill_columns(d_coords, n, seed, d_voxels, cols, ctx->stream));
        ctx->launches += 3;
        return VXP_OK;
    }
    if (n >= kSharedFillMin) {   // one launch; the first chunk of a column computes its heights for the others
        vxp::ColumnScratch cols{};
        const vxp_status s = column_scratch(ctx, n, &cols);
        if (s != VXP_OK) return s;
        VXP_CUDA(ctx, vxp::launch_terrain_fill_shared(d_coords, n, seed, d_voxels, cols, ctx->stream));
        ctx->launches += 1;
        return VXP_OK;
    }
    VXP_CUDA(ctx, vxp::launch_terrain_fill(d_coords, n, seed, d_voxels, ctx->stream));
    ctx->launches += n ? 1 : 0;
    return VXP_OK;
}

vxp_status vxp_pattern_fill_batch(vxp_ctx* ctx, const vxp_coord* d_coords, uint32_t n, vxp_pattern pattern,
                                  uint16_t* d_voxels) {
    if (!ctx || (n && (!d_coords || !d_voxels)) || !aligned16(d_voxels)) return VXP_ERR_BAD_ARG;
    if (pattern != VXP_PATTERN_CHECKER && pattern != VXP_PATTERN_SOLID4) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    ctx->chain_open = false;
    VXP_CUDA(ctx, vxp::launch_pattern_fill(d_coords, n, (int)pattern, d_voxels, ctx->stream));
    ctx->launches += n ? 1 : 0;
    return VXP_OK;
}

vxp_status vxp_mesh_batch(vxp_ctx* ctx, const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n,
                          vxp_mode mode, const vxp_mesh_dev* out) {
    if (!ctx || !valid_mode(mode) || !valid_out(out, n)) return VXP_ERR_BAD_ARG;
    if (n && (!d_voxels || !aligned16(d_voxels))) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    return chained_mesh_launch(ctx, d_voxels, d_nbr6, n, mode == VXP_MODE_GREEDY, *out);
}

vxp_status vxp_terrain_mesh_fused(vxp_ctx* ctx, const vxp_coord* d_coords, uint32_t n, uint64_t seed,
                                  vxp_mode mode, const vxp_mesh_dev* out) {
    if (!ctx || !valid_mode(mode) || !valid_out(out, n) || (n && !d_coords)) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp::ColumnScratch cols{};
    if (n) {
        const vxp_status s = column_scratch(ctx, n, &cols);
        if (s != VXP_OK) return s;
    }
    ctx->chain_open = false;
    VXP_CUDA(ctx, vxp::launch_terrain_mesh(d_coords, n, seed, mode == VXP_MODE_GREEDY, *out, cols, ctx->stream));
    ctx->launches += n ? 4 : 0;          // columns, heights, classification, mesh
    return VXP_OK;
}

// ------------------------------------------------------------------ incremental re-mesh

namespace {
// "on the dirty list" stamps for n chunks: zeroed on allocation and when the epoch wraps
vxp_status next_dirty_flags(vxp_ctx* ctx, uint32_t n, uint32_t* epoch) {
    if (n > ctx->dirty_cap) {
        const size_t want = (size_t)n + n / 4;
        if (ctx->d_dirty_flags) VXP_CUDA(ctx, cudaFree(ctx->d_dirty_flags));
        ctx->d_dirty_flags = nullptr;
        ctx->dirty_cap = 0;
        VXP_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_dirty_flags), want * sizeof(uint32_t)));
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_dirty_flags, 0, want * sizeof(uint32_t), ctx->stream));
        ctx->dirty_cap = want;
        ctx->dirty_epoch = 0;
    }
    if (++ctx->dirty_epoch == 0xFFFFFFFFu) {
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_dirty_flags, 0, ctx->dirty_cap * sizeof(uint32_t), ctx->stream));
        ctx->dirty_epoch = 1;
    }
    *epoch = ctx->dirty_epoch;
    return VXP_OK;
}
uint32_t max_dirty(uint32_t n, uint32_t m) { return (uint64_t)m * 7 < n ? m * 7 : n; }

// serialise against whatever mesh launch of this context may still be running on another stream
vxp_status join_chain(vxp_ctx* ctx) {
    if (ctx->launched && ctx->launch_stream != ctx->stream) {
        VXP_CUDA(ctx, cudaEventRecord(ctx->ev_chain, ctx->launch_stream));
        VXP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chain, 0));
    }
    ctx->launched = true;
    ctx->launch_stream = ctx->stream;
    ctx->chain_open = false;
    return VXP_OK;
}
}  // namespace

vxp_status vxp_mesh_list(vxp_ctx* ctx, const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n,
                         const uint32_t* d_list, uint32_t count, vxp_mode mode, const vxp_mesh_dev* out) {
    if (!ctx || !valid_mode(mode) || !valid_out(out, count)) return VXP_ERR_BAD_ARG;
    if (count && (!d_voxels || !aligned16(d_voxels) || !d_list || n == 0)) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp_status s;
    if ((s = join_chain(ctx)) != VXP_OK) return s;
    VXP_CUDA(ctx, vxp::launch_mesh_list(d_voxels, d_nbr6, n, d_list, nullptr, count, mode == VXP_MODE_GREEDY, *out,
                                        ctx->d_counters + ctx->slot, ctx->stream));
    if (count) {
        ctx->launches += 1;
        ctx->slot ^= 1;
    }
    return VXP_OK;
}

vxp_status vxp_remesh_edits(vxp_ctx* ctx, uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, const vxp_edit* d_edits,
                            uint32_t m, vxp_mode mode, uint32_t* d_dirty, uint32_t* d_dirty_count, const vxp_mesh_dev* out) {
    if (!ctx || !valid_mode(mode) || !d_dirty_count) return VXP_ERR_BAD_ARG;
    const uint32_t cap = max_dirty(n, m);
    if (!valid_out(out, cap)) return VXP_ERR_BAD_ARG;
    if (m && (!d_voxels || !aligned16(d_voxels) || !d_edits || !d_dirty || n == 0)) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp_status s;
    if ((s = join_chain(ctx)) != VXP_OK) return s;
    uint32_t epoch = 0;
    if (m && (s = next_dirty_flags(ctx, n, &epoch)) != VXP_OK) return s;
    VXP_CUDA(ctx, vxp::launch_apply_edits(d_voxels, d_nbr6, n, d_edits, m, ctx->d_dirty_flags, epoch, d_dirty, d_dirty_count,
                                          ctx->stream));
    VXP_CUDA(ctx, vxp::launch_mesh_list(d_voxels, d_nbr6, n, d_dirty, d_dirty_count, m ? cap : 0, mode == VXP_MODE_GREEDY, *out,
                                        ctx->d_counters + ctx->slot, ctx->stream));
    if (m) {
        ctx->launches += 2;
        ctx->slot ^= 1;
    }
    return VXP_OK;
}

// ------------------------------------------------------------------ chunk wire format (SPEC §10)

vxp_status vxp_pack_chunks(vxp_ctx* ctx, const uint16_t* d_voxels, uint32_t n, uint8_t* d_packed, uint64_t capacity_bytes,
                           uint64_t* d_offsets, uint64_t* d_total_bytes) {
    if (!ctx || !d_total_bytes || (n && (!d_voxels || !aligned16(d_voxels) || !d_offsets))) return VXP_ERR_BAD_ARG;
    if (capacity_bytes && (!d_packed || !aligned16(d_packed))) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    ctx->chain_open = false;
    VXP_CUDA(ctx, vxp::launch_pack_chunks(d_voxels, n, d_packed, capacity_bytes, reinterpret_cast<unsigned long long*>(d_offsets),
                                          reinterpret_cast<unsigned long long*>(d_total_bytes), ctx->d_counters + 2, ctx->stream));
    ctx->launches += n ? 1 : 0;
    return VXP_OK;
}

vxp_status vxp_unpack_chunks(vxp_ctx* ctx, const uint8_t* d_packed, uint64_t packed_bytes, const uint64_t* d_offsets, uint32_t n,
                             uint16_t* d_voxels) {
    if (!ctx || (n && (!d_packed || !aligned16(d_packed) || !d_offsets || !d_voxels || !aligned16(d_voxels)))) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    ctx->chain_open = false;
    VXP_CUDA(ctx, vxp::launch_unpack_chunks(d_packed, packed_bytes, reinterpret_cast<const unsigned long long*>(d_offsets), n, d_voxels,
                                            ctx->d_status, ctx->stream));
    ctx->launches += n ? 1 : 0;
    ctx->status_pending = n != 0;
    return VXP_OK;
}

// ------------------------------------------------------------------ level of detail (SPEC §11)

vxp_status vxp_downsample_chunks(vxp_ctx* ctx, const uint16_t* d_voxels, const int32_t* d_child8, uint32_t n_out, uint16_t* d_out) {
    if (!ctx || (n_out && (!d_voxels || !aligned16(d_voxels) || !d_child8 || !d_out || !aligned16(d_out)))) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    ctx->chain_open = false;
    VXP_CUDA(ctx, vxp::launch_downsample_chunks(d_voxels, d_child8, n_out, d_out, ctx->stream));
    ctx->launches += n_out ? 1 : 0;
    return VXP_OK;
}

// ------------------------------------------------------------------ host level

vxp_status vxp_terrain_fill_host(vxp_ctx* ctx, const vxp_coord* coords, uint32_t n, uint64_t seed,
                                 uint16_t* voxels) {
    if (!ctx || (n && (!coords || !voxels))) return VXP_ERR_BAD_ARG;
    if (n == 0) return VXP_OK;
    DeviceGuard g(ctx->device);
    vxp_status s;
    if ((s = grow_device(ctx, &ctx->d_coords, &ctx->coords_cap, n, 1)) != VXP_OK) return s;
    ctx->world_valid = false;            // the voxel arena is re-used
    if ((s = grow_device(ctx, &ctx->d_voxels, &ctx->voxels_cap, n, VXP_CHUNK_VOXELS)) != VXP_OK) return s;
    ctx->chain_open = false;
    VXP_CUDA(ctx, cudaMemcpyAsync(ctx->d_coords, coords, (size_t)n * sizeof(vxp_coord), cudaMemcpyHostToDevice,
                                  ctx->stream));
    if ((s = vxp_terrain_fill_batch(ctx, ctx->d_coords, n, seed, ctx->d_voxels)) != VXP_OK) return s;   // picks the path by batch size
    VXP_CUDA(ctx, cudaMemcpyAsync(voxels, ctx->d_voxels, (size_t)n * VXP_CHUNK_VOXELS * sizeof(uint16_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VXP_OK;
}

// Host voxels in, host mesh out.  The batch is cut into slices; the upload of slice k+1, the meshing of
// slice k and the download of the quads of slice k-1 overlap (PCIe is full duplex and the kernel is ~50x
// faster than either copy), so the call costs about as much as the upload alone.  A slice is meshed as soon
// as every chunk its neighbour table names has arrived.  If the output arena turns out too small the
// kernels are re-run (voxels are on the device by then) through the plain path.
vxp_status vxp_mesh_chunks_host(vxp_ctx* ctx, const uint16_t* voxels, const int32_t* nbr6, uint32_t n,
                                uint32_t n_stored, vxp_mode mode, vxp_mesh_host* out) {
    if (!ctx || !out || !valid_mode(mode) || n_stored < n || (n && !voxels)) return VXP_ERR_BAD_ARG;
    if (nbr6)
        for (size_t i = 0; i < (size_t)n * 6; ++i)
            if (nbr6[i] < -1 || nbr6[i] >= (int64_t)n_stored) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp_status s;
    ctx->world_valid = false;            // the resident world is being replaced: valid again only when this call has succeeded
    if ((s = grow_device(ctx, &ctx->d_voxels, &ctx->voxels_cap, n_stored, VXP_CHUNK_VOXELS)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_nbr6, &ctx->nbr_cap, n, 6)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_meta, &ctx->meta_cap, n, 1)) != VXP_OK) return s;
    if ((s = grow_pinned(ctx, &ctx->h_meta, &ctx->h_meta_cap, n, 1)) != VXP_OK) return s;
    if ((s = grow_quads(ctx, (size_t)n * 512 + 4096)) != VXP_OK) return s;       // first guess; grows on overflow
    if ((s = grow_host_quads(ctx, ctx->quads_cap)) != VXP_OK) return s;
    const int32_t* d_nbr = nbr6 ? ctx->d_nbr6 : nullptr;
    const int greedy = mode == VXP_MODE_GREEDY;
    const size_t chunk_bytes = (size_t)VXP_CHUNK_VOXELS * sizeof(uint16_t);
    uint64_t total = 0;
    bool done = false;
    if (n) {
        // more, smaller slices shorten the part of the call that nothing overlaps (the last slice's kernel and download);
        // measured on 4096 chunks: 8 slices 774 k chunks/s, 16: 786 k, 32: 766 k
        const uint32_t S = n >= 4096 ? 16u : n >= 2048 ? 8u : n >= 256 ? 4u : 1u;
        const uint32_t slice = (n + S - 1) / S;
        // slice whose upload completes the inputs of slice k
        uint32_t need[vxp_ctx::kMaxSlices];
        for (uint32_t k = 0; k < S; ++k) {
            uint32_t hi_ref = k;
            if (nbr6) {
                const uint32_t lo = k * slice, hi = lo + slice < n ? lo + slice : n;
                for (size_t i = (size_t)lo * 6; i < (size_t)hi * 6; ++i)
                    if (nbr6[i] >= 0 && (uint32_t)nbr6[i] < n && (uint32_t)nbr6[i] / slice > hi_ref) hi_ref = (uint32_t)nbr6[i] / slice;
            }
            need[k] = hi_ref < S ? hi_ref : S - 1;
        }
        vxp::BoundaryScratch scratch{nullptr, nullptr, 0};
        if ((s = next_boundary_scratch(ctx, n, &scratch)) != VXP_OK) return s;
        if (ctx->launched && ctx->launch_stream != ctx->stream) {
            VXP_CUDA(ctx, cudaEventRecord(ctx->ev_chain, ctx->launch_stream));
            VXP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chain, 0));
        }
        ctx->launched = true;
        ctx->launch_stream = ctx->stream;
        ctx->chain_open = false;
        const vxp_mesh_dev dev{ctx->d_verts, ctx->host_indices ? ctx->d_indices : nullptr, ctx->d_meta, ctx->d_total, ctx->quads_cap};
        // uploads: neighbour table and read-only halo chunks first, then the slices in order
        VXP_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
        VXP_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
        VXP_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));
        if (nbr6)
            VXP_CUDA(ctx, cudaMemcpyAsync(ctx->d_nbr6, nbr6, (size_t)n * 6 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_in));
        if (n_stored > n &&
            (s = upload(ctx, ctx->d_voxels + (size_t)n * VXP_CHUNK_VOXELS, voxels + (size_t)n * VXP_CHUNK_VOXELS,
                        (size_t)(n_stored - n) * chunk_bytes, ctx->s_in)) != VXP_OK)
            return s;
        // kernels: slice j as soon as slice need[j] is on the device (enqueued right behind that upload, so that with
        // pageable caller memory - where this thread spends the call copying into the staging buffers - the first
        // kernels and downloads do not wait for the last upload); the running total follows each one to the host
        VXP_CUDA(ctx, cudaMemsetAsync(ctx->d_total, 0, sizeof(uint64_t), ctx->stream));
        uint32_t launched = 0;
        for (uint32_t k = 0; k < S; ++k) {
            const uint32_t lo = k * slice, hi = lo + slice < n ? lo + slice : n;
            if (lo < hi && (s = upload(ctx, ctx->d_voxels + (size_t)lo * VXP_CHUNK_VOXELS, voxels + (size_t)lo * VXP_CHUNK_VOXELS,
                                       (size_t)(hi - lo) * chunk_bytes, ctx->s_in)) != VXP_OK)
                return s;
            VXP_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[k], ctx->s_in));
            while (launched < S && need[launched] <= k) {
                const uint32_t j = launched++;
                VXP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[need[j]], 0));
                VXP_CUDA(ctx, vxp::launch_mesh_range(ctx->d_voxels, d_nbr, n, j * slice, (j + 1) * slice, greedy, dev, scratch, ctx->stream));
                ctx->launches += 1;
                VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_totals + j, ctx->d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
                VXP_CUDA(ctx, cudaEventRecord(ctx->ev_k[j], ctx->stream));
            }
        }
        // downloads: the quads of slice k as soon as its kernel is done
        uint64_t prev = 0;
        bool overflow = false;
        for (uint32_t k = 0; k < S; ++k) {
            VXP_CUDA(ctx, cudaEventSynchronize(ctx->ev_k[k]));
            const uint64_t t = ctx->h_totals[k];
            if (t > ctx->quads_cap) { overflow = true; total = t; continue; }
            if (!overflow && t > prev) {
                VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_verts + prev * 4, ctx->d_verts + prev * 4, (t - prev) * 4 * sizeof(uint64_t),
                                              cudaMemcpyDeviceToHost, ctx->s_out));
                if (ctx->host_indices)
                    VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_indices + prev * 6, ctx->d_indices + prev * 6, (t - prev) * 6 * sizeof(uint32_t),
                                                  cudaMemcpyDeviceToHost, ctx->s_out));
            }
            prev = t;
        }
        if (!overflow) {
            total = prev;
            VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_meta, ctx->d_meta, (size_t)n * sizeof(vxp_chunk_meta), cudaMemcpyDeviceToHost, ctx->s_out));
            VXP_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
            done = true;
        } else {
            VXP_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
            VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if ((s = grow_quads(ctx, total)) != VXP_OK) return s;
        }
    }
    auto resident = [&] {                // the world stays on the device for vxp_remesh_edits_host
        ctx->world_n = n;
        ctx->world_stored = n_stored;
        ctx->world_has_nbr = nbr6 != nullptr;
        ctx->world_valid = n > 0;
    };
    if (done || n == 0) {
        out->verts = ctx->h_verts;
        out->indices = ctx->host_indices ? ctx->h_indices : nullptr;
        out->meta = ctx->h_meta;
        out->total_quads = total;
        resident();
        return VXP_OK;
    }
    s = mesh_into_host(
        ctx, n,
        [&](const vxp_mesh_dev& dev) {
            return chained_mesh_launch(ctx, ctx->d_voxels, d_nbr, n, greedy, dev) == VXP_OK ? cudaSuccess : cudaErrorUnknown;
        },
        out);
    if (s == VXP_OK) resident();
    return s;
}

vxp_status vxp_terrain_mesh_host(vxp_ctx* ctx, const vxp_coord* coords, uint32_t n, uint64_t seed,
                                 vxp_mode mode, vxp_mesh_host* out) {
    if (!ctx || !out || !valid_mode(mode) || (n && !coords)) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp_status s;
    if ((s = grow_device(ctx, &ctx->d_coords, &ctx->coords_cap, n, 1)) != VXP_OK) return s;
    if (n)
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->d_coords, coords, (size_t)n * sizeof(vxp_coord),
                                      cudaMemcpyHostToDevice, ctx->stream));
    return mesh_into_host(
        ctx, n,
        [&](const vxp_mesh_dev& dev) {
            vxp::ColumnScratch cols{};
            if (n && column_scratch(ctx, n, &cols) != VXP_OK) return cudaErrorMemoryAllocation;
            ctx->launches += n ? 4 : 0;
            return vxp::launch_terrain_mesh(ctx->d_coords, n, seed, mode == VXP_MODE_GREEDY, dev, cols, ctx->stream);
        },
        out);
}

namespace {
bool valid_record(const uint8_t* packed, uint64_t packed_bytes, uint64_t off) {
    if ((off & 15u) || off > packed_bytes || packed_bytes - off < VXP_PACKED_HEADER_BYTES) return false;
    uint32_t h[3];
    std::memcpy(h, packed + off, sizeof h);
    const uint32_t bits = h[1] & 0xFFFFu, len = h[1] >> 16;
    if (h[0] != 0x31505856u || h[2] != 4096u * bits) return false;
    if (!(bits == 0 || bits == 1 || bits == 2 || bits == 4 || bits == 16)) return false;
    if (bits == 16 ? len != 0 : (len < 1 || len > 16 || len > (1u << bits))) return false;
    return packed_bytes - off - VXP_PACKED_HEADER_BYTES >= h[2];
}
}  // namespace

vxp_status vxp_mesh_packed_host(vxp_ctx* ctx, const uint8_t* packed, uint64_t packed_bytes, const uint64_t* offsets,
                                const int32_t* nbr6, uint32_t n, uint32_t n_stored, vxp_mode mode, vxp_mesh_host* out) {
    if (!ctx || !out || !valid_mode(mode) || n_stored < n || (n_stored && (!packed || !offsets))) return VXP_ERR_BAD_ARG;
    for (uint32_t i = 0; i < n_stored; ++i)
        if (!valid_record(packed, packed_bytes, offsets[i])) return VXP_ERR_BAD_ARG;
    if (nbr6)
        for (size_t i = 0; i < (size_t)n * 6; ++i)
            if (nbr6[i] < -1 || nbr6[i] >= (int64_t)n_stored) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp_status s;
    ctx->world_valid = false;            // valid again only when this call has succeeded
    if ((s = grow_device(ctx, &ctx->d_voxels, &ctx->voxels_cap, n_stored, VXP_CHUNK_VOXELS)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_nbr6, &ctx->nbr_cap, n, 6)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_packed, &ctx->packed_cap, (size_t)packed_bytes + 16, 1)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_offsets, &ctx->offsets_cap, n_stored, 1)) != VXP_OK) return s;
    if ((s = join_chain(ctx)) != VXP_OK) return s;
    if (ctx->status_pending && (s = vxp_ctx_synchronize(ctx)) != VXP_OK) return s;   // an earlier vxp_unpack_chunks verdict stays that call's
    const int32_t* d_nbr = nbr6 ? ctx->d_nbr6 : nullptr;
    if (n_stored) {
        if ((s = upload(ctx, ctx->d_packed, packed, packed_bytes, ctx->stream)) != VXP_OK) return s;
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->d_offsets, offsets, (size_t)n_stored * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        if (nbr6) VXP_CUDA(ctx, cudaMemcpyAsync(ctx->d_nbr6, nbr6, (size_t)n * 6 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        VXP_CUDA(ctx, vxp::launch_unpack_chunks(ctx->d_packed, packed_bytes, ctx->d_offsets, n_stored, ctx->d_voxels, ctx->d_status, ctx->stream));
        ctx->launches += 1;
        ctx->status_pending = true;
    }
    const int greedy = mode == VXP_MODE_GREEDY;
    s = mesh_into_host(
        ctx, n,
        [&](const vxp_mesh_dev& dev) {
            return chained_mesh_launch(ctx, ctx->d_voxels, d_nbr, n, greedy, dev) == VXP_OK ? cudaSuccess : cudaErrorUnknown;
        },
        out);
    if (s == VXP_OK) s = unpack_verdict(ctx);      // mesh_into_host has synchronised: a malformed record fails the call
    if (s == VXP_OK) {
        ctx->world_n = n;
        ctx->world_stored = n_stored;
        ctx->world_has_nbr = nbr6 != nullptr;
        ctx->world_valid = n > 0;
    }
    return s;
}

// Edits go up (12 bytes each), the voxels are patched and the dirty chunks re-meshed on the device, and only the
// dirty list and the new meshes come down.  Two synchronisations: one to learn how many quads to fetch, one at the end.
vxp_status vxp_remesh_edits_host(vxp_ctx* ctx, const vxp_edit* edits, uint32_t m, vxp_mode mode, vxp_mesh_host* out,
                                 const uint32_t** dirty, uint32_t* count) {
    if (!ctx || !out || !dirty || !count || !valid_mode(mode) || (m && !edits)) return VXP_ERR_BAD_ARG;
    if (!ctx->world_valid) return VXP_ERR_BAD_ARG;
    const uint32_t n = ctx->world_n;
    for (uint32_t i = 0; i < m; ++i)
        if (edits[i].chunk >= n || edits[i].x >= 32 || edits[i].y >= 32 || edits[i].z >= 32) return VXP_ERR_BAD_ARG;
    DeviceGuard g(ctx->device);
    vxp_status s;
    const uint32_t cap = max_dirty(n, m);
    if ((s = grow_device(ctx, &ctx->d_edits, &ctx->edits_cap, m, 1)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_dirty, &ctx->dirty_list_cap, cap, 1)) != VXP_OK) return s;
    if ((s = grow_pinned(ctx, &ctx->h_dirty, &ctx->h_dirty_cap, cap, 1)) != VXP_OK) return s;
    if ((s = grow_pinned(ctx, &ctx->h_meta, &ctx->h_meta_cap, cap, 1)) != VXP_OK) return s;
    if ((s = grow_device(ctx, &ctx->d_meta, &ctx->meta_cap, cap, 1)) != VXP_OK) return s;
    if ((s = grow_quads(ctx, (size_t)cap * 1024 + 4096)) != VXP_OK) return s;
    uint64_t total = 0;
    uint32_t cnt = 0;
    if (m) VXP_CUDA(ctx, cudaMemcpyAsync(ctx->d_edits, edits, (size_t)m * sizeof(vxp_edit), cudaMemcpyHostToDevice, ctx->stream));
    for (int attempt = 0; attempt < 2; ++attempt) {
        const vxp_mesh_dev dev{ctx->d_verts, ctx->host_indices ? ctx->d_indices : nullptr, ctx->d_meta, ctx->d_total, ctx->quads_cap};
        if (attempt == 0) {
            s = vxp_remesh_edits(ctx, ctx->d_voxels, ctx->world_has_nbr ? ctx->d_nbr6 : nullptr, n, ctx->d_edits, m, mode,
                                 ctx->d_dirty, ctx->d_dirty_count, &dev);
        } else {   // the arena was too small: the voxels are already patched and the list is on the device
            VXP_CUDA(ctx, vxp::launch_mesh_list(ctx->d_voxels, ctx->world_has_nbr ? ctx->d_nbr6 : nullptr, n, ctx->d_dirty,
                                                ctx->d_dirty_count, cap, mode == VXP_MODE_GREEDY, dev,
                                                ctx->d_counters + ctx->slot, ctx->stream));
            ctx->launches += 1;
            ctx->slot ^= 1;
            s = VXP_OK;
        }
        if (s != VXP_OK) return s;
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_total, ctx->d_total, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_dirty_count, ctx->d_dirty_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (cap) {
            VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_dirty, ctx->d_dirty, (size_t)cap * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_meta, ctx->d_meta, (size_t)cap * sizeof(vxp_chunk_meta), cudaMemcpyDeviceToHost, ctx->stream));
        }
        VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        total = m ? *ctx->h_total : 0;
        cnt = m ? *ctx->h_dirty_count : 0;
        if (total <= ctx->quads_cap) break;
        if (attempt == 1) return VXP_ERR_CAPACITY;
        if ((s = grow_quads(ctx, total)) != VXP_OK) return s;
    }
    if ((s = grow_host_quads(ctx, total)) != VXP_OK) return s;
    if (total) {
        VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_verts, ctx->d_verts, total * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->host_indices)
            VXP_CUDA(ctx, cudaMemcpyAsync(ctx->h_indices, ctx->d_indices, total * 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        VXP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    out->verts = ctx->h_verts;
    out->indices = ctx->host_indices ? ctx->h_indices : nullptr;
    out->meta = ctx->h_meta;
    out->total_quads = total;
    *dirty = ctx->h_dirty;
    *count = cnt;
    return VXP_OK;
}

}  // extern "C"
