// vxp_mesh_common.cuh — building blocks shared by the meshing kernels (SPEC §2-§5): launch-chaining and
// boundary-exchange plumbing, the halo, the quad-record store, the culled face words.
//
// Halo.  The +-Y / +-Z halo rows are 64-byte rows of the neighbours, fetched at kernel start.
// The +-X halo of a chunk is one u16 out of every 64-byte row of the neighbour: 1024 sectors
// per face if gathered from the voxels.  Instead every CTA publishes the two x-boundary bit
// planes of its own chunk (epoch-stamped words, or just an "all air" / "all solid" tag) to a
// context-owned scratch area as soon as its occupancy is known, and a neighbour that finds
// words of the current launch reads those planes; otherwise it gathers from the voxels.
// Nobody ever waits, so there is no ordering requirement between CTAs, and both paths yield
// the same bits.
#pragma once
#include <type_traits>
#include "vxp_device.cuh"
#include "../../include/vxp.h"

namespace vxp {
namespace mesh {

constexpr int kPlaneWords = 32 * kPad;   // 1056 words per padded 32x32 word plane set
constexpr int kSlabBytes = 16384;        // a chunk is streamed in slabs of 8 z-layers
constexpr int kSlabs = 4;

// ---- optional per-phase cycle accounting (profiling build only: -DVXP_PROFILE, libvxp_prof.so)
#ifdef VXP_PROFILE
constexpr int kProfSlots = 16;
constexpr int kProfCtas = 65536;
__device__ unsigned long long g_prof[kProfCtas][kProfSlots];
__device__ unsigned int g_prof_warp[kProfCtas][8];   // per-warp cycles of the greedy sweep
__device__ unsigned long long g_prof_time[kProfCtas][2];   // %globaltimer (ns) at CTA start / end
__device__ __forceinline__ unsigned long long prof_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct Prof {
    long long t;
    bool on;
    __device__ __forceinline__ void start(bool enable) { on = enable && blockIdx.x < kProfCtas; t = clock64(); }
    __device__ __forceinline__ void mark(int slot) {
        // plain stores only (one chunk per CTA, every slot written at most once): a read-modify-write
        // would put a global round trip into every phase it tries to measure
        if (on) { const long long n = clock64(); g_prof[blockIdx.x][slot] = (unsigned long long)(n - t); t = n; }
    }
    __device__ __forceinline__ void count(int slot, unsigned long long v = 1ull) { if (on) g_prof[blockIdx.x][slot] = v; }
};
#else
struct Prof {
    __device__ __forceinline__ void start(bool) {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void count(int, unsigned long long = 1ull) {}
};
#endif

// ---- boundary summaries (context-owned scratch, one entry per meshed chunk)
constexpr uint32_t kTagAir = 1u, kTagSolid = 2u;   // low 2 bits of a tag; the rest is the launch epoch
struct Bnd {
    uint32_t* tags;               // n words: epoch<<2 | kTag*  (only chunks that are all air / all solid)
    unsigned long long* planes;   // n x 2 x 32 words: epoch<<32 | bits; [0] = the chunk's x=0 layer, [1] = x=31 (word z, bit y)
    uint32_t epoch;               // in [1, 2^30)
};
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Quad counter of a launch.  Low kQuadBits bits: quads reserved so far (a chunk's atomic add returns the
// first quad of its range); the bits above count the CTAs that have checked in (MeshOut::cta_inc = kCtaOne;
// 0 when nobody waits for the count: fused path, ranges of the sliced host pipeline).  One word, so the
// check-in of a chunk with quads rides on the reservation it performs anyway and needs no ordering of its own.
constexpr int kQuadBits = 42;                                 // 98 304 quads x 2.9 M chunks (180 GB of voxels) < 2^42
constexpr unsigned long long kQuadMask = (1ull << kQuadBits) - 1ull;
constexpr unsigned long long kCtaOne = 1ull << kQuadBits;     // at most 2^22 CTAs per launch
struct MeshOut {
    unsigned long long* verts;
    uint32_t* indices;
    vxp_chunk_meta* meta;
    unsigned long long* total;     // the counter the chunks reserve from (context-owned slot, or the caller's total_quads)
    unsigned long long cap;
    unsigned long long cta_inc;    // kCtaOne: every CTA checks in exactly once (mesh_kernel); 0: plain quad counter
};
// Reserve Q quads (and check the CTA in); returns the first quad of the range.  Thread 0 only.
template <class SM>
__device__ __forceinline__ unsigned long long reserve_quads(SM& sm, const MeshOut& out, uint32_t Q) {
    sm.counted = 1u;
    return atomicAdd(out.total, (unsigned long long)Q + out.cta_inc) & kQuadMask;
}

// 8 consecutive u16 -> 8 bits (bit j = element j non-zero): 4 VIMNMX.U16x2 + 3 LEA + 2 logic
__device__ __forceinline__ uint32_t nz8(uint4 q) {
    const uint32_t a = __vminu2(q.x, 0x00010001u), b = __vminu2(q.y, 0x00010001u);
    const uint32_t c = __vminu2(q.z, 0x00010001u), d = __vminu2(q.w, 0x00010001u);
    const uint32_t acc = (a + (b << 2)) + ((c + (d << 2)) << 4);   // even elements: bits 0,2,4,6; odd: 16,18,20,22
    return (acc | (acc >> 15)) & 0xFFu;
}

struct ChunkFlags {
    bool any_solid, all_solid, single_id, small_ids;
};

// ---- id bit planes (greedy).  A chunk whose block ids all fit in kIdPlanes bits (small palettes: terrain uses
// 1..3) gets its three equality volumes from bit planes: bit b of every id is pulled out while the tile is converted
// (speculatively: whether the ids fit is only known at the end), and the equality rows are then three xors per
// plane, 32 voxels per instruction (build_masks).  Chunks with larger ids take the per-voxel tile pass
// (equality_to_regs).  Measured on config 3: the tile pass costs a surface chunk 8.0 k cycles; pulling the planes
// out of the tile in a pass of their own 4.5 k; during the conversion 1.5 k.
constexpr int kIdPlanes = 3;
struct IdPlanes {
    uint32_t w[kIdPlanes][kSlabs];   // w[b][s]: bit x = bit b of voxel x of this thread's row of slab s (convert_row_z / convert_row_y)
};
// The low kIdPlanes bits of 8 consecutive u16, one plane byte each (bit j of byte b = bit b of element j), placed
// in byte k of p0 / p1 / p2.  Three steps: the low bytes of the 8 elements are packed into two words (2 PRMT) and
// folded into one (elements 0..3 in the low nibbles of the bytes, 4..7 in the high nibbles); plane b is then the
// 8 bits (0x11111111 << b) of that word, and one multiply gathers them into the top byte: bit 8i + 4h + b times
// 2^(24 - 7i - b) lands on bit 24 + i + 4h, and no two cross terms share a bit position (no carries).
// 14 instructions, three of them on the FMA pipe, instead of 3 x 13.
template <int k>
__device__ __forceinline__ void id_plane_bytes(uint4 q, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    static_assert(kIdPlanes == 3, "three planes are pulled out here");
    const uint32_t lo = __byte_perm(q.x, q.y, 0x6420), hi = __byte_perm(q.z, q.w, 0x6420);
    const uint32_t u = (lo & 0x07070707u) | ((hi & 0x07070707u) << 4);
    constexpr uint32_t sel = k == 0 ? 0x3217u : k == 1 ? 0x3270u : k == 2 ? 0x3710u : 0x7210u;   // byte k <- top byte of the product
    p0 = __byte_perm(p0, (u & 0x11111111u) * 0x01020408u, sel);
    p1 = __byte_perm(p1, (u & 0x22222222u) * 0x00810204u, sel);
    p2 = __byte_perm(p2, (u & 0x44444444u) * 0x00408102u, sel);
}

// ------------------------------------------------------------------ halo
// +-Y / +-Z faces: one 64-byte row of the neighbour per lane, fetched by warps 2..5 at kernel start
// (so the round trip overlaps the tile load) and turned into halo words after the conversion.
struct HaloRows {
    uint4 q[4];
};
__device__ __forceinline__ void issue_yz_halo(HaloRows& h, const uint16_t* __restrict__ voxels, int nb) {
    const int lane = threadIdx.x & 31, f = threadIdx.x >> 5;
    h.q[0] = h.q[1] = h.q[2] = h.q[3] = make_uint4(0, 0, 0, 0);
    if (f < 2 || f >= 6 || nb < 0) return;
    const int row = f == 2 ? lane * 32 + 31 : f == 3 ? lane * 32 : f == 4 ? 31 * 32 + lane : lane;   // (z*32 + y) in the neighbour
    const uint4* r4 = reinterpret_cast<const uint4*>(voxels + (size_t)nb * VXP_CHUNK_VOXELS + row * 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) h.q[k] = __ldg(r4 + k);
}

// +-X faces: the chunk's own x=0 / x=31 boundary planes (word z, bit y) are published to the
// context scratch as self-validating 64-bit words (launch epoch in the high half), so readers
// need no ordering and a single round trip; chunks that are all air / all solid publish one
// tag word instead.
template <class SM>
__device__ __forceinline__ void publish_boundary(const SM& sm, const Bnd& bnd, uint32_t chunk, ChunkFlags fl) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!fl.any_solid || fl.all_solid) {
        if (tid == kThreads - 32) st_relaxed_u32(bnd.tags + chunk, bnd.epoch << 2 | (fl.any_solid ? kTagSolid : kTagAir));
        return;
    }
    unsigned long long* dst = bnd.planes + (size_t)chunk * 64;
    const unsigned long long hi_epoch = (unsigned long long)bnd.epoch << 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = warp * 4 + k;
        const uint32_t c = sm.occ[(z + 1) * 34 + lane + 1];
        const uint32_t lo = __ballot_sync(kFull, c & 1u), hi = __ballot_sync(kFull, c >> 31);
        if (lane == 0) {
            st_relaxed_u64(dst + z, hi_epoch | lo);
            st_relaxed_u64(dst + 32 + z, hi_epoch | hi);
        }
    }
}

// Fill the halo of sm.occ / sm.xh: warp f < 6 produces the 32 words of face f.  Returns (per
// thread) whether any halo bit this thread produced is air.  Contains no barrier.
// nb = neighbour index of face f = warp index (loaded by the caller before the tile wait).
#ifndef VXP_PLANE_POLLS
#define VXP_PLANE_POLLS 16
#endif
constexpr int kPlanePolls = VXP_PLANE_POLLS;   // bounded re-reads of a boundary plane that is not there yet (never a wait-for);
                                               // measured on config 2: round 1 (tile kernel) 1 -> 109.4 us, 3 -> 104.9, 6 -> 104.1, 12 -> 103.5;
                                               // round 2: 4 -> 67.6, 8 -> 66.6, 16 -> 66.0, 32 / 64 -> 65.9
// The halo is produced in two steps so that callers with something else to do can put it between them:
// halo_begin   warps 2..5 write their +-Y / +-Z halo rows; warps 0,1 issue the first read of their x
//              neighbour's tag and boundary plane (two loads in flight together, not waited for)
// halo_finish  warps 0,1 use what came back, re-read a bounded number of times if the words were not there
//              yet, and finally gather from the neighbour's voxels; they write sm.xh
// Both return (per thread) whether any halo bit this thread produced is air.  Neither contains a barrier.
struct XPoll {
    unsigned long long pw;
    uint32_t tag;
    bool armed;
};
template <class SM>
__device__ __forceinline__ bool halo_begin(SM& sm, int nb, uint32_t n, const Bnd& bnd, const HaloRows& rows, XPoll& xp) {
    const int tid = threadIdx.x, lane = tid & 31, f = tid >> 5;
    xp.armed = false;
    xp.tag = 0;
    xp.pw = 0;
    if (f >= 6) return false;
    if (f >= 2) {
        uint32_t word = 0;
        if (nb >= 0) word = nz8(rows.q[0]) | nz8(rows.q[1]) << 8 | nz8(rows.q[2]) << 16 | nz8(rows.q[3]) << 24;
        sm.occ[f == 2 ? (lane + 1) * 34 : f == 3 ? (lane + 1) * 34 + 33 : f == 4 ? lane + 1 : 33 * 34 + lane + 1] = word;
        return word != kFull;
    }
    if (nb >= 0 && bnd.tags && (uint32_t)nb < n) {
        xp.tag = ld_relaxed_u32(bnd.tags + nb);
        xp.pw = ld_relaxed_u64(bnd.planes + (size_t)nb * 64 + (f ^ 1) * 32 + lane);
        xp.armed = true;
    }
    return false;
}
template <class SM>
__device__ __forceinline__ bool halo_finish(SM& sm, const uint16_t* __restrict__ voxels, int nb, uint32_t n, const Bnd& bnd,
                                            XPoll xp, Prof& prof) {
    const int tid = threadIdx.x, lane = tid & 31, f = tid >> 5;
    if (f >= 2) return false;
    uint32_t word = 0;
    if (nb >= 0) {
        bool have = false;
        if (xp.armed) {
            const uint32_t* tagp = bnd.tags + nb;
            const unsigned long long* planep = bnd.planes + (size_t)nb * 64 + (f ^ 1) * 32 + lane;
            for (int t = 0;; ++t) {
                if ((xp.tag >> 2) == bnd.epoch) {                       // all air / all solid neighbour
                    word = (xp.tag & 3u) == kTagSolid ? kFull : 0u;
                    have = true;
                } else if (__all_sync(kFull, (uint32_t)(xp.pw >> 32) == bnd.epoch)) {
                    word = (uint32_t)xp.pw;
                    have = true;
                }
                if (lane == 0) prof.count(15, (unsigned long long)(t + 1));
                if (have || t + 1 >= kPlanePolls) break;
                xp.tag = ld_relaxed_u32(tagp);                          // both loads in flight together
                xp.pw = ld_relaxed_u64(planep);
            }
            if (lane == 0) prof.count(have ? 12 : 13);
        }
        if (!have) {                                      // gather from the neighbour's voxels: one u16 per 64-byte row
            const uint16_t* nv = voxels + (size_t)nb * VXP_CHUNK_VOXELS + (f ? 0 : 31) + 32 * lane;
#pragma unroll 8
            for (int z = 0; z < 32; ++z) {
                const uint32_t v = __ldg(nv + 1024 * z);
                const uint32_t w = __ballot_sync(kFull, v != 0);
                if (lane == z) word = w;
            }
        }
    }
    sm.xh[f][lane] = word;
    return word != kFull;
}
// ------------------------------------------------------------------ active rows
// actz[z] bit y = row (y,z) has a solid voxel with an air neighbour (halo included).
template <class SM>
__device__ __forceinline__ void build_active(SM& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = warp * 4 + k, y = lane;
        const uint32_t* o = sm.occ + (z + 1) * 34 + (y + 1);
        const uint32_t c = o[0];
        const uint32_t lo = (sm.xh[0][z] >> y) & 1u, hi = (sm.xh[1][z] >> y) & 1u;
        const uint32_t buried = o[-1] & o[1] & o[-34] & o[34] & ((c << 1) | lo) & ((c >> 1) | (hi << 31));
        const uint32_t act = __ballot_sync(kFull, (c & ~buried) != 0);
        if (lane == 0) sm.actz[z] = act;
    }
}

// ---- the normal path.  Every instruction on the row recurrence costs its full latency (the
// chain has no parallelism and a CTA has few other warps to hide it), so the sweep proper does the
// bare minimum per row and everything else happens afterwards in parallel:
//
//   sweep_log    (one thread per plane, whole warps in lock step) visits only the rows that can
//                change anything - row v matters iff mask row v or v-1 is non-empty - and logs, for
//                each non-empty row, the cover Cs of the runs that survive into it, in the plane's
//                private segment of the log.  No start rows, no atomics.
//   stage_births (all threads, one log entry each) recomputes the runs that START at the entry's row
//                (cells not covered by survivors, split at id changes) and stages one record per
//                run, height still 1.
//   run_height   (one thread per record, inside the flush) walks the following entries of the plane
//                while the run's first column stays covered.
struct SweepLog {
    uint2* entry;          // [nL] .x = survivors' cover of a visited row, .y = plane<<5 | row | (runs started by the
                           //      plane's earlier rows) << 13; plane segments back to back
    const uint32_t* seg;   // [192] first entry of each plane
    const uint32_t* rows;  // [192] non-empty (= logged) rows of each plane (bit mask)
    const uint32_t* first; // [192] ordinal of the plane's first run
};

// Height of the run of a staged record: rows v+1.. whose log entry keeps column u covered.
__device__ __forceinline__ uint32_t run_height(const SweepLog& log, uint32_t rec) {
    const uint32_t u = rec & 31u, v = (rec >> 5) & 31u, plane = (rec >> 10) & 255u;
    const uint32_t rows = log.rows[plane];
    const uint2* en = log.entry + log.seg[plane] + __popc(rows & ((1u << v) - 1u));   // the entry of row v
    uint32_t h = 1;
    while (v + h < 32u && ((rows >> (v + h)) & 1u) && ((en[h].x >> u) & 1u)) ++h;
    return h;
}

// ------------------------------------------------------------------ staged records -> buffers
__device__ __forceinline__ void st_global_v4u64(void* p, unsigned long long a, unsigned long long b,
                                                unsigned long long c, unsigned long long d) {
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
// chunk-local voxel index of the origin cell of a staged record
__device__ __forceinline__ int rec_origin_voxel(uint32_t rec) {
    const uint32_t u = rec & 31u, v = (rec >> 5) & 31u, s = (rec >> 10) & 31u, axis = (rec >> 16) & 3u;
    return (int)(s << (5u * axis) | u << (axis == 0 ? 5 : 0) | v << (axis == 2 ? 5 : 10));
}
// SPEC §5: the four vertices of a staged record with block id b, as one 32-byte store.
__device__ __forceinline__ void store_quad(unsigned long long* dst, uint32_t rec, uint32_t b) {
    const uint32_t u = rec & 31u, v = (rec >> 5) & 31u, s = (rec >> 10) & 31u, f = (rec >> 15) & 7u;
    const uint32_t w = ((rec >> 18) & 31u) + 1u, h = ((rec >> 23) & 31u) + 1u, axis = f >> 1;
    const uint32_t su = axis == 0 ? 6u : 0u, sv = axis == 2 ? 6u : 12u;          // bit offsets of U, V in word0
    const uint32_t c0 = (s + (f & 1u)) << (6u * axis) | u << su | v << sv | f << 18;
    const uint32_t du = w << su, dv = h << sv;
    const bool flip = (0x19u >> f) & 1u;                                         // f in {0,3,4}: c0,c3,c2,c1
    const uint32_t k1 = c0 + (flip ? dv : du), k3 = c0 + (flip ? du : dv), k2 = c0 + du + dv;
    const unsigned long long w1 = (unsigned long long)(b | ((rec >> 18) & 0x3FFu) << 16) << 32;
    st_global_v4u64(dst, w1 | c0, w1 | (k1 | 1u << 21), w1 | (k2 | 2u << 21), w1 | (k3 | 3u << 21));
}

// Writes quads [ord0, ord0+cnt) (chunk-local ordinals) of the chunk whose range starts at sm.first.
// The first batch of records and block ids is fetched before `before_first()` + the optional
// barrier, so the latency of the range reservation (a contended global atomic) overlaps the
// id gather instead of preceding it.  Returns false if the chunk does not fit the arena.
template <class SM, class RecFn, class IdFn, class Pre>
__device__ __forceinline__ bool flush_staging(SM& sm, RecFn rec_at, uint32_t cnt, uint32_t ord0, const MeshOut& out,
                                              IdFn id_of, Pre before_first, bool barrier) {
    const int tid = threadIdx.x;
    uint32_t rec[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t i = tid + k * kThreads;
        rec[k] = i < cnt ? rec_at(i) : 0u;
        b[k] = i < cnt ? id_of(rec_origin_voxel(rec[k])) : 0u;
    }
    before_first();
    if (barrier) __syncthreads();
    const unsigned long long first = sm.first;
    if (first == ~0ull) return false;
    unsigned long long* const vdst = out.verts + (first + ord0) * 4;
    // vertices: one thread per quad, one 32-byte store (a full sector)
    for (uint32_t i0 = tid; i0 < cnt; i0 += 4 * kThreads) {
        if (i0 != (uint32_t)tid) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = i0 + k * kThreads;
                rec[k] = i < cnt ? rec_at(i) : 0u;
                b[k] = i < cnt ? id_of(rec_origin_voxel(rec[k])) : 0u;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t i = i0 + k * kThreads;
            if (i < cnt) store_quad(vdst + (size_t)i * 4, rec[k], b[k]);
        }
    }
    // indices: 8-byte pairs, three per quad -> a warp stores 256 contiguous bytes
    if (out.indices == nullptr) return true;           // the caller binds a static index buffer (vxp_mesh_dev.indices == NULL)
    uint32_t* const idst = out.indices + (first + ord0) * 6;
    uint32_t j = (uint32_t)tid / 3u, part = (uint32_t)tid - 3u * j;
    for (uint32_t i = tid; i < 3 * cnt; i += kThreads) {
        const uint32_t b4 = 4 * (ord0 + j);
        st_global_v2u32(idst + 2 * (size_t)i, b4 + (part == 0 ? 0u : part + 1u), b4 + (part == 0 ? 1u : part == 1 ? 2u : 0u));
        j += kThreads / 3;                       // 256 = 85*3 + 1
        if (++part == 3u) { part = 0u; ++j; }
    }
    return true;
}

// Thread 0: turn the result of the range reservation into sm.first and the chunk's meta entry.
template <class SM>
__device__ __forceinline__ void publish_range(SM& sm, uint32_t chunk, uint32_t Q, unsigned long long f,
                                              const MeshOut& out) {
    const bool fits = f + Q <= out.cap;
    out.meta[chunk] = vxp_chunk_meta{fits ? (uint32_t)f : 0xFFFFFFFFu, Q};
    sm.first = fits ? f : ~0ull;
}

// ------------------------------------------------------------------ culled: occupancy -> quads -> buffers
// Every visible face is a quad.  No mask arrays, no transposes: thread (warp w, lane y) owns the rows
// (y, z = 4w + k) and derives their face words (bit x = face f of voxel (x,y,z) visible) from the occupancy:
// the six words of a row share its five occupancy loads, and a row of air - half of a surface chunk - costs
// one load and a test.  culled_count: quads | non-empty words << 17 of the thread's words of the x faces
// (kX) or of the y and z faces, which do not depend on the x halo, so mesh_chunk counts them while that halo
// is in flight.  culled_emit: one packed block scan gives every thread its first ordinal and its slot in
// the word list; the non-empty words are listed with the ordinal of their first quad; the threads share the
// list evenly when they stage the quad records.
template <class SM>
__device__ __forceinline__ bool row_words_yz(const SM& sm, int y, int z, uint32_t (&wd)[6]) {
    const uint32_t* o = sm.occ + (z + 1) * 34 + (y + 1);
    const uint32_t c = o[0];
    if (c == 0) return false;
    wd[2] = c & ~o[-1];
    wd[3] = c & ~o[+1];
    wd[4] = c & ~o[-34];
    wd[5] = c & ~o[+34];
    return true;
}
template <class SM>
__device__ __forceinline__ bool row_words_x(const SM& sm, int y, int z, uint32_t (&wd)[6]) {
    const uint32_t c = sm.occ[(z + 1) * 34 + (y + 1)];
    if (c == 0) return false;
    wd[0] = c & ~((c << 1) | ((sm.xh[0][z] >> y) & 1u));
    wd[1] = c & ~((c >> 1) | ((sm.xh[1][z] >> y) & 1u) << 31);
    return true;
}
template <bool kX, class SM>
__device__ __forceinline__ uint32_t culled_count(const SM& sm) {
    const int y = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t wd[6];
        if (!(kX ? row_words_x(sm, y, 4 * w + k, wd) : row_words_yz(sm, y, 4 * w + k, wd))) continue;
#pragma unroll
        for (int f = kX ? 0 : 2; f < (kX ? 2 : 6); ++f) packed += (uint32_t)__popc(wd[f]) + (wd[f] ? 1u << 17 : 0u);
    }
    return packed;
}

// ---- launch chaining.  Consecutive mesh launches of a context may overlap on the device (programmatic
// dependent launch): every CTA lets the next launch start as soon as it is resident itself, so the next
// launch's CTAs fill the slots that the tail of this one leaves empty.  What makes that safe:
//   * launches alternate between two context-owned slots (counter + boundary scratch), so launch k+1 shares
//     nothing with launch k (the caller alternates the output arrays, see vxp_ctx_set_overlap);
//   * the LAST CTA of every launch waits for the previous launch to complete before it lets the next one
//     start, so launch k+2 - which re-uses the slot of launch k - can never begin before k has completed;
//   * the same CTA closes the launch: it waits until every CTA has checked in at the slot counter, hands the
//     quad total to the caller's total_quads and leaves the slot zeroed for launch k+2.  Nobody else waits.
// Without the launch attribute (the default) both instructions are no-ops and launches serialise as usual.
struct Chain {
    unsigned long long* user_total;   // the caller's total_quads; nullptr: no check-in, no hand-over (MeshOut::cta_inc == 0)
    // list mode (incremental re-mesh): CTA b meshes chunk list[b] into meta entry b, for b < *list_count; the
    // grid is an upper bound of the count, which only the device knows.  nullptr: CTA b meshes chunk chunk0 + b.
    const uint32_t* list;
    const uint32_t* list_count;       // nullptr: the whole grid is the list
    // Whole-batch launches deal the chunks out in runs of 16 (x neighbours stay next to each other in time, which
    // the boundary exchange relies on), run g of the grid being run (g * run_stride) % runs of the batch (a small
    // stride coprime to runs): stored
    // worlds come sorted by coordinate, and a stride spreads their heavy (surface) and light (air, buried) layers
    // evenly over the launch instead of sending them through the GPU one after the other.  runs == 0: identity.
    uint32_t runs, run_stride;
};
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

} // namespace mesh
} // namespace vxp
