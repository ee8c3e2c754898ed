// vxp_mesh4.cuh — per-chunk meshing pipeline, generation 4 (SPEC §2-§5).
//
// One CTA (256 threads, ~70 KiB of shared memory, three CTAs per SM) owns one chunk:
//
//   tile      64 KiB, four 16 KiB slabs landed by bulk TMA copies (or written by the fused
//             terrain generator); converted slab by slab as the copies complete
//     -> occ  occupancy bit volume (bit x of word (y,z)) plus a one-voxel halo
//     -> eq   id-equality bit volumes EX/EY/EZ, greedy only, computed from the tile for the
//             rows that carry a visible face and held in registers until the tile is dead
//     -> culled: six face words per row straight from the occupancy, packed block scan, list of
//        non-empty words shared evenly between the threads
//     -> greedy: six face-mask sets in plane orientation (X planes through warp bit transposes),
//        bit-parallel row sweep, one thread per plane, that only logs which runs survive each row;
//        runs are staged and their heights resolved in parallel afterwards
//     -> one 32-byte vertex store per quad, coalesced 8-byte index stores.
//
// After the equality pass the tile is dead and its 64 KiB are re-used for eq / mask /
// staging, which is what lets three CTAs share an SM.
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
namespace v4 {

constexpr int kPlaneWords = 32 * kPad;   // 1056 words per padded 32x32 word plane set
constexpr int kSlabBytes = 16384;        // 8 z-layers
constexpr int kSlabs = 4;

// ---- optional per-phase cycle accounting (profiling build only: -DVXP_PROFILE, libvxp_prof.so)
#ifdef VXP_PROFILE
constexpr int kProfSlots = 16;
constexpr int kProfCtas = 8192;
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

// ---- shared memory
template <bool kGreedy>
struct Post;   // what lives in the tile's 64 KiB once the tile is dead
template <>
struct Post<false> {
    static constexpr int kWords = 6 * 32 * 32;                             // face words of a chunk
    static constexpr int kCap = (kTileBytes - kWords * 8) / 4;             // 4096 staged quad records
    uint2 words[kWords];         // non-empty face words: .x = f<<10 | y<<5 | z | (ordinal of the word's first quad) << 13, .y = bits over x
    uint32_t staging[kCap];
};
template <>
struct Post<true> {
    static constexpr int kCap = (kTileBytes - (192 * kPad + 5 * kPlaneWords) * 4) / 4;   // 4768
    // "H" arrays (same id as the previous cell of the row) and the masks are needed until the quads are
    // staged; the "V" arrays (same id as the cell of the previous row) only by the sweep, so they sit next
    // to the staging area and join it afterwards (kArena words from ey on).
    uint32_t ex[kPlaneWords];    // ex[z*33+y] bit x = (id(x,y,z) == id(x-1,y,z))   (bit 0 unused)
    uint32_t ht[kPlaneWords];    // X planes: ht[x*33+z] bit y = ey transposed
    uint32_t mask[192 * kPad];
    uint32_t ey[kPlaneWords];    // ey[z*33+y] bit x = (id(x,y,z) == id(x,y-1,z))   (row y=0 unused)
    uint32_t ez[kPlaneWords];    // ez[z*33+y] bit x = (id(x,y,z) == id(x,y,z-1))   (layer z=0 unused)
    uint32_t vt[kPlaneWords];    // X planes: vt[x*33+z] bit y = ez transposed
    uint32_t staging[kCap];
    static constexpr int kArena = 3 * kPlaneWords + kCap;
};

template <bool kGreedy>
struct __align__(128) Smem {
    union {
        uint16_t vox[VXP_CHUNK_VOXELS];   // [z][y][x]
        Post<kGreedy> post;
    } tile;
    uint32_t occ[34 * 34];      // occ[(z+1)*34 + (y+1)], bit x; rows/layers -1 and 32 are the halo
    uint32_t xh[2][32];         // xh[0][z] bit y: voxel (-1,y,z) solid;  xh[1][z] bit y: voxel (32,y,z)
    uint32_t actz[32];          // actz[z] bit y: row (y,z) carries at least one visible face
    uint32_t scratch[kWarps + 1];
    uint32_t seg[192];          // greedy: first log entry of each plane
    uint32_t rows[192];         // greedy: visited rows of each plane
    uint32_t first_run[192];    // greedy: ordinal of the first run of each plane
    uint32_t cursor;            // staging fill level
    uint32_t counted;           // thread 0 only: this CTA has checked in at the launch counter (reserve_quads)
    uint32_t ones;              // 0xFFFFFFFF: stand-in equality row for single-id chunks
    unsigned long long first;
    unsigned long long mbar[kSlabs];
};
static_assert(sizeof(Post<false>) <= kTileBytes && sizeof(Post<true>) <= kTileBytes, "post-tile layout overflows the tile");
static_assert(3 * (sizeof(Smem<true>) + 1024) <= 227 * 1024 + 1024, "three CTAs per SM need <= ~75 KiB each");

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
    uint32_t w[kIdPlanes][kSlabs];   // w[b][s]: byte k = bit b of the 8 voxels of this thread's k-th piece of slab s
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

// ------------------------------------------------------------------ tile -> occupancy
// Convert slab s (layers 8s..8s+7, 16 KiB at `slab`) into sm.occ; accumulates the per-thread flag words.
// kPlanes: also pull the id planes out.
template <bool kPlanes, class SM>
__device__ __forceinline__ void convert_rows(SM& sm, const uint4* slab, int s, uint32_t& v_or, uint32_t& v_and, uint32_t& o_and,
                                             IdPlanes& ip) {
    const int tid = threadIdx.x;
    unsigned char* occ_bytes = reinterpret_cast<unsigned char*>(sm.occ);
    uint32_t p0 = 0, p1 = 0, p2 = 0;
    uint4 q[4];
    uint32_t slab_or = 0, slab_and = kFull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q[k] = slab[tid + k * kThreads];                    // uint4 index inside the slab: row*4 + piece; a warp reads 512 contiguous bytes
        slab_or |= q[k].x | q[k].y | q[k].z | q[k].w;
        slab_and &= q[k].x & q[k].y & q[k].z & q[k].w;
    }
    v_or |= slab_or;
    v_and &= slab_and;
    // The warp's rows of this slab: 8 values of y in 4 layers.  In a terrain world most of these blocks hold one
    // value only - air above the surface (38 % of the chunks are nothing else), one kind of rock below it - and
    // then there is nothing to gather bit by bit: the occupancy bytes and the id planes are all ones or all zeros.
    const uint32_t id = slab_or & 0xFFFFu;
    const uint32_t first = __shfl_sync(kFull, slab_or, 0);              // (every lane takes part: not inside the &&)
    const bool same = slab_or == slab_and && id == (slab_or >> 16) && slab_or == first;
    if (__all_sync(kFull, same)) {
        const uint32_t fill = id ? 0xFFu : 0u;
        o_and &= fill;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + k * kThreads;
            const int row = s * 256 + (idx >> 2), p = idx & 3;  // row = z*32 + y
            occ_bytes[(((row >> 5) + 1) * 34 + (row & 31) + 1) * 4 + p] = (unsigned char)fill;
        }
        if constexpr (kPlanes) {
            p0 = (id & 1u) ? kFull : 0u;
            p1 = (id & 2u) ? kFull : 0u;
            p2 = (id & 4u) ? kFull : 0u;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + k * kThreads;
            const int row = s * 256 + (idx >> 2), p = idx & 3;
            const uint32_t bits = nz8(q[k]);
            o_and &= bits;
            occ_bytes[(((row >> 5) + 1) * 34 + (row & 31) + 1) * 4 + p] = (unsigned char)bits;
        }
        if constexpr (kPlanes) {
            id_plane_bytes<0>(q[0], p0, p1, p2);
            id_plane_bytes<1>(q[1], p0, p1, p2);
            id_plane_bytes<2>(q[2], p0, p1, p2);
            id_plane_bytes<3>(q[3], p0, p1, p2);
        }
    }
    if constexpr (kPlanes) {
#pragma unroll
        for (int t = 0; t < kSlabs; ++t)                    // s is an unrolled loop counter of the caller: ip stays in registers
            if (t == s) { ip.w[0][t] = p0; ip.w[1][t] = p1; ip.w[2][t] = p2; }
    }
}
template <bool kGreedy>
__device__ __forceinline__ void convert_slab(Smem<kGreedy>& sm, int s, uint32_t& v_or, uint32_t& v_and, uint32_t& o_and, IdPlanes& ip) {
    convert_rows<kGreedy>(sm, reinterpret_cast<const uint4*>(sm.tile.vox) + s * (kSlabBytes / 16), s, v_or, v_and, o_and, ip);
}
// Store the id planes into the post area (the tile is dead from here on: the caller has passed the barrier
// that follows the conversion).  plane b, word (z*kPad + y), bit x.
__device__ __forceinline__ void store_id_planes(Post<true>& post, const IdPlanes& ip) {
    const int tid = threadIdx.x;
    unsigned char* base = reinterpret_cast<unsigned char*>(post.staging);
#pragma unroll
    for (int s = 0; s < kSlabs; ++s) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + k * kThreads;
            const int row = s * 256 + (idx >> 2), p = idx & 3;
            const int word = (row >> 5) * kPad + (row & 31);
#pragma unroll
            for (int b = 0; b < kIdPlanes; ++b) base[(b * kPlaneWords + word) * 4 + p] = (unsigned char)(ip.w[b][s] >> (8 * k));
        }
    }
}
// Block-wide flags; contains the barrier that publishes the occupancy interior.  Every flag is a predicate each
// thread can evaluate on its own words, so four barrier reductions do (no shared-memory round).
template <class SM>
__device__ __forceinline__ ChunkFlags reduce_flags(SM& sm, uint32_t v_or, uint32_t v_and, uint32_t o_and) {
    const uint32_t ref = reinterpret_cast<const uint32_t*>(sm.tile.vox)[0];   // voxels 0 and 1 of the chunk
    ChunkFlags f;
    f.any_solid = __syncthreads_or(v_or != 0u);
    f.all_solid = __syncthreads_and((o_and & 0xFFu) == 0xFFu);
    // every u16 of the chunk is the same value: every word equals the first one, whose halves are equal
    f.single_id = __syncthreads_and(v_or == ref && v_and == ref && (ref & 0xFFFFu) == (ref >> 16));
    f.small_ids = __syncthreads_and(((v_or | (v_or >> 16)) & 0xFFFFu) < (1u << kIdPlanes));
    return f;
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
#define VXP_PLANE_POLLS 8
#endif
constexpr int kPlanePolls = VXP_PLANE_POLLS;   // bounded re-reads of a boundary plane that is not there yet (never a wait-for);
                                               // measured on config 2: 1 -> 109.4 us, 3 -> 104.9, 6 -> 104.1, 12 -> 103.5
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

// ------------------------------------------------------------------ tile -> EX / EY / EZ (into registers)
// res[k] of a thread with piece p = lane&3 holds, for row (idx>>2) of item k: p=0 the EX word,
// p=1 EY, p=2 EZ.  Rows without a visible face are skipped (their words are never consulted).
template <bool kGreedy>
__device__ __forceinline__ void equality_to_regs(const Smem<kGreedy>& sm, uint32_t (&res)[16]) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int p = lane & 3;
    const uint4* tile4 = reinterpret_cast<const uint4*>(sm.tile.vox);
    const uint32_t selA = (p & 1) ? 0x3715u : 0x6240u;   // 4x4 byte transpose across the 4 lanes of a row
    const uint32_t selB = (p & 2) ? 0x3276u : 0x5410u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int idx = tid + k * kThreads;
        const int row = idx >> 2, z = row >> 5, y = row & 31;
        res[k] = 0;
        // the warp's 8 rows are y0..y0+7 of one layer
        if (((sm.actz[z] >> (y & 24)) & 0xFFu) == 0) continue;   // warp-uniform
        const uint4 q = tile4[idx];
        const uint32_t leftw = __shfl_up_sync(kFull, q.w, 1);
        const uint4 px = make_uint4(prev16(leftw, q.x), prev16(q.x, q.y), prev16(q.y, q.z), prev16(q.z, q.w));
        uint32_t w = (~nz8(xor4(q, px))) & 0xFFu;
        if (y > 0) w |= ((~nz8(xor4(q, tile4[idx - 4]))) & 0xFFu) << 8;
        if (z > 0) w |= ((~nz8(xor4(q, tile4[idx - 128]))) & 0xFFu) << 16;
        uint32_t t = __shfl_xor_sync(kFull, w, 1);
        w = __byte_perm(w, t, selA);
        t = __shfl_xor_sync(kFull, w, 2);
        res[k] = __byte_perm(w, t, selB);
    }
}
__device__ __forceinline__ void equality_store(Post<true>& post, const uint32_t (&res)[16]) {
    const int tid = threadIdx.x, p = tid & 3;
    if (p == 3) return;
    uint32_t* const dst = p == 0 ? post.ex : p == 1 ? post.ey : post.ez;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int row = (tid + k * kThreads) >> 2;
        dst[(row >> 5) * kPad + (row & 31)] = res[k];
    }
}

// ------------------------------------------------------------------ occupancy -> face masks
// Returns the OR of every mask word this thread produced.
// with_planes (greedy, ids that fit kIdPlanes bits): the equality rows EX / EY / EZ of each row are derived here
// from the id bit planes (store_id_planes) instead of having been computed from the tile.
template <bool kGreedy>
__device__ __forceinline__ uint32_t build_masks(Smem<kGreedy>& sm, bool with_eq, bool with_planes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* mask = sm.tile.post.mask;
    uint32_t acc = 0;
    // X planes have u=y, v=z, slice x: their rows are 32x32 bit transposes (lane=y, bit=x) -> (lane=x, bit=y) of what
    // a warp holds for one z.  The transposes of the warp's four layers run together (independent shuffles).
    uint32_t tx[8], te[8];          // tx[k] / tx[4+k]: -X / +X visible faces of layer k;  te[k] / te[4+k]: EY / EZ rows
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp, y = lane;
        const uint32_t* o = sm.occ + (z + 1) * 34 + (y + 1);
        const uint32_t c = o[0];
        const uint32_t my = c & ~o[-1], py = c & ~o[+1], mz = c & ~o[-34], pz = c & ~o[+34];
        mask[(2 * 32 + y) * kPad + z] = my;    // -Y : slice y, row z
        mask[(3 * 32 + y) * kPad + z] = py;    // +Y
        mask[(4 * 32 + z) * kPad + y] = mz;    // -Z : slice z, row y
        mask[(5 * 32 + z) * kPad + y] = pz;    // +Z
        const uint32_t lo = (sm.xh[0][z] >> y) & 1u, hi = (sm.xh[1][z] >> y) & 1u;
        tx[k] = c & ~((c << 1) | lo);                   // -X visible, bit x
        tx[4 + k] = c & ~((c >> 1) | (hi << 31));       // +X visible, bit x
        acc |= my | py | mz | pz | tx[k] | tx[4 + k];
        te[k] = te[4 + k] = 0;
        if constexpr (kGreedy) {
            if (with_eq) {
                Post<true>& P = sm.tile.post;
                if (with_planes) {
                    const uint32_t* B = P.staging + z * kPad + y;
                    uint32_t dx = 0, dy = 0, dz = 0;
#pragma unroll
                    for (int b = 0; b < kIdPlanes; ++b) {
                        const uint32_t w = B[b * kPlaneWords];
                        const uint32_t wy = y > 0 ? B[b * kPlaneWords - 1] : w;       // row y = 0 / layer z = 0 are never consulted
                        const uint32_t wz = z > 0 ? B[b * kPlaneWords - kPad] : w;
                        dx |= w ^ (w << 1);
                        dy |= w ^ wy;
                        dz |= w ^ wz;
                    }
                    te[k] = ~dy;
                    te[4 + k] = ~dz;
                    P.ex[z * kPad + y] = ~dx;
                    P.ey[z * kPad + y] = te[k];
                    P.ez[z * kPad + y] = te[4 + k];
                } else {
                    te[k] = P.ey[z * kPad + y];
                    te[4 + k] = P.ez[z * kPad + y];
                }
            }
        }
    }
    uint32_t xany = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) xany |= tx[k];
    const bool anyx = __any_sync(kFull, xany != 0);     // no x face in the warp's four layers: their X mask rows are 0
    if (anyx) {
        transpose32xN(tx, lane);
        if constexpr (kGreedy) {
            if (with_eq) transpose32xN(te, lane);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp;
        mask[(0 * 32 + lane) * kPad + z] = tx[k];
        mask[(1 * 32 + lane) * kPad + z] = tx[4 + k];
        if constexpr (kGreedy) {
            if (with_eq && anyx) {
                sm.tile.post.ht[lane * kPad + z] = te[k];
                sm.tile.post.vt[lane * kPad + z] = te[4 + k];
            }
        }
    }
    return acc;
}

// ------------------------------------------------------------------ greedy sweep (one thread = one plane)
// Bit-parallel evaluation of the SPEC §4 greedy rule.  After row v-1 every set cell of that row
// belongs to exactly one active run (S = start bits, E = end bits, cover C = mask row v-1).
// Row v: a run survives iff all its cells are set in row v with the id of the cell above
// (G = M & V); survivors are found with one carry chain upwards (segmented AND-reduce) and
// spread back over their run with one carry chain on the bit-reversed words.  Runs that do not
// survive are finished quads; the cells of row v not covered by survivors start new runs,
// split where the id changes (H).  r0..r4 hold the start row of the run that starts at each
// column, bit-sliced.
struct Sweep {
    const uint32_t* m;   // this plane's 32 mask rows
    const uint32_t* hb;  // equality rows along u  (bit u: cell u has the id of cell u-1)
    const uint32_t* vb;  // equality rows along v  (bit u: cell (u,v) has the id of (u,v-1))
    int hs, vs;          // row strides of hb / vb
    uint32_t C, S, E, r0, r1, r2, r3, r4;
    uint32_t base;       // f<<15 | s<<10
};

template <bool kGreedy>
__device__ __forceinline__ void sweep_init(Sweep& st, Smem<kGreedy>& sm, bool with_eq) {
    const int tid = threadIdx.x;
    const int f = tid >> 5, s = tid & 31;
    st.C = st.S = st.E = st.r0 = st.r1 = st.r2 = st.r3 = st.r4 = 0;
    st.base = (uint32_t)f << 15 | (uint32_t)s << 10;
    st.m = sm.tile.post.mask + (tid < 192 ? tid : 0) * kPad;
    st.hb = st.vb = &sm.ones;
    st.hs = st.vs = 0;
    if constexpr (kGreedy) {
        if (with_eq) {
            Post<true>& P = sm.tile.post;
            if (f < 2) { st.hb = P.ht + s * kPad; st.hs = 1; st.vb = P.vt + s * kPad; st.vs = 1; }        // X planes: u=y, v=z
            else if (f < 4) { st.hb = P.ex + s; st.hs = kPad; st.vb = P.ez + s; st.vs = kPad; }           // Y planes: u=x, v=z
            else { st.hb = P.ex + s * kPad; st.hs = 1; st.vb = P.ey + s * kPad; st.vs = 1; }              // Z planes: u=x, v=y
        }
    }
}

// Stage the runs (dS, dE: start / end bits, pairwise in order) that ended before row v.
__device__ __forceinline__ void sweep_emit(const Sweep& st, uint32_t dS, uint32_t dE, int v, uint32_t* staging,
                                           uint32_t cap, uint32_t* cursor) {
    if (dS == 0) return;
    uint32_t pos = atomicAdd(cursor, (uint32_t)__popc(dS));   // one reservation per row step, not per quad
#ifdef VXP_EXP_NO_EMIT
    if (pos != 0xFFFFFFFFu) return;   // timing experiment: the sweep without its per-quad loop (wrong output)
#endif
    const uint32_t tail = st.base | (uint32_t)(v - 1) << 23;
    do {
        const int u = __ffs(dS) - 1, e = __ffs(dE) - 1;
        dS &= dS - 1;
        dE &= dE - 1;
        const uint32_t v0 = ((st.r0 >> u) & 1u) | ((st.r1 >> u) & 1u) << 1 | ((st.r2 >> u) & 1u) << 2 |
                            ((st.r3 >> u) & 1u) << 3 | ((st.r4 >> u) & 1u) << 4;
        // u | v0<<5 | s<<10 | f<<15 | (w-1)<<18 | (h-1)<<23 with h-1 = v-1-v0
        const uint32_t rec = tail + (uint32_t)u + (v0 << 5) + ((uint32_t)(e - u) << 18) - (v0 << 23);
        if (pos < cap) staging[pos] = rec;
        ++pos;
    } while (dS);
}


// The row update proper.  Returns the dead runs (start / end bits) through dS / dE.
__device__ __forceinline__ void sweep_row(Sweep& st, int v, uint32_t Mv, uint32_t Hv, uint32_t Vv, uint32_t& dS,
                                          uint32_t& dE) {
    uint32_t Cs = 0;
    if (st.C) {
        const uint32_t X = st.C & Mv & Vv;
        const uint32_t T = (X & ~st.E) + st.S;
        const uint32_t Es = T & st.E & X;                       // end bits of the surviving runs
        const uint32_t P = __brev(st.C) & ~__brev(st.S);
        const uint32_t Q = P + __brev(Es);
        Cs = __brev(Q ^ P);                                     // their cover
    }
    dS = st.S & ~Cs;
    dE = st.E & ~Cs;
    const uint32_t R = Mv & ~Cs;
    const uint32_t NS = R & ~((R << 1) & Hv);
    const uint32_t NE = R & ~((R >> 1) & (Hv >> 1));
    st.S = (st.S & Cs) | NS;
    st.E = (st.E & Cs) | NE;
    st.C = Mv;
    st.r0 = (st.r0 & ~NS) | ((v & 1) ? NS : 0u);
    st.r1 = (st.r1 & ~NS) | ((v & 2) ? NS : 0u);
    st.r2 = (st.r2 & ~NS) | ((v & 4) ? NS : 0u);
    st.r3 = (st.r3 & ~NS) | ((v & 8) ? NS : 0u);
    st.r4 = (st.r4 & ~NS) | ((v & 16) ? NS : 0u);
}

// One row step (v in [0,32]); v == 32 flushes the runs still active after the last row.
// (Used by the stepped heavy path; the normal path is sweep_plane below.)
__device__ __forceinline__ void sweep_step(Sweep& st, int v, uint32_t* staging, uint32_t cap, uint32_t* cursor) {
    if (v == 32) {
        sweep_emit(st, st.S, st.E, 32, staging, cap, cursor);
        st.S = st.E = st.C = 0;
        return;
    }
    const uint32_t Mv = st.m[v];
    if ((Mv | st.C) == 0) return;
    uint32_t dS, dE;
    const uint32_t r0 = st.r0, r1 = st.r1, r2 = st.r2, r3 = st.r3, r4 = st.r4;   // start rows before the update
    Sweep old = st;
    sweep_row(st, v, Mv, st.hb[v * st.hs], st.vb[v * st.vs], dS, dE);
    old.r0 = r0; old.r1 = r1; old.r2 = r2; old.r3 = r3; old.r4 = r4;
    sweep_emit(old, dS, dE, v, staging, cap, cursor);
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

// Non-empty mask rows of the plane, as a bit mask.  The sweep visits rows rm | rm<<1 (a row matters iff it
// or the row before it is non-empty; row 32 never needs a visit: what is alive after row 31 just ends
// there) and logs the non-empty ones.
__device__ __forceinline__ uint32_t sweep_rows(const Sweep& st) {
    uint32_t rm = 0;
#pragma unroll
    for (int v = 0; v < 32; ++v) rm |= (uint32_t)(st.m[v] != 0) << v;
    return rm;
}

// Returns the number of runs the plane starts (= its quads).
__device__ __forceinline__ uint32_t sweep_log(Sweep& st, uint32_t rm, uint32_t plane, uint2* entry) {
    uint32_t births = 0;
    uint32_t todo = rm | (rm << 1);
    // the warp's trip count is known up front (no vote on the loop-carried path); the body is branch-free for
    // lanes with rows left (predicated), so next row's loads and this row's update overlap
    const int trips = (int)__reduce_max_sync(kFull, (uint32_t)__popc(todo));
    int v = todo ? __ffs(todo) - 1 : 0;
    bool valid = todo != 0;
    todo &= todo - 1;
    uint32_t Mv = st.m[v], Hv = st.hb[v * st.hs], Vv = st.vb[v * st.vs];
#pragma unroll 2
    for (int it = 0; it < trips; ++it) {
        const bool validn = todo != 0;
        const int vn = validn ? __ffs(todo) - 1 : 0;
        todo &= todo - 1;
        const uint32_t Mn = st.m[vn], Hn = st.hb[vn * st.hs], Vn = st.vb[vn * st.vs];
        // survivors of the runs alive after the previous row (C == 0: none, and then X == Es == 0 and Cs == 0)
        const uint32_t X = st.C & Mv & Vv;
        const uint32_t T = (X & ~st.E) + st.S;
        const uint32_t Es = T & st.E & X;                               // end bits of the surviving runs
        const uint32_t P = __brev(st.C) & ~__brev(st.S);
        const uint32_t Cs = __brev((P + __brev(Es)) ^ P);               // their cover
        const uint32_t R = Mv & ~Cs;
        const uint32_t NS = R & ~((R << 1) & Hv);
        const uint32_t NE = R & ~((R >> 1) & (Hv >> 1));
        if (valid) {
            st.S = (st.S & Cs) | NS;
            st.E = (st.E & Cs) | NE;
            st.C = Mv;
            if (Mv) *entry++ = make_uint2(Cs, plane << 5 | (uint32_t)v | births << 13);
            births += __popc(NS);
        }
        v = vn; valid = validn; Mv = Mn; Hv = Hn; Vv = Vn;
    }
    return births;
}

// H row (bit u: cell u has the id of cell u-1) of plane `plane`, row v.
template <bool kGreedy>
__device__ __forceinline__ uint32_t plane_h_row(const Smem<kGreedy>& sm, bool with_eq, uint32_t plane, uint32_t v) {
    if constexpr (kGreedy) {
        if (with_eq) {
            const Post<true>& P = sm.tile.post;
            const uint32_t f = plane >> 5, s = plane & 31u;
            return f < 2 ? P.ht[s * kPad + v] : f < 4 ? P.ex[v * kPad + s] : P.ex[s * kPad + v];
        }
    }
    return kFull;
}

// One record (u | v<<5 | s<<10 | f<<15 | (w-1)<<18, height 1) per run that starts at a logged row, at its
// ordinal: the plane's first ordinal + the runs of the plane's earlier rows + its rank in the row.
template <bool kGreedy>
__device__ __forceinline__ void stage_births(const Smem<kGreedy>& sm, bool with_eq, const SweepLog& log, uint32_t nL,
                                             uint32_t* rec) {
    const uint32_t* mask = sm.tile.post.mask;
    for (uint32_t i = threadIdx.x; i < nL; i += kThreads) {
        const uint2 en = log.entry[i];
        const uint32_t plane = (en.y >> 5) & 255u, v = en.y & 31u;
        const uint32_t R = mask[plane * kPad + v] & ~en.x;           // cells of this row not covered from above
        if (R == 0) continue;
        const uint32_t Hv = plane_h_row(sm, with_eq, plane, v);
        uint32_t NS = R & ~((R << 1) & Hv), NE = R & ~((R >> 1) & (Hv >> 1));
        uint32_t pos = log.first[plane] + (en.y >> 13);
        const uint32_t base = (en.y & 0x1FFFu) << 5;                  // f<<15 | s<<10 | v<<5
        do {
            const uint32_t u = __ffs(NS) - 1, e = __ffs(NE) - 1;
            NS &= NS - 1;
            NE &= NE - 1;
            rec[pos++] = base | u | (e - u) << 18;
        } while (NS);
    }
}
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
template <class IdFn>
__device__ __forceinline__ void culled_emit(Smem<false>& sm, uint32_t chunk, uint32_t packed, const MeshOut& out, IdFn id_of,
                                            Prof& prof) {
    const int tid = threadIdx.x;
    constexpr uint32_t kCap = Post<false>::kCap;
    uint32_t* const staging = sm.tile.post.staging;
    auto nothing = [] {};
    auto staged = [staging](uint32_t i) { return staging[i]; };
    {
        Post<false>& P = sm.tile.post;
        const int y = tid & 31, w = tid >> 5;
        uint32_t tot;
        const uint32_t pre = block_excl_scan(packed, sm.scratch, &tot);
        const uint32_t Q = tot & 0x1FFFFu, nwords = tot >> 17;
        if (Q == 0) {
            if (tid == 0) out.meta[chunk] = vxp_chunk_meta{0u, 0u};
            prof.count(10);
            return;
        }
        prof.count(11);
        unsigned long long fq = 0;
        if (tid == 0) fq = reserve_quads(sm, out, Q);   // in flight while the records are staged
        {
            uint32_t ord = pre & 0x1FFFFu, slot = pre >> 17;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t wd[6];
                const int z = 4 * w + k;
                if (!row_words_yz(sm, y, z, wd)) continue;
                row_words_x(sm, y, z, wd);
#pragma unroll
                for (int f = 0; f < 6; ++f) {
                    if (wd[f]) P.words[slot++] = make_uint2((uint32_t)(f << 10 | y << 5 | z) | ord << 13, wd[f]);
                    ord += __popc(wd[f]);
                }
            }
        }
        __syncthreads();
        prof.mark(5);
        for (uint32_t base = 0; base < Q; base += kCap) {
            const uint32_t end = min(Q, base + kCap);
            for (uint32_t i = tid; i < nwords; i += kThreads) {     // neighbours in the list are words of similar density
                const uint2 d = P.words[i];
                uint32_t bits = d.y, o = d.x >> 13;
                if (o + __popc(bits) <= base || o >= end) continue;
                const uint32_t f = (d.x >> 10) & 7u, y = (d.x >> 5) & 31u, zz = d.x & 31u;
                // u | v<<5 | s<<10 | f<<15:  X planes (s,u,v) = (x,y,z);  Y planes (y,x,z);  Z planes (z,x,y)
                const uint32_t rb = f << 15 | (f < 2 ? (y | zz << 5) : f < 4 ? (zz << 5 | y << 10) : (y << 5 | zz << 10));
                const uint32_t sh = f < 2 ? 10u : 0u;
                while (bits) {
                    const uint32_t x = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (o >= base && o < end) staging[o - base] = rb | x << sh;
                    ++o;
                }
            }
            __syncthreads();
            prof.mark(6);
            bool fits;
            if (base == 0) fits = flush_staging(sm, staged, end, 0, out, id_of, [&] { if (tid == 0) publish_range(sm, chunk, Q, fq, out); }, true);
            else fits = flush_staging(sm, staged, end - base, base, out, id_of, nothing, false);
            if (!fits) return;
            __syncthreads();
            prof.mark(7);
        }
        return;
    }
}

// ------------------------------------------------------------------ masks -> quads -> buffers
template <bool kGreedy, class IdFn>
__device__ __forceinline__ void emit_chunk(Smem<kGreedy>& sm, uint32_t chunk, bool with_eq, const MeshOut& out, IdFn id_of,
                                           Prof& prof) {
    const int tid = threadIdx.x;
    constexpr uint32_t kCap = Post<kGreedy>::kCap;
    uint32_t* const staging = sm.tile.post.staging;
    auto nothing = [] {};
    auto staged = [staging](uint32_t i) { return staging[i]; };
    if constexpr (!kGreedy) {
        (void)with_eq;
        culled_emit(sm, chunk, culled_count<false>(sm) + culled_count<true>(sm), out, id_of, prof);
        return;
    } else {
        // ---- greedy: sweep_log, stage_births, run_height (see above).  The log sits at the end of the
        // staging area; the records go below it if they fit (the V arrays then stay intact for the stepped
        // fallback) and otherwise, the sweep being over, from ey up.
        Post<true>& P = sm.tile.post;
        if (tid == 0) sm.cursor = 0;
        Sweep st;
        sweep_init(st, sm, with_eq);
        const uint32_t rm = tid < 192 ? sweep_rows(st) : 0u;
        uint32_t nL;
        const uint32_t seg = block_excl_scan((uint32_t)__popc(rm), sm.scratch, &nL);   // barriers: cursor = 0 published
        if (tid < 192) { sm.seg[tid] = seg; sm.rows[tid] = rm; }
        const bool log_fits = 2 * nL <= kCap;
        SweepLog log;
        log.entry = reinterpret_cast<uint2*>(staging + kCap - 2 * nL);
        log.seg = sm.seg;
        log.rows = sm.rows;
        log.first = sm.first_run;
        uint32_t Q = 0;
        unsigned long long f = 0;
        if (log_fits) {
            uint32_t births = 0;
#ifdef VXP_PROFILE
            const long long t_sw = clock64();
#endif
            if (tid < 192) births = sweep_log(st, rm, (uint32_t)tid, log.entry + seg);   // warps 0..5, whole warps
#ifdef VXP_PROFILE
            if ((tid & 31) == 0 && blockIdx.x < kProfCtas) g_prof_warp[blockIdx.x][tid >> 5] = (unsigned int)(clock64() - t_sw);
#endif
            const uint32_t first_run = block_excl_scan(births, sm.scratch, &Q);   // barriers: log published
            if (tid < 192) sm.first_run[tid] = first_run;
            if (tid == 0) f = reserve_quads(sm, out, Q);       // in flight while the records are staged
            prof.mark(5);
            const uint32_t cap_small = kCap - 2 * nL, cap_arena = Post<true>::kArena - 2 * nL;
            if (Q <= cap_arena) {                         // block-uniform
                uint32_t* const rec = Q <= cap_small ? staging : P.ey;
                __syncthreads();                          // first_run published; V arrays dead
                stage_births(sm, with_eq, log, nL, rec);
                __syncthreads();
                prof.mark(6);
                flush_staging(sm, [rec, &log](uint32_t i) { const uint32_t r = rec[i]; return r | (run_height(log, r) - 1u) << 23; },
                              Q, 0, out, id_of, [&] { if (tid == 0) publish_range(sm, chunk, Q, f, out); }, true);
                prof.mark(7);
                return;
            }
        } else {
            // the log alone does not fit: count with a cursor-only stepped sweep
            if (tid < 192) {
#pragma unroll 1
                for (int v = 0; v <= 32; ++v) sweep_step(st, v, staging, 0u, &sm.cursor);
            }
            __syncthreads();
            Q = sm.cursor;
            if (tid == 0) f = reserve_quads(sm, out, Q);
        }
        auto publish = [&] { if (tid == 0) publish_range(sm, chunk, Q, f, out); };
        prof.count(14);
        // heavy path (the log or the records do not fit): sweep again row by row, half of the planes per
        // step (<= 96*32 quads), staging through a cursor and flushing whenever the next step might not fit
        publish();
        sweep_init(st, sm, with_eq);
        uint32_t flushed = 0;
        if (tid == 0) sm.cursor = 0;
#pragma unroll 1
        for (int v = 0; v <= 32; ++v) {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                __syncthreads();                          // cursor reset / previous flush done
                if (tid < 192 && (tid >= 96) == (half == 1)) sweep_step(st, v, staging, kCap, &sm.cursor);
                __syncthreads();
                const uint32_t cnt = sm.cursor;
                if (cnt + 96 * 32 > kCap || (v == 32 && half == 1)) {   // the next half step might not fit
                    if (!flush_staging(sm, staged, cnt, flushed, out, id_of, nothing, false)) return;
                    flushed += cnt;
                    __syncthreads();
                    if (tid == 0) sm.cursor = 0;
                }
            }
        }
        prof.mark(7);
    }
}


// ------------------------------------------------------------------ shared tail of both kernels
// Preconditions: occupancy interior + halo complete and published by a barrier; tile intact.
template <bool kGreedy, class IdFn>
__device__ __forceinline__ void mesh_tail(Smem<kGreedy>& sm, uint32_t chunk, ChunkFlags fl, bool with_planes, const MeshOut& out,
                                          IdFn id_of, Prof& prof) {
    const int tid = threadIdx.x;
    bool with_eq = false;
    if constexpr (kGreedy) {
        with_eq = !fl.single_id;
        if (with_eq && !with_planes) {
            build_active(sm);
            __syncthreads();
            uint32_t res[16];
            equality_to_regs(sm, res);
            __syncthreads();                     // the tile is dead from here on
            equality_store(sm.tile.post, res);
        }
    }
    if constexpr (kGreedy) {
        __syncthreads();
        prof.mark(3);
        const bool any = __syncthreads_or(build_masks(sm, with_eq, with_planes) != 0);
        prof.mark(4);
        if (!any) {
            if (tid == 0) out.meta[chunk] = vxp_chunk_meta{0u, 0u};
            prof.count(10);
            return;
        }
        prof.count(11);
    }
    emit_chunk<kGreedy>(sm, chunk, with_eq, out, id_of, prof);
}

// ====================================================================== stored-chunk meshing
// Slabs of the tile a CTA asks for before the first one has landed (measured on B200, config 2:
// 1 -> 120 us, 2 -> 111 us, 4 -> 107 us per 4096 chunks: deeper is better, so the whole tile it is).
#ifndef VXP_SLABS_IN_FLIGHT
#define VXP_SLABS_IN_FLIGHT 4
#endif
constexpr int kSlabsInFlight = VXP_SLABS_IN_FLIGHT;
template <bool kGreedy>
__device__ __forceinline__ void issue_slab(Smem<kGreedy>& sm, const uint16_t* gvox, int s) {
    mbar_expect_tx(&sm.mbar[s], kSlabBytes);
    tma_load_1d(reinterpret_cast<unsigned char*>(sm.tile.vox) + s * kSlabBytes,
                reinterpret_cast<const unsigned char*>(gvox) + s * kSlabBytes, kSlabBytes, &sm.mbar[s]);
}

// One chunk, start to finish; its meta entry is out.meta[mi].  The four tile barriers have been initialised by the caller and complete once
// per chunk: `parity` is the phase to wait for (0 for the first chunk of a CTA, then alternating).
// Every return is block-uniform.
template <bool kGreedy>
__device__ __forceinline__ void mesh_chunk(Smem<kGreedy>& sm, const uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6,
                                           uint32_t n, uint32_t chunk, uint32_t mi, uint32_t parity, const MeshOut& out, const Bnd& bnd) {
    const int tid = threadIdx.x;
    const uint16_t* gvox = voxels + (size_t)chunk * VXP_CHUNK_VOXELS;
    Prof prof;
    prof.start(tid == 0);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kSlabsInFlight; ++s) issue_slab(sm, gvox, s);
    }
    // neighbour of face (warp index), fetched while the tile is in flight
    const int nb = (nbr6 && tid < 192) ? __ldg(nbr6 + 6 * (size_t)chunk + (tid >> 5)) : -1;
    HaloRows rows;
    issue_yz_halo(rows, voxels, nb);

    uint32_t v_or = 0, v_and = kFull, o_and = kFull;
    IdPlanes ip;
#pragma unroll (kGreedy ? kSlabs : 1)
    for (int s = 0; s < kSlabs; ++s) {
        mbar_wait(&sm.mbar[s], parity);
        if (tid == 0 && s + kSlabsInFlight < kSlabs) issue_slab(sm, gvox, s + kSlabsInFlight);
        convert_slab(sm, s, v_or, v_and, o_and, ip);
    }
    prof.mark(0);
    const ChunkFlags fl = reduce_flags(sm, v_or, v_and, o_and);   // barrier: occupancy interior published
    prof.mark(1);
#if defined(VXP_EXP_STOP) && VXP_EXP_STOP == 1
    if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};   // timing experiment: stream + convert only
    return;
#endif
    if (bnd.tags) publish_boundary(sm, bnd, chunk, fl);
    if (!fl.any_solid) {                        // all air: no faces
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(8);
        return;
    }
    if constexpr (!kGreedy) {
        if (!fl.all_solid) {
            // Culled, not a buried candidate: count the faces that do not depend on the x halo (+-Y, +-Z: two
            // thirds of the face words) while the x neighbours' boundary planes are on their way.
            XPoll xp;
            halo_begin(sm, nb, n, bnd, rows, xp);
            __syncthreads();                    // +-Y / +-Z halo rows published
            uint32_t packed = culled_count<false>(sm);
            halo_finish(sm, voxels, nb, n, bnd, xp, prof);
            __syncthreads();                    // x halo published
            prof.mark(2);
            packed += culled_count<true>(sm);
            culled_emit(sm, mi, packed, out, [gvox](int idx) -> uint32_t { return __ldg(gvox + idx); }, prof);
            return;
        }
    }
    // greedy, several ids that all fit the id planes: the planes replace the tile here (it is dead after the
    // conversion barrier), stored while the x neighbours' boundary planes are on their way
    bool with_planes = false;
    XPoll xp;
    const bool air_yz = halo_begin(sm, nb, n, bnd, rows, xp);
    if constexpr (kGreedy) {
        with_planes = !fl.single_id && fl.small_ids;
        if (with_planes) store_id_planes(sm.tile.post, ip);
    }
    const bool air_x = halo_finish(sm, voxels, nb, n, bnd, xp, prof);
    const bool halo_air = __syncthreads_or(air_yz || air_x);   // barrier: halo (and id planes) published
    prof.mark(2);
#if defined(VXP_EXP_STOP) && VXP_EXP_STOP == 2
    if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};   // timing experiment: ... + halo
    return;
#endif
    if (fl.all_solid && !halo_air) {            // buried: no faces
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(9);
        return;
    }
    mesh_tail<kGreedy>(sm, mi, fl, with_planes, out, [gvox](int idx) -> uint32_t { return __ldg(gvox + idx); }, prof);
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

// Meshes chunks [chunk0, hi) of a batch of n, one chunk per CTA (grid = hi - chunk0): the hardware block
// scheduler is the load balancer.  (Persistent CTAs drawing tickets were measured 7-12 % slower: a CTA
// that draws its next chunk early delays that chunk's boundary planes for the neighbours that poll them.)
template <bool kGreedy>
__global__ void __launch_bounds__(kThreads, 3)
mesh_kernel(const uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6, uint32_t n, uint32_t chunk0, uint32_t hi,
            MeshOut out, Bnd bnd, Chain chain) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<kGreedy>& sm = *reinterpret_cast<Smem<kGreedy>*>(smem_raw);
    const bool dealt = chain.runs != 0u && (blockIdx.x >> 4) < chain.runs;
    if (dealt && threadIdx.x == 0)              // one 64-bit modulo per CTA, not per thread
        sm.cursor = (uint32_t)(((unsigned long long)(blockIdx.x >> 4) * chain.run_stride) % chain.runs) * 16u + (blockIdx.x & 15u);
    const bool closer = chain.user_total != nullptr && blockIdx.x == gridDim.x - 1;
    if (threadIdx.x == 0) {
        if (closer) griddep_wait();             // the previous launch has completed (its slot is free for the next one)
        sm.ones = kFull;
        sm.counted = 0u;
#pragma unroll
        for (int s = 0; s < kSlabs; ++s) mbar_init(&sm.mbar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();                            // barrier init visible to all
    griddep_launch_dependents();
    uint32_t chunk = dealt ? sm.cursor : chunk0 + blockIdx.x;
    uint32_t mi = chunk;
    bool active = chunk < hi;
    if (chain.list != nullptr) {
        active = chain.list_count == nullptr || blockIdx.x < __ldg(chain.list_count);
        chunk = active ? __ldg(chain.list + blockIdx.x) : 0u;
        mi = blockIdx.x;
        active = active && chunk < n;
    }
#ifdef VXP_PROFILE
    if (threadIdx.x == 0 && blockIdx.x < kProfCtas) g_prof_time[blockIdx.x][0] = prof_now();
#endif
    if (active) mesh_chunk(sm, voxels, nbr6, n, chunk, mi, 0u, out, bnd);
#ifdef VXP_PROFILE
    if (threadIdx.x == 0 && blockIdx.x < kProfCtas) g_prof_time[blockIdx.x][1] = prof_now();
#endif
    if (threadIdx.x == 0 && chain.user_total != nullptr) {
        if (!sm.counted) red_add_u64(out.total, kCtaOne);          // a chunk without quads checks in here
        if (closer) {
            // sm.counted: this CTA's own reservation has returned, so the counter already includes it
            unsigned long long t;
            while (((t = ld_relaxed_u64(out.total)) >> kQuadBits) != gridDim.x) __nanosleep(64);
            *chain.user_total = t & kQuadMask;
            st_relaxed_u64(out.total, 0ull);    // nobody adds any more: every CTA has checked in
        }
    }
}

} // namespace v4
} // namespace vxp
