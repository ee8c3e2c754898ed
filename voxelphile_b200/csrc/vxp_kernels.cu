// vxp_kernels.cu — sm_100a kernels: terrain fill, pattern fill, stored-chunk meshing,
// fused terrain+mesh.  SPEC.md v1.  Host launchers at the bottom (C++ linkage, used by vxp_api.cu).
#include <climits>
#include <cstring>

#include "vxp_launch.h"
#include "vxp_mesh.cuh"

namespace vxp {

// ====================================================================== terrain fill (SPEC §6)
// One CTA per chunk: lattice gradients of the footprint (100 hashes), 1024 column heights -> shared
// memory, then 4096 16-byte stores.  Rows entirely above the surface of their z-layer, or entirely
// below its dirt, skip the per-voxel classification.
__global__ void __launch_bounds__(kThreads) terrain_fill_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                 uint64_t seed, uint16_t* __restrict__ out) {
    __shared__ TerrainLattice s_lat;
    __shared__ short s_h[1024];
    __shared__ short s_zmin[32], s_zmax[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TerrainKeys tk = terrain_keys(seed);
    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const vxp_coord c = coords[chunk];
        terrain_lattice_fill(s_lat, tk, 32 * c.x, 32 * c.z);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int z = warp + k * kWarps;                 // a warp owns the 32 columns of one z-layer
            const int h = terrain_height_lattice(s_lat, 32 * c.x + lane, 32 * c.z + z);
            s_h[z * 32 + lane] = (short)h;
            const int lo = __reduce_min_sync(kFull, h), hi = __reduce_max_sync(kFull, h);
            if (lane == 0) { s_zmin[z] = (short)lo; s_zmax[z] = (short)hi; }
        }
        __syncthreads();

        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)chunk * VXP_CHUNK_VOXELS);
        const int Y0 = 32 * c.y;
#pragma unroll 4
        for (int j = tid; j < 4096; j += kThreads) {         // j = (z*32 + y)*4 + p
            const int z = j >> 7, Y = Y0 + ((j >> 2) & 31), x0 = (j & 3) * 8;
            uint4 v;
            if (Y > s_zmax[z]) {                             // above every column of this layer: air
                v = make_uint4(0, 0, 0, 0);
            } else if (Y < s_zmin[z] - 3) {                  // below the dirt of every column: stone / slate by Y only
                const uint32_t b = ((Y >> 3) & 1) ? 4u : 3u, bb = b | b << 16;
                v = make_uint4(bb, bb, bb, bb);
            } else {
                const short* hp = s_h + z * 32 + x0;
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) w[e] = terrain_block(hp[2 * e], Y) | terrain_block(hp[2 * e + 1], Y) << 16;
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            st_voxels16(dst + j, v);
        }
        __syncthreads();
    }
}

// ====================================================================== pattern fill (SPEC §7)
__global__ void __launch_bounds__(kThreads) pattern_fill_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                 int pattern, uint16_t* __restrict__ out) {
    const int tid = threadIdx.x;
    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const vxp_coord c = coords[chunk];
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)chunk * VXP_CHUNK_VOXELS);
        for (int j = tid; j < 4096; j += kThreads) {
            const int Z = 32 * c.z + (j >> 7), Y = 32 * c.y + ((j >> 2) & 31), X0 = 32 * c.x + (j & 3) * 8;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                w[e] = pattern_block(pattern, X0 + 2 * e, Y, Z) | pattern_block(pattern, X0 + 2 * e + 1, Y, Z) << 16;
            dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// ====================================================================== fused terrain + mesh
// Terrain is a function of the (x,z) column only, so the chunks of a batch that share a column coordinate
// (cx,cz) share their 34x34 height footprint.  Four launches, chained with programmatic dependent launch (each
// kernel lets its successor become resident right away and waits for its predecessor's results itself), and
// no memsets: the hash table is epoch-stamped, and the last CTA of the mesh kernel hands the quad total to the
// caller and leaves the launch counters zeroed for the next call.
//   column_insert_kernel   one thread per chunk: find-or-insert (cx,cz) in an open-addressing hash table
//                          (atomicCAS on 64-bit keys); the inserting thread draws the column's id
//   column_heights_kernel  persistent CTAs, one column at a time: lattice gradients, 34x34 heights, min / max, column
//                          tops per 32-layer band
//   column_classify_kernel one thread per chunk: chunks entirely above / below the surface (most of a world) are
//                          finished here, the others go on a work list (the ones with many column tops at its front)
//   terrain_mesh_kernel    persistent CTAs draw work items from a ticket counter, turn the heights of the footprint
//                          straight into occupancy / equality words (masks_from_heights_g) and run the emit pipeline
//                          of vxp_mesh.cuh; voxels never exist.
constexpr int kColStride = 34 * 34 + 12;           // shorts per column record: footprint, hmin, hmax, kColBands band counts, 2 pad (8-byte multiple)
constexpr int kColBands = 8;                       // band b = chunk layer (hmin >> 5) + b: how many of the footprint's columns end in it
#ifndef VXP_HEAVY_MIN
#define VXP_HEAVY_MIN 400
#endif
constexpr int kHeavyMin = VXP_HEAVY_MIN;           // a chunk with at least this many column tops goes to the front of the work list
constexpr int kKeyEpochShift = 54;                 // key = epoch << 54 | (cx & 0x7FFFFFF) << 27 | (cz & 0x7FFFFFF)

__device__ __forceinline__ unsigned long long column_key(int cx, int cz, uint32_t epoch) {
    // 27 bits per coordinate cover every chunk whose world coordinates fit an i32 (|32 c| < 2^31)
    return (unsigned long long)epoch << kKeyEpochShift | (unsigned long long)((uint32_t)cx & 0x7FFFFFFu) << 27 |
           (unsigned long long)((uint32_t)cz & 0x7FFFFFFu);
}
__global__ void __launch_bounds__(kThreads) column_insert_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                  ColumnScratch cs) {
    mesh::griddep_launch_dependents();
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const vxp_coord c = coords[i];
    const unsigned long long key = column_key(c.x, c.z, cs.epoch);
    uint32_t slot = (uint32_t)(((key << 10) * 0x9E3779B97F4A7C15ull) >> 32) & cs.mask;   // the epoch does not move a column
    for (;;) {
        unsigned long long old = mesh::ld_relaxed_u64(cs.keys + slot);
        if ((old >> kKeyEpochShift) != cs.epoch) {       // a key of an earlier call (or none): the slot is free
            const unsigned long long seen = atomicCAS(cs.keys + slot, old, key);
            if (seen == old) {                           // we created the column
                const uint32_t id = atomicAdd(cs.count, 1u);
                cs.id_of_slot[slot] = id;
                cs.coord_of_id[id] = make_int2(c.x, c.z);
                break;
            }
            old = seen;                                  // somebody else took it, with a key of this call
        }
        if (old == key) break;
        slot = (slot + 1) & cs.mask;
    }
    cs.slot_of_chunk[i] = slot;
}
// 608 threads: the 1156 heights of a footprint are two per thread (a height is a chain of ~180 dependent instructions, and
// with few columns - an eighth of a world per GPU - the kernel is one wave of that latency)
constexpr int kHeightThreads = 608, kHeightWarps = kHeightThreads / 32;
__global__ void __launch_bounds__(kHeightThreads) column_heights_kernel(ColumnScratch cs, uint64_t seed) {
    __shared__ TerrainLattice s_lat;
    __shared__ int s_red[2 * kHeightWarps], s_band[kColBands];
    mesh::griddep_launch_dependents();
    mesh::griddep_wait();                              // the columns of this call have been inserted
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TerrainKeys tk = terrain_keys(seed);
    const uint32_t ncols = *cs.count;
    for (uint32_t col = blockIdx.x; col < ncols; col += gridDim.x) {
        const int2 c = cs.coord_of_id[col];
        terrain_lattice_fill(s_lat, tk, 32 * c.x - 1, 32 * c.y - 1);
        __syncthreads();
        short* dst = cs.heights + (size_t)col * kColStride;
        int hmin = INT_MAX, hmax = INT_MIN;
        int h[2];
        bool real[2];                                  // corners of the footprint are nobody's neighbour: not computed
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * kHeightThreads;
            const int zz = i / 34 - 1, xx = i % 34 - 1;
            real[k] = i < 34 * 34 && !((zz < 0 || zz > 31) && (xx < 0 || xx > 31));
            h[k] = 0;
            if (real[k]) {
                h[k] = terrain_height_lattice(s_lat, 32 * c.x + xx, 32 * c.y + zz);
                hmin = min(hmin, h[k]);
                hmax = max(hmax, h[k]);
            }
            if (i < 34 * 34) dst[i] = (short)h[k];
        }
        hmin = __reduce_min_sync(kFull, hmin);
        hmax = __reduce_max_sync(kFull, hmax);
        if (lane == 0) { s_red[warp] = hmin; s_red[kHeightWarps + warp] = hmax; }
        if (tid < kColBands) s_band[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kHeightWarps; ++w) { hmin = min(hmin, s_red[w]); hmax = max(hmax, s_red[kHeightWarps + w]); }
        // how many column tops fall into each 32-layer band from hmin's up: what a chunk of this column will cost
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (real[k]) atomicAdd(&s_band[min((h[k] >> 5) - (hmin >> 5), kColBands - 1)], 1);
        __syncthreads();
        if (tid == 0) {
            dst[34 * 34] = (short)hmin;
            dst[34 * 34 + 1] = (short)hmax;
        }
        if (tid < kColBands) dst[34 * 34 + 2 + tid] = (short)s_band[tid];
        __syncthreads();
    }
}

// One thread per chunk: chunks that are all air (including the layer below them) or all solid including every
// halo layer have no visible face and are finished here; the others go on the work list of the mesh kernel.
__global__ void __launch_bounds__(kThreads) column_classify_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                    ColumnScratch cs, vxp_chunk_meta* __restrict__ meta) {
    mesh::griddep_launch_dependents();
    mesh::griddep_wait();                              // heights written
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool surface = false, heavy = false;
    uint32_t id = 0;
    int cy = 0;
    if (i < n) {
        id = cs.id_of_slot[cs.slot_of_chunk[i]];
        cy = coords[i].y;
        const short* col = cs.heights + (size_t)id * kColStride;
        const int hmin = col[34 * 34], hmax = col[34 * 34 + 1], Y0 = 32 * cy;
        surface = !(Y0 > hmax || Y0 + 32 <= hmin);
        if (!surface) meta[i] = vxp_chunk_meta{0u, 0u};
        else heavy = col[34 * 34 + 2 + min(max(cy - (hmin >> 5), 0), kColBands - 1)] >= kHeavyMin;
    }
    // The work list is filled from both ends, one reservation per warp and end: chunks that hold many column tops (the
    // expensive ones: their cost goes with their quads) from the front, the others from the back.  The mesh kernel draws
    // from the front first, so the long items start early and the short ones fill the tail.
    const uint32_t mh = __ballot_sync(kFull, surface && heavy), ml = __ballot_sync(kFull, surface && !heavy);
    if ((mh | ml) == 0) return;
    uint32_t bh = 0, bl = 0;
    if (lane == 0) {
        if (mh) bh = atomicAdd(cs.count + 1, (uint32_t)__popc(mh));
        if (ml) bl = atomicAdd(cs.count + 6, (uint32_t)__popc(ml));
    }
    bh = __shfl_sync(kFull, bh, 0);
    bl = __shfl_sync(kFull, bl, 0);
    const uint32_t below = (1u << lane) - 1u;
    const uint4 entry = make_uint4(i, id, (uint32_t)cy, 0u);   // all the mesh kernel needs to start on the chunk
    if (surface && heavy) cs.work[bh + __popc(mh & below)] = entry;
    if (surface && !heavy) cs.work[n - 1u - (bl + __popc(ml & below))] = entry;
}

// Terrain fill of a large batch: the chunks of a column (cx,cz) share its heights (column_insert_kernel /
// column_heights_kernel above), so a chunk only reads its 32 x 32 heights (2 KiB, L2) and streams 64 KiB out.
// One CTA per chunk, same store loop as terrain_fill_kernel.
__global__ void __launch_bounds__(kThreads) terrain_fill_columns_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                         ColumnScratch cs, uint16_t* __restrict__ out) {
    __shared__ short s_h[1024];
    __shared__ short s_zmin[32], s_zmax[32];
    mesh::griddep_wait();                              // heights written
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n) return;
    const short* col = cs.heights + (size_t)cs.id_of_slot[cs.slot_of_chunk[chunk]] * kColStride;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = warp + k * kWarps;                 // a warp owns the 32 columns of one z-layer
        const int h = col[(z + 1) * 34 + lane + 1];
        s_h[z * 32 + lane] = (short)h;
        const int lo = __reduce_min_sync(kFull, h), hi = __reduce_max_sync(kFull, h);
        if (lane == 0) { s_zmin[z] = (short)lo; s_zmax[z] = (short)hi; }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)chunk * VXP_CHUNK_VOXELS);
    const int Y0 = 32 * coords[chunk].y;
#pragma unroll 4
    for (int j = tid; j < 4096; j += kThreads) {         // j = (z*32 + y)*4 + p
        const int z = j >> 7, Y = Y0 + ((j >> 2) & 31), x0 = (j & 3) * 8;
        uint4 v;
        if (Y > s_zmax[z]) {                             // above every column of this layer: air
            v = make_uint4(0, 0, 0, 0);
        } else if (Y < s_zmin[z] - 3) {                  // below the dirt of every column: stone / slate by Y only
            const uint32_t b = ((Y >> 3) & 1) ? 4u : 3u, bb = b | b << 16;
            v = make_uint4(bb, bb, bb, bb);
        } else {
            const short* hp = s_h + z * 32 + x0;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = terrain_block(hp[2 * e], Y) | terrain_block(hp[2 * e + 1], Y) << 16;
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        st_voxels16(dst + j, v);
    }
}

// Terrain fill of a small or medium batch, one launch: the first chunk of a column (cx,cz) to arrive claims the column in
// the hash table (the same epoch-stamped find-or-insert as column_insert_kernel), computes its 32 x 32 heights - the noise
// is 45 % of a chunk's instructions - and publishes them; the other chunks of the column wait for that word and read
// 2 KiB from L2 instead.  A chunk only ever waits for a CTA that is already running (it has made the claim), so nobody
// can be waiting for a CTA that has not been scheduled.  The ready word (cs.id_of_slot[slot]) is epoch << 22 | (id + 1).
// Leaves cs.count zeroed (the last CTA out), like the fused mesh kernel.
constexpr int kReadyIdBits = 22;
__global__ void __launch_bounds__(kThreads) terrain_fill_shared_kernel(const vxp_coord* __restrict__ coords, uint32_t n, uint64_t seed,
                                                                        ColumnScratch cs, uint32_t stride, uint16_t* __restrict__ out) {
    __shared__ TerrainLattice s_lat;
    __shared__ short s_h[1024];
    __shared__ short s_zmin[32], s_zmax[32];
    __shared__ uint32_t s_col[2];                      // column id, producer?
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTA b takes chunk (b * stride) mod n (stride coprime to n): the chunks of a column, neighbours in a sorted batch,
    // arrive far apart in time, so all but the first find the heights there instead of waiting for them together
    const uint32_t chunk = (uint32_t)(((unsigned long long)blockIdx.x * stride) % n);
    const vxp_coord c = coords[chunk];
    uint32_t slot = 0;
    if (tid == 0) {
        const unsigned long long key = column_key(c.x, c.z, cs.epoch);
        slot = (uint32_t)(((key << 10) * 0x9E3779B97F4A7C15ull) >> 32) & cs.mask;
        bool mine = false;
        for (;;) {
            unsigned long long old = mesh::ld_relaxed_u64(cs.keys + slot);
            if ((old >> kKeyEpochShift) != cs.epoch) {   // a key of an earlier call (or none): the slot is free
                const unsigned long long seen = atomicCAS(cs.keys + slot, old, key);
                if (seen == old) { mine = true; break; }
                old = seen;
            }
            if (old == key) break;
            slot = (slot + 1) & cs.mask;
        }
        uint32_t id;
        if (mine) {
            id = atomicAdd(cs.count, 1u);
        } else {                                         // the claimant is running: its heights are on their way
            uint32_t w;
            while (((w = mesh::ld_relaxed_u32(cs.id_of_slot + slot)) >> kReadyIdBits) != cs.epoch) __nanosleep(100);
            __threadfence();                             // the heights were written before the word
            id = (w & ((1u << kReadyIdBits) - 1u)) - 1u;
        }
        s_col[0] = id;
        s_col[1] = mine ? 1u : 0u;
    }
    __syncthreads();
    const uint32_t id = s_col[0];
    const bool mine = s_col[1] != 0u;                    // block-uniform
    short* col = cs.heights + (size_t)id * kColStride;   // this path keeps the 32 x 32 interior heights at the record's start
    if (mine) {
        const TerrainKeys tk = terrain_keys(seed);
        terrain_lattice_fill(s_lat, tk, 32 * c.x, 32 * c.z);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int z = warp + k * kWarps;             // a warp owns the 32 columns of one z-layer
            const int h = terrain_height_lattice(s_lat, 32 * c.x + lane, 32 * c.z + z);
            s_h[z * 32 + lane] = (short)h;
            col[z * 32 + lane] = (short)h;
        }
        __syncthreads();                                 // every height has been stored ...
        if (tid == 0) {
            __threadfence();                             // ... and is visible before the word that says so
            mesh::st_relaxed_u32(cs.id_of_slot + slot, cs.epoch << kReadyIdBits | (id + 1u));
        }
    } else {
        for (int i = tid; i < 512; i += kThreads)        // 2 KiB, 4 bytes at a time, past L1 (written during this launch)
            reinterpret_cast<uint32_t*>(s_h)[i] = __ldcg(reinterpret_cast<const uint32_t*>(col) + i);
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = warp + k * kWarps;
        const int h = s_h[z * 32 + lane];
        const int lo = __reduce_min_sync(kFull, h), hi = __reduce_max_sync(kFull, h);
        if (lane == 0) { s_zmin[z] = (short)lo; s_zmax[z] = (short)hi; }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)chunk * VXP_CHUNK_VOXELS);
    const int Y0 = 32 * c.y;
#pragma unroll 4
    for (int j = tid; j < 4096; j += kThreads) {         // j = (z*32 + y)*4 + p
        const int z = j >> 7, Y = Y0 + ((j >> 2) & 31), x0 = (j & 3) * 8;
        uint4 v;
        if (Y > s_zmax[z]) {                             // above every column of this layer: air
            v = make_uint4(0, 0, 0, 0);
        } else if (Y < s_zmin[z] - 3) {                  // below the dirt of every column: stone / slate by Y only
            const uint32_t b = ((Y >> 3) & 1) ? 4u : 3u, bb = b | b << 16;
            v = make_uint4(bb, bb, bb, bb);
        } else {
            const short* hp = s_h + z * 32 + x0;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = terrain_block(hp[2 * e], Y) | terrain_block(hp[2 * e + 1], Y) << 16;
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        st_voxels16(dst + j, v);
    }
    if (tid == 0 && atomicAdd(cs.count + 3, 1u) == gridDim.x - 1) {   // last CTA out: leave the counters zeroed
        cs.count[0] = 0u;
        cs.count[3] = 0u;
    }
}

// ---- one surface chunk: footprint -> occupancy / equality words, no voxels in between.
// A terrain column (x, z) is a function of y alone: solid up to its height, grass on top, three layers of dirt, then stone
// and slate in bands of eight.  So the 32 voxels of a column are four words with bit y set - occupancy and the three id
// bit planes - and each is a handful of shifts (where the unfused path writes, reads back and converts 32 voxels).
// Words over y are what the X planes want as they are (face masks: the lanes next door; equality along y: a shift); the
// Y / Z planes want words over x, which 32 x 32 bit transposes across the warp deliver.  s_h[(z+1)*34 + (x+1)].
__device__ __forceinline__ uint32_t bits_le(int n) {                // bits 0..n of a word (none for n < 0, all for n >= 31)
    const int m = min(max(n + 1, 0), 32);
    return m == 32 ? kFull : (1u << m) - 1u;
}
struct ColumnWords {
    uint32_t occ, p0, p1, p2;                                        // bit y: voxel solid / bit b of its block id
};
// t = height of the column relative to the chunk's first layer (the chunk's Y0 is a multiple of 32, so the slate bands
// (Y >> 3) & 1 are the same bits in every chunk); same blocks as terrain_block()
__device__ __forceinline__ ColumnWords column_words(int t) {
    constexpr uint32_t kSlate = 0xFF00FF00u;
    ColumnWords c;
    c.occ = bits_le(t);
    const uint32_t below = bits_le(t - 1), deep = bits_le(t - 4);
    const uint32_t grass = c.occ & ~below, dirt = below & ~deep, stone = deep & ~kSlate;   // ids 1, 2, 3; slate = 4
    c.p0 = grass | stone;
    c.p1 = dirt | stone;
    c.p2 = deep & kSlate;
    return c;
}
// Z halo layers (z = -1, 32) of the occupancy: one transpose each, warps 0 and 1
template <class SM>
__device__ __forceinline__ void z_halo_from_heights(SM& sm, const short* s_h, int Y0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= 2) return;
    const int zz = warp ? 33 : 0;                                     // row of the footprint
    uint32_t w[1] = {bits_le(s_h[zz * 34 + lane + 1] - Y0)};
    transpose32xN(w, lane);
    sm.occ[zz * 34 + lane + 1] = w[0];
}

// Everything emit_greedy() reads, from the heights: occupancy with halo, EX / EY / EZ, the X-plane masks and their
// transposed equality rows, the non-empty rows of every plane.  Returns the OR of all face words.  sm.rows zeroed by the caller.
__device__ __forceinline__ uint32_t masks_from_heights_g(mesh::SmemG& sm, const short* s_h, int Y0) {
    using namespace mesh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PostG& P = sm.tile.post;
    uint32_t ta[8], tb[8];                                            // to transpose: occupancy, EX | EY, EZ (four layers each)
    uint32_t acc = 0, rx0 = 0, rx1 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp, x = lane;
        const short* hp = s_h + (z + 1) * 34 + x + 1;
        const int t = hp[0] - Y0;
        const ColumnWords c = column_words(t), cz = column_words(hp[-34] - Y0);   // this column and its -z neighbour
        uint32_t left = __shfl_up_sync(kFull, c.occ, 1), right = __shfl_down_sync(kFull, c.occ, 1);
        if (x == 0) left = bits_le(hp[-1] - Y0);                      // the halo columns
        if (x == 31) right = bits_le(hp[1] - Y0);
        const uint32_t xm0 = c.occ & ~left, xm1 = c.occ & ~right;     // -X / +X visible, bit y: plane x, row z as it is
        P.xmask[(0 * 32 + x) * kPad + z] = xm0;
        P.xmask[(1 * 32 + x) * kPad + z] = xm1;
        rx0 |= (uint32_t)(xm0 != 0u) << z;
        rx1 |= (uint32_t)(xm1 != 0u) << z;
        acc |= xm0 | xm1;
        uint32_t l0 = __shfl_up_sync(kFull, c.p0, 1), l1 = __shfl_up_sync(kFull, c.p1, 1), l2 = __shfl_up_sync(kFull, c.p2, 1);
        if (x == 0) l0 = l1 = l2 = 0u;
        const uint32_t dx = (c.p0 ^ l0) | (c.p1 ^ l1) | (c.p2 ^ l2);
        const uint32_t dy = ((c.p0 ^ (c.p0 << 1)) | (c.p1 ^ (c.p1 << 1)) | (c.p2 ^ (c.p2 << 1))) & ~1u;
        const uint32_t dz = z > 0 ? (c.p0 ^ cz.p0) | (c.p1 ^ cz.p1) | (c.p2 ^ cz.p2) : 0u;
        P.ht[x * kPad + z] = ~dy;                                     // X planes: u = y, v = z
        P.vt[x * kPad + z] = ~dz;
        ta[k] = c.occ;
        ta[4 + k] = ~dx;
        tb[k] = ~dy;
        tb[4 + k] = ~dz;
        // Y halo rows (y = -1, 32) of layer z, bit x
        const uint32_t b_lo = __ballot_sync(kFull, t >= -1), b_hi = __ballot_sync(kFull, t >= 32);
        if (lane == 0) { sm.occ[(z + 1) * 34] = b_lo; sm.occ[(z + 1) * 34 + 33] = b_hi; }
    }
    transpose32xN(ta, lane);
    transpose32xN(tb, lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                     // lane = y now
        const int z = k * kWarps + warp, y = lane;
        sm.occ[(z + 1) * 34 + y + 1] = ta[k];
        sm.eq[0][z * kPad + y] = ta[4 + k];
        sm.eq[1][z * kPad + y] = tb[k];
        sm.eq[2][z * kPad + y] = tb[4 + k];
    }
    z_halo_from_heights(sm, s_h, Y0);
    if (rx0) atomicOr(&sm.rows[0 * 32 + lane], rx0);
    if (rx1) atomicOr(&sm.rows[1 * 32 + lane], rx1);
    __syncthreads();                                                  // occupancy with halo published
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp, y = lane;
        const uint32_t* o = sm.occ + (z + 1) * 34 + (y + 1);
        const uint32_t c = ta[k];
        const uint32_t my = c & ~o[-1], py = c & ~o[+1], mz = c & ~o[-34], pz = c & ~o[+34];
        acc |= my | py | mz | pz;
        const uint32_t bz0 = __ballot_sync(kFull, mz != 0u), bz1 = __ballot_sync(kFull, pz != 0u);
        if (lane == 0) { sm.rows[4 * 32 + z] = bz0; sm.rows[5 * 32 + z] = bz1; }
        if (my) atomicOr(&sm.rows[2 * 32 + y], 1u << z);
        if (py) atomicOr(&sm.rows[3 * 32 + y], 1u << z);
    }
    return acc;
}

template <class SM>
__device__ __forceinline__ int load_footprint(SM& sm, const ColumnScratch& cs, const uint4& item) {   // item: {chunk, column id, chunk y}
    const int tid = threadIdx.x;
    const short* col = cs.heights + (size_t)item.y * kColStride;
    for (int i = tid; i < 34 * 34 / 2; i += kThreads)     // 2312 bytes, 4 at a time
        reinterpret_cast<uint32_t*>(sm.aux.heights)[i] = __ldg(reinterpret_cast<const uint32_t*>(col) + i);
    return 32 * (int)item.z;
}

__device__ __forceinline__ void terrain_mesh_chunk(mesh::SmemG& sm, const ColumnScratch& cs, const uint4& item, const mesh::MeshOut& out) {
    const int tid = threadIdx.x;
    const uint32_t chunk = item.x;
    mesh::Prof prof;
    prof.start(false);
    const int Y0 = load_footprint(sm, cs, item);
    const short* s_h = sm.aux.heights;
    if (tid < 192) sm.rows[tid] = 0u;
    __syncthreads();
    const bool any = __syncthreads_or(masks_from_heights_g(sm, s_h, Y0) != 0);
    if (!any) {
        if (tid == 0) out.meta[chunk] = vxp_chunk_meta{0u, 0u};
        return;
    }
    // terrain ids are 1..4: the equality rows always come from the id planes
    mesh::emit_greedy(sm, chunk, true, out, [s_h, Y0](int idx) -> uint32_t {
        const int x = idx & 31, y = (idx >> 5) & 31, z = idx >> 10;
        return terrain_block(s_h[(z + 1) * 34 + x + 1], Y0 + y);
    }, prof);
}
__device__ __forceinline__ void terrain_mesh_chunk(mesh::SmemC& sm, const ColumnScratch& cs, const uint4& item, const mesh::MeshOut& out) {
    using namespace mesh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chunk = item.x;
    Prof prof;
    prof.start(false);
    const int Y0 = load_footprint(sm, cs, item);
    const short* s_h = sm.aux.heights;
    __syncthreads();
    // occupancy only: words over y per column, transposed to words over x; halo rows by ballot / transpose / as they are
    uint32_t ta[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp;
        const int t = s_h[(z + 1) * 34 + lane + 1] - Y0;
        ta[k] = bits_le(t);
        const uint32_t b_lo = __ballot_sync(kFull, t >= -1), b_hi = __ballot_sync(kFull, t >= 32);
        if (lane == 0) { sm.occ[(z + 1) * 34] = b_lo; sm.occ[(z + 1) * 34 + 33] = b_hi; }
    }
    transpose32xN(ta, lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) sm.occ[(k * kWarps + warp + 1) * 34 + lane + 1] = ta[k];
    z_halo_from_heights(sm, s_h, Y0);
    if (tid < 64) sm.xh[tid >> 5][tid & 31] = bits_le(s_h[((tid & 31) + 1) * 34 + ((tid >> 5) ? 33 : 0)] - Y0);   // X halos: bit y per z
    __syncthreads();                                  // occupancy and halo published
    culled_emit_c(sm, chunk, culled_count<false>(sm) + culled_count<true>(sm), out, [s_h, Y0](int idx) -> uint32_t {
        const int x = idx & 31, y = (idx >> 5) & 31, z = idx >> 10;
        return terrain_block(s_h[(z + 1) * 34 + x + 1], Y0 + y);
    }, prof);
}

// cs.count: [0] columns, [1] heavy work items, [2] ticket, [3] CTAs done, [4..5] quads reserved (64 bit), [6] light work items: all zero between calls
template <bool kGreedy>
__global__ void __launch_bounds__(kThreads, kGreedy ? mesh::kGreedyCtas : mesh::kCulledCtas)
terrain_mesh_kernel(const vxp_coord* __restrict__ coords, uint32_t n, ColumnScratch cs, mesh::MeshOut out,
                    unsigned long long* __restrict__ user_total) {
    using SM = std::conditional_t<kGreedy, mesh::SmemG, mesh::SmemC>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    __shared__ uint4 s_item[2];
    mesh::griddep_wait();                             // work list complete
    const uint32_t nheavy = cs.count[1], nwork = nheavy + cs.count[6];
    const uint4 none = make_uint4(~0u, 0u, 0u, 0u);
    auto entry_of = [&](uint32_t t) { return cs.work[t < nheavy ? t : n - 1u - (t - nheavy)]; };   // heavy items first (see column_classify_kernel)
    // Work items are drawn from a ticket counter (chunk costs vary widely), two ahead: while item k is meshed, thread 0
    // has the ticket of k+1 in hand, fetches that item's entry and draws the ticket of k+2 - both round trips hide
    // behind the chunk, and an item starts with everything it needs to request its heights.
    uint32_t t_next = 0;
    if (threadIdx.x == 0) {
        if constexpr (kGreedy) { sm.ones = kFull; sm.zero = 0u; }
        const uint32_t t0 = atomicAdd(cs.count + 2, 1u);
        t_next = atomicAdd(cs.count + 2, 1u);
        s_item[0] = t0 < nwork ? entry_of(t0) : none;
    }
    __syncthreads();
    for (uint32_t it = 0;; ++it) {
        const uint4 item = s_item[it & 1];
        if (item.x == ~0u) break;
        uint4 upcoming = none;
        uint32_t t_after = 0;
        if (threadIdx.x == 0) {
            if (t_next < nwork) upcoming = entry_of(t_next);
            t_after = atomicAdd(cs.count + 2, 1u);
        }
#ifdef VXP_PROFILE
        if (threadIdx.x == 0 && item.x < mesh::kProfCtas) mesh::g_prof_time[item.x][0] = mesh::prof_now();
#endif
        terrain_mesh_chunk(sm, cs, item, out);
#ifdef VXP_PROFILE
        if (threadIdx.x == 0 && item.x < mesh::kProfCtas) mesh::g_prof_time[item.x][1] = mesh::prof_now();
#endif
        if (threadIdx.x == 0) {
            s_item[(it + 1) & 1] = upcoming;
            t_next = t_after;
        }
        __syncthreads();                              // shared memory is free again, the next item is visible
    }
    if (threadIdx.x == 0) {
        __threadfence();                              // this CTA's reservations are ordered before its check-out
        if (atomicAdd(cs.count + 3, 1u) == gridDim.x - 1) {   // last CTA out: hand the total over, leave the counters zeroed
            __threadfence();
            *user_total = atomicExch(out.total, 0ull);
            cs.count[0] = 0u;
            cs.count[1] = 0u;
            cs.count[2] = 0u;
            cs.count[3] = 0u;
            cs.count[6] = 0u;
        }
    }
}

// ====================================================================== launchers

template <class K>
static cudaError_t set_smem(K fn, size_t bytes) {
    return cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t configure_device() {
    cudaError_t e = set_smem(mesh::mesh_kernel_g, sizeof(mesh::SmemG));
    if (e == cudaSuccess) e = set_smem(mesh::mesh_kernel_c, sizeof(mesh::SmemC));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel<false>, sizeof(mesh::SmemC));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel<true>, sizeof(mesh::SmemG));
    return e;
}

// A launch that may become resident while its predecessor on the stream is still running (programmatic dependent
// launch); the kernel waits for the predecessor's results itself (griddepcontrol.wait).
template <class... Params, class... Args>
static cudaError_t launch_dependent_threads(void (*kernel)(Params...), uint32_t grid, uint32_t threads, size_t smem, cudaStream_t stream,
                                            Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}
template <class... Params, class... Args>
static cudaError_t launch_dependent(void (*kernel)(Params...), uint32_t grid, size_t smem, cudaStream_t stream, Args... args) {
    return launch_dependent_threads(kernel, grid, (uint32_t)kThreads, smem, stream, args...);
}

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

cudaError_t launch_terrain_fill(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    terrain_fill_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, seed, d_voxels);
    return cudaGetLastError();
}

cudaError_t launch_terrain_fill_shared(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                       const ColumnScratch& cs, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint32_t stride = (uint32_t)(0.6180339887 * n) | 1u;   // the batch in golden-ratio steps
    auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
    while (gcd(stride, n) != 1u) stride += 2;
    terrain_fill_shared_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, seed, cs, stride % n, d_voxels);
    return cudaGetLastError();
}

// columns of the batch -> hash table -> heights (the first two launches of the fused path and of the large-batch fill)
static cudaError_t launch_columns(const vxp_coord* d_coords, uint32_t n, uint64_t seed, const ColumnScratch& cs, cudaStream_t stream) {
    column_insert_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_coords, n, cs);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint32_t sms = (uint32_t)sm_count();
    const uint32_t hgrid = n < 4 * sms ? n : 4 * sms;
    return launch_dependent_threads(column_heights_kernel, hgrid, (uint32_t)kHeightThreads, 0, stream, cs, seed);
}

cudaError_t launch_terrain_fill_columns(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                        const ColumnScratch& cs, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = launch_columns(d_coords, n, seed, cs, stream);
    if (e == cudaSuccess) e = launch_dependent(terrain_fill_columns_kernel, n, 0, stream, d_coords, n, cs, d_voxels);
    // the fused path expects the launch counters zeroed (its own last CTA leaves them so)
    if (e == cudaSuccess) e = cudaMemsetAsync(cs.count, 0, kColumnCounterWords * sizeof(uint32_t), stream);
    return e;
}

cudaError_t launch_pattern_fill(const vxp_coord* d_coords, uint32_t n, int pattern, uint16_t* d_voxels,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    pattern_fill_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, pattern, d_voxels);
    return cudaGetLastError();
}

static mesh::MeshOut mesh_out(const vxp_mesh_dev& o) {
    return mesh::MeshOut{reinterpret_cast<unsigned long long*>(o.verts), o.indices, o.meta,
                         reinterpret_cast<unsigned long long*>(o.total_quads), o.capacity_quads, 0ull};
}

static cudaError_t launch_mesh_kernel(bool greedy, uint32_t grid, int overlap, cudaStream_t stream, const uint16_t* d_voxels,
                                      const int32_t* d_nbr6, uint32_t n, uint32_t lo, uint32_t hi, mesh::MeshOut out, mesh::Bnd bnd,
                                      mesh::Chain chain) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = greedy ? sizeof(mesh::SmemG) : sizeof(mesh::SmemC);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    if (greedy) return cudaLaunchKernelEx(&cfg, mesh::mesh_kernel_g, d_voxels, d_nbr6, n, lo, hi, out, bnd, chain);
    return cudaLaunchKernelEx(&cfg, mesh::mesh_kernel_c, d_voxels, d_nbr6, n, lo, hi, out, bnd, chain);
}

cudaError_t launch_mesh_range(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, uint32_t lo, uint32_t hi,
                              int greedy, const vxp_mesh_dev& o, const BoundaryScratch& scratch, cudaStream_t stream) {
    if (hi > n) hi = n;
    if (lo >= hi) return cudaSuccess;
    const mesh::Bnd bnd{scratch.tags, scratch.planes, scratch.epoch};
    const uint32_t grid = hi - lo;
    const mesh::Chain chain{nullptr, nullptr, nullptr, 0u, 0u};
    return launch_mesh_kernel(greedy != 0, grid, 0, stream, d_voxels, d_nbr6, n, lo, hi, mesh_out(o), bnd, chain);
}

// Stride of the run dealing (see mesh::Chain); a build-time constant so that tuning runs are `make variant` builds.
#ifndef VXP_DEAL_STRIDE
#define VXP_DEAL_STRIDE 17
#endif
cudaError_t launch_mesh(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, int greedy,
                        const vxp_mesh_dev& o, const BoundaryScratch& scratch, unsigned long long* counter, int overlap,
                        cudaStream_t stream) {
    if (n == 0) return cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    if (n > (1u << 22) - 1u) return cudaErrorInvalidValue;   // CTA count field of the launch counter
    const mesh::Bnd bnd{scratch.tags, scratch.planes, scratch.epoch};
    mesh::MeshOut out = mesh_out(o);
    out.total = counter;
    out.cta_inc = mesh::kCtaOne;
    // deal the runs of 16 chunks out with a small stride coprime to their number (see mesh::Chain).  Measured on
    // config 2 / 3 (256 runs): identity 75.7 / 82.4 us, 17 -> 69.5 / 75.4, 5 -> 70.4 / 75.7, 33 -> 70.7 / 76.0,
    // 97 -> 72.3 / 77.0, 159 (0.618 R) -> 71.9 / 76.6: small strides keep +-Y / +-Z neighbours close enough in
    // time for their halo rows to hit L2.
    uint32_t runs = n >> 4, stride = 0;
    if (runs > 34 && VXP_DEAL_STRIDE > 0) {
        auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
        stride = ((uint32_t)VXP_DEAL_STRIDE | 1u) % runs;
        while (gcd(stride, runs) != 1u) stride += 2u;
    } else {
        runs = 0;
    }
    const mesh::Chain chain{reinterpret_cast<unsigned long long*>(o.total_quads), nullptr, nullptr, runs, stride};
    return launch_mesh_kernel(greedy != 0, n, overlap, stream, d_voxels, d_nbr6, n, 0, n, out, bnd, chain);
}

// ---- incremental re-mesh
// One thread per edit: write the voxel, mark the chunk and - for a voxel on a chunk face - the neighbour across
// that face (a symmetric neighbour table is assumed: the chunk that sees this voxel through its halo is nbr6[c][f]).
// flags[c] == epoch means "already on the list"; the thread that flips it appends the chunk.
__global__ void __launch_bounds__(kThreads) apply_edits_kernel(uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6,
                                                                uint32_t n, const vxp_edit* __restrict__ edits, uint32_t m,
                                                                uint32_t* __restrict__ flags, uint32_t epoch,
                                                                uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= m) return;
    const vxp_edit e = edits[i];
    if (e.chunk >= n || e.x >= 32u || e.y >= 32u || e.z >= 32u) return;     // out of range: ignored
    voxels[(size_t)e.chunk * VXP_CHUNK_VOXELS + e.x + 32u * e.y + 1024u * e.z] = e.id;
    auto mark = [&](uint32_t c) {
        if (atomicExch(flags + c, epoch) != epoch) list[atomicAdd(count, 1u)] = c;
    };
    mark(e.chunk);
    if (nbr6 == nullptr) return;
    const int32_t* nb = nbr6 + 6 * (size_t)e.chunk;
    auto across = [&](int f) {
        const int32_t c = __ldg(nb + f);
        if (c >= 0 && (uint32_t)c < n) mark((uint32_t)c);
    };
    if (e.x == 0u) across(0);
    if (e.x == 31u) across(1);
    if (e.y == 0u) across(2);
    if (e.y == 31u) across(3);
    if (e.z == 0u) across(4);
    if (e.z == 31u) across(5);
}

cudaError_t launch_apply_edits(uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, const vxp_edit* d_edits, uint32_t m,
                               uint32_t* d_flags, uint32_t epoch, uint32_t* d_list, uint32_t* d_count, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess || m == 0) return e;
    apply_edits_kernel<<<(m + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_voxels, d_nbr6, n, d_edits, m, d_flags, epoch,
                                                                              d_list, d_count);
    return cudaGetLastError();
}

cudaError_t launch_mesh_list(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, const uint32_t* d_list,
                             const uint32_t* d_count, uint32_t max_count, int greedy, const vxp_mesh_dev& o,
                             unsigned long long* counter, cudaStream_t stream) {
    if (max_count == 0) return cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    if (max_count > (1u << 22) - 1u) return cudaErrorInvalidValue;
    const mesh::Bnd bnd{nullptr, nullptr, 0};      // the x neighbours are not meshed alongside: their boundary is gathered
    mesh::MeshOut out = mesh_out(o);
    out.total = counter;
    out.cta_inc = mesh::kCtaOne;
    const mesh::Chain chain{reinterpret_cast<unsigned long long*>(o.total_quads), d_list, d_count, 0u, 0u};
    return launch_mesh_kernel(greedy != 0, max_count, 0, stream, d_voxels, d_nbr6, n, 0, n, out, bnd, chain);
}

size_t boundary_tag_bytes(uint32_t n) { return (size_t)n * sizeof(uint32_t); }
size_t boundary_plane_bytes(uint32_t n) { return (size_t)n * 64 * sizeof(unsigned long long); }

#ifdef VXP_DEBUG_EXPORTS
// diagnostics of the profiling / tuning builds (make prof, make variant); not in libvxp.so, not part of include/vxp.h
extern "C" int vxp_debug_occupancy(int greedy, int fused) {
    int n = -1;
    cudaError_t e;
    if (fused)
        e = greedy ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, terrain_mesh_kernel<true>, kThreads, sizeof(mesh::SmemG))
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, terrain_mesh_kernel<false>, kThreads, sizeof(mesh::SmemC));
    else
        e = greedy ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mesh::mesh_kernel_g, kThreads, sizeof(mesh::SmemG))
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mesh::mesh_kernel_c, kThreads, sizeof(mesh::SmemC));
    return e == cudaSuccess ? n : -1;
}
extern "C" int vxp_debug_smem_bytes(int greedy) { return greedy ? (int)sizeof(mesh::SmemG) : (int)sizeof(mesh::SmemC); }
#endif

#ifdef VXP_PROFILE
extern "C" int vxp_debug_read_prof_time(unsigned long long* out, int max_ctas) {
    const int ctas = max_ctas < mesh::kProfCtas ? max_ctas : mesh::kProfCtas;
    return cudaMemcpyFromSymbol(out, mesh::g_prof_time, sizeof(unsigned long long) * 2 * (size_t)ctas) == cudaSuccess ? 2 : -1;
}
extern "C" int vxp_debug_read_prof_warp(unsigned int* out, int max_ctas) {
    const int ctas = max_ctas < mesh::kProfCtas ? max_ctas : mesh::kProfCtas;
    return cudaMemcpyFromSymbol(out, mesh::g_prof_warp, sizeof(unsigned int) * 8 * (size_t)ctas) == cudaSuccess ? 8 : -1;
}
// profiling build only: read (and clear) the per-CTA phase counters of the mesh kernels
extern "C" int vxp_debug_read_prof(unsigned long long* out, int max_ctas) {
    const int ctas = max_ctas < mesh::kProfCtas ? max_ctas : mesh::kProfCtas;
    const size_t bytes = sizeof(unsigned long long) * mesh::kProfSlots * (size_t)ctas;
    if (cudaMemcpyFromSymbol(out, mesh::g_prof, bytes) != cudaSuccess) return -1;
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, mesh::g_prof) != cudaSuccess) return -2;
    if (cudaMemset(sym, 0, sizeof(unsigned long long) * mesh::kProfSlots * mesh::kProfCtas) != cudaSuccess) return -3;
    return mesh::kProfSlots;
}
#endif

size_t column_table_slots(uint32_t n) {
    size_t t = 64;
    while (t < 2 * (size_t)n) t <<= 1;
    return t;
}
size_t column_record_bytes() { return kColStride * sizeof(short); }

// Persistent CTAs per SM of the fused mesh kernel: what fits (shared memory), and no more CTAs than work items can feed.
cudaError_t launch_terrain_mesh(const vxp_coord* d_coords, uint32_t n, uint64_t seed, int greedy,
                                const vxp_mesh_dev& o, const ColumnScratch& cs, cudaStream_t stream) {
    if (n == 0) return cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    cudaError_t e = launch_columns(d_coords, n, seed, cs, stream);
    if (e == cudaSuccess)
        e = launch_dependent(column_classify_kernel, (n + kThreads - 1) / kThreads, 0, stream, d_coords, n, cs, o.meta);
    if (e != cudaSuccess) return e;
    // the work list's length is only known on the device: a resident grid strides over it
    const uint32_t per_sm = greedy ? (uint32_t)mesh::kGreedyCtas : (uint32_t)mesh::kCulledCtas;
    const uint32_t cap = per_sm * (uint32_t)sm_count();
    const uint32_t mgrid = n < cap ? n : cap;
    mesh::MeshOut out = mesh_out(o);
    out.total = reinterpret_cast<unsigned long long*>(cs.count + 4);
    unsigned long long* user_total = reinterpret_cast<unsigned long long*>(o.total_quads);
    if (greedy) return launch_dependent(terrain_mesh_kernel<true>, mgrid, sizeof(mesh::SmemG), stream, d_coords, n, cs, out, user_total);
    return launch_dependent(terrain_mesh_kernel<false>, mgrid, sizeof(mesh::SmemC), stream, d_coords, n, cs, out, user_total);
}

} // namespace vxp
