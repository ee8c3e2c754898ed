// vxp_kernels.cu — sm_100a kernels: terrain fill, pattern fill, stored-chunk meshing,
// fused terrain+mesh.  SPEC.md v1.  Host launchers at the bottom (C++ linkage, used by vxp_api.cu).
#include <climits>

#include <cstdlib>
#include <cstring>

#include "vxp_launch.h"
#include "vxp_mesh4.cuh"
#include "vxp_mesh5.cuh"

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
            dst[j] = v;
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

// ====================================================================== stored-chunk meshing
// v4::mesh_kernel lives in vxp_mesh4.cuh (one CTA per chunk, three CTAs per SM).

// ====================================================================== fused terrain + mesh
// Terrain is a function of the (x,z) column only, so the chunks of a batch that share a column coordinate
// (cx,cz) share their 34x34 height footprint.  Three launches:
//   column_insert_kernel   one thread per chunk: find-or-insert (cx,cz) in an open-addressing hash table
//                          (atomicCAS on 64-bit keys); the inserting thread draws the column's id
//   column_heights_kernel  persistent CTAs, one column at a time: lattice gradients, 34x34 heights, min / max
//   terrain_mesh_kernel    one CTA per chunk: read min / max (most chunks of a world are entirely above or
//                          below the surface and stop there), else the footprint, build the tile in shared
//                          memory and run the meshing pipeline of vxp_mesh4.cuh; voxels never reach HBM.
constexpr int kColStride = 34 * 34 + 4;            // shorts per column record: footprint, hmin, hmax, 2 pad (8-byte multiple)
constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ unsigned long long column_key(int cx, int cz) {
    return ((unsigned long long)(uint32_t)cx << 32 | (uint32_t)cz) ^ 0x8000000080000000ull;   // never kEmptyKey for sane coords
}
__global__ void __launch_bounds__(kThreads) column_insert_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                  ColumnScratch cs) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const vxp_coord c = coords[i];
    const unsigned long long key = column_key(c.x, c.z);
    uint32_t slot = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & cs.mask;
    for (;;) {
        const unsigned long long old = atomicCAS(cs.keys + slot, kEmptyKey, key);
        if (old == kEmptyKey) {                      // we created the column
            const uint32_t id = atomicAdd(cs.count, 1u);
            cs.id_of_slot[slot] = id;
            cs.coord_of_id[id] = make_int2(c.x, c.z);
            break;
        }
        if (old == key) break;
        slot = (slot + 1) & cs.mask;
    }
    cs.slot_of_chunk[i] = slot;
}
__global__ void __launch_bounds__(kThreads) column_heights_kernel(ColumnScratch cs, uint64_t seed) {
    __shared__ TerrainLattice s_lat;
    __shared__ int s_red[2 * kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TerrainKeys tk = terrain_keys(seed);
    const uint32_t ncols = *cs.count;
    for (uint32_t col = blockIdx.x; col < ncols; col += gridDim.x) {
        const int2 c = cs.coord_of_id[col];
        terrain_lattice_fill(s_lat, tk, 32 * c.x - 1, 32 * c.y - 1);
        __syncthreads();
        short* dst = cs.heights + (size_t)col * kColStride;
        int hmin = INT_MAX, hmax = INT_MIN;
        for (int i = tid; i < 34 * 34; i += kThreads) {
            const int zz = i / 34 - 1, xx = i % 34 - 1;
            const bool corner = (zz < 0 || zz > 31) && (xx < 0 || xx > 31);
            int h = 0;
            if (!corner) {
                h = terrain_height_lattice(s_lat, 32 * c.x + xx, 32 * c.y + zz);
                hmin = min(hmin, h);
                hmax = max(hmax, h);
            }
            dst[i] = (short)h;
        }
        hmin = __reduce_min_sync(kFull, hmin);
        hmax = __reduce_max_sync(kFull, hmax);
        if (lane == 0) { s_red[warp] = hmin; s_red[kWarps + warp] = hmax; }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) { hmin = min(hmin, s_red[w]); hmax = max(hmax, s_red[kWarps + w]); }
            dst[34 * 34] = (short)hmin;
            dst[34 * 34 + 1] = (short)hmax;
        }
        __syncthreads();
    }
}

// One thread per chunk: chunks that are all air (including the layer below them) or all solid including every
// halo layer have no visible face and are finished here; the others go on the work list of the mesh kernel.
__global__ void __launch_bounds__(kThreads) column_classify_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                    ColumnScratch cs, vxp_chunk_meta* __restrict__ meta) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool surface = false;
    if (i < n) {
        const short* col = cs.heights + (size_t)cs.id_of_slot[cs.slot_of_chunk[i]] * kColStride;
        const int hmin = col[34 * 34], hmax = col[34 * 34 + 1], Y0 = 32 * coords[i].y;
        surface = !(Y0 > hmax || Y0 + 32 <= hmin);
        if (!surface) meta[i] = vxp_chunk_meta{0u, 0u};
    }
    const uint32_t m = __ballot_sync(kFull, surface);          // one list reservation per warp
    if (m == 0) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cs.count + 1, (uint32_t)__popc(m));
    base = __shfl_sync(kFull, base, 0);
    if (surface) cs.slot_of_chunk[n + base + __popc(m & ((1u << lane) - 1u))] = i;   // work list lives behind slot_of_chunk
}

// Terrain fill of a large batch: the chunks of a column (cx,cz) share its heights (column_insert_kernel /
// column_heights_kernel above), so a chunk only reads its 32 x 32 heights (2 KiB, L2) and streams 64 KiB out.
// One CTA per chunk, same store loop as terrain_fill_kernel.
__global__ void __launch_bounds__(kThreads) terrain_fill_columns_kernel(const vxp_coord* __restrict__ coords, uint32_t n,
                                                                         ColumnScratch cs, uint16_t* __restrict__ out) {
    __shared__ short s_h[1024];
    __shared__ short s_zmin[32], s_zmax[32];
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
        dst[j] = v;
    }
}

// One surface chunk: footprint -> tile in shared memory -> the meshing pipeline of vxp_mesh4.cuh.
template <bool kGreedy>
__device__ __forceinline__ void terrain_mesh_chunk(v4::Smem<kGreedy>& sm, short* s_h, const vxp_coord* __restrict__ coords,
                                                   const ColumnScratch& cs, uint32_t chunk, const v4::MeshOut& out) {
    const int tid = threadIdx.x;
    const vxp_coord c = coords[chunk];
    v4::Prof prof;
    prof.start(false);
    const short* col = cs.heights + (size_t)cs.id_of_slot[cs.slot_of_chunk[chunk]] * kColStride;
    const int Y0 = 32 * c.y;
    if (tid == 0) sm.ones = kFull;
    for (int i = tid; i < 34 * 34 / 2; i += kThreads)     // 2312 bytes, 4 at a time
        reinterpret_cast<uint32_t*>(s_h)[i] = __ldg(reinterpret_cast<const uint32_t*>(col) + i);
    __syncthreads();
    // tile
    uint4* tile4 = reinterpret_cast<uint4*>(sm.tile.vox);
#pragma unroll 2
    for (int j = tid; j < 4096; j += kThreads) {
        const int z = j >> 7, Y = Y0 + ((j >> 2) & 31), x0 = (j & 3) * 8;
        const short* hp = s_h + (z + 1) * 34 + x0 + 1;
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = terrain_block(hp[2 * e], Y) | terrain_block(hp[2 * e + 1], Y) << 16;
        tile4[j] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // halo occupancy straight from the heights (solid iff Y <= h)
    for (int i = tid; i < 128; i += kThreads) {       // Y and Z halo rows: word over x
        const int face = i >> 5, r = i & 31;          // 0:-Y 1:+Y 2:-Z 3:+Z
        int zz, Y, dst;
        if (face == 0) { zz = r; Y = Y0 - 1; dst = (r + 1) * 34 + 0; }
        else if (face == 1) { zz = r; Y = Y0 + 32; dst = (r + 1) * 34 + 33; }
        else if (face == 2) { zz = -1; Y = Y0 + r; dst = 0 * 34 + (r + 1); }
        else { zz = 32; Y = Y0 + r; dst = 33 * 34 + (r + 1); }
        const short* hp = s_h + (zz + 1) * 34 + 1;
        uint32_t bits = 0;
#pragma unroll 8
        for (int x = 0; x < 32; ++x) bits |= (uint32_t)(Y <= hp[x]) << x;
        sm.occ[dst] = bits;
    }
    if (tid < 64) {                                   // X halos: word over y for each z
        const int face = tid >> 5, z = tid & 31;
        const int h = s_h[(z + 1) * 34 + (face ? 33 : 0)];
        uint32_t bits = 0;
#pragma unroll 8
        for (int y = 0; y < 32; ++y) bits |= (uint32_t)(Y0 + y <= h) << y;
        sm.xh[face][z] = bits;
    }
    __syncthreads();
    prof.mark(0);

    uint32_t v_or = 0, v_and = kFull, o_and = kFull;
    v4::IdPlanes ip;
#pragma unroll
    for (int s = 0; s < v4::kSlabs; ++s) v4::convert_slab(sm, s, v_or, v_and, o_and, ip);
    const v4::ChunkFlags fl = v4::reduce_flags(sm, v_or, v_and, o_and);   // barrier: the tile has been read
    prof.mark(1);
    bool with_planes = false;
    if constexpr (kGreedy) {
        with_planes = !fl.single_id && fl.small_ids;     // terrain ids are 1..3: always, for a chunk with several ids
        if (with_planes) v4::store_id_planes(sm.tile.post, ip);
    }
    const short* hcol = s_h;
    v4::mesh_tail<kGreedy>(sm, chunk, fl, with_planes, out, [hcol, Y0](int idx) -> uint32_t {
        const int x = idx & 31, y = (idx >> 5) & 31, z = idx >> 10;
        return terrain_block(hcol[(z + 1) * 34 + x + 1], Y0 + y);
    }, prof);
}

template <bool kGreedy>
__global__ void __launch_bounds__(kThreads, 3)
terrain_mesh_kernel(const vxp_coord* __restrict__ coords, uint32_t n, ColumnScratch cs, v4::MeshOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    v4::Smem<kGreedy>& sm = *reinterpret_cast<v4::Smem<kGreedy>*>(smem_raw);
    __shared__ short s_h[34 * 34];   // heights of the column footprint: s_h[(z+1)*34 + (x+1)]  (|h-256| <= 720)
    __shared__ uint32_t s_item[2];
    const uint32_t nwork = cs.count[1];
    if (threadIdx.x == 0) s_item[0] = atomicAdd(cs.count + 2, 1u);
    __syncthreads();
    for (uint32_t it = 0;; ++it) {                    // work items are drawn from a ticket counter: chunk costs vary widely
        const uint32_t item = s_item[it & 1];
        if (item >= nwork) break;
        if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(cs.count + 2, 1u);   // its latency hides behind this chunk
        terrain_mesh_chunk(sm, s_h, coords, cs, cs.slot_of_chunk[n + item], out);
        __syncthreads();                              // shared memory is free again, the next ticket is visible
    }
}

// ---- generation 5 of the fused greedy kernel (vxp_mesh5.cuh: four CTAs per SM): the chunk is generated one
// 16 KiB slab at a time into the ring and converted from there, the rest is v5's pipeline.
__device__ __forceinline__ void terrain_mesh_chunk_g(v5::SmemG& sm, short* s_h, const vxp_coord* __restrict__ coords,
                                                     const ColumnScratch& cs, uint32_t chunk, const v4::MeshOut& out) {
    const int tid = threadIdx.x;
    const vxp_coord c = coords[chunk];
    v4::Prof prof;
    prof.start(false);
    const short* col = cs.heights + (size_t)cs.id_of_slot[cs.slot_of_chunk[chunk]] * kColStride;
    const int Y0 = 32 * c.y;
    if (tid == 0) { sm.ones = kFull; sm.zero = 0u; }
    for (int i = tid; i < 34 * 34 / 2; i += kThreads)
        reinterpret_cast<uint32_t*>(s_h)[i] = __ldg(reinterpret_cast<const uint32_t*>(col) + i);
    __syncthreads();
    // halo occupancy straight from the heights (solid iff Y <= h)
    for (int i = tid; i < 128; i += kThreads) {       // Y and Z halo rows: word over x
        const int face = i >> 5, r = i & 31;          // 0:-Y 1:+Y 2:-Z 3:+Z
        int zz, Y, dst;
        if (face == 0) { zz = r; Y = Y0 - 1; dst = (r + 1) * 34 + 0; }
        else if (face == 1) { zz = r; Y = Y0 + 32; dst = (r + 1) * 34 + 33; }
        else if (face == 2) { zz = -1; Y = Y0 + r; dst = 0 * 34 + (r + 1); }
        else { zz = 32; Y = Y0 + r; dst = 33 * 34 + (r + 1); }
        const short* hp = s_h + (zz + 1) * 34 + 1;
        uint32_t bits = 0;
#pragma unroll 8
        for (int x = 0; x < 32; ++x) bits |= (uint32_t)(Y <= hp[x]) << x;
        sm.occ[dst] = bits;
    }
    if (tid < 64) {                                   // X halos: word over y for each z
        const int face = tid >> 5, z = tid & 31;
        const int h = s_h[(z + 1) * 34 + (face ? 33 : 0)];
        uint32_t bits = 0;
#pragma unroll 8
        for (int y = 0; y < 32; ++y) bits |= (uint32_t)(Y0 + y <= h) << y;
        sm.xh[face][z] = bits;
    }
    uint32_t v_or = 0, v_and = kFull, o_and = kFull;
    v4::IdPlanes ip;
#pragma unroll
    for (int s = 0; s < v4::kSlabs; ++s) {
        uint4* slab4 = reinterpret_cast<uint4*>(sm.tile.ring[s & 1]);
#pragma unroll 2
        for (int k = 0; k < 4; ++k) {
            const int j = s * 1024 + tid + k * kThreads;      // uint4 index inside the chunk: (z*32 + y)*4 + piece
            const int z = j >> 7, Y = Y0 + ((j >> 2) & 31), x0 = (j & 3) * 8;
            const short* hp = s_h + (z + 1) * 34 + x0 + 1;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = terrain_block(hp[2 * e], Y) | terrain_block(hp[2 * e + 1], Y) << 16;
            slab4[tid + k * kThreads] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncthreads();                              // slab s complete (and slab s-1 converted by everybody: its buffer is next)
        v5::convert_ring_slab(sm, s, v_or, v_and, o_and, ip);
    }
    const v4::ChunkFlags fl = v5::reduce_flags_g(sm, v_or, v_and, o_and);   // barriers: occupancy and halo published, ring dead
    const bool with_eq = !fl.single_id;               // terrain ids are 1..4: a chunk with several ids always takes the plane path
    if (with_eq) v5::store_id_planes_g(sm, ip);
    if (tid < 192) sm.rows[tid] = 0u;
    __syncthreads();
    const bool any = __syncthreads_or(v5::build_masks_g(sm, with_eq, with_eq) != 0);
    if (!any) {
        if (tid == 0) out.meta[chunk] = vxp_chunk_meta{0u, 0u};
        return;
    }
    const short* hcol = s_h;
    v5::emit_greedy(sm, chunk, with_eq, out, [hcol, Y0](int idx) -> uint32_t {
        const int x = idx & 31, y = (idx >> 5) & 31, z = idx >> 10;
        return terrain_block(hcol[(z + 1) * 34 + x + 1], Y0 + y);
    }, prof);
}

__global__ void __launch_bounds__(kThreads, 4)
terrain_mesh_kernel_g(const vxp_coord* __restrict__ coords, uint32_t n, ColumnScratch cs, v4::MeshOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    v5::SmemG& sm = *reinterpret_cast<v5::SmemG*>(smem_raw);
    __shared__ short s_h[34 * 34];
    __shared__ uint32_t s_item[2];
    const uint32_t nwork = cs.count[1];
    if (threadIdx.x == 0) s_item[0] = atomicAdd(cs.count + 2, 1u);
    __syncthreads();
    for (uint32_t it = 0;; ++it) {
        const uint32_t item = s_item[it & 1];
        if (item >= nwork) break;
        if (threadIdx.x == 0) s_item[(it + 1) & 1] = atomicAdd(cs.count + 2, 1u);
        terrain_mesh_chunk_g(sm, s_h, coords, cs, cs.slot_of_chunk[n + item], out);
        __syncthreads();
    }
}

// ====================================================================== launchers

template <class K>
static cudaError_t set_smem(K fn, size_t bytes) {
    return cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t configure_device() {
    cudaError_t e = set_smem(v4::mesh_kernel<false>, sizeof(v4::Smem<false>));
    if (e == cudaSuccess) e = set_smem(v4::mesh_kernel<true>, sizeof(v4::Smem<true>));
    if (e == cudaSuccess) e = set_smem(v5::mesh_kernel_g, sizeof(v5::SmemG));
    if (e == cudaSuccess) e = set_smem(v5::mesh_kernel_c, sizeof(v5::SmemC));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel<false>, sizeof(v4::Smem<false>));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel<true>, sizeof(v4::Smem<true>));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel_g, sizeof(v5::SmemG));
    return e;
}

cudaError_t launch_terrain_fill(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    terrain_fill_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, seed, d_voxels);
    return cudaGetLastError();
}

cudaError_t launch_terrain_fill_columns(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                        const ColumnScratch& cs, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(cs.keys, 0xFF, ((size_t)cs.mask + 1) * sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cs.count, 0, 3 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    column_insert_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_coords, n, cs);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t hgrid = n < (uint32_t)(4 * sms) ? n : (uint32_t)(4 * sms);
    column_heights_kernel<<<hgrid, kThreads, 0, stream>>>(cs, seed);
    terrain_fill_columns_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, cs, d_voxels);
    return cudaGetLastError();
}

cudaError_t launch_pattern_fill(const vxp_coord* d_coords, uint32_t n, int pattern, uint16_t* d_voxels,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    pattern_fill_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, pattern, d_voxels);
    return cudaGetLastError();
}

static v4::MeshOut mesh_out(const vxp_mesh_dev& o) {
    return v4::MeshOut{reinterpret_cast<unsigned long long*>(o.verts), o.indices, o.meta,
                       reinterpret_cast<unsigned long long*>(o.total_quads), o.capacity_quads, 0ull};
}

// Greedy batches run generation 5 (four CTAs per SM, vxp_mesh5.cuh); VXP_GREEDY_V4=1 selects generation 4 for A/B runs.
static bool greedy_v4() {
    static const bool v = getenv("VXP_GREEDY_V4") != nullptr;
    return v;
}
static bool culled_v4() {
    static const bool v = getenv("VXP_CULLED_V4") != nullptr;
    return v;
}
template <bool kGreedy>
static cudaError_t launch_mesh_kernel(uint32_t grid, int overlap, cudaStream_t stream, const uint16_t* d_voxels,
                                      const int32_t* d_nbr6, uint32_t n, uint32_t lo, uint32_t hi, v4::MeshOut out, v4::Bnd bnd,
                                      v4::Chain chain) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    const bool g5 = kGreedy && !greedy_v4(), c5 = !kGreedy && !culled_v4();
    cfg.dynamicSmemBytes = g5 ? sizeof(v5::SmemG) : c5 ? sizeof(v5::SmemC) : sizeof(v4::Smem<kGreedy>);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    if (g5) return cudaLaunchKernelEx(&cfg, v5::mesh_kernel_g, d_voxels, d_nbr6, n, lo, hi, out, bnd, chain);
    if (c5) return cudaLaunchKernelEx(&cfg, v5::mesh_kernel_c, d_voxels, d_nbr6, n, lo, hi, out, bnd, chain);
    return cudaLaunchKernelEx(&cfg, v4::mesh_kernel<kGreedy>, d_voxels, d_nbr6, n, lo, hi, out, bnd, chain);
}

cudaError_t launch_mesh_range(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, uint32_t lo, uint32_t hi,
                              int greedy, const vxp_mesh_dev& o, const BoundaryScratch& scratch, cudaStream_t stream) {
    if (hi > n) hi = n;
    if (lo >= hi) return cudaSuccess;
    const v4::Bnd bnd{scratch.tags, scratch.planes, scratch.epoch};
    const uint32_t grid = hi - lo;
    const v4::Chain chain{nullptr, nullptr, nullptr, 0u, 0u};
    return greedy ? launch_mesh_kernel<true>(grid, 0, stream, d_voxels, d_nbr6, n, lo, hi, mesh_out(o), bnd, chain)
                  : launch_mesh_kernel<false>(grid, 0, stream, d_voxels, d_nbr6, n, lo, hi, mesh_out(o), bnd, chain);
}

cudaError_t launch_mesh(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, int greedy,
                        const vxp_mesh_dev& o, const BoundaryScratch& scratch, unsigned long long* counter, int overlap,
                        cudaStream_t stream) {
    if (n == 0) return cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    if (n > (1u << 22) - 1u) return cudaErrorInvalidValue;   // CTA count field of the launch counter
    const v4::Bnd bnd{scratch.tags, scratch.planes, scratch.epoch};
    v4::MeshOut out = mesh_out(o);
    out.total = counter;
    out.cta_inc = v4::kCtaOne;
    // deal the runs of 16 chunks out with a small stride coprime to their number (see v4::Chain).  Measured on
    // config 2 / 3 (256 runs): identity 75.7 / 82.4 us, 17 -> 69.5 / 75.4, 5 -> 70.4 / 75.7, 33 -> 70.7 / 76.0,
    // 97 -> 72.3 / 77.0, 159 (0.618 R) -> 71.9 / 76.6: small strides keep +-Y / +-Z neighbours close enough in
    // time for their halo rows to hit L2.
    // (the two environment variables are for tuning runs only and are read once)
    static const bool no_deal = getenv("VXP_NO_DEAL") != nullptr;
    static const uint32_t stride_override = getenv("VXP_DEAL_STRIDE") ? (uint32_t)atoi(getenv("VXP_DEAL_STRIDE")) | 1u : 0u;
    uint32_t runs = n >> 4, stride = 0;
    if (runs > 34 && !no_deal) {
        auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
        stride = stride_override ? stride_override % runs : 17u;
        while (gcd(stride, runs) != 1u) stride += 2u;
    } else {
        runs = 0;
    }
    const v4::Chain chain{reinterpret_cast<unsigned long long*>(o.total_quads), nullptr, nullptr, runs, stride};
    return greedy ? launch_mesh_kernel<true>(n, overlap, stream, d_voxels, d_nbr6, n, 0, n, out, bnd, chain)
                  : launch_mesh_kernel<false>(n, overlap, stream, d_voxels, d_nbr6, n, 0, n, out, bnd, chain);
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
    const v4::Bnd bnd{nullptr, nullptr, 0};      // the x neighbours are not meshed alongside: their boundary is gathered
    v4::MeshOut out = mesh_out(o);
    out.total = counter;
    out.cta_inc = v4::kCtaOne;
    const v4::Chain chain{reinterpret_cast<unsigned long long*>(o.total_quads), d_list, d_count, 0u, 0u};
    return greedy ? launch_mesh_kernel<true>(max_count, 0, stream, d_voxels, d_nbr6, n, 0, n, out, bnd, chain)
                  : launch_mesh_kernel<false>(max_count, 0, stream, d_voxels, d_nbr6, n, 0, n, out, bnd, chain);
}

size_t boundary_tag_bytes(uint32_t n) { return (size_t)n * sizeof(uint32_t); }
size_t boundary_plane_bytes(uint32_t n) { return (size_t)n * 64 * sizeof(unsigned long long); }

// diagnostics (not part of include/vxp.h): resident CTAs per SM of a mesh kernel on the current device
extern "C" int vxp_debug_occupancy(int greedy, int fused) {
    int n = -1;
    cudaError_t e;
    if (fused)
        e = greedy ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, terrain_mesh_kernel<true>, kThreads, sizeof(v4::Smem<true>))
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, terrain_mesh_kernel<false>, kThreads, sizeof(v4::Smem<false>));
    else
        e = greedy ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v5::mesh_kernel_g, kThreads, sizeof(v5::SmemG))
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v5::mesh_kernel_c, kThreads, sizeof(v5::SmemC));
    return e == cudaSuccess ? n : -1;
}
extern "C" int vxp_debug_smem_bytes(int greedy) { return greedy ? (int)sizeof(v5::SmemG) : (int)sizeof(v5::SmemC); }

#ifdef VXP_PROFILE
extern "C" int vxp_debug_read_prof_time(unsigned long long* out, int max_ctas) {
    const int ctas = max_ctas < v4::kProfCtas ? max_ctas : v4::kProfCtas;
    return cudaMemcpyFromSymbol(out, v4::g_prof_time, sizeof(unsigned long long) * 2 * (size_t)ctas) == cudaSuccess ? 2 : -1;
}
extern "C" int vxp_debug_read_prof_warp(unsigned int* out, int max_ctas) {
    const int ctas = max_ctas < v4::kProfCtas ? max_ctas : v4::kProfCtas;
    return cudaMemcpyFromSymbol(out, v4::g_prof_warp, sizeof(unsigned int) * 8 * (size_t)ctas) == cudaSuccess ? 8 : -1;
}
// profiling build only: read (and clear) the per-CTA phase counters of the mesh kernels
extern "C" int vxp_debug_read_prof(unsigned long long* out, int max_ctas) {
    const int ctas = max_ctas < v4::kProfCtas ? max_ctas : v4::kProfCtas;
    const size_t bytes = sizeof(unsigned long long) * v4::kProfSlots * (size_t)ctas;
    if (cudaMemcpyFromSymbol(out, v4::g_prof, bytes) != cudaSuccess) return -1;
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, v4::g_prof) != cudaSuccess) return -2;
    if (cudaMemset(sym, 0, sizeof(unsigned long long) * v4::kProfSlots * v4::kProfCtas) != cudaSuccess) return -3;
    return v4::kProfSlots;
}
#endif

size_t column_table_slots(uint32_t n) {
    size_t t = 64;
    while (t < 2 * (size_t)n) t <<= 1;
    return t;
}
size_t column_record_bytes() { return kColStride * sizeof(short); }

cudaError_t launch_terrain_mesh(const vxp_coord* d_coords, uint32_t n, uint64_t seed, int greedy,
                                const vxp_mesh_dev& o, const ColumnScratch& cs, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    if (e != cudaSuccess || n == 0) return e;
    e = cudaMemsetAsync(cs.keys, 0xFF, ((size_t)cs.mask + 1) * sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cs.count, 0, 3 * sizeof(uint32_t), stream);   // columns, work items, ticket
    if (e != cudaSuccess) return e;
    column_insert_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_coords, n, cs);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t hgrid = n < (uint32_t)(4 * sms) ? n : (uint32_t)(4 * sms);
    column_heights_kernel<<<hgrid, kThreads, 0, stream>>>(cs, seed);
    column_classify_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, stream>>>(d_coords, n, cs, o.meta);
    // the work list's length is only known on the device: a resident grid strides over it
    // generation 5 (four CTAs per SM) pays off when every CTA draws many work items: world of 274 625 chunks 881 -> 850 us,
    // but a 4096-chunk batch (1540 work items over 592 CTAs instead of 444) 94 -> 113 us
    const bool g5 = greedy && !greedy_v4() && (n >= 32768u || getenv("VXP_FUSED_V5") != nullptr);
    const uint32_t per_sm = g5 ? 4u : 3u;
    const uint32_t mgrid = n < per_sm * (uint32_t)sms ? n : per_sm * (uint32_t)sms;
    if (g5)
        terrain_mesh_kernel_g<<<mgrid, kThreads, sizeof(v5::SmemG), stream>>>(d_coords, n, cs, mesh_out(o));
    else if (greedy)
        terrain_mesh_kernel<true><<<mgrid, kThreads, sizeof(v4::Smem<true>), stream>>>(d_coords, n, cs, mesh_out(o));
    else
        terrain_mesh_kernel<false><<<mgrid, kThreads, sizeof(v4::Smem<false>), stream>>>(d_coords, n, cs, mesh_out(o));
    return cudaGetLastError();
}

} // namespace vxp
