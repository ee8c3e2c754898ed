// vxp_kernels.cu — sm_100a kernels: terrain fill, pattern fill, stored-chunk meshing,
// fused terrain+mesh.  SPEC.md v1.  Host launchers at the bottom (C++ linkage, used by vxp_api.cu).
#include <climits>

#include <cstdlib>
#include <cstring>

#include "vxp_launch.h"
#include "vxp_mesh4.cuh"

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
// The tile is generated in shared memory from the height field; voxels never reach HBM and the
// halo comes straight from the heights of the 34x34 column footprint.
template <bool kGreedy>
__global__ void __launch_bounds__(kThreads, 3)
terrain_mesh_kernel(const vxp_coord* __restrict__ coords, uint32_t n, uint64_t seed, v4::MeshOut out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    v4::Smem<kGreedy>& sm = *reinterpret_cast<v4::Smem<kGreedy>*>(smem_raw);
    __shared__ short s_h[34 * 34];   // heights of the column footprint: s_h[(z+1)*34 + (x+1)]  (|h-256| <= 720)
    __shared__ int s_red[2 * kWarps];
    __shared__ TerrainLattice s_lat;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n) return;
    const vxp_coord c = coords[chunk];
    const TerrainKeys tk = terrain_keys(seed);
    v4::Prof prof;
    prof.start(tid == 0);
    if (tid == 0) sm.ones = kFull;
    terrain_lattice_fill(s_lat, tk, 32 * c.x - 1, 32 * c.z - 1);
    __syncthreads();

    int hmin = INT_MAX, hmax = INT_MIN;
    for (int i = tid; i < 34 * 34; i += kThreads) {
        const int zz = i / 34 - 1, xx = i % 34 - 1;
        const bool corner = (zz < 0 || zz > 31) && (xx < 0 || xx > 31);
        int h = 0;
        if (!corner) {
            h = terrain_height_lattice(s_lat, 32 * c.x + xx, 32 * c.z + zz);
            hmin = min(hmin, h);
            hmax = max(hmax, h);
        }
        s_h[i] = (short)h;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        hmin = min(hmin, __shfl_xor_sync(kFull, hmin, d));
        hmax = max(hmax, __shfl_xor_sync(kFull, hmax, d));
    }
    if (lane == 0) { s_red[warp] = hmin; s_red[kWarps + warp] = hmax; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { hmin = min(hmin, s_red[w]); hmax = max(hmax, s_red[kWarps + w]); }

    const int Y0 = 32 * c.y;
    // all air (incl. the layer below it) or all solid incl. every halo layer: no visible face
    if (Y0 > hmax || Y0 + 32 <= hmin) {
        if (tid == 0) out.meta[chunk] = vxp_chunk_meta{0u, 0u};
        return;
    }
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
#pragma unroll 1
    for (int s = 0; s < v4::kSlabs; ++s) v4::convert_slab(sm, s, v_or, v_and, o_and);
    const v4::ChunkFlags fl = v4::reduce_flags(sm, v_or, v_and, o_and);
    prof.mark(1);
    const short* hcol = s_h;
    v4::mesh_tail<kGreedy>(sm, chunk, fl, out, [hcol, Y0](int idx) -> uint32_t {
        const int x = idx & 31, y = (idx >> 5) & 31, z = idx >> 10;
        return terrain_block(hcol[(z + 1) * 34 + x + 1], Y0 + y);
    }, prof);
}

// ====================================================================== launchers

template <class K>
static cudaError_t set_smem(K fn, size_t bytes) {
    return cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t configure_device() {
    cudaError_t e = set_smem(v4::mesh_kernel<false>, sizeof(v4::Smem<false>));
    if (e == cudaSuccess) e = set_smem(v4::mesh_kernel<true>, sizeof(v4::Smem<true>));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel<false>, sizeof(v4::Smem<false>));
    if (e == cudaSuccess) e = set_smem(terrain_mesh_kernel<true>, sizeof(v4::Smem<true>));
    return e;
}

cudaError_t launch_terrain_fill(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    terrain_fill_kernel<<<n, kThreads, 0, stream>>>(d_coords, n, seed, d_voxels);
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
                       reinterpret_cast<unsigned long long*>(o.total_quads), o.capacity_quads};
}

cudaError_t launch_mesh_range(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, uint32_t lo, uint32_t hi,
                              int greedy, const vxp_mesh_dev& o, const BoundaryScratch& scratch, cudaStream_t stream) {
    if (hi > n) hi = n;
    if (lo >= hi) return cudaSuccess;
    const v4::Bnd bnd{scratch.tags, scratch.planes, scratch.epoch};
    if (greedy)
        v4::mesh_kernel<true><<<hi - lo, kThreads, sizeof(v4::Smem<true>), stream>>>(d_voxels, d_nbr6, n, lo, mesh_out(o), bnd);
    else
        v4::mesh_kernel<false><<<hi - lo, kThreads, sizeof(v4::Smem<false>), stream>>>(d_voxels, d_nbr6, n, lo, mesh_out(o), bnd);
    return cudaGetLastError();
}

cudaError_t launch_mesh(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, int greedy,
                        const vxp_mesh_dev& o, const BoundaryScratch& scratch, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    if (e != cudaSuccess || n == 0) return e;
    return launch_mesh_range(d_voxels, d_nbr6, n, 0, n, greedy, o, scratch, stream);
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
        e = greedy ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v4::mesh_kernel<true>, kThreads, sizeof(v4::Smem<true>))
                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v4::mesh_kernel<false>, kThreads, sizeof(v4::Smem<false>));
    return e == cudaSuccess ? n : -1;
}
extern "C" int vxp_debug_smem_bytes(int greedy) { return greedy ? (int)sizeof(v4::Smem<true>) : (int)sizeof(v4::Smem<false>); }

#ifdef VXP_PROFILE
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

cudaError_t launch_terrain_mesh(const vxp_coord* d_coords, uint32_t n, uint64_t seed, int greedy,
                                const vxp_mesh_dev& o, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(o.total_quads, 0, sizeof(uint64_t), stream);
    if (e != cudaSuccess || n == 0) return e;
    if (greedy)
        terrain_mesh_kernel<true><<<n, kThreads, sizeof(v4::Smem<true>), stream>>>(d_coords, n, seed, mesh_out(o));
    else
        terrain_mesh_kernel<false><<<n, kThreads, sizeof(v4::Smem<false>), stream>>>(d_coords, n, seed, mesh_out(o));
    return cudaGetLastError();
}

} // namespace vxp
