// vxp_mesh.cuh — per-chunk meshing pipeline in shared memory (SPEC §2-§5).
//
// One CTA owns one chunk at a time:
//   tile (64 KiB, filled by a bulk TMA or by the fused terrain generator)
//     -> occupancy bit volume (+ 1-voxel halo from the six neighbours)
//     -> id-equality bit volumes EX/EY/EZ (greedy only, skipped for uniform chunks)
//     -> six face-mask sets in plane orientation (X planes via warp bit-transposes)
//     -> culled: popcount + block scan;  greedy: bitwise row sweep, one thread per plane
//     -> 32-bit quad records staged in shared memory (re-using the dead tile)
//     -> coalesced 16-byte vertex / 8-byte index stores.
#pragma once
#include "vxp_device.cuh"
#include "../../include/vxp.h"

namespace vxp {

constexpr int kPlaneWords = 32 * kPad;                          // 1056 words per padded 32x32 plane set
constexpr int kStageCap = (kTileBytes - 2 * kPlaneWords * 4) / 4; // 14272 staged quad records

struct __align__(128) MeshSmem {
    union {
        uint16_t vox[VXP_CHUNK_VOXELS]; // [z][y][x], TMA destination
        struct {
            uint32_t ht[kPlaneWords];    // X planes: ht[x*33+z] bit y = (id(x,y,z) == id(x,y-1,z))
            uint32_t vt[kPlaneWords];    // X planes: vt[x*33+z] bit y = (id(x,y,z) == id(x,y,z-1))
            uint32_t staging[kStageCap]; // quad records awaiting expansion
        } post;
    } tile;
    uint32_t occ[34 * 34];      // occ[(z+1)*34 + (y+1)], bit x; rows/layers -1 and 32 are the halo
    uint32_t xh[2][32];         // xh[0][z] bit y: voxel (-1,y,z) solid;  xh[1][z] bit y: voxel (32,y,z)
    uint32_t ex[kPlaneWords];   // ex[z*33+y] bit x = (id(x,y,z) == id(x-1,y,z))   (bit 0 unused)
    uint32_t ey[kPlaneWords];   // ey[z*33+y] bit x = (id(x,y,z) == id(x,y-1,z))   (row y=0 unused)
    uint32_t ez[kPlaneWords];   // ez[z*33+y] bit x = (id(x,y,z) == id(x,y,z-1))   (layer z=0 unused)
    uint32_t mask[192 * kPad];  // mask[(f*32+s)*33 + v] bit u: face f of cell (u,v) in slice s visible
    uint32_t scratch[kWarps + 1];
    uint32_t cursor;            // staging fill level
    uint32_t ones;              // 0xFFFFFFFF: stand-in equality row for uniform chunks
    unsigned long long mbar;    // TMA completion barrier
};

struct MeshOut {
    unsigned long long* verts;
    uint32_t* indices;
    vxp_chunk_meta* meta;
    unsigned long long* total;
    unsigned long long cap;
};

// ------------------------------------------------------------------ halo
// Occupancy of the six neighbour boundary layers, read straight from global
// memory (L2-resident: the neighbour is some other CTA's tile).
__device__ __forceinline__ void load_halo(MeshSmem& sm, const uint16_t* __restrict__ voxels,
                                          const int32_t* __restrict__ nbr6, uint32_t chunk) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int nb[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) nb[f] = nbr6 ? __ldg(nbr6 + 6 * (size_t)chunk + f) : -1;

    // Y and Z halos: 4 faces x 32 rows x 4 pieces of 8 voxels
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int j = tid + k * kThreads;
        const int face = 2 + (j >> 7), r = (j >> 2) & 31, p = j & 3;
        int n, src, dst;
        if (face == 2) { n = nb[2]; src = (r * 32 + 31) * 32; dst = (r + 1) * 34 + 0; }       // y=-1 <- nbr y=31
        else if (face == 3) { n = nb[3]; src = (r * 32 + 0) * 32; dst = (r + 1) * 34 + 33; }  // y=32 <- nbr y=0
        else if (face == 4) { n = nb[4]; src = (31 * 32 + r) * 32; dst = 0 * 34 + (r + 1); }  // z=-1 <- nbr z=31
        else { n = nb[5]; src = (0 * 32 + r) * 32; dst = 33 * 34 + (r + 1); }                 // z=32 <- nbr z=0
        uint32_t bits = 0;
        if (n >= 0) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(voxels + (size_t)n * VXP_CHUNK_VOXELS + src) + p);
            bits = nonzero8(q) << (8 * p);
        }
        bits |= __shfl_xor_sync(kFull, bits, 1);
        bits |= __shfl_xor_sync(kFull, bits, 2);
        if (p == 0) sm.occ[dst] = bits;
    }
    // X halos: 2 faces x 32 layers, lane = y
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int item = warp * 8 + k;
        const int face = item >> 5, z = item & 31;
        const int n = nb[face];
        uint32_t v = 0;
        if (n >= 0) {
            const int x = face ? 0 : 31;
            v = __ldg(voxels + (size_t)n * VXP_CHUNK_VOXELS + x + 32 * lane + 1024 * z);
        }
        const uint32_t word = __ballot_sync(kFull, v != 0);
        if (lane == 0) sm.xh[face][z] = word;
    }
}

// ------------------------------------------------------------------ tile -> occupancy
// Returns (block-uniform) true iff all 32768 voxels carry the same id.
__device__ __forceinline__ bool build_occupancy(MeshSmem& sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = lane & 3;
    const uint32_t v0 = sm.tile.vox[0];
    const uint32_t first = v0 | v0 << 16;
    uint32_t diff = 0;
    const uint4* tile4 = reinterpret_cast<const uint4*>(sm.tile.vox);
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
        const int rg = it * kWarps + warp;           // row group: 8 consecutive rows
        const int row = rg * 8 + (lane >> 2);        // row = z*32 + y
        const uint4 q = tile4[row * 4 + p];          // warp reads 512 contiguous bytes
        diff |= (q.x ^ first) | (q.y ^ first) | (q.z ^ first) | (q.w ^ first);
        uint32_t bits = nonzero8(q) << (8 * p);
        bits |= __shfl_xor_sync(kFull, bits, 1);
        bits |= __shfl_xor_sync(kFull, bits, 2);
        if (p == 0) sm.occ[((row >> 5) + 1) * 34 + (row & 31) + 1] = bits;
    }
    return __syncthreads_or(diff != 0) == 0;
}

// ------------------------------------------------------------------ tile -> EX / EY / EZ
__device__ __forceinline__ void build_equality(MeshSmem& sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = lane & 3;
    const uint4* tile4 = reinterpret_cast<const uint4*>(sm.tile.vox);
    // 4x4 byte transpose across the 4 lanes of a row: selectors per lane
    const uint32_t selA = (p & 1) ? 0x3715u : 0x6240u;
    const uint32_t selB = (p & 2) ? 0x3276u : 0x5410u;
    uint32_t* const dstvol = (p == 0) ? sm.ex : (p == 1) ? sm.ey : sm.ez;
#pragma unroll 2
    for (int it = 0; it < 16; ++it) {
        const int rg = it * kWarps + warp;
        const int row = rg * 8 + (lane >> 2);
        const int z = row >> 5, y = row & 31;
        const uint4 q = tile4[row * 4 + p];
        // x-1 neighbour: shift the u16 stream by one element, pulling from the lane on the left
        const uint32_t leftw = __shfl_up_sync(kFull, q.w, 1);
        const uint4 px = make_uint4(prev16(leftw, q.x), prev16(q.x, q.y), prev16(q.y, q.z), prev16(q.z, q.w));
        uint32_t w = (~nonzero8(xor4(q, px))) & 0xFFu;
        if (y > 0) w |= ((~nonzero8(xor4(q, tile4[(row - 1) * 4 + p]))) & 0xFFu) << 8;
        if (z > 0) w |= ((~nonzero8(xor4(q, tile4[(row - 32) * 4 + p]))) & 0xFFu) << 16;
        // lanes p=0..3 hold [ex8,ey8,ez8,-] for x-pieces 0..3  ->  lane 0: ex word, 1: ey, 2: ez
        uint32_t t = __shfl_xor_sync(kFull, w, 1);
        w = __byte_perm(w, t, selA);
        t = __shfl_xor_sync(kFull, w, 2);
        w = __byte_perm(w, t, selB);
        if (p < 3) dstvol[z * kPad + y] = w;
    }
}

// ------------------------------------------------------------------ occupancy -> face masks
template <bool kGreedy>
__device__ __forceinline__ void build_masks(MeshSmem& sm, bool uniform) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int z = it * kWarps + warp, y = lane;
        const uint32_t* o = sm.occ + (z + 1) * 34 + (y + 1);
        const uint32_t c = o[0];
        sm.mask[(2 * 32 + y) * kPad + z] = c & ~o[-1];   // -Y : slice y, row z
        sm.mask[(3 * 32 + y) * kPad + z] = c & ~o[+1];   // +Y
        sm.mask[(4 * 32 + z) * kPad + y] = c & ~o[-34];  // -Z : slice z, row y
        sm.mask[(5 * 32 + z) * kPad + y] = c & ~o[+34];  // +Z
        const uint32_t lo = (sm.xh[0][z] >> y) & 1u, hi = (sm.xh[1][z] >> y) & 1u;
        uint32_t vxm = c & ~((c << 1) | lo);             // -X visible, bit x
        uint32_t vxp_ = c & ~((c >> 1) | (hi << 31));    // +X visible, bit x
        // transpose (lane=y, bit=x) -> (lane=x, bit=y): X planes have u=y, v=z, slice x
        if (__any_sync(kFull, vxm != 0)) vxm = transpose32(vxm, lane);
        if (__any_sync(kFull, vxp_ != 0)) vxp_ = transpose32(vxp_, lane);
        sm.mask[(0 * 32 + lane) * kPad + z] = vxm;
        sm.mask[(1 * 32 + lane) * kPad + z] = vxp_;
        if (kGreedy && !uniform) {
            sm.tile.post.ht[lane * kPad + z] = transpose32(sm.ey[z * kPad + y], lane);
            sm.tile.post.vt[lane * kPad + z] = transpose32(sm.ez[z * kPad + y], lane);
        }
    }
}

// ------------------------------------------------------------------ greedy sweep (one thread = one plane)
struct Sweep {
    uint32_t* m;         // this plane's 32 mask rows (consumed in place)
    const uint32_t* hb;  // equality rows along u  (bit u: cell u has the id of cell u-1)
    const uint32_t* vb;  // equality rows along v  (bit u: cell (u,v) has the id of (u,v-1))
    int hs, vs;          // row strides of hb / vb
    uint32_t R;          // remaining bits of the current row
    uint32_t pending;    // quad produced but not yet staged
    uint32_t base;       // f<<15 | s<<10
    int v;
    bool has_pending, done;
};

__device__ __forceinline__ void sweep_init(Sweep& st, MeshSmem& sm, bool uniform) {
    const int tid = threadIdx.x;
    st.done = tid >= 192;
    st.has_pending = false;
    st.pending = 0;
    st.v = 0;
    st.R = 0;
    st.m = sm.mask;
    st.hb = st.vb = &sm.ones;
    st.hs = st.vs = 0;
    st.base = 0;
    if (st.done) return;
    const int f = tid >> 5, s = tid & 31;
    st.base = (uint32_t)f << 15 | (uint32_t)s << 10;
    st.m = sm.mask + tid * kPad;
    st.R = st.m[0];
    if (!uniform) {
        if (f < 2) {            // X planes: u=y, v=z
            st.hb = sm.tile.post.ht + s * kPad; st.hs = 1;
            st.vb = sm.tile.post.vt + s * kPad; st.vs = 1;
        } else if (f < 4) {     // Y planes: u=x, v=z, slice y=s
            st.hb = sm.ex + s; st.hs = kPad;
            st.vb = sm.ez + s; st.vs = kPad;
        } else {                // Z planes: u=x, v=y, slice z=s
            st.hb = sm.ex + s * kPad; st.hs = 1;
            st.vb = sm.ey + s * kPad; st.vs = 1;
        }
    }
}

// Advance to the next quad of this plane (SPEC §4 greedy rule); false when the plane is exhausted.
__device__ __forceinline__ bool sweep_next(Sweep& st, uint32_t& rec) {
    while (st.R == 0) {
        if (++st.v >= 32) { st.done = true; return false; }
        st.R = st.m[st.v];
    }
    const int v = st.v;
    const int u = __ffs(st.R) - 1;
    const uint32_t t = st.R >> u;                        // bit 0 set
    const uint32_t nt = ~t;
    const int ones = nt ? (__ffs(nt) - 1) : 32;          // run of set cells from u
    // id boundaries to the right of u: bit k of bnd = cell u+1+k differs from its left neighbour
    const uint32_t eqrow = st.hb[v * st.hs];
    const uint32_t bnd = (u == 31) ? 0u : ((~eqrow) >> (u + 1));
    const int lim = bnd ? __ffs(bnd) : 33;               // width allowed by ids
    const int w = min(ones, lim);
    const uint32_t run = (w == 32 ? kFull : ((1u << w) - 1u)) << u;
    int h = 1;
    while (v + h < 32) {
        const uint32_t m2 = st.m[v + h];
        if ((m2 & run) != run) break;
        if ((st.vb[(v + h) * st.vs] & run) != run) break;
        st.m[v + h] = m2 & ~run;
        ++h;
    }
    st.R &= ~run;
    rec = st.base | (uint32_t)u | (uint32_t)v << 5 | (uint32_t)(w - 1) << 18 | (uint32_t)(h - 1) << 23;
    return true;
}

// Stage quads until the plane is done or the staging buffer is full.
// Returns the number of quads this call produced (staged or left pending).
__device__ __forceinline__ uint32_t sweep_produce(Sweep& st, MeshSmem& sm) {
    uint32_t produced = 0;
    while (true) {
        if (st.has_pending) {
            const uint32_t pos = atomicAdd(&sm.cursor, 1u);
            if (pos >= (uint32_t)kStageCap) break;       // full: keep it pending for the next round
            sm.tile.post.staging[pos] = st.pending;
            st.has_pending = false;
        }
        if (st.done) break;
        uint32_t rec;
        if (!sweep_next(st, rec)) break;
        st.pending = rec;
        st.has_pending = true;
        ++produced;
    }
    return produced;
}
// Count-only continuation (no staging): finish the plane, return the number of further quads.
__device__ __forceinline__ uint32_t sweep_count_rest(Sweep& st) {
    uint32_t n = 0, rec;
    while (!st.done && sweep_next(st, rec)) ++n;
    return n;
}

// ------------------------------------------------------------------ expansion: staged records -> buffers
// Writes quads [ord0, ord0+cnt) (chunk-local ordinals) of the chunk whose range starts at `first`.
template <class IdFn>
__device__ __forceinline__ void flush_staging(const MeshSmem& sm, uint32_t cnt, unsigned long long first,
                                              uint32_t ord0, const MeshOut& out, IdFn id_of) {
    const int tid = threadIdx.x;
    // vertices: two threads per quad, 16 B each -> a warp stores 512 contiguous bytes
    for (uint32_t i = tid; i < 2 * cnt; i += kThreads) {
        const uint32_t j = i >> 1, half = i & 1;
        const Quad q = unpack_rec(sm.tile.post.staging[j]);
        const uint32_t b = id_of(quad_origin_voxel(q));
        const unsigned long long va = make_vertex(q, b, 2 * half), vb = make_vertex(q, b, 2 * half + 1);
        st_global_v2u64(out.verts + (first + ord0 + j) * 4 + 2 * half, va, vb);
    }
    // indices: three threads per quad, 8 B each -> a warp stores 256 contiguous bytes
    for (uint32_t i = tid; i < 3 * cnt; i += kThreads) {
        const uint32_t j = i / 3, part = i - 3 * j;
        const uint32_t b4 = 4 * (ord0 + j);
        const uint32_t a = b4 + (part == 0 ? 0u : part == 1 ? 2u : 3u);
        const uint32_t b = b4 + (part == 0 ? 1u : part == 1 ? 2u : 0u);
        st_global_v2u32(out.indices + (first + ord0 + j) * 6 + 2 * part, a, b);
    }
}

// Reserve the chunk's range in the output arena (one thread), broadcast through shared memory.
__device__ __forceinline__ bool reserve_range(MeshSmem& sm, uint32_t chunk, uint32_t Q, const MeshOut& out,
                                              unsigned long long& first) {
    __shared__ unsigned long long s_first;
    if (threadIdx.x == 0) {
        unsigned long long f = Q ? atomicAdd(out.total, (unsigned long long)Q) : 0ull;
        const bool fits = f + Q <= out.cap;
        out.meta[chunk] = vxp_chunk_meta{fits ? (uint32_t)f : 0xFFFFFFFFu, Q};
        s_first = fits ? f : ~0ull;
    }
    __syncthreads();
    first = s_first;
    (void)sm;
    return first != ~0ull;
}

// ------------------------------------------------------------------ masks -> quads -> buffers
template <bool kGreedy, class IdFn>
__device__ __forceinline__ void emit_chunk(MeshSmem& sm, uint32_t chunk, bool uniform, const MeshOut& out,
                                           IdFn id_of) {
    const int tid = threadIdx.x;
    unsigned long long first;
    if (!kGreedy) {
        // ---- culled: every set bit is a quad; thread t owns rows t, t+256, ... (24 rows)
        uint32_t mine = 0;
#pragma unroll 4
        for (int r = tid; r < 192 * 32; r += kThreads) mine += __popc(sm.mask[(r >> 5) * kPad + (r & 31)]);
        uint32_t Q;
        const uint32_t pre = block_excl_scan(mine, sm.scratch, &Q);
        if (!reserve_range(sm, chunk, Q, out, first) || Q == 0) return;
        for (uint32_t base = 0; base < Q; base += kStageCap) {
            const uint32_t end = min(Q, base + (uint32_t)kStageCap);
            uint32_t ord = pre;
            for (int r = tid; r < 192 * 32 && ord < end; r += kThreads) {
                uint32_t bits = sm.mask[(r >> 5) * kPad + (r & 31)];
                const uint32_t c = __popc(bits);
                if (ord + c > base) {
                    const uint32_t rbase = (uint32_t)(r >> 5) << 10 | (uint32_t)(r & 31) << 5; // f<<15|s<<10|v<<5
                    uint32_t o = ord;
                    while (bits) {
                        const int u = __ffs(bits) - 1;
                        bits &= bits - 1;
                        if (o >= base && o < end) sm.tile.post.staging[o - base] = rbase | (uint32_t)u;
                        ++o;
                    }
                }
                ord += c;
            }
            __syncthreads();
            flush_staging(sm, end - base, first, base, out, id_of);
            __syncthreads();
        }
        return;
    }
    // ---- greedy: optimistic single sweep; staging overflow falls back to count + re-sweep in rounds
    if (tid == 0) sm.cursor = 0;
    __syncthreads();
    Sweep st;
    sweep_init(st, sm, uniform);
    uint32_t produced = sweep_produce(st, sm);
    if (__syncthreads_and(st.done && !st.has_pending)) {
        const uint32_t Q = sm.cursor;
        if (!reserve_range(sm, chunk, Q, out, first) || Q == 0) return;
        flush_staging(sm, Q, first, 0, out, id_of);
        return;
    }
    // heavy path (Q > kStageCap): finish counting, rebuild the consumed masks, sweep again in rounds
    produced += sweep_count_rest(st);
    uint32_t Q;
    block_excl_scan(produced, sm.scratch, &Q);
    const bool fits = reserve_range(sm, chunk, Q, out, first);
    if (!fits) return;
    build_masks<true>(sm, uniform);
    if (tid == 0) sm.cursor = 0;
    __syncthreads();
    sweep_init(st, sm, uniform);
    uint32_t flushed = 0;
    while (true) {
        sweep_produce(st, sm);
        const bool all_done = __syncthreads_and(st.done && !st.has_pending);
        const uint32_t cnt = min(sm.cursor, (uint32_t)kStageCap);
        flush_staging(sm, cnt, first, flushed, out, id_of);
        flushed += cnt;
        __syncthreads();
        if (all_done) break;
        if (tid == 0) sm.cursor = 0;
        __syncthreads();
    }
}

} // namespace vxp
