// vxp_device.cuh — device-side primitives shared by the sm_100a kernels.
// Implements SPEC.md v1 (builder-defined; the reference has no code for this
// path — SURVEY.md §0).  Everything here is integer / bitwise.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vxp {

constexpr int kThreads = 256;          // 8 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kPad = 33;               // row stride (words) of 32x32 word planes: conflict-free both ways
constexpr uint32_t kTileBytes = 65536; // one chunk: 32^3 u16
constexpr uint32_t kFull = 0xFFFFFFFFu;

// ---------------------------------------------------------------- TMA / mbarrier (PTX)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes -> visible to the async proxy (before a TMA overwrites / reads smem)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk TMA: global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- u16 -> bit tricks

__device__ __forceinline__ uint4 xor4(uint4 a, uint4 b) {
    return make_uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w);
}
// (hi:lo) viewed as u16 stream; returns the word that starts one u16 *before* `cur`
__device__ __forceinline__ uint32_t prev16(uint32_t prev, uint32_t cur) {
    return __funnelshift_l(prev, cur, 16); // (cur << 16) | (prev >> 16)
}

// 32x32 bit-matrix transpose across a warp: lane i holds row i on entry and column i on exit
// (exit bit j of lane i == entry bit i of lane j).
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = (j == 16)  ? 0x0000FFFFu
                           : (j == 8) ? 0x00FF00FFu
                           : (j == 4) ? 0x0F0F0F0Fu
                           : (j == 2) ? 0x33333333u
                                      : 0x55555555u;
        uint32_t y = __shfl_xor_sync(kFull, x, j);
        if (lane & j)
            x = (x & ~m) | ((y >> j) & m);
        else
            x = (x & m) | ((y << j) & ~m);
    }
    return x;
}

// N independent transposes, step by step: the N shuffles of a step are in flight together.  Per word and step one
// shuffle, one rotate and one bit select: the lane keeps the half of its bits selected by `keep` and takes the
// others from its partner's word rotated by j towards them (what a rotate wraps around is masked off).
template <int N>
__device__ __forceinline__ void transpose32xN(uint32_t (&x)[N], int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = (j == 16)  ? 0x0000FFFFu
                           : (j == 8) ? 0x00FF00FFu
                           : (j == 4) ? 0x0F0F0F0Fu
                           : (j == 2) ? 0x33333333u
                                      : 0x55555555u;
        const bool up = (lane & j) != 0;
        const uint32_t keep = up ? ~m : m;          // up: keep the high halves, take (partner >> j) into the low ones
        const uint32_t rot = up ? j : 32 - j;       // rotate right by j = shift right; by 32-j = shift left by j
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const uint32_t y = __shfl_xor_sync(kFull, x[k], j);
            x[k] = (x[k] & keep) | (__funnelshift_r(y, y, rot) & ~keep);
        }
    }
}

// block-wide helpers (kThreads threads, scratch: kWarps+1 words of shared memory)
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
// exclusive prefix of v over the block in thread order; *total gets the block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* scratch, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        uint32_t s = scratch[w];
        if (w < warp) base += s;
        sum += s;
    }
    *total = sum;
    __syncthreads();
    return base + inc - v;
}

// ---------------------------------------------------------------- SPEC §6 terrain (exact integer)

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}
struct TerrainKeys {
    uint32_t k[4];
};
__device__ __forceinline__ TerrainKeys terrain_keys(uint64_t seed) {
    TerrainKeys t;
#pragma unroll
    for (int o = 0; o < 4; ++o)
        t.k[o] = mix32((uint32_t)(seed >> 32) ^ mix32((uint32_t)seed + 0x9E3779B9u * (uint32_t)(o + 1)));
    return t;
}
__device__ __forceinline__ int grad_dot(uint32_t key, int i, int j, int ax, int az) {
    uint32_t g = mix32(mix32(key + (uint32_t)i * 0x85EBCA6Bu) ^ ((uint32_t)j * 0xC2B2AE35u)) & 7u;
    // [(1,0),(-1,0),(0,1),(0,-1),(1,1),(-1,1),(1,-1),(-1,-1)]
    int gx = (g < 4) ? ((g < 2) ? (g == 0 ? 1 : -1) : 0) : ((g & 1) ? -1 : 1);
    int gz = (g < 4) ? ((g < 2) ? 0 : (g == 2 ? 1 : -1)) : ((g & 2) ? -1 : 1);
    return gx * ax + gz * az;
}
__device__ __forceinline__ int fade16(int k) { return (k * k * (768 - 2 * k)) >> 8; }

__device__ __forceinline__ int terrain_height(const TerrainKeys& tk, int X, int Z) {
    long long sum = 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int S = 7 - o;
        const int A = 192 >> o;
        int Xo = X + 37 * o, Zo = Z + 59 * o;
        int ix = Xo >> S, iz = Zo >> S; // arithmetic shift = floor
        int dx = (Xo & ((1 << S) - 1)) << (8 - S);
        int dz = (Zo & ((1 << S) - 1)) << (8 - S);
        int d00 = grad_dot(tk.k[o], ix, iz, dx, dz);
        int d10 = grad_dot(tk.k[o], ix + 1, iz, dx - 256, dz);
        int d01 = grad_dot(tk.k[o], ix, iz + 1, dx, dz - 256);
        int d11 = grad_dot(tk.k[o], ix + 1, iz + 1, dx - 256, dz - 256);
        int fx = fade16(dx), fz = fade16(dz);
        int n0 = d00 * 65536 + fx * (d10 - d00);
        int n1 = d01 * 65536 + fx * (d11 - d01);
        long long n = (long long)n0 * 65536 + (long long)fz * (long long)(n1 - n0);
        sum += n * A;
    }
    return 256 + (int)(sum >> 40);
}
// ---- the same height function with the lattice gradients of a chunk's footprint hashed once.
// A footprint of at most 34 columns per axis touches at most 5 lattice points per axis in every
// octave (cells are 128, 64, 32, 16 wide), so 4 x 5 x 5 = 100 gradient hashes serve all of its
// columns; per column only the interpolation is left.  Bit-exact with terrain_height().
struct TerrainLattice {
    int ix0[4], iz0[4];          // lattice index of the footprint origin, per octave
    short g[4][5][5];            // gradient of lattice point (ix0+i, iz0+j): gx in the low byte, gz in the high byte
};
// threads 0..99 of the CTA; X0, Z0 = world coordinates of the footprint origin; needs a barrier afterwards
__device__ __forceinline__ void terrain_lattice_fill(TerrainLattice& L, const TerrainKeys& tk, int X0, int Z0) {
    const int t = threadIdx.x;
    if (t >= 100) return;
    const int o = t / 25, r = t - 25 * o, i = r / 5, j = r - 5 * i;
    const int S = 7 - o;
    const int ix0 = (X0 + 37 * o) >> S, iz0 = (Z0 + 59 * o) >> S;
    const uint32_t g = mix32(mix32(tk.k[o] + (uint32_t)(ix0 + i) * 0x85EBCA6Bu) ^ ((uint32_t)(iz0 + j) * 0xC2B2AE35u)) & 7u;
    const int gx = (g < 4) ? ((g < 2) ? (g == 0 ? 1 : -1) : 0) : ((g & 1) ? -1 : 1);
    const int gz = (g < 4) ? ((g < 2) ? 0 : (g == 2 ? 1 : -1)) : ((g & 2) ? -1 : 1);
    L.g[o][i][j] = (short)((gx & 0xFF) | (gz << 8));
    if (r == 0) { L.ix0[o] = ix0; L.iz0[o] = iz0; }
}
__device__ __forceinline__ int lattice_dot(short g, int ax, int az) {
    return (int)(signed char)(g & 0xFF) * ax + (int)(g >> 8) * az;
}
// The same dot product as one IDP.2A: `axz` holds (ax, az) as two int16 (ax low), g has gx in byte 0 and gz in byte 1.
__device__ __forceinline__ int lattice_dot2(unsigned short g, uint32_t axz) { return __dp2a_lo((int)axz, (int)(uint32_t)g, 0); }
__device__ __forceinline__ int terrain_height_lattice(const TerrainLattice& L, int X, int Z) {
    long long sum = 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int S = 7 - o;
        const int A = 192 >> o;
        const int Xo = X + 37 * o, Zo = Z + 59 * o;
        const int i = (Xo >> S) - L.ix0[o], j = (Zo >> S) - L.iz0[o];
        const int dx = (Xo & ((1 << S) - 1)) << (8 - S);
        const int dz = (Zo & ((1 << S) - 1)) << (8 - S);
        // (dx, dz) as two int16 in one word; dx - 256 is dx + 0xFF00 in 16 bits (0 <= dx < 256: no carry out of the half)
        const uint32_t a00 = (uint32_t)dx | (uint32_t)dz << 16;
        const int d00 = lattice_dot2((unsigned short)L.g[o][i][j], a00);
        const int d10 = lattice_dot2((unsigned short)L.g[o][i + 1][j], a00 + 0x0000FF00u);
        const int d01 = lattice_dot2((unsigned short)L.g[o][i][j + 1], a00 + 0xFF000000u);
        const int d11 = lattice_dot2((unsigned short)L.g[o][i + 1][j + 1], a00 + 0xFF00FF00u);
        const int fx = fade16(dx), fz = fade16(dz);
        const int n0 = d00 * 65536 + fx * (d10 - d00);
        const int n1 = d01 * 65536 + fx * (d11 - d01);
        const long long n = (long long)n0 * 65536 + (long long)fz * (long long)(n1 - n0);
        sum += n * A;
    }
    return 256 + (int)(sum >> 40);
}
__device__ __forceinline__ uint32_t terrain_block(int h, int Y) {
    if (Y > h) return 0u;
    if (Y == h) return 1u;
    if (Y >= h - 3) return 2u;
    return ((Y >> 3) & 1) ? 4u : 3u;
}
__device__ __forceinline__ uint32_t pattern_block(int pattern, int X, int Y, int Z) {
    if (pattern == 0) return (uint32_t)((X + Y + Z) & 1);
    return 1u + (mix32(mix32(mix32((uint32_t)X) ^ (uint32_t)Y) ^ (uint32_t)Z) & 3u);
}

// ---------------------------------------------------------------- SPEC §5 vertices

// staged quad record: u | v<<5 | s<<10 | f<<15 | (w-1)<<18 | (h-1)<<23   (id is looked up at expansion)
__device__ __forceinline__ uint32_t make_rec(int f, int s, int u, int v, int w, int h) {
    return (uint32_t)u | (uint32_t)v << 5 | (uint32_t)s << 10 | (uint32_t)f << 15 |
           (uint32_t)(w - 1) << 18 | (uint32_t)(h - 1) << 23;
}
struct Quad {
    int u, v, s, f, w, h;
};
__device__ __forceinline__ Quad unpack_rec(uint32_t r) {
    Quad q;
    q.u = r & 31;
    q.v = (r >> 5) & 31;
    q.s = (r >> 10) & 31;
    q.f = (r >> 15) & 7;
    q.w = ((r >> 18) & 31) + 1;
    q.h = ((r >> 23) & 31) + 1;
    return q;
}
// chunk-local voxel index of the quad's origin cell
__device__ __forceinline__ int quad_origin_voxel(const Quad& q) {
    const int axis = q.f >> 1;
    int x = axis == 0 ? q.s : q.u;
    int y = axis == 0 ? q.u : (axis == 1 ? q.s : q.v);
    int z = axis == 2 ? q.s : q.v;
    return x + 32 * y + 1024 * z;
}
__device__ __forceinline__ unsigned long long make_vertex(const Quad& q, uint32_t b, int k) {
    const bool flip = (0x19u >> q.f) & 1u;      // f in {0,3,4}
    const int c = flip ? ((4 - k) & 3) : k;     // corner id c0..c3
    const int cu = q.u + ((c == 1 || c == 2) ? q.w : 0);
    const int cv = q.v + ((c >= 2) ? q.h : 0);
    const int along = q.s + (q.f & 1);
    const int axis = q.f >> 1;
    const uint32_t x = axis == 0 ? along : cu;
    const uint32_t y = axis == 0 ? cu : (axis == 1 ? along : cv);
    const uint32_t z = axis == 2 ? along : cv;
    const uint32_t w0 = x | y << 6 | z << 12 | (uint32_t)q.f << 18 | (uint32_t)k << 21;
    const uint32_t w1 = b | (uint32_t)(q.w - 1) << 16 | (uint32_t)(q.h - 1) << 21;
    return (unsigned long long)w0 | (unsigned long long)w1 << 32;
}

__device__ __forceinline__ void st_global_v2u64(void* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_global_v2u32(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// 16 bytes of a voxel array that the writing kernel does not read again: streaming store (evict-first in L2; measured on
// 4096 chunks: unpack 72 -> 65 us, terrain fill 59.7 -> 59.1 us, LOD unchanged)
__device__ __forceinline__ void st_voxels16(uint4* p, const uint4& v) { __stcs(p, v); }

} // namespace vxp
