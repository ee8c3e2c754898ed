// vxp_mesh.cuh — the two meshing kernels for stored chunks (SPEC §2-§5): one CTA (256 threads) per chunk.
//
// The chunk is STREAMED through a ring of 16 KiB slabs (bulk TMA copies, the next one issued as soon as a buffer
// has been converted); nothing needs the voxels once the occupancy bits and the id bit planes are out, so what a
// CTA keeps is small and several CTAs share an SM (greedy: four, culled: five):
//
//   ring (2 x 16 KiB) ->  greedy: X-plane face masks, transposed equality rows of the X planes, staging / sweep log
//                         culled: window of non-empty face words, staging
//   eq   (12.4 KiB)       greedy: id planes (registers while streaming, stored for the chunks that use them);
//                         the equality rows EX / EY / EZ afterwards (in place)
//   occ  (4.5 KiB)        occupancy with halo; the Y / Z plane face masks are derived from it on the fly
//                         (M = occ & ~occ_neighbour: two loads instead of one stored word)
//   greedy: eq, the carry of the second pass and the row tables are dead while the chunk streams: a third ring buffer
//   lies over them, so three of the four slabs are in flight from the start.
//
// Greedy chunks whose ids do not fit the id planes stream a second time (L2 hits) and compute their equality
// rows per voxel, slab by slab.  Halo exchange, quad-record store and launch chaining: vxp_mesh_common.cuh.
//
// Greedy sweep (SPEC §4 evaluated bit-parallel over the 32 columns of a plane).  After row v-1 every set cell of
// that row belongs to exactly one active run (S = start bits, E = end bits, cover C = mask row v-1).  Row v: a
// run survives iff all its cells are set in row v with the id of the cell above (X = C & M & V); survivors are
// found with one carry chain upwards (segmented AND-reduce: T = (X & ~E) + S, Es = T & E & X) and spread back over
// their run with one carry chain on the bit-reversed words.  Runs that do not survive are finished quads; the
// cells of row v not covered by survivors start new runs, split where the id changes (H).
#pragma once
#include "vxp_mesh_common.cuh"

namespace vxp {
namespace mesh {

// CTAs per SM the kernels are compiled for (and the fused path sizes its persistent grid with)
#ifndef VXP_GREEDY_CTAS
#define VXP_GREEDY_CTAS 4
#endif
#ifndef VXP_CULLED_CTAS
#define VXP_CULLED_CTAS 5
#endif
constexpr int kGreedyCtas = VXP_GREEDY_CTAS, kCulledCtas = VXP_CULLED_CTAS;
// what a chunk keeps besides the pipeline's own arrays: the carried layer of the second streaming pass (greedy, stored
// chunks with large ids) or the column's height footprint (fused terrain path) - never both
union Aux {
    uint4 carry[128];
    short heights[34 * 34 + 4];
};

struct PostG {   // what lives in the ring once the chunk has been streamed
    static constexpr int kCap = (2 * kSlabBytes - (64 * kPad + 2 * kPlaneWords) * 4) / 4;   // 3968 words
    uint32_t xmask[64 * kPad];   // X planes only: xmask[(f*32 + x)*33 + z] bit y
    uint32_t ht[kPlaneWords];    // X planes: ht[x*33+z] bit y = EY transposed
    uint32_t vt[kPlaneWords];    // X planes: vt[x*33+z] bit y = EZ transposed
    uint32_t staging[kCap];
};
struct __align__(128) SmemG {
    union {
        unsigned char ring[2][kSlabBytes];
        PostG post;
    } tile;
    // eq .. first_run are dead while the chunk streams (the id planes are in registers until it has been read): the
    // third ring buffer lies over them
    uint32_t eq[3][kPlaneWords];   // [b][z*33+y]: id plane b while streaming; afterwards EX / EY / EZ (same layout as v4's ex / ey / ez)
    Aux aux;                       // second streaming pass: the last layer of the previous slab / fused path: heights
    uint32_t seg[192];
    uint32_t rows[192];
    uint32_t first_run[192];
    uint32_t occ[34 * 34];
    uint32_t xh[2][32];
    uint32_t actz[32];
    uint32_t scratch[kWarps + 1];
    uint32_t cursor;
    uint32_t counted;
    uint32_t ones;                 // 0xFFFFFFFF: stand-in equality row for single-id chunks
    uint32_t zero;                 // 0: stand-in "neighbour" row of the stored X masks
    uint32_t ref;                  // voxels 0 and 1 of the chunk
    unsigned long long first;
    unsigned long long mbar[3];
};
static_assert(sizeof(PostG) <= 2 * kSlabBytes, "post layout overflows the ring");
static_assert(offsetof(SmemG, eq) == 2 * kSlabBytes && offsetof(SmemG, occ) >= 3 * kSlabBytes, "the third ring buffer overlays eq .. first_run");
#ifndef VXP_RING
#define VXP_RING 3
#endif
constexpr int kRingG = VXP_RING;   // ring buffers of the greedy kernel's first pass (2: the tile only)
// per CTA: the struct + 128 bytes of static shared memory (fused kernel) + 1 KiB the driver reserves; 227 KiB per SM
static_assert(kGreedyCtas * (sizeof(SmemG) + 128 + 1024) <= 227 * 1024, "shared memory of the greedy kernels exceeds the CTAs-per-SM target");

// ------------------------------------------------------------------ streaming
// ring buffer b: 0, 1 = the tile, 2 = over eq .. first_run
__device__ __forceinline__ unsigned char* ring_buffer(SmemG& sm, int b) { return reinterpret_cast<unsigned char*>(&sm) + b * kSlabBytes; }
__device__ __forceinline__ void issue_ring_slab(SmemG& sm, const uint16_t* gvox, int s, int b) {
    mbar_expect_tx(&sm.mbar[b], kSlabBytes);
    tma_load_1d(ring_buffer(sm, b), reinterpret_cast<const unsigned char*>(gvox) + (size_t)s * kSlabBytes, kSlabBytes, &sm.mbar[b]);
}

// Slab s (in ring buffer b) -> occupancy words and id-plane words.  A thread converts one 64-byte row of the slab
// (32 voxels -> one word of each kind, so one 4-byte store instead of four byte stores); a warp takes 8 values of y in
// 4 layers, the shape that is most often one value in a terrain world (air above, rock below).  The four 16-byte pieces
// of a row are read in an order rotated by lane / 2, which makes a quarter-warp touch every bank once, and the
// finished words are rotated back.
__device__ __forceinline__ int convert_row_z(int s) { return 8 * s + 4 * ((int)threadIdx.x >> 7) + (((int)threadIdx.x >> 3) & 3); }
__device__ __forceinline__ int convert_row_y() { return 8 * (((int)threadIdx.x >> 5) & 3) + ((int)threadIdx.x & 7); }
__device__ __forceinline__ void convert_ring_slab(SmemG& sm, int s, int b, uint32_t& v_or, uint32_t& v_and, uint32_t& o_and, IdPlanes& ip) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int j = (lane >> 1) & 3;
    const int z = convert_row_z(s), y = convert_row_y();
    const uint4* row = reinterpret_cast<const uint4*>(ring_buffer(sm, b)) + ((z & 7) * 32 + y) * 4;
    uint4 q[4];                                                // q[k] = piece (k + j) & 3 of the row
    uint32_t slab_or = 0, slab_and = kFull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q[k] = row[(k + j) & 3];
        slab_or |= q[k].x | q[k].y | q[k].z | q[k].w;
        slab_and &= q[k].x & q[k].y & q[k].z & q[k].w;
    }
    if (s == 0 && tid == 0) sm.ref = q[0].x;                   // thread 0 reads its pieces in order
    v_or |= slab_or;
    v_and &= slab_and;
    const uint32_t id = slab_or & 0xFFFFu;
    const uint32_t first = __shfl_sync(kFull, slab_or, 0);
    const bool same = slab_or == slab_and && id == (slab_or >> 16) && slab_or == first;
    const bool uniform = __all_sync(kFull, same);              // the warp's 8 x 4 rows hold one value: nothing to gather bit by bit
    uint32_t p0 = 0, p1 = 0, p2 = 0, ob = 0;
    if (uniform) {
        ob = id ? kFull : 0u;
        p0 = (id & 1u) ? kFull : 0u;
        p1 = (id & 2u) ? kFull : 0u;
        p2 = (id & 4u) ? kFull : 0u;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) ob |= nz8(q[k]) << (8 * k);
        id_plane_bytes<0>(q[0], p0, p1, p2);
        id_plane_bytes<1>(q[1], p0, p1, p2);
        id_plane_bytes<2>(q[2], p0, p1, p2);
        id_plane_bytes<3>(q[3], p0, p1, p2);
        const uint32_t rot = 8u * (uint32_t)j;                 // byte k holds piece (k + j) & 3: rotate it there
        ob = __funnelshift_l(ob, ob, rot);
        p0 = __funnelshift_l(p0, p0, rot);
        p1 = __funnelshift_l(p1, p1, rot);
        p2 = __funnelshift_l(p2, p2, rot);
    }
    o_and &= ob;
    sm.occ[(z + 1) * 34 + y + 1] = ob;
#pragma unroll
    for (int t = 0; t < kSlabs; ++t)                        // s is an unrolled loop counter of the caller: ip stays in registers
        if (t == s) { ip.w[0][t] = p0; ip.w[1][t] = p1; ip.w[2][t] = p2; }
}
// The id planes of a chunk that will use them (several ids, all small): registers -> sm.eq[b][z*33+y], bit x.
__device__ __forceinline__ void store_id_planes_g(SmemG& sm, const IdPlanes& ip) {
#pragma unroll
    for (int s = 0; s < kSlabs; ++s) {
        const int word = convert_row_z(s) * kPad + convert_row_y();
#pragma unroll
        for (int b = 0; b < kIdPlanes; ++b) sm.eq[b][word] = ip.w[b][s];
    }
}

__device__ __forceinline__ ChunkFlags reduce_flags_g(SmemG& sm, uint32_t v_or, uint32_t v_and, uint32_t o_and) {
    ChunkFlags f;
    f.any_solid = __syncthreads_or(v_or != 0u);                // barrier: occupancy, planes and sm.ref published
    const uint32_t ref = sm.ref;
    f.all_solid = __syncthreads_and(o_and == kFull);
    f.single_id = __syncthreads_and(v_or == ref && v_and == ref && (ref & 0xFFFFu) == (ref >> 16));
    f.small_ids = __syncthreads_and(((v_or | (v_or >> 16)) & 0xFFFFu) < (1u << kIdPlanes));
    return f;
}

// Second pass for chunks whose ids do not fit the planes: stream the chunk again and compute EX / EY / EZ of the
// rows that carry a visible face from the voxels themselves (slab by slab; the layer below a
// slab's first one comes from sm.carry).  `phase` = number of completed phases of each ring barrier so far (2).
__device__ __forceinline__ void equality_second_pass(SmemG& sm, const uint16_t* gvox) {
    // completed phases of the two barriers after the first pass: ring of 2: 2 and 2; ring of 3: 2 (slabs 0, 3) and 1
    constexpr uint32_t kDone0 = 2u, kDone1 = kRingG == 3 ? 1u : 2u;
    const int tid = threadIdx.x, lane = tid & 31;
    const int p = lane & 3;
    const uint32_t selA = (p & 1) ? 0x3715u : 0x6240u;   // 4x4 byte transpose across the 4 lanes of a row
    const uint32_t selB = (p & 2) ? 0x3276u : 0x5410u;
    if (tid == 0) {
        issue_ring_slab(sm, gvox, 0, 0);
        issue_ring_slab(sm, gvox, 1, 1);
    }
#pragma unroll 1
    for (int s = 0; s < kSlabs; ++s) {
        mbar_wait(&sm.mbar[s & 1], (((s & 1) ? kDone1 : kDone0) + (uint32_t)(s >> 1)) & 1u);
        const uint4* slab = reinterpret_cast<const uint4*>(ring_buffer(sm, s & 1));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = tid + k * kThreads;                  // within the slab
            const int row = s * 256 + (idx >> 2), z = row >> 5, y = row & 31;
            if (((sm.actz[z] >> (y & 24)) & 0xFFu) == 0) continue;   // warp-uniform: the warp's 8 rows are y0..y0+7 of one layer
            const uint4 q = slab[idx];
            const uint32_t leftw = __shfl_up_sync(kFull, q.w, 1);
            const uint4 px = make_uint4(prev16(leftw, q.x), prev16(q.x, q.y), prev16(q.y, q.z), prev16(q.z, q.w));
            uint32_t w = (~nz8(xor4(q, px))) & 0xFFu;
            if (y > 0) w |= ((~nz8(xor4(q, slab[idx - 4]))) & 0xFFu) << 8;
            if (z > 0) {
                const uint4 below = (z & 7) ? slab[idx - 128] : sm.aux.carry[idx & 127];
                w |= ((~nz8(xor4(q, below))) & 0xFFu) << 16;
            }
            uint32_t t = __shfl_xor_sync(kFull, w, 1);
            w = __byte_perm(w, t, selA);
            t = __shfl_xor_sync(kFull, w, 2);
            w = __byte_perm(w, t, selB);
            if (p < 3) sm.eq[p][z * kPad + y] = w;               // p = 0 EX, 1 EY, 2 EZ
        }
        __syncthreads();                                         // everybody has read the carried layer and this slab
        if (s + 1 < kSlabs) {
            if (tid < 128) sm.aux.carry[tid] = slab[7 * 128 + tid];  // layer 8s+7 for the first layer of slab s+1
            __syncthreads();
            if (tid == 0 && s + 2 < kSlabs) issue_ring_slab(sm, gvox, s + 2, s & 1);
        }
    }
}

// ------------------------------------------------------------------ planes -> equality rows, occupancy -> X masks
// Plane path: EX / EY / EZ of every row from the id planes (three xors per plane), written back over the planes
// (through registers: the rows read their -y / -z neighbours).  Both paths: the X-plane masks and the transposed
// equality rows of the X planes (32x32 bit transposes across the warp).  Returns the OR of all face words.
__device__ __forceinline__ uint32_t build_masks_g(SmemG& sm, bool with_eq, bool with_planes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t acc = 0;
    uint32_t tx[8], te[8], ex[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp, y = lane;
        const uint32_t* o = sm.occ + (z + 1) * 34 + (y + 1);
        const uint32_t c = o[0];
        const uint32_t lo = (sm.xh[0][z] >> y) & 1u, hi = (sm.xh[1][z] >> y) & 1u;
        tx[k] = c & ~((c << 1) | lo);                   // -X visible, bit x
        tx[4 + k] = c & ~((c >> 1) | (hi << 31));       // +X visible, bit x
        const uint32_t my = c & ~o[-1], py = c & ~o[+1], mz = c & ~o[-34], pz = c & ~o[+34];
        acc |= my | py | mz | pz | tx[k] | tx[4 + k];
        // non-empty rows of the planes (what the sweep visits): Z planes (slice z, row y) by ballot, Y planes
        // (slice y, row z) bit by bit; sm.rows was zeroed by the caller
        const uint32_t bz0 = __ballot_sync(kFull, mz != 0u), bz1 = __ballot_sync(kFull, pz != 0u);
        if (lane == 0) { sm.rows[4 * 32 + z] = bz0; sm.rows[5 * 32 + z] = bz1; }
        if (my) atomicOr(&sm.rows[2 * 32 + y], 1u << z);
        if (py) atomicOr(&sm.rows[3 * 32 + y], 1u << z);
        te[k] = te[4 + k] = ex[k] = 0;
        if (with_eq) {
            if (with_planes) {
                const uint32_t* B = &sm.eq[0][z * kPad + y];
                uint32_t dx = 0, dy = 0, dz = 0;
#pragma unroll
                for (int b = 0; b < kIdPlanes; ++b) {
                    const uint32_t w = B[b * kPlaneWords];
                    const uint32_t wy = y > 0 ? B[b * kPlaneWords - 1] : w;
                    const uint32_t wz = z > 0 ? B[b * kPlaneWords - kPad] : w;
                    dx |= w ^ (w << 1);
                    dy |= w ^ wy;
                    dz |= w ^ wz;
                }
                ex[k] = ~dx;
                te[k] = ~dy;
                te[4 + k] = ~dz;
            } else {
                te[k] = sm.eq[1][z * kPad + y];
                te[4 + k] = sm.eq[2][z * kPad + y];
            }
        }
    }
    if (with_eq && with_planes) {
        __syncthreads();                                // every row has read the planes it needs
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int z = k * kWarps + warp, y = lane;
            sm.eq[0][z * kPad + y] = ex[k];
            sm.eq[1][z * kPad + y] = te[k];
            sm.eq[2][z * kPad + y] = te[4 + k];
        }
    }
    uint32_t xany = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) xany |= tx[k];
    const bool anyx = __any_sync(kFull, xany != 0);
    if (anyx) {
        transpose32xN(tx, lane);
        if (with_eq) transpose32xN(te, lane);
    }
    PostG& P = sm.tile.post;
    uint32_t rx0 = 0, rx1 = 0;                           // X planes (slice x = lane, row z) after the transposes
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int z = k * kWarps + warp;
        rx0 |= (uint32_t)(tx[k] != 0u) << z;
        rx1 |= (uint32_t)(tx[4 + k] != 0u) << z;
        P.xmask[(0 * 32 + lane) * kPad + z] = tx[k];
        P.xmask[(1 * 32 + lane) * kPad + z] = tx[4 + k];
        if (with_eq && anyx) {
            P.ht[lane * kPad + z] = te[k];
            P.vt[lane * kPad + z] = te[4 + k];
        }
    }
    if (rx0) atomicOr(&sm.rows[0 * 32 + lane], rx0);
    if (rx1) atomicOr(&sm.rows[1 * 32 + lane], rx1);
    return acc;
}

// ------------------------------------------------------------------ face-mask rows
// Row v of plane (f, s): stored for the X planes, occ & ~occ_neighbour for the others.
struct MaskRows {
    const uint32_t* c;   // centre word of row 0
    const uint32_t* nb;  // neighbour word of row 0 (&sm.zero for stored masks)
    int cs, ns;          // row strides
    __device__ __forceinline__ uint32_t row(int v) const { return c[v * cs] & ~nb[v * ns]; }
};
__device__ __forceinline__ MaskRows mask_rows(const SmemG& sm, uint32_t plane) {
    const uint32_t f = plane >> 5, s = plane & 31u;
    MaskRows m;
    if (f < 2) { m.c = sm.tile.post.xmask + plane * kPad; m.cs = 1; m.nb = &sm.zero; m.ns = 0; }
    else if (f < 4) { m.c = sm.occ + 34 + (s + 1); m.cs = 34; m.nb = m.c + (f == 2 ? -1 : 1); m.ns = 34; }      // slice y = s, row z
    else { m.c = sm.occ + (s + 1) * 34 + 1; m.cs = 1; m.nb = m.c + (f == 4 ? -34 : 34); m.ns = 1; }             // slice z = s, row y
    return m;
}

// ------------------------------------------------------------------ sweep (v4's recurrence on MaskRows)
struct SweepG {
    MaskRows m;
    const uint32_t* hb;
    const uint32_t* vb;
    int hs, vs;
    uint32_t C, S, E, r0, r1, r2, r3, r4;
    uint32_t base;
};
__device__ __forceinline__ void sweep_init_g(SweepG& st, SmemG& sm, bool with_eq) {
    const int tid = threadIdx.x;
    const int f = tid >> 5, s = tid & 31;
    st.C = st.S = st.E = st.r0 = st.r1 = st.r2 = st.r3 = st.r4 = 0;
    st.base = (uint32_t)f << 15 | (uint32_t)s << 10;
    st.m = mask_rows(sm, tid < 192 ? (uint32_t)tid : 0u);
    st.hb = st.vb = &sm.ones;
    st.hs = st.vs = 0;
    if (with_eq) {
        PostG& P = sm.tile.post;
        if (f < 2) { st.hb = P.ht + s * kPad; st.hs = 1; st.vb = P.vt + s * kPad; st.vs = 1; }                  // X planes: u=y, v=z
        else if (f < 4) { st.hb = sm.eq[0] + s; st.hs = kPad; st.vb = sm.eq[2] + s; st.vs = kPad; }             // Y planes: u=x, v=z
        else { st.hb = sm.eq[0] + s * kPad; st.hs = 1; st.vb = sm.eq[1] + s * kPad; st.vs = 1; }                // Z planes: u=x, v=y
    }
}
__device__ __forceinline__ void sweep_emit_g(const SweepG& st, uint32_t dS, uint32_t dE, int v, uint32_t* staging, uint32_t cap,
                                             uint32_t* cursor) {
    if (dS == 0) return;
    uint32_t pos = atomicAdd(cursor, (uint32_t)__popc(dS));
    const uint32_t tail = st.base | (uint32_t)(v - 1) << 23;
    do {
        const int u = __ffs(dS) - 1, e = __ffs(dE) - 1;
        dS &= dS - 1;
        dE &= dE - 1;
        const uint32_t v0 = ((st.r0 >> u) & 1u) | ((st.r1 >> u) & 1u) << 1 | ((st.r2 >> u) & 1u) << 2 |
                            ((st.r3 >> u) & 1u) << 3 | ((st.r4 >> u) & 1u) << 4;
        const uint32_t rec = tail + (uint32_t)u + (v0 << 5) + ((uint32_t)(e - u) << 18) - (v0 << 23);
        if (pos < cap) staging[pos] = rec;
        ++pos;
    } while (dS);
}
__device__ __forceinline__ void sweep_row_g(SweepG& st, int v, uint32_t Mv, uint32_t Hv, uint32_t Vv, uint32_t& dS, uint32_t& dE) {
    uint32_t Cs = 0;
    if (st.C) {
        const uint32_t X = st.C & Mv & Vv;
        const uint32_t T = (X & ~st.E) + st.S;
        const uint32_t Es = T & st.E & X;
        const uint32_t P = __brev(st.C) & ~__brev(st.S);
        const uint32_t Q = P + __brev(Es);
        Cs = __brev(Q ^ P);
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
__device__ __forceinline__ void sweep_step_g(SweepG& st, int v, uint32_t* staging, uint32_t cap, uint32_t* cursor) {
    if (v == 32) {
        sweep_emit_g(st, st.S, st.E, 32, staging, cap, cursor);
        st.S = st.E = st.C = 0;
        return;
    }
    const uint32_t Mv = st.m.row(v);
    if ((Mv | st.C) == 0) return;
    uint32_t dS, dE;
    SweepG old = st;
    sweep_row_g(st, v, Mv, st.hb[v * st.hs], st.vb[v * st.vs], dS, dE);
    sweep_emit_g(old, dS, dE, v, staging, cap, cursor);
}
__device__ __forceinline__ uint32_t sweep_rows_g(const SweepG& st) {
    uint32_t rm = 0;
#pragma unroll
    for (int v = 0; v < 32; ++v) rm |= (uint32_t)(st.m.row(v) != 0) << v;
    return rm;
}
template <bool kStore = true>
__device__ __forceinline__ uint32_t sweep_log_g(SweepG& st, uint32_t rm, uint32_t plane, uint2* entry) {
    uint32_t births = 0;
    uint32_t todo = rm | (rm << 1);
    const int trips = (int)__reduce_max_sync(kFull, (uint32_t)__popc(todo));
    int v = todo ? __ffs(todo) - 1 : 0;
    bool valid = todo != 0;
    todo &= todo - 1;
    uint32_t Mv = st.m.row(v), Hv = st.hb[v * st.hs], Vv = st.vb[v * st.vs];
#pragma unroll 2
    for (int it = 0; it < trips; ++it) {
        const bool validn = todo != 0;
        const int vn = validn ? __ffs(todo) - 1 : 0;
        todo &= todo - 1;
        const uint32_t Mn = st.m.row(vn), Hn = st.hb[vn * st.hs], Vn = st.vb[vn * st.vs];
        const uint32_t X = st.C & Mv & Vv;
        const uint32_t T = (X & ~st.E) + st.S;
        const uint32_t Es = T & st.E & X;
        const uint32_t P = __brev(st.C) & ~__brev(st.S);
        const uint32_t Cs = __brev((P + __brev(Es)) ^ P);
        const uint32_t R = Mv & ~Cs;
        const uint32_t NS = R & ~((R << 1) & Hv);
        const uint32_t NE = R & ~((R >> 1) & (Hv >> 1));
        if (valid) {
            st.S = (st.S & Cs) | NS;
            st.E = (st.E & Cs) | NE;
            st.C = Mv;
            if (kStore && Mv) *entry++ = make_uint2(Cs, plane << 5 | (uint32_t)v | births << 13);
            births += __popc(NS);
        }
        v = vn; valid = validn; Mv = Mn; Hv = Hn; Vv = Vn;
    }
    return births;
}
__device__ __forceinline__ uint32_t plane_h_row_g(const SmemG& sm, bool with_eq, uint32_t plane, uint32_t v) {
    if (!with_eq) return kFull;
    const uint32_t f = plane >> 5, s = plane & 31u;
    return f < 2 ? sm.tile.post.ht[s * kPad + v] : f < 4 ? sm.eq[0][v * kPad + s] : sm.eq[0][s * kPad + v];
}
// Staged records in up to two pieces: ordinals [0, na) at a, the rest at b.
struct RecStore {
    uint32_t* a;
    uint32_t na;
    uint32_t* b;
    __device__ __forceinline__ uint32_t& at(uint32_t i) const { return i < na ? a[i] : b[i - na]; }
};
// Only the records with ordinals in [lo, hi) are staged, at ordinal - lo.
__device__ __forceinline__ void stage_births_g(const SmemG& sm, bool with_eq, const SweepLog& log, uint32_t nL, const RecStore& rec,
                                               uint32_t lo = 0u, uint32_t hi = 0xFFFFFFFFu) {
    for (uint32_t i = threadIdx.x; i < nL; i += kThreads) {
        const uint2 en = log.entry[i];
        const uint32_t plane = (en.y >> 5) & 255u, v = en.y & 31u;
        const uint32_t R = mask_rows(sm, plane).row((int)v) & ~en.x;
        if (R == 0) continue;
        const uint32_t Hv = plane_h_row_g(sm, with_eq, plane, v);
        uint32_t NS = R & ~((R << 1) & Hv), NE = R & ~((R >> 1) & (Hv >> 1));
        uint32_t pos = log.first[plane] + (en.y >> 13);
        if (pos >= hi || pos + __popc(NS) <= lo) continue;
        const uint32_t base = (en.y & 0x1FFFu) << 5;
        do {
            const uint32_t u = __ffs(NS) - 1, e = __ffs(NE) - 1;
            NS &= NS - 1;
            NE &= NE - 1;
            if (pos >= lo && pos < hi) rec.at(pos - lo) = base | u | (e - u) << 18;
            ++pos;
        } while (NS);
    }
}

// ------------------------------------------------------------------ chunks whose sweep log does not fit
// The planes are swept in two groups - the -X, -Y, -Z planes, then the others - if the log of either fits with room for
// a window of at least 1024 records below it, and each group is staged and flushed window by window before the next
// one is swept (the V rows stay alive, so the records cannot spill into them).  The chunk's range has been reserved and
// published by the caller; ordinals run on across the groups.  Returns false, having done nothing, if a half does not
// fit either (the caller's stepped sweep takes those; for a terrain chunk it costs a quarter more than this way).
template <class IdFn>
__device__ __forceinline__ bool emit_greedy_halves(SmemG& sm, bool with_eq, const MeshOut& out, IdFn id_of) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t kCap = PostG::kCap;
    uint32_t* const staging = sm.tile.post.staging;
    auto nothing = [] {};
    uint32_t rows_of[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) rows_of[f] = __reduce_add_sync(kFull, (uint32_t)__popc(sm.rows[f * 32 + lane]));
    const uint32_t even = rows_of[0] + rows_of[2] + rows_of[4], odd = rows_of[1] + rows_of[3] + rows_of[5];
    if (2u * max(even, odd) + 1024u > kCap) return false;              // block-uniform
    const uint32_t rm = tid < 192 ? sm.rows[tid] : 0u;
    uint32_t base = 0;                                                 // ordinal of the group's first quad
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        const bool member = tid < 192 && (warp & 1) == g;              // warp-uniform
        uint32_t nL = 0, seg = 0;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            if ((f & 1) != g) continue;
            if (f < warp) seg += rows_of[f];
            nL += rows_of[f];
        }
        seg += warp_incl_scan((uint32_t)__popc(rm), lane) - (uint32_t)__popc(rm);
        __syncthreads();                              // the previous group has been flushed: log, staging, seg and scratch are free
        if (member) sm.seg[tid] = seg;
        SweepLog log;
        log.entry = reinterpret_cast<uint2*>(staging + kCap - 2 * nL);
        log.seg = sm.seg;
        log.rows = sm.rows;
        log.first = sm.first_run;
        uint32_t births = 0;
        if (member) {
            SweepG st;
            sweep_init_g(st, sm, with_eq);
            births = sweep_log_g(st, rm, (uint32_t)tid, log.entry + seg);
        }
        const uint32_t incl = warp_incl_scan(births, lane);
        if (lane == 31) sm.scratch[warp] = incl;
        __syncthreads();
        uint32_t first_run = base + incl - births, Qg = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = sm.scratch[w];
            if (w < warp) first_run += t;
            Qg += t;
        }
        if (member) sm.first_run[tid] = first_run;
        const uint32_t window = kCap - 2 * nL;
        const RecStore rec{staging, window, staging};
        for (uint32_t lo = 0; lo < Qg; lo += window) {
            const uint32_t hi = min(Qg, lo + window);
            __syncthreads();                          // first_run published, the sweep is over / the previous window has been flushed
            stage_births_g(sm, with_eq, log, nL, rec, base + lo, base + hi);
            __syncthreads();
            if (!flush_staging(sm, [rec, &log](uint32_t i) { const uint32_t r = rec.at(i); return r | (run_height(log, r) - 1u) << 23; },
                               hi - lo, base + lo, out, id_of, nothing, false))
                return true;
        }
        base += Qg;
    }
    return true;
}

// ------------------------------------------------------------------ masks -> quads -> buffers
template <class IdFn>
__device__ __forceinline__ void emit_greedy(SmemG& sm, uint32_t mi, bool with_eq, const MeshOut& out, IdFn id_of, Prof& prof) {
    const int tid = threadIdx.x;
    constexpr uint32_t kCap = PostG::kCap;
    uint32_t* const staging = sm.tile.post.staging;
    auto nothing = [] {};
    auto staged = [staging](uint32_t i) { return staging[i]; };
    if (tid == 0) sm.cursor = 0;
    SweepG st;
    sweep_init_g(st, sm, with_eq);
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t rm = tid < 192 ? sm.rows[tid] : 0u;   // from build_masks_g
    // first log entry of every plane = exclusive prefix of the planes' non-empty rows.  Every warp adds up the six
    // faces for itself (six loads, six warp reductions): no barrier, where a block-wide scan needs two.
    uint32_t nL = 0, seg = 0;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        const uint32_t tot = __reduce_add_sync(kFull, (uint32_t)__popc(sm.rows[f * 32 + lane]));
        nL += tot;
        if (f < warp) seg += tot;
    }
    seg += warp_incl_scan((uint32_t)__popc(rm), lane) - (uint32_t)__popc(rm);
    if (tid < 192) sm.seg[tid] = seg;
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
        if (tid < 192) births = sweep_log_g(st, rm, (uint32_t)tid, log.entry + seg);
#ifdef VXP_PROFILE
        if (lane == 0 && blockIdx.x < kProfCtas) g_prof_warp[blockIdx.x][warp] = (unsigned int)(clock64() - t_sw);
#endif
        // ordinal of every plane's first run: block-wide exclusive scan of the births, with one barrier (the log
        // is published by it too; sm.scratch is not written again before the chunk ends)
        const uint32_t incl = warp_incl_scan(births, lane);
        if (lane == 31) sm.scratch[warp] = incl;
        __syncthreads();
        uint32_t first_run = incl - births;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t t = sm.scratch[w];
            if (w < warp) first_run += t;
            Q += t;
        }
        if (tid < 192) sm.first_run[tid] = first_run;
        if (tid == 0) f = reserve_quads(sm, out, Q);
        prof.mark(5);
        // The records go below the log if they fit; otherwise - the sweep being over, its V rows are dead - the first
        // 2 x 1056 of them into EY / EZ (contiguous) and the rest from vt (which sits right below the staging area)
        // up to the log.
        RecStore rec{staging, Q, staging};            // block-uniform
        bool fits = Q <= kCap - 2 * nL;
        if (!fits) {
            rec = RecStore{sm.eq[1], 2u * kPlaneWords, sm.tile.post.vt};
            fits = Q <= 2u * kPlaneWords + (uint32_t)kPlaneWords + kCap - 2 * nL;
        }
        if (fits) {
            __syncthreads();                          // first_run published; the sweep (and with it every read of a V row) is over
            stage_births_g(sm, with_eq, log, nL, rec);
            __syncthreads();
            prof.mark(6);
            flush_staging(sm, [rec, &log](uint32_t i) { const uint32_t r = rec.at(i); return r | (run_height(log, r) - 1u) << 23; },
                          Q, 0, out, id_of, [&] { if (tid == 0) publish_range(sm, mi, Q, f, out); }, true);
            prof.mark(7);
            return;
        }
        // More records than fit at once, but the log is there: stage and flush them in windows of ordinals (no second sweep).
        const uint32_t window = 3u * kPlaneWords + kCap - 2 * nL;     // rec is the split store
        if (window >= 1024u) {
            prof.count(14);
            for (uint32_t lo = 0; lo < Q; lo += window) {
                const uint32_t hi = min(Q, lo + window);
                __syncthreads();                      // first round: first_run published, sweep over; later: previous flush done
                stage_births_g(sm, with_eq, log, nL, rec, lo, hi);
                __syncthreads();
                bool ok;
                if (lo == 0) ok = flush_staging(sm, [rec, &log](uint32_t i) { const uint32_t r = rec.at(i); return r | (run_height(log, r) - 1u) << 23; },
                                                hi, 0, out, id_of, [&] { if (tid == 0) publish_range(sm, mi, Q, f, out); }, true);
                else ok = flush_staging(sm, [rec, &log](uint32_t i) { const uint32_t r = rec.at(i); return r | (run_height(log, r) - 1u) << 23; },
                                        hi - lo, lo, out, id_of, nothing, false);
                if (!ok) return;
            }
            prof.mark(7);
            return;
        }
    } else {
        // the log does not fit: count the quads with a sweep that logs nothing (every run that starts is one), then emit by halves
        uint32_t births = 0;
        if (tid < 192) births = sweep_log_g<false>(st, rm, (uint32_t)tid, nullptr);
        const uint32_t incl = warp_incl_scan(births, lane);
        if (lane == 31) sm.scratch[warp] = incl;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kWarps; ++w) Q += sm.scratch[w];
        if (tid == 0) f = reserve_quads(sm, out, Q);
    }
    // the log or the records do not fit: the planes in two halves if that fits, else a stepped sweep that stages
    // through a cursor, half of the planes per step, flushing in rounds
    prof.count(14);
    if (tid == 0) publish_range(sm, mi, Q, f, out);
    __syncthreads();                                  // sm.first published; everybody is done with the log and the cursor
    if (emit_greedy_halves(sm, with_eq, out, id_of)) {
        prof.mark(7);
        return;
    }
    sweep_init_g(st, sm, with_eq);
    uint32_t flushed = 0;
    if (tid == 0) sm.cursor = 0;
#pragma unroll 1
    for (int v = 0; v <= 32; ++v) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            __syncthreads();
            if (tid < 192 && (tid >= 96) == (half == 1)) sweep_step_g(st, v, staging, kCap, &sm.cursor);
            __syncthreads();
            const uint32_t cnt = sm.cursor;
            if (cnt + 96 * 32 > kCap || (v == 32 && half == 1)) {
                if (!flush_staging(sm, staged, cnt, flushed, out, id_of, nothing, false)) return;
                flushed += cnt;
                __syncthreads();
                if (tid == 0) sm.cursor = 0;
            }
        }
    }
    prof.mark(7);
}

// ------------------------------------------------------------------ one chunk
__device__ __forceinline__ void mesh_chunk_g(SmemG& sm, const uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6,
                                             uint32_t n, uint32_t chunk, uint32_t mi, const MeshOut& out, const Bnd& bnd) {
    const int tid = threadIdx.x;
    const uint16_t* gvox = voxels + (size_t)chunk * VXP_CHUNK_VOXELS;
    Prof prof;
    prof.start(tid == 0);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRingG; ++s) issue_ring_slab(sm, gvox, s, s);
    }
    const int nb = (nbr6 && tid < 192) ? __ldg(nbr6 + 6 * (size_t)chunk + (tid >> 5)) : -1;
    HaloRows rows;
    issue_yz_halo(rows, voxels, nb);

    uint32_t v_or = 0, v_and = kFull, o_and = kFull;
    IdPlanes ip;
#pragma unroll
    for (int s = 0; s < kSlabs; ++s) {
        mbar_wait(&sm.mbar[s % kRingG], (uint32_t)(s / kRingG) & 1u);
        convert_ring_slab(sm, s, s % kRingG, v_or, v_and, o_and, ip);
        if (s + kRingG < kSlabs) {
            __syncthreads();                            // the buffer has been read by everybody
            if (tid == 0) issue_ring_slab(sm, gvox, s + kRingG, s % kRingG);
        }
    }
    prof.mark(0);
    const ChunkFlags fl = reduce_flags_g(sm, v_or, v_and, o_and);
    prof.mark(1);
    if (bnd.tags) publish_boundary(sm, bnd, chunk, fl);
    if (!fl.any_solid) {
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(8);
        return;
    }
    const bool with_eq = !fl.single_id;
    const bool with_planes = with_eq && fl.small_ids;
    XPoll xp;
    const bool air_yz = halo_begin(sm, nb, n, bnd, rows, xp);
    if (with_planes) store_id_planes_g(sm, ip);         // while the x neighbours' boundary planes are on their way
    const bool air_x = halo_finish(sm, voxels, nb, n, bnd, xp, prof);
    const bool halo_air = __syncthreads_or(air_yz || air_x);   // barrier: halo (and id planes) published
    prof.mark(2);
    if (fl.all_solid && !halo_air) {
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(9);
        return;
    }
    if (with_eq && !with_planes) {
        build_active(sm);
        __syncthreads();
        equality_second_pass(sm, gvox);                 // ends with a barrier
    }
    if (tid < 192) sm.rows[tid] = 0u;
    __syncthreads();
    prof.mark(3);
    const bool any = __syncthreads_or(build_masks_g(sm, with_eq, with_planes) != 0);
    prof.mark(4);
    if (!any) {
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(10);
        return;
    }
    prof.count(11);
    emit_greedy(sm, mi, with_eq, out, [gvox](int idx) -> uint32_t { return __ldg(gvox + idx); }, prof);
}

// Meshes chunks [chunk0, hi) of a batch of n, one chunk per CTA (grid = hi - chunk0): the hardware block scheduler
// is the load balancer.  (Persistent CTAs drawing tickets were measured 7-12 % slower: a CTA that draws its next
// chunk early delays that chunk's boundary planes for the neighbours that poll them.)  Launch chaining: Chain.
__global__ void __launch_bounds__(kThreads, kGreedyCtas)
mesh_kernel_g(const uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6, uint32_t n, uint32_t chunk0, uint32_t hi,
              MeshOut out, Bnd bnd, Chain chain) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemG& sm = *reinterpret_cast<SmemG*>(smem_raw);
    const bool dealt = chain.runs != 0u && (blockIdx.x >> 4) < chain.runs;
    if (dealt && threadIdx.x == 0)
        sm.cursor = (uint32_t)(((unsigned long long)(blockIdx.x >> 4) * chain.run_stride) % chain.runs) * 16u + (blockIdx.x & 15u);
    const bool closer = chain.user_total != nullptr && blockIdx.x == gridDim.x - 1;
    if (threadIdx.x == 0) {
        if (closer) griddep_wait();
        sm.ones = kFull;
        sm.zero = 0u;
        sm.counted = 0u;
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        mbar_init(&sm.mbar[2], 1);
        fence_mbar_init();
    }
    __syncthreads();
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
    if (active) mesh_chunk_g(sm, voxels, nbr6, n, chunk, mi, out, bnd);
#ifdef VXP_PROFILE
    if (threadIdx.x == 0 && blockIdx.x < kProfCtas) g_prof_time[blockIdx.x][1] = prof_now();
#endif
    if (threadIdx.x == 0 && chain.user_total != nullptr) {
        if (!sm.counted) red_add_u64(out.total, kCtaOne);
        if (closer) {
            unsigned long long t;
            while (((t = ld_relaxed_u64(out.total)) >> kQuadBits) != gridDim.x) __nanosleep(64);
            *chain.user_total = t & kQuadMask;
            st_relaxed_u64(out.total, 0ull);
        }
    }
}

// ====================================================================== culled meshing, streamed: five CTAs per SM
// The culled kernel needs no ids at all while it scans (block ids are gathered from the chunk in L2 when the
// quads are written), so it streams the chunk through the same two-slab ring and keeps only the occupancy.  The
// non-empty face words are listed in a window of 2048 words (a terrain surface chunk has ~1500 non-
// empty face words; a chunk with more is listed and flushed window by window) next to the 4096-record staging area:
// 32 KiB in the ring + occupancy = 37.6 KiB per CTA.
struct PostC {
    static constexpr int kWordCap = 2048;
    static constexpr int kCap = (2 * kSlabBytes - kWordCap * 8) / 4;   // 4096 staged quad records
    uint2 words[kWordCap];       // .x = f<<10 | y<<5 | z | (ordinal of the word's first quad) << 13, .y = bits over x
    uint32_t staging[kCap];
};
struct __align__(128) SmemC {
    union {
        unsigned char ring[2][kSlabBytes];
        PostC post;
    } tile;
    uint32_t occ[34 * 34];
    Aux aux;                       // fused path: heights
    uint32_t xh[2][32];
    uint32_t scratch[kWarps + 1];
    uint32_t cursor;
    uint32_t counted;
    unsigned long long first;
    unsigned long long mbar[2];
};
static_assert(sizeof(PostC) <= 2 * kSlabBytes, "post layout overflows the ring");
static_assert(kCulledCtas * (sizeof(SmemC) + 128 + 1024) <= 227 * 1024, "shared memory of the culled kernels exceeds the CTAs-per-SM target");

__device__ __forceinline__ void issue_ring_slab_c(SmemC& sm, const uint16_t* gvox, int s) {
    mbar_expect_tx(&sm.mbar[s & 1], kSlabBytes);
    tma_load_1d(sm.tile.ring[s & 1], reinterpret_cast<const unsigned char*>(gvox) + (size_t)s * kSlabBytes, kSlabBytes,
                &sm.mbar[s & 1]);
}
// Slab s -> occupancy bytes; v_or / o_and accumulate "any voxel" / "all voxels" per thread.
__device__ __forceinline__ void convert_ring_slab_c(SmemC& sm, int s, uint32_t& v_or, uint32_t& o_and) {
    // one 64-byte row per thread, read and stored as in convert_ring_slab (greedy)
    const int lane = threadIdx.x & 31;
    const int j = (lane >> 1) & 3;
    const int z = convert_row_z(s), y = convert_row_y();
    const uint4* row = reinterpret_cast<const uint4*>(sm.tile.ring[s & 1]) + ((z & 7) * 32 + y) * 4;
    uint4 q[4];
    uint32_t slab_or = 0, slab_and = kFull;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q[k] = row[(k + j) & 3];
        slab_or |= q[k].x | q[k].y | q[k].z | q[k].w;
        slab_and &= q[k].x & q[k].y & q[k].z & q[k].w;
    }
    v_or |= slab_or;
    // the warp's 8 x 4 rows hold one value (air, or one kind of rock): nothing to gather bit by bit
    const uint32_t first = __shfl_sync(kFull, slab_or, 0);
    const bool same = slab_or == slab_and && (slab_or & 0xFFFFu) == (slab_or >> 16) && slab_or == first;
    uint32_t ob = 0;
    if (__all_sync(kFull, same)) {
        ob = slab_or ? kFull : 0u;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) ob |= nz8(q[k]) << (8 * k);
        ob = __funnelshift_l(ob, ob, 8u * (uint32_t)j);
    }
    o_and &= ob;
    sm.occ[(z + 1) * 34 + y + 1] = ob;
}

// One packed block scan gives every thread its first ordinal and its slot in the word list; the non-empty words are
// listed (window by window, PostC::kWordCap words at a time) with the ordinal of their first quad; the threads
// share the list evenly when they stage the quad records.
template <class IdFn>
__device__ __forceinline__ void culled_emit_c(SmemC& sm, uint32_t mi, uint32_t packed, const MeshOut& out, IdFn id_of, Prof& prof) {
    const int tid = threadIdx.x;
    constexpr uint32_t kCap = PostC::kCap, kWordCap = PostC::kWordCap;
    PostC& P = sm.tile.post;
    uint32_t* const staging = P.staging;
    auto nothing = [] {};
    auto staged = [staging](uint32_t i) { return staging[i]; };
    const int y = tid & 31, w = tid >> 5;
    uint32_t tot;
    const uint32_t pre = block_excl_scan(packed, sm.scratch, &tot);
    const uint32_t Q = tot & 0x1FFFFu, nwords = tot >> 17;
    if (Q == 0) {
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(10);
        return;
    }
    prof.count(11);
    unsigned long long fq = 0;
    if (tid == 0) fq = reserve_quads(sm, out, Q);   // in flight while the records are staged
    bool first_flush = true;
    for (uint32_t wbase = 0; wbase < nwords; wbase += kWordCap) {
        const uint32_t wend = min(nwords, wbase + kWordCap);
        if (wbase) __syncthreads();                 // the previous window's words have been expanded
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
                    if (wd[f]) {
                        if (slot >= wbase && slot < wend) P.words[slot - wbase] = make_uint2((uint32_t)(f << 10 | y << 5 | z) | ord << 13, wd[f]);
                        ++slot;
                    }
                    ord += __popc(wd[f]);
                }
            }
        }
        __syncthreads();
        prof.mark(5);
        const uint32_t wcount = wend - wbase;
        const uint32_t ord_lo = P.words[0].x >> 13;
        const uint2 lastw = P.words[wcount - 1];
        const uint32_t ord_hi = (lastw.x >> 13) + __popc(lastw.y);
        for (uint32_t base = ord_lo; base < ord_hi; base += kCap) {
            const uint32_t end = min(ord_hi, base + kCap);
            for (uint32_t i = tid; i < wcount; i += kThreads) {     // neighbours in the list are words of similar density
                const uint2 d = P.words[i];
                uint32_t bits = d.y, o = d.x >> 13;
                if (o + __popc(bits) <= base || o >= end) continue;
                const uint32_t f = (d.x >> 10) & 7u, yy = (d.x >> 5) & 31u, zz = d.x & 31u;
                // u | v<<5 | s<<10 | f<<15:  X planes (s,u,v) = (x,y,z);  Y planes (y,x,z);  Z planes (z,x,y)
                const uint32_t rb = f << 15 | (f < 2 ? (yy | zz << 5) : f < 4 ? (zz << 5 | yy << 10) : (yy << 5 | zz << 10));
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
            if (first_flush) fits = flush_staging(sm, staged, end - base, base, out, id_of, [&] { if (tid == 0) publish_range(sm, mi, Q, fq, out); }, true);
            else fits = flush_staging(sm, staged, end - base, base, out, id_of, nothing, false);
            first_flush = false;
            if (!fits) return;
            __syncthreads();
            prof.mark(7);
        }
    }
}

__device__ __forceinline__ void mesh_chunk_c(SmemC& sm, const uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6,
                                             uint32_t n, uint32_t chunk, uint32_t mi, const MeshOut& out, const Bnd& bnd) {
    const int tid = threadIdx.x;
    const uint16_t* gvox = voxels + (size_t)chunk * VXP_CHUNK_VOXELS;
    Prof prof;
    prof.start(tid == 0);
    if (tid == 0) {
        issue_ring_slab_c(sm, gvox, 0);
        issue_ring_slab_c(sm, gvox, 1);
    }
    const int nb = (nbr6 && tid < 192) ? __ldg(nbr6 + 6 * (size_t)chunk + (tid >> 5)) : -1;
    HaloRows rows;
    issue_yz_halo(rows, voxels, nb);
    uint32_t v_or = 0, o_and = kFull;
#pragma unroll 1
    for (int s = 0; s < kSlabs; ++s) {
        mbar_wait(&sm.mbar[s & 1], (uint32_t)(s >> 1) & 1u);
        convert_ring_slab_c(sm, s, v_or, o_and);
        if (s + 2 < kSlabs) {
            __syncthreads();                            // the buffer has been read by everybody
            if (tid == 0) issue_ring_slab_c(sm, gvox, s + 2);
        }
    }
    prof.mark(0);
    ChunkFlags fl;
    fl.any_solid = __syncthreads_or(v_or != 0u);        // barrier: occupancy interior published, ring dead
    fl.all_solid = __syncthreads_and(o_and == kFull);
    fl.single_id = fl.small_ids = false;
    prof.mark(1);
    if (bnd.tags) publish_boundary(sm, bnd, chunk, fl);
    if (!fl.any_solid) {
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(8);
        return;
    }
    auto id_of = [gvox](int idx) -> uint32_t { return __ldg(gvox + idx); };
    XPoll xp;
    if (!fl.all_solid) {
        // count the faces that do not depend on the x halo (+-Y, +-Z) while the x neighbours' boundary planes are on their way
        halo_begin(sm, nb, n, bnd, rows, xp);
        __syncthreads();                                // +-Y / +-Z halo rows published
        uint32_t packed = culled_count<false>(sm);
        halo_finish(sm, voxels, nb, n, bnd, xp, prof);
        __syncthreads();                                // x halo published
        prof.mark(2);
        packed += culled_count<true>(sm);
        culled_emit_c(sm, mi, packed, out, id_of, prof);
        return;
    }
    const bool air_yz = halo_begin(sm, nb, n, bnd, rows, xp);
    const bool air_x = halo_finish(sm, voxels, nb, n, bnd, xp, prof);
    const bool halo_air = __syncthreads_or(air_yz || air_x);   // barrier: halo published
    prof.mark(2);
    if (!halo_air) {                                    // buried: no faces
        if (tid == 0) out.meta[mi] = vxp_chunk_meta{0u, 0u};
        prof.count(9);
        return;
    }
    culled_emit_c(sm, mi, culled_count<false>(sm) + culled_count<true>(sm), out, id_of, prof);
}

// Same contract and launch chaining as mesh_kernel_g.
// CTAs per SM the kernel is compiled for (measured on config 2: 3 with a whole-tile layout: 69.5 us; 4: 65.4 us at 61 registers;
// 5: 62.4 us at 48 registers, no spills; the shared memory would allow six, the registers would not)
__global__ void __launch_bounds__(kThreads, kCulledCtas)
mesh_kernel_c(const uint16_t* __restrict__ voxels, const int32_t* __restrict__ nbr6, uint32_t n, uint32_t chunk0, uint32_t hi,
              MeshOut out, Bnd bnd, Chain chain) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemC& sm = *reinterpret_cast<SmemC*>(smem_raw);
    const bool dealt = chain.runs != 0u && (blockIdx.x >> 4) < chain.runs;
    if (dealt && threadIdx.x == 0)
        sm.cursor = (uint32_t)(((unsigned long long)(blockIdx.x >> 4) * chain.run_stride) % chain.runs) * 16u + (blockIdx.x & 15u);
    const bool closer = chain.user_total != nullptr && blockIdx.x == gridDim.x - 1;
    if (threadIdx.x == 0) {
        if (closer) griddep_wait();
        sm.counted = 0u;
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        mbar_init(&sm.mbar[2], 1);
        fence_mbar_init();
    }
    __syncthreads();
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
    if (active) mesh_chunk_c(sm, voxels, nbr6, n, chunk, mi, out, bnd);
#ifdef VXP_PROFILE
    if (threadIdx.x == 0 && blockIdx.x < kProfCtas) g_prof_time[blockIdx.x][1] = prof_now();
#endif
    if (threadIdx.x == 0 && chain.user_total != nullptr) {
        if (!sm.counted) red_add_u64(out.total, kCtaOne);
        if (closer) {
            unsigned long long t;
            while (((t = ld_relaxed_u64(out.total)) >> kQuadBits) != gridDim.x) __nanosleep(64);
            *chain.user_total = t & kQuadMask;
            st_relaxed_u64(out.total, 0ull);
        }
    }
}

} // namespace mesh
} // namespace vxp
