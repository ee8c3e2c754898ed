// vxp_pack.cu — SPEC §10 chunk wire format on the GPU: palette + bit-packed indices.
// One CTA per chunk for both directions.  Pure integer / byte work, HBM-bound on the raw side.
#include <cuda_runtime.h>

#include "vxp_launch.h"
#include "vxp_device.cuh"

namespace vxp {

namespace {
constexpr uint32_t kMagic = 0x31505856u;      // "VXP1"
constexpr int kHeaderBytes = 48;              // 16-byte header + 16-entry palette
constexpr int kSetSlots = 64;
constexpr int kCursorBits = 42;                // byte cursor of a pack launch: 4 TiB, 2^22 chunks
constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;

struct PackShared {
    uint32_t set[kSetSlots];     // open-addressing set of the ids met so far
    uint32_t distinct;           // may run past 17: only "more than 16" matters
    uint32_t palette[16];        // ascending
    uint32_t len, bits;
    unsigned long long offset;
};

__device__ __forceinline__ void set_insert(PackShared& sh, uint32_t id) {
    uint32_t slot = (id * 0x9E3779B1u) >> 26;                    // 64 slots
    for (int probe = 0; probe < kSetSlots; ++probe) {
        const uint32_t old = atomicCAS(&sh.set[slot], kEmptySlot, id);
        if (old == id) return;
        if (old == kEmptySlot) { atomicAdd(&sh.distinct, 1u); return; }
        if (sh.distinct > 16u) return;                             // the chunk goes out raw anyway
        slot = (slot + 1) & (kSetSlots - 1);
    }
}

// index of id v in the ascending palette = number of entries below it
__device__ __forceinline__ uint32_t palette_index(const uint32_t* pal, uint32_t len, uint32_t v) {
    uint32_t idx = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) idx += (j < (int)len && pal[j] < v) ? 1u : 0u;
    return idx;
}
}  // namespace

// One CTA per chunk, two passes over its 64 KiB (the second one hits L2).
// Pass 1 finds the chunk's distinct ids.  Block ids are small numbers in practice, so the set is first collected as a
// 64-bit mask (bit v = id v occurs): a thread ORs its 128 voxels into two registers - one shift per voxel, one per group
// of eight equal voxels - a warp reduction and two shared-memory atomics per warp merge them, and both the ascending
// palette SPEC §10 asks for and every id's index in it fall out of the mask (rank of a bit = population count below it)
// with no search at all.  A chunk with an id >= 64 takes the general path: an open-addressing set in shared memory, a
// rank sort of its at most 16 values, a palette search per run of equal voxels.
// Pass 2 turns 8 voxels into `bits` bytes.
__global__ void __launch_bounds__(kThreads) pack_chunks_kernel(const uint16_t* __restrict__ voxels, uint32_t n,
                                                                unsigned char* __restrict__ packed, unsigned long long capacity,
                                                                unsigned long long* __restrict__ offsets,
                                                                unsigned long long* __restrict__ cursor,
                                                                unsigned long long* __restrict__ total) {
    __shared__ PackShared sh;
    __shared__ uint32_t s_mask[2], s_large;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n) return;
    const uint4* src = reinterpret_cast<const uint4*>(voxels + (size_t)chunk * VXP_CHUNK_VOXELS);
    if (tid < kSetSlots) sh.set[tid] = kEmptySlot;
    if (tid == 0) { sh.distinct = 0; s_mask[0] = s_mask[1] = 0u; s_large = 0u; }
    __syncthreads();
    // pass 1, small ids: the mask
    {
        uint32_t mlo = 0, mhi = 0, all = 0;
        auto add = [&](uint32_t v) {
            const uint32_t bit = 1u << (v & 31u);
            if (v & 32u) mhi |= bit; else mlo |= bit;
        };
#pragma unroll 1
        for (int k0 = 0; k0 < 16; k0 += 8) {                       // eight 16-byte loads in flight per thread (the pass is latency-bound otherwise)
            uint4 qq[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) qq[k] = __ldg(src + tid + (k0 + k) * kThreads);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 q = qq[k];
                all |= q.x | q.y | q.z | q.w;
                if (q.x == q.y && q.x == q.z && q.x == q.w && (q.x & 0xFFFFu) == (q.x >> 16)) {
                    add(q.x & 0x3Fu);                              // eight equal voxels (ids >= 64 are caught by `all`)
                } else {
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) { add(w[e] & 0x3Fu); add((w[e] >> 16) & 0x3Fu); }
                }
            }
        }
        mlo = __reduce_or_sync(kFull, mlo);
        mhi = __reduce_or_sync(kFull, mhi);
        all = __reduce_or_sync(kFull, all);
        if (lane == 0) {
            if (mlo) atomicOr(&s_mask[0], mlo);
            if (mhi) atomicOr(&s_mask[1], mhi);
            if ((all | (all >> 16)) & 0xFFC0u) atomicOr(&s_large, 1u);
        }
    }
    __syncthreads();
    const uint32_t mlo = s_mask[0], mhi = s_mask[1];
    const uint32_t plo = (uint32_t)__popc(mlo);
    bool small = s_large == 0u && plo + (uint32_t)__popc(mhi) <= 16u;   // block-uniform
    if (tid < 16) sh.palette[tid] = 0;
    if (!small) {
        // general path: the set of ids (a thread skips repeats of the value it inserted last: chunks are mostly runs)
        uint32_t last = kEmptySlot;
#pragma unroll 1
        for (int k = 0; k < 16; ++k) {
            const uint4 q = __ldg(src + tid + k * kThreads);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t a = w[e] & 0xFFFFu, b = w[e] >> 16;
                if (a != last) { set_insert(sh, a); last = a; }
                if (b != last) { set_insert(sh, b); last = b; }
            }
        }
    }
    __syncthreads();
    const uint32_t distinct = small ? plo + (uint32_t)__popc(mhi) : sh.distinct;
    const bool raw = distinct > 16u;
    if (small) {
        if (tid < 64) {                                            // palette entry of rank r = the r-th set bit of the mask
            const uint32_t m = tid < 32 ? mlo : mhi, below = m & ((1u << (tid & 31)) - 1u);
            if ((m >> (tid & 31)) & 1u) sh.palette[(tid < 32 ? 0u : plo) + (uint32_t)__popc(below)] = (uint32_t)tid;
        }
    } else if (!raw && tid < kSetSlots) {                          // rank sort of at most 16 values
        const uint32_t v = sh.set[tid];
        if (v != kEmptySlot) {
            uint32_t rank = 0;
            for (int j = 0; j < kSetSlots; ++j) rank += (sh.set[j] != kEmptySlot && sh.set[j] < v) ? 1u : 0u;
            sh.palette[rank] = v;
        }
    }
    if (tid == 0) {
        const uint32_t bits = raw ? 16u : distinct <= 1u ? 0u : distinct <= 2u ? 1u : distinct <= 4u ? 2u : 4u;
        const unsigned long long size = kHeaderBytes + 4096ull * bits;
        // one atomic per chunk reserves the record and checks the CTA in (count above bit 42): the last one to reserve
        // knows the batch's size, hands it to the caller and leaves the cursor zeroed - no memset in front of the launch
        const unsigned long long old = atomicAdd(cursor, size | 1ull << kCursorBits);
        const unsigned long long off = old & ((1ull << kCursorBits) - 1ull);
        if ((old >> kCursorBits) == n - 1u) {
            *total = off + size;
            *cursor = 0ull;
        }
        sh.bits = bits;
        sh.len = raw ? 0u : distinct;
        sh.offset = off + size <= capacity ? off : ~0ull;
        offsets[chunk] = sh.offset;
    }
    __syncthreads();
    if (sh.offset == ~0ull) return;                                // does not fit: *total tells the caller how much is needed
    unsigned char* rec = packed + sh.offset;
    const uint32_t bits = sh.bits, len = sh.len;
    if (tid == 0) {
        uint32_t* h = reinterpret_cast<uint32_t*>(rec);
        h[0] = kMagic;
        h[1] = bits | len << 16;
        h[2] = 4096u * bits;
        h[3] = 0u;
    }
    if (tid < 8) reinterpret_cast<uint32_t*>(rec + 16)[tid] = sh.palette[2 * tid] | sh.palette[2 * tid + 1] << 16;
    if (bits == 0u) return;
    unsigned char* data = rec + kHeaderBytes;
    // index of id v in the palette: small ids: the number of mask bits below bit v
    auto index_of = [&](uint32_t v) -> uint32_t {
        if (small) return ((v & 32u) ? plo : 0u) + (uint32_t)__popc(((v & 32u) ? mhi : mlo) & ((1u << (v & 31u)) - 1u));
        return palette_index(sh.palette, len, v);
    };
    const uint32_t spread = bits == 1u ? 0xFFu : bits == 2u ? 0x5555u : 0x11111111u;
    // pass 2 (L2 hits): 8 voxels -> 8 indices -> `bits` bytes
#pragma unroll 1
    for (int k0 = 0; k0 < 16; k0 += 4) {
    uint4 qq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) qq[k] = __ldg(src + tid + (k0 + k) * kThreads);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = tid + (k0 + k) * kThreads;
        const uint4 q = qq[k];
        if (bits == 16u) { reinterpret_cast<uint4*>(data)[j] = q; continue; }
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t out = 0;
        // chunks are mostly runs of one id: look an id up only when it differs from the previous voxel's
        uint32_t lastv = w[0] & 0xFFFFu, lasti = index_of(lastv);
        if (q.x == q.y && q.x == q.z && q.x == q.w && lastv == (q.x >> 16)) {   // eight equal voxels: one look-up, one multiply
            out = lasti * spread;
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t v = (w[t >> 1] >> (16 * (t & 1))) & 0xFFFFu;
                if (v != lastv) { lastv = v; lasti = index_of(v); }
                out |= lasti << (bits * t);
            }
        }
        if (bits == 1u) data[j] = (unsigned char)out;
        else if (bits == 2u) reinterpret_cast<unsigned short*>(data)[j] = (unsigned short)out;
        else reinterpret_cast<uint32_t*>(data)[j] = out;
    }
    }
}

// status[0] is set to 1 if any record is malformed (offset outside the buffer - vxp_pack_chunks marks records that did
// not fit with ~0 -, misaligned, bad magic / bits / palette length, data running past the end); such chunks come out as air.
// One CTA per chunk, eight CTAs per SM: offset -> header -> data is a chain of three dependent global round trips in front
// of 64 KiB of stores, which only other resident chunks can hide.  Every thread reads and validates the header itself
// (the same words for all of them: one broadcast load, no shared-memory round for it); only the palette goes through
// shared memory.  (Persistent CTAs prefetching the next record's header were measured slower, 78 us against 71 us per
// 4096 chunks: records cost between 16 stores and a full decode, and a static stride over the batch balances them badly.)
struct UnpackHead {
    unsigned long long off;
    uint32_t h0, h1, h2, pal;   // magic, bits | len << 16, data bytes, palette word (threads 0..7)
    bool in_buffer;
};
__device__ __forceinline__ UnpackHead unpack_fetch(const unsigned char* __restrict__ packed, unsigned long long packed_bytes,
                                                   const unsigned long long* __restrict__ offsets, uint32_t chunk, uint32_t n) {
    UnpackHead u{};
    if (chunk >= n) return u;
    u.off = __ldg(offsets + chunk);
    u.in_buffer = (u.off & 15ull) == 0ull && u.off <= packed_bytes && packed_bytes - u.off >= (unsigned long long)kHeaderBytes;
    if (u.in_buffer) {
        const uint32_t* h = reinterpret_cast<const uint32_t*>(packed + u.off);
        u.h0 = __ldg(h);
        u.h1 = __ldg(h + 1);
        u.h2 = __ldg(h + 2);
        if (threadIdx.x < 8) u.pal = __ldg(h + 4 + threadIdx.x);
    }
    return u;
}
__global__ void __launch_bounds__(kThreads, 8) unpack_chunks_kernel(const unsigned char* __restrict__ packed, unsigned long long packed_bytes,
                                                                     const unsigned long long* __restrict__ offsets, uint32_t n,
                                                                     uint16_t* __restrict__ voxels, uint32_t* __restrict__ status) {
    __shared__ uint32_t s_pal[1][16];
    const int tid = threadIdx.x;
    const uint32_t chunk = blockIdx.x;
    constexpr int kPer = 16;                                       // 16-byte stores per thread
    if (chunk >= n) return;
    const UnpackHead cur = unpack_fetch(packed, packed_bytes, offsets, chunk, n);
    constexpr int buf = 0;
    {
        // every thread validates the header it read itself (the same words for all of them): no broadcast, no barrier for it
        uint32_t bits = cur.h1 & 0xFFFFu;
        const uint32_t len = cur.h1 >> 16;
        const bool ok = cur.in_buffer && cur.h0 == kMagic && (bits == 0u || bits == 1u || bits == 2u || bits == 4u || bits == 16u) &&
                        cur.h2 == 4096u * bits && (bits == 16u ? len == 0u : (len >= 1u && len <= 16u && len <= (1u << bits))) &&
                        packed_bytes - cur.off - (unsigned long long)kHeaderBytes >= (unsigned long long)cur.h2;
        if (!ok) {
            if (tid == 0) atomicExch(status, 1u);
            bits = 0xFFu;
        }
        if (tid < 8) {
            s_pal[buf][2 * tid] = cur.pal & 0xFFFFu;
            s_pal[buf][2 * tid + 1] = cur.pal >> 16;
        }
        __syncthreads();                                       // palette published
        const uint32_t* pal = s_pal[buf];
        uint4* dst = reinterpret_cast<uint4*>(voxels + (size_t)chunk * VXP_CHUNK_VOXELS);
        const unsigned char* data = packed + (ok ? cur.off : 0ull) + kHeaderBytes;
        if (bits == 0u || bits == 0xFFu) {
            const uint32_t b = bits == 0u ? pal[0] : 0u, bb = b | b << 16;
#pragma unroll
            for (int k = 0; k < kPer; ++k) st_voxels16(dst + tid + k * kThreads, make_uint4(bb, bb, bb, bb));
        } else if (bits == 16u) {
#pragma unroll
            for (int k = 0; k < kPer; ++k) st_voxels16(dst + tid + k * kThreads, __ldg(reinterpret_cast<const uint4*>(data) + tid + k * kThreads));
        } else {
            const uint32_t mask = (1u << bits) - 1u;
            constexpr int kFlight = 8;                             // loads in flight per thread, then as many stores
#pragma unroll 1
            for (int k0 = 0; k0 < kPer; k0 += kFlight) {
                uint32_t in[kFlight];
#pragma unroll
                for (int k = 0; k < kFlight; ++k) {
                    const int j = tid + (k0 + k) * kThreads;
                    in[k] = bits == 1u ? data[j] : bits == 2u ? reinterpret_cast<const unsigned short*>(data)[j]
                                                              : reinterpret_cast<const uint32_t*>(data)[j];
                }
#pragma unroll
                for (int k = 0; k < kFlight; ++k) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = pal[(in[k] >> (bits * (2 * e))) & mask] | pal[(in[k] >> (bits * (2 * e + 1))) & mask] << 16;
                    st_voxels16(dst + tid + (k0 + k) * kThreads, make_uint4(w[0], w[1], w[2], w[3]));
                }
            }
        }
    }
}

cudaError_t launch_pack_chunks(const uint16_t* d_voxels, uint32_t n, unsigned char* d_packed, unsigned long long capacity,
                               unsigned long long* d_offsets, unsigned long long* d_total, unsigned long long* d_cursor,
                               cudaStream_t stream) {
    if (n == 0) return cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream);
    if (n > (1u << 22)) return cudaErrorInvalidValue;
    pack_chunks_kernel<<<n, kThreads, 0, stream>>>(d_voxels, n, d_packed, capacity, d_offsets, d_cursor, d_total);
    return cudaGetLastError();
}

cudaError_t launch_unpack_chunks(const unsigned char* d_packed, unsigned long long packed_bytes, const unsigned long long* d_offsets,
                                 uint32_t n, uint16_t* d_voxels, uint32_t* d_status, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;                    // d_status is sticky: the caller clears it when it has read a verdict
    unpack_chunks_kernel<<<n, kThreads, 0, stream>>>(d_packed, packed_bytes, d_offsets, n, d_voxels, d_status);
    return cudaGetLastError();
}

}  // namespace vxp

// ====================================================================== LOD (SPEC §11)
// One output chunk = 2 x 2 x 2 stored chunks at half resolution.  Four CTAs per output chunk (a quarter of its z-layers
// each: a grid of whole output chunks is a few hundred CTAs for a typical batch, less than one even wave); a thread
// produces 8 output voxels (one 16-byte store) from 4 rows x 16 input voxels.  HBM-bound: 8 bytes read per byte written.
namespace vxp {
namespace {
// first non-air of (a, b) per u16 lane pair -> one u16: a if a != 0 else b
__device__ __forceinline__ uint32_t first_solid(uint32_t a, uint32_t b) { return a ? a : b; }
}  // namespace

__global__ void __launch_bounds__(kThreads) downsample_chunks_kernel(const uint16_t* __restrict__ voxels,
                                                                      const int32_t* __restrict__ child8, uint32_t n_out,
                                                                      uint16_t* __restrict__ out) {
    const uint32_t oc = blockIdx.x >> 2, part = blockIdx.x & 3u;
    if (oc >= n_out) return;
    __shared__ int s_child[8];
    if (threadIdx.x < 8) s_child[threadIdx.x] = child8[8 * (size_t)oc + threadIdx.x];
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)oc * VXP_CHUNK_VOXELS);
    for (int j = part * 1024 + threadIdx.x; j < (int)(part + 1) * 1024; j += kThreads) {   // j = (z*32 + y)*4 + p
        const int z = j >> 7, y = (j >> 2) & 31, p = j & 3;        // output voxels x = 8p .. 8p+7
        const int c = (p >> 1) | (y >> 4) << 1 | (z >> 4) << 2;    // child: dx + 2 dy + 4 dz
        const int ch = s_child[c];
        uint4 r = make_uint4(0, 0, 0, 0);
        if (ch >= 0) {
            const uint16_t* src = voxels + (size_t)ch * VXP_CHUNK_VOXELS;
            const int xi = (16 * p) & 31, yi = (2 * y) & 31, zi = (2 * z) & 31;
            uint32_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            // priority: y high first, then z low, then x low (SPEC §11)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int yy = yi + 1 - (t >> 1), zz = zi + (t & 1);
                const uint4* row = reinterpret_cast<const uint4*>(src + xi + 32 * yy + 1024 * zz);
                const uint4 a = __ldg(row), b = __ldg(row + 1);    // 16 input voxels
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {                      // input voxels 2e (low half), 2e+1 (high half) -> output voxel e
                    const uint32_t v = first_solid(w[e] & 0xFFFFu, w[e] >> 16);
                    o[e] = first_solid(o[e], v);
                }
            }
            r = make_uint4(o[0] | o[1] << 16, o[2] | o[3] << 16, o[4] | o[5] << 16, o[6] | o[7] << 16);
        }
        st_voxels16(dst + j, r);
    }
}

cudaError_t launch_downsample_chunks(const uint16_t* d_voxels, const int32_t* d_child8, uint32_t n_out, uint16_t* d_out,
                                     cudaStream_t stream) {
    if (n_out == 0) return cudaSuccess;
    downsample_chunks_kernel<<<4 * n_out, kThreads, 0, stream>>>(d_voxels, d_child8, n_out, d_out);
    return cudaGetLastError();
}
}  // namespace vxp
