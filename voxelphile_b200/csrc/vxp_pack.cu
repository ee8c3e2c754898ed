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

__global__ void __launch_bounds__(kThreads) pack_chunks_kernel(const uint16_t* __restrict__ voxels, uint32_t n,
                                                                unsigned char* __restrict__ packed, unsigned long long capacity,
                                                                unsigned long long* __restrict__ offsets,
                                                                unsigned long long* __restrict__ cursor) {
    __shared__ PackShared sh;
    const int tid = threadIdx.x;
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n) return;
    const uint4* src = reinterpret_cast<const uint4*>(voxels + (size_t)chunk * VXP_CHUNK_VOXELS);
    if (tid < kSetSlots) sh.set[tid] = kEmptySlot;
    if (tid == 0) sh.distinct = 0;
    __syncthreads();
    // pass 1: the set of ids (a thread skips repeats of the value it inserted last: chunks are mostly runs)
    uint32_t last = kEmptySlot;
#pragma unroll 4
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
    __syncthreads();
    const uint32_t distinct = sh.distinct;
    const bool raw = distinct > 16u;
    if (tid < 16) sh.palette[tid] = 0;
    __syncthreads();
    if (!raw && tid < kSetSlots) {                                 // rank sort of at most 16 values
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
        const unsigned long long off = atomicAdd(cursor, size);
        sh.bits = bits;
        sh.len = raw ? 0u : distinct;
        sh.offset = off + size <= capacity ? off : ~0ull;
        offsets[chunk] = sh.offset;
    }
    __syncthreads();
    if (sh.offset == ~0ull) return;                                // does not fit: *cursor tells the caller how much is needed
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
    // pass 2 (L2 hits): 8 voxels -> 8 indices -> `bits` bytes
    for (int k = 0; k < 16; ++k) {
        const int j = tid + k * kThreads;
        const uint4 q = __ldg(src + j);
        if (bits == 16u) { reinterpret_cast<uint4*>(data)[j] = q; continue; }
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t out = 0;
        // chunks are mostly runs of one id: look an id up only when it differs from the previous voxel's
        uint32_t lastv = w[0] & 0xFFFFu, lasti = palette_index(sh.palette, len, lastv);
        if (q.x == q.y && q.x == q.z && q.x == q.w && lastv == (q.x >> 16)) {   // eight equal voxels: one look-up, one multiply
            out = lasti * (bits == 1u ? 0xFFu : bits == 2u ? 0x5555u : 0x11111111u);
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t v = (t & 1) ? w[t >> 1] >> 16 : w[t >> 1] & 0xFFFFu;
                if (v != lastv) { lastv = v; lasti = palette_index(sh.palette, len, v); }
                out |= lasti << (bits * t);
            }
        }
        if (bits == 1u) data[j] = (unsigned char)out;
        else if (bits == 2u) reinterpret_cast<unsigned short*>(data)[j] = (unsigned short)out;
        else reinterpret_cast<uint32_t*>(data)[j] = out;
    }
}

// status[0] is set to 1 if any record is malformed (offset outside the buffer - vxp_pack_chunks marks records that did
// not fit with ~0 -, misaligned, bad magic / bits / palette length, data running past the end); such chunks come out as air.
__global__ void __launch_bounds__(kThreads) unpack_chunks_kernel(const unsigned char* __restrict__ packed, unsigned long long packed_bytes,
                                                                  const unsigned long long* __restrict__ offsets, uint32_t n,
                                                                  uint16_t* __restrict__ voxels, uint32_t* __restrict__ status) {
    __shared__ uint32_t s_pal[16];
    __shared__ uint32_t s_bits;
    const int tid = threadIdx.x;
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n) return;
    const unsigned long long off = offsets[chunk];
    const bool in_buffer = (off & 15ull) == 0ull && off <= packed_bytes && packed_bytes - off >= (unsigned long long)kHeaderBytes;   // block-uniform
    const unsigned char* rec = packed + (in_buffer ? off : 0ull);
    if (tid == 0) {
        bool ok = in_buffer;
        uint32_t bits = 0;
        if (ok) {
            const uint32_t* h = reinterpret_cast<const uint32_t*>(rec);
            bits = h[1] & 0xFFFFu;
            const uint32_t len = h[1] >> 16;
            ok = h[0] == kMagic && (bits == 0u || bits == 1u || bits == 2u || bits == 4u || bits == 16u) &&
                 h[2] == 4096u * bits && (bits == 16u ? len == 0u : (len >= 1u && len <= 16u && len <= (1u << bits))) &&
                 packed_bytes - off - (unsigned long long)kHeaderBytes >= (unsigned long long)h[2];
        }
        if (!ok) atomicExch(status, 1u);
        s_bits = ok ? bits : 0xFFu;
    }
    if (tid < 8 && in_buffer) {
        const uint32_t p = reinterpret_cast<const uint32_t*>(rec + 16)[tid];
        s_pal[2 * tid] = p & 0xFFFFu;
        s_pal[2 * tid + 1] = p >> 16;
    }
    __syncthreads();
    const uint32_t bits = s_bits;
    uint4* dst = reinterpret_cast<uint4*>(voxels + (size_t)chunk * VXP_CHUNK_VOXELS);
    const unsigned char* data = rec + kHeaderBytes;
    if (bits == 0u || bits == 0xFFu) {
        const uint32_t b = bits == 0u ? s_pal[0] : 0u, bb = b | b << 16;
        for (int k = 0; k < 16; ++k) dst[tid + k * kThreads] = make_uint4(bb, bb, bb, bb);
        return;
    }
    if (bits == 16u) {
#pragma unroll 8
        for (int k = 0; k < 16; ++k) dst[tid + k * kThreads] = __ldg(reinterpret_cast<const uint4*>(data) + tid + k * kThreads);
        return;
    }
    const uint32_t mask = (1u << bits) - 1u;
#pragma unroll 1
    for (int k0 = 0; k0 < 16; k0 += 8) {                  // eight loads in flight per thread, then eight stores
        uint32_t in[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = tid + (k0 + k) * kThreads;
            in[k] = bits == 1u ? data[j] : bits == 2u ? reinterpret_cast<const unsigned short*>(data)[j]
                                                      : reinterpret_cast<const uint32_t*>(data)[j];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                w[e] = s_pal[(in[k] >> (bits * (2 * e))) & mask] | s_pal[(in[k] >> (bits * (2 * e + 1))) & mask] << 16;
            dst[tid + (k0 + k) * kThreads] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

cudaError_t launch_pack_chunks(const uint16_t* d_voxels, uint32_t n, unsigned char* d_packed, unsigned long long capacity,
                               unsigned long long* d_offsets, unsigned long long* d_total, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess || n == 0) return e;
    pack_chunks_kernel<<<n, kThreads, 0, stream>>>(d_voxels, n, d_packed, capacity, d_offsets, d_total);
    return cudaGetLastError();
}

cudaError_t launch_unpack_chunks(const unsigned char* d_packed, unsigned long long packed_bytes, const unsigned long long* d_offsets,
                                 uint32_t n, uint16_t* d_voxels, uint32_t* d_status, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_status, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess || n == 0) return e;
    unpack_chunks_kernel<<<n, kThreads, 0, stream>>>(d_packed, packed_bytes, d_offsets, n, d_voxels, d_status);
    return cudaGetLastError();
}

}  // namespace vxp

// ====================================================================== LOD (SPEC §11)
// One output chunk = 2 x 2 x 2 stored chunks at half resolution.  One CTA per output chunk; a thread produces 8
// output voxels (one 16-byte store) from 4 rows x 16 input voxels.  HBM-bound: 8 bytes read per byte written.
namespace vxp {
namespace {
// first non-air of (a, b) per u16 lane pair -> one u16: a if a != 0 else b
__device__ __forceinline__ uint32_t first_solid(uint32_t a, uint32_t b) { return a ? a : b; }
}  // namespace

__global__ void __launch_bounds__(kThreads) downsample_chunks_kernel(const uint16_t* __restrict__ voxels,
                                                                      const int32_t* __restrict__ child8, uint32_t n_out,
                                                                      uint16_t* __restrict__ out) {
    const uint32_t oc = blockIdx.x;
    if (oc >= n_out) return;
    __shared__ int s_child[8];
    if (threadIdx.x < 8) s_child[threadIdx.x] = child8[8 * (size_t)oc + threadIdx.x];
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)oc * VXP_CHUNK_VOXELS);
    for (int j = threadIdx.x; j < 4096; j += kThreads) {          // j = (z*32 + y)*4 + p
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
        dst[j] = r;
    }
}

cudaError_t launch_downsample_chunks(const uint16_t* d_voxels, const int32_t* d_child8, uint32_t n_out, uint16_t* d_out,
                                     cudaStream_t stream) {
    if (n_out == 0) return cudaSuccess;
    downsample_chunks_kernel<<<n_out, kThreads, 0, stream>>>(d_voxels, d_child8, n_out, d_out);
    return cudaGetLastError();
}
}  // namespace vxp
