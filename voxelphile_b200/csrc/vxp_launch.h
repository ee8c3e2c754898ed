// vxp_launch.h — host-side launchers of the sm_100a kernels (internal, C++ linkage).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vxp.h"

namespace vxp {
// Context-owned scratch of the mesh kernel (vxp_mesh_common.cuh, "halo"): per meshed chunk one tag word and
// two 32-word x-boundary bit planes stored as epoch-stamped 64-bit words.  Both areas must be zero-filled
// when allocated; epoch is a per-context launch counter in [1, 2^30) (the areas are re-zeroed when it
// wraps).  tags == nullptr disables the exchange.
struct BoundaryScratch {
    uint32_t* tags;
    unsigned long long* planes;
    uint32_t epoch;
};
size_t boundary_tag_bytes(uint32_t n);
size_t boundary_plane_bytes(uint32_t n);

// Context-owned scratch of the fused terrain+mesh path: the chunks of a batch that share a column (cx,cz)
// share one height footprint.  Sized for a batch of n chunks: keys / id_of_slot have column_table_slots(n)
// entries (a power of two, mask = slots - 1), slot_of_chunk has n, work (the list of the chunks that cross the
// surface, with what the mesh kernel needs to start on one) up to n, coord_of_id n, heights n records of column_record_bytes(), count
// kColumnCounterWords words (columns, work items, ticket, CTAs done, 64-bit quad counter; 8-byte aligned).
// keys and count must be zero-filled when allocated.  keys are stamped with `epoch` (in [1, 1023], advanced by
// the owner per call; the owner re-zeroes keys when it wraps), so no call starts with a memset; the fused mesh
// kernel leaves count zeroed.
constexpr int kColumnCounterWords = 8;
struct ColumnScratch {
    unsigned long long* keys;
    uint32_t* id_of_slot;
    uint32_t* slot_of_chunk;
    uint4* work;                 // work list of the fused mesh kernel: {chunk, column id, chunk y, 0}
    int2* coord_of_id;
    short* heights;
    uint32_t* count;
    uint32_t mask;
    uint32_t epoch;
};
size_t column_table_slots(uint32_t n);
size_t column_record_bytes();

// per-device one-time setup (opt-in shared memory size); call with the device current
cudaError_t configure_device();
cudaError_t launch_terrain_fill(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                cudaStream_t stream);
// one launch; the chunks of a column share its heights through the column scratch (n <= 2^22 - 2)
cudaError_t launch_terrain_fill_shared(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                       const ColumnScratch& cs, cudaStream_t stream);
// Large batches: the chunks of a column (cx,cz) share one height footprint (three launches; columns: scratch for n chunks).
cudaError_t launch_terrain_fill_columns(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                        const ColumnScratch& columns, cudaStream_t stream);
cudaError_t launch_pattern_fill(const vxp_coord* d_coords, uint32_t n, int pattern, uint16_t* d_voxels,
                                cudaStream_t stream);
// One whole-batch launch.  `counter` is a context-owned, zero-initialised 8-byte launch counter (vxp_mesh_common.cuh,
// "launch chaining"): the kernel reserves quad ranges from it, hands the total to *out.total_quads and leaves it
// zeroed, so no memset precedes the kernel.  overlap != 0 launches with programmatic stream serialisation: the
// kernel may start while the previous mesh launch on the stream is still draining.  The caller must then alternate
// between two counters / scratch areas / output arrays.
cudaError_t launch_mesh(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, int greedy,
                        const vxp_mesh_dev& out, const BoundaryScratch& scratch, unsigned long long* counter, int overlap,
                        cudaStream_t stream);
// chunks [lo, hi) of a batch of n; does NOT zero *out.total_quads (ranges of one batch accumulate into it)
cudaError_t launch_mesh_range(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, uint32_t lo, uint32_t hi,
                              int greedy, const vxp_mesh_dev& out, const BoundaryScratch& scratch, cudaStream_t stream);
// Incremental re-mesh.  launch_apply_edits writes the voxels and builds the list of chunks whose
// mesh they invalidate (d_flags: n words, context-owned, zeroed on allocation; epoch distinguishes calls; d_count
// is zeroed here).  launch_mesh_list meshes chunk d_list[b] into meta entry b for b < *d_count <= max_count
// (max_count CTAs are launched; the count stays on the device; d_count == nullptr: the count is max_count).
cudaError_t launch_apply_edits(uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, const vxp_edit* d_edits, uint32_t m,
                               uint32_t* d_flags, uint32_t epoch, uint32_t* d_list, uint32_t* d_count, cudaStream_t stream);
cudaError_t launch_mesh_list(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, const uint32_t* d_list,
                             const uint32_t* d_count, uint32_t max_count, int greedy, const vxp_mesh_dev& out,
                             unsigned long long* counter, cudaStream_t stream);
// Chunk wire format (SPEC §10; vxp_pack.cu).  d_total ends as the bytes the batch needs (records that do
// not fit get offset ~0); d_cursor: 8 bytes of context scratch, zero between launches (the kernel leaves it so).
// d_status: one sticky word, 1 = malformed record (the caller clears it when it has read a verdict).
cudaError_t launch_pack_chunks(const uint16_t* d_voxels, uint32_t n, unsigned char* d_packed, unsigned long long capacity,
                               unsigned long long* d_offsets, unsigned long long* d_total, unsigned long long* d_cursor,
                               cudaStream_t stream);
cudaError_t launch_unpack_chunks(const unsigned char* d_packed, unsigned long long packed_bytes, const unsigned long long* d_offsets,
                                 uint32_t n, uint16_t* d_voxels, uint32_t* d_status, cudaStream_t stream);
// LOD (SPEC §11): out[i] = half-resolution chunk of the 2x2x2 stored chunks child8[i][dx + 2 dy + 4 dz] (-1 = air).
cudaError_t launch_downsample_chunks(const uint16_t* d_voxels, const int32_t* d_child8, uint32_t n_out, uint16_t* d_out,
                                     cudaStream_t stream);
cudaError_t launch_terrain_mesh(const vxp_coord* d_coords, uint32_t n, uint64_t seed, int greedy,
                                const vxp_mesh_dev& out, const ColumnScratch& columns, cudaStream_t stream);
}  // namespace vxp
