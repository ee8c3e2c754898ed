/*
 * vxp.h — C ABI of libvxp.so: B200 (sm_100a) chunk terrain fill and meshing.
 *
 * Boundary note.  The reference snapshot (voxelphile/voxelphile, crate
 * `xenotech`) exposes NO interface for this path: its only public items are
 * `xenotech::main` (src/lib.rs:9), the login enums (src/lib.rs:31-45) and
 * `User` (src/lib.rs:47); `client()`/`server()` (src/lib.rs:109,111), where a
 * world/mesher would attach, are empty.  There is therefore nothing for these
 * entry points to "replace"; they define the boundary that a
 * Rust `extern "C"` block would bind — see INTEGRATION.md and rust/vxp-sys.
 * The one convention carried over is the reference's payload-free error kinds
 * returned by value (src/net/udp/mod.rs:16-28): every call returns a
 * vxp_status and nothing throws across the ABI.
 *
 * Semantics of every buffer and bit are fixed by SPEC.md (v1).
 * There is no CPU fallback: without a CUDA device every call fails.
 *
 * Threading: a vxp_ctx is used by one thread at a time; distinct contexts (one
 * per GPU) may be driven concurrently.  Device-level calls are asynchronous on
 * the context's stream; host-level calls synchronise before returning.
 */
#ifndef VXP_H
#define VXP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXP_SPEC_VERSION 1
#define VXP_CHUNK_DIM 32
#define VXP_CHUNK_VOXELS 32768
#define VXP_MAX_QUADS_PER_CHUNK 98304

typedef struct vxp_ctx vxp_ctx;

typedef enum vxp_status {
    VXP_OK = 0,
    VXP_ERR_BAD_ARG = 1,
    VXP_ERR_NO_DEVICE = 2,
    VXP_ERR_OOM = 3,
    VXP_ERR_CAPACITY = 4, /* output arena too small; total_quads holds the need */
    VXP_ERR_CUDA = 5
} vxp_status;

typedef enum vxp_mode { VXP_MODE_CULLED = 0, VXP_MODE_GREEDY = 1 } vxp_mode;
typedef enum vxp_pattern { VXP_PATTERN_CHECKER = 0, VXP_PATTERN_SOLID4 = 1 } vxp_pattern;

typedef struct vxp_coord { int32_t x, y, z; } vxp_coord;           /* chunk coordinate */
typedef struct vxp_chunk_meta { uint32_t first_quad, quad_count; } vxp_chunk_meta;
/* One voxel edit: voxel (x,y,z) of stored chunk `chunk` becomes `id` (0 = air).  12 bytes. */
typedef struct vxp_edit { uint32_t chunk; uint16_t x, y, z, id; } vxp_edit;

/* Batch mesh in DEVICE memory (SPEC §5).  All four arrays are caller-owned.
 * verts must be 32-byte aligned (one quad = one 32-byte store), indices 8-byte aligned.
 * indices may be NULL: the index buffer is then not written at all (SPEC §5 makes it a pure function of the
 * quad ordinal, 4q + {0,1,2,2,3,0}; a renderer can bind one static index buffer for every mesh) - 24 of the
 * 56 bytes per quad less to write. */
typedef struct vxp_mesh_dev {
    uint64_t *verts;          /* 4 * capacity_quads vertices, 8 B each        */
    uint32_t *indices;        /* 6 * capacity_quads indices, or NULL           */
    vxp_chunk_meta *meta;     /* n entries                                     */
    uint64_t *total_quads;    /* one counter; written by the call (need not be zeroed) */
    uint64_t capacity_quads;
} vxp_mesh_dev;

/* Batch mesh in HOST memory, owned by the context (pinned); valid until the
 * next host-level call on the same context or vxp_ctx_destroy. */
typedef struct vxp_mesh_host {
    const uint64_t *verts;
    const uint32_t *indices;  /* NULL after vxp_ctx_set_host_indices(ctx, 0) */
    const vxp_chunk_meta *meta;
    uint64_t total_quads;
} vxp_mesh_host;

/* ---- context ----------------------------------------------------------- */
vxp_status vxp_ctx_create(int device, vxp_ctx **out);
void vxp_ctx_destroy(vxp_ctx *ctx);
/* Run all of this context's work on the caller's CUDA stream (a cudaStream_t).
 * NULL means what it means to CUDA: the legacy default stream.
 * vxp_ctx_reset_stream returns to the context's own non-blocking stream. */
vxp_status vxp_ctx_set_stream(vxp_ctx *ctx, void *cuda_stream);
vxp_status vxp_ctx_reset_stream(vxp_ctx *ctx);
/* Launch overlap (off by default).  When enabled, a vxp_mesh_batch call may start executing while the previous
 * vxp_mesh_batch call on the same context and stream is still finishing (programmatic dependent launch): the
 * next batch's chunks fill the SMs that the tail of the previous batch leaves idle.  At most two calls are ever
 * in flight; call k+2 never starts before call k has completed.  The caller promises, while it is enabled:
 *   - consecutive vxp_mesh_batch calls use disjoint output arrays (verts, indices, meta, total_quads), e.g.
 *     two vxp_mesh_dev used alternately; the arrays of call k may be re-used by call k+2;
 *   - the voxels and neighbour table of a call were complete before the PREVIOUS vxp_mesh_batch call was
 *     enqueued (a call that follows any other vxp call, vxp_ctx_synchronize or vxp_ctx_set_overlap never
 *     starts early, so freshly produced inputs are safe behind one of those).
 * Results are identical either way. */
vxp_status vxp_ctx_set_overlap(vxp_ctx *ctx, int enabled);
/* Host-level calls: whether the index buffer is produced and downloaded (default 1).  With 0, vxp_mesh_host.indices
 * is NULL, the kernels skip the index writes and 24 of the 56 bytes per quad do not cross PCIe. */
vxp_status vxp_ctx_set_host_indices(vxp_ctx *ctx, int enabled);
vxp_status vxp_ctx_synchronize(vxp_ctx *ctx);
const char *vxp_status_str(vxp_status s);
/* Detail of the last VXP_ERR_CUDA on this context ("" if none). */
const char *vxp_last_error(const vxp_ctx *ctx);
/* Kernels launched by this context since creation (bench bookkeeping). */
uint64_t vxp_launch_count(const vxp_ctx *ctx);
int vxp_spec_version(void);

/* ---- host helpers (no device work) ------------------------------------- */
/* nbr[i][f] = index of the chunk at coords[i] + dir_f, or -1 (SPEC §1). */
vxp_status vxp_nbr6_from_coords(const vxp_coord *coords, uint32_t n, int32_t *nbr6);
/* Pinned host memory for the host-level calls' inputs. */
vxp_status vxp_host_alloc(size_t bytes, void **out);
void vxp_host_free(void *p);

/* ---- device level: every pointer is device memory ---------------------- */
/* SPEC §6: d_voxels[i] = terrain of chunk d_coords[i]. */
vxp_status vxp_terrain_fill_batch(vxp_ctx *ctx, const vxp_coord *d_coords, uint32_t n,
                                  uint64_t seed, uint16_t *d_voxels);
/* SPEC §7 test patterns. */
vxp_status vxp_pattern_fill_batch(vxp_ctx *ctx, const vxp_coord *d_coords, uint32_t n,
                                  vxp_pattern pattern, uint16_t *d_voxels);
/* SPEC §2-§5: mesh the first n chunks stored at d_voxels.  d_nbr6 (n x 6) may be
 * NULL (all sides air); its entries may name any chunk stored at d_voxels,
 * including read-only "halo" chunks at index >= n that are not meshed
 * themselves (multi-GPU seam layers).  Overflow of out->capacity_quads is only
 * visible in *out->total_quads and meta[i].first_quad == 0xFFFFFFFF (the call
 * itself is asynchronous). */
vxp_status vxp_mesh_batch(vxp_ctx *ctx, const uint16_t *d_voxels, const int32_t *d_nbr6,
                          uint32_t n, vxp_mode mode, const vxp_mesh_dev *out);
/* Terrain + mesh without materialising voxels: chunk d_coords[i] is meshed
 * against the infinite terrain of `seed` (all six neighbours are terrain). */
vxp_status vxp_terrain_mesh_fused(vxp_ctx *ctx, const vxp_coord *d_coords, uint32_t n,
                                  uint64_t seed, vxp_mode mode, const vxp_mesh_dev *out);

/* ---- incremental re-mesh: small batches, latency path -- */
/* Mesh the `count` chunks named by d_list (indices into d_voxels, each < n; d_nbr6 is the n x 6 table of the
 * whole batch).  out->meta has `count` entries: entry k belongs to chunk d_list[k]. */
vxp_status vxp_mesh_list(vxp_ctx *ctx, const uint16_t *d_voxels, const int32_t *d_nbr6, uint32_t n,
                         const uint32_t *d_list, uint32_t count, vxp_mode mode, const vxp_mesh_dev *out);
/* Apply m voxel edits to the stored chunks and re-mesh exactly the chunks they invalidate: every edited chunk
 * and, for an edited voxel on a chunk face, the neighbour across that face (d_nbr6 must be symmetric for the
 * edited chunks: nbr6[nbr6[c][f]][f^1] == c).  No host round trip: the dirty list and its length stay on the
 * device.  d_dirty gets the invalidated chunk indices (arbitrary order, no duplicates; room for min(n, 7m)
 * entries), *d_dirty_count their number; out->meta has min(n, 7m) entries, entry k belongs to chunk d_dirty[k].
 * Edits of one call must address distinct voxels (of two edits of the same voxel either may win); edits whose
 * chunk or coordinates are out of range are ignored. */
vxp_status vxp_remesh_edits(vxp_ctx *ctx, uint16_t *d_voxels, const int32_t *d_nbr6, uint32_t n,
                            const vxp_edit *d_edits, uint32_t m, vxp_mode mode,
                            uint32_t *d_dirty, uint32_t *d_dirty_count, const vxp_mesh_dev *out);

/* ---- chunk wire format (SPEC §10): palette + bit-packed indices ---- */
/* A packed batch is a byte buffer of 16-byte-aligned records (48-byte header + palette, then 0 / 4 / 8 / 16 / 64 KiB
 * of indices for 1 / 2 / 3-4 / 5-16 / more distinct ids) plus offsets[i] = byte offset of chunk i's record. */
#define VXP_PACKED_HEADER_BYTES 48
#define VXP_PACKED_MAX_RECORD_BYTES (48 + 65536)
/* Compress n stored chunks on the device.  Records are placed in arbitrary order (d_offsets says where); the bytes
 * of a record depend on its chunk only.  *d_total_bytes ends as the bytes the batch needs: if that exceeds
 * capacity_bytes, the records that did not fit have offset UINT64_MAX.  d_packed must be 16-byte aligned. */
vxp_status vxp_pack_chunks(vxp_ctx *ctx, const uint16_t *d_voxels, uint32_t n, uint8_t *d_packed,
                           uint64_t capacity_bytes, uint64_t *d_offsets, uint64_t *d_total_bytes);
/* Expand n records of the packed buffer (packed_bytes long) into d_voxels (n * 32768 u16).  Malformed records - bad
 * header, or an offset or data that does not lie inside the buffer, such as the UINT64_MAX vxp_pack_chunks gives a
 * record that did not fit - come out as air and make the NEXT vxp_ctx_synchronize return VXP_ERR_BAD_ARG (calls queue
 * without synchronising; the verdict covers every vxp_unpack_chunks call since the previous synchronize). */
vxp_status vxp_unpack_chunks(vxp_ctx *ctx, const uint8_t *d_packed, uint64_t packed_bytes, const uint64_t *d_offsets,
                             uint32_t n, uint16_t *d_voxels);

/* ---- level of detail (SPEC §11) ------------------------ */
/* d_out[i] = the 2 x 2 x 2 stored chunks d_child8[i][dx + 2*dy + 4*dz] (indices into d_voxels, -1 = air) at half
 * resolution: output voxel (x,y,z) is the first non-air voxel of its 2 x 2 x 2 block in the order y high, z low,
 * x low, air if there is none.  The result is an ordinary chunk: mesh it with vxp_mesh_batch, apply the call
 * again for the next level.  d_out must not overlap the chunks it reads. */
vxp_status vxp_downsample_chunks(vxp_ctx *ctx, const uint16_t *d_voxels, const int32_t *d_child8, uint32_t n_out,
                                 uint16_t *d_out);

/* ---- host level: host pointers in, context-owned pinned results out ---- */
vxp_status vxp_terrain_fill_host(vxp_ctx *ctx, const vxp_coord *coords, uint32_t n,
                                 uint64_t seed, uint16_t *voxels);
/* voxels holds n_stored >= n chunks; the first n are meshed, nbr6 (n x 6) indexes
 * all n_stored of them. */
vxp_status vxp_mesh_chunks_host(vxp_ctx *ctx, const uint16_t *voxels, const int32_t *nbr6,
                                uint32_t n, uint32_t n_stored, vxp_mode mode, vxp_mesh_host *out);
vxp_status vxp_terrain_mesh_host(vxp_ctx *ctx, const vxp_coord *coords, uint32_t n,
                                 uint64_t seed, vxp_mode mode, vxp_mesh_host *out);
/* vxp_mesh_chunks_host for a packed batch: only the packed bytes cross PCIe (9 x fewer for a terrain world), the
 * chunks are expanded on the device.  n_stored records (the first n are meshed, nbr6 is n x 6 and indexes all of
 * them); every record is validated here (VXP_ERR_BAD_ARG).  The expanded world stays resident like after
 * vxp_mesh_chunks_host. */
vxp_status vxp_mesh_packed_host(vxp_ctx *ctx, const uint8_t *packed, uint64_t packed_bytes, const uint64_t *offsets,
                                const int32_t *nbr6, uint32_t n, uint32_t n_stored, vxp_mode mode, vxp_mesh_host *out);
/* Edit the world that the last vxp_mesh_chunks_host / vxp_mesh_packed_host call left resident on the device (its voxels and neighbour
 * table) and re-mesh what the edits invalidate: 12 bytes per edit go up, only the new meshes come down.
 * *dirty (context-owned, pinned, *count entries) names the re-meshed chunks; out->meta[k] belongs to (*dirty)[k].
 * VXP_ERR_BAD_ARG if no world is resident or an edit is out of range. */
vxp_status vxp_remesh_edits_host(vxp_ctx *ctx, const vxp_edit *edits, uint32_t m, vxp_mode mode,
                                 vxp_mesh_host *out, const uint32_t **dirty, uint32_t *count);

#ifdef __cplusplus
}
#endif
#endif /* VXP_H */
