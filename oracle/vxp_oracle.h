/*
 * vxp_oracle.h — CPU oracle for SPEC.md v1.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: the reference snapshot (/root/reference, crate `xenotech`,
 * src/lib.rs:1-111, src/net/udp/mod.rs:1-91) contains no chunk, mesher or
 * noise code and no tests or golden vectors.  This file is
 * therefore a builder-authored normative implementation of SPEC.md, not a
 * restatement of reference code.  It is pinned only by the hand-derived
 * known-answer tests in tests/ and by an independent pure-Python restatement
 * of the declarative spec (tests/spec_py.py) that generated tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libvxp.so) never
 * links or calls it.
 */
#ifndef VXP_ORACLE_H
#define VXP_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXO_N 32
#define VXO_VOXELS 32768
#define VXO_MAX_QUADS 98304 /* 16384 solid voxels x 6 faces (checkerboard) */

enum { VXO_CULLED = 0, VXO_GREEDY = 1 };
enum { VXO_PATTERN_CHECKER = 0, VXO_PATTERN_SOLID4 = 1 };

/* SPEC §6 */
uint32_t vxo_mix32(uint32_t h);
int32_t vxo_height(uint64_t seed, int32_t X, int32_t Z);
uint16_t vxo_block(uint64_t seed, int32_t X, int32_t Y, int32_t Z);
void vxo_terrain_fill(uint64_t seed, int32_t cx, int32_t cy, int32_t cz, uint16_t *out);
void vxo_terrain_fill_batch(uint64_t seed, const int32_t *coords, uint32_t n, int threads,
                            uint16_t *out);
/* SPEC §7 */
void vxo_pattern_fill(int pattern, int32_t cx, int32_t cy, int32_t cz, uint16_t *out);

/* SPEC §4: mesh one chunk into (unsorted) canonical keys; nb[f] may be NULL (= air).
 * keys must hold VXO_MAX_QUADS entries.  Returns the quad count. */
uint32_t vxo_mesh_chunk_keys(const uint16_t *vox, const uint16_t *const nb[6], int mode,
                             uint64_t *keys);
void vxo_sort_keys(uint64_t *keys, uint32_t n);
uint64_t vxo_mix64(uint64_t z);
uint64_t vxo_hash_keys(const uint64_t *keys, uint32_t n);
/* SPEC §5: vertices and indices of the quad `key` with chunk-local ordinal q. */
void vxo_expand_quad(uint64_t key, uint32_t q, uint64_t verts[4], uint32_t idx[6]);

/* Batch mesher with the same output contract as vxp_mesh_batch (SPEC §5).
 * Returns 0, or 1 if some chunk did not fit in cap_quads. */
int vxo_mesh_batch(const uint16_t *vox, const int32_t *nbr, uint32_t n, int mode, int threads,
                   uint64_t *verts, uint32_t *indices, uint64_t cap_quads, uint32_t *meta,
                   uint64_t *total_quads);
/* Per-chunk canonical hashes + counts only (no buffers). */
void vxo_mesh_batch_hash(const uint16_t *vox, const int32_t *nbr, uint32_t n, int mode,
                         int threads, uint64_t *hashes, uint32_t *counts);
int vxo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
