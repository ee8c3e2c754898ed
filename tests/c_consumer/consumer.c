/*
 * consumer.c — a plain C11 program that includes include/vxp.h and LINKS libvxp.so (no dlopen, no Python):
 * what a Rust `extern "C"` block in the reference crate would bind (the dependency would be
 * declared at /root/reference/Cargo.toml:6).  Test infrastructure (tests/test_abi.py builds and runs it).
 *
 * Runs BASELINE config 1 and two neighbours of it through the host-level calls: chunk coordinates ->
 * vxp_terrain_fill_host -> vxp_mesh_chunks_host (both modes) and vxp_terrain_mesh_host, and prints per chunk
 *     <call> <mode> <chunk> <quad_count> <sum of mix64(canonical key)>
 * for the test to compare with the oracle.  Exit code 2: no CUDA device (the library has no CPU fallback).
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include "vxp.h"

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* SPEC §5 inverse: canonical key of a quad from its first vertex (always corner c0) */
static uint64_t key_of(uint64_t v0) {
    const uint32_t w0 = (uint32_t)v0, w1 = (uint32_t)(v0 >> 32);
    const uint32_t p[3] = {w0 & 63u, (w0 >> 6) & 63u, (w0 >> 12) & 63u};
    const uint32_t f = (w0 >> 18) & 7u, axis = f >> 1;
    const uint64_t s = p[axis] - (f & 1u), u = p[axis == 0 ? 1 : 0], v = p[axis == 2 ? 1 : 2];
    return u | v << 6 | s << 12 | (uint64_t)f << 18 | (uint64_t)((w1 >> 16) & 31u) << 21 |
           (uint64_t)((w1 >> 21) & 31u) << 26 | (uint64_t)(w1 & 0xFFFFu) << 31;
}

static void report(const char *call, int mode, const vxp_mesh_host *m, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t h = 0;
        const vxp_chunk_meta e = m->meta[i];
        for (uint32_t q = 0; q < e.quad_count; ++q) h += mix64(key_of(m->verts[4 * ((uint64_t)e.first_quad + q)]));
        printf("%s %d %" PRIu32 " %" PRIu32 " %" PRIu64 "\n", call, mode, i, e.quad_count, h);
    }
}

#define CHECK(expr)                                                                              \
    do {                                                                                         \
        vxp_status s_ = (expr);                                                                  \
        if (s_ != VXP_OK) {                                                                      \
            fprintf(stderr, "%s: %s (%s)\n", #expr, vxp_status_str(s_), ctx ? vxp_last_error(ctx) : ""); \
            return s_ == VXP_ERR_NO_DEVICE ? 2 : 1;                                              \
        }                                                                                        \
    } while (0)

int main(void) {
    enum { N = 3 };
    const vxp_coord coords[N] = {{0, 0, 0}, {0, 8, 0}, {1, 8, 0}};
    const uint64_t seed = 42;
    vxp_ctx *ctx = NULL;
    if (vxp_spec_version() != VXP_SPEC_VERSION) return 1;
    CHECK(vxp_ctx_create(0, &ctx));
    int32_t nbr[N][6];
    CHECK(vxp_nbr6_from_coords(coords, N, &nbr[0][0]));
    uint16_t *vox = malloc((size_t)N * VXP_CHUNK_VOXELS * sizeof *vox);
    if (!vox) return 1;
    CHECK(vxp_terrain_fill_host(ctx, coords, N, seed, vox));
    for (int mode = VXP_MODE_CULLED; mode <= VXP_MODE_GREEDY; ++mode) {
        vxp_mesh_host m;
        CHECK(vxp_mesh_chunks_host(ctx, vox, &nbr[0][0], N, N, (vxp_mode)mode, &m));
        report("stored", mode, &m, N);
        CHECK(vxp_terrain_mesh_host(ctx, coords, N, seed, (vxp_mode)mode, &m));
        report("fused", mode, &m, N);
    }
    free(vox);
    vxp_ctx_destroy(ctx);
    return 0;
}
