/*
 * vxp_oracle.c — CPU oracle for SPEC.md v1.  TEST INFRASTRUCTURE ONLY; see
 * vxp_oracle.h for the "parity unpinned" statement.  Scalar, written for
 * clarity; shares no code with voxelphile_b200/csrc.
 */
#include "vxp_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------- tiny pthread parallel-for */

typedef void (*vxo_body_fn)(void *ctx, int64_t i, uint64_t *scratch_keys);
typedef struct {
    vxo_body_fn body;
    void *ctx;
    int64_t n, grain;
    atomic_llong next;
} vxo_job;

static void *vxo_worker(void *arg) {
    vxo_job *job = (vxo_job *)arg;
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * VXO_MAX_QUADS);
    for (;;) {
        int64_t lo = atomic_fetch_add(&job->next, job->grain);
        if (lo >= job->n) break;
        int64_t hi = lo + job->grain < job->n ? lo + job->grain : job->n;
        for (int64_t i = lo; i < hi; ++i) job->body(job->ctx, i, keys);
    }
    free(keys);
    return NULL;
}

static void vxo_parallel_for(int64_t n, int threads, int64_t grain, vxo_body_fn body, void *ctx) {
    vxo_job job = {body, ctx, n, grain, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, vxo_worker, &job);
    vxo_worker(&job);
    for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
}

/* ---------------------------------------------------------------- SPEC §6 */

uint32_t vxo_mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

static const int8_t GRAD[8][2] = {{1, 0}, {-1, 0}, {0, 1},  {0, -1},
                                  {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};

static int32_t grad_dot(uint32_t key, int32_t i, int32_t j, int32_t dx, int32_t dz) {
    uint32_t g = vxo_mix32(vxo_mix32(key + (uint32_t)i * 0x85EBCA6Bu) ^ ((uint32_t)j * 0xC2B2AE35u)) & 7u;
    return GRAD[g][0] * dx + GRAD[g][1] * dz;
}

static int32_t fade(int32_t k) { return (k * k * (768 - 2 * k)) >> 8; }

/* floor(a / 2^s) for signed a without relying on the shift of a negative value */
static int64_t floor_shr64(int64_t a, int s) {
    int64_t d = (int64_t)1 << s;
    int64_t q = a / d;
    if ((a % d) != 0 && a < 0) q -= 1;
    return q;
}
static int32_t floor_shr32(int32_t a, int s) { return (int32_t)floor_shr64(a, s); }

int32_t vxo_height(uint64_t seed, int32_t X, int32_t Z) {
    int64_t sum = 0;
    for (int o = 0; o < 4; ++o) {
        int S = 7 - o;
        int32_t A = 192 >> o;
        uint32_t key = vxo_mix32((uint32_t)(seed >> 32) ^
                                 vxo_mix32((uint32_t)seed + 0x9E3779B9u * (uint32_t)(o + 1)));
        int32_t Xo = X + 37 * o, Zo = Z + 59 * o;
        int32_t ix = floor_shr32(Xo, S), iz = floor_shr32(Zo, S);
        int32_t dx = (Xo - ix * (1 << S)) << (8 - S);
        int32_t dz = (Zo - iz * (1 << S)) << (8 - S);
        int32_t d00 = grad_dot(key, ix, iz, dx, dz);
        int32_t d10 = grad_dot(key, ix + 1, iz, dx - 256, dz);
        int32_t d01 = grad_dot(key, ix, iz + 1, dx, dz - 256);
        int32_t d11 = grad_dot(key, ix + 1, iz + 1, dx - 256, dz - 256);
        int32_t fx = fade(dx), fz = fade(dz);
        int32_t n0 = d00 * 65536 + fx * (d10 - d00);
        int32_t n1 = d01 * 65536 + fx * (d11 - d01);
        int64_t n = (int64_t)n0 * 65536 + (int64_t)fz * (int64_t)(n1 - n0);
        sum += n * A;
    }
    return 256 + (int32_t)floor_shr64(sum, 40);
}

static uint16_t block_from_height(int32_t h, int32_t Y) {
    if (Y > h) return 0;
    if (Y == h) return 1;
    if (Y >= h - 3) return 2;
    return (floor_shr32(Y, 3) & 1) ? 4 : 3;
}

uint16_t vxo_block(uint64_t seed, int32_t X, int32_t Y, int32_t Z) {
    return block_from_height(vxo_height(seed, X, Z), Y);
}

void vxo_terrain_fill(uint64_t seed, int32_t cx, int32_t cy, int32_t cz, uint16_t *out) {
    for (int z = 0; z < VXO_N; ++z)
        for (int x = 0; x < VXO_N; ++x) {
            int32_t h = vxo_height(seed, 32 * cx + x, 32 * cz + z);
            for (int y = 0; y < VXO_N; ++y)
                out[x + 32 * y + 1024 * z] = block_from_height(h, 32 * cy + y);
        }
}

typedef struct {
    uint64_t seed;
    const int32_t *coords;
    uint16_t *out;
} terrain_ctx;

static void terrain_body(void *p, int64_t i, uint64_t *scratch) {
    terrain_ctx *c = (terrain_ctx *)p;
    (void)scratch;
    vxo_terrain_fill(c->seed, c->coords[3 * i], c->coords[3 * i + 1], c->coords[3 * i + 2],
                     c->out + (size_t)i * VXO_VOXELS);
}

void vxo_terrain_fill_batch(uint64_t seed, const int32_t *coords, uint32_t n, int threads,
                            uint16_t *out) {
    terrain_ctx c = {seed, coords, out};
    vxo_parallel_for(n, threads, 8, terrain_body, &c);
}

/* ---------------------------------------------------------------- SPEC §7 */

void vxo_pattern_fill(int pattern, int32_t cx, int32_t cy, int32_t cz, uint16_t *out) {
    for (int z = 0; z < VXO_N; ++z)
        for (int y = 0; y < VXO_N; ++y)
            for (int x = 0; x < VXO_N; ++x) {
                int32_t X = 32 * cx + x, Y = 32 * cy + y, Z = 32 * cz + z;
                uint16_t id;
                if (pattern == VXO_PATTERN_CHECKER)
                    id = ((X + Y + Z) & 1) ? 1 : 0;
                else
                    id = (uint16_t)(1 + (vxo_mix32(vxo_mix32(vxo_mix32((uint32_t)X) ^ (uint32_t)Y) ^
                                                   (uint32_t)Z) & 3u));
                out[x + 32 * y + 1024 * z] = id;
            }
}

/* ------------------------------------------------------------ SPEC §2-§4 */

static const int DIR[6][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}};

/* voxel at p, where at most one coordinate is one step outside the chunk */
static uint16_t voxel_at(const uint16_t *vox, const uint16_t *const nb[6], int x, int y, int z) {
    const uint16_t *src = vox;
    if (x < 0) { src = nb[0]; x += 32; }
    else if (x > 31) { src = nb[1]; x -= 32; }
    else if (y < 0) { src = nb[2]; y += 32; }
    else if (y > 31) { src = nb[3]; y -= 32; }
    else if (z < 0) { src = nb[4]; z += 32; }
    else if (z > 31) { src = nb[5]; z -= 32; }
    return src ? src[x + 32 * y + 1024 * z] : 0;
}

/* (a, u, v) -> (x, y, z) per SPEC §3 */
static void plane_to_xyz(int axis, int s, int u, int v, int *x, int *y, int *z) {
    if (axis == 0) { *x = s; *y = u; *z = v; }
    else if (axis == 1) { *x = u; *y = s; *z = v; }
    else { *x = u; *y = v; *z = s; }
}

static uint64_t make_key(int f, int s, uint16_t b, int u, int v, int w, int h) {
    return (uint64_t)u | (uint64_t)v << 6 | (uint64_t)s << 12 | (uint64_t)f << 18 |
           (uint64_t)(w - 1) << 21 | (uint64_t)(h - 1) << 26 | (uint64_t)b << 31;
}

uint32_t vxo_mesh_chunk_keys(const uint16_t *vox, const uint16_t *const nb[6], int mode,
                             uint64_t *keys) {
    uint32_t nq = 0;
    uint16_t cell[32][32]; /* [v][u]: id of the visible face, 0 if none */
    for (int f = 0; f < 6; ++f) {
        int axis = f >> 1;
        for (int s = 0; s < 32; ++s) {
            int any = 0;
            for (int v = 0; v < 32; ++v)
                for (int u = 0; u < 32; ++u) {
                    int x, y, z;
                    plane_to_xyz(axis, s, u, v, &x, &y, &z);
                    uint16_t id = vox[x + 32 * y + 1024 * z];
                    uint16_t out = 0;
                    if (id != 0 &&
                        voxel_at(vox, nb, x + DIR[f][0], y + DIR[f][1], z + DIR[f][2]) == 0)
                        out = id;
                    cell[v][u] = out;
                    any |= out;
                }
            if (!any) continue;
            for (int v = 0; v < 32; ++v)
                for (int u = 0; u < 32; ++u) {
                    uint16_t b = cell[v][u];
                    if (!b) continue;
                    int w = 1, h = 1;
                    if (mode == VXO_GREEDY) {
                        while (u + w < 32 && cell[v][u + w] == b) ++w;
                        for (;;) {
                            if (v + h >= 32) break;
                            int ok = 1;
                            for (int k = 0; k < w; ++k)
                                if (cell[v + h][u + k] != b) { ok = 0; break; }
                            if (!ok) break;
                            ++h;
                        }
                    }
                    for (int j = 0; j < h; ++j)
                        for (int k = 0; k < w; ++k) cell[v + j][u + k] = 0;
                    keys[nq++] = make_key(f, s, b, u, v, w, h);
                }
        }
    }
    return nq;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
void vxo_sort_keys(uint64_t *keys, uint32_t n) { qsort(keys, n, sizeof(uint64_t), cmp_u64); }

uint64_t vxo_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

uint64_t vxo_hash_keys(const uint64_t *keys, uint32_t n) {
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) acc += vxo_mix64(keys[i]);
    return acc;
}

/* ---------------------------------------------------------------- SPEC §5 */

void vxo_expand_quad(uint64_t key, uint32_t q, uint64_t verts[4], uint32_t idx[6]) {
    int u = (int)(key & 63), v = (int)(key >> 6 & 63), s = (int)(key >> 12 & 63);
    int f = (int)(key >> 18 & 7), w = (int)(key >> 21 & 31) + 1, h = (int)(key >> 26 & 31) + 1;
    uint32_t b = (uint32_t)(key >> 31 & 0xFFFF);
    int axis = f >> 1, along = s + (f & 1);
    int cu[4] = {u, u + w, u + w, u}, cv[4] = {v, v, v + h, v + h};
    int flip = (f == 0 || f == 3 || f == 4);
    for (int k = 0; k < 4; ++k) {
        int c = flip ? (4 - k) & 3 : k; /* c0,c3,c2,c1 when flipped */
        int x, y, z;
        plane_to_xyz(axis, along, cu[c], cv[c], &x, &y, &z);
        uint32_t w0 = (uint32_t)x | (uint32_t)y << 6 | (uint32_t)z << 12 | (uint32_t)f << 18 |
                      (uint32_t)k << 21;
        uint32_t w1 = b | (uint32_t)(w - 1) << 16 | (uint32_t)(h - 1) << 21;
        verts[k] = (uint64_t)w0 | (uint64_t)w1 << 32;
    }
    static const uint32_t pat[6] = {0, 1, 2, 2, 3, 0};
    for (int k = 0; k < 6; ++k) idx[k] = 4 * q + pat[k];
}

/* ------------------------------------------------------------- batch API */

static void gather_nb(const uint16_t *vox, const int32_t *nbr, uint32_t i,
                      const uint16_t *nb[6]) {
    for (int f = 0; f < 6; ++f) {
        int32_t j = nbr ? nbr[6 * (size_t)i + f] : -1;
        nb[f] = j < 0 ? NULL : vox + (size_t)j * VXO_VOXELS;
    }
}

typedef struct {
    const uint16_t *vox;
    const int32_t *nbr;
    int mode;
    uint64_t *verts;
    uint32_t *indices;
    uint64_t cap_quads;
    uint32_t *meta;
    atomic_ullong cursor;
    atomic_int overflow;
    uint64_t *hashes;
    uint32_t *counts;
} mesh_ctx;

static void mesh_body(void *p, int64_t i, uint64_t *keys) {
    mesh_ctx *c = (mesh_ctx *)p;
    const uint16_t *nb[6];
    gather_nb(c->vox, c->nbr, (uint32_t)i, nb);
    uint32_t cnt = vxo_mesh_chunk_keys(c->vox + (size_t)i * VXO_VOXELS, nb, c->mode, keys);
    uint64_t first = atomic_fetch_add(&c->cursor, cnt);
    c->meta[2 * i + 1] = cnt;
    if (first + cnt > c->cap_quads) {
        c->meta[2 * i] = 0xFFFFFFFFu;
        atomic_store(&c->overflow, 1);
        return;
    }
    c->meta[2 * i] = (uint32_t)first;
    for (uint32_t q = 0; q < cnt; ++q)
        vxo_expand_quad(keys[q], q, c->verts + 4 * (first + q), c->indices + 6 * (first + q));
}

int vxo_mesh_batch(const uint16_t *vox, const int32_t *nbr, uint32_t n, int mode, int threads,
                   uint64_t *verts, uint32_t *indices, uint64_t cap_quads, uint32_t *meta,
                   uint64_t *total_quads) {
    mesh_ctx c = {vox, nbr, mode, verts, indices, cap_quads, meta, 0, 0, NULL, NULL};
    vxo_parallel_for(n, threads, 4, mesh_body, &c);
    *total_quads = atomic_load(&c.cursor);
    return atomic_load(&c.overflow);
}

static void hash_body(void *p, int64_t i, uint64_t *keys) {
    mesh_ctx *c = (mesh_ctx *)p;
    const uint16_t *nb[6];
    gather_nb(c->vox, c->nbr, (uint32_t)i, nb);
    uint32_t cnt = vxo_mesh_chunk_keys(c->vox + (size_t)i * VXO_VOXELS, nb, c->mode, keys);
    c->counts[i] = cnt;
    c->hashes[i] = vxo_hash_keys(keys, cnt);
}

void vxo_mesh_batch_hash(const uint16_t *vox, const int32_t *nbr, uint32_t n, int mode,
                         int threads, uint64_t *hashes, uint32_t *counts) {
    mesh_ctx c = {vox, nbr, mode, NULL, NULL, 0, NULL, 0, 0, hashes, counts};
    vxo_parallel_for(n, threads, 4, hash_body, &c);
}


/* ------------------------------------------- terrain world (fused path, config 5) */

typedef struct {
    uint64_t seed;
    int32_t x0, y0, w; /* origin and row width of the layer, halo ring included */
    int32_t cz;
    uint16_t *dst;
} layer_ctx;

static void layer_body(void *p, int64_t i, uint64_t *scratch) {
    layer_ctx *c = (layer_ctx *)p;
    (void)scratch;
    vxo_terrain_fill(c->seed, c->x0 + (int32_t)(i % c->w), c->y0 + (int32_t)(i / c->w), c->cz,
                     c->dst + (size_t)i * VXO_VOXELS);
}

typedef struct {
    const uint16_t *prev, *cur, *next; /* layers cz-1, cz, cz+1, each (nx+2) x (ny+2) chunks */
    int32_t nx, w;
    int mode;
    uint64_t *hashes;
    uint32_t *counts;
} world_ctx;

static void world_body(void *p, int64_t i, uint64_t *keys) {
    world_ctx *c = (world_ctx *)p;
    size_t j = (size_t)(i % c->nx + 1) + (size_t)c->w * (size_t)(i / c->nx + 1);
    const uint16_t *nb[6] = {c->cur + (j - 1) * VXO_VOXELS,          c->cur + (j + 1) * VXO_VOXELS,
                             c->cur + (j - (size_t)c->w) * VXO_VOXELS, c->cur + (j + (size_t)c->w) * VXO_VOXELS,
                             c->prev + j * VXO_VOXELS,                c->next + j * VXO_VOXELS};
    uint32_t cnt = vxo_mesh_chunk_keys(c->cur + j * VXO_VOXELS, nb, c->mode, keys);
    c->counts[i] = cnt;
    c->hashes[i] = vxo_hash_keys(keys, cnt);
}

int vxo_terrain_world_hash(uint64_t seed, const int32_t lo[3], const int32_t hi[3], int mode, int threads,
                           uint64_t *hashes, uint32_t *counts) {
    const int32_t nx = hi[0] - lo[0], ny = hi[1] - lo[1], nz = hi[2] - lo[2];
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    const int32_t w = nx + 2;
    const size_t layer_chunks = (size_t)w * (size_t)(ny + 2);
    uint16_t *buf[3];
    for (int k = 0; k < 3; ++k) {
        buf[k] = (uint16_t *)malloc(layer_chunks * VXO_VOXELS * sizeof(uint16_t));
        if (!buf[k]) {
            for (int j = 0; j < k; ++j) free(buf[j]);
            return -1;
        }
    }
    /* every chunk is generated with the plain per-chunk fill, neighbours included: no shortcut of any kind */
    layer_ctx lc = {seed, lo[0] - 1, lo[1] - 1, w, 0, NULL};
    for (int k = 0; k < 2; ++k) { /* layers lo.z - 1 and lo.z */
        lc.cz = lo[2] - 1 + k;
        lc.dst = buf[k];
        vxo_parallel_for((int64_t)layer_chunks, threads, 8, layer_body, &lc);
    }
    for (int32_t z = 0; z < nz; ++z) {
        uint16_t *prev = buf[z % 3], *cur = buf[(z + 1) % 3], *next = buf[(z + 2) % 3];
        lc.cz = lo[2] + z + 1;
        lc.dst = next;
        vxo_parallel_for((int64_t)layer_chunks, threads, 8, layer_body, &lc);
        world_ctx wc = {prev, cur, next, nx, w, mode, hashes + (size_t)z * nx * ny, counts + (size_t)z * nx * ny};
        vxo_parallel_for((int64_t)nx * ny, threads, 4, world_body, &wc);
    }
    for (int k = 0; k < 3; ++k) free(buf[k]);
    return 0;
}

int vxo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}
