//! Raw FFI to libvxp.so — one declaration per entry point of `include/vxp.h`.
//! UNCOMPILED: the build image has no rustc; keep in sync with the header by hand.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct vxp_ctx {
    _private: [u8; 0],
}

pub type vxp_status = c_int;
pub const VXP_OK: vxp_status = 0;
pub const VXP_ERR_BAD_ARG: vxp_status = 1;
pub const VXP_ERR_NO_DEVICE: vxp_status = 2;
pub const VXP_ERR_OOM: vxp_status = 3;
pub const VXP_ERR_CAPACITY: vxp_status = 4;
pub const VXP_ERR_CUDA: vxp_status = 5;

pub type vxp_mode = c_int;
pub const VXP_MODE_CULLED: vxp_mode = 0;
pub const VXP_MODE_GREEDY: vxp_mode = 1;

pub type vxp_pattern = c_int;

pub const VXP_CHUNK_VOXELS: usize = 32768;

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq, Hash)]
pub struct vxp_coord {
    pub x: i32,
    pub y: i32,
    pub z: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct vxp_chunk_meta {
    pub first_quad: u32,
    pub quad_count: u32,
}

/// One voxel edit: voxel (x,y,z) of stored chunk `chunk` becomes `id` (0 = air).  12 bytes.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct vxp_edit {
    pub chunk: u32,
    pub x: u16,
    pub y: u16,
    pub z: u16,
    pub id: u16,
}

#[repr(C)]
pub struct vxp_mesh_dev {
    pub verts: *mut u64,
    pub indices: *mut u32,
    pub meta: *mut vxp_chunk_meta,
    pub total_quads: *mut u64,
    pub capacity_quads: u64,
}

#[repr(C)]
pub struct vxp_mesh_host {
    pub verts: *const u64,
    pub indices: *const u32,
    pub meta: *const vxp_chunk_meta,
    pub total_quads: u64,
}

extern "C" {
    pub fn vxp_ctx_create(device: c_int, out: *mut *mut vxp_ctx) -> vxp_status;
    pub fn vxp_ctx_destroy(ctx: *mut vxp_ctx);
    pub fn vxp_ctx_set_stream(ctx: *mut vxp_ctx, cuda_stream: *mut c_void) -> vxp_status;
    pub fn vxp_ctx_reset_stream(ctx: *mut vxp_ctx) -> vxp_status;
    pub fn vxp_ctx_set_overlap(ctx: *mut vxp_ctx, enabled: c_int) -> vxp_status;
    pub fn vxp_ctx_set_host_indices(ctx: *mut vxp_ctx, enabled: c_int) -> vxp_status;
    pub fn vxp_ctx_synchronize(ctx: *mut vxp_ctx) -> vxp_status;
    pub fn vxp_status_str(s: vxp_status) -> *const c_char;
    pub fn vxp_last_error(ctx: *const vxp_ctx) -> *const c_char;
    pub fn vxp_launch_count(ctx: *const vxp_ctx) -> u64;
    pub fn vxp_spec_version() -> c_int;

    pub fn vxp_nbr6_from_coords(coords: *const vxp_coord, n: u32, nbr6: *mut i32) -> vxp_status;
    pub fn vxp_host_alloc(bytes: usize, out: *mut *mut c_void) -> vxp_status;
    pub fn vxp_host_free(p: *mut c_void);

    pub fn vxp_terrain_fill_batch(ctx: *mut vxp_ctx, d_coords: *const vxp_coord, n: u32, seed: u64,
                                  d_voxels: *mut u16) -> vxp_status;
    pub fn vxp_pattern_fill_batch(ctx: *mut vxp_ctx, d_coords: *const vxp_coord, n: u32,
                                  pattern: vxp_pattern, d_voxels: *mut u16) -> vxp_status;
    pub fn vxp_mesh_batch(ctx: *mut vxp_ctx, d_voxels: *const u16, d_nbr6: *const i32, n: u32,
                          mode: vxp_mode, out: *const vxp_mesh_dev) -> vxp_status;
    pub fn vxp_terrain_mesh_fused(ctx: *mut vxp_ctx, d_coords: *const vxp_coord, n: u32, seed: u64,
                                  mode: vxp_mode, out: *const vxp_mesh_dev) -> vxp_status;

    pub fn vxp_mesh_list(ctx: *mut vxp_ctx, d_voxels: *const u16, d_nbr6: *const i32, n: u32, d_list: *const u32,
                         count: u32, mode: vxp_mode, out: *const vxp_mesh_dev) -> vxp_status;
    pub fn vxp_remesh_edits(ctx: *mut vxp_ctx, d_voxels: *mut u16, d_nbr6: *const i32, n: u32,
                            d_edits: *const vxp_edit, m: u32, mode: vxp_mode, d_dirty: *mut u32,
                            d_dirty_count: *mut u32, out: *const vxp_mesh_dev) -> vxp_status;

    pub fn vxp_pack_chunks(ctx: *mut vxp_ctx, d_voxels: *const u16, n: u32, d_packed: *mut u8, capacity_bytes: u64,
                           d_offsets: *mut u64, d_total_bytes: *mut u64) -> vxp_status;
    pub fn vxp_unpack_chunks(ctx: *mut vxp_ctx, d_packed: *const u8, packed_bytes: u64, d_offsets: *const u64, n: u32,
                             d_voxels: *mut u16) -> vxp_status;

    pub fn vxp_downsample_chunks(ctx: *mut vxp_ctx, d_voxels: *const u16, d_child8: *const i32, n_out: u32,
                                 d_out: *mut u16) -> vxp_status;

    pub fn vxp_terrain_fill_host(ctx: *mut vxp_ctx, coords: *const vxp_coord, n: u32, seed: u64,
                                 voxels: *mut u16) -> vxp_status;
    pub fn vxp_mesh_chunks_host(ctx: *mut vxp_ctx, voxels: *const u16, nbr6: *const i32, n: u32,
                                n_stored: u32, mode: vxp_mode, out: *mut vxp_mesh_host) -> vxp_status;
    pub fn vxp_terrain_mesh_host(ctx: *mut vxp_ctx, coords: *const vxp_coord, n: u32, seed: u64,
                                 mode: vxp_mode, out: *mut vxp_mesh_host) -> vxp_status;
    pub fn vxp_mesh_packed_host(ctx: *mut vxp_ctx, packed: *const u8, packed_bytes: u64, offsets: *const u64,
                                nbr6: *const i32, n: u32, n_stored: u32, mode: vxp_mode,
                                out: *mut vxp_mesh_host) -> vxp_status;
    pub fn vxp_remesh_edits_host(ctx: *mut vxp_ctx, edits: *const vxp_edit, m: u32, mode: vxp_mode,
                                 out: *mut vxp_mesh_host, dirty: *mut *const u32, count: *mut u32) -> vxp_status;
}
