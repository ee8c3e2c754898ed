//! Safe chunk / mesher surface over `vxp-sys`: chunk coordinate -> voxel array -> mesh buffers.
//! UNCOMPILED (no Rust toolchain in the build image); mirrors `voxelphile_b200/mesher.py`, which
//! is the tested host mirror.  Style follows the reference: payload-free error enums through
//! `Result` (src/net/udp/mod.rs:16-28, src/lib.rs:37-45), `pub type` aliases (src/net/udp/mod.rs:14).

use std::ptr;
use vxp_sys as sys;

pub type BlockId = u16;
pub type ChunkCoord = sys::vxp_coord;
pub const CHUNK_DIM: usize = 32;
pub const CHUNK_VOXELS: usize = sys::VXP_CHUNK_VOXELS;

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum MeshError {
    BadArg,
    NoDevice,
    OutOfMemory,
    Capacity,
    Cuda,
}

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum MeshMode {
    Culled,
    Greedy,
}

fn check(s: sys::vxp_status) -> Result<(), MeshError> {
    match s {
        sys::VXP_OK => Ok(()),
        sys::VXP_ERR_BAD_ARG => Err(MeshError::BadArg),
        sys::VXP_ERR_NO_DEVICE => Err(MeshError::NoDevice),
        sys::VXP_ERR_OOM => Err(MeshError::OutOfMemory),
        sys::VXP_ERR_CAPACITY => Err(MeshError::Capacity),
        _ => Err(MeshError::Cuda),
    }
}

/// 32^3 block ids, x fastest (SPEC.md §1).
pub struct Chunk {
    pub coord: ChunkCoord,
    pub voxels: Box<[BlockId; CHUNK_VOXELS]>,
}

/// Packed vertices (8 B) and indices of one chunk (SPEC.md §5).  `indices` is empty after
/// [`Mesher::set_indices`]`(false)`: SPEC §5 makes them `4q + {0,1,2,2,3,0}`, one static buffer serves every mesh.
pub struct Mesh {
    pub vertices: Vec<u64>,
    pub indices: Vec<u32>,
}

pub struct Mesher {
    ctx: *mut sys::vxp_ctx,
}

// One context is driven by one thread at a time; it may move between threads.
unsafe impl Send for Mesher {}

impl Mesher {
    pub fn new(device: i32) -> Result<Self, MeshError> {
        let mut ctx = ptr::null_mut();
        check(unsafe { sys::vxp_ctx_create(device, &mut ctx) })?;
        Ok(Mesher { ctx })
    }

    /// Whether host results carry an index buffer (default true; `vxp_ctx_set_host_indices`).
    pub fn set_indices(&mut self, enabled: bool) -> Result<(), MeshError> {
        check(unsafe { sys::vxp_ctx_set_host_indices(self.ctx, enabled as i32) })
    }

    /// Neighbour table of a coordinate list (-1 = no chunk on that side).
    pub fn neighbours(coords: &[ChunkCoord]) -> Result<Vec<[i32; 6]>, MeshError> {
        let mut out = vec![[-1i32; 6]; coords.len()];
        check(unsafe {
            sys::vxp_nbr6_from_coords(coords.as_ptr(), coords.len() as u32, out.as_mut_ptr() as *mut i32)
        })?;
        Ok(out)
    }

    /// Terrain of `coords` for `seed` (SPEC.md §6).
    pub fn generate(&mut self, coords: &[ChunkCoord], seed: u64) -> Result<Vec<Chunk>, MeshError> {
        let mut flat = vec![0u16; coords.len() * CHUNK_VOXELS];
        check(unsafe {
            sys::vxp_terrain_fill_host(self.ctx, coords.as_ptr(), coords.len() as u32, seed, flat.as_mut_ptr())
        })?;
        Ok(coords
            .iter()
            .zip(flat.chunks_exact(CHUNK_VOXELS))
            .map(|(c, v)| {
                let mut b = Box::new([0u16; CHUNK_VOXELS]);
                b.copy_from_slice(v);
                Chunk { coord: *c, voxels: b }
            })
            .collect())
    }

    /// Mesh stored chunks; neighbours are looked up by coordinate inside the batch.
    pub fn mesh(&mut self, chunks: &[Chunk], mode: MeshMode) -> Result<Vec<Mesh>, MeshError> {
        let coords: Vec<ChunkCoord> = chunks.iter().map(|c| c.coord).collect();
        let nbr = Self::neighbours(&coords)?;
        let mut flat = Vec::with_capacity(chunks.len() * CHUNK_VOXELS);
        for c in chunks {
            flat.extend_from_slice(&c.voxels[..]);
        }
        let mut host = sys::vxp_mesh_host { verts: ptr::null(), indices: ptr::null(), meta: ptr::null(), total_quads: 0 };
        check(unsafe {
            sys::vxp_mesh_chunks_host(self.ctx, flat.as_ptr(), nbr.as_ptr() as *const i32, chunks.len() as u32,
                                      chunks.len() as u32, mode_of(mode), &mut host)
        })?;
        Ok(unsafe { split(&host, chunks.len()) })
    }

    /// Terrain + mesh without materialising voxels.
    pub fn generate_and_mesh(&mut self, coords: &[ChunkCoord], seed: u64, mode: MeshMode)
                             -> Result<Vec<Mesh>, MeshError> {
        let mut host = sys::vxp_mesh_host { verts: ptr::null(), indices: ptr::null(), meta: ptr::null(), total_quads: 0 };
        check(unsafe {
            sys::vxp_terrain_mesh_host(self.ctx, coords.as_ptr(), coords.len() as u32, seed, mode_of(mode), &mut host)
        })?;
        Ok(unsafe { split(&host, coords.len()) })
    }
}

/// A voxel edit addressed by position in the batch last passed to [`Mesher::mesh`].
#[derive(Clone, Copy, Debug)]
pub struct Edit {
    pub chunk: u32,
    pub x: u16,
    pub y: u16,
    pub z: u16,
    pub block: u16,
}

impl Mesher {
    /// Edit the world that the last `mesh` call left resident on the device and re-mesh what the
    /// edits invalidate (SPEC §9).  Returns (index of the chunk in that batch, its new mesh) pairs.
    pub fn remesh(&mut self, edits: &[Edit], mode: MeshMode) -> Result<Vec<(u32, Mesh)>, MeshError> {
        let raw: Vec<sys::vxp_edit> =
            edits.iter().map(|e| sys::vxp_edit { chunk: e.chunk, x: e.x, y: e.y, z: e.z, id: e.block }).collect();
        let mut host = sys::vxp_mesh_host { verts: ptr::null(), indices: ptr::null(), meta: ptr::null(), total_quads: 0 };
        let (mut dirty, mut count) = (ptr::null::<u32>(), 0u32);
        check(unsafe {
            sys::vxp_remesh_edits_host(self.ctx, raw.as_ptr(), raw.len() as u32, mode_of(mode), &mut host, &mut dirty, &mut count)
        })?;
        let ids = unsafe { std::slice::from_raw_parts(dirty, count as usize) };
        Ok(ids.iter().copied().zip(unsafe { split(&host, count as usize) }).collect())
    }
}

fn mode_of(m: MeshMode) -> sys::vxp_mode {
    match m {
        MeshMode::Culled => sys::VXP_MODE_CULLED,
        MeshMode::Greedy => sys::VXP_MODE_GREEDY,
    }
}

unsafe fn split(host: &sys::vxp_mesh_host, n: usize) -> Vec<Mesh> {
    let meta = std::slice::from_raw_parts(host.meta, n);
    meta.iter()
        .map(|m| {
            let (first, cnt) = (m.first_quad as usize, m.quad_count as usize);
            if cnt == 0 {
                return Mesh { vertices: Vec::new(), indices: Vec::new() };
            }
            Mesh {
                vertices: std::slice::from_raw_parts(host.verts.add(4 * first), 4 * cnt).to_vec(),
                indices: if host.indices.is_null() { Vec::new() }
                         else { std::slice::from_raw_parts(host.indices.add(6 * first), 6 * cnt).to_vec() },
            }
        })
        .collect()
}

impl Drop for Mesher {
    fn drop(&mut self) {
        unsafe { sys::vxp_ctx_destroy(self.ctx) }
    }
}
