// vxp_launch.h — host-side launchers of the sm_100a kernels (internal, C++ linkage).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vxp.h"

namespace vxp {
// per-device one-time setup (opt-in shared memory size); call with the device current
cudaError_t configure_device();
cudaError_t launch_terrain_fill(const vxp_coord* d_coords, uint32_t n, uint64_t seed, uint16_t* d_voxels,
                                cudaStream_t stream);
cudaError_t launch_pattern_fill(const vxp_coord* d_coords, uint32_t n, int pattern, uint16_t* d_voxels,
                                cudaStream_t stream);
// both zero *total_quads first (one memset node), then launch one kernel
cudaError_t launch_mesh(const uint16_t* d_voxels, const int32_t* d_nbr6, uint32_t n, int greedy,
                        const vxp_mesh_dev& out, cudaStream_t stream);
cudaError_t launch_terrain_mesh(const vxp_coord* d_coords, uint32_t n, uint64_t seed, int greedy,
                                const vxp_mesh_dev& out, cudaStream_t stream);
size_t mesh_smem_bytes();
}  // namespace vxp
