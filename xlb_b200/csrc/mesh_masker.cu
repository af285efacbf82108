// MeshBoundaryMasker on the device (replaces xlb/operator/boundary_masker/mesh_boundary_masker.py:49-236, which leans on
// Warp's BVH: wp.Mesh, wp.mesh_query_aabb).  Surface voxelisation without a tree:
//   1. one warp per TRIANGLE walks the voxels under the triangle's bounding box and runs the exact triangle / unit-box test
//      (mesh_math.cuh) — the same (triangle, voxel) pairs the reference visits voxel-first through its BVH query — marking
//      hits in a byte volume that is one cell larger than the grid on every side (the reference also tests neighbour voxels
//      that lie outside the grid, L182-188);
//   2. one thread per CELL: marked -> bc_mask = 255 (solid, skipped by the stepper); otherwise every direction l whose
//      neighbour voxel is marked makes the cell a boundary cell: bc_mask = id, missing[opp[l]] = true (L176-188).
// Work is O(sum of bounding-box volumes) instead of O(cells x 27 BVH queries).
#include "lbm_math.cuh"
#include "mesh_math.cuh"

namespace xlbn {

struct MeshGeom {
  int nx, ny, nz;  // grid extents; the solid volume is (nx+2)(ny+2)(nz+2), voxel (i,j,k) at [(i+1),(j+1),(k+1)]
};

__device__ __forceinline__ long long pad_index(const MeshGeom& g, int i, int j, int k) {
  return ((long long)(i + 1) * (g.ny + 2) + (j + 1)) * (g.nz + 2) + (k + 1);
}

template <int EDGE_TEST>
__global__ void mesh_mark_kernel(const float* __restrict__ verts, long long n_tri, MeshGeom g, uint8_t* __restrict__ solid) {
  const long long tri = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (tri >= n_tri) return;
  const float* p = verts + tri * 9;
  // the reference's vertex order: v0 = eval(1,0) = first, v1 = eval(0,1) = second, v2 = eval(0,0) = third vertex (L138-140)
  float v0[3] = {p[0], p[1], p[2]}, v1[3] = {p[3], p[4], p[5]}, v2[3] = {p[6], p[7], p[8]};
  TriSetup t;
  tri_setup<EDGE_TEST>(v0, v1, v2, t);
  int lo[3], hi[3];
  const int n[3] = {g.nx, g.ny, g.nz};
  if (!tri_voxel_range(t, n, lo, hi)) return;
  const int ry = hi[1] - lo[1] + 1, rz = hi[2] - lo[2] + 1;
  const long long total = (long long)(hi[0] - lo[0] + 1) * ry * rz;
  for (long long w = lane; w < total; w += 32) {
    const int i = lo[0] + (int)(w / ((long long)ry * rz)), j = lo[1] + (int)((w / rz) % ry), k = lo[2] + (int)(w % rz);
    const float low[3] = {(float)i, (float)j, (float)k};
    if (tri_box_overlap(t, low)) solid[pad_index(g, i, j, k)] = 1;
  }
}

template <class L>
__global__ void mesh_classify_kernel(MeshGeom g, const uint8_t* __restrict__ solid, int bc_id, uint8_t* __restrict__ bc_mask, uint8_t* __restrict__ missing) {
  const long long cells = (long long)g.nx * g.ny * g.nz;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  const int k = (int)(c % g.nz), j = (int)((c / g.nz) % g.ny), i = (int)(c / ((long long)g.nz * g.ny));
  if (solid[pad_index(g, i, j, k)]) {
    bc_mask[c] = 255;  // L170-172
    return;
  }
  XLBN_FOR(L::Q, l)
    if constexpr (l > 0) {
      if (solid[pad_index(g, i + L::c(0, l), j + L::c(1, l), k + L::c(2, l))]) {  // L182-188
        bc_mask[c] = (uint8_t)bc_id;
        missing[(long long)L::opp(l) * cells + c] = 1;
      }
    }
  XLBN_END
}

}  // namespace xlbn

using namespace xlbn;

extern "C" int xlbn_mask_mesh(int lattice, const float* vertices, long long n_triangles, int bc_id, int edge_test, const int32_t dims[3],
                              uint8_t* bc_mask, uint8_t* missing, uint8_t* solid_scratch, void* stream) {
  if (lattice != XLBN_D3Q19 && lattice != XLBN_D3Q27) return fail(XLBN_E_UNSUPPORTED, "xlbn_mask_mesh: 3-D lattices only (mesh_boundary_masker.py:27-28)");
  if (!vertices || !dims || !bc_mask || !missing || !solid_scratch) return fail(XLBN_E_ARG, "xlbn_mask_mesh: NULL argument");
  if (n_triangles < 0 || bc_id < 1 || bc_id > 254) return fail(XLBN_E_ARG, "xlbn_mask_mesh: n_triangles %lld, id %d", n_triangles, bc_id);
  if (edge_test != XLBN_MESH_SCHWARZ_SEIDEL && edge_test != XLBN_MESH_REFERENCE_LITERAL) return fail(XLBN_E_ARG, "xlbn_mask_mesh: edge_test %d", edge_test);
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(XLBN_E_SHAPE, "xlbn_mask_mesh: dims %d %d %d", dims[0], dims[1], dims[2]);
  const MeshGeom g = {dims[0], dims[1], dims[2]};
  cudaStream_t st = (cudaStream_t)stream;
  const long long padded = (long long)(g.nx + 2) * (g.ny + 2) * (g.nz + 2), cells = (long long)g.nx * g.ny * g.nz;
  XLBN_CUDA_OK(cudaMemsetAsync(solid_scratch, 0, (size_t)padded, st));
  if (n_triangles > 0) {
    const long long threads = n_triangles * 32;
    const unsigned blocks = (unsigned)((threads + 127) / 128);
    if (edge_test == XLBN_MESH_SCHWARZ_SEIDEL) mesh_mark_kernel<XLBN_MESH_SCHWARZ_SEIDEL_><<<blocks, 128, 0, st>>>(vertices, n_triangles, g, solid_scratch);
    else mesh_mark_kernel<XLBN_MESH_REFERENCE_LITERAL_><<<blocks, 128, 0, st>>>(vertices, n_triangles, g, solid_scratch);
    XLBN_LAUNCH_OK("mesh_mark_kernel");
  }
  const unsigned cblocks = (unsigned)((cells + 255) / 256);
  if (lattice == XLBN_D3Q19) mesh_classify_kernel<D3Q19><<<cblocks, 256, 0, st>>>(g, solid_scratch, bc_id, bc_mask, missing);
  else mesh_classify_kernel<D3Q27><<<cblocks, 256, 0, st>>>(g, solid_scratch, bc_id, bc_mask, missing);
  XLBN_LAUNCH_OK("mesh_classify_kernel");
  return 0;
}
