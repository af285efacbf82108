// TEST HARNESS — the mesh voxeliser's per-triangle code (xlb_b200/csrc/mesh_math.cuh: tri_setup, tri_voxel_range,
// tri_box_overlap — what mesh_mark_kernel runs per warp) compiled for the host.  Built with -fmad=false like the library TU.
#define XLBN_HOST_MIRROR 1
#include "../../xlb_b200/csrc/mesh_math.cuh"

using namespace xlbn;

template <int EDGE_TEST>
static void mark(const float* verts, long n_tri, const int* n, unsigned char* solid) {
  for (long tri = 0; tri < n_tri; ++tri) {
    const float* p = verts + tri * 9;
    TriSetup t;
    tri_setup<EDGE_TEST>(p, p + 3, p + 6, t);
    int lo[3], hi[3];
    if (!tri_voxel_range(t, n, lo, hi)) continue;
    for (int i = lo[0]; i <= hi[0]; ++i)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int k = lo[2]; k <= hi[2]; ++k) {
          const float low[3] = {(float)i, (float)j, (float)k};
          if (tri_box_overlap(t, low)) solid[((long)(i + 1) * (n[1] + 2) + (j + 1)) * (n[2] + 2) + (k + 1)] = 1;
        }
  }
}

// solid: (nx+2)(ny+2)(nz+2) bytes, zero-initialised by the caller
extern "C" int mirror_mesh_solid(const float* verts, long n_tri, const int* dims, int edge_test, unsigned char* solid) {
  if (edge_test == 0) mark<XLBN_MESH_SCHWARZ_SEIDEL_>(verts, n_tri, dims, solid);
  else mark<XLBN_MESH_REFERENCE_LITERAL_>(verts, n_tri, dims, solid);
  return 0;
}
