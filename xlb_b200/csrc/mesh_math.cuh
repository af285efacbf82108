// Triangle / unit-voxel overlap test of the mesh voxeliser (replaces the functions pre_compute, triangle_box_intersect,
// mesh_voxel_intersect of xlb/operator/boundary_masker/mesh_boundary_masker.py:60-148).
//
// The test is Schwarz & Seidel, "Fast parallel surface and solid voxelization on GPUs" (2010), which the reference cites: a
// triangle overlaps the box [low, low + 1]^3 iff (a) its plane separates the box's two extreme corners along the normal and
// (b) in each of the three axis-aligned projections no edge function puts the box outside the projected triangle.
//
// EDGE_TEST selects how (b) is evaluated:
//   XLBN_MESH_SCHWARZ_SEIDEL  the published form: edge normal n_e = sgn * (-e[ax1], e[ax0]), offset
//                             d_e = -(n_e . v[ax0, ax1]) + max(0, n_e.x) + max(0, n_e.y).
//   XLBN_MESH_REFERENCE_LITERAL  the reference's pre_compute AS WRITTEN (L78-97): both components of the edge normal are
//                             taken from e[ax0] and both offsets from v[ax0], so the vertex terms cancel and the test
//                             degenerates to  sgn * e[ax0] * (low[ax1] - low[ax0]) + |e[ax0]| >= 0  — only voxels with
//                             |i-j|, |j-k|, |k-i| <= 1 can pass.  Kept so that the masks can be
//                             compared bit for bit with the reference's kernel; not useful for simulations.
// fp32 throughout, one rounding per operation in the reference's order (this TU is built with -fmad=false).
#pragma once

#include "common.cuh"

namespace xlbn {

enum { XLBN_MESH_SCHWARZ_SEIDEL_ = 0, XLBN_MESH_REFERENCE_LITERAL_ = 1 };

struct TriSetup {
  float n[3];         // unit normal: normalize(cross(v1 - v0, v2 - v0))   (wp.mesh_eval_face_normal)
  float dist1, dist2;
  float ne0[3][3], ne1[3][3], de[3][3];  // [edge i][projection axis ax0]
  float lo[3], hi[3];                    // bounding box of the triangle
  bool valid;                            // length(normal) > 0  (L114)
};

XLBN_MATH float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
XLBN_MATH float max0(float x) { return x > 0.0f ? x : 0.0f; }

// pre_compute (mesh_boundary_masker.py:65-99) for the triangle (v0, v1, v2)
template <int EDGE_TEST>
XLBN_MATH void tri_setup(const float* v0, const float* v1, const float* v2, TriSetup& t) {
  const float* v[3] = {v0, v1, v2};
  float a[3], b[3], c[3];
  for (int k = 0; k < 3; ++k) {
    a[k] = v1[k] - v0[k];
    b[k] = v2[k] - v0[k];
    t.lo[k] = fminf(v0[k], fminf(v1[k], v2[k]));
    t.hi[k] = fmaxf(v0[k], fmaxf(v1[k], v2[k]));
  }
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  const float len = sqrtf(dot3(c, c));
  t.valid = len > 0.0f;
  for (int k = 0; k < 3; ++k) t.n[k] = t.valid ? c[k] / len : 0.0f;
  float corner[3], d1[3], d2[3];
  for (int k = 0; k < 3; ++k) {
    corner[k] = t.n[k] > 0.0f ? 1.0f : 0.0f;
    d1[k] = corner[k] - v0[k];
    d2[k] = (1.0f - corner[k]) - v0[k];
  }
  t.dist1 = dot3(t.n, d1);
  t.dist2 = dot3(t.n, d2);
  float e[3][3];  // edges[i] = v[(i+1)%3] - v[i]
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) e[i][k] = v[(i + 1) % 3][k] - v[i][k];
  for (int ax0 = 0; ax0 < 3; ++ax0) {
    const int ax1 = (ax0 + 1) % 3, ax2 = (ax0 + 2) % 3;
    const float sgn = t.n[ax2] < 0.0f ? -1.0f : 1.0f;
    for (int i = 0; i < 3; ++i) {
      if (EDGE_TEST == XLBN_MESH_REFERENCE_LITERAL_) {
        t.ne0[i][ax0] = -1.0f * sgn * e[i][ax0];
        t.ne1[i][ax0] = sgn * e[i][ax0];
        t.de[i][ax0] = (-1.0f * (t.ne0[i][ax0] * v[i][ax0] + t.ne1[i][ax0] * v[i][ax0]) + max0(t.ne0[i][ax0])) + max0(t.ne1[i][ax0]);
      } else {
        t.ne0[i][ax0] = -1.0f * sgn * e[i][ax1];
        t.ne1[i][ax0] = sgn * e[i][ax0];
        t.de[i][ax0] = (-1.0f * (t.ne0[i][ax0] * v[i][ax0] + t.ne1[i][ax0] * v[i][ax1]) + max0(t.ne0[i][ax0])) + max0(t.ne1[i][ax0]);
      }
    }
  }
}

// triangle_box_intersect (L110-128) for the unit box at `low`, preceded by the bounding-box overlap that
// wp.mesh_query_aabb applies (inclusive: L136)
XLBN_MATH bool tri_box_overlap(const TriSetup& t, const float* low) {
  for (int k = 0; k < 3; ++k)
    if (t.lo[k] > low[k] + 1.0f || t.hi[k] < low[k]) return false;
  if (!t.valid) return false;
  const float nl = dot3(t.n, low);
  if (!((nl + t.dist1) * (nl + t.dist2) <= 0.0f)) return false;
  bool hit = true;
  for (int ax0 = 0; ax0 < 3; ++ax0) {
    const int ax1 = (ax0 + 1) % 3;
    for (int i = 0; i < 3; ++i) hit = hit && ((t.ne0[i][ax0] * low[ax0] + t.ne1[i][ax0] * low[ax1]) + t.de[i][ax0] >= 0.0f);
  }
  return hit;
}

// Voxels whose box [i, i+1] overlaps the triangle's bounding box inclusively, clipped to the padded volume [-1, n]
// (the reference also queries neighbour voxels one cell outside the grid, L182-188).  false: nothing to visit.
XLBN_MATH bool tri_voxel_range(const TriSetup& t, const int* n, int* lo, int* hi) {
  for (int k = 0; k < 3; ++k) {
    const int a = (int)ceilf(t.lo[k] - 1.0f), b = (int)floorf(t.hi[k]);
    lo[k] = a < -1 ? -1 : a;
    hi[k] = b > n[k] ? n[k] : b;
    if (lo[k] > hi[k]) return false;
  }
  return true;
}

}  // namespace xlbn
