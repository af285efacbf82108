/* TEST INFRASTRUCTURE (oracle) — C/OpenMP restatement of the reference's fused step, used (a) as a fast second oracle
 * for full-size parity runs (C1: 128^3 x 1000 steps) and (b) as the multi-threaded CPU baseline in bench.py.
 * Never linked into or called by the product (xlb_b200).  Built by oracle/Makefile into oracle/liblbm_ref.so. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum {
  BC_NONE = 0, BC_EQUILIBRIUM = 1, BC_DO_NOTHING = 2, BC_HALFWAY = 3, BC_FULLWAY = 4, BC_ZOUHE_VELOCITY = 5,
  BC_ZOUHE_PRESSURE = 6, BC_REGULARIZED_VELOCITY = 7, BC_REGULARIZED_PRESSURE = 8, BC_OUTFLOW = 9
};

typedef struct LbmDesc {
  int d, q, nx, ny, nz;
  int collision;   /* 0 BGK, 1 KBC, 2 SmagorinskyLESBGK */
  int compute;     /* 1 f32, 2 f64 */
  int store;       /* 0 f16, 1 f32, 2 f64 */
  double omega;
  int c[3 * 27];   /* c[a*27 + l] */
  int opp[27];
  double w[27];
  double cc[27 * 6];
  double qi[27 * 6];
  int bc_kind[256];
  double bc_rho[256];
  double bc_u[256 * 3];
  int has_force;      /* ForcedCollision + ExactDifference (forced_collision.py:34-39) */
  double force[3];
  double smagorinsky; /* SmagorinskyLESBGK coefficient (smagorinsky_les_bgk.py:24) */
} LbmDesc;

#define REAL float
#define SUFFIX _f32
#include "lbm_ref_core.h"
#undef REAL
#undef SUFFIX
#define REAL double
#define SUFFIX _f64
#include "lbm_ref_core.h"
#undef REAL
#undef SUFFIX

/* Runs nsteps steps with the caller's swap convention; returns 0 if the final populations are in fa, 1 if in fb. */
int lbm_ref_run(const LbmDesc* d, void* fa, void* fb, const unsigned char* bc_mask, const unsigned char* missing, int nsteps, int nthreads) {
  void *f0 = fa, *f1 = fb;
  for (int t = 0; t < nsteps; ++t) {
    if (d->compute == 1) step_f32(d, f0, f1, bc_mask, missing, nthreads);
    else step_f64(d, f0, f1, bc_mask, missing, nthreads);
    void* tmp = f0;
    f0 = f1;
    f1 = tmp;
  }
  return f0 == fa ? 0 : 1;
}

int lbm_ref_sizeof_desc(void) { return (int)sizeof(LbmDesc); }
