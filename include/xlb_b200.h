/*
 * xlb_b200 — C ABI of the B200-native (sm_100a) fused lattice-Boltzmann step.
 *
 * This header is the drop-in boundary.  The reference (Autodesk/XLB) has no native code: every
 * operator is a Python class whose `__call__` dispatches to `jax_implementation` / `warp_implementation`
 * (reference: xlb/operator/operator.py:39-74) and the device work is `wp.launch(kernel, inputs=[...], dim=...)`
 * or one jitted XLA executable.  Each entry point below replaces exactly one such launch site; the reference
 * file:line it stands in for is cited next to it.  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add inside the corresponding `@Operator.register_backend(...)` method.
 *
 * Conventions
 *   - plain C types only: device pointers as void*, sizes as int / long long, a CUDA stream as void* (cudaStream_t;
 *     NULL = legacy default stream).  No torch / warp / jax types.
 *   - arrays are the reference's layout: [cardinality][nx][ny][nz], C-contiguous (z unit-stride, x slowest)
 *     (reference: xlb/grid/warp_grid.py:17-32).  2-D fields ([q][nx][ny] or [q][nx][ny][1]) are passed with
 *     dims = {nx, ny, 1}; the library maps them internally so that the unit-stride axis is the thread axis.
 *   - every function returns 0 on success; < 0 for argument / shape / dtype / unsupported-combination errors
 *     (XLBN_E_*); > 0 is a cudaError_t.  xlbn_last_error() returns a thread-local human-readable message.
 *   - all launches are asynchronous on the given stream; nothing synchronises except xlbn_halo_* setup calls
 *     documented as such.  The library keeps no global mutable state besides the thread-local error string.
 */
#ifndef XLB_B200_H
#define XLB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XLBN_VERSION 100 /* 0.1.0 */

/* ---- enums ------------------------------------------------------------------------------------------------ */

/* lattice; index order of the discrete velocities is the reference's (SURVEY.md Appendix A;
 * reference: xlb/velocity_set/d2q9.py:18-21, d3q19.py:19, d3q27.py:19) */
enum xlbn_lattice { XLBN_D2Q9 = 0, XLBN_D3Q19 = 1, XLBN_D3Q27 = 2 };

/* array element types (reference: xlb/precision_policy.py:8-43) */
enum xlbn_dtype { XLBN_F16 = 0, XLBN_F32 = 1, XLBN_F64 = 2, XLBN_U8 = 3, XLBN_BOOL = 4 };

/* collision operators (reference: xlb/operator/collision/bgk.py:17-34, kbc.py:40-100, smagorinsky_les_bgk.py:37-90).
 * XLBN_COLLISION_FORCED is a FLAG or-ed onto a base operator: ForcedCollision with the ExactDifference forcing scheme
 * (collision/forced_collision.py:34-39, force/exact_difference_force.py:45-85).  Steppers get it through
 * xlbn_stepper_set_force; the stand-alone xlbn_collide_ext takes it in `collision`. */
enum xlbn_collision { XLBN_BGK = 0, XLBN_KBC = 1, XLBN_SMAGORINSKY_LES_BGK = 2, XLBN_COLLISION_FORCED = 4 };

/* boundary-condition kinds (reference: xlb/operator/boundary_condition/bc_*.py) */
enum xlbn_bc_kind {
  XLBN_BC_NONE = 0,
  XLBN_BC_EQUILIBRIUM = 1,          /* bc_equilibrium.py:58-86            (streaming step) */
  XLBN_BC_DO_NOTHING = 2,           /* bc_do_nothing.py:44-63             (streaming step) */
  XLBN_BC_HALFWAY_BOUNCE_BACK = 3,  /* bc_halfway_bounce_back.py:50-85    (streaming step) */
  XLBN_BC_FULLWAY_BOUNCE_BACK = 4,  /* bc_fullway_bounce_back.py:44-72    (collision step) */
  XLBN_BC_ZOUHE_VELOCITY = 5,       /* bc_zouhe.py:279-311                (streaming step, aux = normal velocity) */
  XLBN_BC_ZOUHE_PRESSURE = 6,       /* bc_zouhe.py:313-338                (streaming step, aux = density) */
  XLBN_BC_REGULARIZED_VELOCITY = 7, /* bc_regularized.py:134-168 */
  XLBN_BC_REGULARIZED_PRESSURE = 8, /* bc_regularized.py:170-202 */
  XLBN_BC_EXTRAPOLATION_OUTFLOW = 9 /* bc_extrapolation_outflow.py:152-195 (streaming part + post-collision aux) */
};

/* masker algorithm: the reference has two that differ on BC-free domain-face cells */
enum xlbn_mask_mode {
  XLBN_MASK_WARP = 0, /* indices_boundary_masker.py:103-224 */
  XLBN_MASK_JAX = 1   /* indices_boundary_masker.py:45-101  */
};

enum xlbn_error {
  XLBN_OK = 0,
  XLBN_E_ARG = -1,         /* null pointer, bad enum, negative size */
  XLBN_E_SHAPE = -2,       /* dims inconsistent with the operation */
  XLBN_E_DTYPE = -3,       /* dtype not accepted for this argument */
  XLBN_E_UNSUPPORTED = -4, /* e.g. KBC on D3Q19 (reference raises too: kbc.py:71-72, 184-185) */
  XLBN_E_STATE = -5        /* halo handle not connected, etc. */
};

/* ---- descriptors -------------------------------------------------------------------------------------------- */

/* One boundary condition of a stepper: the id written into bc_mask (registry order in the reference,
 * boundary_condition_registry.py:19-27; 1..254, 0 = fluid, 255 = solid/skip) and its parameters. */
typedef struct xlbn_bc_desc {
  int32_t id;
  int32_t kind;  /* xlbn_bc_kind */
  double rho;    /* EquilibriumBC density  (bc_equilibrium.py:42) */
  double u[3];   /* EquilibriumBC velocity (bc_equilibrium.py:43); u[2] ignored in 2-D */
} xlbn_bc_desc;

typedef struct xlbn_stepper_desc {
  int32_t lattice;       /* xlbn_lattice */
  int32_t collision;     /* xlbn_collision */
  int32_t compute_dtype; /* XLBN_F32 | XLBN_F64           (PrecisionPolicy.compute_precision) */
  int32_t store_dtype;   /* XLBN_F16 | XLBN_F32 | XLBN_F64 (PrecisionPolicy.store_precision) */
  int32_t n_bc;
  int32_t cells_per_thread; /* tuning knob: 0 = library default; 1, 2, 4, 8 = scalar path, that many z-cells per thread;
                               102, 104 = packed fp32x2 pair path (FFMA2; fp32 compute, fp32/fp16 storage only);
                               202 = half2-state pair path (FP32FP16 BGK only; the default for that policy);
                               203 = 202 with a leaner boundary variant for warps whose boundary cells are all
                                     FullwayBounceBack (same results; a tuning candidate);
                               301 = KBC only: register-lean formulation of the collision, one cell per thread (same algebra
                                     as kbc.py:268-296, feq recomputed per pass instead of held; rounding-level differences).
                                     This is what 0 selects for a KBC stepper, ForcedCollision(KBC) included (the ExactDifference
                                     term is added in the formulation's last pass).
                               300 = KBC only: the literal three-array formulation with the reference's own roundings (IEEE divisions,
                                     nothing fused): bit-identical to the reference kernel; the parity form, ~2x slower
                               402 = FP32FP16 BGK only: the persistent TMA-fed tile kernel (csrc/step_tile.cuh), what 0 selects where the
                                     slab can be tiled: 1024-cell tiles if they fit the plane, else 512-cell tiles;
                                     404 = 512-cell tiles; 403 = 512-cell tiles at three CTAs per SM (tuning variant).
                                     Needs nz | 512, nz % 8 == 0, ny % (512 / nz) == 0, no halo handle on the call
                               501 = BGK with fp32 storage (FP32FP32, FP64FP32), 3-D lattices: the scalar tile kernel — the same TMA-fed
                                     persistent pipeline with one cell per consumer thread (512-cell tiles, one CTA per SM); what 0
                                     selects for D3Q19 FP32FP32 where the slab can be tiled.  502 = two CTAs per SM (fp32 compute only).
                                     Needs nz | 512, nz % 16 == 0, ny % (512 / nz) == 0, no halo handle on the call */
  const xlbn_bc_desc* bcs;  /* n_bc entries, copied */
} xlbn_stepper_desc;

typedef struct xlbn_stepper xlbn_stepper; /* opaque: owns only the device BC table (8 KiB) */
typedef struct xlbn_halo xlbn_halo;       /* opaque: ghost planes + flags of one x-slab, peer-mapped to its 2 neighbours */

/* Field geometry of one call.  dims = extents of the (local slab of the) arrays.  [x_begin, x_begin + x_count) is the
 * range of x-planes to update (whole array: 0, nx); used for the boundary-plane / interior split when the halo
 * exchange is overlapped with the interior update. */
typedef struct xlbn_domain {
  int32_t nx, ny, nz;
  int32_t x_begin, x_count;
} xlbn_domain;

/* ---- library -------------------------------------------------------------------------------------------------- */

int xlbn_version(void);
const char* xlbn_last_error(void);

/* Lattice tables as compiled into the kernels, for cross-checking against the host tables.
 * c: [3*q] (kernel order cx[q], cy[q], cz[q]; 2-D: cz = 0), w: [q], opp: [q].  Returns q, or < 0. */
int xlbn_lattice_tables(int lattice, int32_t* c, double* w, int32_t* opp);

/* ---- the hot path: fused pull-stream + BC + collide + store ------------------------------------------------ */

/* Replaces: IncompressibleNavierStokesStepper._construct_warp (closure-specialised kernel, nse_stepper.py:245-383). */
int xlbn_stepper_create(const xlbn_stepper_desc* desc, xlbn_stepper** out);
int xlbn_stepper_destroy(xlbn_stepper* s);
/* Constant body force: the stepper's collision becomes ForcedCollision(collision, "exact_difference", force)
 * (nse_stepper.py:45-46; forced_collision.py:34-39: f_out += feq(rho, u + force) - feq(rho, u) after the collision, on
 * every cell that collides).  force: host double[3] (2-D lattices read the first two), NULL removes the force. */
int xlbn_stepper_set_force(xlbn_stepper* s, const double* force);
/* Smagorinsky coefficient of an XLBN_SMAGORINSKY_LES_BGK stepper (smagorinsky_les_bgk.py:24; default 0.17). */
int xlbn_stepper_set_smagorinsky(xlbn_stepper* s, double coefficient);

/* Publish `omega` for the steps that follow (stream-ordered, no host synchronisation).  Optional on one stream: xlbn_step
 * does it itself when omega differs from the previous call's (omega is a per-call argument of the reference kernel,
 * nse_stepper.py:351, and may change every step).  REQUIRED before one step is issued as several xlbn_step calls on
 * DIFFERENT streams (the slab path: face planes on a side stream, interior on the caller's): call it on a stream all of
 * them are ordered after.  Also the form to capture inside a CUDA graph that must not depend on earlier calls. */
int xlbn_stepper_prepare(xlbn_stepper* s, double omega, void* stream);

/* One time step.  Replaces the launch in IncompressibleNavierStokesStepper.warp_implementation
 * (nse_stepper.py:385-392; kernel body 344-381) and the jitted jax_implementation_pull (147-192).
 *   f0        in : populations at step t, store dtype, [q][nx][ny][nz]
 *                  (aux-recovery writes the prescribed value back into f0[0, cell] of Zou-He/Regularized cells exactly
 *                  like nse_stepper.py:318-342 — the only writes to f0)
 *   f1        out: populations at step t+1 (cells with bc_mask == 255 are not written, nse_stepper.py:356-358);
 *                  in: f1[0, cell] of Zou-He/Regularized cells holds the prescribed value (boundary_condition.py:151)
 *   bc_mask      : uint8 [nx][ny][nz]
 *   missing_bits : uint32 [nx][ny][nz], bit l set <=> missing_mask[l, cell] (see xlbn_pack_missing); may be NULL
 *                  when the stepper has no BC that reads it
 *   halo         : NULL (x is periodic inside the array) or the slab's halo handle: pulls across the slab faces read
 *                  the ghost planes of parity `timestep & 1`, and the outgoing populations of planes 0 / nx-1 are
 *                  ALSO stored straight into the neighbours' ghost planes of parity `(timestep + 1) & 1` through
 *                  peer-mapped pointers (fused compute + NVLink transfer, no separate exchange pass).
 */
int xlbn_step(xlbn_stepper* s, const void* f0, void* f1, const uint8_t* bc_mask, const uint32_t* missing_bits,
              const xlbn_domain* dom, double omega, int timestep, xlbn_halo* halo, void* stream);

/* ---- masks ---------------------------------------------------------------------------------------------------- */

/* Replaces IndicesBoundaryMasker (indices_boundary_masker.py:45-101 JAX, 103-224 Warp) for ONE boundary condition;
 * call once per BC in list order.  indices: int32 [3][n] device array of GLOBAL cell coordinates (2-D: z row = 0),
 * needs_padding: the BC's flag (boundary_condition.py:56).  global_dims / start: extents of the whole domain and the
 * global coordinate of local cell (0,0,0) (slab decomposition; = local dims and {0,0,0} on one GPU).
 * XLBN_MASK_JAX additionally needs a zero-initialised uint8 scratch `solid` of the local extents plus one halo cell on
 * every side, i.e. (nx+2)(ny+2)(nz+2) bytes (nz+2 -> 1 in 2-D is NOT applied: pass nz = 1 and the library pads it),
 * and a final xlbn_mask_finalize_jax call after the last BC; entries that were already set in the caller's missing mask are streamed
 * like the reference does (L56-63, 92) when a copy of that mask is passed as `incoming`. */
int xlbn_mask_indices(int lattice, int mode, const int32_t* indices, long long n, int bc_id, int needs_padding,
                      const int32_t global_dims[3], const int32_t start[3], const int32_t local_dims[3],
                      uint8_t* bc_mask, uint8_t* missing /* bool [q][nx][ny][nz] */, uint8_t* solid, void* stream);
int xlbn_mask_finalize_jax(int lattice, const int32_t global_dims[3], const int32_t start[3], const int32_t local_dims[3],
                           uint8_t* missing, const uint8_t* solid, const uint8_t* incoming /* copy of the caller's mask, or NULL if it was all false */,
                           void* stream);

/* Replaces MeshBoundaryMasker.warp_implementation (mesh_boundary_masker.py:49-236; 3-D only) for ONE mesh-based BC: surface
 * voxelisation of a triangle soup.  vertices: device float32 [n_triangles][3 vertices][3], grid units, the whole mesh inside
 * [0, dims) (the caller checks, as L209-216).  Voxels [i,i+1]^3 that a triangle overlaps become solid (bc_mask = 255); every
 * other cell with a solid neighbour in direction l gets bc_mask = bc_id and missing[opp[l]] = true.  Other cells are not touched.
 * edge_test: XLBN_MESH_SCHWARZ_SEIDEL = the published overlap test the reference cites; XLBN_MESH_REFERENCE_LITERAL = the
 * reference's pre_compute exactly as written (degenerate edge functions; for bit comparison with the reference only).
 * solid_scratch: device bytes, (nx+2)(ny+2)(nz+2); overwritten. */
enum xlbn_mesh_edge_test { XLBN_MESH_SCHWARZ_SEIDEL = 0, XLBN_MESH_REFERENCE_LITERAL = 1 };
int xlbn_mask_mesh(int lattice, const float* vertices, long long n_triangles, int bc_id, int edge_test, const int32_t dims[3],
                   uint8_t* bc_mask, uint8_t* missing /* bool [q][nx][ny][nz] */, uint8_t* solid_scratch, void* stream);

/* bool [q][n_cells] -> uint32 [n_cells] bitmask consumed by xlbn_step. */
int xlbn_pack_missing(int q, const uint8_t* missing, uint32_t* bits, long long n_cells, void* stream);

/* ---- stand-alone operators (API / test parity; same device functions as the fused kernel) -------------------- */
/* Arrays carry their own dtype code (xlbn_dtype F16/F32/F64); arithmetic is done in compute_dtype. dims = {nx,ny,nz}. */

/* Stream.warp_implementation / jax_implementation (stream.py:18-51, 85-113) */
int xlbn_stream(int lattice, const void* f_in, void* f_out, int dtype, const int32_t dims[3], void* stream);
/* QuadraticEquilibrium (quadratic_equilibrium.py:18-25, 63-97) */
int xlbn_equilibrium(int lattice, int compute_dtype, const void* rho, int rho_dtype, const void* u, int u_dtype, void* f,
                     int f_dtype, const int32_t dims[3], void* stream);
/* Macroscopic / ZeroMoment / FirstMoment (macroscopic.py:21-65, zero_moment.py, first_moment.py). rho or u may be NULL. */
int xlbn_macroscopic(int lattice, int compute_dtype, const void* f, int f_dtype, void* rho, int rho_dtype, void* u,
                     int u_dtype, const int32_t dims[3], void* stream);
/* FirstMoment with a given rho (first_moment.py:14-18) */
int xlbn_first_moment(int lattice, int compute_dtype, const void* f, int f_dtype, const void* rho, int rho_dtype, void* u,
                      int u_dtype, const int32_t dims[3], void* stream);
/* SecondMoment (second_moment.py:35-105): pi [d(d+1)/2][...] */
int xlbn_second_moment(int lattice, int compute_dtype, const void* f, int f_dtype, void* pi, int pi_dtype,
                       const int32_t dims[3], void* stream);
/* BGK / KBC (bgk.py:17-80, kbc.py:40-100, 299-347). rho is read only by KBC. */
int xlbn_collide(int lattice, int collision, int compute_dtype, const void* f, int f_dtype, const void* feq, int feq_dtype,
                 void* fout, int fout_dtype, const void* rho, int rho_dtype, double omega, const int32_t dims[3],
                 void* stream);
/* Any collision operator incl. SmagorinskyLESBGK and the ForcedCollision wrapper (smagorinsky_les_bgk.py:92-138,
 * forced_collision.py:41-103).  collision = base | XLBN_COLLISION_FORCED; rho is read by KBC and by the forcing, u
 * ([d][...]) by the forcing only; force: host double[3] or NULL. */
int xlbn_collide_ext(int lattice, int collision, int compute_dtype, const void* f, int f_dtype, const void* feq, int feq_dtype,
                     void* fout, int fout_dtype, const void* rho, int rho_dtype, const void* u, int u_dtype, double omega,
                     const double* force, double smagorinsky, const int32_t dims[3], void* stream);
/* ExactDifference stand-alone (exact_difference_force.py:87-125): fout = f_postcollision + feq(rho, u + force) - feq. */
int xlbn_exact_difference(int lattice, int compute_dtype, const void* f_postcollision, int f_dtype, const void* feq, int feq_dtype,
                          void* fout, int fout_dtype, const void* rho, int rho_dtype, const void* u, int u_dtype,
                          const double* force, const int32_t dims[3], void* stream);
/* Generic stand-alone BC kernel (boundary_condition.py:83-117): cells with bc_mask == bc.id get the BC's functional,
 * all others keep f_post.  `missing` is the reference's bool [q][...] array. f_pre / f_post share `dtype`. */
int xlbn_bc_apply(int lattice, int compute_dtype, const xlbn_bc_desc* bc, const void* f_pre, void* f_post, int dtype,
                  const uint8_t* bc_mask, const uint8_t* missing, const int32_t dims[3], void* stream);

/* MomentumTransfer (operator/force/momentum_transfer.py:51-90, 108-176): net momentum-exchange force on the solid
 * behind the no-slip BC `bc`, summed over its edge cells.  f0 = post-collision populations, f1 = the second buffer
 * (only read for the aux value of Zou-He-type BCs; may be NULL).  force: device double[3], overwritten. */
int xlbn_momentum_transfer(int lattice, int compute_dtype, const xlbn_bc_desc* bc, const void* f0, const void* f1, int dtype,
                           const uint8_t* bc_mask, const uint8_t* missing, const int32_t dims[3], double* force, void* stream);

/* ---- x-slab halo (multi-GPU; one process per GPU) ------------------------------------------------------------- */
/* Replaces the two lax.ppermute collectives of distribute.py:23-44 / parallel_operator.py:68-80.
 * Ghost storage per slab: 2 parities x 2 faces x n_dir x ny x nz store-dtype values (n_dir = 5 / 9 / 3), plus two
 * step counters written by the neighbours.  Allocated with cudaMalloc so that it can be exported with CUDA IPC. */
int xlbn_halo_create(int lattice, int store_dtype, int ny, int nz, xlbn_halo** out);
int xlbn_halo_destroy(xlbn_halo* h);
/* 64-byte CUDA IPC handle of this slab's ghost block, to be sent to both neighbours (any transport). */
int xlbn_halo_export(xlbn_halo* h, unsigned char handle[64]);
/* Map the neighbours' ghost blocks.  lo = rank-1 (ring), hi = rank+1.  `same_process` != 0: the handles are raw
 * device pointers (8 bytes) of halos living in this process (single-process multi-GPU, tests). Synchronous. */
int xlbn_halo_connect(xlbn_halo* h, const unsigned char lo_handle[64], const unsigned char hi_handle[64], int same_process);
/* Fill this slab's OWN outgoing planes for step `timestep` from f (used once before the first step, and by the
 * non-fused fallback): writes f's plane nx-1 (c_x = +1 populations) into the hi neighbour's ghost and plane 0
 * (c_x = -1) into the lo neighbour's ghost, parity timestep & 1. */
int xlbn_halo_push(xlbn_halo* h, const void* f, const xlbn_domain* dom, int timestep, void* stream);
/* Tell both neighbours that this slab's ghost contributions for step `timestep` are complete (stream-ordered). */
int xlbn_halo_signal(xlbn_halo* h, int timestep, void* stream);
/* Make `stream` wait until both neighbours have signalled `timestep` (device-side spin, no host sync). */
int xlbn_halo_wait(xlbn_halo* h, int timestep, void* stream);
/* Dead-neighbour detection.  A device-side wait never hangs the GPU: after the timeout (default 120 s, environment
 * XLBN_HALO_TIMEOUT_S, or this call) it gives up and marks the handle; from then on xlbn_step (with this halo) and every
 * xlbn_halo_push / _signal / _wait return XLBN_E_STATE — the step that timed out read stale ghosts, so its result and
 * everything after it is invalid and the run must stop.  xlbn_halo_timed_out: 1 if marked, 0 if not (no device sync). */
int xlbn_halo_set_timeout(xlbn_halo* h, double seconds);
int xlbn_halo_timed_out(xlbn_halo* h);
/* Raw pointers for diagnostics/tests: ghost plane block [parity][face(0=lo,1=hi)][n_dir][ny][nz]. */
int xlbn_halo_ghost_ptr(xlbn_halo* h, void** ptr, long long* bytes);

#ifdef __cplusplus
}
#endif
#endif /* XLB_B200_H */
