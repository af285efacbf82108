/* TEST INFRASTRUCTURE (oracle) — included twice by lbm_ref.c with REAL = float / double.
 * Per-cell restatement in C of the reference's fused Warp kernel (xlb/operator/stepper/nse_stepper.py:344-381) and the
 * functionals it calls; each block cites the reference lines it follows.  Lattice tables come from the caller
 * (oracle/lbm_numpy.py:Lattice), so this file is lattice-agnostic except for the KBC shear split. */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

static inline REAL FN(load)(const void* p, int store, long long i) {
  switch (store) {
    case 0: return (REAL)((const _Float16*)p)[i];
    case 1: return (REAL)((const float*)p)[i];
    default: return (REAL)((const double*)p)[i];
  }
}
static inline void FN(store)(void* p, int store, long long i, REAL v) {
  switch (store) {
    case 0: ((_Float16*)p)[i] = (_Float16)v; break;
    case 1: ((float*)p)[i] = (float)v; break;
    default: ((double*)p)[i] = (double)v; break;
  }
}

static inline REAL FN(sqrt_)(REAL v) { return sizeof(REAL) == 4 ? (REAL)sqrtf((float)v) : (REAL)sqrt((double)v); }

/* quadratic_equilibrium.py:35-60 */
static void FN(equilibrium)(const LbmDesc* d, REAL rho, const REAL* u, REAL* feq) {
  REAL uu = 0;
  for (int a = 0; a < d->d; ++a) uu += u[a] * u[a];
  const REAL usqr = (REAL)1.5 * uu;
  for (int l = 0; l < d->q; ++l) {
    REAL cu = 0;
    for (int a = 0; a < d->d; ++a) {
      if (d->c[a * 27 + l] == 1) cu += u[a];
      else if (d->c[a * 27 + l] == -1) cu -= u[a];
    }
    cu *= (REAL)3.0;
    feq[l] = rho * (REAL)d->w[l] * ((REAL)1.0 + cu * ((REAL)1.0 + (REAL)0.5 * cu) - usqr);
  }
}

/* zero_moment.py:32-37, first_moment.py:26-38 */
static void FN(macroscopic)(const LbmDesc* d, const REAL* f, REAL* rho, REAL* u) {
  REAL r = 0;
  for (int l = 0; l < d->q; ++l) r += f[l];
  for (int a = 0; a < d->d; ++a) u[a] = 0;
  for (int l = 0; l < d->q; ++l)
    for (int a = 0; a < d->d; ++a) {
      if (d->c[a * 27 + l] == 1) u[a] += f[l];
      else if (d->c[a * 27 + l] == -1) u[a] -= f[l];
    }
  for (int a = 0; a < d->d; ++a) u[a] /= r;
  *rho = r;
}

/* second_moment.py:67-78 */
static void FN(second_moment)(const LbmDesc* d, const REAL* f, REAL* pi) {
  const int nt = d->d * (d->d + 1) / 2;
  for (int t = 0; t < nt; ++t) {
    pi[t] = 0;
    for (int l = 0; l < d->q; ++l) pi[t] += (REAL)d->cc[l * 6 + t] * f[l];
  }
}

/* kbc.py:188-296 */
static void FN(collide_kbc)(const LbmDesc* d, const REAL* f, const REAL* feq, REAL rho, REAL omega, REAL* out) {
  REAL fneq[27], s[27], pi[6];
  for (int l = 0; l < d->q; ++l) {
    fneq[l] = f[l] - feq[l];
    s[l] = 0;
  }
  FN(second_moment)(d, fneq, pi);
  if (d->d == 3) {
    const REAL nxz = pi[0] - pi[5], nyz = pi[3] - pi[5];
    s[9] = s[18] = ((REAL)2.0 * nxz - nyz) / (REAL)6.0;
    s[3] = s[6] = (-nxz + (REAL)2.0 * nyz) / (REAL)6.0;
    s[1] = s[2] = (-nxz - nyz) / (REAL)6.0;
    s[12] = s[24] = pi[1] / (REAL)4.0;
    s[21] = s[15] = -pi[1] / (REAL)4.0;
    s[10] = s[20] = pi[2] / (REAL)4.0;
    s[19] = s[11] = -pi[2] / (REAL)4.0;
    s[8] = s[4] = pi[4] / (REAL)4.0;
    s[7] = s[5] = -pi[4] / (REAL)4.0;
  } else {
    const REAL n = pi[0] - pi[2];
    s[3] = s[6] = n;
    s[2] = s[1] = -n;
    s[8] = s[7] = pi[1];
    s[4] = s[5] = -pi[1];
  }
  for (int l = 0; l < d->q; ++l) s[l] = (d->d == 3) ? s[l] * rho : s[l] * rho / (REAL)4.0; /* delta_s */
  const REAL beta = (REAL)0.5 * omega, inv_beta = (REAL)1.0 / beta;
  REAL sp1 = 0, sp2 = 0;
  for (int l = 0; l < d->q; ++l) {
    const REAL dh = fneq[l] - s[l];
    const REAL temp = dh / feq[l];
    sp1 += temp * s[l];
    sp2 += temp * dh;
  }
  const REAL gamma = inv_beta - ((REAL)2.0 - inv_beta) * sp1 / ((REAL)1e-32 + sp2);
  for (int l = 0; l < d->q; ++l) out[l] = f[l] - beta * ((REAL)2.0 * s[l] + gamma * (fneq[l] - s[l]));
}

/* helper_functions_bc.py:75-86 / bc_extrapolation_outflow.py:139-149: first missing axis-aligned direction */
static void FN(normal)(const LbmDesc* d, const unsigned char* miss, int* n) {
  n[0] = n[1] = n[2] = 0;
  for (int l = 0; l < d->q; ++l) {
    const int sp = abs(d->c[l]) + abs(d->c[27 + l]) + abs(d->c[54 + l]);
    if (miss[l] && sp == 1) {
      for (int a = 0; a < 3; ++a) n[a] = -d->c[a * 27 + l];
      return;
    }
  }
}

/* bc_zouhe.py:279-338, bc_regularized.py:134-202, helper_functions_bc.py:61-122 */
static void FN(zouhe)(const LbmDesc* d, int kind, REAL aux, const unsigned char* miss, REAL* f) {
  int ni[3];
  FN(normal)(d, miss, ni);
  REAL nrm[3] = {(REAL)ni[0], (REAL)ni[1], (REAL)ni[2]}, u[3] = {0, 0, 0}, feq[27];
  REAL known = 0, middle = 0;
  for (int l = 0; l < d->q; ++l) {
    if (miss[d->opp[l]]) known += (REAL)2.0 * f[l];
    else if (!miss[l]) middle += f[l];
  }
  const REAL fsum = known + middle;
  REAL rho;
  if (kind == BC_ZOUHE_VELOCITY || kind == BC_REGULARIZED_VELOCITY) {
    REAL unormal = 0;
    for (int a = 0; a < d->d; ++a) u[a] = -aux * nrm[a];
    for (int a = 0; a < d->d; ++a) unormal += u[a] * nrm[a];
    rho = fsum / ((REAL)1.0 + unormal);
  } else {
    rho = aux;
    const REAL unormal = -(REAL)1.0 + fsum / rho;
    for (int a = 0; a < d->d; ++a) u[a] = unormal * nrm[a];
  }
  FN(equilibrium)(d, rho, u, feq);
  for (int l = 0; l < d->q; ++l)
    if (miss[l]) f[l] = f[d->opp[l]] + feq[l] - feq[d->opp[l]];
  if (kind == BC_REGULARIZED_VELOCITY || kind == BC_REGULARIZED_PRESSURE) {
    REAL fneq[27], pi[6];
    const int nt = d->d * (d->d + 1) / 2;
    for (int l = 0; l < d->q; ++l) fneq[l] = f[l] - feq[l];
    FN(second_moment)(d, fneq, pi);
    for (int l = 0; l < d->q; ++l) {
      REAL qipi = 0;
      for (int t = 0; t < nt; ++t) qipi += (REAL)d->qi[l * 6 + t] * pi[t];
      f[l] = feq[l] + (REAL)4.5 * (REAL)d->w[l] * qipi;
    }
  }
}

static inline long long FN(cell_of)(const LbmDesc* d, int x, int y, int z) { return ((long long)x * d->ny + y) * d->nz + z; }

/* one step over the whole grid: nse_stepper.py:344-381 */
static void FN(step)(const LbmDesc* d, void* f0, void* f1, const unsigned char* bc_mask, const unsigned char* missing, int nthreads) {
  const long long n = (long long)d->nx * d->ny * d->nz;
  const REAL omega = (REAL)d->omega;
  const REAL cs = (REAL)(1.0 / sqrt(3.0));
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads)
  for (int x = 0; x < d->nx; ++x)
    for (int y = 0; y < d->ny; ++y)
      for (int z = 0; z < d->nz; ++z) {
        const long long cell = FN(cell_of)(d, x, y, z);
        const int id = bc_mask[cell];
        if (id == 255) continue; /* L356-358 */
        REAL f[27], fpre[27], out[27];
        unsigned char miss[27];
        /* pull with periodic wrap: stream.py:59-82 */
        for (int l = 0; l < d->q; ++l) {
          int xs = x - d->c[l], ys = y - d->c[27 + l], zs = z - d->c[54 + l];
          xs = xs < 0 ? d->nx - 1 : (xs >= d->nx ? 0 : xs);
          ys = ys < 0 ? d->ny - 1 : (ys >= d->ny ? 0 : ys);
          zs = zs < 0 ? d->nz - 1 : (zs >= d->nz ? 0 : zs);
          f[l] = FN(load)(f0, d->store, (long long)l * n + FN(cell_of)(d, xs, ys, zs));
        }
        const int kind = id ? d->bc_kind[id] : BC_NONE;
        if (kind != BC_NONE) { /* thread data: L296-316 */
          for (int l = 0; l < d->q; ++l) {
            fpre[l] = FN(load)(f0, d->store, (long long)l * n + cell);
            miss[l] = missing[(long long)l * n + cell];
          }
        }
        /* streaming-step BCs: L367 */
        REAL aux_raw = 0;
        switch (kind) {
          case BC_EQUILIBRIUM: { /* bc_equilibrium.py:76-86 */
            REAL u[3] = {(REAL)d->bc_u[id * 3], (REAL)d->bc_u[id * 3 + 1], (REAL)d->bc_u[id * 3 + 2]};
            FN(equilibrium)(d, (REAL)d->bc_rho[id], u, f);
          } break;
          case BC_DO_NOTHING: /* bc_do_nothing.py:52-63 */
            for (int l = 0; l < d->q; ++l) f[l] = fpre[l];
            break;
          case BC_HALFWAY: /* bc_halfway_bounce_back.py:68-85 */
          case BC_OUTFLOW: /* bc_extrapolation_outflow.py:152-170 */
            for (int l = 0; l < d->q; ++l)
              if (miss[l]) f[l] = fpre[d->opp[l]];
            break;
          case BC_ZOUHE_VELOCITY:
          case BC_ZOUHE_PRESSURE:
          case BC_REGULARIZED_VELOCITY:
          case BC_REGULARIZED_PRESSURE:
            aux_raw = FN(load)(f1, d->store, cell); /* prescribed value in f_1[0, cell]: bc_zouhe.py:302 */
            FN(zouhe)(d, kind, aux_raw, miss, f);
            break;
          default: break;
        }
        /* macroscopic -> equilibrium -> collision: L369-371 */
        REAL rho, u[3] = {0, 0, 0}, feq[27];
        FN(macroscopic)(d, f, &rho, u);
        FN(equilibrium)(d, rho, u, feq);
        if (d->collision == 0) {
          for (int l = 0; l < d->q; ++l) out[l] = f[l] - omega * (f[l] - feq[l]); /* bgk.py:30-34 */
        } else if (d->collision == 1) {
          FN(collide_kbc)(d, f, feq, rho, omega, out);
        } else { /* smagorinsky_les_bgk.py:37-90: 'strain' from squared fneq selected by the SIGNED sum of c_l */
          REAL strain = 0;
          for (int l = 0; l < d->q; ++l) {
            const int cs_l = d->c[l] + d->c[27 + l] + d->c[54 + l];
            const REAL fneq = f[l] - feq[l];
            if (cs_l == 1) strain += fneq * fneq;
            if (cs_l >= 2) strain += (REAL)2.0 * fneq * fneq;
          }
          const REAL tau0 = (REAL)1.0 / omega, coef = (REAL)d->smagorinsky;
          const REAL tau = tau0 + (REAL)0.5 * (FN(sqrt_)(tau0 * tau0 + (REAL)36.0 * (coef * coef) * FN(sqrt_)(strain)) - tau0);
          const REAL inv_tau = (REAL)1.0 / tau;
          for (int l = 0; l < d->q; ++l) out[l] = f[l] - inv_tau * (f[l] - feq[l]);
        }
        if (d->has_force) { /* exact_difference_force.py:79-84: f += feq(rho, u + F) - feq(rho, u) */
          REAL uf[3] = {0, 0, 0}, feq_force[27];
          for (int a = 0; a < d->d; ++a) uf[a] = u[a] + (REAL)d->force[a];
          FN(equilibrium)(d, rho, uf, feq_force);
          for (int l = 0; l < d->q; ++l) out[l] += feq_force[l] - feq[l];
        }
        /* collision-step BCs and outflow aux: L374, L286-293 */
        if (kind == BC_FULLWAY) { /* bc_fullway_bounce_back.py:60-72 */
          for (int l = 0; l < d->q; ++l) out[l] = f[d->opp[l]];
        } else if (kind == BC_OUTFLOW) { /* bc_extrapolation_outflow.py:172-195 */
          int nv[3];
          FN(normal)(d, miss, nv);
          for (int l = 0; l < d->q; ++l)
            if (miss[l]) {
              const int px = x - (d->c[l] + nv[0]), py = y - (d->c[27 + l] + nv[1]), pz = z - (d->c[54 + l] + nv[2]);
              const REAL f_aux = FN(load)(f0, d->store, (long long)l * n + FN(cell_of)(d, px, py, pz));
              out[d->opp[l]] = ((REAL)1.0 - cs) * f[l] + cs * f_aux;
            }
        }
        /* aux recovery: L318-342 (only the rest slot carries data for the in-scope BCs) */
        if (kind >= BC_ZOUHE_VELOCITY && kind <= BC_REGULARIZED_PRESSURE) FN(store)(f0, d->store, cell, aux_raw);
        for (int l = 0; l < d->q; ++l) FN(store)(f1, d->store, (long long)l * n + cell, out[l]); /* L380-381 */
      }
}

#undef FN
#undef CAT
#undef CAT_
