// Per-cell LBM algebra in registers, in the compute dtype TC.  One set of device functions shared by the fused step
// kernel, the boundary-cell path and the stand-alone operator kernels, so all of them produce the same numbers.
// The operation ORDER follows the reference's Warp functionals (sequential in l), cited per function.
#pragma once

#include "lattice.cuh"

namespace xlbn {

#define XLBN_DEV XLBN_MATH
#define XLBN_FOR(N, var) static_for<N>([&](auto var##_) { constexpr int var = decltype(var##_)::value;
#define XLBN_END });

// rho = sum_l f_l ; u_d = (sum_{c=+1} f - sum_{c=-1} f) / rho
// (reference: zero_moment.py:32-37, first_moment.py:26-38, macroscopic.py:43-47)
// FAST: u = (sum c f) * (1/rho) instead of three divisions (<= 1 ulp difference); used where the kernel is issue-bound
template <class L, class TC, bool FAST = false>
XLBN_DEV void macroscopic(const TC (&f)[L::Q], TC& rho, TC (&u)[L::D]) {
  rho = TC(0);
  XLBN_FOR(L::Q, l) rho += f[l]; XLBN_END
  XLBN_FOR(L::D, d) u[d] = TC(0); XLBN_END
  XLBN_FOR(L::Q, l)
    XLBN_FOR(L::D, d)
      if constexpr (L::c(d, l) == 1) u[d] += f[l];
      else if constexpr (L::c(d, l) == -1) u[d] -= f[l];
    XLBN_END
  XLBN_END
  if constexpr (FAST) {
    const TC inv = rcp_(rho);
    XLBN_FOR(L::D, d) u[d] = u[d] * inv; XLBN_END
  } else {
    XLBN_FOR(L::D, d) u[d] /= rho; XLBN_END
  }
}

// feq_l = rho w_l (1 + cu (1 + 0.5 cu) - usqr), cu = 3 c_l.u, usqr = 1.5 u.u
// (reference: quadratic_equilibrium.py:35-60)
// ROUNDINGS.  The library is compiled with -fmad=false: the compiler never contracts a*b+c, every fused operation is an explicit
// fma_() in the source.  The BGK chain (macroscopic -> equilibrium -> BGK, and the BC functionals) uses NONE: one IEEE rounding per
// operation, in the reference's operation order — exactly what the reference's functionals evaluate — so that fp32 results are
// bit-identical to the reference kernel as restated by oracle/lbm_ref.c, and fp16 storage sees the same roundings step after step.
template <class L, class TC>
XLBN_DEV void equilibrium(TC rho, const TC (&u)[L::D], TC (&feq)[L::Q]) {
  TC uu = u[0] * u[0];
  XLBN_FOR(L::D - 1, d) uu = uu + u[d + 1] * u[d + 1]; XLBN_END
  const TC usqr = TC(1.5) * uu;
  XLBN_FOR(L::Q, l)
    TC cu = TC(0);
    XLBN_FOR(L::D, d)
      if constexpr (L::c(d, l) == 1) cu += u[d];
      else if constexpr (L::c(d, l) == -1) cu -= u[d];
    XLBN_END
    cu *= TC(3.0);
    feq[l] = rho * TC(L::w(l)) * (TC(1.0) + cu * (TC(1.0) + TC(0.5) * cu) - usqr);  // quadratic_equilibrium.py:58, one rounding per operation
  XLBN_END
}

// Pi_t = sum_l cc[l][t] f_l   (reference: second_moment.py:67-78)
template <class L, class TC>
XLBN_DEV void second_moment(const TC (&f)[L::Q], TC (&pi)[L::NT]) {
  XLBN_FOR(L::NT, t)
    pi[t] = TC(0);
    XLBN_FOR(L::Q, l)
      if constexpr (L::cc(l, t) == 1) pi[t] += f[l];
      else if constexpr (L::cc(l, t) == -1) pi[t] -= f[l];
    XLBN_END
  XLBN_END
}

// BGK: f - omega (f - feq)   (reference: bgk.py:30-34)
template <class L, class TC>
XLBN_DEV void collide_bgk(const TC (&f)[L::Q], const TC (&feq)[L::Q], TC omega, TC (&out)[L::Q]) {
  XLBN_FOR(L::Q, l)
    const TC fneq = f[l] - feq[l];
    out[l] = f[l] - omega * fneq;  // bgk.py:33, unfused
  XLBN_END
}

// KBC shear part of fneq (reference: kbc.py:188-250).  s_l for the D3Q27 / D2Q9 lattices.
template <class L, class TC, bool FAST = false>
XLBN_DEV void kbc_shear(const TC (&fneq)[L::Q], TC (&s)[L::Q]) {
  TC pi[L::NT];
  second_moment<L, TC>(fneq, pi);
  XLBN_FOR(L::Q, l) s[l] = TC(0); XLBN_END
  if constexpr (L::ID == XLBN_D3Q27) {
    const TC nxz = pi[0] - pi[5];
    const TC nyz = pi[3] - pi[5];
    if constexpr (FAST) {  // x / 6 as x * (1/6): one rounding of the constant instead of a division
      s[9] = (TC(2.0) * nxz - nyz) * TC(1.0 / 6.0);
      s[3] = (-nxz + TC(2.0) * nyz) * TC(1.0 / 6.0);
      s[1] = (-nxz - nyz) * TC(1.0 / 6.0);
    } else {
      s[9] = (TC(2.0) * nxz - nyz) / TC(6.0);
      s[3] = (-nxz + TC(2.0) * nyz) / TC(6.0);
      s[1] = (-nxz - nyz) / TC(6.0);
    }
    s[18] = s[9];
    s[6] = s[3];
    s[2] = s[1];
    s[12] = pi[1] / TC(4.0);
    s[24] = s[12];
    s[21] = -pi[1] / TC(4.0);
    s[15] = s[21];
    s[10] = pi[2] / TC(4.0);
    s[20] = s[10];
    s[19] = -pi[2] / TC(4.0);
    s[11] = s[19];
    s[8] = pi[4] / TC(4.0);
    s[4] = s[8];
    s[7] = -pi[4] / TC(4.0);
    s[5] = s[7];
  } else if constexpr (L::ID == XLBN_D2Q9) {
    const TC n = pi[0] - pi[2];
    s[3] = n;
    s[6] = n;
    s[2] = -n;
    s[1] = -n;
    s[8] = pi[1];
    s[4] = -pi[1];
    s[5] = -pi[1];
    s[7] = pi[1];
  }
}

// KBC (reference: kbc.py:268-296): entropic stabiliser gamma from the scalar products <dh|ds>, <dh|dh> weighted by 1/feq.
// FAST: the 27 divisions dh/feq (they only feed the scalar stabiliser gamma) become dh * rcp(feq)
template <class L, class TC, bool FAST = false>
XLBN_DEV void collide_kbc(const TC (&f)[L::Q], const TC (&feq)[L::Q], TC rho, TC omega, TC (&out)[L::Q]) {
  TC fneq[L::Q], ds[L::Q];
  XLBN_FOR(L::Q, l) fneq[l] = f[l] - feq[l]; XLBN_END
  kbc_shear<L, TC, FAST>(fneq, ds);
  XLBN_FOR(L::Q, l)
    if constexpr (L::D == 3) ds[l] = ds[l] * rho;
    else ds[l] = ds[l] * rho / TC(4.0);
  XLBN_END
  const TC beta = TC(0.5) * omega;
  const TC inv_beta = TC(1.0) / beta;
  TC sp1 = TC(0), sp2 = TC(0);
  XLBN_FOR(L::Q, l)
    const TC dh = fneq[l] - ds[l];
    if constexpr (FAST) {
      const TC temp = dh * rcp_approx_(feq[l]);
      sp1 = fma_(temp, ds[l], sp1);
      sp2 = fma_(temp, dh, sp2);
    } else {  // the reference's roundings, one per operation (kbc.py:284-291)
      const TC temp = dh / feq[l];
      sp1 = sp1 + temp * ds[l];
      sp2 = sp2 + temp * dh;
    }
  XLBN_END
  const TC gamma = inv_beta - (TC(2.0) - inv_beta) * sp1 / (TC(1e-32) + sp2);
  XLBN_FOR(L::Q, l)
    const TC dh = fneq[l] - ds[l];
    if constexpr (FAST) out[l] = fma_(-beta, fma_(gamma, dh, TC(2.0) * ds[l]), f[l]);
    else out[l] = f[l] - beta * (TC(2.0) * ds[l] + gamma * dh);  // kbc.py:294-296
  XLBN_END
}

// macroscopic -> equilibrium -> collision on one cell, in place (reference: nse_stepper.py:369-371).
// FAST selects reciprocal-based divisions (see macroscopic / collide_kbc); the fused kernel uses it where it is
// issue-bound (KBC, and the packed pair path), never for fp64.
template <class L, int COLL, class TC, bool FAST = false>
XLBN_DEV void collide_cell(TC (&f)[L::Q], TC omega) {
  TC rho, u[L::D], feq[L::Q], out[L::Q];
  macroscopic<L, TC, FAST>(f, rho, u);
  equilibrium<L, TC>(rho, u, feq);
  if constexpr (COLL == XLBN_BGK) collide_bgk<L, TC>(f, feq, omega, out);
  else collide_kbc<L, TC, FAST>(f, feq, rho, omega, out);
  XLBN_FOR(L::Q, l) f[l] = out[l]; XLBN_END
}

// ---- extended collision models ------------------------------------------------------------------
// COLL = base operator (low two bits) | XLBN_COLLISION_FORCED.
template <int COLL>
constexpr int kBaseCollision = COLL & 3;
template <int COLL>
constexpr bool kForcedCollision = (COLL & XLBN_COLLISION_FORCED) != 0;

XLBN_MATH float sqrt_(float x) { return sqrtf(x); }
XLBN_MATH double sqrt_(double x) { return sqrt(x); }

// SmagorinskyLESBGK (reference: smagorinsky_les_bgk.py:37-90; a Warp functional only, which reads c[2, l]: 3-D lattices).
// Restated literally: the 'strain' is a sum of squared non-equilibrium populations selected by the SIGNED component
// sum of c_l (== 1: weight 1, >= 2: weight 2), in l order; tau = tau0 + (sqrt(tau0^2 + 36 C^2 sqrt(strain)) - tau0)/2.
template <class L, class TC>
XLBN_DEV void collide_smagorinsky(const TC (&f)[L::Q], const TC (&feq)[L::Q], TC omega, TC coef, TC (&out)[L::Q]) {
  static_assert(L::D == 3, "SmagorinskyLESBGK: 3-D lattices only (the reference functional indexes c[2, l])");
  TC strain = TC(0);
  XLBN_FOR(L::Q, l)
    constexpr int csum = L::c(0, l) + L::c(1, l) + L::c(2, l);
    if constexpr (csum >= 1) {
      const TC fneq = f[l] - feq[l];
      if constexpr (csum == 1) strain = fma_(fneq, fneq, strain);
      else strain = fma_(TC(2.0) * fneq, fneq, strain);
    }
  XLBN_END
  const TC tau0 = TC(1.0) / omega;
  const TC tau = tau0 + TC(0.5) * (sqrt_(fma_(tau0, tau0, TC(36.0) * (coef * coef) * sqrt_(strain))) - tau0);
  const TC inv_tau = TC(1.0) / tau;
  XLBN_FOR(L::Q, l) out[l] = fma_(-inv_tau, f[l] - feq[l], f[l]); XLBN_END
}

// ExactDifference forcing (reference: exact_difference_force.py:79-84): out += feq(rho, u + F) - feq(rho, u)
template <class L, class TC>
XLBN_DEV void exact_difference(TC rho, const TC (&u)[L::D], const TC (&feq)[L::Q], const TC (&force)[L::D], TC (&out)[L::Q]) {
  TC uf[L::D], feq_force[L::Q];
  XLBN_FOR(L::D, d) uf[d] = u[d] + force[d]; XLBN_END
  equilibrium<L, TC>(rho, uf, feq_force);
  XLBN_FOR(L::Q, l) out[l] += feq_force[l] - feq[l]; XLBN_END
}

// ---- KBC, register-lean formulation (tuning variant, opt-in: cells_per_thread = 301) ---------------------------------------
// Same algebra as collide_kbc with FAST divisions, arranged as three passes over the populations that RECOMPUTE feq_l
// instead of holding feq[], fneq[], ds[] (3 q-vectors) next to f[]: six distinct shear values instead of a q-vector, the
// shear-free populations (rest + 8 corners on D3Q27) skipped, and an opaque launder() on (rho, u, usqr) between the passes so
// that the compiler cannot CSE the recomputation back into q live values.  ptxas, D3Q27 fp32: 72 registers / 32 B spilled
// against 80 / 1 KiB for collide_kbc, for +9 % instructions (profiles/round2_prep/).  Differs from collide_kbc by rounding
// only (x/4 -> x*0.25 is exact; the output is f - beta*gamma*fneq - beta*(2-gamma)*ds instead of f - beta*(2 ds + gamma dh)).
constexpr int kLeanKbc = 8;  // internal flag or-ed onto XLBN_KBC in the COLL template argument
// cells_per_thread = 300: the literal KBC formulation with the reference's roundings (IEEE divisions, nothing fused): bit-identical to
// the reference kernel, ~35 divisions per cell (the parity form; the lean form above is the fast default)
constexpr int kExactKbc = 16;

XLBN_DEV void launder(float& x) {
#if !XLBN_ON_HOST
  asm volatile("" : "+f"(x));
#endif
}
XLBN_DEV void launder(double& x) {
#if !XLBN_ON_HOST
  asm volatile("" : "+d"(x));
#endif
}

// one population's equilibrium from (rho, u, usqr)   (quadratic_equilibrium.py:35-60, same expression as equilibrium())
template <class L, class TC, int l>
XLBN_DEV TC feq_one(TC rho, const TC (&u)[L::D], TC usqr) {
  TC cu = TC(0);
  XLBN_FOR(L::D, d)
    if constexpr (L::c(d, l) == 1) cu += u[d];
    else if constexpr (L::c(d, l) == -1) cu -= u[d];
  XLBN_END
  cu *= TC(3.0);
  return rho * TC(L::w(l)) * (fma_(cu, fma_(TC(0.5), cu, TC(1.0)), TC(1.0)) - usqr);
}

// delta_s of population l from the distinct shear values (already multiplied by rho [/4 in 2-D]); kbc.py:213-250, SURVEY Appendix B
template <class L, class TC, int l>
XLBN_DEV TC kbc_ds(const TC (&sv)[6]) {
  if constexpr (L::ID == XLBN_D3Q27) {
    if constexpr (l == 9 || l == 18) return sv[0];
    else if constexpr (l == 3 || l == 6) return sv[1];
    else if constexpr (l == 1 || l == 2) return sv[2];
    else if constexpr (l == 12 || l == 24) return sv[3];
    else if constexpr (l == 21 || l == 15) return -sv[3];
    else if constexpr (l == 10 || l == 20) return sv[4];
    else if constexpr (l == 19 || l == 11) return -sv[4];
    else if constexpr (l == 8 || l == 4) return sv[5];
    else if constexpr (l == 7 || l == 5) return -sv[5];
    else return TC(0);
  } else {
    if constexpr (l == 3 || l == 6) return sv[0];
    else if constexpr (l == 1 || l == 2) return -sv[0];
    else if constexpr (l == 7 || l == 8) return sv[1];
    else if constexpr (l == 4 || l == 5) return -sv[1];
    else return TC(0);
  }
}
template <class L, int l>
__host__ __device__ constexpr bool kbc_has_shear() {
  if constexpr (L::ID == XLBN_D3Q27) return l != 0 && L::speed(l) != 3;
  else return l != 0;
}

// FORCED: ForcedCollision's ExactDifference term (exact_difference_force.py:79-84), + feq(rho, u + F) - feq(rho, u), is added in pass 3,
// where feq_l(rho, u) is at hand anyway: one more feq_one per population instead of a second q-vector.
template <class L, class TC, bool FAST, bool FORCED = false>
XLBN_DEV void collide_kbc_lean(TC (&f)[L::Q], TC omega, const TC* force = nullptr) {
  static_assert(L::ID == XLBN_D3Q27 || L::ID == XLBN_D2Q9, "KBC: D3Q27 and D2Q9 only (kbc.py:71-72)");
  TC rho, u[L::D];
  macroscopic<L, TC, FAST>(f, rho, u);
  TC uu = u[0] * u[0];
  XLBN_FOR(L::D - 1, d) uu = fma_(u[d + 1], u[d + 1], uu); XLBN_END
  TC usqr = TC(1.5) * uu;
  // pass 1: Pi_neq
  TC pi[L::NT];
  XLBN_FOR(L::NT, t) pi[t] = TC(0); XLBN_END
  XLBN_FOR(L::Q, l)
    const TC fneq = f[l] - feq_one<L, TC, l>(rho, u, usqr);
    XLBN_FOR(L::NT, t)
      if constexpr (L::cc(l, t) == 1) pi[t] += fneq;
      else if constexpr (L::cc(l, t) == -1) pi[t] -= fneq;
    XLBN_END
  XLBN_END
  TC sv[6];
  if constexpr (L::ID == XLBN_D3Q27) {
    const TC nxz = pi[0] - pi[5], nyz = pi[3] - pi[5];
    sv[0] = (TC(2.0) * nxz - nyz) * TC(1.0 / 6.0) * rho;
    sv[1] = (-nxz + TC(2.0) * nyz) * TC(1.0 / 6.0) * rho;
    sv[2] = (-nxz - nyz) * TC(1.0 / 6.0) * rho;
    sv[3] = pi[1] * TC(0.25) * rho;
    sv[4] = pi[2] * TC(0.25) * rho;
    sv[5] = pi[4] * TC(0.25) * rho;
  } else {
    sv[0] = (pi[0] - pi[2]) * rho * TC(0.25);
    sv[1] = pi[1] * rho * TC(0.25);
    sv[2] = sv[3] = sv[4] = sv[5] = TC(0);
  }
  launder(rho);
  launder(usqr);
  XLBN_FOR(L::D, d) launder(u[d]); XLBN_END
  // pass 2: entropic scalar products
  TC sp1 = TC(0), sp2 = TC(0);
  XLBN_FOR(L::Q, l)
    const TC feq = feq_one<L, TC, l>(rho, u, usqr);
    const TC fneq = f[l] - feq;
    const TC r = rcp_approx_(feq);
    if constexpr (kbc_has_shear<L, l>()) {
      const TC ds = kbc_ds<L, TC, l>(sv);
      const TC dh = fneq - ds;
      const TC temp = dh * r;
      sp1 = fma_(temp, ds, sp1);
      sp2 = fma_(temp, dh, sp2);
    } else {
      sp2 = fma_(fneq * r, fneq, sp2);
    }
  XLBN_END
  const TC beta = TC(0.5) * omega;
  const TC inv_beta = TC(1.0) / beta;
  const TC gamma = inv_beta - (TC(2.0) - inv_beta) * sp1 / (TC(1e-32) + sp2);
  // pass 3: f - beta (2 ds + gamma dh) = f - beta gamma fneq - beta (2 - gamma) ds
  launder(rho);
  launder(usqr);
  XLBN_FOR(L::D, d) launder(u[d]); XLBN_END
  const TC bg = beta * gamma, b2 = beta * (TC(2.0) - gamma);
  TC uf[L::D], usqr_f = TC(0);
  if constexpr (FORCED) {
    XLBN_FOR(L::D, d) uf[d] = u[d] + force[d]; XLBN_END
    TC uuf = uf[0] * uf[0];
    XLBN_FOR(L::D - 1, d) uuf = fma_(uf[d + 1], uf[d + 1], uuf); XLBN_END
    usqr_f = TC(1.5) * uuf;
  }
  XLBN_FOR(L::Q, l)
    const TC feq = feq_one<L, TC, l>(rho, u, usqr);
    const TC fneq = f[l] - feq;
    TC out = fma_(-bg, fneq, f[l]);
    if constexpr (kbc_has_shear<L, l>()) out = fma_(-b2, kbc_ds<L, TC, l>(sv), out);
    if constexpr (FORCED) out += feq_one<L, TC, l>(rho, uf, usqr_f) - feq;
    f[l] = out;
  XLBN_END
}

// macroscopic -> equilibrium -> collision [-> forcing] on one cell, in place, for every operator incl. the extended ones
// (reference: nse_stepper.py:369-371 with self.collision = ForcedCollision(...), L45-46).
template <class L, int COLL, class TC, bool FAST = false>
XLBN_DEV void collide_cell_ext(TC (&f)[L::Q], TC omega, const double* force, double smagorinsky) {
  if constexpr ((COLL & kLeanKbc) != 0) {
    static_assert(kBaseCollision<COLL> == XLBN_KBC, "the lean formulation exists for KBC only");
    if constexpr (kForcedCollision<COLL>) {
      TC fv[L::D];
      XLBN_FOR(L::D, d) fv[d] = (TC)force[d]; XLBN_END
      collide_kbc_lean<L, TC, FAST, true>(f, omega, fv);
    } else {
      collide_kbc_lean<L, TC, FAST>(f, omega);
    }
    return;
  }
  TC rho, u[L::D], feq[L::Q], out[L::Q];
  macroscopic<L, TC, FAST>(f, rho, u);
  equilibrium<L, TC>(rho, u, feq);
  if constexpr (kBaseCollision<COLL> == XLBN_BGK) collide_bgk<L, TC>(f, feq, omega, out);
  else if constexpr (kBaseCollision<COLL> == XLBN_KBC) collide_kbc<L, TC, FAST>(f, feq, rho, omega, out);
  else collide_smagorinsky<L, TC>(f, feq, omega, (TC)smagorinsky, out);
  if constexpr (kForcedCollision<COLL>) {
    TC fv[L::D];
    XLBN_FOR(L::D, d) fv[d] = (TC)force[d]; XLBN_END
    exact_difference<L, TC>(rho, u, feq, fv, out);
  }
  XLBN_FOR(L::Q, l) f[l] = out[l]; XLBN_END
}

// ---- boundary-condition functionals -------------------------------------------------------------------------------
// `miss` bit l <=> missing_mask[l, cell].

// First missing axis-aligned direction in index order gives the outward normal n = -c_l
// (reference: helper_functions_bc.py:75-86, bc_extrapolation_outflow.py:139-149).
template <class L>
XLBN_DEV void bc_normal(uint32_t miss, int (&n)[L::D]) {
  bool found = false;
  XLBN_FOR(L::D, d) n[d] = 0; XLBN_END
  XLBN_FOR(L::Q, l)
    if constexpr (L::is_main(l)) {
      if (!found && ((miss >> l) & 1u)) {
        found = true;
        XLBN_FOR(L::D, d) n[d] = -L::c(d, l); XLBN_END
      }
    }
  XLBN_END
}

// fsum = 2 sum_known f + sum_middle f; known: missing[opp[l]], middle: neither (reference: helper_functions_bc.py:61-73)
template <class L, class TC>
XLBN_DEV TC bc_fsum(const TC (&f)[L::Q], uint32_t miss) {
  TC known = TC(0), middle = TC(0);
  XLBN_FOR(L::Q, l)
    if ((miss >> L::opp(l)) & 1u) known += TC(2.0) * f[l];
    else if (!((miss >> l) & 1u)) middle += f[l];
  XLBN_END
  return known + middle;
}

// missing l (sequential, in place): f[l] = f[opp[l]] + feq[l] - feq[opp[l]]  (reference: helper_functions_bc.py:88-97)
template <class L, class TC>
XLBN_DEV void bc_bounceback_nonequilibrium(TC (&f)[L::Q], const TC (&feq)[L::Q], uint32_t miss) {
  XLBN_FOR(L::Q, l)
    if ((miss >> l) & 1u) f[l] = f[L::opp(l)] + feq[l] - feq[L::opp(l)];
  XLBN_END
}

// all l: f[l] = feq[l] + 4.5 w_l (Q_l : Pi_neq)   (reference: helper_functions_bc.py:99-122)
template <class L, class TC>
XLBN_DEV void bc_regularize(TC (&f)[L::Q], const TC (&feq)[L::Q]) {
  TC fneq[L::Q], pi[L::NT];
  XLBN_FOR(L::Q, l) fneq[l] = f[l] - feq[l]; XLBN_END
  second_moment<L, TC>(fneq, pi);
  XLBN_FOR(L::Q, l)
    TC qipi = TC(0);
    XLBN_FOR(L::NT, t) qipi += TC(L::qi(l, t)) * pi[t]; XLBN_END
    f[l] = feq[l] + TC(4.5) * TC(L::w(l)) * qipi;
  XLBN_END
}

// Zou-He / Regularized on the post-stream populations (reference: bc_zouhe.py:279-338, bc_regularized.py:134-202).
// `aux` is the per-cell prescribed scalar: normal velocity magnitude (velocity type) or density (pressure type).
template <class L, class TC>
XLBN_DEV void bc_zouhe(int kind, TC aux, uint32_t miss, TC (&f)[L::Q]) {
  int ni[L::D];
  bc_normal<L>(miss, ni);
  TC nrm[L::D], u[L::D], feq[L::Q];
  XLBN_FOR(L::D, d) nrm[d] = TC(ni[d]); XLBN_END
  const TC fsum = bc_fsum<L, TC>(f, miss);
  TC rho;
  if (kind == XLBN_BC_ZOUHE_VELOCITY || kind == XLBN_BC_REGULARIZED_VELOCITY) {
    TC unormal = TC(0);
    XLBN_FOR(L::D, d) u[d] = -aux * nrm[d]; XLBN_END
    XLBN_FOR(L::D, d) unormal += u[d] * nrm[d]; XLBN_END
    rho = fsum / (TC(1.0) + unormal);
  } else {
    rho = aux;
    const TC unormal = -TC(1.0) + fsum / rho;
    XLBN_FOR(L::D, d) u[d] = unormal * nrm[d]; XLBN_END
  }
  equilibrium<L, TC>(rho, u, feq);
  bc_bounceback_nonequilibrium<L, TC>(f, feq, miss);
  if (kind == XLBN_BC_REGULARIZED_VELOCITY || kind == XLBN_BC_REGULARIZED_PRESSURE) bc_regularize<L, TC>(f, feq);
}

// missing l: f[l] = f_pre[opp[l]]   (reference: bc_halfway_bounce_back.py:68-85, bc_extrapolation_outflow.py:152-170)
template <class L, class TC>
XLBN_DEV void bc_take_opposite_of_pre(const TC (&fpre)[L::Q], uint32_t miss, TC (&f)[L::Q]) {
  XLBN_FOR(L::Q, l)
    if ((miss >> l) & 1u) f[l] = fpre[L::opp(l)];
  XLBN_END
}

inline __host__ __device__ bool bc_kind_needs_aux(int kind) { return kind >= XLBN_BC_ZOUHE_VELOCITY && kind <= XLBN_BC_REGULARIZED_PRESSURE; }
inline __host__ __device__ bool bc_kind_needs_fpre(int kind) {
  return kind == XLBN_BC_DO_NOTHING || kind == XLBN_BC_HALFWAY_BOUNCE_BACK || kind == XLBN_BC_EXTRAPOLATION_OUTFLOW;
}
inline __host__ __device__ bool bc_kind_needs_missing(int kind) {
  return kind == XLBN_BC_HALFWAY_BOUNCE_BACK || kind >= XLBN_BC_ZOUHE_VELOCITY;
}

}  // namespace xlbn
