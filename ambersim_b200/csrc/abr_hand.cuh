// abr_hand.cuh — the path-decomposed world step for FIXED-BASE trees with joint equalities ("hand" kernels).
//
// The reference's own bundled robot and its own sampler test (ambersim/models/barrett_hand/bh280.xml;
// tests/trajopt/test_predictive_sampler.py:17-52: bh280, 100 samples x horizon 10, contacts disabled) is a palm welded
// to the world with three finger chains and four joint-coupling equalities (bh280.xml:196-199), one of which ties two
// different fingers together. The limb kernels (abr_limb.cuh) need a floating base and no equalities, so this model ran
// on the generic shared-memory kernels. Here one LANE owns one root-to-leaf chain of at most NL hinge / slide joints
// hanging off a static body; chains share no moving body, so the joint-space inertia is block diagonal: each lane keeps
// its own NL x NL triangle in registers and the smooth dynamics need no cross-lane traffic at all.
//
// Equality rows couple dofs of one or of two lanes. They are kept as GLOBAL rows, replicated in every lane of the
// world: a lane holds its slice of each row's Jacobian (Je[r][d], zero where the row does not touch the lane), row
// scalars (position, J x, force) are summed over the lanes with a butterfly, and the Newton system
//   H = blockdiag(B_lane) + sum_r D_r u_r u_r',   B_lane = M_lane + limit rows (diagonal)
// is gathered into every lane (order 4 NL <= 12) and factored densely there, like MJX's dense Cholesky of H
// (SURVEY App. A.8); the rest of the solver (costs, line search) works on the lanes' own slices.
//
// Formulas follow SURVEY.md Appendix A (mjx smooth / constraint / solver / forward.euler); spatial quantities are taken
// about the world origin instead of the subtree CoM (any reference point gives the same M, bias forces and Jacobians).
#ifndef ABR_HAND_CUH_
#define ABR_HAND_CUH_

#include "abr_limb.cuh"

namespace abr {
namespace hand {

using limb::kStride;
constexpr int kNE = 4;     // global equality rows per world
constexpr int kLanes = 4;  // lanes per world (chains padded with dummy lanes)
// per-lane table, word (slot, lane) at T[slot * kStride + lane]
struct Map {
  int NL;
  __host__ __device__ constexpr int body(int p) const { return limb::kBodyW * (p - 1); }            // p = 1..NL, as limb::kBodyW
  __host__ __device__ constexpr int jnt(int p) const { return limb::kBodyW * NL + limb::kJntW * (p - 1); }
  __host__ __device__ constexpr int base() const { return limb::kBodyW * NL + limb::kJntW * NL; }  // world pose of the static parent: pos3 quat4
  __host__ __device__ constexpr int eq(int r) const { return base() + 7 + (kRowPrm + 5 + 1) * r; }  // prm[10] polycoef[5] valid
  __host__ __device__ constexpr int ijnt(int p) const { return eq(kNE) + limb::kJntI * (p - 1); }
  __host__ __device__ constexpr int ieq(int r) const { return ijnt(NL + 1) + 2 * r; }               // chain position of joint1 / joint2 in this lane (0 = none)
  __host__ __device__ constexpr int total() const { return ieq(kNE); }
};

#define HTF(slot) (C.T[(slot) * kStride])
#define HTI(slot) (__float_as_int(C.T[(slot) * kStride]))

template <int NL> struct Lane {
  float q[NL], v[NL], warm[NL], ctrl[NL], a[NL];
};
struct Cfg {
  const float* T;
  float dt, grav[3], tol, ls_tol, meaninertia;
  int iterations, ls_iterations, disableflags, nefc, nv, lg;
};
using NoShare = limb::ShareT<0>;  // nothing is shared between lanes: limb::ldl_factor / ldl_solve reduce to a plain local L'DL

template <int N> __device__ __forceinline__ float dotn(const float (&a)[N], const float (&b)[N]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; i++) s = fmaf(a[i], b[i], s);
  return s;
}
// out = M x for the lane's dense symmetric block (packed lower triangle)
template <int N> __device__ __forceinline__ void mul_m(const float (&M)[N * (N + 1) / 2], const float (&x)[N], float (&out)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < N; j++) s = fmaf(M[j <= i ? limb::TR(i, j) : limb::TR(j, i)], x[j], s);
    out[i] = s;
  }
}
// mjx.forward for one lane; M, fs, fc are returned for the implicit-damping Euler step
template <int NL> __device__ __forceinline__ void forward(Lane<NL>& s, const Cfg& C, float (&M)[NL * (NL + 1) / 2], float (&fs)[NL], float (&fc)[NL]) {
  using namespace limb;
  constexpr int N = NL, NTRI = N * (N + 1) / 2;
  constexpr Map mp{NL};
  const NoShare S{};
  const int lg = C.lg;
  // ---------------------------------------------------------------- kinematics (chain off a static body) + cinert + cdof
  float cinert[NL][10], cdof[N][6];
  int jflags[NL];
  {
    float pos[3] = {HTF(mp.base()), HTF(mp.base() + 1), HTF(mp.base() + 2)};
    float quat[4] = {HTF(mp.base() + 3), HTF(mp.base() + 4), HTF(mp.base() + 5), HTF(mp.base() + 6)};
    float Rp[9];
    q_to_mat(quat, Rp);
#pragma unroll
    for (int p = 1; p <= NL; p++) {
      const int fl = HTI(mp.ijnt(p));
      jflags[p - 1] = fl;
      const int type = fl & kJTypeMask;
      const float ca[3] = {HTF(mp.body(p)), HTF(mp.body(p) + 1), HTF(mp.body(p) + 2)};
      const float bq[4] = {HTF(mp.body(p) + 3), HTF(mp.body(p) + 4), HTF(mp.body(p) + 5), HTF(mp.body(p) + 6)};
      const float bj[4] = {HTF(mp.body(p) + 10), HTF(mp.body(p) + 11), HTF(mp.body(p) + 12), HTF(mp.body(p) + 13)};
      const float cx[3] = {HTF(mp.body(p) + 21), HTF(mp.body(p) + 22), HTF(mp.body(p) + 23)};
      const float jp[3] = {HTF(mp.jnt(p)), HTF(mp.jnt(p) + 1), HTF(mp.jnt(p) + 2)};
      float r[3], anchor[3], axis[3];
      m_rot(Rp, ca, r);
#pragma unroll
      for (int i = 0; i < 3; i++) anchor[i] = pos[i] + r[i];
      m_rot(Rp, cx, axis);
      const float dq = s.q[p - 1] - HTF(mp.jnt(p) + 6);
      float sn, cs, qloc[4], qn[4], Rn[9];
      sincos_bf(((type == kJHinge) ? dq : 0.f) * 0.5f, sn, cs);
#pragma unroll
      for (int i = 0; i < 4; i++) qloc[i] = fmaf(sn, bj[i], cs * bq[i]);
      q_mul(quat, qloc, qn);
      q_to_mat(qn, Rn);
      m_rot(Rn, jp, r);
      const float sl = (type == kJSlide) ? dq : 0.f;
#pragma unroll
      for (int i = 0; i < 3; i++) pos[i] = anchor[i] - r[i] + axis[i] * sl;
#pragma unroll
      for (int i = 0; i < 4; i++) quat[i] = qn[i];
#pragma unroll
      for (int i = 0; i < 9; i++) Rp[i] = Rn[i];
      const float ip[3] = {HTF(mp.body(p) + 7), HTF(mp.body(p) + 8), HTF(mp.body(p) + 9)};
      const float Ib[6] = {HTF(mp.body(p) + 15), HTF(mp.body(p) + 16), HTF(mp.body(p) + 17), HTF(mp.body(p) + 18), HTF(mp.body(p) + 19), HTF(mp.body(p) + 20)};
      float off[3], irot[6];
      m_rot(Rn, ip, r);
#pragma unroll
      for (int i = 0; i < 3; i++) off[i] = pos[i] + r[i];  // inertial frame origin relative to the world origin (the reference point)
      rot_inertia(Rn, Ib, irot);
      const float ms = HTF(mp.body(p) + 14);
      const float oo = v_dot(off, off);
      constexpr int ra[6] = {0, 1, 2, 0, 0, 1}, cb[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
      for (int e = 0; e < 6; e++) cinert[p - 1][e] = irot[e] + ms * ((ra[e] == cb[e] ? oo : 0.f) - off[ra[e]] * off[cb[e]]);
      cinert[p - 1][6] = off[0] * ms; cinert[p - 1][7] = off[1] * ms; cinert[p - 1][8] = off[2] * ms; cinert[p - 1][9] = ms;
      const bool hinge = type == kJHinge;
      const float o2[3] = {-anchor[0], -anchor[1], -anchor[2]};
      float cr[3];
      v_cross(axis, o2, cr);
#pragma unroll
      for (int i = 0; i < 3; i++) { cdof[p - 1][i] = hinge ? axis[i] : 0.f; cdof[p - 1][3 + i] = hinge ? cr[i] : axis[i]; }
    }
  }
  // ---------------------------------------------------------------- crb + M, factor
  {
    float up[10];
#pragma unroll
    for (int i = 0; i < 10; i++) up[i] = 0.f;
#pragma unroll
    for (int p = NL; p >= 1; p--) {
      const int i = p - 1;
      float crb[10], buf[6];
#pragma unroll
      for (int e = 0; e < 10; e++) { crb[e] = cinert[i][e] + up[e]; up[e] = crb[e]; }
      inert_mul(crb, cdof[i], buf);
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (j <= i) {
          float t = cdof[j][0] * buf[0] + cdof[j][1] * buf[1] + cdof[j][2] * buf[2] + cdof[j][3] * buf[3] + cdof[j][4] * buf[4] + cdof[j][5] * buf[5];
          if (i == j) t += HTF(mp.jnt(p) + 10);
          M[TR(i, j)] = t;
        }
      }
    }
  }
  float F[NTRI], invD[N];
#pragma unroll
  for (int e = 0; e < NTRI; e++) F[e] = M[e];
  ldl_factor<N>(F, invD, S);
  // ---------------------------------------------------------------- velocity pass: com_vel + rne + passive + actuation
  {
    float cfrc[NL][6], cdd[6];
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool grav = !(C.disableflags & ABR_DSBL_GRAVITY);
    float ca[6] = {0.f, 0.f, 0.f, grav ? -C.grav[0] : 0.f, grav ? -C.grav[1] : 0.f, grav ? -C.grav[2] : 0.f};
#pragma unroll
    for (int d = 0; d < N; d++) {
      motion_cross(cv, cdof[d], cdd);
      const float qd = s.v[d];
#pragma unroll
      for (int i = 0; i < 6; i++) { cv[i] = fmaf(cdof[d][i], qd, cv[i]); ca[i] = fmaf(cdd[i], qd, ca[i]); }
      float f1[6], f2[6], f3[6];
      inert_mul(cinert[d], ca, f1);
      inert_mul(cinert[d], cv, f2);
      motion_cross_force(cv, f2, f3);
#pragma unroll
      for (int i = 0; i < 6; i++) cfrc[d][i] = f1[i] + f3[i];
    }
    float up[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool passive_on = !(C.disableflags & ABR_DSBL_PASSIVE);
    const bool act_on = !(C.disableflags & ABR_DSBL_ACTUATION);
#pragma unroll
    for (int p = NL; p >= 1; p--) {
      const int d = p - 1;
      float f[6];
#pragma unroll
      for (int i = 0; i < 6; i++) { f[i] = cfrc[d][i] + up[i]; up[i] = f[i]; }
      const float bias = cdof[d][0] * f[0] + cdof[d][1] * f[1] + cdof[d][2] * f[2] + cdof[d][3] * f[3] + cdof[d][4] * f[4] + cdof[d][5] * f[5];
      const int fl = jflags[d];
      const float q = s.q[d];
      float t = 0.f;
      if (passive_on) t = -HTF(mp.jnt(p) + 8) * (q - HTF(mp.jnt(p) + 7)) - HTF(mp.jnt(p) + 9) * s.v[d];
      const float* prm = &HTF(mp.jnt(p) + 24);  // ctrlrange2 forcerange2 gainprm3 biasprm3 gear
      const int af = fl >> kJActShift;
      float ct = s.ctrl[d];
      if ((af & 1) && !(C.disableflags & ABR_DSBL_CLAMPCTRL)) ct = fminf(fmaxf(ct, prm[0]), prm[1 * kStride]);
      const float gear = prm[10 * kStride];
      const float len = q * gear, vel = s.v[d] * gear;
      float gain = prm[4 * kStride];
      if (af & 4) gain += prm[5 * kStride] * len + prm[6 * kStride] * vel;
      float bs = 0.f;
      if (af & 8) bs = prm[7 * kStride] + prm[8 * kStride] * len + prm[9 * kStride] * vel;
      float af_ = gain * ct + bs;
      if (af & 2) af_ = fminf(fmaxf(af_, prm[2 * kStride]), prm[3 * kStride]);
      af_ *= gear;
      t += ((fl & kJAct) && act_on) ? af_ : 0.f;
      fs[d] = t - bias;
    }
  }
  // ---------------------------------------------------------------- qacc_smooth
  float as[N];
#pragma unroll
  for (int i = 0; i < N; i++) as[i] = fs[i];
  ldl_solve<N>(F, invD, as, S);
  if (C.nefc == 0) {  // no constraint rows: MJX returns qacc_smooth without entering the solver, qacc_warmstart keeps its value
#pragma unroll
    for (int i = 0; i < N; i++) { s.a[i] = as[i]; fc[i] = 0.f; }
    return;
  }
  // ---------------------------------------------------------------- constraint rows: global equality rows, local limit rows
  float Je[kNE][N], eD[kNE], earef[kNE];
#pragma unroll
  for (int r = 0; r < kNE; r++) {
    const int e1 = HTI(mp.ieq(r)), e2 = HTI(mp.ieq(r) + 1);
    const bool valid = HTF(mp.eq(r) + kRowPrm + 5) != 0.f;
    const float* data = &HTF(mp.eq(r) + kRowPrm);
    float part = 0.f, c2 = 0.f;
#pragma unroll
    for (int p = 1; p <= NL; p++) {
      const float dif = s.q[p - 1] - HTF(mp.jnt(p) + 6);
      if (e1 == p) part += dif;
      if (e2 == p) {
        const float d2 = dif * dif, d3 = d2 * dif, d4 = d3 * dif;
        part -= data[1 * kStride] * dif + data[2 * kStride] * d2 + data[3 * kStride] * d3 + data[4 * kStride] * d4;
        c2 = -(data[1 * kStride] + data[2 * kStride] * dif * 2.f + data[3 * kStride] * d2 * 3.f + data[4 * kStride] * d3 * 4.f);
      }
    }
    float jvel = 0.f;
#pragma unroll
    for (int d = 0; d < N; d++) {
      Je[r][d] = ((e1 == d + 1) ? 1.f : 0.f) + ((e2 == d + 1) ? c2 : 0.f);
      jvel = fmaf(Je[r][d], s.v[d], jvel);
    }
    const float pos = gall(part, lg) - data[0];  // polycoef[0] once per row
    jvel = gall(jvel, lg);
    float D, aref;
    row_kbi<false, kStride>(&HTF(mp.eq(r)), pos, jvel, HTF(mp.eq(r) + 7), valid, D, aref);
    eD[r] = D; earef[r] = aref;
  }
  float lD[NL], laref[NL], lsg[NL];
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    const int d = p - 1;
    const float q = s.q[d];
    const float dmin = q - HTF(mp.jnt(p) + 11), dmax = HTF(mp.jnt(p) + 12) - q;
    const float pos = fminf(dmin, dmax) - HTF(mp.jnt(p) + 13);
    const bool active = (pos < 0.f) && (jflags[d] & kJLimited);
    const float sg = (dmin < dmax) ? 1.f : -1.f;
    lsg[d] = active ? sg : 0.f;
    row_kbi<false, kStride>(&HTF(mp.jnt(p) + 14), pos, sg * s.v[d], HTF(mp.jnt(p) + 14 + 7), active, lD[d], laref[d]);
  }
  // ---------------------------------------------------------------- solver.solve (Newton)
  // state of a point: x, M x (local), limit residuals (local), equality residuals (replicated)
  auto residuals = [&](const float (&x)[N], float (&Jl)[NL], float (&Jq)[kNE]) {
#pragma unroll
    for (int d = 0; d < N; d++) Jl[d] = lsg[d] * x[d] - laref[d];
#pragma unroll
    for (int r = 0; r < kNE; r++) Jq[r] = gall(dotn<N>(Je[r], x), lg) - earef[r];
  };
  auto point_cost = [&](const float (&x)[N], const float (&Mx)[N], const float (&Jl)[NL], const float (&Jq)[kNE], float& gauss) {
    float sc = 0.f, g = 0.f;
#pragma unroll
    for (int d = 0; d < N; d++) {
      if (Jl[d] < 0.f) sc = fmaf(lD[d] * Jl[d], Jl[d], sc);
      g = fmaf(Mx[d] - fs[d], x[d] - as[d], g);
    }
    sc = gall(sc, lg); g = gall(g, lg);
#pragma unroll
    for (int r = 0; r < kNE; r++) sc = fmaf(eD[r] * Jq[r], Jq[r], sc);  // equality rows are always active
    gauss = 0.5f * g;
    return 0.5f * sc + 0.5f * g;
  };
  float Ma[N], Jl[NL], Jq[kNE], gauss, cost;
#pragma unroll
  for (int d = 0; d < N; d++) { s.a[d] = as[d]; Ma[d] = fs[d]; }
  residuals(as, Jl, Jq);
  cost = point_cost(as, Ma, Jl, Jq, gauss);
  if (!(C.disableflags & ABR_DSBL_WARMSTART)) {
    float Mw[N], Jlw[NL], Jqw[kNE], g2;
    mul_m<N>(M, s.warm, Mw);
    residuals(s.warm, Jlw, Jqw);
    const float c2 = point_cost(s.warm, Mw, Jlw, Jqw, g2);
    const bool use = c2 < cost;
    cost = use ? c2 : cost; gauss = use ? g2 : gauss;
#pragma unroll
    for (int d = 0; d < N; d++) { s.a[d] = use ? s.warm[d] : s.a[d]; Ma[d] = use ? Mw[d] : Ma[d]; Jl[d] = use ? Jlw[d] : Jl[d]; }
#pragma unroll
    for (int r = 0; r < kNE; r++) Jq[r] = use ? Jqw[r] : Jq[r];
  }
  float prev_cost = INFINITY;
  const float scale = 1.f / (C.meaninertia * (float)max(1, C.nv));
  bool live = true;
  for (int niter = 0;; niter++) {
    // qfrc_constraint at the current point
#pragma unroll
    for (int d = 0; d < N; d++) {
      float t = (Jl[d] < 0.f) ? -lD[d] * Jl[d] * lsg[d] : 0.f;
#pragma unroll
      for (int r = 0; r < kNE; r++) t = fmaf(-eD[r] * Jq[r], Je[r][d], t);
      fc[d] = t;
    }
    if (niter >= C.iterations) break;
    float grad[N];
#pragma unroll
    for (int d = 0; d < N; d++) grad[d] = Ma[d] - fs[d] - fc[d];
    // Newton direction: H = blockdiag(M_lane + limit rows) + sum_r D_r u_r u_r' is assembled REPLICATED in every lane of the world
    // (order kLanes * NL) and factored densely, as MJX factors its dense H. (The Woodbury route through the lanes' own factors is
    // cheaper but cancels catastrophically in float32: the equality rows are stiff and the finger inertias tiny.)
    constexpr int NF = kLanes * NL, NFT = NF * (NF + 1) / 2;
    float Hf[NFT], hD[NF], gf[NF];
    const int lane0 = (threadIdx.x & 31) & ~(kLanes - 1);
#pragma unroll
    for (int e = 0; e < NFT; e++) Hf[e] = 0.f;
#pragma unroll
    for (int l = 0; l < kLanes; l++) {
#pragma unroll
      for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) {
          if (j <= i) {
            float v = M[TR(i, j)];
            if (i == j && Jl[i] < 0.f) v = fmaf(lD[i] * lsg[i], lsg[i], v);
            Hf[TR(NL * l + i, NL * l + j)] = __shfl_sync(ABR_FULL, v, lane0 + l);
          }
        }
        gf[NL * l + i] = __shfl_sync(ABR_FULL, grad[i], lane0 + l);
      }
    }
#pragma unroll
    for (int r = 0; r < kNE; r++) {
      float uf[NF];
#pragma unroll
      for (int l = 0; l < kLanes; l++)
#pragma unroll
        for (int d = 0; d < N; d++) uf[NL * l + d] = __shfl_sync(ABR_FULL, Je[r][d], lane0 + l);
#pragma unroll
      for (int i = 0; i < NF; i++) {
        const float t = eD[r] * uf[i];
#pragma unroll
        for (int j = 0; j < NF; j++)
          if (j <= i) Hf[TR(i, j)] = fmaf(t, uf[j], Hf[TR(i, j)]);
      }
    }
    ldl_factor<NF>(Hf, hD, S);
    ldl_solve<NF>(Hf, hD, gf, S);
    float mg[N];
    const int gl = (threadIdx.x & 31) & (kLanes - 1);
#pragma unroll
    for (int d = 0; d < N; d++) {
      float v = gf[d];
#pragma unroll
      for (int l = 1; l < kLanes; l++) v = (gl == l) ? gf[NL * l + d] : v;
      mg[d] = v;
    }
    if (C.iterations != 1) {
      float gn = gall(dotn<N>(grad, grad), lg);
      bool done = scale * (prev_cost - cost) < C.tol;
      done = done || (scale * sqrtf(gn) < C.tol);
      live = live && !done;
      if (!__any_sync(ABR_FULL, live)) break;
    }
    // ---- exact line search (solver._linesearch): equality rows are always active, so they are part of the constant quadratic
    float search[N], mv[N], jvl[NL], jvq[kNE];
#pragma unroll
    for (int d = 0; d < N; d++) search[d] = -mg[d];
    mul_m<N>(M, search, mv);
#pragma unroll
    for (int d = 0; d < N; d++) jvl[d] = lsg[d] * search[d];
#pragma unroll
    for (int r = 0; r < kNE; r++) jvq[r] = gall(dotn<N>(Je[r], search), lg);
    float sn = gall(dotn<N>(search, search), lg), sMa = gall(dotn<N>(search, Ma), lg), sq = gall(dotn<N>(search, fs), lg), smv = gall(dotn<N>(search, mv), lg);
    const float smag = sqrt_fast(sn) * C.meaninertia * (float)max(1, C.nv);
    const float gtol = C.tol * C.ls_tol * smag;
    float qg0 = gauss, qg1 = sMa - sq, qg2 = 0.5f * smv;
#pragma unroll
    for (int r = 0; r < kNE; r++) {
      qg0 = fmaf(0.5f * eD[r] * Jq[r], Jq[r], qg0); qg1 = fmaf(eD[r] * jvq[r], Jq[r], qg1); qg2 = fmaf(0.5f * eD[r] * jvq[r], jvq[r], qg2);
    }
    float la0[NL], la1[NL], la2[NL];
#pragma unroll
    for (int d = 0; d < NL; d++) {
      const float ja = Jl[d], w = jvl[d], Dr = lD[d];
      la0[d] = 0.5f * ja * ja * Dr; la1[d] = w * ja * Dr; la2[d] = 0.5f * w * w * Dr;
    }
#define HLS_EVAL(al) ls_eval<NL, 0>(Jl, jvl, la1, la2, (al), qg1, qg2, lg)
    const LSP p0 = ls_eval<NL, 0, true>(Jl, jvl, la1, la2, 0.f, qg1, qg2, lg);
    const LSP l0 = HLS_EVAL(-safe_div_fast(p0.d0, p0.d1));
    const bool lesser = l0.d0 < p0.d0;
    LSP hi = lesser ? p0 : l0;
    LSP lo = lesser ? l0 : p0;
    bool swap = true;
    int it = 0;
    while (true) {
      bool done = it >= C.ls_iterations;
      done = done || !swap;
      done = done || ((lo.d0 < 0.f) && (lo.d0 > -gtol));
      done = done || ((hi.d0 > 0.f) && (hi.d0 < gtol));
      if (!__any_sync(ABR_FULL, !done)) break;
      const LSP lo_next = HLS_EVAL(lo.alpha - safe_div_fast(lo.d0, lo.d1));
      const LSP hi_next = HLS_EVAL(hi.alpha - safe_div_fast(hi.d0, hi.d1));
      const LSP mid = HLS_EVAL(0.5f * (lo.alpha + hi.alpha));
      if (!done) {
        const bool swap_lo_next = (lo.d0 > 0.f) || (lo.d0 < lo_next.d0);
        if (swap_lo_next) lo = lo_next;
        const bool swap_lo_mid = (mid.d0 < 0.f) && (lo.d0 < mid.d0);
        if (swap_lo_mid) lo = mid;
        const bool swap_hi_next = (hi.d0 < 0.f) || (hi.d0 > hi_next.d0);
        if (swap_hi_next) hi = hi_next;
        const bool swap_hi_mid = (mid.d0 > 0.f) && (hi.d0 > mid.d0);
        if (swap_hi_mid) hi = mid;
        swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
        it++;
      }
    }
#undef HLS_EVAL
    const float c_p0 = ls_cost<NL, 0, true>(Jl, jvl, la0, p0, qg0, lg), c_lo = ls_cost<NL, 0>(Jl, jvl, la0, lo, qg0, lg), c_hi = ls_cost<NL, 0>(Jl, jvl, la0, hi, qg0, lg);
    const bool improved = (c_lo < c_p0) || (c_hi < c_p0);
    const float alpha = (improved && live) ? ((c_lo < c_hi) ? lo.alpha : hi.alpha) : 0.f;
#pragma unroll
    for (int d = 0; d < N; d++) { s.a[d] = fmaf(search[d], alpha, s.a[d]); Ma[d] = fmaf(mv[d], alpha, Ma[d]); Jl[d] = fmaf(jvl[d], alpha, Jl[d]); }
#pragma unroll
    for (int r = 0; r < kNE; r++) Jq[r] = fmaf(jvq[r], alpha, Jq[r]);
    if (C.iterations != 1) {
      float g2;
      const float c2 = point_cost(s.a, Ma, Jl, Jq, g2);
      if (live) { prev_cost = cost; cost = c2; gauss = g2; }
    }
  }
#pragma unroll
  for (int d = 0; d < N; d++) s.warm[d] = s.a[d];
}

// forward.euler (+ implicit joint damping unless EULERDAMP is disabled)
template <int NL> __device__ __forceinline__ void euler(Lane<NL>& s, const Cfg& C, float (&M)[NL * (NL + 1) / 2], const float (&fs)[NL], const float (&fc)[NL]) {
  constexpr Map mp{NL};
  const float dt = C.dt;
  if (!(C.disableflags & ABR_DSBL_EULERDAMP)) {
    float hD[NL], rhs[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
      M[limb::TR(i, i)] += HTF(mp.jnt(i + 1) + 9) * dt;
      rhs[i] = fs[i] + fc[i];
    }
    const NoShare S{};
    limb::ldl_factor<NL>(M, hD, S);
    limb::ldl_solve<NL>(M, hD, rhs, S);
#pragma unroll
    for (int i = 0; i < NL; i++) s.a[i] = rhs[i];
  }
#pragma unroll
  for (int d = 0; d < NL; d++) { s.v[d] = fmaf(s.a[d], dt, s.v[d]); s.q[d] = fmaf(dt, s.v[d], s.q[d]); }
}

__device__ __forceinline__ Cfg make_cfg(const Layout& L, const float* T, int g) {
  Cfg C;
  C.T = T + g;
  C.dt = L.timestep; C.grav[0] = L.gravity[0]; C.grav[1] = L.gravity[1]; C.grav[2] = L.gravity[2];
  C.tol = L.tolerance; C.ls_tol = L.ls_tolerance; C.meaninertia = L.meaninertia;
  C.iterations = L.iterations; C.ls_iterations = L.ls_iterations; C.disableflags = L.disableflags; C.nefc = L.nefc; C.nv = L.nv; C.lg = L.lg2G;
  return C;
}

// shoot (shooting.py:22-48) / the sampler's rollouts (shooting.py:140-153) for fixed-base chains
template <int NL>
__global__ void __launch_bounds__(128) k_hand_rollout(const __grid_constant__ Layout L, const __grid_constant__ RolloutArgs A) {
  extern __shared__ __align__(16) float smem[];
  constexpr Map mp{NL};
  constexpr int NTRI = NL * (NL + 1) / 2;
  const int ntab = mp.total() * kStride;
  const int nx = L.nx, nu = L.nu, nq = L.nq, Nh = A.N;
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) smem[i] = A.blob[L.f_htab + i];
  float* cqd = smem + ntab; float* cqf = cqd + nx; float* crd = cqf + nx; float* cxg = crd + nu;
  if (A.cost.enabled) {
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { cqd[i] = A.cost.qd[i]; cqf[i] = A.cost.qfd[i]; cxg[i] = A.cost.xg[i]; }
    for (int i = threadIdx.x; i < nu; i += blockDim.x) crd[i] = A.cost.rd[i];
  }
  __syncthreads();
  const int lg = L.lg2G;
  const int g = threadIdx.x & ((1 << lg) - 1);
  const int wraw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> lg);
  const bool valid = wraw < A.nworld;
  const int w = valid ? wraw : A.nworld - 1;
  const Cfg C = make_cfg(L, smem, g);
  const bool sampler = A.mode == 1;
  int prob = w, sample = 0;
  if (sampler) {
    if (A.sample_ids) { prob = w; sample = A.sample_ids[w]; }
    else { prob = w / A.S; sample = A.sample_offset + (w - prob * A.S); }
  }
  const float* x0 = A.x0 + (size_t)(sampler ? prob : w) * A.x0_stride;
  Lane<NL> s;
  int gd[NL], gq[NL], ga[NL];
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    gd[p - 1] = HTI(mp.ijnt(p) + 1); gq[p - 1] = HTI(mp.ijnt(p) + 2); ga[p - 1] = HTI(mp.ijnt(p) + 3);
    s.q[p - 1] = (gd[p - 1] >= 0) ? x0[gq[p - 1]] : 0.f;
    s.v[p - 1] = (gd[p - 1] >= 0) ? x0[nq + gd[p - 1]] : 0.f;
    s.warm[p - 1] = 0.f; s.ctrl[p - 1] = 0.f; s.a[p - 1] = 0.f;
  }
  float* xs = A.xs_out ? A.xs_out + (size_t)w * (Nh + 1) * nx : nullptr;
  // running quadratic cost with diagonal weights (cost.py:62-85); every state entry belongs to exactly one lane
  auto quad_x = [&](bool terminal) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < NL; d++) {
      if (gd[d] >= 0) {
        const float* wq = terminal ? cqf : cqd;
        const float e = s.q[d] - cxg[gq[d]]; acc = fmaf(wq[gq[d]] * e, e, acc);
        const float e2 = s.v[d] - cxg[nq + gd[d]]; acc = fmaf(wq[nq + gd[d]] * e2, e2, acc);
      }
    }
    return acc;
  };
  auto store_x = [&](float* x) {
#pragma unroll
    for (int d = 0; d < NL; d++)
      if (gd[d] >= 0) { x[gq[d]] = s.q[d]; x[nq + gd[d]] = s.v[d]; }
  };
  float cacc = 0.f;
  if (xs && valid) store_x(xs);
  if (A.cost.enabled) cacc += quad_x(Nh == 0);
  // t = -1 is mjx.forward with ctrl = 0, which seeds qacc_warmstart (shooting.py:36)
#pragma unroll 1
  for (int t = -1; t < Nh; t++) {
    if (t >= 0) {
#pragma unroll
      for (int d = 0; d < NL; d++) {
        float u = 0.f;
        if (ga[d] >= 0) {
          if (!sampler) {
            u = A.us[(size_t)w * A.us_stride + (size_t)t * nu + ga[d]];
          } else {
            float nz = 0.f;
            if (sample > 0) {
              if (A.noise) nz = A.noise[(((size_t)prob * (A.S_total - 1) + (sample - 1)) * Nh + t) * nu + ga[d]];
              else nz = limb::philox_normal_ool(A.seed, (uint32_t)sample, (uint32_t)prob, (uint32_t)(t * nu + ga[d]));
            }
            const float vv = A.us[(size_t)prob * A.us_stride + (size_t)t * nu + ga[d]] + nz * A.stdev;
            u = fminf(fmaxf(vv, HTF(mp.jnt(d + 1) + 24)), HTF(mp.jnt(d + 1) + 25));  // clip to actuator_ctrlrange (shooting.py:146-148)
          }
          if (A.us_out && valid) A.us_out[((size_t)w * Nh + t) * nu + ga[d]] = u;
          if (A.cost.enabled) cacc = fmaf(crd[ga[d]] * u, u, cacc);
        }
        s.ctrl[d] = u;
      }
    }
    float M[NTRI], fs[NL], fc[NL];
    forward<NL>(s, C, M, fs, fc);
    if (t >= 0) {
      euler<NL>(s, C, M, fs, fc);
      if (xs && valid) store_x(xs + (size_t)(t + 1) * nx);
      if (A.cost.enabled) cacc += quad_x(t == Nh - 1);
    }
  }
  if (A.costs_out) {
    cacc = limb::gall(cacc, lg);
    if (valid && g == 0) A.costs_out[w] = 0.5f * cacc;
  }
}

// MjxEnv.pipeline_init / pipeline_step (rl/base.py:81-96) with the auto-reset blend, for fixed-base chains (the reference's own
// RL example is the pendulum, rl/pendulum/swingup.py). The fused task epilogue and per-env randomisation stay on the generic kernels.
template <int NL>
__global__ void __launch_bounds__(128) k_hand_env(const __grid_constant__ Layout L, const __grid_constant__ EnvArgs A) {
  extern __shared__ __align__(16) float smem[];
  constexpr Map mp{NL};
  constexpr int NTRI = NL * (NL + 1) / 2;
  const int ntab = mp.total() * kStride;
  const int nu = L.nu, nq = L.nq, nv = L.nv;
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) smem[i] = A.blob[L.f_htab + i];
  __syncthreads();
  const int lg = L.lg2G;
  const int g = threadIdx.x & ((1 << lg) - 1);
  const int wraw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> lg);
  const bool valid = wraw < A.E;
  const int w = valid ? wraw : A.E - 1;
  const Cfg C = make_cfg(L, smem, g);
  const bool reset = A.reset_mask && A.reset_mask[w];
  const float* sq = (reset ? A.first_qpos : A.qpos) + (size_t)w * nq;
  const float* sv = (reset ? A.first_qvel : A.qvel) + (size_t)w * nv;
  const float* sw = reset ? A.first_warm : A.warm;
  if (sw) sw += (size_t)w * nv;
  Lane<NL> s;
  int gd[NL], gq[NL];
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    gd[p - 1] = HTI(mp.ijnt(p) + 1); gq[p - 1] = HTI(mp.ijnt(p) + 2);
    const int ga = HTI(mp.ijnt(p) + 3);
    s.q[p - 1] = (gd[p - 1] >= 0) ? sq[gq[p - 1]] : 0.f;
    s.v[p - 1] = (gd[p - 1] >= 0) ? sv[gd[p - 1]] : 0.f;
    s.warm[p - 1] = (gd[p - 1] >= 0 && sw) ? sw[gd[p - 1]] : 0.f;
    s.ctrl[p - 1] = (ga >= 0 && A.ctrl) ? A.ctrl[(size_t)w * nu + ga] : 0.f;
    s.a[p - 1] = 0.f;
  }
  const int nfw = A.forward_only ? 1 : A.nsubsteps;
#pragma unroll 1
  for (int it = 0; it < nfw; it++) {
    float M[NTRI], fs[NL], fc[NL];
    forward<NL>(s, C, M, fs, fc);
    if (!A.forward_only) euler<NL>(s, C, M, fs, fc);
  }
  if (valid) {
    float* oq = A.qpos + (size_t)w * nq; float* ov = A.qvel + (size_t)w * nv;
    float* ow = A.warm ? A.warm + (size_t)w * nv : nullptr; float* oa = A.qacc ? A.qacc + (size_t)w * nv : nullptr;
#pragma unroll
    for (int d = 0; d < NL; d++) {
      if (gd[d] >= 0) {
        oq[gq[d]] = s.q[d];
        if (!A.forward_only) ov[gd[d]] = s.v[d];
        if (ow) ow[gd[d]] = s.warm[d];
        if (oa) oa[gd[d]] = s.a[d];
      }
    }
    if (A.time && g == 0) {
      const float t0 = reset ? 0.f : A.time[w];
      A.time[w] = A.forward_only ? t0 : t0 + L.timestep * (float)A.nsubsteps;
    }
  }
}

template <int NL> int launch_hand_env_t(const Layout& L, const EnvArgs& a, cudaStream_t st) {
  constexpr Map mp{NL};
  const long threads = (long)a.E << L.lg2G;
  const int tpb = threads <= 32 ? 32 : (threads <= 64 * 148 ? 64 : 128);
  const size_t sm = sizeof(float) * ((size_t)mp.total() * kStride);
  cudaError_t e = cudaFuncSetAttribute(k_hand_env<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  k_hand_env<NL><<<(int)((threads + tpb - 1) / tpb), tpb, sm, st>>>(L, a);
  return (int)cudaGetLastError();
}

template <int NL> int launch_hand_rollout_t(const Layout& L, const RolloutArgs& a, cudaStream_t st) {
  constexpr Map mp{NL};
  const long threads = (long)a.nworld << L.lg2G;
  const int tpb = threads <= 32 ? 32 : (threads <= 64 * 148 ? 64 : 128);  // small solves: spread the warps over the SMs
  const size_t sm = sizeof(float) * ((size_t)mp.total() * kStride + 3 * L.nx + L.nu);
  cudaError_t e = cudaFuncSetAttribute(k_hand_rollout<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)((threads + tpb - 1) / tpb);
  k_hand_rollout<NL><<<grid, tpb, sm, st>>>(L, a);
  return (int)cudaGetLastError();
}

#undef HTF
#undef HTI

}  // namespace hand

int launch_hand_rollout_3(const Layout& L, const RolloutArgs& a, cudaStream_t st);
int launch_hand_env_3(const Layout& L, const EnvArgs& a, cudaStream_t st);

}  // namespace abr
#endif
