// abr_limb.cuh — the path-decomposed ("limb") world step: the fast path of the engine.
//
// A floating-base articulated body whose other bodies carry one hinge/slide joint each (quadruped,
// biped, exoskeleton, hand on a free base ...) is cut into its root-to-leaf paths. One LANE owns
// one path: trunk (free joint, 6 dofs) + NL chain bodies, so every local dof is an ancestor of all
// later ones and the lane's joint-space inertia / Newton Hessian is a plain dense lower triangle
// of order N = 6 + NL with compile-time indices. The whole world step is straight-line code over
// per-lane arrays that the compiler keeps in registers: no index tables, no shared-memory state,
// no __syncwarp, no divergence between lanes (per-lane differences are selects).
//
// Bodies shared by several paths (the trunk by all G lanes; e.g. a torso by its two arm paths)
// are processed redundantly and stay bit-identical across their lanes. Quantities that flow up the
// tree (composite inertia, RNE forces, L'DL Schur updates, M v and J'f partial sums, the subtree
// CoM) are held as per-lane SHARES on shared bodies and summed over the sharing lane group with
// xor-butterfly shuffles right before they are consumed ("merge"); a contribution produced on a
// shared body is added by that body's owner lane only, so it is counted once. G worlds-lanes are
// packed 32/G worlds to a warp.
//
// Replaces the third-party mjx.step behind ambersim/trajopt/shooting.py:41 and
// ambersim/rl/base.py:93 for eligible models (see build_limb in abr_engine.cu); everything else
// runs on the generic G-lane kernel (abr_step.cuh). Formulas follow SURVEY.md Appendix A.
#ifndef ABR_LIMB_CUH_
#define ABR_LIMB_CUH_

#include <stdlib.h>

#include "abr_kernels.cuh"

namespace abr {
namespace limb {

// ---- per-lane model table: word (slot, lane) lives at T[slot * kStride + lane]
constexpr int kStride = 8;  // max lanes per world
// In shared memory the table is lane-major: word (slot, lane) at S[lane * tpad(total) + slot], every lane's column 16-byte aligned and
// every block of Map a multiple of 4 words, so that neighbouring slots load as LDS.64 / LDS.128; tpad = 4 (mod 32) keeps the columns of up
// to 8 lanes on disjoint banks.
constexpr int kLS = 1;  // slot stride in shared memory
__host__ __device__ constexpr int tpad(int total) { return ((total - 4 + 31) / 32) * 32 + 4; }
constexpr int kBodyW = 24;  // anchor in the parent frame (body_pos + R(body_quat) jnt_pos)3 | body_quat4 | body_ipos3 | body_quat o (0, jnt_axis)4 |
                            // mass | inertia tensor in the body frame (xx yy zz xy xz yz)6 | joint axis in the parent frame3
constexpr int kJntW = 36;   // jnt_pos3 jnt_axis3 qpos0 qpos_spring stiffness damping armature range2 margin | limit prm[10] | act prm[11]
constexpr int kConW = 36;   // geom_pos3 radius | con prm[14] | plane normal3 point3 frame9
constexpr int kJntI = 4;    // flags, global dof, global qpos adr, global actuator
constexpr int kConI = 2;    // local body position of the sphere (-1 = empty slot), condim
struct Map {
  int NL, NC;
  __host__ __device__ constexpr int body(int p) const { return kBodyW * p; }
  __host__ __device__ constexpr int jnt(int p) const { return kBodyW * (NL + 1) + kJntW * (p - 1); }
  __host__ __device__ constexpr int trunk() const { return kBodyW * (NL + 1) + kJntW * NL; }  // damping[6] armature[6]
  __host__ __device__ constexpr int con(int c) const { return trunk() + 12 + kConW * c; }
  __host__ __device__ constexpr int ijnt(int p) const { return con(NC) + kJntI * (p - 1); }
  __host__ __device__ constexpr int ish() const { return ijnt(NL + 1); }  // own bits, level bits
  __host__ __device__ constexpr int icon(int c) const { return ish() + 2 + kConI * c; }
  __host__ __device__ constexpr int total() const { return icon(NC); }
};
// joint flag bits
constexpr int kJHinge = 1, kJSlide = 2, kJTypeMask = 3, kJLimited = 4, kJAct = 8, kJActShift = 4;

__host__ __device__ constexpr int PD(int d) { return d < 6 ? 0 : d - 5; }   // dof -> position
__host__ __device__ constexpr int FD(int p) { return p == 0 ? 0 : 5 + p; }  // first dof of a position
__host__ __device__ constexpr int LD(int p) { return 5 + p; }               // last dof of a position
__host__ __device__ constexpr int TR(int i, int j) { return i * (i + 1) / 2 + j; }

// per-lane sharing info. LGC >= 0 fixes the sharing PATTERN at compile time: LGC holds 2 bits per chain
// position = log2 of the largest lane group sharing a body there (position 0, the trunk, is shared by all
// 2^(LGC & 3) lanes), so merges at unshared positions and their level tests fold away; which lanes share and
// who owns is still read per lane. LGC = 2 is the flat quadruped pattern (only the trunk shared, by 4 lanes),
// LGC = 86 the biped's (trunk by 4, three torso positions by 2). LGC = -1 reads the pattern from the table.
template <int LGC> struct ShareT {
  int own;   // bit p: this lane is the owner of its body at position p (always set on private bodies)
  int lvl;   // 2 bits per position: log2 of the lane-group size sharing the body
  int mx;    // 2 bits per position: max of lvl over the lanes (uniform)
  int lg_;   // log2(G)
  __device__ __forceinline__ int m(int p) const { if (LGC >= 0) return (LGC >> (2 * p)) & 3; return (mx >> (2 * p)) & 3; }
  __device__ __forceinline__ bool o(int p) const { if (LGC >= 0 && m(p) == 0) return true; return (own >> p) & 1; }
  __device__ __forceinline__ int l(int p) const {
    if (LGC >= 0 && m(p) == 0) return 0;
    if (LGC >= 0 && p == 0) return LGC & 3;
    return (lvl >> (2 * p)) & 3;
  }
  __device__ __forceinline__ int lg() const { return LGC >= 0 ? (LGC & 3) : lg_; }
};

// sum over the lane group of level `lv` (per lane, uniform within a group); mxl is uniform
__device__ __forceinline__ float gs(float x, int lv, int mxl) {
  if (mxl > 0) { const float y = __shfl_xor_sync(ABR_FULL, x, 1); x = (lv > 0) ? x + y : x; }
  if (mxl > 1) { const float y = __shfl_xor_sync(ABR_FULL, x, 2); x = (lv > 1) ? x + y : x; }
  if (mxl > 2) { const float y = __shfl_xor_sync(ABR_FULL, x, 4); x = (lv > 2) ? x + y : x; }
  return x;
}
__device__ __forceinline__ float gall(float x, int lg) { return gs(x, lg, lg); }


// 1/x as the bare MUFU.RCP (about 1 ulp): the operands here (clamped pivots, impedances in (0,1), regularisers >= mjMINVAL,
// line-search curvatures) never reach the magnitudes whose scaling fix-ups make the compiler's `1.f / x` nine instructions long
// ABR_PRECISE_MATH (`make precise`): correctly rounded reciprocal / square root and libdevice's sincosf instead, to measure what the
// fast forms change and cost (tools/precise_vs_fast.py, profiles/r2_precise_math.txt)
__device__ __forceinline__ float rcp_fast(float x) {
#ifdef ABR_PRECISE_MATH
  return __frcp_rn(x);
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
__device__ __forceinline__ float sqrt_fast(float x) {
#ifdef ABR_PRECISE_MATH
  return __fsqrt_rn(x);
#else
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}
__device__ __forceinline__ float safe_div_fast(float a, float b) { return a * rcp_fast(b + ((b == 0.f) ? kMinVal : 0.f)); }

// sin/cos with a three-constant Cody-Waite reduction and Cephes minimax polynomials (about 1 ulp for
// |x| < 1e5): branch-free, so the straight-line step carries no Payne-Hanek slow path per call site.
__device__ __forceinline__ void sincos_bf(float x, float& sn, float& cs) {
#ifdef ABR_PRECISE_MATH
  sincosf(x, &sn, &cs);
  return;
#endif
  const float j = rintf(x * 0.636619772367581343f);
  float r = fmaf(-j, 1.57079625129699707031f, x);
  r = fmaf(-j, 7.54978941586159635335e-08f, r);
  r = fmaf(-j, 5.39030252995776476554e-15f, r);
  const int q = (int)j;
  const float r2 = r * r;
  const float sp = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
  const float cp = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f), fmaf(-0.5f, r2, 1.f));
  const float a = (q & 1) ? cp : sp, b = (q & 1) ? sp : cp;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}
__device__ __forceinline__ void axis_angle_quat_bf(const float* axis, float angle, float* q) {
  float sn, cs;
  sincos_bf(angle * 0.5f, sn, cs);
  q[0] = cs; q[1] = axis[0] * sn; q[2] = axis[1] * sn; q[3] = axis[2] * sn;
}
// x / max(|x|, tiny) with one reciprocal
template <int K> __device__ __forceinline__ float normalize_k(float (&x)[K]) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < K; i++) ss = fmaf(x[i], x[i], ss);
  const float nrm = sqrt_fast(ss);
  const float inv = rcp_fast((nrm == 0.f) ? 1e-6f : nrm);
#pragma unroll
  for (int i = 0; i < K; i++) x[i] *= inv;
  return nrm;
}
// The trunk's translational dofs are the unit twists e_3, e_4, e_5 (cdof[q] = [0 0 0 | e_q], q < 3). Multiplying by their
// zeros cannot be folded by the compiler (IEEE: 0 * x is not 0 for every x), so the loops over dofs go through these two
// helpers, which after unrolling reduce to a register pick for q < 3 (same values: only exact zeros are dropped).
template <int N> __device__ __forceinline__ float cdof_dot(const float (&cdof)[N][6], int d, const float* f) {
  if (d < 3) return f[3 + d];
  return cdof[d][0] * f[0] + cdof[d][1] * f[1] + cdof[d][2] * f[2] + cdof[d][3] * f[3] + cdof[d][4] * f[4] + cdof[d][5] * f[5];
}
template <int N> __device__ __forceinline__ void cdof_axpy(const float (&cdof)[N][6], int d, float x, float (&V)[6]) {
  if (d < 3) { V[3 + d] += x; return; }
#pragma unroll
  for (int i = 0; i < 6; i++) V[i] = fmaf(cdof[d][i], x, V[i]);
}
// inert_mul(I, [0 | v_lin]): the products with the exact zeros of the angular part dropped
__device__ __forceinline__ void inert_mul_lin(const float* I, const float* v, float* r) {
  r[0] = fmaf(-I[8], v[4], I[7] * v[5]);
  r[1] = fmaf(-I[6], v[5], I[8] * v[3]);
  r[2] = fmaf(-I[7], v[3], I[6] * v[4]);
  r[3] = I[9] * v[3]; r[4] = I[9] * v[4]; r[5] = I[9] * v[5];
}
// inert_mul(I, e_{3+q}): the momentum of a unit translation along axis q
__device__ __forceinline__ void inert_mul_unit(const float* I, int q, float* r) {
  // cross(I + 6, e_q) with the zero products dropped
  r[0] = (q == 1) ? -I[8] : ((q == 2) ? I[7] : 0.f);
  r[1] = (q == 0) ? I[8] : ((q == 2) ? -I[6] : 0.f);
  r[2] = (q == 0) ? -I[7] : ((q == 1) ? I[6] : 0.f);
  r[3] = (q == 0) ? I[9] : 0.f; r[4] = (q == 1) ? I[9] : 0.f; r[5] = (q == 2) ? I[9] : 0.f;
}
__device__ __forceinline__ void m_rot(const float (&R)[9], const float* v, float* r) {
  r[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  r[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  r[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
// R Ib R' for a symmetric Ib = (xx yy zz xy xz yz): the body-frame inertia tensor in the world frame, same packing
__device__ __forceinline__ void rot_inertia(const float (&R)[9], const float* I, float* out) {
  float T[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    T[3 * r + 0] = R[3 * r] * I[0] + R[3 * r + 1] * I[3] + R[3 * r + 2] * I[4];
    T[3 * r + 1] = R[3 * r] * I[3] + R[3 * r + 1] * I[1] + R[3 * r + 2] * I[5];
    T[3 * r + 2] = R[3 * r] * I[4] + R[3 * r + 1] * I[5] + R[3 * r + 2] * I[2];
  }
  constexpr int ra[6] = {0, 1, 2, 0, 0, 1}, cb[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int e = 0; e < 6; e++) out[e] = T[3 * ra[e]] * R[3 * cb[e]] + T[3 * ra[e] + 1] * R[3 * cb[e] + 1] + T[3 * ra[e] + 2] * R[3 * cb[e] + 2];
}
static __device__ __noinline__ float pow_cold(float x, float p) { return powf(x, p); }
// the sampler's counter-based normal: out of line so the shoot-mode loop body does not carry three copies of it
static __device__ __noinline__ float philox_normal_ool(unsigned long long seed, uint32_t sample, uint32_t problem, uint32_t index) {
  return philox_normal(seed, sample, problem, index);
}

// ------------------------------------------------------------------------------ dense local L'DL
// A: packed lower triangle in "share" form on shared rows. Leaves-first elimination (MuJoCo's
// L'DL order): after the call A holds D on the diagonal and D*L below it, replicated on shared rows.
template <int N, class SH> __device__ __forceinline__ void ldl_factor(float (&A)[N * (N + 1) / 2], float (&invD)[N], const SH& S) {
#pragma unroll
  for (int k = N - 1; k >= 0; k--) {
    const int p = PD(k);
    if (k == LD(p) && S.m(p)) {  // merge the rows of position p before they are used as pivots
      const int lv = S.l(p), mxl = S.m(p);
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++)
          if (PD(i) == p && j <= i) A[TR(i, j)] = gs(A[TR(i, j)], lv, mxl);
    }
    const float d = rcp_fast(fmaxf(A[TR(k, k)], kMinVal));
    invD[k] = d;
    const bool ownp = S.o(p);
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (i < k) {
        float t = A[TR(k, i)] * d;
        if (PD(i) != p) t = ownp ? t : 0.f;  // shallower rows hold shares: the owner adds the update
#pragma unroll
        for (int j = 0; j < N; j++)
          if (j <= i) A[TR(i, j)] = fmaf(-t, A[TR(k, j)], A[TR(i, j)]);
      }
    }
  }
}
// x <- (L'DL)^-1 x ; x comes in full (replicated on shared dofs) and leaves full
template <int N, class SH> __device__ __forceinline__ void ldl_solve(const float (&A)[N * (N + 1) / 2], const float (&invD)[N], float (&x)[N], const SH& S) {
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = S.o(PD(i)) ? x[i] : 0.f;  // to share form
#pragma unroll
  for (int k = N - 1; k >= 0; k--) {
    const int p = PD(k);
    if (k == LD(p) && S.m(p)) {
      const int lv = S.l(p), mxl = S.m(p);
#pragma unroll
      for (int i = 0; i < N; i++)
        if (PD(i) == p) x[i] = gs(x[i], lv, mxl);
    }
    const float t0 = x[k] * invD[k];
    const bool ownp = S.o(p);
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (i < k) {
        float t = A[TR(k, i)] * t0;
        if (PD(i) != p) t = ownp ? t : 0.f;
        x[i] -= t;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < N; k++) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++)
      if (i < k) acc = fmaf(A[TR(k, i)], x[i], acc);
    x[k] = (x[k] - acc) * invD[k];
  }
}
// out = M x, M full (replicated on shared rows), x full
template <int N, class SH> __device__ __forceinline__ void mul_m(const float (&M)[N * (N + 1) / 2], const float (&x)[N], float (&out)[N], const SH& S) {
  float up[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    float lo = 0.f;
    float u = 0.f;
#pragma unroll
    for (int j = 0; j < N; j++) {
      if (i < 3 && j < 3 && i != j) continue;  // the translational block of M is m * identity: exact zeros off its diagonal
      if (j <= i) lo = fmaf(M[TR(i, j)], x[j], lo);
      else if (PD(j) == PD(i)) lo = fmaf(M[TR(j, i)], x[j], lo);
      else u += S.o(PD(j)) ? M[TR(j, i)] * x[j] : 0.f;
    }
    out[i] = lo; up[i] = u;
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    const int p = PD(i);
    if (S.m(p)) up[i] = gs(up[i], S.l(p), S.m(p));
    out[i] += up[i];
  }
}

// ------------------------------------------------------------------------------ compile-time specialisation
// Code that a launch never executes still sits in the horizon loop's instruction stream and costs fetch bandwidth
// (measured: the implicit-damping Euler block, the multi-iteration blocks, the trajectory stores and the other
// mode's control path together cost 8 % at every batch size). SPEC < 0 keeps every decision at run time (any options);
// SPEC >= 0 fixes them: bit 0 = eulerdamp enabled, bit 1 = more than one Newton iteration, bit 2 = trajectories /
// controls are written out, bit 3 = sampler mode (controls generated from the guess + noise). The launcher picks the
// variant from the model's options and the call's arguments; the built fast variants have bit 1 clear (both values of bit 0) and also
// assume the rest of the common configuration: no disable flag set other than eulerdamp, the default impedance power 2 on every row, hinge joints only.
template <int SPEC> struct Spec {
  static constexpr bool gen = SPEC < 0;
  __device__ __forceinline__ static bool edamp(int disableflags) { return gen ? !(disableflags & ABR_DSBL_EULERDAMP) : ((SPEC & 1) != 0); }
  __device__ __forceinline__ static bool multi(int iterations) { return gen ? iterations != 1 : ((SPEC & 2) != 0); }
  __device__ __forceinline__ static bool out(const void* p) { return gen ? p != nullptr : (((SPEC & 4) != 0) && p != nullptr); }
  __device__ __forceinline__ static bool sampler(int mode) { return gen ? mode == 1 : ((SPEC & 8) != 0); }
  __device__ __forceinline__ static int flags(int disableflags) { return gen ? disableflags : ((SPEC & 1) ? 0 : (int)ABR_DSBL_EULERDAMP); }
  static constexpr bool pow2 = SPEC >= 0;
  static constexpr bool hinges = SPEC >= 0;  // every chain joint is a hinge (padding positions have a zero axis either way)
};

// probe (ABR_PROBE_STEPSYNC >= 2): CTA barriers at four uniform points inside the step as well (negative: profiles/r2_stepsync.txt)
#if defined(ABR_PROBE_STEPSYNC) && (ABR_PROBE_STEPSYNC + 0) >= 2
#define ABR_MIDSYNC() __syncthreads()
#else
#define ABR_MIDSYNC() ((void)0)
#endif
// ------------------------------------------------------------------------------ lane state
template <int NL, int NC> struct Lane {
  static constexpr int NP = NL + 1, N = 6 + NL, NTRI = N * (N + 1) / 2, NR = NL + 4 * NC;
  float qt[7];     // trunk qpos: position, quaternion
  float qc[NL];    // chain joint positions
  float v[N];      // qvel (trunk 6, chain)
  float warm[N];   // qacc_warmstart
  float ctrl[NL];  // control of the chain joint's actuator
  float a[N];      // qacc of the last forward
};

template <int LGC> struct LaneCfg {
  const float* T;  // table + lane-in-group
  ShareT<LGC> S;
  float dt, grav[3], mass, inv_mass, tol, ls_tol, meaninertia;
  float frs, acs, dps, ars;  // per-world domain randomisation: scales of contact friction, actuator strength, joint damping, joint armature (1 = the model's own)
  int iterations, ls_iterations, disableflags, nefc, nv;
};

#define LTF(slot) (C.T[(slot) * kLS])
#define LTI(slot) (__float_as_int(C.T[(slot) * kLS]))

// kbi: impedance, D and aref of one active row (constraint._kbi / _row in SURVEY App. A.7), split so that the four pyramid
// rows of a contact (same penetration, same solref / solimp) evaluate the impedance once
struct Kbi { float b, kip, g; };  // damping b, stiffness * impedance * pos, (1 - imp) / imp
template <bool POW2, int ST = kLS> __device__ __forceinline__ Kbi kbi_imp(const float* prm /* smem, slot stride ST (the hand tables: kStride) */, float pos) {
  const float k = prm[0 * ST], dmin = prm[2 * ST], dmax = prm[3 * ST];
  const float iw = prm[4 * ST], mid = prm[5 * ST];
  const float x = fabsf(pos) * iw;
  float ia, ib;
  if (POW2) {
    ia = x * x; const float t = 1.f - x; ib = t * t;
  } else {
    const float power = prm[6 * ST];
    if (power == 2.f) { ia = x * x; const float t = 1.f - x; ib = t * t; }
    else if (power == 1.f) { ia = x; ib = 1.f - x; }
    else { ia = pow_cold(x, power); ib = pow_cold(1.f - x, power); }
  }
  const float y = (x < mid) ? prm[8 * ST] * ia : 1.f - prm[9 * ST] * ib;
  float imp = dmin + y * (dmax - dmin);
  imp = fminf(fmaxf(imp, dmin), dmax);
  if (x > 1.f) imp = dmax;
  Kbi o;
  o.b = prm[1 * ST]; o.kip = k * imp * pos; o.g = (1.f - imp) * rcp_fast(imp);
  return o;
}
__device__ __forceinline__ void kbi_row(const Kbi& q, float jvel, float invw, bool active, float& D, float& aref) {
  const float R = fmaxf(invw * q.g, kMinVal);
  D = active ? rcp_fast(R) : 0.f;
  aref = active ? -q.b * jvel - q.kip : 0.f;
}
template <bool POW2, int ST = kLS> __device__ __forceinline__ void row_kbi(const float* prm, float pos, float jvel, float invw, bool active, float& D, float& aref) {
  kbi_row(kbi_imp<POW2, ST>(prm, pos), jvel, invw, active, D, aref);
}

struct LSP { float alpha, d0, d1, q1, q2; };  // the point's cost is alpha^2 q2 + alpha q1 + q0(alpha): q0 is only summed for the points whose cost is read

// constraint rows of one lane: NL joint-limit rows (chain dof 6 + r), then 4 pyramid rows per contact slot
// CB ("contact body") = every contact of a lane sits on the lane's leaf body and touches one plane: the contact
// Jacobian is then J_c = [g_c ; fr]' cdof with g_c[k] = off_c x fr[k], so instead of the 3 x N basis per contact
// the lane keeps 9 floats per contact and works through 6-vectors (body twist, total wrench, a 6x6 weight).
template <int NL, int NC, bool CB> struct Rows {
  static constexpr int N = 6 + NL, NR = NL + 4 * NC, NCC = NC > 0 ? NC : 1;
  float D[NR], aref[NR], lsg[NL];
  float B[CB ? 1 : NCC][3][CB ? 1 : N];  // contact Jacobian basis: normal, tangent 1, tangent 2
  float g[CB ? NCC : 1][3][3];           // CB: (contact point - leaf body origin) x frame_k
  float fr[3][3];                        // CB: the plane's contact frame
  float sref[3];                         // CB: leaf body origin - subtree CoM. The contact algebra is done about the leaf body's
                                         // origin (levers of a few cm) and shifted to / from the CoM-referenced twists once per 6-vector:
                                         // about the CoM the 6x6 weight would carry D * (1 m)^2 terms that cancel to D * (5 cm)^2
  float mu1[NCC], mu2[NCC];              // 0 on condim-1 slots
};
// CB: per-contact (normal, tangent1, tangent2) components of J x from the leaf body's twist V = sum_d cdof[d] x[d]
template <int NL, int NC> __device__ __forceinline__ void contact_bv(const Rows<NL, NC, true>& R, const float (&cdof)[6 + NL][6], const float (&x)[6 + NL],
                                                                     float (&bv)[NC > 0 ? NC : 1][3]) {
  float V[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int d = 0; d < 6 + NL; d++) cdof_axpy(cdof, d, x[d], V);
  float sh[3], lin[3];
  v_cross(V, R.sref, sh);  // linear velocity at the leaf body's origin
#pragma unroll
  for (int k = 0; k < 3; k++) lin[k] = R.fr[k][0] * (V[3] + sh[0]) + R.fr[k][1] * (V[4] + sh[1]) + R.fr[k][2] * (V[5] + sh[2]);
#pragma unroll
  for (int c = 0; c < NC; c++)
#pragma unroll
    for (int k = 0; k < 3; k++) bv[c][k] = fmaf(R.g[c][k][0], V[0], fmaf(R.g[c][k][1], V[1], fmaf(R.g[c][k][2], V[2], lin[k])));
}
// out = J x for the structured Jacobian
template <int NL, int NC, bool CB> __device__ __forceinline__ void mul_j(const Rows<NL, NC, CB>& R, const float (&cdof)[6 + NL][6], const float (&x)[6 + NL],
                                                                   float (&out)[NL + 4 * NC]) {
  constexpr int N = 6 + NL;
#pragma unroll
  for (int r = 0; r < NL; r++) out[r] = R.lsg[r] * x[6 + r];
  float bv[NC > 0 ? NC : 1][3];
  if constexpr (CB) {
    contact_bv<NL, NC>(R, cdof, x, bv);
  } else {
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float t = 0.f;
#pragma unroll
        for (int d = 0; d < N; d++) t = fmaf(R.B[c][k][d], x[d], t);
        bv[c][k] = t;
      }
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    out[NL + 4 * c + 0] = fmaf(R.mu1[c], bv[c][1], bv[c][0]);
    out[NL + 4 * c + 1] = fmaf(-R.mu1[c], bv[c][1], bv[c][0]);
    out[NL + 4 * c + 2] = fmaf(R.mu2[c], bv[c][2], bv[c][0]);
    out[NL + 4 * c + 3] = fmaf(-R.mu2[c], bv[c][2], bv[c][0]);
  }
}
// cost = 0.5 sum_active D Jaref^2 + 0.5 (Ma - fs).(a - as); uniform over the world's lanes
template <int NL, int NC, bool CB, class SH, bool AT_AS = false> __device__ __forceinline__ float solver_cost(const Rows<NL, NC, CB>& R, const SH& S, const float (&x)[6 + NL], const float (&Mx)[6 + NL],
                                                                       const float (&Jx)[NL + 4 * NC], const float (&fs)[6 + NL], const float (&as)[6 + NL], float& gauss) {
  float sc = 0.f, g = 0.f;
#pragma unroll
  for (int r = 0; r < NL + 4 * NC; r++)
    if (Jx[r] < 0.f) sc = fmaf(R.D[r] * Jx[r], Jx[r], sc);
  if (!AT_AS) {  // at x = qacc_smooth the Gauss term (M x - f).(x - as) is exactly zero
#pragma unroll
    for (int d = 0; d < 6 + NL; d++)
      if (S.o(PD(d))) g = fmaf(Mx[d] - fs[d], x[d] - as[d], g);
    g = gall(g, S.lg());
  }
  sc = gall(sc, S.lg());
  gauss = 0.5f * g;
  return 0.5f * sc + 0.5f * g;
}
// one point of the exact line search: (d0, d1) at alpha. a0/a1/a2 are the per-row quadratic
// coefficients 0.5 D ja^2, D ja jv, 0.5 D jv^2 (a row counts while ja + alpha jv < 0).
// Rows below R0 are skipped: the caller passes R0 = NL when no lane of the warp has an active joint-limit row (their
// coefficients are exact zeros then, so the sums are the same numbers).
// Z: the point alpha = 0 (the products with that exact zero, which the compiler may not fold, are left out).
template <int NR, int R0, bool Z = false> __device__ __forceinline__ LSP ls_eval(const float (&Jaref)[NR], const float (&jv)[NR], const float (&a1)[NR],
                                                         const float (&a2)[NR], float alpha, float qg1, float qg2, int lg) {
  float q1 = 0.f, q2 = 0.f;
#pragma unroll
  for (int r = R0; r < NR; r++) {
    if ((Z ? Jaref[r] : fmaf(alpha, jv[r], Jaref[r])) < 0.f) { q1 += a1[r]; q2 += a2[r]; }  // predicated adds, no selects
  }
  q1 = gall(q1, lg) + qg1; q2 = gall(q2, lg) + qg2;
  LSP pt;
  pt.alpha = Z ? 0.f : alpha; pt.q1 = q1; pt.q2 = q2;
  pt.d0 = Z ? q1 : 2.f * alpha * q2 + q1;
  pt.d1 = 2.f * q2 + ((q2 == 0.f) ? kMinVal : 0.f);
  return pt;
}
// the cost of a point (the loop's decisions only read d0 / d1, so q0 is summed for the three points whose cost is compared)
template <int NR, int R0, bool Z = false> __device__ __forceinline__ float ls_cost(const float (&Jaref)[NR], const float (&jv)[NR], const float (&a0)[NR], const LSP& pt, float qg0, int lg) {
  float q0 = 0.f;
#pragma unroll
  for (int r = R0; r < NR; r++)
    if ((Z ? Jaref[r] : fmaf(pt.alpha, jv[r], Jaref[r])) < 0.f) q0 += a0[r];
  q0 = gall(q0, lg) + qg0;
  if (Z) return q0;
  return pt.alpha * pt.alpha * pt.q2 + pt.alpha * pt.q1 + q0;
}

// solver._linesearch: the exact line search along jv from Jaref; returns the step length. D are the rows' weights (0 on inactive
// rows), (qg0, qg1, qg2) the Gauss part of the cost along the search direction. The trip count is warp-uniform.
template <int NR, int R0> __device__ __forceinline__ float line_search(const float (&Jaref)[NR], const float (&jv)[NR], const float (&D)[NR], float qg0, float qg1,
                                                                     float qg2, float gtol, int ls_iterations, int lg, bool live) {
  float la0[NR], la1[NR], la2[NR];
#pragma unroll
  for (int r = R0; r < NR; r++) {
    const float ja = Jaref[r], w = jv[r], Dr = D[r];
    la0[r] = 0.5f * ja * ja * Dr; la1[r] = w * ja * Dr; la2[r] = 0.5f * w * w * Dr;
  }
#define LS_EVAL(al) ls_eval<NR, R0>(Jaref, jv, la1, la2, (al), qg1, qg2, lg)
  const LSP p0 = ls_eval<NR, R0, true>(Jaref, jv, la1, la2, 0.f, qg1, qg2, lg);
  const LSP l0 = LS_EVAL(-safe_div_fast(p0.d0, p0.d1));
  const bool lesser = l0.d0 < p0.d0;
  LSP hi = lesser ? p0 : l0;
  LSP lo = lesser ? l0 : p0;
  bool swap = true;
  int it = 0;
  while (true) {
    bool done = it >= ls_iterations;
    done = done || !swap;
    done = done || ((lo.d0 < 0.f) && (lo.d0 > -gtol));
    done = done || ((hi.d0 > 0.f) && (hi.d0 < gtol));
    if (!__any_sync(ABR_FULL, !done)) break;
    const LSP lo_next = LS_EVAL(lo.alpha - safe_div_fast(lo.d0, lo.d1));
    const LSP hi_next = LS_EVAL(hi.alpha - safe_div_fast(hi.d0, hi.d1));
    const LSP mid = LS_EVAL(0.5f * (lo.alpha + hi.alpha));
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
#undef LS_EVAL
  const float c_p0 = ls_cost<NR, R0, true>(Jaref, jv, la0, p0, qg0, lg), c_lo = ls_cost<NR, R0>(Jaref, jv, la0, lo, qg0, lg), c_hi = ls_cost<NR, R0>(Jaref, jv, la0, hi, qg0, lg);
  const bool improved = (c_lo < c_p0) || (c_hi < c_p0);
  return (improved && live) ? ((c_lo < c_hi) ? lo.alpha : hi.alpha) : 0.f;
}

// mjx.forward for one lane: on exit s.a = qacc, s.warm = qacc; M, fs, fc are returned for the
// implicit-damping Euler variant.
template <int NL, int NC, int LGC, bool CB, int SPEC>
__device__ __forceinline__ void forward(Lane<NL, NC>& s, const LaneCfg<LGC>& C, float (&M)[(6 + NL) * (7 + NL) / 2], float (&fs)[6 + NL], float (&fc)[6 + NL]) {
  constexpr int NP = NL + 1, N = 6 + NL, NTRI = N * (N + 1) / 2, NR = NL + 4 * NC;
  constexpr Map mp{NL, NC};
  const ShareT<LGC>& S = C.S;
  // ---------------------------------------------------------------- kinematics (smooth.kinematics)
  // One rotation matrix per body. Everything constant in the PARENT frame is folded into the table (joint anchor
  // body_pos + R(body_quat) jnt_pos, joint axis R(body_quat) jnt_axis, and body_quat o (0, jnt_axis), so that
  // body_quat o q_joint = cos(a/2) body_quat + sin(a/2) [body_quat o (0, axis)]); the body's inertia is carried as its
  // body-frame tensor, so the inertial-frame quaternion never has to be composed at run time.
  float xpos[NP][3], xquat[NP][4], xipos[NP][3], xanc[NP][3], xax[NP][3], irot[NP][6];
  float R0[9], Rp[9];
  {
    float q[4] = {s.qt[3], s.qt[4], s.qt[5], s.qt[6]};
    normalize_k<4>(q);
    s.qt[3] = q[0]; s.qt[4] = q[1]; s.qt[5] = q[2]; s.qt[6] = q[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { xpos[0][i] = s.qt[i]; xanc[0][i] = s.qt[i]; xax[0][i] = (i == 2) ? 1.f : 0.f; }
#pragma unroll
    for (int i = 0; i < 4; i++) xquat[0][i] = q[i];
    q_to_mat(q, R0);
#pragma unroll
    for (int i = 0; i < 9; i++) Rp[i] = R0[i];
    const float ip[3] = {LTF(mp.body(0) + 7), LTF(mp.body(0) + 8), LTF(mp.body(0) + 9)};
    const float Ib[6] = {LTF(mp.body(0) + 15), LTF(mp.body(0) + 16), LTF(mp.body(0) + 17), LTF(mp.body(0) + 18), LTF(mp.body(0) + 19), LTF(mp.body(0) + 20)};
    float r[3];
    m_rot(R0, ip, r);
#pragma unroll
    for (int i = 0; i < 3; i++) xipos[0][i] = xpos[0][i] + r[i];
    rot_inertia(R0, Ib, irot[0]);
  }
  int jflags[NP];
  jflags[0] = 0;
#pragma unroll
  for (int p = 1; p < NP; p++) {
    const int fl = LTI(mp.ijnt(p));
    jflags[p] = fl;
    const int type = fl & kJTypeMask;
    const float ca[3] = {LTF(mp.body(p)), LTF(mp.body(p) + 1), LTF(mp.body(p) + 2)};
    const float bq[4] = {LTF(mp.body(p) + 3), LTF(mp.body(p) + 4), LTF(mp.body(p) + 5), LTF(mp.body(p) + 6)};
    const float bj[4] = {LTF(mp.body(p) + 10), LTF(mp.body(p) + 11), LTF(mp.body(p) + 12), LTF(mp.body(p) + 13)};
    const float cx[3] = {LTF(mp.body(p) + 21), LTF(mp.body(p) + 22), LTF(mp.body(p) + 23)};
    const float jp[3] = {LTF(mp.jnt(p)), LTF(mp.jnt(p) + 1), LTF(mp.jnt(p) + 2)};
    float r[3], pos[3], axis[3], anchor[3];
    m_rot(Rp, ca, r);
#pragma unroll
    for (int i = 0; i < 3; i++) anchor[i] = xpos[p - 1][i] + r[i];
    m_rot(Rp, cx, axis);
    const float dq = s.qc[p - 1] - LTF(mp.jnt(p) + 6);
    // hinge: rotate about the joint axis and re-anchor; slide: translate along the axis
    float sn, cs, qloc[4], qn[4], Rn[9];
    sincos_bf(((Spec<SPEC>::hinges || type == kJHinge) ? dq : 0.f) * 0.5f, sn, cs);
#pragma unroll
    for (int i = 0; i < 4; i++) qloc[i] = fmaf(sn, bj[i], cs * bq[i]);
    q_mul(xquat[p - 1], qloc, qn);
    q_to_mat(qn, Rn);
    m_rot(Rn, jp, r);
    const float sl = (!Spec<SPEC>::hinges && type == kJSlide) ? dq : 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) pos[i] = anchor[i] - r[i] + axis[i] * sl;
#pragma unroll
    for (int i = 0; i < 3; i++) { xpos[p][i] = pos[i]; xanc[p][i] = anchor[i]; xax[p][i] = axis[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) xquat[p][i] = qn[i];
    const float ip[3] = {LTF(mp.body(p) + 7), LTF(mp.body(p) + 8), LTF(mp.body(p) + 9)};
    const float Ib[6] = {LTF(mp.body(p) + 15), LTF(mp.body(p) + 16), LTF(mp.body(p) + 17), LTF(mp.body(p) + 18), LTF(mp.body(p) + 19), LTF(mp.body(p) + 20)};
    m_rot(Rn, ip, r);
#pragma unroll
    for (int i = 0; i < 3; i++) xipos[p][i] = pos[i] + r[i];
    rot_inertia(Rn, Ib, irot[p]);
#pragma unroll
    for (int i = 0; i < 9; i++) Rp[i] = Rn[i];
  }
  // ---------------------------------------------------------------- com_pos: subtree CoM, cinert, cdof
  float com[3];
  {
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const float ms = S.o(p) ? LTF(mp.body(p) + 14) : 0.f;
      sx = fmaf(xipos[p][0], ms, sx); sy = fmaf(xipos[p][1], ms, sy); sz = fmaf(xipos[p][2], ms, sz);
    }
    sx = gall(sx, S.lg()); sy = gall(sy, S.lg()); sz = gall(sz, S.lg());
    const float im = C.inv_mass;
    com[0] = sx * im; com[1] = sy * im; com[2] = sz * im;
  }
  float cinert[NP][10];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    const float off[3] = {xipos[p][0] - com[0], xipos[p][1] - com[1], xipos[p][2] - com[2]};
    const float ms = LTF(mp.body(p) + 14);
    const float oo = v_dot(off, off);
    constexpr int ra[6] = {0, 1, 2, 0, 0, 1}, cb[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; e++) {
      const int r_ = ra[e], c_ = cb[e];
      cinert[p][e] = irot[p][e] + ms * ((r_ == c_ ? oo : 0.f) - off[r_] * off[c_]);
    }
    cinert[p][6] = off[0] * ms; cinert[p][7] = off[1] * ms; cinert[p][8] = off[2] * ms; cinert[p][9] = ms;
  }
  float cdof[N][6];
  {
    const float off[3] = {com[0] - xanc[0][0], com[1] - xanc[0][1], com[2] - xanc[0][2]};
#pragma unroll
    for (int q = 0; q < 3; q++) {
#pragma unroll
      for (int i = 0; i < 6; i++) cdof[q][i] = (i == 3 + q) ? 1.f : 0.f;
      const float ax[3] = {R0[q], R0[3 + q], R0[6 + q]};
      float cr[3];
      v_cross(ax, off, cr);
      cdof[3 + q][0] = ax[0]; cdof[3 + q][1] = ax[1]; cdof[3 + q][2] = ax[2];
      cdof[3 + q][3] = cr[0]; cdof[3 + q][4] = cr[1]; cdof[3 + q][5] = cr[2];
    }
  }
#pragma unroll
  for (int p = 1; p < NP; p++) {
    const int d = 5 + p;
    const bool hinge = Spec<SPEC>::hinges || (jflags[p] & kJTypeMask) == kJHinge;
    const float off[3] = {com[0] - xanc[p][0], com[1] - xanc[p][1], com[2] - xanc[p][2]};
    float cr[3];
    v_cross(xax[p], off, cr);
#pragma unroll
    for (int i = 0; i < 3; i++) { cdof[d][i] = hinge ? xax[p][i] : 0.f; cdof[d][3 + i] = hinge ? cr[i] : xax[p][i]; }
  }
  // ---------------------------------------------------------------- collision (plane - sphere) + Jacobian basis
  ABR_MIDSYNC();
  Rows<NL, NC, CB> R;
  float cdist[NC > 0 ? NC : 1];
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const int pc = LTI(mp.icon(c));
    {
      const bool pyr = LTI(mp.icon(c) + 1) == 3;
      R.mu1[c] = pyr ? C.frs * LTF(mp.con(c) + 4 + 11) : 0.f; R.mu2[c] = pyr ? C.frs * LTF(mp.con(c) + 4 + 12) : 0.f;
    }
    float bx[3] = {xpos[0][0], xpos[0][1], xpos[0][2]}, bq[4] = {xquat[0][0], xquat[0][1], xquat[0][2], xquat[0][3]};
#pragma unroll
    for (int p = 1; p < NP; p++) {
      if (pc == p) {
#pragma unroll
        for (int i = 0; i < 3; i++) bx[i] = xpos[p][i];
#pragma unroll
        for (int i = 0; i < 4; i++) bq[i] = xquat[p][i];
      }
    }
    const float gp[3] = {LTF(mp.con(c)), LTF(mp.con(c) + 1), LTF(mp.con(c) + 2)};
    const float rad = LTF(mp.con(c) + 3);
    const float n[3] = {LTF(mp.con(c) + 18), LTF(mp.con(c) + 19), LTF(mp.con(c) + 20)};
    const float pp[3] = {LTF(mp.con(c) + 21), LTF(mp.con(c) + 22), LTF(mp.con(c) + 23)};
    float r[3];
    q_rot(gp, bq, r);
    const float sp[3] = {bx[0] + r[0], bx[1] + r[1], bx[2] + r[2]};
    const float dd[3] = {sp[0] - pp[0], sp[1] - pp[1], sp[2] - pp[2]};
    const float dist = v_dot(dd, n) - rad;
    cdist[c] = dist;
    const float kk = rad + 0.5f * dist;
    const float off[3] = {sp[0] - n[0] * kk - com[0], sp[1] - n[1] * kk - com[1], sp[2] - n[2] * kk - com[2]};
    if constexpr (CB) {
      const float offb[3] = {r[0] - n[0] * kk, r[1] - n[1] * kk, r[2] - n[2] * kk};  // contact point - body origin
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float fr[3] = {LTF(mp.con(c) + 24 + 3 * k), LTF(mp.con(c) + 25 + 3 * k), LTF(mp.con(c) + 26 + 3 * k)};
        v_cross(offb, fr, R.g[c][k]);
        if (c == 0) { R.fr[k][0] = fr[0]; R.fr[k][1] = fr[1]; R.fr[k][2] = fr[2]; R.sref[k] = bx[k] - com[k]; }
      }
    } else {
      const int ld = (pc < 0) ? -1 : 5 + pc;
#pragma unroll
      for (int d = 0; d < N; d++) {
        float cr[3] = {0.f, 0.f, 0.f};
        if (d >= 3) v_cross(cdof[d], off, cr);  // a translational trunk dof moves the contact point by its own unit vector
        const float jp[3] = {cdof[d][3] + cr[0], cdof[d][4] + cr[1], cdof[d][5] + cr[2]};
        const bool on = d <= ld;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float fr[3] = {LTF(mp.con(c) + 24 + 3 * k), LTF(mp.con(c) + 25 + 3 * k), LTF(mp.con(c) + 26 + 3 * k)};
          R.B[c][k][d] = on ? ((d < 3) ? fr[d] : v_dot(fr, jp)) : 0.f;
        }
      }
    }
  }
  // ---------------------------------------------------------------- velocity pass: com_vel + rne + passive + actuation
  // Run BEFORE crb + M on the long-chain families (cinert is then dead when the factorisation needs its registers: biped +3 %) and
  // after the factorisation on the short-chain ones (Barkour class: the other order costs 2 %, profiles/r2_attempt_stage_order.txt).
  constexpr bool kVelFirst = NL >= 5;
  auto velocity_pass = [&]() {
    float cvel[NP][6], cfrc[NP][6], cdd[N][6];
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool grav = !(C.disableflags & ABR_DSBL_GRAVITY);
    float ca[6] = {0.f, 0.f, 0.f, grav ? -C.grav[0] : 0.f, grav ? -C.grav[1] : 0.f, grav ? -C.grav[2] : 0.f};
    // trunk (free joint): translations first, then rotations against the translated cvel
#pragma unroll
    for (int q = 0; q < 3; q++) {
      cv[3 + q] = s.v[q];
#pragma unroll
      for (int i = 0; i < 6; i++) cdd[q][i] = 0.f;
    }
    // cv is a pure translation here, so motion_cross(cv, cdof[q]) = [0 | v_lin x axis_q]: the products with cv's exact zeros
    // (which the compiler may not fold) are dropped, as are the updates of ca's angular part by them
#pragma unroll
    for (int q = 3; q < 6; q++) {
      cdd[q][0] = 0.f; cdd[q][1] = 0.f; cdd[q][2] = 0.f;
      v_cross(cv + 3, cdof[q], cdd[q] + 3);
    }
#pragma unroll
    for (int q = 3; q < 6; q++)
#pragma unroll
      for (int i = 0; i < 6; i++) cv[i] = fmaf(cdof[q][i], s.v[q], cv[i]);
#pragma unroll
    for (int q = 3; q < 6; q++)
#pragma unroll
      for (int i = 3; i < 6; i++) ca[i] = fmaf(cdd[q][i], s.v[q], ca[i]);
#pragma unroll
    for (int p = 0; p < NP; p++) {
      if (p > 0) {
        const int d = 5 + p;
        motion_cross(cv, cdof[d], cdd[d]);
        const float qd = s.v[d];
#pragma unroll
        for (int i = 0; i < 6; i++) { cv[i] = fmaf(cdof[d][i], qd, cv[i]); ca[i] = fmaf(cdd[d][i], qd, ca[i]); }
      }
#pragma unroll
      for (int i = 0; i < 6; i++) cvel[p][i] = cv[i];
      float f1[6], f2[6], f3[6];
      if (p == 0) inert_mul_lin(cinert[p], ca, f1); else inert_mul(cinert[p], ca, f1);  // the trunk's ca has no angular part
      inert_mul(cinert[p], cv, f2);
      motion_cross_force(cv, f2, f3);
#pragma unroll
      for (int i = 0; i < 6; i++) cfrc[p][i] = f1[i] + f3[i];
    }
    (void)cvel;
    float up[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool passive_on = !(C.disableflags & ABR_DSBL_PASSIVE);
    const bool act_on = !(C.disableflags & ABR_DSBL_ACTUATION);
#pragma unroll
    for (int p = NP - 1; p >= 0; p--) {
      if (S.m(p)) {
#pragma unroll
        for (int i = 0; i < 6; i++) up[i] = gs(up[i], S.l(p), S.m(p));
      }
      float f[6];
#pragma unroll
      for (int i = 0; i < 6; i++) f[i] = cfrc[p][i] + up[i];
#pragma unroll
      for (int d = 0; d < N; d++) {
        if (PD(d) != p) continue;
        const float bias = cdof_dot(cdof, d, f);
        float t = 0.f;
        if (p == 0) {
          if (passive_on) t = -C.dps * LTF(mp.trunk() + d) * s.v[d];
        } else {
          const int fl = jflags[p];
          const float q = s.qc[p - 1];
          if (passive_on) t = -LTF(mp.jnt(p) + 8) * (q - LTF(mp.jnt(p) + 7)) - C.dps * LTF(mp.jnt(p) + 9) * s.v[d];
          // actuator (transmission + fwd_actuation); prm: ctrlrange2 forcerange2 gainprm3 biasprm3 gear
          const float* prm = &LTF(mp.jnt(p) + 24);
          const int af = fl >> kJActShift;
          float ct = s.ctrl[p - 1];
          if ((af & 1) && !(C.disableflags & ABR_DSBL_CLAMPCTRL)) ct = fminf(fmaxf(ct, prm[0]), prm[1 * kLS]);
          const float gear = prm[10 * kLS];
          const float len = q * gear, vel = s.v[d] * gear;
          float gain = prm[4 * kLS];
          if (af & 4) gain += prm[5 * kLS] * len + prm[6 * kLS] * vel;
          float bs = 0.f;
          if (af & 8) bs = prm[7 * kLS] + prm[8 * kLS] * len + prm[9 * kLS] * vel;
          float af_ = (gain * ct + bs) * C.acs;
          if (af & 2) af_ = fminf(fmaxf(af_, prm[2 * kLS]), prm[3 * kLS]);
          af_ *= gear;
          t += ((fl & kJAct) && act_on) ? af_ : 0.f;
        }
        fs[d] = t - bias;
      }
      const bool ownp = S.o(p);
#pragma unroll
      for (int i = 0; i < 6; i++) up[i] = ownp ? f[i] : 0.f;
    }
  };
  if constexpr (kVelFirst) velocity_pass();
  // ---------------------------------------------------------------- crb + M (smooth.crb, support.make_m)
  ABR_MIDSYNC();
  {
    float crb[10], up[10];
#pragma unroll
    for (int i = 0; i < 10; i++) up[i] = 0.f;
#pragma unroll
    for (int p = NP - 1; p >= 0; p--) {
      if (S.m(p)) {
#pragma unroll
        for (int i = 0; i < 10; i++) up[i] = gs(up[i], S.l(p), S.m(p));
      }
#pragma unroll
      for (int i = 0; i < 10; i++) crb[i] = cinert[p][i] + up[i];
      // rows of M for the dofs of this position
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (PD(i) == p) {
          float buf[6];
          if (i < 3) inert_mul_unit(crb, i, buf); else inert_mul(crb, cdof[i], buf);
#pragma unroll
          for (int j = 0; j < N; j++) {
            if (j <= i) {
              float t = cdof_dot(cdof, j, buf);
              if (i == j) t += C.ars * ((i < 6) ? LTF(mp.trunk() + 6 + i) : LTF(mp.jnt(PD(i)) + 10));
              M[TR(i, j)] = t;
            }
          }
        }
      }
      const bool ownp = S.o(p);
#pragma unroll
      for (int i = 0; i < 10; i++) up[i] = ownp ? crb[i] : 0.f;
    }
  }
  // ---------------------------------------------------------------- factor M
  float F[NTRI], invD[N];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++)
      if (j <= i) F[TR(i, j)] = S.o(PD(i)) ? M[TR(i, j)] : 0.f;
  ldl_factor<N>(F, invD, S);
  if constexpr (!kVelFirst) velocity_pass();
  // ---------------------------------------------------------------- qacc_smooth
  float as[N];
#pragma unroll
  for (int i = 0; i < N; i++) as[i] = fs[i];
  ldl_solve<N>(F, invD, as, S);
  if (C.nefc == 0) {
#pragma unroll
    for (int i = 0; i < N; i++) { s.a[i] = as[i]; s.warm[i] = as[i]; fc[i] = 0.f; }
    return;
  }
  // ---------------------------------------------------------------- constraint rows (make_constraint)
  ABR_MIDSYNC();
  // joint limits: the impedance / reference acceleration of a row is only evaluated when some lane of the warp has an active
  // limit (warp-uniform branch); otherwise the rows are the exact zeros an inactive row gets anyway
  bool anylim;
  {
    float lpos[NL], lsgn[NL];
    bool lact[NL], any = false;
#pragma unroll
    for (int p = 1; p < NP; p++) {
      const float q = s.qc[p - 1];
      const float dmin = q - LTF(mp.jnt(p) + 11), dmax = LTF(mp.jnt(p) + 12) - q;
      lpos[p - 1] = fminf(dmin, dmax) - LTF(mp.jnt(p) + 13);
      lact[p - 1] = (lpos[p - 1] < 0.f) && (jflags[p] & kJLimited) && S.o(p);
      lsgn[p - 1] = (dmin < dmax) ? 1.f : -1.f;
      any = any || lact[p - 1];
    }
    anylim = __any_sync(ABR_FULL, any);
    if (anylim) {
#pragma unroll
      for (int p = 1; p < NP; p++) {
        const int d = 5 + p, r = p - 1;
        R.lsg[r] = lact[r] ? lsgn[r] : 0.f;
        row_kbi<Spec<SPEC>::pow2>(&LTF(mp.jnt(p) + 14), lpos[r], lsgn[r] * s.v[d], LTF(mp.jnt(p) + 14 + 7), lact[r], R.D[r], R.aref[r]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < NL; r++) { R.lsg[r] = 0.f; R.D[r] = 0.f; R.aref[r] = 0.f; }
    }
  }
  {
    float bva[NC > 0 ? NC : 1][3];
    if constexpr (CB) {
      contact_bv<NL, NC>(R, cdof, s.v, bva);
    } else {
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
          float t = 0.f;
#pragma unroll
          for (int d = 0; d < N; d++) t = fmaf(R.B[c][k][d], s.v[d], t);
          bva[c][k] = t;
        }
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const float* prm = &LTF(mp.con(c) + 4);
      const float pos = cdist[c] - prm[13 * kLS];
      const bool act0 = (pos < 0.f) && (LTI(mp.icon(c)) >= 0);
      const bool pyr = LTI(mp.icon(c) + 1) == 3;
      const Kbi kq = kbi_imp<Spec<SPEC>::pow2>(prm, pos);
#pragma unroll
      for (int sub = 0; sub < 4; sub++) {
        const int r = NL + 4 * c + sub;
        const float mu = (sub < 2) ? R.mu1[c] : R.mu2[c];
        const float jvel = bva[c][0] + ((sub & 1) ? -mu : mu) * bva[c][1 + (sub >> 1)];
        float invw = (sub < 2 || !pyr) ? prm[7 * kLS] : prm[10 * kLS];
        if (pyr && C.frs != 1.f) {  // the pyramid's invweight is t (1 + mu^2) 2 mu^2 / impratio: rescale it with the friction
          const float mu0 = prm[(sub < 2 ? 11 : 12) * kLS];
          invw *= C.frs * C.frs * (1.f + mu * mu) / (1.f + mu0 * mu0);
        }
        kbi_row(kq, jvel, invw, act0 && (pyr || sub == 0), R.D[r], R.aref[r]);
      }
    }
  }
  // ---------------------------------------------------------------- solver.solve (Newton)
  ABR_MIDSYNC();
  float Ma[N], Jaref[NR];
  float gauss, cost;
  // M qacc_smooth is qfrc_smooth itself (as = M^-1 fs): MJX multiplies it out again and gets fs back up to rounding
#pragma unroll
  for (int d = 0; d < N; d++) Ma[d] = fs[d];
  mul_j<NL, NC, CB>(R, cdof, as, Jaref);
#pragma unroll
  for (int r = 0; r < NR; r++) Jaref[r] -= R.aref[r];
  cost = solver_cost<NL, NC, CB, ShareT<LGC>, true>(R, S, as, Ma, Jaref, fs, as, gauss);
#pragma unroll
  for (int d = 0; d < N; d++) s.a[d] = as[d];
  if (!(C.disableflags & ABR_DSBL_WARMSTART)) {
    float Mw[N], Jw[NR], g2;
    mul_m<N>(M, s.warm, Mw, S);
    mul_j<NL, NC, CB>(R, cdof, s.warm, Jw);
#pragma unroll
    for (int r = 0; r < NR; r++) Jw[r] -= R.aref[r];
    const float c2 = solver_cost<NL, NC, CB>(R, S, s.warm, Mw, Jw, fs, as, g2);
    const bool use = c2 < cost;
    cost = use ? c2 : cost; gauss = use ? g2 : gauss;
#pragma unroll
    for (int d = 0; d < N; d++) { s.a[d] = use ? s.warm[d] : s.a[d]; Ma[d] = use ? Mw[d] : Ma[d]; }
#pragma unroll
    for (int r = 0; r < NR; r++) Jaref[r] = use ? Jw[r] : Jaref[r];
  }
  float prev_cost = INFINITY;
  const float scale = 1.f / (C.meaninertia * (float)max(1, C.nv));
  bool live = true;
  const bool need_fc = Spec<SPEC>::edamp(C.disableflags);  // qfrc_constraint of the final point is only read by the implicit-damping Euler step
  for (int niter = 0;; niter++) {
    // efc_force and qfrc_constraint at the current point (skipped at the final point when nothing reads it)
    if (niter < C.iterations || need_fc) {
      float up[N];
#pragma unroll
      for (int d = 0; d < N; d++) up[d] = 0.f;
#pragma unroll
      for (int r = 0; r < NL; r++) { const float f = (Jaref[r] < 0.f) ? -R.D[r] * Jaref[r] : 0.f; up[6 + r] = R.lsg[r] * f; }
      float Wt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, Fl[3] = {0.f, 0.f, 0.f};  // CB: total wrench on the leaf body
#pragma unroll
      for (int c = 0; c < NC; c++) {
        float f[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const int r = NL + 4 * c + k; f[k] = (Jaref[r] < 0.f) ? -R.D[r] * Jaref[r] : 0.f; }
        const float mu1 = R.mu1[c], mu2 = R.mu2[c];
        const float F[3] = {f[0] + f[1] + f[2] + f[3], mu1 * (f[0] - f[1]), mu2 * (f[2] - f[3])};
        if constexpr (CB) {
#pragma unroll
          for (int k = 0; k < 3; k++) {
            Fl[k] += F[k];
#pragma unroll
            for (int i = 0; i < 3; i++) Wt[i] = fmaf(R.g[c][k][i], F[k], Wt[i]);
          }
        } else {
#pragma unroll
          for (int d = 0; d < N; d++) up[d] = fmaf(R.B[c][2][d], F[2], fmaf(R.B[c][1][d], F[1], fmaf(R.B[c][0][d], F[0], up[d])));
        }
      }
      if constexpr (CB && NC > 0) {
#pragma unroll
        for (int i = 0; i < 3; i++) Wt[3 + i] = R.fr[0][i] * Fl[0] + R.fr[1][i] * Fl[1] + R.fr[2][i] * Fl[2];
        {
          float sh[3];
          v_cross(R.sref, Wt + 3, sh);  // torque about the CoM = torque about the body origin + sref x force
#pragma unroll
          for (int i = 0; i < 3; i++) Wt[i] += sh[i];
        }
#pragma unroll
        for (int d = 0; d < N; d++)
          up[d] += cdof_dot(cdof, d, Wt);
      }
#pragma unroll
      for (int d = 0; d < N; d++) {
        const int p = PD(d);
        fc[d] = S.m(p) ? gs(up[d], S.l(p), S.m(p)) : up[d];
      }
    }
    if (niter >= C.iterations) break;
    float grad[N];
#pragma unroll
    for (int d = 0; d < N; d++) grad[d] = Ma[d] - fs[d] - fc[d];
    // Hessian H = M + J' diag(D active) J in share form, then L'DL
    float H[NTRI], hD[N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++)
        if (j <= i) H[TR(i, j)] = S.o(PD(i)) ? M[TR(i, j)] : 0.f;
#pragma unroll
    for (int r = 0; r < NL; r++)
      if (Jaref[r] < 0.f) H[TR(6 + r, 6 + r)] = fmaf(R.D[r] * R.lsg[r], R.lsg[r], H[TR(6 + r, 6 + r)]);
    float K[21];  // CB: 6x6 weight of the leaf body's twist, sum_c [g_c; fr] W_c [g_c; fr]'
    if constexpr (CB) {
#pragma unroll
      for (int e = 0; e < 21; e++) K[e] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < NC; c++) {
      const int r0 = NL + 4 * c;
      const float mu1 = R.mu1[c], mu2 = R.mu2[c];
      const float w0 = (Jaref[r0] < 0.f) ? R.D[r0] : 0.f, w1 = (Jaref[r0 + 1] < 0.f) ? R.D[r0 + 1] : 0.f;
      const float w2 = (Jaref[r0 + 2] < 0.f) ? R.D[r0 + 2] : 0.f, w3 = (Jaref[r0 + 3] < 0.f) ? R.D[r0 + 3] : 0.f;
      const float W00 = w0 + w1 + w2 + w3, W01 = mu1 * (w0 - w1), W02 = mu2 * (w2 - w3);
      const float W11 = mu1 * mu1 * (w0 + w1), W22 = mu2 * mu2 * (w2 + w3);
      if constexpr (CB) {
        float w[3][6], U[3][6];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
          for (int i = 0; i < 3; i++) { w[k][i] = R.g[c][k][i]; w[k][3 + i] = R.fr[k][i]; }
#pragma unroll
        for (int i = 0; i < 6; i++) {
          U[0][i] = W00 * w[0][i] + W01 * w[1][i] + W02 * w[2][i];
          U[1][i] = W01 * w[0][i] + W11 * w[1][i];
          U[2][i] = W02 * w[0][i] + W22 * w[2][i];
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
          for (int j = 0; j < 6; j++)
            if (j <= i) K[TR(i, j)] = fmaf(w[2][i], U[2][j], fmaf(w[1][i], U[1][j], fmaf(w[0][i], U[0][j], K[TR(i, j)])));
      } else {
#pragma unroll
        for (int j = 0; j < N; j++) {
          const float b0 = R.B[c][0][j], b1 = R.B[c][1][j], b2 = R.B[c][2][j];
          const float o0 = W00 * b0 + W01 * b1 + W02 * b2, o1 = W01 * b0 + W11 * b1, o2 = W02 * b0 + W22 * b2;
#pragma unroll
          for (int i = 0; i < N; i++)
            if (i >= j) H[TR(i, j)] = fmaf(R.B[c][2][i], o2, fmaf(R.B[c][1][i], o1, fmaf(R.B[c][0][i], o0, H[TR(i, j)])));
        }
      }
    }
    if constexpr (CB && NC > 0) {
#pragma unroll
      for (int j = 0; j < N; j++) {
        float cj[6], KC[6], sh[3] = {0.f, 0.f, 0.f};
        if (j >= 3) v_cross(cdof[j], R.sref, sh);  // dof j's twist at the leaf body's origin (a pure translation does not change)
#pragma unroll
        for (int b = 0; b < 3; b++) { cj[b] = cdof[j][b]; cj[3 + b] = cdof[j][3 + b] + sh[b]; }
#pragma unroll
        for (int a = 0; a < 6; a++) {
          float t = 0.f;
#pragma unroll
          for (int b = 0; b < 6; b++) t = fmaf(K[a >= b ? TR(a, b) : TR(b, a)], cj[b], t);
          KC[a] = (j < 3) ? K[a >= 3 + j ? TR(a, 3 + j) : TR(3 + j, a)] : t;
        }
        v_cross(R.sref, KC + 3, sh);  // back to the CoM reference: the "torque" part picks up sref x "force" part
#pragma unroll
        for (int b = 0; b < 3; b++) KC[b] += sh[b];
#pragma unroll
        for (int i = 0; i < N; i++)
          if (i >= j) H[TR(i, j)] += cdof_dot(cdof, i, KC);
      }
    }
    ldl_factor<N>(H, hD, S);
    float mg[N];
#pragma unroll
    for (int d = 0; d < N; d++) mg[d] = grad[d];
    ldl_solve<N>(H, hD, mg, S);
    if (Spec<SPEC>::multi(C.iterations)) {
      float gn = 0.f;
#pragma unroll
      for (int d = 0; d < N; d++) gn += S.o(PD(d)) ? grad[d] * grad[d] : 0.f;
      gn = gall(gn, S.lg());
      bool done = scale * (prev_cost - cost) < C.tol;
      done = done || (scale * sqrtf(gn) < C.tol);
      live = live && !done;
      if (!__any_sync(ABR_FULL, live)) break;
    }
    float search[N], mv[N], jv[NR];
#pragma unroll
    for (int d = 0; d < N; d++) search[d] = -mg[d];
    // ---- exact line search (solver._linesearch)
    mul_m<N>(M, search, mv, S);
    mul_j<NL, NC, CB>(R, cdof, search, jv);
    float sn = 0.f, sMa = 0.f, sq = 0.f, smv = 0.f;
#pragma unroll
    for (int d = 0; d < N; d++) {
      if (S.o(PD(d))) { const float t = search[d]; sn += t * t; sMa += t * Ma[d]; sq += t * fs[d]; smv += t * mv[d]; }
    }
    sn = gall(sn, S.lg()); sMa = gall(sMa, S.lg()); sq = gall(sq, S.lg()); smv = gall(smv, S.lg());
    const float smag = sqrt_fast(sn) * C.meaninertia * (float)max(1, C.nv);
    const float gtol = C.tol * C.ls_tol * smag;
    const float qg0 = gauss, qg1 = sMa - sq, qg2 = 0.5f * smv;
    // joint-limit rows are rarely active: when no lane of the warp has one, the line search runs over the contact rows only
    const float alpha = anylim ? line_search<NR, 0>(Jaref, jv, R.D, qg0, qg1, qg2, gtol, C.ls_iterations, S.lg(), live)
                               : line_search<NR, (NL < NR ? NL : 0)>(Jaref, jv, R.D, qg0, qg1, qg2, gtol, C.ls_iterations, S.lg(), live);
#pragma unroll
    for (int d = 0; d < N; d++) { s.a[d] = fmaf(search[d], alpha, s.a[d]); Ma[d] = fmaf(mv[d], alpha, Ma[d]); }
#pragma unroll
    for (int r = 0; r < NR; r++) Jaref[r] = fmaf(jv[r], alpha, Jaref[r]);
    if (Spec<SPEC>::multi(C.iterations)) {
      float g2;
      const float c2 = solver_cost<NL, NC, CB>(R, S, s.a, Ma, Jaref, fs, as, g2);
      if (live) { prev_cost = cost; cost = c2; gauss = g2; }
    }
  }
#pragma unroll
  for (int d = 0; d < N; d++) s.warm[d] = s.a[d];
}

// forward.euler (+ implicit joint damping unless EULERDAMP is disabled) and position integration
template <int NL, int NC, int LGC, int SPEC>
__device__ __forceinline__ void euler(Lane<NL, NC>& s, const LaneCfg<LGC>& C, float (&M)[(6 + NL) * (7 + NL) / 2], const float (&fs)[6 + NL], const float (&fc)[6 + NL]) {
  constexpr int N = 6 + NL;
  constexpr Map mp{NL, NC};
  const float dt = C.dt;
  if (Spec<SPEC>::edamp(C.disableflags)) {
    float hD[N], rhs[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      const float damp = C.dps * ((i < 6) ? LTF(mp.trunk() + i) : LTF(mp.jnt(PD(i)) + 9));
      M[TR(i, i)] += damp * dt;
      rhs[i] = fs[i] + fc[i];
    }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++)
        if (j <= i) M[TR(i, j)] = C.S.o(PD(i)) ? M[TR(i, j)] : 0.f;
    ldl_factor<N>(M, hD, C.S);
    ldl_solve<N>(M, hD, rhs, C.S);
#pragma unroll
    for (int i = 0; i < N; i++) s.a[i] = rhs[i];
  }
#pragma unroll
  for (int d = 0; d < N; d++) s.v[d] = fmaf(s.a[d], dt, s.v[d]);
  // free joint
  s.qt[0] = fmaf(dt, s.v[0], s.qt[0]); s.qt[1] = fmaf(dt, s.v[1], s.qt[1]); s.qt[2] = fmaf(dt, s.v[2], s.qt[2]);
  float w[3] = {s.v[3], s.v[4], s.v[5]};
  const float nrm = normalize_k<3>(w);
  float ql[4], qn[4];
  axis_angle_quat_bf(w, dt * nrm, ql);
  q_mul(s.qt + 3, ql, qn);
  normalize_k<4>(qn);
  s.qt[3] = qn[0]; s.qt[4] = qn[1]; s.qt[5] = qn[2]; s.qt[6] = qn[3];
#pragma unroll
  for (int p = 1; p <= NL; p++) s.qc[p - 1] = fmaf(dt, s.v[5 + p], s.qc[p - 1]);
}

// ------------------------------------------------------------------------------ kernels
constexpr int kTPB = 32;      // smallest CTA: one warp (32/G worlds)
#ifndef ABR_LIMB_MAXTPB
#define ABR_LIMB_MAXTPB 256
#endif
constexpr int kMaxTPB = ABR_LIMB_MAXTPB;  // largest CTA: 8 warps = every register of an SM at 255 registers per thread
// CTA size of a launch ("instruction delivery"). The straight-line step does not fit the SM's instruction
// cache, so every SM streams it from the GPC-level cache each step; the warps of one SM run in loose lockstep and share that
// stream, while the SMs of a GPC compete for it. A small batch therefore runs faster on FEWER SMs with MORE warps each:
// the per-GPC instruction traffic drops with the number of streaming SMs. ABR_LIMB_TPB overrides the policy (probes).
// The long-chain families (NL >= 6: the step spills and streams > 128 KB of code) run a wave of 7- / 6-warp CTAs in 0.955 / 0.94
// of the time of a wave of 8-warp CTAs (one CTA per SM at 255 registers), so their launches take the CTA size that minimises
// waves x time per wave. The model reproduces every point of profiles/r2_tpb_sweep3.txt to 2 % (C3: 2048 warps = 256 CTAs of 8
// warps or 293 of 7, two waves either way: +4.4 %; 8192 warps: 7 waves of 8 beat 8 waves of 7). The Barkour class does not
// follow it (its second, partial wave runs faster than a full one) and keeps 8 warps beyond 518 warps (7 below: see the end of pick_tpb).
constexpr int kNumSM = 148;  // B200
inline int pick_tpb(long nwarps, int nl = 0) {
  const char* ev = getenv("ABR_LIMB_TPB");
  const int env = ev ? atoi(ev) : 0;
  if (env >= 32 && env <= kMaxTPB && env % 32 == 0) return env;
  if (nwarps <= 16) return 128;  // measured: profiles/r2_tpb_sweep*.txt (a handful of warps: one per sub-partition)
  if (nl >= 6 && kMaxTPB == 256) {
    const int wpc[3] = {8, 7, 6};
    const double per_wave[3] = {1.0, 0.955, 0.94};
    int best = 8; double best_t = 1e30;
    for (int i = 0; i < 3; i++) {
      const long ctas = (nwarps + wpc[i] - 1) / wpc[i];
      const double t = (double)((ctas + kNumSM - 1) / kNumSM) * per_wave[i];
      if (t < best_t - 1e-9) { best_t = t; best = wpc[i]; }
    }
    return best * 32;
  }
  // short-chain families, a single wave: 7-warp CTAs run 1 - 1.8 % faster than 8-warp ones (profiles/r2_tpb_sweep4.txt: 2048 / 3072 /
  // 4096 worlds of the Barkour class; from 8192 worlds = 1024 warps on, 8 warps win again)
  if (kMaxTPB == 256 && nwarps <= 7L * kNumSM / 2) return 224;
  return kMaxTPB;
}
#ifndef ABR_LIMB_MINB
#define ABR_LIMB_MINB 1  // resident CTAs per SM the register allocation must allow (1 = up to 255 registers)
#endif

template <int NL, int NC, int LGC, int SPEC = -1> __device__ __forceinline__ LaneCfg<LGC> make_cfg(const Layout& L, const float* T, int g) {
  constexpr Map mp{NL, NC};
  LaneCfg<LGC> C;
  C.T = (const float*)__builtin_assume_aligned(T + g * tpad(mp.total()), 16);
  C.S.own = LTI(mp.ish()); C.S.lvl = LTI(mp.ish() + 1); C.S.mx = L.l_mx; C.S.lg_ = L.lg2G;
  C.dt = L.timestep; C.grav[0] = L.gravity[0]; C.grav[1] = L.gravity[1]; C.grav[2] = L.gravity[2];
  C.frs = 1.f; C.acs = 1.f; C.dps = 1.f; C.ars = 1.f;
  C.mass = L.l_mass; C.inv_mass = 1.f / L.l_mass; C.tol = L.tolerance; C.ls_tol = L.ls_tolerance; C.meaninertia = L.meaninertia;
  C.iterations = Spec<SPEC>::gen ? L.iterations : 1; C.ls_iterations = L.ls_iterations; C.disableflags = Spec<SPEC>::flags(L.disableflags); C.nefc = L.nefc; C.nv = L.nv;
  return C;
}

// this lane's share of sum_i w_i (x_i - xg_i)^2 over the state entries it owns (cost.py:62-85, diagonal Q)
template <int NL, int NC, int LGC> __device__ __forceinline__ float quad_x_diag(const Lane<NL, NC>& s, const LaneCfg<LGC>& C, const float* w, const float* xg, int nq) {
  constexpr Map mp{NL, NC};
  float acc = 0.f;
  if (C.S.o(0)) {
#pragma unroll
    for (int i = 0; i < 7; i++) { const float e = s.qt[i] - xg[i]; acc = fmaf(w[i] * e, e, acc); }
#pragma unroll
    for (int i = 0; i < 6; i++) { const float e = s.v[i] - xg[nq + i]; acc = fmaf(w[nq + i] * e, e, acc); }
  }
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2);
    if (gd >= 0 && C.S.o(p)) {
      const float e = s.qc[p - 1] - xg[gq]; acc = fmaf(w[gq] * e, e, acc);
      const float e2 = s.v[5 + p] - xg[nq + gd]; acc = fmaf(w[nq + gd] * e2, e2, acc);
    }
  }
  return acc;
}
// The same sum from a per-lane table in shared memory built once per launch (rollout kernel): slot k of lane l holds
// {running weight, terminal weight, goal} at ct[(3 k + {0,1,2}) * 32 + l] with zero weights where the lane does not own the
// entry, so the per-step evaluation is two loads with immediate offsets + three FP instructions per state entry.
// Slots: 0..6 trunk qpos, 7..12 trunk qvel, then (q, v) per chain position; after them one control weight per position.
template <int NL> __device__ __forceinline__ constexpr int cost_slots() { return 13 + 2 * NL; }
template <int NL, int NC> __device__ __forceinline__ float quad_x_tab(const Lane<NL, NC>& s, const float* ct /* + lane + (terminal ? 32 : 0) */, const float* cg /* goal column + lane */) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 7; i++) { const float e = s.qt[i] - cg[(3 * i) * 32]; acc = fmaf(ct[(3 * i) * 32] * e, e, acc); }
#pragma unroll
  for (int i = 0; i < 6; i++) { const float e = s.v[i] - cg[(3 * (7 + i)) * 32]; acc = fmaf(ct[(3 * (7 + i)) * 32] * e, e, acc); }
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    const int k = 13 + 2 * (p - 1);
    const float e = s.qc[p - 1] - cg[(3 * k) * 32]; acc = fmaf(ct[(3 * k) * 32] * e, e, acc);
    const float e2 = s.v[5 + p] - cg[(3 * (k + 1)) * 32]; acc = fmaf(ct[(3 * (k + 1)) * 32] * e2, e2, acc);
  }
  return acc;
}
template <int NL, int NC, int LGC> __device__ __forceinline__ void store_x(const Lane<NL, NC>& s, const LaneCfg<LGC>& C, float* x, int nq) {
  constexpr Map mp{NL, NC};
  if (C.S.o(0)) {
#pragma unroll
    for (int i = 0; i < 7; i++) x[i] = s.qt[i];
#pragma unroll
    for (int i = 0; i < 6; i++) x[nq + i] = s.v[i];
  }
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2);
    if (gd >= 0 && C.S.o(p)) { x[gq] = s.qc[p - 1]; x[nq + gd] = s.v[5 + p]; }
  }
}

// shoot (shooting.py:22-48) / the sampler's rollouts (shooting.py:140-153) on the limb path
template <int NL, int NC, int LGC, bool CB, int SPEC>
__global__ void __launch_bounds__(kMaxTPB, ABR_LIMB_MINB) k_limb_rollout(const __grid_constant__ Layout L, const __grid_constant__ RolloutArgs A) {
  extern __shared__ __align__(16) float smem[];
  constexpr Map mp{NL, NC};
  constexpr int N = 6 + NL, NTRI = N * (N + 1) / 2;
  const int ntab = mp.total() * kStride;
  const int nx = L.nx, nu = L.nu, nq = L.nq, Nh = A.N;
  constexpr int TOTP = tpad(mp.total());
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) smem[(i % kStride) * TOTP + i / kStride] = A.blob[L.f_ltab + i];
  float* cqd = smem + TOTP * kStride; float* cqf = cqd + nx; float* crd = cqf + nx; float* cxg = crd + nu;
  if (A.cost.enabled) {
    for (int i = threadIdx.x; i < nx; i += blockDim.x) { cqd[i] = A.cost.qd[i]; cqf[i] = A.cost.qfd[i]; cxg[i] = A.cost.xg[i]; }
    for (int i = threadIdx.x; i < nu; i += blockDim.x) crd[i] = A.cost.rd[i];
  }
  __syncthreads();
#ifdef ABR_PROBE_SKEW  // probe: the second warp of each sub-partition starts ABR_PROBE_SKEW cycles late (one-off, no per-step cost)
  if ((threadIdx.x >> 5) & 4) { const long long t0 = clock64(); while (clock64() - t0 < (long long)(ABR_PROBE_SKEW)) {} }
#endif
  const int lg = (LGC >= 0) ? (LGC & 3) : L.lg2G;
  const int g = threadIdx.x & ((1 << lg) - 1);
  const int wraw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> lg);
  const bool valid = wraw < A.nworld;
  const int w = valid ? wraw : A.nworld - 1;
  const LaneCfg<LGC> C = make_cfg<NL, NC, LGC, SPEC>(L, smem, g);
  int prob = w, sample = 0;
  if (Spec<SPEC>::sampler(A.mode)) {
    if (A.sample_ids) { prob = w; sample = A.sample_ids[w]; }
    else { prob = w / A.S; sample = A.sample_offset + (w - prob * A.S); }
  }
  const int t_first = A.t_begin, t_last = A.t_end;
  const bool resume = t_first > 0;
  const int cstride = nq + 2 * L.nv + 8;
  float* carry = A.carry ? A.carry + (size_t)w * cstride : nullptr;
  // a resumed slice reads the state its predecessor left in `carry` (same layout as x0, then qacc_warmstart)
  const float* x0 = resume ? carry : A.x0 + (size_t)(Spec<SPEC>::sampler(A.mode) ? prob : w) * A.x0_stride;
  Lane<NL, NC> s;
#pragma unroll
  for (int i = 0; i < 7; i++) s.qt[i] = x0[i];
#pragma unroll
  for (int i = 0; i < 6; i++) { s.v[i] = x0[nq + i]; s.warm[i] = resume ? x0[nx + i] : 0.f; s.a[i] = 0.f; }
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2);
    s.qc[p - 1] = (gd >= 0) ? x0[gq] : 0.f;
    s.v[5 + p] = (gd >= 0) ? x0[nq + gd] : 0.f;
    s.warm[5 + p] = (gd >= 0 && resume) ? x0[nx + gd] : 0.f;
    s.ctrl[p - 1] = 0.f;
    s.a[5 + p] = 0.f;
  }
  float* xs = Spec<SPEC>::out(A.xs_out) ? A.xs_out + (size_t)w * (Nh + 1) * nx : nullptr;
  // per-lane cost table (see quad_x_tab): weights of entries this lane does not own are zero
  constexpr int NS = cost_slots<NL>();
  float* ctab = cxg + nx + (threadIdx.x >> 5) * ((3 * NS + NL) * 32) + (threadIdx.x & 31);  // one table per warp
  if (A.cost.enabled) {
    const bool own0 = C.S.o(0);
#pragma unroll
    for (int i = 0; i < 13; i++) {
      const int gi = (i < 7) ? i : nq + (i - 7);
      ctab[(3 * i) * 32] = own0 ? cqd[gi] : 0.f; ctab[(3 * i + 1) * 32] = own0 ? cqf[gi] : 0.f; ctab[(3 * i + 2) * 32] = cxg[gi];
    }
#pragma unroll
    for (int p = 1; p <= NL; p++) {
      const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2), ga = LTI(mp.ijnt(p) + 3);
      const bool own = gd >= 0 && C.S.o(p);
      const int k = 13 + 2 * (p - 1);
      ctab[(3 * k) * 32] = own ? cqd[gq] : 0.f; ctab[(3 * k + 1) * 32] = own ? cqf[gq] : 0.f; ctab[(3 * k + 2) * 32] = own ? cxg[gq] : 0.f;
      ctab[(3 * k + 3) * 32] = own ? cqd[nq + gd] : 0.f; ctab[(3 * k + 4) * 32] = own ? cqf[nq + gd] : 0.f; ctab[(3 * k + 5) * 32] = own ? cxg[nq + gd] : 0.f;
      ctab[(3 * NS + (p - 1)) * 32] = (ga >= 0 && C.S.o(p)) ? crd[ga] : 0.f;
    }
  }
  const float* cgoal = ctab + 2 * 32;
  // explicit controls are fetched one step ahead with cp.async into a two-deep per-thread slot (no registers held across the
  // step, no exposed L2 / HBM latency at the top of it): word (buffer b, position p) of a thread at upre[(b * NL + p - 1) * blockDim.x]
  float* upre = cxg + nx + (blockDim.x >> 5) * ((3 * NS + NL) * 32) + threadIdx.x;
  const bool prefetch = !Spec<SPEC>::sampler(A.mode);
  auto fetch_u = [&](int t) {
#pragma unroll
    for (int p = 1; p <= NL; p++) {
      const int ga = LTI(mp.ijnt(p) + 3);
      if (ga >= 0) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(upre + ((t & 1) * NL + (p - 1)) * blockDim.x);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(A.us + (size_t)w * A.us_stride + (size_t)t * nu + ga) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (prefetch && (resume ? t_first : 0) < t_last) fetch_u(resume ? t_first : 0);
  float cacc = 0.f;
  if (!resume) {
    if (xs && valid) store_x<NL, NC, LGC>(s, C, xs, nq);
    if (A.cost.enabled) cacc += quad_x_tab<NL, NC>(s, ctab + (Nh > 0 ? 0 : 32), cgoal);
  } else {
    cacc = carry[nx + L.nv + g];
  }
  // t = -1 is mjx.forward with ctrl = 0, which seeds qacc_warmstart (shooting.py:36)
#pragma unroll 1
  for (int t = resume ? t_first : -1; t < t_last; t++) {
    // One CTA barrier per step re-aligns the warps of the SM: they share one instruction stream the better the tighter their
    // lockstep (profiles/r2_stepsync.txt: +1.5 % at 4096 worlds, +1 % at 65 536; a half-step skew costs 22 %, barriers at four
    // more points inside the step cost more than they return). Every thread of the CTA runs every step (invalid worlds are clamped).
    // Short-chain families only: the long-chain ones (their step is beyond the 128 KB instruction-delivery tier) lose 2 - 3 % to it.
#ifndef ABR_NO_STEPSYNC
    if constexpr (NL < 5) __syncthreads();
#endif
    if (t >= 0) {
      if (prefetch) asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int p = 1; p <= NL; p++) {
        const int ga = LTI(mp.ijnt(p) + 3);
        float u = 0.f;
        if (ga >= 0) {
          if (!Spec<SPEC>::sampler(A.mode)) {
            u = upre[((t & 1) * NL + (p - 1)) * blockDim.x];
          } else {
            float nz = 0.f;
            if (sample > 0) {
              if (A.noise) nz = A.noise[(((size_t)prob * (A.S_total - 1) + (sample - 1)) * Nh + t) * nu + ga];
              else nz = philox_normal_ool(A.seed, (uint32_t)sample, (uint32_t)prob, (uint32_t)(t * nu + ga));
            }
            const float vv = A.us[(size_t)prob * A.us_stride + (size_t)t * nu + ga] + nz * A.stdev;
            u = fminf(fmaxf(vv, LTF(mp.jnt(p) + 24)), LTF(mp.jnt(p) + 25));  // clip to actuator_ctrlrange (shooting.py:146-148)
          }
          if (C.S.o(p)) {
            if (Spec<SPEC>::out(A.us_out) && valid) A.us_out[((size_t)w * Nh + t) * nu + ga] = u;
          }
        }
        s.ctrl[p - 1] = u;
        if (A.cost.enabled) cacc = fmaf(ctab[(3 * NS + (p - 1)) * 32] * u, u, cacc);
      }
      if (prefetch && t + 1 < t_last) fetch_u(t + 1);
    }
    float M[NTRI], fs[N], fc[N];
    forward<NL, NC, LGC, CB, SPEC>(s, C, M, fs, fc);
    if (t >= 0) {
      euler<NL, NC, LGC, SPEC>(s, C, M, fs, fc);
      if (xs && valid) store_x<NL, NC, LGC>(s, C, xs + (size_t)(t + 1) * nx, nq);
      if (A.cost.enabled) cacc += quad_x_tab<NL, NC>(s, ctab + ((t == Nh - 1) ? 32 : 0), cgoal);
    }
  }
  if (t_last < Nh) {  // hand the state to the next slice
    if (valid && carry) {
      store_x<NL, NC, LGC>(s, C, carry, nq);
      if (C.S.o(0)) {
#pragma unroll
        for (int i = 0; i < 6; i++) carry[nx + i] = s.warm[i];
      }
#pragma unroll
      for (int p = 1; p <= NL; p++) {
        const int gd = LTI(mp.ijnt(p) + 1);
        if (gd >= 0 && C.S.o(p)) carry[nx + gd] = s.warm[5 + p];
      }
      carry[nx + L.nv + g] = cacc;
    }
    return;
  }
  if (A.costs_out) {
    cacc = gall(cacc, lg);
    if (valid && g == 0) A.costs_out[w] = 0.5f * cacc;
  }
}

// MjxEnv.pipeline_init / pipeline_step (rl/base.py:81-96) with the auto-reset blend, on the limb path
template <int NL, int NC, int LGC, bool CB, int SPEC>
__global__ void __launch_bounds__(kMaxTPB, ABR_LIMB_MINB) k_limb_env(const __grid_constant__ Layout L, const __grid_constant__ EnvArgs A) {
  extern __shared__ __align__(16) float smem[];
  constexpr Map mp{NL, NC};
  constexpr int N = 6 + NL, NTRI = N * (N + 1) / 2;
  const int ntab = mp.total() * kStride;
  const int nu = L.nu, nq = L.nq, nv = L.nv;
  constexpr int TOTP = tpad(mp.total());
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) smem[(i % kStride) * TOTP + i / kStride] = A.blob[L.f_ltab + i];
  __syncthreads();
  const int lg = (LGC >= 0) ? (LGC & 3) : L.lg2G;
  const int g = threadIdx.x & ((1 << lg) - 1);
  const int wraw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> lg);
  const bool valid = wraw < A.E;
  const int w = valid ? wraw : A.E - 1;
  LaneCfg<LGC> C = make_cfg<NL, NC, LGC, SPEC>(L, smem, g);
  if (A.dr) {
    const float* r = A.dr + (size_t)A.dr_n * w;
    C.frs = r[0]; C.acs = r[1];
    if (A.dr_n >= 4) { C.dps = r[2]; C.ars = r[3]; }
  }
  const bool reset = A.reset_mask && A.reset_mask[w];
  const bool was_done = A.t_steps && A.t_done[w];  // AutoResetWrapper zeroes the counter of an env that finished last step
  const float* sq = (reset ? A.first_qpos : A.qpos) + (size_t)w * nq;
  const float* sv = (reset ? A.first_qvel : A.qvel) + (size_t)w * nv;
  const float* sw = reset ? A.first_warm : A.warm;
  if (sw) sw += (size_t)w * nv;
  Lane<NL, NC> s;
#pragma unroll
  for (int i = 0; i < 7; i++) s.qt[i] = sq[i];
#pragma unroll
  for (int i = 0; i < 6; i++) { s.v[i] = sv[i]; s.warm[i] = sw ? sw[i] : 0.f; s.a[i] = 0.f; }
#pragma unroll
  for (int p = 1; p <= NL; p++) {
    const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2), ga = LTI(mp.ijnt(p) + 3);
    s.qc[p - 1] = (gd >= 0) ? sq[gq] : 0.f;
    s.v[5 + p] = (gd >= 0) ? sv[gd] : 0.f;
    s.warm[5 + p] = (gd >= 0 && sw) ? sw[gd] : 0.f;
    s.ctrl[p - 1] = (ga >= 0 && A.ctrl) ? A.ctrl[(size_t)w * nu + ga] : 0.f;
    s.a[5 + p] = 0.f;
  }
  const int nfw = A.forward_only ? 1 : A.nsubsteps;
#pragma unroll 1
  for (int it = 0; it < nfw; it++) {
    float M[NTRI], fs[N], fc[N];
    forward<NL, NC, LGC, CB, SPEC>(s, C, M, fs, fc);
    if (!A.forward_only) euler<NL, NC, LGC, SPEC>(s, C, M, fs, fc);
  }
  bool fin = false;
  if (A.t_steps) {  // EpisodeWrapper + AutoResetWrapper semantics around obs / reward / done of the stepped state
    float r = quad_x_diag<NL, NC, LGC>(s, C, A.t_qd, A.t_xg, nq);
#pragma unroll
    for (int p = 1; p <= NL; p++) {
      const int ga = LTI(mp.ijnt(p) + 3);
      if (ga >= 0 && C.S.o(p)) r = fmaf(A.t_rd[ga] * s.ctrl[p - 1], s.ctrl[p - 1], r);
    }
    r = gall(r, lg);
    const int st = ((reset || was_done) ? 0 : A.t_steps[w]) + 1;
    const bool term = !(s.qt[2] >= A.t_zmin);  // a non-finite height terminates too
    const bool trunc = A.t_max_steps > 0 && st >= A.t_max_steps;
    fin = term || trunc;
    if (valid && g == 0) {
      A.t_reward[w] = -0.5f * r; A.t_done[w] = fin ? 1 : 0; A.t_steps[w] = st;
      if (A.t_trunc) A.t_trunc[w] = (trunc && !term) ? 1 : 0;
    }
    if (fin) {  // leave the launch already reset: where(done, first_state, state)
      const float* fq = A.first_qpos + (size_t)w * nq; const float* fv = A.first_qvel + (size_t)w * nv;
      const float* fw = A.first_warm ? A.first_warm + (size_t)w * nv : nullptr;
#pragma unroll
      for (int i = 0; i < 7; i++) s.qt[i] = fq[i];
#pragma unroll
      for (int i = 0; i < 6; i++) { s.v[i] = fv[i]; s.warm[i] = fw ? fw[i] : 0.f; }
#pragma unroll
      for (int p = 1; p <= NL; p++) {
        const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2);
        s.qc[p - 1] = (gd >= 0) ? fq[gq] : 0.f;
        s.v[5 + p] = (gd >= 0) ? fv[gd] : 0.f;
        s.warm[5 + p] = (gd >= 0 && fw) ? fw[gd] : 0.f;
      }
    }
    if (valid && A.t_obs) store_x<NL, NC, LGC>(s, C, A.t_obs + (size_t)w * (nq + nv), nq);
  }
  if (valid) {
    float* oq = A.qpos + (size_t)w * nq; float* ov = A.qvel + (size_t)w * nv;
    float* ow = A.warm ? A.warm + (size_t)w * nv : nullptr; float* oa = A.qacc ? A.qacc + (size_t)w * nv : nullptr;
    if (C.S.o(0)) {
#pragma unroll
      for (int i = 0; i < 7; i++) oq[i] = s.qt[i];
#pragma unroll
      for (int i = 0; i < 6; i++) {
        if (!A.forward_only) ov[i] = s.v[i];
        if (ow) ow[i] = s.warm[i];
        if (oa) oa[i] = s.a[i];
      }
      if (A.time) {
        const float t0 = reset ? 0.f : A.time[w];
        A.time[w] = fin ? 0.f : (A.forward_only ? t0 : t0 + L.timestep * (float)A.nsubsteps);
      }
    }
#pragma unroll
    for (int p = 1; p <= NL; p++) {
      const int gd = LTI(mp.ijnt(p) + 1), gq = LTI(mp.ijnt(p) + 2);
      if (gd >= 0 && C.S.o(p)) {
        oq[gq] = s.qc[p - 1];
        if (!A.forward_only) ov[gd] = s.v[5 + p];
        if (ow) ow[gd] = s.warm[5 + p];
        if (oa) oa[gd] = s.a[5 + p];
      }
    }
  }
}

// extra_floats: per CTA; extra_per_thread: per thread of the CTA
template <int NL, int NC, class Args, class K> int launch_limb(K kern, const Layout& L, const Args& a, int nworld, int extra_floats, int extra_per_thread, cudaStream_t st) {
  constexpr Map mp{NL, NC};
  const long threads = (long)nworld << L.lg2G;
  const int tpb = pick_tpb((threads + 31) / 32, NL);
  const size_t sm = sizeof(float) * ((size_t)tpad(mp.total()) * kStride + extra_floats + (size_t)extra_per_thread * tpb);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * ((size_t)tpad(mp.total()) * kStride + extra_floats + (size_t)extra_per_thread * kMaxTPB)));
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)((threads + tpb - 1) / tpb);
  static const int cl = [] { const char* e = getenv("ABR_LIMB_CLUSTER"); return e ? atoi(e) : 0; }();  // probe: co-scheduled CTA groups
  if (cl > 1 && cl <= 8) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((grid + cl - 1) / cl * cl)); cfg.blockDim = dim3((unsigned)tpb); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, L, a);
    return e != cudaSuccess ? (int)e : (int)cudaGetLastError();
  }
  kern<<<grid, tpb, sm, st>>>(L, a);
  return (int)cudaGetLastError();
}

#undef LTF
#undef LTI

}  // namespace limb

// LGC: -1 = general sharing pattern (suffix g), 2 = flat 4-lane pattern (suffix f2), 86 = biped pattern (suffix b).
// STAG: sg = general (SPEC -1), s0 / s4 / s8 / s12 = the fast variants (see limb::Spec).
#define ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, STAG) int launch_limb_rollout_##NL##_##NC##_##TAG##_##STAG(const Layout&, const RolloutArgs&, cudaStream_t);
#define ABR_DECLARE_LIMB_ENV(NL, NC, TAG, STAG) int launch_limb_env_##NL##_##NC##_##TAG##_##STAG(const Layout&, const EnvArgs&, cudaStream_t);
#define ABR_DEFINE_LIMB_ROLLOUT(NL, NC, LGC, CB, TAG, SPEC, STAG)                                                         \
  int launch_limb_rollout_##NL##_##NC##_##TAG##_##STAG(const Layout& L, const RolloutArgs& a, cudaStream_t st) {           \
    return limb::launch_limb<NL, NC>(limb::k_limb_rollout<NL, NC, LGC, CB, SPEC>, L, a, a.nworld,                          \
                                     3 * L.nx + L.nu, 3 * (13 + 2 * NL) + 3 * NL, st);                                        \
  }
#define ABR_DEFINE_LIMB_ENV(NL, NC, LGC, CB, TAG, SPEC, STAG)                                                             \
  int launch_limb_env_##NL##_##NC##_##TAG##_##STAG(const Layout& L, const EnvArgs& a, cudaStream_t st) {                   \
    return limb::launch_limb<NL, NC>(limb::k_limb_env<NL, NC, LGC, CB, SPEC>, L, a, a.E, 0, 0, st);                           \
  }
#define ABR_ALIAS_LIMB_ROLLOUT(NL, NC, TAG, STAG)                                                                           \
  int launch_limb_rollout_##NL##_##NC##_##TAG##_##STAG(const Layout& L, const RolloutArgs& a, cudaStream_t st) {           \
    return launch_limb_rollout_##NL##_##NC##_##TAG##_sg(L, a, st);                                                        \
  }
#define ABR_ALIAS_LIMB_ENV(NL, NC, TAG, STAG)                                                                               \
  int launch_limb_env_##NL##_##NC##_##TAG##_##STAG(const Layout& L, const EnvArgs& a, cudaStream_t st) {                   \
    return launch_limb_env_##NL##_##NC##_##TAG##_sg(L, a, st);                                                            \
  }
#define ABR_DECLARE_LIMB_FAMILY_FAST(NL, NC, TAG)                                                                         \
  ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, sg) ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s0) ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s4) \
  ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s8) ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s12)                                    \
  ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s1) ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s5)                                     \
  ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s9) ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, s13)                                    \
  ABR_DECLARE_LIMB_ENV(NL, NC, TAG, sg) ABR_DECLARE_LIMB_ENV(NL, NC, TAG, s0) ABR_DECLARE_LIMB_ENV(NL, NC, TAG, s1)
#define ABR_DECLARE_LIMB_FAMILY_GENERAL(NL, NC, TAG) ABR_DECLARE_LIMB_ROLLOUT(NL, NC, TAG, sg) ABR_DECLARE_LIMB_ENV(NL, NC, TAG, sg)
// families (compile-time sharing patterns): f2 = flat 4 lanes (quadruped; with 6-joint chains: a humanoid without waist joints), f3 = flat 8
// lanes (hexapod), l2 = flat 2 lanes (legs only: exoskeleton / biped without arms), b = biped with a torso chain shared by the arm paths;
// g = any pattern, read from the lane table (3x slower: the merges and their level tests are then run-time work at every position)
ABR_DECLARE_LIMB_FAMILY_FAST(3, 1, f2)
ABR_DECLARE_LIMB_FAMILY_FAST(3, 1, f3)
ABR_DECLARE_LIMB_FAMILY_GENERAL(3, 1, g)
ABR_DECLARE_LIMB_FAMILY_GENERAL(6, 4, g)
ABR_DECLARE_LIMB_FAMILY_FAST(6, 4, b)
ABR_DECLARE_LIMB_FAMILY_FAST(6, 4, l2)
ABR_DECLARE_LIMB_FAMILY_FAST(6, 4, f2)

}  // namespace abr
#endif
