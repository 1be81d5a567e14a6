// abr_oracle.cc — CPU ORACLE. TEST INFRASTRUCTURE ONLY: nothing in the product path
// (ambersim_b200/, the C-ABI library) may include, link or call this file. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY UNPINNED: the arithmetic of the hot path lives in third-party `mujoco-mjx` / `mujoco`
// (reference pyproject.toml:24-25, `mujoco>=3.0.0`, `mujoco-mjx>=3.0.0`, un-vendored, un-pinned;
// call sites ambersim/trajopt/shooting.py:34,36,41 and ambersim/rl/base.py:52,83,85,93).
// Neither package nor JAX can be installed in this image and the reference holds no golden
// vectors for this path, so this file restates the PUBLISHED MJX 3.0.1-3.1.x
// algorithm (dense path, pyramidal cones, static constraint sizes) as summarised in SURVEY.md
// Appendix A, and is validated by physics self-checks (tests/test_oracle_physics.py) instead of
// reference outputs. tools/dump_mjx_golden.py pins it off-box when MJX is available.
//
// Structure follows mjx/_src: forward.py (forward/step/euler/rungekutta4), smooth.py (kinematics,
// com_pos, crb, factor_m, com_vel, rne, transmission), passive.py, collision_primitive.py,
// constraint.py (make_constraint, _kbi), solver.py (solve, _linesearch), support.py (jac, mul_m).
// Templated on the scalar: double (ground truth), float (the reference's own precision), and an
// op-counting scalar that yields the algorithmic FLOP figure.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <memory>
#include <vector>

#include "../include/abr.h"

namespace {

constexpr double kMinVal = 1e-15;  // mjMINVAL
constexpr double kMinImp = 1e-4;   // mjMINIMP
constexpr double kMaxImp = 0.9999; // mjMAXIMP

// ----------------------------------------------------------------------------- op counter
struct Cnt {
  double v;
  static thread_local long long n;
  Cnt() : v(0) {}
  Cnt(double x) : v(x) {}
  explicit operator double() const { return v; }
};
thread_local long long Cnt::n = 0;
inline Cnt operator+(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { Cnt::n++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt& operator+=(Cnt& a, Cnt b) { Cnt::n++; a.v += b.v; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { Cnt::n++; a.v -= b.v; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { Cnt::n++; a.v *= b.v; return a; }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
inline bool operator==(Cnt a, Cnt b) { return a.v == b.v; }

inline double Sqrt(double x) { return std::sqrt(x); }
inline float Sqrt(float x) { return std::sqrt(x); }
inline Cnt Sqrt(Cnt x) { Cnt::n++; return Cnt(std::sqrt(x.v)); }
inline double Sin(double x) { return std::sin(x); }
inline float Sin(float x) { return std::sin(x); }
inline Cnt Sin(Cnt x) { Cnt::n++; return Cnt(std::sin(x.v)); }
inline double Cos(double x) { return std::cos(x); }
inline float Cos(float x) { return std::cos(x); }
inline Cnt Cos(Cnt x) { Cnt::n++; return Cnt(std::cos(x.v)); }
inline double Pow(double x, double y) { return std::pow(x, y); }
inline float Pow(float x, float y) { return std::pow(x, y); }
inline Cnt Pow(Cnt x, Cnt y) { Cnt::n++; return Cnt(std::pow(x.v, y.v)); }
inline double Abs(double x) { return std::fabs(x); }
inline float Abs(float x) { return std::fabs(x); }
inline Cnt Abs(Cnt x) { return Cnt(std::fabs(x.v)); }
template <class T> inline T Max(T a, T b) { return a > b ? a : b; }
template <class T> inline T Min(T a, T b) { return a < b ? a : b; }
template <class T> inline T Clip(T x, T lo, T hi) { return Min(Max(x, lo), hi); }
inline double ToD(double x) { return x; }
inline double ToD(float x) { return x; }
inline double ToD(Cnt x) { return x.v; }

// ----------------------------------------------------------------------------- small math
// mjx/_src/math.py
template <class T> inline void cross(const T* a, const T* b, T* r) {
  T r0 = a[1] * b[2] - a[2] * b[1];
  T r1 = a[2] * b[0] - a[0] * b[2];
  T r2 = a[0] * b[1] - a[1] * b[0];
  r[0] = r0; r[1] = r1; r[2] = r2;
}
template <class T> inline T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> inline void quat_mul(const T* u, const T* v, T* r) {
  T r0 = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  T r1 = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  T r2 = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  T r3 = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3;
}
// math.rotate: r = 2 (u.v) u + (s^2 - u.u) v + 2 s (u x v)
template <class T> inline void rotate(const T* vec, const T* quat, T* r) {
  const T s = quat[0];
  const T* u = quat + 1;
  T uv = dot3(u, vec), uu = dot3(u, u);
  T c[3];
  cross(u, vec, c);
  T out[3];
  for (int i = 0; i < 3; i++) out[i] = T(2) * (uv * u[i]) + (s * s - uu) * vec[i] + T(2) * s * c[i];
  r[0] = out[0]; r[1] = out[1]; r[2] = out[2];
}
template <class T> inline void quat_to_mat(const T* q, T* m) {
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  T q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = T(2) * (q12 - q03);     m[2] = T(2) * (q13 + q02);
  m[3] = T(2) * (q12 + q03);     m[4] = q00 - q11 + q22 - q33; m[5] = T(2) * (q23 - q01);
  m[6] = T(2) * (q13 - q02);     m[7] = T(2) * (q23 + q01);     m[8] = q00 - q11 - q22 + q33;
}
template <class T> inline void axis_angle_to_quat(const T* axis, T angle, T* q) {
  T s = Sin(angle * T(0.5)), c = Cos(angle * T(0.5));
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
// math.normalize_with_norm: x / (n + 1e-6 * (n == 0))
template <class T> inline T normalize(T* x, int n) {
  T s = T(0);
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  T nrm = Sqrt(s);
  T den = (nrm == T(0)) ? T(1e-6) : nrm;
  for (int i = 0; i < n; i++) x[i] = x[i] / den;
  return nrm;
}
// math.inert_mul: cinert = (Ixx,Iyy,Izz,Ixy,Ixz,Iyz, m*off[3], m)
template <class T> inline void inert_mul(const T* I, const T* v, T* r) {
  T ang[3], vel[3], c1[3], c2[3];
  ang[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2];
  ang[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2];
  ang[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2];
  cross(I + 6, v + 3, c1);
  cross(I + 6, v, c2);
  for (int i = 0; i < 3; i++) { ang[i] = ang[i] + c1[i]; vel[i] = I[9] * v[3 + i] - c2[i]; }
  for (int i = 0; i < 3; i++) { r[i] = ang[i]; r[3 + i] = vel[i]; }
}
// math.motion_cross(u, v)
template <class T> inline void motion_cross(const T* u, const T* v, T* r) {
  T a[3], b[3], c[3];
  cross(u, v, a);
  cross(u + 3, v, b);
  cross(u, v + 3, c);
  for (int i = 0; i < 3; i++) { r[i] = a[i]; r[3 + i] = b[i] + c[i]; }
}
// math.motion_cross_force(v, f)
template <class T> inline void motion_cross_force(const T* v, const T* f, T* r) {
  T a[3], b[3], c[3];
  cross(v, f, a);
  cross(v + 3, f + 3, b);
  cross(v, f + 3, c);
  for (int i = 0; i < 3; i++) { r[i] = a[i] + b[i]; r[3 + i] = c[i]; }
}
// math.make_frame
template <class T> inline void make_frame(const T* a_in, T* frame) {
  T a[3] = {a_in[0], a_in[1], a_in[2]};
  normalize(a, 3);
  T b[3] = {T(0), T(0), T(0)};
  if (T(-0.5) < a[1] && a[1] < T(0.5)) b[1] = T(1); else b[2] = T(1);
  T ab = dot3(a, b);
  for (int i = 0; i < 3; i++) b[i] = b[i] - a[i] * ab;
  normalize(b, 3);
  T c[3];
  cross(a, b, c);
  for (int i = 0; i < 3; i++) { frame[i] = a[i]; frame[3 + i] = b[i]; frame[6 + i] = c[i]; }
}

// dense Cholesky (lower), jax.scipy.linalg.cho_factor / cho_solve
template <class T> bool cholesky(const T* A, T* L, int n) {
  bool ok = true;
  for (int i = 0; i < n * n; i++) L[i] = T(0);
  for (int j = 0; j < n; j++) {
    T s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (!(s > T(0))) ok = false;
    T d = Sqrt(s);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      T t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / d;
    }
  }
  return ok;
}
template <class T> void cho_solve(const T* L, const T* b, T* x, int n) {
  std::vector<T> y(n);
  for (int i = 0; i < n; i++) {
    T s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    T s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}


// ----------------------------------------------------------------------------- convex collision helpers
// mjx/_src/math.py + collision_convex.py [MEMORY, MJX 3.1.x; parity unpinned]: the geometric primitives of sphere - convex,
// capsule - convex and convex - convex collision.
template <class T> inline void closest_segment_point(const T* a, const T* b, const T* pt, T* out) {
  T ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  T t = Clip(dot3(pa, ab) / (dot3(ab, ab) + T(1e-6)), T(0), T(1));
  for (int i = 0; i < 3; i++) out[i] = a[i] + t * ab[i];
}
// math.closest_segment_to_segment_points
template <class T> inline void closest_segment_to_segment_points(const T* a0, const T* a1, const T* b0, const T* b1, T* best_a, T* best_b) {
  T da[3], db[3];
  for (int i = 0; i < 3; i++) { da[i] = a1[i] - a0[i]; db[i] = b1[i] - b0[i]; }
  const T la = normalize(da, 3), lb = normalize(db, 3);
  const T ha = la * T(0.5), hb = lb * T(0.5);
  T am[3], bm[3], tr[3];
  for (int i = 0; i < 3; i++) { am[i] = a0[i] + da[i] * ha; bm[i] = b0[i] + db[i] * hb; tr[i] = am[i] - bm[i]; }
  const T dd = dot3(da, db), dat = dot3(da, tr), dbt = dot3(db, tr);
  const T denom = T(1) - dd * dd;
  const T ota = (-dat + dd * dbt) / (denom + T(1e-6));
  const T otb = dbt + ota * dd;
  const T ta = Clip(ota, -ha, ha), tb = Clip(otb, -hb, hb);
  T pa[3], pb[3];
  for (int i = 0; i < 3; i++) { pa[i] = am[i] + da[i] * ta; pb[i] = bm[i] + db[i] * tb; }
  T na[3], nb[3];
  closest_segment_point(a0, a1, pb, na);
  closest_segment_point(b0, b1, pa, nb);
  T e1[3] = {na[0] - pb[0], na[1] - pb[1], na[2] - pb[2]}, e2[3] = {pa[0] - nb[0], pa[1] - nb[1], pa[2] - nb[2]};
  const bool first = dot3(e1, e1) < dot3(e2, e2);
  for (int i = 0; i < 3; i++) { best_a[i] = first ? na[i] : pa[i]; best_b[i] = first ? pb[i] : nb[i]; }
}
template <class T> inline void project_pt_onto_plane(const T* pt, const T* plane_pt, const T* n, T* out) {
  T d[3] = {pt[0] - plane_pt[0], pt[1] - plane_pt[1], pt[2] - plane_pt[2]};
  const T dist = dot3(d, n);
  for (int i = 0; i < 3; i++) out[i] = pt[i] - dist * n[i];
}
// collision_convex._closest_segment_point_plane: where the segment (a, b) meets the plane, clipped to the segment
template <class T> inline void closest_segment_point_plane(const T* a, const T* b, const T* p0, const T* n, T* out) {
  T ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  const T d = dot3(p0, n), denom = dot3(n, ab);
  T t = (d - dot3(n, a)) / (denom + ((denom == T(0)) ? T(1e-6) : T(0)));
  t = Clip(t, T(0), T(1));
  for (int i = 0; i < 3; i++) out[i] = a[i] + t * ab[i];
}
// collision_convex._clip_edge_to_planes: the segment (p0, p1) cut by np side planes (point, outward normal); `ok` is false when
// the segment lies entirely in front of one plane (then the original end points are returned) or the cut end points crossed
template <class T> inline void clip_edge_to_planes(const T* p0, const T* p1, const T* plane_pts, const T* plane_normals, int np, T* out0, T* out1, bool* ok) {
  T e01[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e10[3] = {-e01[0], -e01[1], -e01[2]};
  bool any_both = false;
  T best0 = T(-1e30), best1 = T(-1e30);
  T n0[3] = {p0[0], p0[1], p0[2]}, n1[3] = {p1[0], p1[1], p1[2]};
  for (int k = 0; k < np; k++) {
    const T* pp = plane_pts + 3 * k; const T* pn = plane_normals + 3 * k;
    T d0[3] = {p0[0] - pp[0], p0[1] - pp[1], p0[2] - pp[2]}, d1[3] = {p1[0] - pp[0], p1[1] - pp[1], p1[2] - pp[2]};
    const bool f0 = dot3(d0, pn) > T(1e-6), f1 = dot3(d1, pn) > T(1e-6);
    any_both = any_both || (f0 && f1);
    T cp[3];
    closest_segment_point_plane(p0, p1, pp, pn, cp);
    // the candidate for an end point is the cut point where that end is in front of the plane, the end point itself otherwise;
    // the one furthest along the edge wins (first maximum)
    T c0[3], c1[3];
    for (int i = 0; i < 3; i++) { c0[i] = f0 ? cp[i] : p0[i]; c1[i] = f1 ? cp[i] : p1[i]; }
    T r0[3] = {c0[0] - p0[0], c0[1] - p0[1], c0[2] - p0[2]}, r1[3] = {c1[0] - p1[0], c1[1] - p1[1], c1[2] - p1[2]};
    const T s0 = dot3(r0, e01), s1 = dot3(r1, e10);
    if (s0 > best0) { best0 = s0; for (int i = 0; i < 3; i++) n0[i] = c0[i]; }
    if (s1 > best1) { best1 = s1; for (int i = 0; i < 3; i++) n1[i] = c1[i]; }
  }
  bool mask = !any_both;
  for (int i = 0; i < 3; i++) { out0[i] = mask ? n0[i] : p0[i]; out1[i] = mask ? n1[i] : p1[i]; }
  T dn[3] = {out0[0] - out1[0], out0[1] - out1[1], out0[2] - out1[2]};
  if (dot3(e10, dn) < T(0)) mask = false;  // (p0 - p1).(new0 - new1) < 0: the cut points crossed
  *ok = mask;
}
// collision_convex._manifold_points: up to four well-spread points of a masked planar point set (first maximum wins, as jnp.argmax)
template <class T> inline void manifold_points(const T* pts, const bool* mask, int n, const T* normal, int idx[4]) {
  auto dm = [&](int k) { return mask[k] ? T(0) : T(-1e6); };
  auto argmax = [&](auto score) { int best = 0; T bv = score(0); for (int k = 1; k < n; k++) { T v = score(k); if (v > bv) { bv = v; best = k; } } return best; };
  const int ia = argmax([&](int k) { return dm(k); });
  const T* a = pts + 3 * ia;
  const int ib = argmax([&](int k) { T d[3] = {a[0] - pts[3 * k], a[1] - pts[3 * k + 1], a[2] - pts[3 * k + 2]}; return dot3(d, d) + dm(k); });
  const T* b = pts + 3 * ib;
  T amb[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]}, ab[3];
  cross(normal, amb, ab);
  const int ic = argmax([&](int k) { T d[3] = {a[0] - pts[3 * k], a[1] - pts[3 * k + 1], a[2] - pts[3 * k + 2]}; return Abs(dot3(d, ab)) + dm(k); });
  const T* c = pts + 3 * ic;
  T amc[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]}, bmc[3] = {b[0] - c[0], b[1] - c[1], b[2] - c[2]}, ac[3], bc[3];
  cross(normal, amc, ac);
  cross(normal, bmc, bc);
  int id = 0;
  T bv = T(-1e30);
  for (int h = 0; h < 2; h++)
    for (int k = 0; k < n; k++) {
      const T* o = h == 0 ? b : a;
      const T* ax = h == 0 ? bc : ac;
      T d[3] = {o[0] - pts[3 * k], o[1] - pts[3 * k + 1], o[2] - pts[3 * k + 2]};
      const T v = Abs(dot3(d, ax)) + dm(k);
      if (v > bv) { bv = v; id = k; }
    }
  idx[0] = ia; idx[1] = ib; idx[2] = ic; idx[3] = id;
}

// ----------------------------------------------------------------------------- the simulator
struct Sizes { int ncon, ne, nl, nefc; };

inline int pair_ncon(int kind) {
  if (kind == ABR_PAIR_PLANE_CONVEX || kind == ABR_PAIR_CONVEX_CONVEX) return 4;
  return (kind == ABR_PAIR_PLANE_CAPSULE || kind == ABR_PAIR_CAPSULE_CONVEX) ? 2 : 1;
}

Sizes compute_sizes(const AbrModelHost* m) {
  Sizes s{0, 0, 0, 0};
  int dis = m->opt.disableflags;
  if (dis & ABR_DSBL_CONSTRAINT) return s;
  if (!(dis & ABR_DSBL_EQUALITY))
    for (int i = 0; i < m->neq; i++) if (m->eq_active[i]) s.ne++;
  if (!(dis & ABR_DSBL_LIMIT))
    for (int j = 0; j < m->njnt; j++)
      if (m->jnt_limited[j] && (m->jnt_type[j] == ABR_JNT_HINGE || m->jnt_type[j] == ABR_JNT_SLIDE)) s.nl++;
  int ncrow = 0;
  if (!(dis & ABR_DSBL_CONTACT))
    for (int p = 0; p < m->npair; p++) {
      s.ncon += pair_ncon(m->pair_kind[p]);
      ncrow += pair_ncon(m->pair_kind[p]) * (m->pair_condim[p] == 1 ? 1 : 4);
    }
  s.nefc = s.ne + s.nl + ncrow;
  return s;
}

template <class T> struct Sim {
  const AbrModelHost* m;
  Sizes sz;
  bool literal_post = false;  // also run MJX's unused post-iteration gradient/Hessian (FLOP count)
  int nq, nv, nu, nbody, njnt, ngeom;
  // persistent state
  std::vector<T> qpos, qvel, ctrl, qacc_warmstart;
  T time = T(0);
  // derived
  std::vector<T> xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, geom_xpos, geom_xmat;
  std::vector<T> subtree_com, cinert, cdof, crb, qM, qL;
  std::vector<T> con_dist, con_pos, con_frame; std::vector<int> con_pair;
  std::vector<T> efc_J, efc_D, efc_aref, efc_pos;
  std::vector<T> act_length, act_velocity, act_force, qfrc_actuator;
  std::vector<T> cvel, cdof_dot, qfrc_passive, qfrc_bias, qfrc_smooth, qacc_smooth, qacc;
  std::vector<T> qfrc_constraint, efc_force;
  int solver_niter = 0;
  // model promoted to T
  std::vector<T> mf;  // unused placeholder

  explicit Sim(const AbrModelHost* model) : m(model) {
    sz = compute_sizes(m);
    nq = m->nq; nv = m->nv; nu = m->nu; nbody = m->nbody; njnt = m->njnt; ngeom = m->ngeom;
    qpos.assign(nq, T(0)); qvel.assign(nv, T(0)); ctrl.assign(nu, T(0)); qacc_warmstart.assign(nv, T(0));
    for (int i = 0; i < nq; i++) qpos[i] = T(m->qpos0[i]);  // make_data: qpos = qpos0
    xpos.assign(nbody * 3, T(0)); xquat.assign(nbody * 4, T(0)); xmat.assign(nbody * 9, T(0));
    xipos.assign(nbody * 3, T(0)); ximat.assign(nbody * 9, T(0));
    xanchor.assign(njnt * 3, T(0)); xaxis.assign(njnt * 3, T(0));
    geom_xpos.assign(ngeom * 3, T(0)); geom_xmat.assign(ngeom * 9, T(0));
    subtree_com.assign(nbody * 3, T(0)); cinert.assign(nbody * 10, T(0)); crb.assign(nbody * 10, T(0));
    cdof.assign(nv * 6, T(0)); qM.assign(nv * nv, T(0)); qL.assign(nv * nv, T(0));
    con_dist.assign(sz.ncon, T(0)); con_pos.assign(sz.ncon * 3, T(0)); con_frame.assign(sz.ncon * 9, T(0));
    con_pair.assign(sz.ncon, 0);
    efc_J.assign(sz.nefc * nv, T(0)); efc_D.assign(sz.nefc, T(0)); efc_aref.assign(sz.nefc, T(0));
    efc_pos.assign(sz.nefc, T(0));
    act_length.assign(nu, T(0)); act_velocity.assign(nu, T(0)); act_force.assign(nu, T(0));
    qfrc_actuator.assign(nv, T(0));
    cvel.assign(nbody * 6, T(0)); cdof_dot.assign(nv * 6, T(0));
    qfrc_passive.assign(nv, T(0)); qfrc_bias.assign(nv, T(0)); qfrc_smooth.assign(nv, T(0));
    qacc_smooth.assign(nv, T(0)); qacc.assign(nv, T(0)); qfrc_constraint.assign(nv, T(0));
    efc_force.assign(sz.nefc, T(0));
  }

  T F(const float* a, int i) const { return T(double(a[i])); }

  // ---- smooth.kinematics
  void kinematics() {
    for (int b = 0; b < nbody; b++) {
      T pos[3], quat[4];
      if (b == 0) {
        pos[0] = pos[1] = pos[2] = T(0);
        quat[0] = T(1); quat[1] = quat[2] = quat[3] = T(0);
      } else {
        int p = m->body_parentid[b];
        T bp[3] = {F(m->body_pos, 3 * b), F(m->body_pos, 3 * b + 1), F(m->body_pos, 3 * b + 2)};
        T bq[4] = {F(m->body_quat, 4 * b), F(m->body_quat, 4 * b + 1), F(m->body_quat, 4 * b + 2), F(m->body_quat, 4 * b + 3)};
        T r[3];
        rotate(bp, &xquat[4 * p], r);
        for (int i = 0; i < 3; i++) pos[i] = xpos[3 * p + i] + r[i];
        quat_mul(&xquat[4 * p], bq, quat);
      }
      for (int k = 0; k < m->body_jntnum[b]; k++) {
        int j = m->body_jntadr[b] + k;
        int a = m->jnt_qposadr[j];
        T jp[3] = {F(m->jnt_pos, 3 * j), F(m->jnt_pos, 3 * j + 1), F(m->jnt_pos, 3 * j + 2)};
        T ja[3] = {F(m->jnt_axis, 3 * j), F(m->jnt_axis, 3 * j + 1), F(m->jnt_axis, 3 * j + 2)};
        T anchor[3], axis[3];
        int type = m->jnt_type[j];
        if (type == ABR_JNT_FREE) {
          for (int i = 0; i < 3; i++) anchor[i] = qpos[a + i];
          axis[0] = T(0); axis[1] = T(0); axis[2] = T(1);
          for (int i = 0; i < 3; i++) pos[i] = qpos[a + i];
          for (int i = 0; i < 4; i++) quat[i] = qpos[a + 3 + i];
          normalize(quat, 4);
          for (int i = 0; i < 4; i++) qpos[a + 3 + i] = quat[i];  // normalised quat written back
        } else {
          T r[3];
          rotate(jp, quat, r);
          for (int i = 0; i < 3; i++) anchor[i] = r[i] + pos[i];
          rotate(ja, quat, axis);
          if (type == ABR_JNT_HINGE) {
            T ql[4], qn[4];
            axis_angle_to_quat(ja, qpos[a] - F(m->qpos0, a), ql);
            quat_mul(quat, ql, qn);
            for (int i = 0; i < 4; i++) quat[i] = qn[i];
            rotate(jp, quat, r);
            for (int i = 0; i < 3; i++) pos[i] = anchor[i] - r[i];
          } else if (type == ABR_JNT_SLIDE) {
            T dq = qpos[a] - F(m->qpos0, a);
            for (int i = 0; i < 3; i++) pos[i] = pos[i] + axis[i] * dq;
          }
        }
        for (int i = 0; i < 3; i++) { xanchor[3 * j + i] = anchor[i]; xaxis[3 * j + i] = axis[i]; }
      }
      for (int i = 0; i < 3; i++) xpos[3 * b + i] = pos[i];
      for (int i = 0; i < 4; i++) xquat[4 * b + i] = quat[i];
      quat_to_mat(quat, &xmat[9 * b]);
      T ip[3] = {F(m->body_ipos, 3 * b), F(m->body_ipos, 3 * b + 1), F(m->body_ipos, 3 * b + 2)};
      T iq[4] = {F(m->body_iquat, 4 * b), F(m->body_iquat, 4 * b + 1), F(m->body_iquat, 4 * b + 2), F(m->body_iquat, 4 * b + 3)};
      T r[3], q2[4];
      rotate(ip, quat, r);
      for (int i = 0; i < 3; i++) xipos[3 * b + i] = pos[i] + r[i];
      quat_mul(quat, iq, q2);
      quat_to_mat(q2, &ximat[9 * b]);
    }
    for (int g = 0; g < ngeom; g++) {
      int b = m->geom_bodyid[g];
      T gp[3] = {F(m->geom_pos, 3 * g), F(m->geom_pos, 3 * g + 1), F(m->geom_pos, 3 * g + 2)};
      T gq[4] = {F(m->geom_quat, 4 * g), F(m->geom_quat, 4 * g + 1), F(m->geom_quat, 4 * g + 2), F(m->geom_quat, 4 * g + 3)};
      T r[3], q2[4];
      rotate(gp, &xquat[4 * b], r);
      for (int i = 0; i < 3; i++) geom_xpos[3 * g + i] = xpos[3 * b + i] + r[i];
      quat_mul(&xquat[4 * b], gq, q2);
      quat_to_mat(q2, &geom_xmat[9 * g]);
    }
  }

  // ---- smooth.com_pos
  void com_pos() {
    std::vector<T> pos(nbody * 3), mass(nbody);
    for (int b = 0; b < nbody; b++) {
      mass[b] = F(m->body_mass, b);
      for (int i = 0; i < 3; i++) pos[3 * b + i] = xipos[3 * b + i] * mass[b];
    }
    for (int b = nbody - 1; b > 0; b--) {
      int p = m->body_parentid[b];
      for (int i = 0; i < 3; i++) pos[3 * p + i] += pos[3 * b + i];
      mass[p] += mass[b];
    }
    for (int b = 0; b < nbody; b++)
      for (int i = 0; i < 3; i++)
        subtree_com[3 * b + i] = (mass[b] < T(kMinVal)) ? xipos[3 * b + i] : pos[3 * b + i] / mass[b];
    for (int b = 0; b < nbody; b++) {
      const T* rc = &subtree_com[3 * m->body_rootid[b]];
      T off[3];
      for (int i = 0; i < 3; i++) off[i] = xipos[3 * b + i] - rc[i];
      T ms = F(m->body_mass, b);
      const T* R = &ximat[9 * b];
      T in[3] = {F(m->body_inertia, 3 * b), F(m->body_inertia, 3 * b + 1), F(m->body_inertia, 3 * b + 2)};
      // (ximat * inert) @ ximat.T + h h^T mass, h = cross(off, -I)  =>  mass (|off|^2 I - off off^T)
      T I[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          T s = T(0);
          for (int k = 0; k < 3; k++) s += R[3 * r + k] * in[k] * R[3 * c + k];
          I[3 * r + c] = s;
        }
      T oo = dot3(off, off);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) I[3 * r + c] += ms * ((r == c ? oo : T(0)) - off[r] * off[c]);
      T* ci = &cinert[10 * b];
      ci[0] = I[0]; ci[1] = I[4]; ci[2] = I[8]; ci[3] = I[1]; ci[4] = I[2]; ci[5] = I[5];
      for (int i = 0; i < 3; i++) ci[6 + i] = off[i] * ms;
      ci[9] = ms;
    }
    for (int j = 0; j < njnt; j++) {
      int b = m->jnt_bodyid[j];
      int d = m->jnt_dofadr[j];
      const T* rc = &subtree_com[3 * m->body_rootid[b]];
      T off[3];
      for (int i = 0; i < 3; i++) off[i] = rc[i] - xanchor[3 * j + i];
      int type = m->jnt_type[j];
      if (type == ABR_JNT_FREE) {
        for (int k = 0; k < 3; k++)
          for (int i = 0; i < 6; i++) cdof[6 * (d + k) + i] = (i == 3 + k) ? T(1) : T(0);
        for (int k = 0; k < 3; k++) {
          T a[3] = {xmat[9 * b + k], xmat[9 * b + 3 + k], xmat[9 * b + 6 + k]};  // column k of xmat
          T c[3];
          cross(a, off, c);
          for (int i = 0; i < 3; i++) { cdof[6 * (d + 3 + k) + i] = a[i]; cdof[6 * (d + 3 + k) + 3 + i] = c[i]; }
        }
      } else if (type == ABR_JNT_HINGE) {
        T c[3];
        cross(&xaxis[3 * j], off, c);
        for (int i = 0; i < 3; i++) { cdof[6 * d + i] = xaxis[3 * j + i]; cdof[6 * d + 3 + i] = c[i]; }
      } else if (type == ABR_JNT_SLIDE) {
        for (int i = 0; i < 3; i++) { cdof[6 * d + i] = T(0); cdof[6 * d + 3 + i] = xaxis[3 * j + i]; }
      }
    }
  }

  // ---- smooth.crb + support.make_m (dense) + smooth.factor_m (dense Cholesky)
  void crb_and_factor() {
    for (int i = 0; i < nbody * 10; i++) crb[i] = cinert[i];
    for (int b = nbody - 1; b > 0; b--) {
      int p = m->body_parentid[b];
      for (int i = 0; i < 10; i++) crb[10 * p + i] += crb[10 * b + i];
    }
    for (int i = 0; i < 10; i++) crb[i] = T(0);
    for (int i = 0; i < nv * nv; i++) qM[i] = T(0);
    for (int i = 0; i < nv; i++) {
      T buf[6];
      inert_mul(&crb[10 * m->dof_bodyid[i]], &cdof[6 * i], buf);
      for (int j = i; j >= 0; j = m->dof_parentid[j]) {
        T s = T(0);
        for (int k = 0; k < 6; k++) s += cdof[6 * j + k] * buf[k];
        qM[i * nv + j] = s;
        qM[j * nv + i] = s;
      }
      qM[i * nv + i] += F(m->dof_armature, i);
    }
    cholesky(qM.data(), qL.data(), nv);
  }

  void mul_m(const T* v, T* r) const {
    for (int i = 0; i < nv; i++) {
      T s = T(0);
      for (int j = 0; j < nv; j++) s += qM[i * nv + j] * v[j];
      r[i] = s;
    }
  }

  // ---- collision_driver.collision (static pairs) + collision_primitive.py
  void plane_sphere_core(const T* n, const T* ppos, const T* spos, T radius, T* dist, T* pos) {
    T d[3];
    for (int i = 0; i < 3; i++) d[i] = spos[i] - ppos[i];
    *dist = dot3(d, n) - radius;
    for (int i = 0; i < 3; i++) pos[i] = spos[i] - n[i] * (radius + T(0.5) * (*dist));
  }
  void collision() {
    int c = 0;
    if (sz.ncon == 0) return;
    for (int p = 0; p < m->npair; p++) {
      int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
      int kind = m->pair_kind[p];
      const T* p1 = &geom_xpos[3 * g1];
      const T* p2 = &geom_xpos[3 * g2];
      const T* m1 = &geom_xmat[9 * g1];
      const T* m2 = &geom_xmat[9 * g2];
      if (kind == ABR_PAIR_PLANE_SPHERE) {
        T n[3] = {m1[2], m1[5], m1[8]};
        plane_sphere_core(n, p1, p2, F(m->geom_size, 3 * g2), &con_dist[c], &con_pos[3 * c]);
        make_frame(n, &con_frame[9 * c]);
        con_pair[c++] = p;
      } else if (kind == ABR_PAIR_PLANE_CAPSULE) {
        T n[3] = {m1[2], m1[5], m1[8]};
        T axis[3] = {m2[2], m2[5], m2[8]};
        T na = dot3(n, axis);
        T b[3];
        for (int i = 0; i < 3; i++) b[i] = axis[i] - n[i] * na;
        T bn = normalize(b, 3);
        if (bn < T(0.5)) {
          b[0] = b[1] = b[2] = T(0);
          if (T(-0.5) < n[1] && n[1] < T(0.5)) b[1] = T(1); else b[2] = T(1);
        }
        T cr[3];
        cross(n, b, cr);
        T half = F(m->geom_size, 3 * g2 + 1), rad = F(m->geom_size, 3 * g2);
        for (int s = 0; s < 2; s++) {
          T sp[3];
          for (int i = 0; i < 3; i++) sp[i] = p2[i] + (s == 0 ? axis[i] * half : -(axis[i] * half));
          plane_sphere_core(n, p1, sp, rad, &con_dist[c], &con_pos[3 * c]);
          for (int i = 0; i < 3; i++) { con_frame[9 * c + i] = n[i]; con_frame[9 * c + 3 + i] = b[i]; con_frame[9 * c + 6 + i] = cr[i]; }
          con_pair[c++] = p;
        }
      } else if (kind == ABR_PAIR_PLANE_CONVEX) {
        // mjx collision_convex.plane_convex [MEMORY, MJX 3.1.x]: everything in the convex geom's frame. support = penetration
        // depth of each vertex; the manifold is picked among the vertices within 1 mm of the deepest one by _manifold_points
        // (a = first masked vertex, b = furthest from a, c = furthest from the line ab, d = furthest from the edges ac / bc;
        // jnp.argmax = FIRST maximum); a vertex picked twice gives an inactive contact (dist = 1).
        const int va = m->geom_vertadr[g2], nvt = m->geom_vertnum[g2];
        T n[3] = {m1[2], m1[5], m1[8]};
        T dp[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
        T pl[3], nl[3];  // plane point and normal in the convex frame: mat' (.)
        for (int i = 0; i < 3; i++) { pl[i] = m2[i] * dp[0] + m2[3 + i] * dp[1] + m2[6 + i] * dp[2]; nl[i] = m2[i] * n[0] + m2[3 + i] * n[1] + m2[6 + i] * n[2]; }
        std::vector<T> vx(3 * nvt), support(nvt), mask(nvt);
        T smax = T(-1e30);
        for (int k = 0; k < nvt; k++) {
          for (int i = 0; i < 3; i++) vx[3 * k + i] = F(m->vert, 3 * (va + k) + i);
          T d[3] = {pl[0] - vx[3 * k], pl[1] - vx[3 * k + 1], pl[2] - vx[3 * k + 2]};
          support[k] = dot3(d, nl);
          if (support[k] > smax) smax = support[k];
        }
        const T thr = Max(T(0), smax - T(1e-3));
        for (int k = 0; k < nvt; k++) mask[k] = (support[k] > thr) ? T(0) : T(-1e6);
        auto argmax = [&](auto score) { int best = 0; T bv = score(0); for (int k = 1; k < nvt; k++) { T v = score(k); if (v > bv) { bv = v; best = k; } } return best; };
        const int ia = argmax([&](int k) { return mask[k]; });
        const T* a = &vx[3 * ia];
        const int ib = argmax([&](int k) { T d[3] = {a[0] - vx[3 * k], a[1] - vx[3 * k + 1], a[2] - vx[3 * k + 2]}; return dot3(d, d) + mask[k]; });
        const T* b = &vx[3 * ib];
        T amb[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]}, ab[3];
        cross(nl, amb, ab);
        const int ic = argmax([&](int k) { T d[3] = {a[0] - vx[3 * k], a[1] - vx[3 * k + 1], a[2] - vx[3 * k + 2]}; return Abs(dot3(d, ab)) + mask[k]; });
        const T* cc = &vx[3 * ic];
        T amc[3] = {a[0] - cc[0], a[1] - cc[1], a[2] - cc[2]}, bmc[3] = {b[0] - cc[0], b[1] - cc[1], b[2] - cc[2]}, ac[3], bc[3];
        cross(nl, amc, ac);
        cross(nl, bmc, bc);
        // concatenate([dist_bp, dist_ap]).argmax() % nvert: the b-edge scores come first
        int id = 0;
        {
          T bv = T(-1e30);
          for (int h = 0; h < 2; h++)
            for (int k = 0; k < nvt; k++) {
              const T* o = h == 0 ? b : a;
              const T* ax = h == 0 ? bc : ac;
              T d[3] = {o[0] - vx[3 * k], o[1] - vx[3 * k + 1], o[2] - vx[3 * k + 2]};
              T v = Abs(dot3(d, ax)) + mask[k];
              if (v > bv) { bv = v; id = k; }
            }
        }
        const int idx[4] = {ia, ib, ic, id};
        T fr[9];
        make_frame(n, fr);
        for (int q = 0; q < 4; q++) {
          bool unique = true;
          for (int j = 0; j < q; j++) if (idx[j] == idx[q]) unique = false;
          const T dist = unique ? -support[idx[q]] : T(1);
          const T* v = &vx[3 * idx[q]];
          for (int i = 0; i < 3; i++) {
            const T wp = p2[i] + m2[3 * i] * v[0] + m2[3 * i + 1] * v[1] + m2[3 * i + 2] * v[2];
            con_pos[3 * c + i] = wp - T(0.5) * dist * n[i];
          }
          con_dist[c] = dist;
          for (int i = 0; i < 9; i++) con_frame[9 * c + i] = fr[i];
          con_pair[c++] = p;
        }
      } else if (kind == ABR_PAIR_SPHERE_CONVEX) {
        // collision_convex.sphere_convex [MEMORY, MJX 3.1.x]: in the convex geom's frame. The face with the least penetration among
        // those the sphere reaches behind; the closest point of that polygon to the sphere centre; one contact.
        T dp[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, sp[3];
        for (int i = 0; i < 3; i++) sp[i] = m2[i] * dp[0] + m2[3 + i] * dp[1] + m2[6 + i] * dp[2];
        const T rad = F(m->geom_size, 3 * g1);
        const int fa = m->geom_faceadr[g2], nf = m->geom_facenum[g2], va = m->geom_vertadr[g2];
        int best = 0;
        T bestv = T(-1e30);
        for (int f = 0; f < nf; f++) {
          T nrm[3], v0[3], d[3];
          const int k0 = m->face_vert[m->face_vertadr[fa + f]];
          for (int i = 0; i < 3; i++) { nrm[i] = F(m->face_normal, 3 * (fa + f) + i); v0[i] = F(m->vert, 3 * (va + k0) + i); d[i] = sp[i] - v0[i]; }
          T sup = dot3(d, nrm) - rad;
          if (sup >= T(0)) sup = T(-1e12);
          if (sup > bestv) { bestv = sup; best = f; }
        }
        const int pa = m->face_vertadr[fa + best], pn = m->face_vertnum[fa + best];
        std::vector<T> P(3 * pn);
        T nrm[3];
        for (int i = 0; i < 3; i++) nrm[i] = F(m->face_normal, 3 * (fa + best) + i);
        for (int k = 0; k < pn; k++) for (int i = 0; i < 3; i++) P[3 * k + i] = F(m->vert, 3 * (va + m->face_vert[pa + k]) + i);
        T pt[3];
        project_pt_onto_plane(sp, &P[0], nrm, pt);
        bool inside = true;
        int eidx = 0;
        T ebest = T(1e30);
        for (int k = 0; k < pn; k++) {
          const T* e0 = &P[3 * ((k + pn - 1) % pn)]; const T* e1 = &P[3 * k];
          T ed[3] = {e1[0] - e0[0], e1[1] - e0[1], e1[2] - e0[2]}, en[3], d[3] = {pt[0] - e0[0], pt[1] - e0[1], pt[2] - e0[2]};
          cross(ed, nrm, en);
          T dist = dot3(d, en);
          if (!(dist <= T(0))) inside = false;
          const bool degenerate = en[0] == T(0) && en[1] == T(0) && en[2] == T(0);
          if (degenerate || dist < T(0)) dist = T(1e12);
          if (dist < ebest) { ebest = dist; eidx = k; }
        }
        if (!inside) closest_segment_point(&P[3 * ((eidx + pn - 1) % pn)], &P[3 * eidx], pt, pt);
        T n[3] = {pt[0] - sp[0], pt[1] - sp[1], pt[2] - sp[2]};
        const T d = normalize(n, 3);
        T pl[3];
        for (int i = 0; i < 3; i++) pl[i] = T(0.5) * (pt[i] + sp[i] + n[i] * rad);
        T nw[3];
        for (int i = 0; i < 3; i++) {
          nw[i] = m2[3 * i] * n[0] + m2[3 * i + 1] * n[1] + m2[3 * i + 2] * n[2];
          con_pos[3 * c + i] = p2[i] + m2[3 * i] * pl[0] + m2[3 * i + 1] * pl[1] + m2[3 * i + 2] * pl[2];
        }
        con_dist[c] = d - rad;
        make_frame(nw, &con_frame[9 * c]);
        con_pair[c++] = p;
      } else if (kind == ABR_PAIR_CAPSULE_CONVEX) {
        // collision_convex.capsule_convex [MEMORY, MJX 3.1.x]: in the convex geom's frame. The capsule's axis segment clipped to the
        // side planes of the best face gives two face contacts; a face edge closer to the axis than the radius replaces the first
        // one by an edge contact (and switches the second off).
        const T rad = F(m->geom_size, 3 * g1), half = F(m->geom_size, 3 * g1 + 1);
        T dp[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, cp[3], ax[3];
        const T axw[3] = {m1[2], m1[5], m1[8]};
        for (int i = 0; i < 3; i++) { cp[i] = m2[i] * dp[0] + m2[3 + i] * dp[1] + m2[6 + i] * dp[2]; ax[i] = m2[i] * axw[0] + m2[3 + i] * axw[1] + m2[6 + i] * axw[2]; }
        T c0[3], c1[3];
        for (int i = 0; i < 3; i++) { c0[i] = cp[i] - ax[i] * half; c1[i] = cp[i] + ax[i] * half; }
        const int fa = m->geom_faceadr[g2], nf = m->geom_facenum[g2], va = m->geom_vertadr[g2];
        int best = 0;
        T bestv = T(-1e30);
        bool has_support = true;
        for (int f = 0; f < nf; f++) {
          T nrm[3], v0[3], d0[3], d1[3];
          const int k0 = m->face_vert[m->face_vertadr[fa + f]];
          for (int i = 0; i < 3; i++) { nrm[i] = F(m->face_normal, 3 * (fa + f) + i); v0[i] = F(m->vert, 3 * (va + k0) + i); d0[i] = c0[i] - v0[i]; d1[i] = c1[i] - v0[i]; }
          T sup = Min(dot3(d0, nrm), dot3(d1, nrm)) - rad;
          if (!(sup < T(0))) has_support = false;
          if (sup >= T(0)) sup = T(-1e12);
          if (sup > bestv) { bestv = sup; best = f; }
        }
        const int pa = m->face_vertadr[fa + best], pn = m->face_vertnum[fa + best];
        std::vector<T> P(3 * pn), E0(3 * pn), EN(3 * pn);
        T nrm[3];
        for (int i = 0; i < 3; i++) nrm[i] = F(m->face_normal, 3 * (fa + best) + i);
        for (int k = 0; k < pn; k++) for (int i = 0; i < 3; i++) P[3 * k + i] = F(m->vert, 3 * (va + m->face_vert[pa + k]) + i);
        for (int k = 0; k < pn; k++) {
          const T* e0 = &P[3 * ((k + pn - 1) % pn)]; const T* e1 = &P[3 * k];
          T ed[3] = {e1[0] - e0[0], e1[1] - e0[1], e1[2] - e0[2]};
          cross(ed, nrm, &EN[3 * k]);
          for (int i = 0; i < 3; i++) E0[3 * k + i] = e0[i];
        }
        T q[2][3];
        bool ok;
        clip_edge_to_planes(c0, c1, E0.data(), EN.data(), pn, q[0], q[1], &ok);
        T pos[2][3], nl[2][3], dist[2];
        for (int k = 0; k < 2; k++) {
          T fp[3];
          for (int i = 0; i < 3; i++) q[k][i] = q[k][i] - nrm[i] * rad;
          project_pt_onto_plane(q[k], &P[0], nrm, fp);
          T df[3] = {fp[0] - q[k][0], fp[1] - q[k][1], fp[2] - q[k][2]};
          const T pen = (ok && has_support) ? dot3(df, nrm) : T(-1);
          dist[k] = -pen;
          for (int i = 0; i < 3; i++) { pos[k][i] = T(0.5) * (q[k][i] + fp[i]); nl[k][i] = -nrm[i]; }
        }
        // edge contact: the face edge closest to the capsule's axis
        T ebest = T(1e30), ec[3] = {T(0), T(0), T(0)}, cc[3] = {T(0), T(0), T(0)};
        for (int k = 0; k < pn; k++) {
          T a[3], b[3];
          closest_segment_to_segment_points(&E0[3 * k], &P[3 * k], c0, c1, a, b);
          T d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
          const T dd = dot3(d, d);
          if (dd < ebest) { ebest = dd; for (int i = 0; i < 3; i++) { ec[i] = a[i]; cc[i] = b[i]; } }
        }
        T ea[3] = {cc[0] - ec[0], cc[1] - ec[1], cc[2] - ec[2]};
        const T edist = normalize(ea, 3);
        if (rad - edist > T(0)) {
          for (int i = 0; i < 3; i++) { pos[0][i] = T(0.5) * (ec[i] + cc[i] - ea[i] * rad); nl[0][i] = -ea[i]; }
          dist[0] = -(rad - edist);
          dist[1] = T(1);
        }
        for (int k = 0; k < 2; k++) {
          T nw[3];
          for (int i = 0; i < 3; i++) {
            nw[i] = m2[3 * i] * nl[k][0] + m2[3 * i + 1] * nl[k][1] + m2[3 * i + 2] * nl[k][2];
            con_pos[3 * c + i] = p2[i] + m2[3 * i] * pos[k][0] + m2[3 * i + 1] * pos[k][1] + m2[3 * i + 2] * pos[k][2];
          }
          con_dist[c] = dist[k];
          make_frame(nw, &con_frame[9 * c]);
          con_pair[c++] = p;
        }
      } else if (kind == ABR_PAIR_CONVEX_CONVEX) {
        // collision_convex.convex_convex [MEMORY, MJX 3.1.x]: separating-axis test over the face normals of both hulls and the
        // cross products of their edges (world frame), then either the incident face clipped against the reference face (up to
        // four contacts, _create_contact_manifold) or the closest points of the two edges (one contact). Face axes are preferred
        // to an edge axis that is not better by more than 1e-5 m (a deviation: see DESIGN.md).
        const int nva = m->geom_vertnum[g1], nvb = m->geom_vertnum[g2];
        std::vector<T> VA(3 * nva), VB(3 * nvb);
        auto to_world = [&](int g, const T* pg, const T* mg, int k, T* out) {
          const int va = m->geom_vertadr[g];
          T v[3] = {F(m->vert, 3 * (va + k)), F(m->vert, 3 * (va + k) + 1), F(m->vert, 3 * (va + k) + 2)};
          for (int i = 0; i < 3; i++) out[i] = pg[i] + mg[3 * i] * v[0] + mg[3 * i + 1] * v[1] + mg[3 * i + 2] * v[2];
        };
        for (int k = 0; k < nva; k++) to_world(g1, p1, m1, k, &VA[3 * k]);
        for (int k = 0; k < nvb; k++) to_world(g2, p2, m2, k, &VB[3 * k]);
        const int fa = m->geom_faceadr[g1], nfa = m->geom_facenum[g1], fb = m->geom_faceadr[g2], nfb = m->geom_facenum[g2];
        const int ea = m->geom_edgeadr[g1], nea = m->geom_edgenum[g1], eb = m->geom_edgeadr[g2], neb = m->geom_edgenum[g2];
        auto face_normal_w = [&](int f, const T* mg, T* out) {
          T nl[3] = {F(m->face_normal, 3 * f), F(m->face_normal, 3 * f + 1), F(m->face_normal, 3 * f + 2)};
          for (int i = 0; i < 3; i++) out[i] = mg[3 * i] * nl[0] + mg[3 * i + 1] * nl[1] + mg[3 * i + 2] * nl[2];
        };
        auto axis_dist = [&](const T* axis, T* sign) {
          T amax = T(-1e30), amin = T(1e30), bmax = T(-1e30), bmin = T(1e30);
          for (int k = 0; k < nva; k++) { const T v = dot3(axis, &VA[3 * k]); amax = Max(amax, v); amin = Min(amin, v); }
          for (int k = 0; k < nvb; k++) { const T v = dot3(axis, &VB[3 * k]); bmax = Max(bmax, v); bmin = Min(bmin, v); }
          const T d1 = amax - bmin, d2 = bmax - amin;
          *sign = (d1 > d2) ? T(-1) : T(1);
          return Min(d1, d2);
        };
        T fbest = T(1e30), fsign = T(1), faxis[3] = {T(0), T(0), T(1)};
        for (int f = 0; f < nfa + nfb; f++) {
          T ax[3], sg;
          if (f < nfa) face_normal_w(fa + f, m1, ax); else face_normal_w(fb + f - nfa, m2, ax);
          const T d = axis_dist(ax, &sg);
          if (d < fbest) { fbest = d; fsign = sg; for (int i = 0; i < 3; i++) faxis[i] = ax[i]; }
        }
        T ebest = T(1e30), epair = T(-1e30), esign = T(1), eaxis[3] = {T(0), T(0), T(1)};
        int ei = 0, ej = 0;
        for (int j = 0; j < neb; j++)
          for (int i = 0; i < nea; i++) {
            const T* a0 = &VA[3 * m->edge_vert[2 * (ea + i)]]; const T* a1 = &VA[3 * m->edge_vert[2 * (ea + i) + 1]];
            const T* b0 = &VB[3 * m->edge_vert[2 * (eb + j)]]; const T* b1 = &VB[3 * m->edge_vert[2 * (eb + j) + 1]];
            T da[3] = {a0[0] - a1[0], a0[1] - a1[1], a0[2] - a1[2]}, db[3] = {b0[0] - b1[0], b0[1] - b1[1], b0[2] - b1[2]};
            normalize(da, 3);
            normalize(db, 3);
            T ax[3];
            cross(da, db, ax);
            if (dot3(ax, ax) < T(1e-6)) continue;  // (nearly) parallel edges span no axis
            normalize(ax, 3);
            T sg;
            const T d = axis_dist(ax, &sg);
            // parallel edges (a box has four per direction) span the same axis and tie on it: among those the SUPPORTING pair is
            // taken, the one whose own separation along the axis equals the hulls' (a deviation: see DESIGN.md)
            T ma[3] = {a0[0] + a1[0], a0[1] + a1[1], a0[2] + a1[2]}, mb[3] = {b0[0] + b1[0], b0[1] + b1[1], b0[2] + b1[2]};
            const T dpair = sg * T(0.5) * (dot3(ax, ma) - dot3(ax, mb));
            if (d < ebest - T(1e-6) || (d < ebest + T(1e-6) && dpair > epair)) {
              ebest = Min(ebest, d); epair = dpair; esign = sg; ei = i; ej = j;
              for (int k = 0; k < 3; k++) eaxis[k] = ax[k];
            }
          }
        const bool edge_contact = ebest < fbest - T(1e-5);
        T nab[3];  // from geom 1 to geom 2
        for (int i = 0; i < 3; i++) nab[i] = edge_contact ? esign * eaxis[i] : fsign * faxis[i];
        T cdist[4] = {T(1), T(1), T(1), T(1)}, cpos[4][3];
        for (int k = 0; k < 4; k++) for (int i = 0; i < 3; i++) cpos[k][i] = T(0.5) * (p1[i] + p2[i]);
        if (edge_contact) {
          T ca[3], cb[3];
          closest_segment_to_segment_points(&VA[3 * m->edge_vert[2 * (ea + ei)]], &VA[3 * m->edge_vert[2 * (ea + ei) + 1]],
                                            &VB[3 * m->edge_vert[2 * (eb + ej)]], &VB[3 * m->edge_vert[2 * (eb + ej) + 1]], ca, cb);
          T d[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]};
          cdist[0] = dot3(d, nab);
          for (int i = 0; i < 3; i++) cpos[0][i] = T(0.5) * (ca[i] + cb[i]);
        } else {
          // the faces most aligned with the axis; the better aligned one is the reference (clipping) face
          int ia = 0, ib = 0;
          T da = T(-1e30), db = T(-1e30);
          for (int f = 0; f < nfa; f++) { T nn[3]; face_normal_w(fa + f, m1, nn); const T v = dot3(nn, nab); if (v > da) { da = v; ia = f; } }
          for (int f = 0; f < nfb; f++) { T nn[3]; face_normal_w(fb + f, m2, nn); const T v = -dot3(nn, nab); if (v > db) { db = v; ib = f; } }
          const bool ref_a = da >= db;
          const int rf = ref_a ? fa + ia : fb + ib, sf = ref_a ? fb + ib : fa + ia;
          const T* RV = ref_a ? VA.data() : VB.data(); const T* SV = ref_a ? VB.data() : VA.data();
          T rn[3], sn[3];
          face_normal_w(rf, ref_a ? m1 : m2, rn);
          face_normal_w(sf, ref_a ? m2 : m1, sn);
          const int nr = m->face_vertnum[rf], ns = m->face_vertnum[sf];
          std::vector<T> RP(3 * nr), SP(3 * ns), RE0(3 * nr), REN(3 * nr), SE0(3 * ns), SEN(3 * ns);
          for (int k = 0; k < nr; k++) for (int i = 0; i < 3; i++) RP[3 * k + i] = RV[3 * m->face_vert[m->face_vertadr[rf] + k] + i];
          for (int k = 0; k < ns; k++) for (int i = 0; i < 3; i++) SP[3 * k + i] = SV[3 * m->face_vert[m->face_vertadr[sf] + k] + i];
          auto side_planes = [&](const std::vector<T>& Pp, int n, const T* nn, std::vector<T>& E0, std::vector<T>& EN) {
            for (int k = 0; k < n; k++) {
              const T* e0 = &Pp[3 * ((k + n - 1) % n)]; const T* e1 = &Pp[3 * k];
              T ed[3] = {e1[0] - e0[0], e1[1] - e0[1], e1[2] - e0[2]};
              cross(ed, nn, &EN[3 * k]);  // outward for a counter-clockwise polygon
              for (int i = 0; i < 3; i++) E0[3 * k + i] = e0[i];
            }
          };
          side_planes(RP, nr, rn, RE0, REN);
          side_planes(SP, ns, sn, SE0, SEN);
          // collision_convex._clip: subject edges against the reference side planes, then the reference edges (carried onto the
          // subject plane along the reference normal) against the subject's side planes
          const int npt = 2 * (ns + nr);
          std::vector<T> inc(3 * npt), ref(3 * npt);
          std::vector<char> mk(npt);
          for (int k = 0; k < ns; k++) {
            bool ok;
            clip_edge_to_planes(&SE0[3 * k], &SP[3 * k], RE0.data(), REN.data(), nr, &inc[3 * (2 * k)], &inc[3 * (2 * k + 1)], &ok);
            mk[2 * k] = mk[2 * k + 1] = ok;
          }
          {
            const T dpl = dot3(&SP[0], sn), denom = dot3(rn, sn);
            std::vector<T> Q0(3 * nr), Q1(3 * nr);
            for (int k = 0; k < nr; k++)
              for (int h = 0; h < 2; h++) {
                const T* src = h == 0 ? &RE0[3 * k] : &RP[3 * k];
                const T t = (dpl - dot3(src, sn)) / (denom + ((denom == T(0)) ? T(1e-6) : T(0)));
                T* dst = h == 0 ? &Q0[3 * k] : &Q1[3 * k];
                for (int i = 0; i < 3; i++) dst[i] = src[i] + t * rn[i];
              }
            for (int k = 0; k < nr; k++) {
              bool ok;
              clip_edge_to_planes(&Q0[3 * k], &Q1[3 * k], SE0.data(), SEN.data(), ns, &inc[3 * (2 * (ns + k))], &inc[3 * (2 * (ns + k) + 1)], &ok);
              mk[2 * (ns + k)] = mk[2 * (ns + k) + 1] = ok;
            }
          }
          std::vector<char> mask(npt);
          std::unique_ptr<bool[]> mb(new bool[npt]);
          for (int k = 0; k < npt; k++) {
            project_pt_onto_plane(&inc[3 * k], &RP[0], rn, &ref[3 * k]);
            T d[3] = {inc[3 * k] - RP[0], inc[3 * k + 1] - RP[1], inc[3 * k + 2] - RP[2]};
            mb[k] = mk[k] && (-dot3(d, rn) > T(1e-6));  // the incident point lies behind the reference face
          }
          int idx[4];
          manifold_points(ref.data(), mb.get(), npt, rn, idx);
          for (int q = 0; q < 4; q++) {
            bool unique = true;
            for (int j = 0; j < q; j++) if (idx[j] == idx[q]) unique = false;
            T d[3] = {inc[3 * idx[q]] - ref[3 * idx[q]], inc[3 * idx[q] + 1] - ref[3 * idx[q] + 1], inc[3 * idx[q] + 2] - ref[3 * idx[q] + 2]};
            const T pen = -dot3(d, rn);
            cdist[q] = (mb[idx[q]] && unique) ? -pen : T(1);
            for (int i = 0; i < 3; i++) cpos[q][i] = T(0.5) * (inc[3 * idx[q] + i] + ref[3 * idx[q] + i]);
          }
        }
        T fr[9];
        make_frame(nab, fr);
        for (int q = 0; q < 4; q++) {
          con_dist[c] = cdist[q];
          for (int i = 0; i < 3; i++) con_pos[3 * c + i] = cpos[q][i];
          for (int i = 0; i < 9; i++) con_frame[9 * c + i] = fr[i];
          con_pair[c++] = p;
        }
      } else if (kind == ABR_PAIR_SPHERE_SPHERE) {
        T n[3];
        for (int i = 0; i < 3; i++) n[i] = p2[i] - p1[i];
        T dist = normalize(n, 3);
        if (dist == T(0)) { n[0] = T(1); n[1] = T(0); n[2] = T(0); }
        T r1 = F(m->geom_size, 3 * g1), r2 = F(m->geom_size, 3 * g2);
        dist = dist - (r1 + r2);
        con_dist[c] = dist;
        for (int i = 0; i < 3; i++) con_pos[3 * c + i] = p1[i] + n[i] * (r1 + dist * T(0.5));
        make_frame(n, &con_frame[9 * c]);
        con_pair[c++] = p;
      } else {
        // sphere-capsule / capsule-capsule: closest points on segments, then sphere-sphere
        T r1 = F(m->geom_size, 3 * g1), r2 = F(m->geom_size, 3 * g2);
        T a1[3], a2[3];
        if (kind == ABR_PAIR_SPHERE_CAPSULE) {
          T axis[3] = {m2[2], m2[5], m2[8]};
          T half = F(m->geom_size, 3 * g2 + 1);
          T d[3];
          for (int i = 0; i < 3; i++) d[i] = p1[i] - p2[i];
          T t = Clip(dot3(d, axis), -half, half);
          for (int i = 0; i < 3; i++) { a1[i] = p1[i]; a2[i] = p2[i] + axis[i] * t; }
        } else {
          T ax1[3] = {m1[2], m1[5], m1[8]}, ax2[3] = {m2[2], m2[5], m2[8]};
          T h1 = F(m->geom_size, 3 * g1 + 1), h2 = F(m->geom_size, 3 * g2 + 1);
          // math.closest_segment_to_segment_points
          T d[3];
          for (int i = 0; i < 3; i++) d[i] = p1[i] - p2[i];
          T dab = dot3(ax1, ax2), d1 = dot3(d, ax1), d2 = dot3(d, ax2);
          T den = T(1) - dab * dab;
          T t1 = (den == T(0)) ? T(0) : (dab * d2 - d1) / den;
          t1 = Clip(t1, -h1, h1);
          T t2 = Clip(t1 * dab + d2, -h2, h2);
          t1 = Clip(t2 * dab - d1, -h1, h1);
          for (int i = 0; i < 3; i++) { a1[i] = p1[i] + ax1[i] * t1; a2[i] = p2[i] + ax2[i] * t2; }
        }
        T n[3];
        for (int i = 0; i < 3; i++) n[i] = a2[i] - a1[i];
        T dist = normalize(n, 3);
        if (dist == T(0)) { n[0] = T(1); n[1] = T(0); n[2] = T(0); }
        dist = dist - (r1 + r2);
        con_dist[c] = dist;
        for (int i = 0; i < 3; i++) con_pos[3 * c + i] = a1[i] + n[i] * (r1 + dist * T(0.5));
        make_frame(n, &con_frame[9 * c]);
        con_pair[c++] = p;
      }
    }
  }

  // ---- constraint._kbi + row finalisation
  void finish_row(int r, T pos, T invweight, const float* solref, const float* solimp, bool active) {
    if (!active) {  // MJX multiplies the whole row by `active`: it stays inert
      for (int k = 0; k < nv; k++) efc_J[r * nv + k] = T(0);
      efc_D[r] = T(0); efc_aref[r] = T(0); efc_pos[r] = T(0);
      return;
    }
    T timeconst = T(double(solref[0])), dampratio = T(double(solref[1]));
    if (!(m->opt.disableflags & ABR_DSBL_REFSAFE)) timeconst = Max(timeconst, T(2) * T(double(m->opt.timestep)));
    T dmin = Clip(T(double(solimp[0])), T(kMinImp), T(kMaxImp));
    T dmax = Clip(T(double(solimp[1])), T(kMinImp), T(kMaxImp));
    T width = Max(T(kMinVal), T(double(solimp[2])));
    T mid = Clip(T(double(solimp[3])), T(kMinImp), T(kMaxImp));
    T power = Max(T(1), T(double(solimp[4])));
    T k = T(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
    T b = T(2) / (dmax * timeconst);
    if (solref[0] <= 0) k = -T(double(solref[0])) / (dmax * dmax);
    if (solref[1] <= 0) b = -T(double(solref[1])) / dmax;
    T imp_x = Abs(pos) / width;
    T imp_a = (T(1) / Pow(mid, power - T(1))) * Pow(imp_x, power);
    T imp_b = T(1) - (T(1) / Pow(T(1) - mid, power - T(1))) * Pow(T(1) - imp_x, power);
    T imp_y = (imp_x < mid) ? imp_a : imp_b;
    T imp = dmin + imp_y * (dmax - dmin);
    imp = Clip(imp, dmin, dmax);
    if (imp_x > T(1)) imp = dmax;
    T R = Max(invweight * (T(1) - imp) / imp, T(kMinVal));
    T jv = T(0);
    for (int d = 0; d < nv; d++) jv += efc_J[r * nv + d] * qvel[d];
    efc_D[r] = T(1) / R;
    efc_aref[r] = -b * jv - k * imp * pos;
    efc_pos[r] = pos;
  }

  // support.jac (translational part), dense over all dofs with ancestor mask
  void jacp_body(const T* point, int body, T* jac /*[nv,3]*/) const {
    for (int i = 0; i < nv * 3; i++) jac[i] = T(0);
    // ancestor mask: dofs of bodies on the path body -> root
    std::vector<char> mask(nbody, 0);
    for (int b = body; b > 0; b = m->body_parentid[b]) mask[b] = 1;
    const T* rc = &subtree_com[3 * m->body_rootid[body]];
    T off[3];
    for (int i = 0; i < 3; i++) off[i] = point[i] - rc[i];
    for (int d = 0; d < nv; d++) {
      if (!mask[m->dof_bodyid[d]]) continue;
      T c[3];
      cross(&cdof[6 * d], off, c);
      for (int i = 0; i < 3; i++) jac[3 * d + i] = cdof[6 * d + 3 + i] + c[i];
    }
  }

  // ---- constraint.make_constraint
  void make_constraint() {
    if (sz.nefc == 0) return;
    int dis = m->opt.disableflags;
    int r = 0;
    for (int i = 0; i < sz.nefc * nv; i++) efc_J[i] = T(0);
    if (!(dis & ABR_DSBL_EQUALITY)) {
      for (int e = 0; e < m->neq; e++) {
        if (!m->eq_active[e]) continue;
        int j1 = m->eq_obj1id[e], j2 = m->eq_obj2id[e];
        const float* data = m->eq_data + 11 * e;
        int a1 = m->jnt_qposadr[j1], d1 = m->jnt_dofadr[j1];
        T pos1 = qpos[a1] - F(m->qpos0, a1);
        T pos, inv;
        if (j2 >= 0) {
          int a2 = m->jnt_qposadr[j2], d2 = m->jnt_dofadr[j2];
          T dif = qpos[a2] - F(m->qpos0, a2);
          T pw[5];
          pw[0] = T(1);
          for (int k = 1; k < 5; k++) pw[k] = pw[k - 1] * dif;
          T poly = T(0), deriv = T(0);
          for (int k = 0; k < 5; k++) poly += T(double(data[k])) * pw[k];
          for (int k = 1; k < 5; k++) deriv += T(double(data[k])) * pw[k - 1] * T(double(k));
          pos = pos1 - poly;
          efc_J[r * nv + d2] = -deriv;
          efc_J[r * nv + d1] = T(1);
          inv = F(m->dof_invweight0, d1) + F(m->dof_invweight0, d2);
        } else {
          pos = pos1 - T(double(data[0]));
          efc_J[r * nv + d1] = T(1);
          inv = F(m->dof_invweight0, d1);
        }
        finish_row(r, pos, inv, m->eq_solref + 2 * e, m->eq_solimp + 5 * e, true);
        r++;
      }
    }
    if (!(dis & ABR_DSBL_LIMIT)) {
      for (int j = 0; j < njnt; j++) {
        if (!m->jnt_limited[j]) continue;
        int type = m->jnt_type[j];
        if (type != ABR_JNT_HINGE && type != ABR_JNT_SLIDE) continue;
        int a = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
        T dmin = qpos[a] - F(m->jnt_range, 2 * j), dmax = F(m->jnt_range, 2 * j + 1) - qpos[a];
        T pos = Min(dmin, dmax) - F(m->jnt_margin, j);
        bool active = pos < T(0);
        efc_J[r * nv + d] = (dmin < dmax) ? T(1) : T(-1);
        finish_row(r, pos, F(m->dof_invweight0, d), m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, active);
        r++;
      }
    }
    if (!(dis & ABR_DSBL_CONTACT)) {
      std::vector<T> j1(nv * 3), j2(nv * 3), diff(3 * nv);
      for (int c = 0; c < sz.ncon; c++) {
        int p = con_pair[c];
        int b1 = m->geom_bodyid[m->pair_geom1[p]], b2 = m->geom_bodyid[m->pair_geom2[p]];
        T dist = con_dist[c] - F(m->pair_includemargin, p);
        bool active = dist < T(0);
        jacp_body(&con_pos[3 * c], b1, j1.data());
        jacp_body(&con_pos[3 * c], b2, j2.data());
        // diff_con = frame @ (jac2 - jac1)^T : 3 x nv
        const T* fr = &con_frame[9 * c];
        for (int k = 0; k < 3; k++)
          for (int d = 0; d < nv; d++) {
            T s = T(0);
            for (int i = 0; i < 3; i++) s += fr[3 * k + i] * (j2[3 * d + i] - j1[3 * d + i]);
            diff[k * nv + d] = s;
          }
        T t = F(m->body_invweight0, 2 * b1) + F(m->body_invweight0, 2 * b2);
        if (m->pair_condim[p] == 1) {
          for (int d = 0; d < nv; d++) efc_J[r * nv + d] = diff[d];
          finish_row(r, dist, t, m->pair_solref + 2 * p, m->pair_solimp + 5 * p, active);
          r++;
        } else {
          for (int k = 0; k < 2; k++) {
            T fri = F(m->pair_friction, 5 * p + k);
            for (int s = 0; s < 2; s++) {
              T f = (s == 0) ? fri : -fri;
              for (int d = 0; d < nv; d++) efc_J[r * nv + d] = diff[d] + diff[(1 + k) * nv + d] * f;
              T inv = (t + f * f * t) * T(2) * f * f / T(double(m->opt.impratio));
              finish_row(r, dist, inv, m->pair_solref + 2 * p, m->pair_solimp + 5 * p, active);
              r++;
            }
          }
        }
      }
    }
  }

  // ---- smooth.transmission (joint transmissions, hinge/slide)
  void transmission() {
    for (int u = 0; u < nu; u++) {
      int j = m->actuator_trnid[u];
      act_length[u] = qpos[m->jnt_qposadr[j]] * F(m->actuator_gear, u);
    }
  }

  void fwd_position() {
    kinematics();
    com_pos();
    crb_and_factor();
    collision();
    make_constraint();
    transmission();
  }

  // ---- fwd_velocity: actuator_velocity, com_vel, passive, rne
  void fwd_velocity() {
    for (int u = 0; u < nu; u++) {
      int j = m->actuator_trnid[u];
      act_velocity[u] = F(m->actuator_gear, u) * qvel[m->jnt_dofadr[j]];
    }
    // com_vel
    for (int b = 0; b < nbody; b++) {
      T cv[6];
      int p = m->body_parentid[b];
      for (int i = 0; i < 6; i++) cv[i] = (b == 0) ? T(0) : cvel[6 * p + i];
      for (int k = 0; k < m->body_jntnum[b]; k++) {
        int j = m->body_jntadr[b] + k;
        int d = m->jnt_dofadr[j];
        if (m->jnt_type[j] == ABR_JNT_FREE) {
          for (int q = 0; q < 3; q++)
            for (int i = 0; i < 6; i++) cv[i] += cdof[6 * (d + q) + i] * qvel[d + q];
          for (int q = 0; q < 3; q++) {
            for (int i = 0; i < 6; i++) cdof_dot[6 * (d + q) + i] = T(0);
            motion_cross(cv, &cdof[6 * (d + 3 + q)], &cdof_dot[6 * (d + 3 + q)]);
          }
          for (int q = 3; q < 6; q++)
            for (int i = 0; i < 6; i++) cv[i] += cdof[6 * (d + q) + i] * qvel[d + q];
        } else {
          motion_cross(cv, &cdof[6 * d], &cdof_dot[6 * d]);
          for (int i = 0; i < 6; i++) cv[i] += cdof[6 * d + i] * qvel[d];
        }
      }
      for (int i = 0; i < 6; i++) cvel[6 * b + i] = cv[i];
    }
    // passive
    for (int d = 0; d < nv; d++) qfrc_passive[d] = T(0);
    if (!(m->opt.disableflags & ABR_DSBL_PASSIVE)) {
      for (int j = 0; j < njnt; j++) {
        int type = m->jnt_type[j];
        if (type == ABR_JNT_HINGE || type == ABR_JNT_SLIDE) {
          int a = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
          qfrc_passive[d] = -F(m->jnt_stiffness, j) * (qpos[a] - F(m->qpos_spring, a));
        }
      }
      for (int d = 0; d < nv; d++) qfrc_passive[d] -= F(m->dof_damping, d) * qvel[d];
    }
    // rne
    std::vector<T> cacc(nbody * 6), cfrc(nbody * 6);
    for (int b = 0; b < nbody; b++) {
      T* a = &cacc[6 * b];
      if (b == 0) {
        for (int i = 0; i < 3; i++) {
          a[i] = T(0);
          a[3 + i] = (m->opt.disableflags & ABR_DSBL_GRAVITY) ? T(0) : -T(double(m->opt.gravity[i]));
        }
      } else {
        int p = m->body_parentid[b];
        for (int i = 0; i < 6; i++) a[i] = cacc[6 * p + i];
      }
      for (int k = 0; k < m->body_dofnum[b]; k++) {
        int d = m->body_dofadr[b] + k;
        for (int i = 0; i < 6; i++) a[i] += cdof_dot[6 * d + i] * qvel[d];
      }
    }
    for (int b = 0; b < nbody; b++) {
      T f1[6], f2[6], f3[6];
      inert_mul(&cinert[10 * b], &cacc[6 * b], f1);
      inert_mul(&cinert[10 * b], &cvel[6 * b], f2);
      motion_cross_force(&cvel[6 * b], f2, f3);
      for (int i = 0; i < 6; i++) cfrc[6 * b + i] = f1[i] + f3[i];
    }
    for (int b = nbody - 1; b > 0; b--) {
      int p = m->body_parentid[b];
      for (int i = 0; i < 6; i++) cfrc[6 * p + i] += cfrc[6 * b + i];
    }
    for (int d = 0; d < nv; d++) {
      T s = T(0);
      const T* f = &cfrc[6 * m->dof_bodyid[d]];
      for (int i = 0; i < 6; i++) s += cdof[6 * d + i] * f[i];
      qfrc_bias[d] = s;
    }
  }

  // ---- forward.fwd_actuation
  void fwd_actuation() {
    for (int d = 0; d < nv; d++) qfrc_actuator[d] = T(0);
    if (nu == 0 || (m->opt.disableflags & ABR_DSBL_ACTUATION)) {
      for (int u = 0; u < nu; u++) act_force[u] = T(0);
      return;
    }
    for (int u = 0; u < nu; u++) {
      T c = ctrl[u];
      if (!(m->opt.disableflags & ABR_DSBL_CLAMPCTRL) && m->actuator_ctrllimited[u])
        c = Clip(c, F(m->actuator_ctrlrange, 2 * u), F(m->actuator_ctrlrange, 2 * u + 1));
      T gain = F(m->actuator_gainprm, 3 * u);
      if (m->actuator_gaintype[u] == ABR_GAIN_AFFINE)
        gain = gain + F(m->actuator_gainprm, 3 * u + 1) * act_length[u] + F(m->actuator_gainprm, 3 * u + 2) * act_velocity[u];
      T bias = T(0);
      if (m->actuator_biastype[u] == ABR_BIAS_AFFINE)
        bias = F(m->actuator_biasprm, 3 * u) + F(m->actuator_biasprm, 3 * u + 1) * act_length[u] +
               F(m->actuator_biasprm, 3 * u + 2) * act_velocity[u];
      T f = gain * c + bias;
      if (m->actuator_forcelimited[u])
        f = Clip(f, F(m->actuator_forcerange, 2 * u), F(m->actuator_forcerange, 2 * u + 1));
      act_force[u] = f;
      int j = m->actuator_trnid[u];
      qfrc_actuator[m->jnt_dofadr[j]] += F(m->actuator_gear, u) * f;
    }
  }

  void fwd_acceleration() {
    for (int d = 0; d < nv; d++) qfrc_smooth[d] = qfrc_passive[d] - qfrc_bias[d] + qfrc_actuator[d];
    cho_solve(qL.data(), qfrc_smooth.data(), qacc_smooth.data(), nv);
  }

  // ---- solver.py
  struct Ctx {
    std::vector<T> qacc, qfrc_constraint, Jaref, efc_force, Ma, grad, Mgrad, search;
    T gauss, cost, prev_cost;
    int niter;
  };
  struct LSPoint { T alpha, cost, d0, d1; };

  bool row_always_active(int r) const { return r < sz.ne; }

  void update_constraint(Ctx& c) {
    const int nefc = sz.nefc;
    T s = T(0);
    for (int r = 0; r < nefc; r++) {
      bool act = row_always_active(r) || c.Jaref[r] < T(0);
      c.efc_force[r] = act ? efc_D[r] * -c.Jaref[r] : T(0);
      if (act) s += efc_D[r] * c.Jaref[r] * c.Jaref[r];
    }
    for (int d = 0; d < nv; d++) {
      T q = T(0);
      for (int r = 0; r < nefc; r++) q += efc_J[r * nv + d] * c.efc_force[r];
      c.qfrc_constraint[d] = q;
    }
    T g = T(0);
    for (int d = 0; d < nv; d++) g += (c.Ma[d] - qfrc_smooth[d]) * (c.qacc[d] - qacc_smooth[d]);
    c.gauss = T(0.5) * g;
    c.prev_cost = c.cost;
    c.cost = T(0.5) * s + c.gauss;
  }

  void update_gradient(Ctx& c) {
    const int nefc = sz.nefc;
    for (int d = 0; d < nv; d++) c.grad[d] = c.Ma[d] - qfrc_smooth[d] - c.qfrc_constraint[d];
    if (m->opt.solver == ABR_SOLVER_CG) {
      cho_solve(qL.data(), c.grad.data(), c.Mgrad.data(), nv);
    } else {
      std::vector<T> H(qM), L(nv * nv);
      for (int r = 0; r < nefc; r++) {
        bool act = row_always_active(r) || c.Jaref[r] < T(0);
        if (!act) continue;
        for (int i = 0; i < nv; i++) {
          T ji = efc_J[r * nv + i] * efc_D[r];
          for (int j = 0; j < nv; j++) H[i * nv + j] += ji * efc_J[r * nv + j];
        }
      }
      cholesky(H.data(), L.data(), nv);
      cho_solve(L.data(), c.grad.data(), c.Mgrad.data(), nv);
    }
  }

  void ctx_create(Ctx& c, const std::vector<T>& a, bool grad) {
    const int nefc = sz.nefc;
    c.qacc = a;
    c.qfrc_constraint.assign(nv, T(0)); c.efc_force.assign(nefc, T(0));
    c.Jaref.assign(nefc, T(0)); c.Ma.assign(nv, T(0));
    c.grad.assign(nv, T(0)); c.Mgrad.assign(nv, T(0)); c.search.assign(nv, T(0));
    for (int r = 0; r < nefc; r++) {
      T s = T(0);
      for (int d = 0; d < nv; d++) s += efc_J[r * nv + d] * a[d];
      c.Jaref[r] = s - efc_aref[r];
    }
    mul_m(a.data(), c.Ma.data());
    c.gauss = T(0); c.cost = T(INFINITY); c.prev_cost = T(0); c.niter = 0;
    update_constraint(c);
    if (grad) {
      update_gradient(c);
      for (int d = 0; d < nv; d++) c.search[d] = -c.Mgrad[d];
    }
  }

  LSPoint ls_point(const Ctx& c, T alpha, const std::vector<T>& jv, const std::vector<T>& quad, const T* qg) {
    const int nefc = sz.nefc;
    T q0 = qg[0], q1 = qg[1], q2 = qg[2];
    for (int r = 0; r < nefc; r++) {
      T x = c.Jaref[r] + alpha * jv[r];
      bool act = row_always_active(r) || x < T(0);
      if (act) { q0 += quad[3 * r]; q1 += quad[3 * r + 1]; q2 += quad[3 * r + 2]; }
    }
    LSPoint p;
    p.alpha = alpha;
    p.cost = alpha * alpha * q2 + alpha * q1 + q0;
    p.d0 = T(2) * alpha * q2 + q1;
    p.d1 = T(2) * q2 + ((q2 == T(0)) ? T(kMinVal) : T(0));
    return p;
  }
  static T safe_div(T a, T b) { return a / (b + ((b == T(0)) ? T(kMinVal) : T(0))); }

  void linesearch(Ctx& c) {
    const int nefc = sz.nefc;
    T sn = T(0);
    for (int d = 0; d < nv; d++) sn += c.search[d] * c.search[d];
    T smag = Sqrt(sn) * T(double(m->opt.meaninertia)) * T(double(std::max(1, nv)));
    T gtol = T(double(m->opt.tolerance)) * T(double(m->opt.ls_tolerance)) * smag;
    std::vector<T> mv(nv), jv(nefc), quad(3 * nefc);
    mul_m(c.search.data(), mv.data());
    for (int r = 0; r < nefc; r++) {
      T s = T(0);
      for (int d = 0; d < nv; d++) s += efc_J[r * nv + d] * c.search[d];
      jv[r] = s;
    }
    T qg[3];
    T sMa = T(0), sq = T(0), smv = T(0);
    for (int d = 0; d < nv; d++) { sMa += c.search[d] * c.Ma[d]; sq += c.search[d] * qfrc_smooth[d]; smv += c.search[d] * mv[d]; }
    qg[0] = c.gauss; qg[1] = sMa - sq; qg[2] = T(0.5) * smv;
    for (int r = 0; r < nefc; r++) {
      quad[3 * r] = T(0.5) * c.Jaref[r] * c.Jaref[r] * efc_D[r];
      quad[3 * r + 1] = jv[r] * c.Jaref[r] * efc_D[r];
      quad[3 * r + 2] = T(0.5) * jv[r] * jv[r] * efc_D[r];
    }
    LSPoint p0 = ls_point(c, T(0), jv, quad, qg);
    LSPoint lo0 = ls_point(c, -safe_div(p0.d0, p0.d1), jv, quad, qg);
    bool lesser = lo0.d0 < p0.d0;
    LSPoint hi = lesser ? p0 : lo0;
    LSPoint lo = lesser ? lo0 : p0;
    bool swap = true;
    int it = 0;
    while (true) {
      bool done = it >= m->opt.ls_iterations;
      done = done || !swap;
      done = done || ((lo.d0 < T(0)) && (lo.d0 > -gtol));
      done = done || ((hi.d0 > T(0)) && (hi.d0 < gtol));
      if (done) break;
      LSPoint lo_next = ls_point(c, lo.alpha - safe_div(lo.d0, lo.d1), jv, quad, qg);
      LSPoint hi_next = ls_point(c, hi.alpha - safe_div(hi.d0, hi.d1), jv, quad, qg);
      LSPoint mid = ls_point(c, T(0.5) * (lo.alpha + hi.alpha), jv, quad, qg);
      bool swap_lo_next = (lo.d0 > T(0)) || (lo.d0 < lo_next.d0);
      if (swap_lo_next) lo = lo_next;
      bool swap_lo_mid = (mid.d0 < T(0)) && (lo.d0 < mid.d0);
      if (swap_lo_mid) lo = mid;
      bool swap_hi_next = (hi.d0 < T(0)) || (hi.d0 > hi_next.d0);
      if (swap_hi_next) hi = hi_next;
      bool swap_hi_mid = (mid.d0 > T(0)) && (hi.d0 > mid.d0);
      if (swap_hi_mid) hi = mid;
      swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
      it++;
    }
    bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
    T alpha = (lo.cost < hi.cost) ? lo.alpha : hi.alpha;
    if (improved) {
      for (int d = 0; d < nv; d++) { c.qacc[d] += c.search[d] * alpha; c.Ma[d] += mv[d] * alpha; }
      for (int r = 0; r < nefc; r++) c.Jaref[r] += jv[r] * alpha;
    }
  }

  T rescale(T v) const { return v / (T(double(m->opt.meaninertia)) * T(double(std::max(1, nv)))); }

  void solver_body(Ctx& c, bool need_post) {
    linesearch(c);
    std::vector<T> prev_grad = c.grad, prev_Mgrad = c.Mgrad;
    update_constraint(c);
    if (need_post) {
      update_gradient(c);
      if (m->opt.solver == ABR_SOLVER_NEWTON) {
        for (int d = 0; d < nv; d++) c.search[d] = -c.Mgrad[d];
      } else {
        T num = T(0), den = T(0);
        for (int d = 0; d < nv; d++) { num += c.grad[d] * (c.Mgrad[d] - prev_Mgrad[d]); den += prev_grad[d] * prev_Mgrad[d]; }
        T beta = Max(T(0), num / Max(T(kMinVal), den));
        for (int d = 0; d < nv; d++) c.search[d] = -c.Mgrad[d] + beta * c.search[d];
      }
    }
    c.niter++;
  }

  void solve() {
    std::vector<T> start = qacc_smooth;
    if (!(m->opt.disableflags & ABR_DSBL_WARMSTART)) {
      Ctx warm, smth;
      ctx_create(warm, qacc_warmstart, false);
      ctx_create(smth, qacc_smooth, false);
      if (warm.cost < smth.cost) start = qacc_warmstart;
    }
    Ctx c;
    ctx_create(c, start, true);
    if (m->opt.iterations == 1) {
      // MJX also recomputes gradient/Hessian after the single iteration; results do not depend on it
      solver_body(c, literal_post);
    } else {
      while (true) {
        T improvement = rescale(c.prev_cost - c.cost);
        T gn = T(0);
        for (int d = 0; d < nv; d++) gn += c.grad[d] * c.grad[d];
        T gradient = rescale(Sqrt(gn));
        bool done = c.niter >= m->opt.iterations;
        done = done || improvement < T(double(m->opt.tolerance));
        done = done || gradient < T(double(m->opt.tolerance));
        if (done) break;
        solver_body(c, true);
      }
    }
    qacc = c.qacc;
    qacc_warmstart = c.qacc;
    qfrc_constraint = c.qfrc_constraint;
    efc_force = c.efc_force;
    solver_niter = c.niter;
  }

  // ---- forward.forward
  void forward() {
    fwd_position();
    fwd_velocity();
    fwd_actuation();
    fwd_acceleration();
    if (sz.nefc == 0) {
      qacc = qacc_smooth;
      for (int d = 0; d < nv; d++) qfrc_constraint[d] = T(0);
      return;
    }
    solve();
  }

  // forward._integrate_pos
  void integrate_pos(std::vector<T>& q, const std::vector<T>& v, T dt) const {
    for (int j = 0; j < njnt; j++) {
      int a = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
      if (m->jnt_type[j] == ABR_JNT_FREE) {
        for (int i = 0; i < 3; i++) q[a + i] = q[a + i] + dt * v[d + i];
        T w[3] = {v[d + 3], v[d + 4], v[d + 5]};
        T nrm = normalize(w, 3);
        T ql[4], qn[4];
        axis_angle_to_quat(w, dt * nrm, ql);
        quat_mul(&q[a + 3], ql, qn);
        normalize(qn, 4);
        for (int i = 0; i < 4; i++) q[a + 3 + i] = qn[i];
      } else {
        q[a] = q[a] + dt * v[d];
      }
    }
  }

  // forward._advance
  void advance(const std::vector<T>& qacc_in, const std::vector<T>* qvel_in) {
    T dt = T(double(m->opt.timestep));
    for (int d = 0; d < nv; d++) qvel[d] = qvel[d] + qacc_in[d] * dt;
    integrate_pos(qpos, qvel_in ? *qvel_in : qvel, dt);
    time = time + dt;
  }

  void euler() {
    std::vector<T> a = qacc;
    if (!(m->opt.disableflags & ABR_DSBL_EULERDAMP)) {
      T dt = T(double(m->opt.timestep));
      std::vector<T> Mh(qM), L(nv * nv), f(nv);
      for (int d = 0; d < nv; d++) Mh[d * nv + d] += F(m->dof_damping, d) * dt;
      cholesky(Mh.data(), L.data(), nv);
      for (int d = 0; d < nv; d++) f[d] = qfrc_smooth[d] + qfrc_constraint[d];
      cho_solve(L.data(), f.data(), a.data(), nv);
    }
    advance(a, nullptr);
  }

  void rungekutta4() {
    static const double A[3] = {0.5, 0.5, 1.0};
    static const double B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    T dt = T(double(m->opt.timestep));
    std::vector<T> qpos0 = qpos, qvel0 = qvel, warm0 = qacc_warmstart;
    T time0 = time;
    std::vector<T> kqvel = qvel, sv(nv), sa(nv);
    for (int d = 0; d < nv; d++) { sv[d] = T(B[0]) * kqvel[d]; sa[d] = T(B[0]) * qacc[d]; }
    T c = T(0);
    for (int s = 0; s < 3; s++) {
      std::vector<T> dqvel(nv), dqacc(nv);
      for (int d = 0; d < nv; d++) { dqvel[d] = T(A[s]) * kqvel[d]; dqacc[d] = T(A[s]) * qacc[d]; }
      std::vector<T> kqpos = qpos0;
      integrate_pos(kqpos, dqvel, dt);
      for (int d = 0; d < nv; d++) kqvel[d] = qvel0[d] + dqacc[d] * dt;
      qpos = kqpos; qvel = kqvel;
      c = T(A[s]);
      time = time0 + c * dt;
      forward();
      for (int d = 0; d < nv; d++) { sv[d] += T(B[s + 1]) * kqvel[d]; sa[d] += T(B[s + 1]) * qacc[d]; }
    }
    qpos = qpos0; qvel = qvel0; qacc_warmstart = warm0; time = time0;
    advance(sa, &sv);
  }

  void step() {
    forward();
    if (m->opt.integrator == ABR_INT_RK4) rungekutta4(); else euler();
  }
};

// ----------------------------------------------------------------------------- C interface
thread_local std::map<std::string, std::vector<double>> g_dump;

template <class T> void dump_vec(const char* name, const std::vector<T>& v) {
  std::vector<double> d(v.size());
  for (size_t i = 0; i < v.size(); i++) d[i] = ToD(v[i]);
  g_dump[name] = d;
}

template <class T> void set_state(Sim<T>& s, const double* qpos, const double* qvel, const double* ctrl, const double* warm) {
  for (int i = 0; i < s.nq; i++) s.qpos[i] = T(qpos[i]);
  for (int i = 0; i < s.nv; i++) s.qvel[i] = T(qvel[i]);
  for (int i = 0; i < s.nu; i++) s.ctrl[i] = ctrl ? T(ctrl[i]) : T(0);
  for (int i = 0; i < s.nv; i++) s.qacc_warmstart[i] = warm ? T(warm[i]) : T(0);
}

template <class T> int forward_dump(const AbrModelHost* m, const double* qpos, const double* qvel, const double* ctrl, const double* warm) {
  Sim<T> s(m);
  set_state(s, qpos, qvel, ctrl, warm);
  s.forward();
  g_dump.clear();
  dump_vec("qpos", s.qpos); dump_vec("xpos", s.xpos); dump_vec("xquat", s.xquat); dump_vec("xmat", s.xmat);
  dump_vec("xipos", s.xipos); dump_vec("ximat", s.ximat); dump_vec("xanchor", s.xanchor); dump_vec("xaxis", s.xaxis);
  dump_vec("geom_xpos", s.geom_xpos); dump_vec("geom_xmat", s.geom_xmat);
  dump_vec("subtree_com", s.subtree_com); dump_vec("cinert", s.cinert); dump_vec("cdof", s.cdof); dump_vec("crb", s.crb);
  dump_vec("qM", s.qM); dump_vec("qLD", s.qL); dump_vec("contact_dist", s.con_dist); dump_vec("contact_pos", s.con_pos);
  dump_vec("contact_frame", s.con_frame); dump_vec("efc_J", s.efc_J); dump_vec("efc_D", s.efc_D); dump_vec("efc_aref", s.efc_aref);
  dump_vec("efc_pos", s.efc_pos);
  dump_vec("actuator_length", s.act_length); dump_vec("actuator_velocity", s.act_velocity); dump_vec("actuator_force", s.act_force);
  dump_vec("qfrc_actuator", s.qfrc_actuator); dump_vec("cvel", s.cvel); dump_vec("cdof_dot", s.cdof_dot);
  dump_vec("qfrc_passive", s.qfrc_passive); dump_vec("qfrc_bias", s.qfrc_bias); dump_vec("qfrc_smooth", s.qfrc_smooth);
  dump_vec("qacc_smooth", s.qacc_smooth); dump_vec("qacc", s.qacc); dump_vec("qfrc_constraint", s.qfrc_constraint);
  dump_vec("efc_force", s.efc_force); dump_vec("qacc_warmstart", s.qacc_warmstart);
  g_dump["solver_niter"] = {double(s.solver_niter)};
  return 0;
}

// shoot(): make_data, set x0, forward (ctrl=0), then N steps (shooting.py:22-48)
template <class T> void rollout_one(const AbrModelHost* m, const double* x0, const double* us, int N, double* xs, double* xfinal) {
  Sim<T> s(m);
  const int nq = m->nq, nv = m->nv, nu = m->nu, nx = nq + nv;
  for (int i = 0; i < nq; i++) s.qpos[i] = T(x0[i]);
  for (int i = 0; i < nv; i++) s.qvel[i] = T(x0[nq + i]);
  s.forward();
  if (xs) for (int i = 0; i < nx; i++) xs[i] = x0[i];
  for (int t = 0; t < N; t++) {
    for (int u = 0; u < nu; u++) s.ctrl[u] = T(us[t * nu + u]);
    s.step();
    if (xs) {
      for (int i = 0; i < nq; i++) xs[(t + 1) * nx + i] = ToD(s.qpos[i]);
      for (int i = 0; i < nv; i++) xs[(t + 1) * nx + nq + i] = ToD(s.qvel[i]);
    }
  }
  if (xfinal) {
    for (int i = 0; i < nq; i++) xfinal[i] = ToD(s.qpos[i]);
    for (int i = 0; i < nv; i++) xfinal[nq + i] = ToD(s.qvel[i]);
  }
}

}  // namespace

extern "C" {

int orc_sizes(const AbrModelHost* m, int* ncon, int* ne, int* nl, int* nefc) {
  Sizes s = compute_sizes(m);
  *ncon = s.ncon; *ne = s.ne; *nl = s.nl; *nefc = s.nefc;
  return 0;
}

// mjx.forward on one world; intermediates retrievable with orc_get. prec: 0 = float64, 1 = float32
int orc_forward(const AbrModelHost* m, int prec, const double* qpos, const double* qvel, const double* ctrl, const double* warm) {
  return prec == 0 ? forward_dump<double>(m, qpos, qvel, ctrl, warm) : forward_dump<float>(m, qpos, qvel, ctrl, warm);
}

int orc_get(const char* name, double* out, int cap, int* n) {
  auto it = g_dump.find(name);
  if (it == g_dump.end()) return -1;
  *n = int(it->second.size());
  if (*n > cap) return -2;
  std::memcpy(out, it->second.data(), sizeof(double) * it->second.size());
  return 0;
}

// nsteps x mjx.step in place with ctrl held (rl/base.py:88-96)
int orc_step(const AbrModelHost* m, int prec, double* qpos, double* qvel, double* warm, double* time, const double* ctrl, int nsteps) {
  auto run = [&](auto tag) {
    using T = decltype(tag);
    Sim<T> s(m);
    set_state(s, qpos, qvel, ctrl, warm);
    s.time = T(*time);
    for (int k = 0; k < nsteps; k++) s.step();
    for (int i = 0; i < s.nq; i++) qpos[i] = ToD(s.qpos[i]);
    for (int i = 0; i < s.nv; i++) { qvel[i] = ToD(s.qvel[i]); warm[i] = ToD(s.qacc_warmstart[i]); }
    *time = ToD(s.time);
  };
  if (prec == 0) run(double(0)); else run(float(0));
  return 0;
}

// shoot for nworld worlds, optionally threaded (CPU baseline). xs nullable [nworld,N+1,nx];
// xfinal nullable [nworld,nx].
int orc_rollout_batch(const AbrModelHost* m, int prec, const double* x0, int x0_stride, const double* us, int us_stride,
                      int nworld, int N, double* xs, double* xfinal, int nthreads) {
  const int nx = m->nq + m->nv;
  auto work = [&](int w0, int w1) {
    for (int w = w0; w < w1; w++) {
      const double* x = x0 + size_t(w) * x0_stride;
      const double* u = us + size_t(w) * us_stride;
      double* xo = xs ? xs + size_t(w) * (N + 1) * nx : nullptr;
      double* xf = xfinal ? xfinal + size_t(w) * nx : nullptr;
      if (prec == 0) rollout_one<double>(m, x, u, N, xo, xf); else rollout_one<float>(m, x, u, N, xo, xf);
    }
  };
  if (nthreads <= 1) { work(0, nworld); return 0; }
  std::vector<std::thread> th;
  int per = (nworld + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    int a = t * per, b = std::min(nworld, a + per);
    if (a < b) th.emplace_back(work, a, b);
  }
  for (auto& t : th) t.join();
  return 0;
}

// algorithmic FLOPs of one mjx.step from the given state (add=sub=mul=div=sqrt=sin=cos=pow=1,
// compare/select/abs/neg=0; an FMA is a mul + an add = 2). literal=1 also counts MJX's unused
// post-iteration gradient/Hessian when iterations==1.
long long orc_count_flops_step(const AbrModelHost* m, const double* qpos, const double* qvel, const double* ctrl,
                               const double* warm, int literal) {
  Sim<Cnt> s(m);
  s.literal_post = literal != 0;
  set_state(s, qpos, qvel, ctrl, warm);
  Cnt::n = 0;
  s.step();
  return Cnt::n;
}

}  // extern "C"
