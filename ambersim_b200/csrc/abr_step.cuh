// abr_step.cuh — the fused world-step: every stage of mjx.forward / mjx.step as a device
// function executed by a group of G lanes (G in {4,8,16,32}, one or several worlds per warp).
// All per-world state lives in shared memory (Layout::w_* offsets); lanes stride over the
// stage's natural items (bodies of one tree level, dofs, packed matrix entries, constraint
// rows) and meet at __syncwarp(). Nothing here touches global memory.
//
// Replaces the third-party mjx.step behind ambersim/trajopt/shooting.py:41 and
// ambersim/rl/base.py:93; stage order and formulas follow SURVEY.md Appendix A (MJX 3.0.1-3.1.x).
#ifndef ABR_STEP_CUH_
#define ABR_STEP_CUH_

#include <cuda_runtime.h>
#include <math.h>

#include "abr_layout.h"
#include "abr.h"

namespace abr {

#define ABR_FULL 0xffffffffu
#define MF(name) (c.mf + c.L.f_##name)
#define MI(name) (c.mi + c.L.i_##name)
#define WF(name) (c.W + c.L.w_##name)

constexpr float kMinVal = 1e-15f;

struct Ctx {
  const Layout& L;
  const float* mf;  // model float pool (shared memory)
  const int* mi;    // model int pool (shared memory)
  float* W;         // this world's shared-memory region
  int lane;         // lane within the group
  float frs = 1.f;  // per-world domain randomisation: contact friction scale ...
  float acs = 1.f;  // ... actuator strength scale,
  float dps = 1.f;  // ... joint damping scale
  float ars = 1.f;  // ... and joint armature scale (abr_env_set_randomization_ex)
};

// ------------------------------------------------------------------------------ small math
__device__ __forceinline__ void v_cross(const float* a, const float* b, float* r) {
  float r0 = a[1] * b[2] - a[2] * b[1];
  float r1 = a[2] * b[0] - a[0] * b[2];
  float r2 = a[0] * b[1] - a[1] * b[0];
  r[0] = r0; r[1] = r1; r[2] = r2;
}
__device__ __forceinline__ float v_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void q_mul(const float* u, const float* v, float* r) {
  float r0 = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  float r1 = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  float r2 = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  float r3 = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3;
}
// MJX math.rotate
__device__ __forceinline__ void q_rot(const float* v, const float* q, float* r) {
  const float s = q[0];
  const float* u = q + 1;
  float uv = v_dot(u, v), uu = v_dot(u, u);
  float cx[3];
  v_cross(u, v, cx);
  float k = s * s - uu;
  float o0 = 2.f * (uv * u[0]) + k * v[0] + 2.f * s * cx[0];
  float o1 = 2.f * (uv * u[1]) + k * v[1] + 2.f * s * cx[1];
  float o2 = 2.f * (uv * u[2]) + k * v[2] + 2.f * s * cx[2];
  r[0] = o0; r[1] = o1; r[2] = o2;
}
__device__ __forceinline__ void q_to_mat(const float* q, float* m) {
  float q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  float q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  float q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2.f * (q12 - q03);      m[2] = 2.f * (q13 + q02);
  m[3] = 2.f * (q12 + q03);      m[4] = q00 - q11 + q22 - q33; m[5] = 2.f * (q23 - q01);
  m[6] = 2.f * (q13 - q02);      m[7] = 2.f * (q23 + q01);      m[8] = q00 - q11 - q22 + q33;
}
__device__ __forceinline__ void axis_angle_quat(const float* axis, float angle, float* q) {
  float s, cs;
  sincosf(angle * 0.5f, &s, &cs);
  q[0] = cs; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
__device__ __forceinline__ float v_normalize(float* x, int n) {
  float s = 0.f;
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  float nrm = sqrtf(s);
  float den = (nrm == 0.f) ? 1e-6f : nrm;
  for (int i = 0; i < n; i++) x[i] = x[i] / den;
  return nrm;
}
// cinert = (Ixx,Iyy,Izz,Ixy,Ixz,Iyz, m*off[3], m)
__device__ __forceinline__ void inert_mul(const float* I, const float* v, float* r) {
  // angular: I3 w + (m c) x v_lin ; linear: m v_lin - (m c) x w. One FMA chain per component (no separate cross + add).
  const float r0 = fmaf(-I[8], v[4], fmaf(I[7], v[5], fmaf(I[4], v[2], fmaf(I[3], v[1], I[0] * v[0]))));
  const float r1 = fmaf(-I[6], v[5], fmaf(I[8], v[3], fmaf(I[5], v[2], fmaf(I[1], v[1], I[3] * v[0]))));
  const float r2 = fmaf(-I[7], v[3], fmaf(I[6], v[4], fmaf(I[2], v[2], fmaf(I[5], v[1], I[4] * v[0]))));
  const float l0 = fmaf(I[8], v[1], fmaf(-I[7], v[2], I[9] * v[3]));
  const float l1 = fmaf(I[6], v[2], fmaf(-I[8], v[0], I[9] * v[4]));
  const float l2 = fmaf(I[7], v[0], fmaf(-I[6], v[1], I[9] * v[5]));
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = l0; r[4] = l1; r[5] = l2;
}
__device__ __forceinline__ void motion_cross(const float* u, const float* v, float* r) {
  float a[3], b[3], cc[3];
  v_cross(u, v, a);
  v_cross(u + 3, v, b);
  v_cross(u, v + 3, cc);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + cc[0]; r[4] = b[1] + cc[1]; r[5] = b[2] + cc[2];
}
__device__ __forceinline__ void motion_cross_force(const float* v, const float* f, float* r) {
  float a[3], b[3], cc[3];
  v_cross(v, f, a);
  v_cross(v + 3, f + 3, b);
  v_cross(v, f + 3, cc);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = cc[0]; r[4] = cc[1]; r[5] = cc[2];
}
__device__ __forceinline__ void make_frame(const float* a_in, float* fr) {
  float a[3] = {a_in[0], a_in[1], a_in[2]};
  v_normalize(a, 3);
  float b[3] = {0.f, 0.f, 0.f};
  if (-0.5f < a[1] && a[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
  float ab = v_dot(a, b);
  b[0] -= a[0] * ab; b[1] -= a[1] * ab; b[2] -= a[2] * ab;
  v_normalize(b, 3);
  float cc[3];
  v_cross(a, b, cc);
  for (int i = 0; i < 3; i++) { fr[i] = a[i]; fr[3 + i] = b[i]; fr[6 + i] = cc[i]; }
}

template <int G> __device__ __forceinline__ float gsum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(ABR_FULL, v, o);
  return v;
}
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// ------------------------------------------------------------------------------ kinematics
// smooth.kinematics: root -> leaves by tree level; one lane per body of the level.
template <int G> __device__ void stage_kinematics(const Ctx& c) {
  const Layout& L = c.L;
  float* xpos = WF(xpos); float* xquat = WF(xquat); float* xipos = WF(xipos);
  float* xanchor = WF(xanchor); float* xaxis = WF(xaxis); float* qpos = WF(qpos);
  if (c.lane == 0) {  // world body (the pose region may be aliased, so re-seed every step)
    xpos[0] = 0.f; xpos[1] = 0.f; xpos[2] = 0.f;
    xquat[0] = 1.f; xquat[1] = 0.f; xquat[2] = 0.f; xquat[3] = 0.f;
  }
  __syncwarp();
  for (int lev = 1; lev <= L.depth; lev++) {
    const int beg = MI(level_adr)[lev], end = MI(level_adr)[lev + 1];
    for (int k = beg + c.lane; k < end; k += G) {
      const int b = MI(level_body)[k];
      const int p = MI(body_parent)[b];
      float pos[3], quat[4], r[3];
      q_rot(MF(body_pos) + 3 * b, xquat + 4 * p, r);
      pos[0] = xpos[3 * p] + r[0]; pos[1] = xpos[3 * p + 1] + r[1]; pos[2] = xpos[3 * p + 2] + r[2];
      q_mul(xquat + 4 * p, MF(body_quat) + 4 * b, quat);
      const int jn = MI(body_jntnum)[b], ja = MI(body_jntadr)[b];
      for (int jj = 0; jj < jn; jj++) {
        const int j = ja + jj;
        const int a = MI(jnt_qposadr)[j];
        const int type = MI(jnt_type)[j];
        float anchor[3], axis[3];
        if (type == ABR_JNT_FREE) {
          anchor[0] = qpos[a]; anchor[1] = qpos[a + 1]; anchor[2] = qpos[a + 2];
          axis[0] = 0.f; axis[1] = 0.f; axis[2] = 1.f;
          pos[0] = anchor[0]; pos[1] = anchor[1]; pos[2] = anchor[2];
          quat[0] = qpos[a + 3]; quat[1] = qpos[a + 4]; quat[2] = qpos[a + 5]; quat[3] = qpos[a + 6];
          v_normalize(quat, 4);
          qpos[a + 3] = quat[0]; qpos[a + 4] = quat[1]; qpos[a + 5] = quat[2]; qpos[a + 6] = quat[3];
        } else {
          const float* jp = MF(jnt_pos) + 3 * j;
          const float* jx = MF(jnt_axis) + 3 * j;
          q_rot(jp, quat, r);
          anchor[0] = r[0] + pos[0]; anchor[1] = r[1] + pos[1]; anchor[2] = r[2] + pos[2];
          q_rot(jx, quat, axis);
          const float dq = qpos[a] - MF(qpos0)[a];
          if (type == ABR_JNT_HINGE) {
            float ql[4], qn[4];
            axis_angle_quat(jx, dq, ql);
            q_mul(quat, ql, qn);
            quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
            q_rot(jp, quat, r);
            pos[0] = anchor[0] - r[0]; pos[1] = anchor[1] - r[1]; pos[2] = anchor[2] - r[2];
          } else {  // slide
            pos[0] += axis[0] * dq; pos[1] += axis[1] * dq; pos[2] += axis[2] * dq;
          }
        }
        for (int i = 0; i < 3; i++) { xanchor[3 * j + i] = anchor[i]; xaxis[3 * j + i] = axis[i]; }
      }
      for (int i = 0; i < 3; i++) xpos[3 * b + i] = pos[i];
      for (int i = 0; i < 4; i++) xquat[4 * b + i] = quat[i];
      q_rot(MF(body_ipos) + 3 * b, quat, r);
      for (int i = 0; i < 3; i++) xipos[3 * b + i] = pos[i] + r[i];
    }
    __syncwarp();
  }
}

// smooth.com_pos: subtree CoM of each kinematic root, cinert per body, cdof per dof.
template <int G> __device__ void stage_com(const Ctx& c) {
  const Layout& L = c.L;
  const float* xipos = WF(xipos);
  float* rootcom = WF(rootcom);
  for (int r = 0; r < L.nroot; r++) {
    float sx = 0.f, sy = 0.f, sz = 0.f, sm = 0.f;
    for (int b = 1 + c.lane; b < L.nbody; b += G) {
      if (MI(body_rootslot)[b] == r) {
        const float ms = MF(body_mass)[b];
        sx += xipos[3 * b] * ms; sy += xipos[3 * b + 1] * ms; sz += xipos[3 * b + 2] * ms; sm += ms;
      }
    }
    sx = gsum<G>(sx); sy = gsum<G>(sy); sz = gsum<G>(sz); sm = gsum<G>(sm);
    if (c.lane == 0) {
      const int rb = MI(root_body)[r];
      if (sm < kMinVal) { rootcom[3 * r] = xipos[3 * rb]; rootcom[3 * r + 1] = xipos[3 * rb + 1]; rootcom[3 * r + 2] = xipos[3 * rb + 2]; }
      else { rootcom[3 * r] = sx / sm; rootcom[3 * r + 1] = sy / sm; rootcom[3 * r + 2] = sz / sm; }
    }
  }
  __syncwarp();
  float* cinert = WF(cinert);
  const float* xquat = WF(xquat);
  if (c.lane < 10) cinert[c.lane] = 0.f;
  if (G < 10 && c.lane == 0) for (int i = 0; i < 10; i++) cinert[i] = 0.f;
  for (int b = 1 + c.lane; b < L.nbody; b += G) {
    const float* rc = rootcom + 3 * MI(body_rootslot)[b];
    float off[3] = {xipos[3 * b] - rc[0], xipos[3 * b + 1] - rc[1], xipos[3 * b + 2] - rc[2]};
    const float ms = MF(body_mass)[b];
    float q2[4], R[9];
    q_mul(xquat + 4 * b, MF(body_iquat) + 4 * b, q2);
    q_to_mat(q2, R);
    const float* in = MF(body_inertia) + 3 * b;
    float oo = v_dot(off, off);
    float I[6];  // xx yy zz xy xz yz
    const int ra[6] = {0, 1, 2, 0, 0, 1}, cb[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; e++) {
      const int r_ = ra[e], c_ = cb[e];
      float s = R[3 * r_] * in[0] * R[3 * c_] + R[3 * r_ + 1] * in[1] * R[3 * c_ + 1] + R[3 * r_ + 2] * in[2] * R[3 * c_ + 2];
      s += ms * ((r_ == c_ ? oo : 0.f) - off[r_] * off[c_]);
      I[e] = s;
    }
    float* ci = cinert + 10 * b;
    for (int e = 0; e < 6; e++) ci[e] = I[e];
    ci[6] = off[0] * ms; ci[7] = off[1] * ms; ci[8] = off[2] * ms; ci[9] = ms;
  }
  float* cdof = WF(cdof);
  const float* xanchor = WF(xanchor); const float* xaxis = WF(xaxis);
  for (int d = c.lane; d < L.nv; d += G) {
    const int j = MI(dof_jnt)[d];
    const int b = MI(jnt_body)[j];
    const int type = MI(jnt_type)[j];
    const float* rc = rootcom + 3 * MI(body_rootslot)[b];
    float off[3] = {rc[0] - xanchor[3 * j], rc[1] - xanchor[3 * j + 1], rc[2] - xanchor[3 * j + 2]};
    float* cd = cdof + 6 * d;
    if (type == ABR_JNT_FREE) {
      const int q = d - MI(jnt_dofadr)[j];
      if (q < 3) {
        for (int i = 0; i < 6; i++) cd[i] = (i == 3 + q) ? 1.f : 0.f;
      } else {
        float R[9], a[3], cr[3];
        q_to_mat(xquat + 4 * b, R);
        const int k = q - 3;
        a[0] = R[k]; a[1] = R[3 + k]; a[2] = R[6 + k];
        v_cross(a, off, cr);
        cd[0] = a[0]; cd[1] = a[1]; cd[2] = a[2]; cd[3] = cr[0]; cd[4] = cr[1]; cd[5] = cr[2];
      }
    } else if (type == ABR_JNT_HINGE) {
      float cr[3];
      v_cross(xaxis + 3 * j, off, cr);
      cd[0] = xaxis[3 * j]; cd[1] = xaxis[3 * j + 1]; cd[2] = xaxis[3 * j + 2];
      cd[3] = cr[0]; cd[4] = cr[1]; cd[5] = cr[2];
    } else {
      cd[0] = 0.f; cd[1] = 0.f; cd[2] = 0.f;
      cd[3] = xaxis[3 * j]; cd[4] = xaxis[3 * j + 1]; cd[5] = xaxis[3 * j + 2];
    }
  }
  __syncwarp();
}


// ------------------------------------------------------------------------------ convex collision primitives
// mjx math.py / collision_convex.py: the pieces of sphere - convex, capsule - convex and convex - convex collision.
// Everything works on small per-lane arrays: a polygon has at most ABR_MAX_FACE_VERTS corners.
__device__ __forceinline__ void closest_segment_point(const float* a, const float* b, const float* pt, float* out) {
  const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  const float t = fminf(fmaxf(v_dot(pa, ab) / (v_dot(ab, ab) + 1e-6f), 0.f), 1.f);
  for (int i = 0; i < 3; i++) out[i] = a[i] + t * ab[i];
}
__device__ inline void closest_segment_to_segment_points(const float* a0, const float* a1, const float* b0, const float* b1, float* best_a, float* best_b) {
  float da[3], db[3];
  for (int i = 0; i < 3; i++) { da[i] = a1[i] - a0[i]; db[i] = b1[i] - b0[i]; }
  const float ha = 0.5f * v_normalize(da, 3), hb = 0.5f * v_normalize(db, 3);
  float am[3], bm[3], tr[3];
  for (int i = 0; i < 3; i++) { am[i] = a0[i] + da[i] * ha; bm[i] = b0[i] + db[i] * hb; tr[i] = am[i] - bm[i]; }
  const float dd = v_dot(da, db), dat = v_dot(da, tr), dbt = v_dot(db, tr);
  const float ota = (-dat + dd * dbt) / (1.f - dd * dd + 1e-6f);
  const float otb = dbt + ota * dd;
  const float ta = fminf(fmaxf(ota, -ha), ha), tb = fminf(fmaxf(otb, -hb), hb);
  float pa[3], pb[3], na[3], nb[3];
  for (int i = 0; i < 3; i++) { pa[i] = am[i] + da[i] * ta; pb[i] = bm[i] + db[i] * tb; }
  closest_segment_point(a0, a1, pb, na);
  closest_segment_point(b0, b1, pa, nb);
  const float e1[3] = {na[0] - pb[0], na[1] - pb[1], na[2] - pb[2]}, e2[3] = {pa[0] - nb[0], pa[1] - nb[1], pa[2] - nb[2]};
  const bool first = v_dot(e1, e1) < v_dot(e2, e2);
  for (int i = 0; i < 3; i++) { best_a[i] = first ? na[i] : pa[i]; best_b[i] = first ? pb[i] : nb[i]; }
}
__device__ __forceinline__ void project_pt_onto_plane(const float* pt, const float* plane_pt, const float* n, float* out) {
  const float d[3] = {pt[0] - plane_pt[0], pt[1] - plane_pt[1], pt[2] - plane_pt[2]};
  const float dist = v_dot(d, n);
  for (int i = 0; i < 3; i++) out[i] = pt[i] - dist * n[i];
}
// unit vector with one reciprocal square root (the separating-axis loops normalise three vectors per edge pair; a zero vector stays zero)
__device__ __forceinline__ void v_unit3(float* x) {
  const float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  const float inv = (ss > 0.f) ? rsqrtf(ss) : 0.f;
  x[0] *= inv; x[1] *= inv; x[2] *= inv;
}
// _clip_edge_to_planes against the side planes of the counter-clockwise polygon P (np corners, normal nrm): plane k passes through
// corner k-1 with the outward normal (P[k] - P[k-1]) x nrm
__device__ inline bool clip_edge_to_poly(const float* p0, const float* p1, const float* P, int np, const float* nrm, float* out0, float* out1) {
  const float e01[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e10[3] = {-e01[0], -e01[1], -e01[2]};
  bool any_both = false;
  float best0 = -1e30f, best1 = -1e30f;
  float n0[3] = {p0[0], p0[1], p0[2]}, n1[3] = {p1[0], p1[1], p1[2]};
#pragma unroll 1
  for (int k = 0; k < np; k++) {
    const float* pp = P + 3 * ((k + np - 1) % np); const float* pe = P + 3 * k;
    const float ed[3] = {pe[0] - pp[0], pe[1] - pp[1], pe[2] - pp[2]};
    float pn[3];
    v_cross(ed, nrm, pn);
    const float d0[3] = {p0[0] - pp[0], p0[1] - pp[1], p0[2] - pp[2]}, d1[3] = {p1[0] - pp[0], p1[1] - pp[1], p1[2] - pp[2]};
    const bool f0 = v_dot(d0, pn) > 1e-6f, f1 = v_dot(d1, pn) > 1e-6f;
    any_both = any_both || (f0 && f1);
    // _closest_segment_point_plane
    const float dpl = v_dot(pp, pn), denom = v_dot(pn, e01);
    const float t = fminf(fmaxf((dpl - v_dot(pn, p0)) / (denom + ((denom == 0.f) ? 1e-6f : 0.f)), 0.f), 1.f);
    float c0[3], c1[3];
    for (int i = 0; i < 3; i++) { const float cp = p0[i] + t * e01[i]; c0[i] = f0 ? cp : p0[i]; c1[i] = f1 ? cp : p1[i]; }
    const float r0[3] = {c0[0] - p0[0], c0[1] - p0[1], c0[2] - p0[2]}, r1[3] = {c1[0] - p1[0], c1[1] - p1[1], c1[2] - p1[2]};
    const float s0 = v_dot(r0, e01), s1 = v_dot(r1, e10);
    if (s0 > best0) { best0 = s0; for (int i = 0; i < 3; i++) n0[i] = c0[i]; }
    if (s1 > best1) { best1 = s1; for (int i = 0; i < 3; i++) n1[i] = c1[i]; }
  }
  bool mask = !any_both;
  for (int i = 0; i < 3; i++) { out0[i] = mask ? n0[i] : p0[i]; out1[i] = mask ? n1[i] : p1[i]; }
  const float dn[3] = {out0[0] - out1[0], out0[1] - out1[1], out0[2] - out1[2]};
  if (v_dot(e10, dn) < 0.f) mask = false;
  return mask;
}
// _manifold_points over n points (stride 3) with a bit mask; first maximum wins
__device__ inline void manifold_points(const float* pts, unsigned mask, int n, const float* normal, int* idx) {
  auto dm = [&](int k) { return ((mask >> k) & 1u) ? 0.f : -1e6f; };
  int ia = 0, ib = 0, ic = 0, id = 0;
  float bv = -1e30f;
#pragma unroll 1
  for (int k = 0; k < n; k++) { const float v = dm(k); if (v > bv) { bv = v; ia = k; } }
  const float a[3] = {pts[3 * ia], pts[3 * ia + 1], pts[3 * ia + 2]};
  bv = -1e30f;
#pragma unroll 1
  for (int k = 0; k < n; k++) { const float d[3] = {a[0] - pts[3 * k], a[1] - pts[3 * k + 1], a[2] - pts[3 * k + 2]}; const float v = v_dot(d, d) + dm(k); if (v > bv) { bv = v; ib = k; } }
  const float b[3] = {pts[3 * ib], pts[3 * ib + 1], pts[3 * ib + 2]};
  const float amb[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
  float ab[3];
  v_cross(normal, amb, ab);
  bv = -1e30f;
#pragma unroll 1
  for (int k = 0; k < n; k++) { const float d[3] = {a[0] - pts[3 * k], a[1] - pts[3 * k + 1], a[2] - pts[3 * k + 2]}; const float v = fabsf(v_dot(d, ab)) + dm(k); if (v > bv) { bv = v; ic = k; } }
  const float cc[3] = {pts[3 * ic], pts[3 * ic + 1], pts[3 * ic + 2]};
  const float amc[3] = {a[0] - cc[0], a[1] - cc[1], a[2] - cc[2]}, bmc[3] = {b[0] - cc[0], b[1] - cc[1], b[2] - cc[2]};
  float ac[3], bc[3];
  v_cross(normal, amc, ac);
  v_cross(normal, bmc, bc);
  bv = -1e30f;
#pragma unroll 1
  for (int h = 0; h < 2; h++) {
#pragma unroll 1
    for (int k = 0; k < n; k++) {
      const float* o = h == 0 ? b : a;
      const float* ax = h == 0 ? bc : ac;
      const float d[3] = {o[0] - pts[3 * k], o[1] - pts[3 * k + 1], o[2] - pts[3 * k + 2]};
      const float v = fabsf(v_dot(d, ax)) + dm(k);
      if (v > bv) { bv = v; id = k; }
    }
  }
  idx[0] = ia; idx[1] = ib; idx[2] = ic; idx[3] = id;
}

// collision_driver.collision (static pairs, primitives) + contact Jacobian basis B[c][3][nv]
// (normal, tangent1, tangent2 rows of frame @ (jacp(body2) - jacp(body1))).
template <int G> __device__ void stage_collision(const Ctx& c) {
  const Layout& L = c.L;
  if (L.ncon == 0) return;
  const float* xpos = WF(xpos); const float* xquat = WF(xquat);
  float* cdist = WF(cdist); float* cpos = WF(cpos); float* cframe = WF(cframe);
  for (int ci = c.lane; ci < L.ncon; ci += G) {
    const int p = MI(con_pair)[ci], sub = MI(con_sub)[ci];
    const int g1 = MI(pair_g1)[p], g2 = MI(pair_g2)[p], kind = MI(pair_kind)[p];
    const int b1 = MI(geom_body)[g1], b2 = MI(geom_body)[g2];
    float p1[3], p2[3], q1[4], q2[4], r[3];
    q_rot(MF(geom_pos) + 3 * g1, xquat + 4 * b1, r);
    p1[0] = xpos[3 * b1] + r[0]; p1[1] = xpos[3 * b1 + 1] + r[1]; p1[2] = xpos[3 * b1 + 2] + r[2];
    q_rot(MF(geom_pos) + 3 * g2, xquat + 4 * b2, r);
    p2[0] = xpos[3 * b2] + r[0]; p2[1] = xpos[3 * b2 + 1] + r[1]; p2[2] = xpos[3 * b2 + 2] + r[2];
    q_mul(xquat + 4 * b1, MF(geom_quat) + 4 * g1, q1);
    q_mul(xquat + 4 * b2, MF(geom_quat) + 4 * g2, q2);
    float m1[9], m2[9];
    q_to_mat(q1, m1);
    q_to_mat(q2, m2);
    const float* s1 = MF(geom_size) + 3 * g1;
    const float* s2 = MF(geom_size) + 3 * g2;
    float dist, pos[3], fr[9];
    if (kind == ABR_PAIR_PLANE_CONVEX) {
      // mjx collision_convex.plane_convex: vertices in the convex geom's frame,
      // support = penetration depth, manifold of up to 4 vertices among those within 1 mm of the deepest (first maximum wins, like
      // jnp.argmax); a vertex picked twice yields an inactive contact (dist = 1). Each of the pair's 4 contact slots redoes the
      // selection (the slots run on different lanes) and keeps its own vertex.
      const int va = MI(geom_vertadr)[g2], nvt = MI(geom_vertnum)[g2];
      const float* V = MF(vert) + 3 * va;
      const float n[3] = {m1[2], m1[5], m1[8]};
      const float dp[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
      float pl[3], nl[3];
      for (int i = 0; i < 3; i++) { pl[i] = m2[i] * dp[0] + m2[3 + i] * dp[1] + m2[6 + i] * dp[2]; nl[i] = m2[i] * n[0] + m2[3 + i] * n[1] + m2[6 + i] * n[2]; }
      float smax = -1e30f;
#pragma unroll 1
      for (int k = 0; k < nvt; k++) { const float d[3] = {pl[0] - V[3 * k], pl[1] - V[3 * k + 1], pl[2] - V[3 * k + 2]}; smax = fmaxf(smax, v_dot(d, nl)); }
      const float thr = fmaxf(0.f, smax - 1e-3f);
      auto sup = [&](int k) { const float d[3] = {pl[0] - V[3 * k], pl[1] - V[3 * k + 1], pl[2] - V[3 * k + 2]}; return v_dot(d, nl); };
      auto msk = [&](int k) { return (sup(k) > thr) ? 0.f : -1e6f; };
      int idx[4] = {0, 0, 0, 0};
      float bv = -1e30f;
#pragma unroll 1
      for (int k = 0; k < nvt; k++) { const float v = msk(k); if (v > bv) { bv = v; idx[0] = k; } }
      const float a[3] = {V[3 * idx[0]], V[3 * idx[0] + 1], V[3 * idx[0] + 2]};
      bv = -1e30f;
#pragma unroll 1
      for (int k = 0; k < nvt; k++) {
        const float d[3] = {a[0] - V[3 * k], a[1] - V[3 * k + 1], a[2] - V[3 * k + 2]};
        const float v = v_dot(d, d) + msk(k);
        if (v > bv) { bv = v; idx[1] = k; }
      }
      const float b[3] = {V[3 * idx[1]], V[3 * idx[1] + 1], V[3 * idx[1] + 2]};
      const float amb[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
      float ab[3];
      v_cross(nl, amb, ab);
      bv = -1e30f;
#pragma unroll 1
      for (int k = 0; k < nvt; k++) {
        const float d[3] = {a[0] - V[3 * k], a[1] - V[3 * k + 1], a[2] - V[3 * k + 2]};
        const float v = fabsf(v_dot(d, ab)) + msk(k);
        if (v > bv) { bv = v; idx[2] = k; }
      }
      const float cc[3] = {V[3 * idx[2]], V[3 * idx[2] + 1], V[3 * idx[2] + 2]};
      const float amc[3] = {a[0] - cc[0], a[1] - cc[1], a[2] - cc[2]}, bmc[3] = {b[0] - cc[0], b[1] - cc[1], b[2] - cc[2]};
      float ac[3], bc[3];
      v_cross(nl, amc, ac);
      v_cross(nl, bmc, bc);
      bv = -1e30f;
#pragma unroll 1
      for (int h = 0; h < 2; h++) {  // concatenate([dist_bp, dist_ap]).argmax() % nvert
#pragma unroll 1
        for (int k = 0; k < nvt; k++) {
          const float* o = h == 0 ? b : a;
          const float* ax = h == 0 ? bc : ac;
          const float d[3] = {o[0] - V[3 * k], o[1] - V[3 * k + 1], o[2] - V[3 * k + 2]};
          const float v = fabsf(v_dot(d, ax)) + msk(k);
          if (v > bv) { bv = v; idx[3] = k; }
        }
      }
      int mine = idx[0];
      bool unique = true;
      for (int q = 1; q < 4; q++)
        if (sub == q) { mine = idx[q]; for (int j = 0; j < q; j++) if (idx[j] == idx[q]) unique = false; }
      dist = unique ? -sup(mine) : 1.f;
      const float* v = V + 3 * mine;
      for (int i = 0; i < 3; i++) pos[i] = p2[i] + m2[3 * i] * v[0] + m2[3 * i + 1] * v[1] + m2[3 * i + 2] * v[2] - 0.5f * dist * n[i];
      make_frame(n, fr);
    } else if (kind == ABR_PAIR_SPHERE_CONVEX) {
      // collision_convex.sphere_convex, in the convex geom's frame: the face with the least penetration among those the sphere
      // reaches behind, then the closest point of that polygon to the sphere centre
      const float dp[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
      float sp[3];
      for (int i = 0; i < 3; i++) sp[i] = m2[i] * dp[0] + m2[3 + i] * dp[1] + m2[6 + i] * dp[2];
      const float rad = s1[0];
      const int fa = MI(geom_faceadr)[g2], nf = MI(geom_facenum)[g2];
      const float* V = MF(vert) + 3 * MI(geom_vertadr)[g2];
      int best = 0;
      float bestv = -1e30f;
#pragma unroll 1
      for (int f = 0; f < nf; f++) {
        const float* nn = MF(face_normal) + 3 * (fa + f);
        const float* v0 = V + 3 * MI(face_vert)[MI(face_vertadr)[fa + f]];
        const float d[3] = {sp[0] - v0[0], sp[1] - v0[1], sp[2] - v0[2]};
        float sup = v_dot(d, nn) - rad;
        if (sup >= 0.f) sup = -1e12f;
        if (sup > bestv) { bestv = sup; best = f; }
      }
      const int pa = MI(face_vertadr)[fa + best], pn = MI(face_vertnum)[fa + best];
      const float* nn = MF(face_normal) + 3 * (fa + best);
      float pt[3];
      project_pt_onto_plane(sp, V + 3 * MI(face_vert)[pa], nn, pt);
      bool inside = true;
      int eidx = 0;
      float ebest = 1e30f;
#pragma unroll 1
      for (int k = 0; k < pn; k++) {
        const float* e0 = V + 3 * MI(face_vert)[pa + (k + pn - 1) % pn]; const float* e1 = V + 3 * MI(face_vert)[pa + k];
        const float ed[3] = {e1[0] - e0[0], e1[1] - e0[1], e1[2] - e0[2]}, d[3] = {pt[0] - e0[0], pt[1] - e0[1], pt[2] - e0[2]};
        float en[3];
        v_cross(ed, nn, en);
        float dd = v_dot(d, en);
        if (!(dd <= 0.f)) inside = false;
        if ((en[0] == 0.f && en[1] == 0.f && en[2] == 0.f) || dd < 0.f) dd = 1e12f;
        if (dd < ebest) { ebest = dd; eidx = k; }
      }
      if (!inside) {
        float q[3];
        closest_segment_point(V + 3 * MI(face_vert)[pa + (eidx + pn - 1) % pn], V + 3 * MI(face_vert)[pa + eidx], pt, q);
        pt[0] = q[0]; pt[1] = q[1]; pt[2] = q[2];
      }
      float n[3] = {pt[0] - sp[0], pt[1] - sp[1], pt[2] - sp[2]};
      const float d = v_normalize(n, 3);
      float pl[3], nw[3];
      for (int i = 0; i < 3; i++) pl[i] = 0.5f * (pt[i] + sp[i] + n[i] * rad);
      for (int i = 0; i < 3; i++) {
        nw[i] = m2[3 * i] * n[0] + m2[3 * i + 1] * n[1] + m2[3 * i + 2] * n[2];
        pos[i] = p2[i] + m2[3 * i] * pl[0] + m2[3 * i + 1] * pl[1] + m2[3 * i + 2] * pl[2];
      }
      dist = d - rad;
      make_frame(nw, fr);
    } else if (kind == ABR_PAIR_CAPSULE_CONVEX) {
      // collision_convex.capsule_convex, in the convex geom's frame: the axis segment clipped to the side planes of the best face
      // gives two face contacts; a face edge closer to the axis than the radius replaces the first by an edge contact
      const float rad = s1[0], half = s1[1];
      const float dp[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, axw[3] = {m1[2], m1[5], m1[8]};
      float cp[3], ax[3], c0[3], c1[3];
      for (int i = 0; i < 3; i++) { cp[i] = m2[i] * dp[0] + m2[3 + i] * dp[1] + m2[6 + i] * dp[2]; ax[i] = m2[i] * axw[0] + m2[3 + i] * axw[1] + m2[6 + i] * axw[2]; }
      for (int i = 0; i < 3; i++) { c0[i] = cp[i] - ax[i] * half; c1[i] = cp[i] + ax[i] * half; }
      const int fa = MI(geom_faceadr)[g2], nf = MI(geom_facenum)[g2];
      const float* V = MF(vert) + 3 * MI(geom_vertadr)[g2];
      int best = 0;
      float bestv = -1e30f;
      bool has_support = true;
#pragma unroll 1
      for (int f = 0; f < nf; f++) {
        const float* nn = MF(face_normal) + 3 * (fa + f);
        const float* v0 = V + 3 * MI(face_vert)[MI(face_vertadr)[fa + f]];
        const float d0[3] = {c0[0] - v0[0], c0[1] - v0[1], c0[2] - v0[2]}, d1[3] = {c1[0] - v0[0], c1[1] - v0[1], c1[2] - v0[2]};
        float sup = fminf(v_dot(d0, nn), v_dot(d1, nn)) - rad;
        if (!(sup < 0.f)) has_support = false;
        if (sup >= 0.f) sup = -1e12f;
        if (sup > bestv) { bestv = sup; best = f; }
      }
      const int pa = MI(face_vertadr)[fa + best], pn = MI(face_vertnum)[fa + best];
      const float* nn = MF(face_normal) + 3 * (fa + best);
      float P[3 * ABR_MAX_FACE_VERTS];
      for (int k = 0; k < pn; k++) for (int i = 0; i < 3; i++) P[3 * k + i] = V[3 * MI(face_vert)[pa + k] + i];
      float q[2][3];
      const bool ok = clip_edge_to_poly(c0, c1, P, pn, nn, q[0], q[1]);
      float fp[3];
      for (int i = 0; i < 3; i++) q[sub][i] -= nn[i] * rad;
      project_pt_onto_plane(q[sub], P, nn, fp);
      const float df[3] = {fp[0] - q[sub][0], fp[1] - q[sub][1], fp[2] - q[sub][2]};
      dist = -((ok && has_support) ? v_dot(df, nn) : -1.f);
      float pl[3], nl[3];
      for (int i = 0; i < 3; i++) { pl[i] = 0.5f * (q[sub][i] + fp[i]); nl[i] = -nn[i]; }
      float ebest = 1e30f, ec[3] = {0.f, 0.f, 0.f}, cc[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int k = 0; k < pn; k++) {
        float a[3], b[3];
        closest_segment_to_segment_points(P + 3 * ((k + pn - 1) % pn), P + 3 * k, c0, c1, a, b);
        const float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
        const float dd = v_dot(d, d);
        if (dd < ebest) { ebest = dd; for (int i = 0; i < 3; i++) { ec[i] = a[i]; cc[i] = b[i]; } }
      }
      float ea[3] = {cc[0] - ec[0], cc[1] - ec[1], cc[2] - ec[2]};
      const float edist = v_normalize(ea, 3);
      if (rad - edist > 0.f) {
        if (sub == 0) {
          for (int i = 0; i < 3; i++) { pl[i] = 0.5f * (ec[i] + cc[i] - ea[i] * rad); nl[i] = -ea[i]; }
          dist = -(rad - edist);
        } else {
          dist = 1.f;
        }
      }
      float nw[3];
      for (int i = 0; i < 3; i++) {
        nw[i] = m2[3 * i] * nl[0] + m2[3 * i + 1] * nl[1] + m2[3 * i + 2] * nl[2];
        pos[i] = p2[i] + m2[3 * i] * pl[0] + m2[3 * i + 1] * pl[1] + m2[3 * i + 2] * pl[2];
      }
      make_frame(nw, fr);
    } else if (kind == ABR_PAIR_CONVEX_CONVEX) {
      // collision_convex.convex_convex: separating axes (face normals of both hulls, cross products of their edges) in the world
      // frame, then the incident face clipped against the reference face (_create_contact_manifold, up to 4 contacts) or the closest
      // points of the two edges (1 contact). An axis is carried into each hull's own frame, so the vertices are never transformed.
      const float* VA = MF(vert) + 3 * MI(geom_vertadr)[g1]; const float* VB = MF(vert) + 3 * MI(geom_vertadr)[g2];
      const int nva = MI(geom_vertnum)[g1], nvb = MI(geom_vertnum)[g2];
      const int fa = MI(geom_faceadr)[g1], nfa = MI(geom_facenum)[g1], fb = MI(geom_faceadr)[g2], nfb = MI(geom_facenum)[g2];
      const int ea = MI(geom_edgeadr)[g1], nea = MI(geom_edgenum)[g1], eb = MI(geom_edgeadr)[g2], neb = MI(geom_edgenum)[g2];
      auto to_world = [&](const float* pg, const float* mg, const float* v, float* out) {
        for (int i = 0; i < 3; i++) out[i] = pg[i] + mg[3 * i] * v[0] + mg[3 * i + 1] * v[1] + mg[3 * i + 2] * v[2];
      };
      auto rot_world = [&](const float* mg, const float* v, float* out) {
        for (int i = 0; i < 3; i++) out[i] = mg[3 * i] * v[0] + mg[3 * i + 1] * v[1] + mg[3 * i + 2] * v[2];
      };
      auto axis_dist = [&](const float* axis, float& sign) {
        float la[3], lb[3];
        for (int i = 0; i < 3; i++) { la[i] = m1[i] * axis[0] + m1[3 + i] * axis[1] + m1[6 + i] * axis[2]; lb[i] = m2[i] * axis[0] + m2[3 + i] * axis[1] + m2[6 + i] * axis[2]; }
        const float oa = v_dot(axis, p1), ob = v_dot(axis, p2);
        float amax = -1e30f, amin = 1e30f, bmax = -1e30f, bmin = 1e30f;
#pragma unroll 1
        for (int k = 0; k < nva; k++) { const float v = oa + v_dot(la, VA + 3 * k); amax = fmaxf(amax, v); amin = fminf(amin, v); }
#pragma unroll 1
        for (int k = 0; k < nvb; k++) { const float v = ob + v_dot(lb, VB + 3 * k); bmax = fmaxf(bmax, v); bmin = fminf(bmin, v); }
        const float d1 = amax - bmin, d2 = bmax - amin;
        sign = (d1 > d2) ? -1.f : 1.f;
        return fminf(d1, d2);
      };
      float fbest = 1e30f, fsign = 1.f, faxis[3] = {0.f, 0.f, 1.f};
#pragma unroll 1
      for (int f = 0; f < nfa + nfb; f++) {
        float ax[3], sg;
        if (f < nfa) rot_world(m1, MF(face_normal) + 3 * (fa + f), ax); else rot_world(m2, MF(face_normal) + 3 * (fb + f - nfa), ax);
        const float d = axis_dist(ax, sg);
        if (d < fbest) { fbest = d; fsign = sg; faxis[0] = ax[0]; faxis[1] = ax[1]; faxis[2] = ax[2]; }
      }
      float ebest = 1e30f, epair = -1e30f, esign = 1.f, eaxis[3] = {0.f, 0.f, 1.f};
      int ei = 0, ej = 0;
#pragma unroll 1
      for (int j = 0; j < neb; j++) {
        float b0[3], b1[3];
        to_world(p2, m2, VB + 3 * MI(edge_vert)[2 * (eb + j)], b0);
        to_world(p2, m2, VB + 3 * MI(edge_vert)[2 * (eb + j) + 1], b1);
        float db[3] = {b0[0] - b1[0], b0[1] - b1[1], b0[2] - b1[2]};
        v_unit3(db);
#pragma unroll 1
        for (int i = 0; i < nea; i++) {
          // the edge's direction needs the rotation only; its end points are placed in the world when the pair is a candidate
          const float* la0 = VA + 3 * MI(edge_vert)[2 * (ea + i)]; const float* la1 = VA + 3 * MI(edge_vert)[2 * (ea + i) + 1];
          const float dl[3] = {la0[0] - la1[0], la0[1] - la1[1], la0[2] - la1[2]};
          float da[3];
          rot_world(m1, dl, da);
          v_unit3(da);
          float ax[3], sg;
          v_cross(da, db, ax);
          if (v_dot(ax, ax) < 1e-6f) continue;
          v_unit3(ax);
          const float d = axis_dist(ax, sg);
          if (!(d < ebest + 1e-6f)) continue;
          // parallel edges tie on their common axis: the supporting pair (its own separation along the axis = the hulls') is taken
          float a0[3], a1[3];
          to_world(p1, m1, la0, a0);
          to_world(p1, m1, la1, a1);
          const float ma[3] = {a0[0] + a1[0], a0[1] + a1[1], a0[2] + a1[2]}, mb[3] = {b0[0] + b1[0], b0[1] + b1[1], b0[2] + b1[2]};
          const float dpair = sg * 0.5f * (v_dot(ax, ma) - v_dot(ax, mb));
          if (d < ebest - 1e-6f || dpair > epair) {
            ebest = fminf(ebest, d); epair = dpair; esign = sg; ei = i; ej = j; eaxis[0] = ax[0]; eaxis[1] = ax[1]; eaxis[2] = ax[2];
          }
        }
      }
      const bool edge_contact = ebest < fbest - 1e-5f;
      float nab[3];
      for (int i = 0; i < 3; i++) nab[i] = edge_contact ? esign * eaxis[i] : fsign * faxis[i];
      dist = 1.f;
      for (int i = 0; i < 3; i++) pos[i] = 0.5f * (p1[i] + p2[i]);
      if (edge_contact) {
        if (sub == 0) {
          float a0[3], a1[3], b0[3], b1[3], ca[3], cb[3];
          to_world(p1, m1, VA + 3 * MI(edge_vert)[2 * (ea + ei)], a0); to_world(p1, m1, VA + 3 * MI(edge_vert)[2 * (ea + ei) + 1], a1);
          to_world(p2, m2, VB + 3 * MI(edge_vert)[2 * (eb + ej)], b0); to_world(p2, m2, VB + 3 * MI(edge_vert)[2 * (eb + ej) + 1], b1);
          closest_segment_to_segment_points(a0, a1, b0, b1, ca, cb);
          const float d[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]};
          dist = v_dot(d, nab);
          for (int i = 0; i < 3; i++) pos[i] = 0.5f * (ca[i] + cb[i]);
        }
      } else {
        int ia = 0, ib = 0;
        float da = -1e30f, db = -1e30f;
#pragma unroll 1
        for (int f = 0; f < nfa; f++) { float nn[3]; rot_world(m1, MF(face_normal) + 3 * (fa + f), nn); const float v = v_dot(nn, nab); if (v > da) { da = v; ia = f; } }
#pragma unroll 1
        for (int f = 0; f < nfb; f++) { float nn[3]; rot_world(m2, MF(face_normal) + 3 * (fb + f), nn); const float v = -v_dot(nn, nab); if (v > db) { db = v; ib = f; } }
        const bool ref_a = da >= db;
        const int rf = ref_a ? fa + ia : fb + ib, sf = ref_a ? fb + ib : fa + ia;
        const float* RVl = ref_a ? VA : VB; const float* SVl = ref_a ? VB : VA;
        const float* rp = ref_a ? p1 : p2; const float* rm = ref_a ? m1 : m2;
        const float* spp = ref_a ? p2 : p1; const float* sm = ref_a ? m2 : m1;
        float rn[3], sn[3];
        rot_world(rm, MF(face_normal) + 3 * rf, rn);
        rot_world(sm, MF(face_normal) + 3 * sf, sn);
        const int nr = MI(face_vertnum)[rf], ns = MI(face_vertnum)[sf];
        float RP[3 * ABR_MAX_FACE_VERTS], SP[3 * ABR_MAX_FACE_VERTS], Q[3 * ABR_MAX_FACE_VERTS], inc[3 * 4 * ABR_MAX_FACE_VERTS];
        for (int k = 0; k < nr; k++) to_world(rp, rm, RVl + 3 * MI(face_vert)[MI(face_vertadr)[rf] + k], RP + 3 * k);
        for (int k = 0; k < ns; k++) to_world(spp, sm, SVl + 3 * MI(face_vert)[MI(face_vertadr)[sf] + k], SP + 3 * k);
        unsigned mk = 0u;
#pragma unroll 1
        for (int k = 0; k < ns; k++) {
          const bool ok = clip_edge_to_poly(SP + 3 * ((k + ns - 1) % ns), SP + 3 * k, RP, nr, rn, inc + 3 * (2 * k), inc + 3 * (2 * k + 1));
          if (ok) mk |= 3u << (2 * k);
        }
        {
          const float dpl = v_dot(SP, sn), denom = v_dot(rn, sn);
          for (int k = 0; k < nr; k++) {
            const float t = (dpl - v_dot(RP + 3 * k, sn)) / (denom + ((denom == 0.f) ? 1e-6f : 0.f));
            for (int i = 0; i < 3; i++) Q[3 * k + i] = RP[3 * k + i] + t * rn[i];
          }
#pragma unroll 1
          for (int k = 0; k < nr; k++) {
            const bool ok = clip_edge_to_poly(Q + 3 * ((k + nr - 1) % nr), Q + 3 * k, SP, ns, sn, inc + 3 * (2 * (ns + k)), inc + 3 * (2 * (ns + k) + 1));
            if (ok) mk |= 3u << (2 * (ns + k));
          }
        }
        const int npt = 2 * (ns + nr);
        // the points on the reference plane replace the incident ones in `Q`-sized scratch: keep both (inc: incident, ref: projected)
        float ref[3 * 4 * ABR_MAX_FACE_VERTS];
        unsigned mb = 0u;
#pragma unroll 1
        for (int k = 0; k < npt; k++) {
          project_pt_onto_plane(inc + 3 * k, RP, rn, ref + 3 * k);
          const float d[3] = {inc[3 * k] - RP[0], inc[3 * k + 1] - RP[1], inc[3 * k + 2] - RP[2]};
          if (((mk >> k) & 1u) && (-v_dot(d, rn) > 1e-6f)) mb |= 1u << k;
        }
        int idx[4];
        manifold_points(ref, mb, npt, rn, idx);
        bool unique = true;
        for (int j = 0; j < sub; j++) if (idx[j] == idx[sub]) unique = false;
        const int me = idx[sub];
        const float d[3] = {inc[3 * me] - ref[3 * me], inc[3 * me + 1] - ref[3 * me + 1], inc[3 * me + 2] - ref[3 * me + 2]};
        dist = (((mb >> me) & 1u) && unique) ? v_dot(d, rn) : 1.f;
        for (int i = 0; i < 3; i++) pos[i] = 0.5f * (inc[3 * me + i] + ref[3 * me + i]);
      }
      make_frame(nab, fr);
    } else if (kind == ABR_PAIR_PLANE_SPHERE || kind == ABR_PAIR_PLANE_CAPSULE) {
      float n[3] = {m1[2], m1[5], m1[8]};
      float sp[3] = {p2[0], p2[1], p2[2]};
      float rad = s2[0];
      if (kind == ABR_PAIR_PLANE_CAPSULE) {
        float axis[3] = {m2[2], m2[5], m2[8]};
        float na = v_dot(n, axis);
        float b[3] = {axis[0] - n[0] * na, axis[1] - n[1] * na, axis[2] - n[2] * na};
        float bn = v_normalize(b, 3);
        if (bn < 0.5f) {
          b[0] = 0.f; b[1] = 0.f; b[2] = 0.f;
          if (-0.5f < n[1] && n[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
        }
        float cr[3];
        v_cross(n, b, cr);
        for (int i = 0; i < 3; i++) { fr[i] = n[i]; fr[3 + i] = b[i]; fr[6 + i] = cr[i]; }
        const float h = (sub == 0) ? s2[1] : -s2[1];
        sp[0] += axis[0] * h; sp[1] += axis[1] * h; sp[2] += axis[2] * h;
      } else {
        make_frame(n, fr);
      }
      float d[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
      dist = v_dot(d, n) - rad;
      for (int i = 0; i < 3; i++) pos[i] = sp[i] - n[i] * (rad + 0.5f * dist);
    } else {
      float a1[3] = {p1[0], p1[1], p1[2]}, a2[3] = {p2[0], p2[1], p2[2]};
      if (kind == ABR_PAIR_SPHERE_CAPSULE) {
        float axis[3] = {m2[2], m2[5], m2[8]};
        float d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
        float t = fminf(fmaxf(v_dot(d, axis), -s2[1]), s2[1]);
        for (int i = 0; i < 3; i++) a2[i] = p2[i] + axis[i] * t;
      } else if (kind == ABR_PAIR_CAPSULE_CAPSULE) {
        float ax1[3] = {m1[2], m1[5], m1[8]}, ax2[3] = {m2[2], m2[5], m2[8]};
        float d[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
        float dab = v_dot(ax1, ax2), d1 = v_dot(d, ax1), d2 = v_dot(d, ax2);
        float den = 1.f - dab * dab;
        float t1 = (den == 0.f) ? 0.f : (dab * d2 - d1) / den;
        t1 = fminf(fmaxf(t1, -s1[1]), s1[1]);
        float t2 = fminf(fmaxf(t1 * dab + d2, -s2[1]), s2[1]);
        t1 = fminf(fmaxf(t2 * dab - d1, -s1[1]), s1[1]);
        for (int i = 0; i < 3; i++) { a1[i] = p1[i] + ax1[i] * t1; a2[i] = p2[i] + ax2[i] * t2; }
      }
      float n[3] = {a2[0] - a1[0], a2[1] - a1[1], a2[2] - a1[2]};
      dist = v_normalize(n, 3);
      if (dist == 0.f) { n[0] = 1.f; n[1] = 0.f; n[2] = 0.f; }
      dist -= (s1[0] + s2[0]);
      for (int i = 0; i < 3; i++) pos[i] = a1[i] + n[i] * (s1[0] + dist * 0.5f);
      make_frame(n, fr);
    }
    cdist[ci] = dist;
    for (int i = 0; i < 3; i++) cpos[3 * ci + i] = pos[i];
    for (int i = 0; i < 9; i++) cframe[9 * ci + i] = fr[i];
  }
  __syncwarp();
  // Jacobian basis: item = (contact, dof on the contact's chain); other columns stay 0 from init
  float* B = WF(B);
  const float* cdof = WF(cdof); const float* rootcom = WF(rootcom);
  const int nv = L.nv;
  const int ncd = MI(cd_adr)[L.ncon];
#pragma unroll 1
  for (int it = c.lane; it < ncd; it += G) {
    const int cd = MI(cd_dof)[it];
    const int ci = cd >> 16, d = cd & 0xffff;
    const int b = MI(dof_body)[d];
    const float* rc = rootcom + 3 * MI(body_rootslot)[b];
    float off[3] = {cpos[3 * ci] - rc[0], cpos[3 * ci + 1] - rc[1], cpos[3 * ci + 2] - rc[2]};
    float cr[3];
    v_cross(cdof + 6 * d, off, cr);
    float jp[3] = {cdof[6 * d + 3] + cr[0], cdof[6 * d + 4] + cr[1], cdof[6 * d + 5] + cr[2]};
    const float sg = (MI(con_dofmask)[ci * nv + d] == 1) ? 1.f : -1.f;
    const float* fr = cframe + 9 * ci;
    B[(3 * ci) * nv + d] = sg * v_dot(fr, jp);
    B[(3 * ci + 1) * nv + d] = sg * v_dot(fr + 3, jp);
    B[(3 * ci + 2) * nv + d] = sg * v_dot(fr + 6, jp);
  }
  __syncwarp();
}

// smooth.crb + support.make_m: composite inertias (leaves -> root, parents pull their children)
// and the packed lower-triangular joint-space inertia M.
template <int G> __device__ void stage_crb(const Ctx& c) {
  const Layout& L = c.L;
  float* crb = WF(crb); const float* cinert = WF(cinert);
  for (int i = 10 + c.lane; i < 10 * L.nbody; i += G) crb[i] = cinert[i];
  __syncwarp();
  for (int lev = L.depth - 1; lev >= 1; lev--) {
    const int beg = MI(level_adr)[lev], end = MI(level_adr)[lev + 1];
    for (int k = beg + c.lane; k < end; k += G) {
      const int p = MI(level_body)[k];
      const int ca = MI(body_childadr)[p], cn = MI(body_childnum)[p];
      if (cn) {
        float acc[10];
        for (int i = 0; i < 10; i++) acc[i] = crb[10 * p + i];
        for (int q = 0; q < cn; q++) {
          const int ch = MI(child)[ca + q];
          for (int i = 0; i < 10; i++) acc[i] += crb[10 * ch + i];
        }
        for (int i = 0; i < 10; i++) crb[10 * p + i] = acc[i];
      }
    }
    __syncwarp();
  }
  float* buf = WF(buf); const float* cdof = WF(cdof);
  for (int d = c.lane; d < L.nv; d += G) inert_mul(crb + 10 * MI(dof_body)[d], cdof + 6 * d, buf + 6 * d);
  __syncwarp();
  float* M = WF(M);
  for (int k = c.lane; k < L.nmpair; k += G) {
    const int ij = MI(mpair)[k];
    const int i = ij >> 16, j = ij & 0xffff;
    const float* a = cdof + 6 * j; const float* b = buf + 6 * i;
    float s = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
    if (i == j) s += c.ars * MF(dof_armature)[i];
    M[L.sparse ? k : tri(i) + j] = s;
  }
  __syncwarp();
}

// dense Cholesky of a packed lower-triangular matrix, in place (jax cho_factor on the dense
// path). Column k: every lane owning a row i >= k forms s_i = A[i,k] - sum_p A[i,p] A[k,p] and,
// redundantly, the pivot s_k. Pointer-only signature so it can stay a real function (called for
// M, H and the implicit-damping matrix) without dragging the Layout along.
template <int G> __device__ __noinline__ void chol_factor(float* A, int n, int lane) {
  for (int k = 0; k < n; k++) {
    const float* rk = A + tri(k);
    float sk = rk[k];
    for (int p = 0; p < k; p++) sk -= rk[p] * rk[p];
    const float d = sqrtf(fmaxf(sk, kMinVal));  // pivot clamp as in MuJoCo's mj_cholFactor
    const float inv = 1.f / d;
    float vals[(64 + G - 1) / G];
    int cnt = 0;
    for (int i = k + lane; i < n; i += G, cnt++) {
      const float* ri = A + tri(i);
      float s = ri[k];
      for (int p = 0; p < k; p++) s -= ri[p] * rk[p];
      vals[cnt] = (i == k) ? d : s * inv;
    }
    __syncwarp();  // all reads of column k's inputs done before anyone overwrites
    cnt = 0;
    for (int i = k + lane; i < n; i += G, cnt++) A[tri(i) + k] = vals[cnt];
    __syncwarp();
  }
}

// x = (L L^T)^-1 b, column-oriented substitutions. y is an n-float scratch; x may alias b.
template <int G> __device__ __noinline__ void chol_solve(const float* A, const float* b, float* x, float* y, int n, int lane) {
  for (int i = lane; i < n; i += G) y[i] = b[i];
  __syncwarp();
  for (int k = 0; k < n; k++) {
    const float yk = y[k] / A[tri(k) + k];
    __syncwarp();
    if (lane == 0) y[k] = yk;
    for (int i = k + 1 + lane; i < n; i += G) y[i] -= A[tri(i) + k] * yk;
    __syncwarp();
  }
  for (int k = n - 1; k >= 0; k--) {
    const float xk = y[k] / A[tri(k) + k];
    __syncwarp();
    if (lane == 0) y[k] = xk;
    const float* rk = A + tri(k);
    for (int i = lane; i < k; i += G) y[i] -= rk[i] * xk;
    __syncwarp();
  }
  for (int i = lane; i < n; i += G) x[i] = y[i];
  __syncwarp();
}

// out = M v (packed symmetric). No trailing sync: caller syncs.
template <int G> __device__ __noinline__ void mul_m(const float* M, const float* v, float* out, int n, int lane) {
  for (int i = lane; i < n; i += G) {
    const float* ri = M + tri(i);
    float s = 0.f;
    for (int j = 0; j <= i; j++) s += ri[j] * v[j];
    for (int j = i + 1; j < n; j++) s += M[tri(j) + i] * v[j];
    out[i] = s;
  }
}

// ------------------------------------------------------------------------------ tree-sparse L'DL
// When every constraint couples dofs of one ancestor chain only, H = M + J'DJ has M's
// branch-induced sparsity and eliminating dofs leaves-first (MuJoCo's L'DL order) creates no
// fill-in. Entries (i,j), j ancestor-or-self of i, are stored in row-chain order (diagonal first);
// dofs of equal height in the dof tree form one elimination stage, and every update is a gather
// over a host-built task list, so a stage needs two __syncwarp() and no atomics.
// After sp_factor: A holds D on the diagonal and D*L off it (L[k,i] = A[k,i] * invD[k]).
template <int G> __device__ __forceinline__ void sp_factor(const Ctx& c, float* A) {
  const Layout& L = c.L;
  float* invD = WF(invD);
  const int* sp_adr = MI(sp_adr);
  for (int s = 0; s < L.nstage; s++) {
    const int d0 = MI(st_adr)[s], d1 = MI(st_adr)[s + 1];
#pragma unroll 1
    for (int q = d0 + c.lane; q < d1; q += G) { const int k = MI(st_dof)[q]; invD[k] = 1.f / fmaxf(A[sp_adr[k]], kMinVal); }
    __syncwarp();
    const int t0 = MI(fu_tadr)[s], t1 = MI(fu_tadr)[s + 1];
#pragma unroll 1
    for (int t = t0 + c.lane; t < t1; t += G) {
      float acc = 0.f;
      const int q1 = MI(fu_cadr)[t + 1];
#pragma unroll 1
      for (int q = MI(fu_cadr)[t]; q < q1; q++) {
        const int src = MI(fu_src)[q];
        acc += A[src >> 16] * A[src & 0xffff] * invD[MI(fu_k)[q]];
      }
      A[MI(fu_tgt)[t]] -= acc;
    }
    __syncwarp();
  }
}

// x = (L' D L)^-1 b with the factor above; x may alias b; WF(y) is the scratch.
template <int G> __device__ __forceinline__ void sp_solve(const Ctx& c, const float* A, const float* b, float* x) {
  const Layout& L = c.L;
  float* z = WF(y); const float* invD = WF(invD);
  const int* sp_adr = MI(sp_adr);
#pragma unroll 1
  for (int i = c.lane; i < L.nv; i += G) z[i] = b[i];
  __syncwarp();
  for (int s = 0; s < L.nstage; s++) {  // leaves -> root:  z = D^-1 L^-T b
    const int d0 = MI(st_adr)[s], d1 = MI(st_adr)[s + 1];
#pragma unroll 1
    for (int q = d0 + c.lane; q < d1; q += G) { const int k = MI(st_dof)[q]; z[k] *= invD[k]; }
    __syncwarp();
    const int t0 = MI(sb_tadr)[s], t1 = MI(sb_tadr)[s + 1];
#pragma unroll 1
    for (int t = t0 + c.lane; t < t1; t += G) {
      float acc = 0.f;
      const int q1 = MI(sb_cadr)[t + 1];
#pragma unroll 1
      for (int q = MI(sb_cadr)[t]; q < q1; q++) { const int src = MI(sb_src)[q]; acc += A[src >> 16] * z[src & 0xffff]; }
      z[MI(sb_tgt)[t]] -= acc;
    }
    __syncwarp();
  }
  for (int s = L.nstage - 1; s >= 0; s--) {  // root -> leaves:  x = L^-1 z
    const int d0 = MI(st_adr)[s], d1 = MI(st_adr)[s + 1];
#pragma unroll 1
    for (int q = d0 + c.lane; q < d1; q += G) {
      const int k = MI(st_dof)[q];
      float acc = 0.f;
      const int e1 = sp_adr[k + 1];
#pragma unroll 1
      for (int e = sp_adr[k] + 1; e < e1; e++) acc += A[e] * z[MI(mpair)[e] & 0xffff];
      z[k] -= invD[k] * acc;
    }
    __syncwarp();
  }
#pragma unroll 1
  for (int i = c.lane; i < L.nv; i += G) x[i] = z[i];
  __syncwarp();
}

// out = M v for the sparse storage. No trailing sync: caller syncs.
template <int G> __device__ __forceinline__ void sp_mul_m(const Ctx& c, const float* M, const float* v, float* out) {
#pragma unroll 1
  for (int i = c.lane; i < c.L.nv; i += G) {
    float acc = 0.f;
    const int q1 = MI(mm_adr)[i + 1];
#pragma unroll 1
    for (int q = MI(mm_adr)[i]; q < q1; q++) { const int src = MI(mm_src)[q]; acc += M[src >> 16] * v[src & 0xffff]; }
    out[i] = acc;
  }
}

// dispatchers: tree-sparse path or dense packed-Cholesky fallback
template <int G> __device__ __forceinline__ void la_factor(const Ctx& c, float* A) {
  if (c.L.sparse) sp_factor<G>(c, A); else chol_factor<G>(A, c.L.nv, c.lane);
}
template <int G> __device__ __forceinline__ void la_solve(const Ctx& c, const float* A, const float* b, float* x) {
  if (c.L.sparse) sp_solve<G>(c, A, b, x); else chol_solve<G>(A, b, x, WF(y), c.L.nv, c.lane);
}
template <int G> __device__ __forceinline__ void la_mul_m(const Ctx& c, const float* v, float* out) {
  if (c.L.sparse) sp_mul_m<G>(c, WF(M), v, out); else mul_m<G>(WF(M), v, out, c.L.nv, c.lane);
}

// out[r] = (J v)[r] for the structured Jacobian (equality coefficients, limit signs, contact
// basis B). bv is a 3*ncon scratch. Ends synced.
template <int G> __device__ void mul_j(const Ctx& c, const float* v, float* out) {
  const Layout& L = c.L;
  const int nv = L.nv;
  float* bv = WF(bv); const float* B = WF(B);
#pragma unroll 1
  for (int it = c.lane; it < 3 * L.ncon; it += G) {
    const float* row = B + it * nv;
    const int ci = it / 3;
    float s = 0.f;
    const int q1 = MI(cd_adr)[ci + 1];
#pragma unroll 1
    for (int q = MI(cd_adr)[ci]; q < q1; q++) { const int d = MI(cd_dof)[q] & 0xffff; s += row[d] * v[d]; }
    bv[it] = s;
  }
  __syncwarp();
  const float* eqc = WF(eqc); const float* lims = WF(lims);
#pragma unroll 1
  for (int r = c.lane; r < L.nefc; r += G) {
    const int info = MI(row_info)[r];
    const int kind = info & 3, idx = (info >> 2) & 0x3ffff, sub = info >> 20;
    float s;
    if (kind == 0) {
      const int j1 = MI(eq_j1)[idx], j2 = MI(eq_j2)[idx];
      s = v[MI(jnt_dofadr)[j1]];
      if (j2 >= 0) s += eqc[idx] * v[MI(jnt_dofadr)[j2]];
    } else if (kind == 1) {
      s = lims[idx] * v[MI(jnt_dofadr)[MI(lim_jnt)[idx]]];
    } else {
      s = bv[3 * idx];
      if (MI(con_condim)[idx] == 3) {
        const float* prm = MF(con_prm) + kConPrm * MI(con_pair)[idx];
        const float mu = c.frs * prm[11 + (sub >> 1)];
        s += ((sub & 1) ? -mu : mu) * bv[3 * idx + 1 + (sub >> 1)];
      }
    }
    out[r] = s;
  }
  __syncwarp();
}

// out = J^T f. Fc is a 3*ncon scratch. Ends synced.
template <int G> __device__ void mul_jt(const Ctx& c, const float* f, float* out) {
  const Layout& L = c.L;
  const int nv = L.nv;
  float* Fc = WF(Fc);
  for (int ci = c.lane; ci < L.ncon; ci += G) {
    const int r0 = MI(con_row)[ci];
    if (MI(con_condim)[ci] == 3) {
      const float* prm = MF(con_prm) + kConPrm * MI(con_pair)[ci];
      const float f0 = f[r0], f1 = f[r0 + 1], f2 = f[r0 + 2], f3 = f[r0 + 3];
      Fc[3 * ci] = f0 + f1 + f2 + f3;
      Fc[3 * ci + 1] = c.frs * prm[11] * (f0 - f1);
      Fc[3 * ci + 2] = c.frs * prm[12] * (f2 - f3);
    } else {
      Fc[3 * ci] = f[r0]; Fc[3 * ci + 1] = 0.f; Fc[3 * ci + 2] = 0.f;
    }
  }
  __syncwarp();
  const float* eqc = WF(eqc); const float* lims = WF(lims); const float* B = WF(B);
  for (int d = c.lane; d < nv; d += G) {
    float s = 0.f;
    for (int e = 0; e < L.ne; e++) {
      const int j1 = MI(eq_j1)[e], j2 = MI(eq_j2)[e];
      if (MI(jnt_dofadr)[j1] == d) s += f[e];
      if (j2 >= 0 && MI(jnt_dofadr)[j2] == d) s += eqc[e] * f[e];
    }
    const int lr = MI(dof_limrow)[d];
    if (lr >= 0) s += lims[lr] * f[L.ne + lr];
    for (int it = 0; it < 3 * L.ncon; it++) s += B[it * nv + d] * Fc[it];
    out[d] = s;
  }
  __syncwarp();
}

// constraint.make_constraint: equality coefficients, limit signs, efc_D, efc_aref per row.
template <int G> __device__ void stage_rows(const Ctx& c) {
  const Layout& L = c.L;
  if (L.nefc == 0) return;
  const float* qpos = WF(qpos); const float* qvel = WF(qvel);
  float* eqc = WF(eqc); float* lims = WF(lims); float* D = WF(D); float* aref = WF(aref);
  float* bv = WF(bv); const float* B = WF(B);
  const int nv = L.nv;
#pragma unroll 1
  for (int it = c.lane; it < 3 * L.ncon; it += G) {
    const float* row = B + it * nv;
    const int ci = it / 3;
    float s = 0.f;
    const int q1 = MI(cd_adr)[ci + 1];
#pragma unroll 1
    for (int q = MI(cd_adr)[ci]; q < q1; q++) { const int d = MI(cd_dof)[q] & 0xffff; s += row[d] * qvel[d]; }
    bv[it] = s;
  }
  __syncwarp();
  for (int r = c.lane; r < L.nefc; r += G) {
    const int info = MI(row_info)[r];
    const int kind = info & 3, idx = (info >> 2) & 0x3ffff, sub = info >> 20;
    const float* prm;
    float pos, jvel, invw;
    bool active = true;
    if (kind == 0) {
      const int j1 = MI(eq_j1)[idx], j2 = MI(eq_j2)[idx];
      const int a1 = MI(jnt_qposadr)[j1];
      const float* data = MF(eq_data) + 5 * idx;
      const float pos1 = qpos[a1] - MF(qpos0)[a1];
      prm = MF(eq_prm) + kRowPrm * idx;
      jvel = qvel[MI(jnt_dofadr)[j1]];
      if (j2 >= 0) {
        const int a2 = MI(jnt_qposadr)[j2];
        const float dif = qpos[a2] - MF(qpos0)[a2];
        const float d2 = dif * dif, d3 = d2 * dif, d4 = d3 * dif;
        const float poly = data[0] + data[1] * dif + data[2] * d2 + data[3] * d3 + data[4] * d4;
        const float deriv = data[1] + data[2] * dif * 2.f + data[3] * d2 * 3.f + data[4] * d3 * 4.f;
        pos = pos1 - poly;
        eqc[idx] = -deriv;
        jvel += -deriv * qvel[MI(jnt_dofadr)[j2]];
      } else {
        pos = pos1 - data[0];
        eqc[idx] = 0.f;
      }
      invw = prm[7];
    } else if (kind == 1) {
      const int j = MI(lim_jnt)[idx];
      const int a = MI(jnt_qposadr)[j];
      const float dmin = qpos[a] - MF(jnt_range)[2 * j], dmax = MF(jnt_range)[2 * j + 1] - qpos[a];
      pos = fminf(dmin, dmax) - MF(jnt_margin)[j];
      active = pos < 0.f;
      const float sg = (dmin < dmax) ? 1.f : -1.f;
      lims[idx] = active ? sg : 0.f;
      jvel = sg * qvel[MI(jnt_dofadr)[j]];
      prm = MF(lim_prm) + kRowPrm * idx;
      invw = prm[7];
    } else {
      prm = MF(con_prm) + kConPrm * MI(con_pair)[idx];
      pos = WF(cdist)[idx] - prm[13];
      active = pos < 0.f;
      jvel = bv[3 * idx];
      invw = prm[7];
      if (MI(con_condim)[idx] == 3) {
        const int k = sub >> 1;
        const float mu0 = prm[11 + k], mu = c.frs * mu0;
        jvel += ((sub & 1) ? -mu : mu) * bv[3 * idx + 1 + k];
        invw = (k == 0) ? prm[7] : prm[10];
        // the pyramid's invweight is t (1 + mu^2) 2 mu^2 / impratio: rescale it with the friction
        if (c.frs != 1.f) invw *= c.frs * c.frs * (1.f + mu * mu) / (1.f + mu0 * mu0);
      }
    }
    float Dr = 0.f, ar = 0.f;
    if (active) {
      const float dmin = prm[2], dmax = prm[3], mid = prm[5], power = prm[6];
      const float x = fabsf(pos) * prm[4];
      float ia, ib;
      if (power == 2.f) { ia = x * x; const float t = 1.f - x; ib = t * t; }
      else if (power == 1.f) { ia = x; ib = 1.f - x; }
      else { ia = powf(x, power); ib = powf(1.f - x, power); }
      const float y = (x < mid) ? prm[8] * ia : 1.f - prm[9] * ib;
      float imp = dmin + y * (dmax - dmin);
      imp = fminf(fmaxf(imp, dmin), dmax);
      if (x > 1.f) imp = dmax;
      const float R = fmaxf(invw * (1.f - imp) / imp, kMinVal);
      Dr = 1.f / R;
      ar = -prm[1] * jvel - prm[0] * imp * pos;
    }
    D[r] = Dr;
    aref[r] = ar;
  }
  __syncwarp();
}

// fwd_velocity + fwd_actuation + fwd_acceleration's qfrc_smooth:
// com_vel (cvel, cdof_dot) and the RNE acceleration pass fused root -> leaves, body forces,
// leaves -> root accumulation, qfrc_bias, passive and actuator forces.
template <int G> __device__ void stage_velocity(const Ctx& c) {
  const Layout& L = c.L;
  const float* cdof = WF(cdof); const float* qvel = WF(qvel); const float* qpos = WF(qpos);
  float* cvel = WF(cvel); float* cacc = WF(cacc); float* cdd = WF(cdofdot);
  const bool grav = !(L.disableflags & ABR_DSBL_GRAVITY);
  if (c.lane == 0) {
    for (int i = 0; i < 6; i++) cvel[i] = 0.f;
    cacc[0] = 0.f; cacc[1] = 0.f; cacc[2] = 0.f;
    cacc[3] = grav ? -L.gravity[0] : 0.f; cacc[4] = grav ? -L.gravity[1] : 0.f; cacc[5] = grav ? -L.gravity[2] : 0.f;
  }
  __syncwarp();
  for (int lev = 1; lev <= L.depth; lev++) {
    const int beg = MI(level_adr)[lev], end = MI(level_adr)[lev + 1];
    for (int k = beg + c.lane; k < end; k += G) {
      const int b = MI(level_body)[k];
      const int p = MI(body_parent)[b];
      float cv[6], ca[6];
      for (int i = 0; i < 6; i++) { cv[i] = cvel[6 * p + i]; ca[i] = cacc[6 * p + i]; }
      const int jn = MI(body_jntnum)[b], ja = MI(body_jntadr)[b];
      for (int jj = 0; jj < jn; jj++) {
        const int j = ja + jj;
        const int d = MI(jnt_dofadr)[j];
        if (MI(jnt_type)[j] == ABR_JNT_FREE) {
          for (int q = 0; q < 3; q++) {
            for (int i = 0; i < 6; i++) { cv[i] += cdof[6 * (d + q) + i] * qvel[d + q]; cdd[6 * (d + q) + i] = 0.f; }
          }
          for (int q = 3; q < 6; q++) motion_cross(cv, cdof + 6 * (d + q), cdd + 6 * (d + q));
          for (int q = 3; q < 6; q++)
            for (int i = 0; i < 6; i++) cv[i] += cdof[6 * (d + q) + i] * qvel[d + q];
          for (int q = 3; q < 6; q++)
            for (int i = 0; i < 6; i++) ca[i] += cdd[6 * (d + q) + i] * qvel[d + q];
        } else {
          float dd[6];
          motion_cross(cv, cdof + 6 * d, dd);
          const float qd = qvel[d];
          for (int i = 0; i < 6; i++) { cdd[6 * d + i] = dd[i]; cv[i] += cdof[6 * d + i] * qd; ca[i] += dd[i] * qd; }
        }
      }
      for (int i = 0; i < 6; i++) { cvel[6 * b + i] = cv[i]; cacc[6 * b + i] = ca[i]; }
    }
    __syncwarp();
  }
  // body forces, in place of cacc
  const float* cinert = WF(cinert);
  for (int b = 1 + c.lane; b < L.nbody; b += G) {
    float f1[6], f2[6], f3[6];
    inert_mul(cinert + 10 * b, cacc + 6 * b, f1);
    inert_mul(cinert + 10 * b, cvel + 6 * b, f2);
    motion_cross_force(cvel + 6 * b, f2, f3);
    for (int i = 0; i < 6; i++) cacc[6 * b + i] = f1[i] + f3[i];
  }
  __syncwarp();
  for (int lev = L.depth - 1; lev >= 1; lev--) {
    const int beg = MI(level_adr)[lev], end = MI(level_adr)[lev + 1];
    for (int k = beg + c.lane; k < end; k += G) {
      const int p = MI(level_body)[k];
      const int ca = MI(body_childadr)[p], cn = MI(body_childnum)[p];
      if (cn) {
        float acc[6];
        for (int i = 0; i < 6; i++) acc[i] = cacc[6 * p + i];
        for (int q = 0; q < cn; q++) {
          const int ch = MI(child)[ca + q];
          for (int i = 0; i < 6; i++) acc[i] += cacc[6 * ch + i];
        }
        for (int i = 0; i < 6; i++) cacc[6 * p + i] = acc[i];
      }
    }
    __syncwarp();
  }
  // actuator forces (transmission + fwd_actuation)
  float* actf = WF(actf); const float* ctrl = WF(ctrl);
  const bool act_on = !(L.disableflags & ABR_DSBL_ACTUATION);
  for (int u = c.lane; u < L.nu; u += G) {
    const float* prm = MF(act_prm) + kActPrm * u;
    const int flags = MI(act_flags)[u];
    const int j = MI(act_jnt)[u];
    float f = 0.f;
    if (act_on) {
      float ct = ctrl[u];
      if ((flags & 1) && !(L.disableflags & ABR_DSBL_CLAMPCTRL)) ct = fminf(fmaxf(ct, prm[0]), prm[1]);
      const float len = qpos[MI(jnt_qposadr)[j]] * prm[10];
      const float vel = qvel[MI(jnt_dofadr)[j]] * prm[10];
      float gain = prm[4];
      if (flags & 4) gain += prm[5] * len + prm[6] * vel;
      float bias = 0.f;
      if (flags & 8) bias = prm[7] + prm[8] * len + prm[9] * vel;
      f = (gain * ct + bias) * c.acs;
      if (flags & 2) f = fminf(fmaxf(f, prm[2]), prm[3]);
      f *= prm[10];
    }
    actf[u] = f;
  }
  __syncwarp();
  float* fs = WF(fs);
  const bool passive_on = !(L.disableflags & ABR_DSBL_PASSIVE);
  for (int d = c.lane; d < L.nv; d += G) {
    const float* cd = cdof + 6 * d;
    const float* f = cacc + 6 * MI(dof_body)[d];
    const float bias = cd[0] * f[0] + cd[1] * f[1] + cd[2] * f[2] + cd[3] * f[3] + cd[4] * f[4] + cd[5] * f[5];
    float s = 0.f;
    if (passive_on) {
      const int j = MI(dof_jnt)[d];
      if (MI(jnt_type)[j] != ABR_JNT_FREE) {
        const int a = MI(jnt_qposadr)[j];
        s = -MF(jnt_stiffness)[j] * (qpos[a] - MF(qpos_spring)[a]);
      }
      s -= c.dps * MF(dof_damping)[d] * qvel[d];
    }
    s -= bias;
    const int aa = MI(dof_actadr)[d], an = MI(dof_actnum)[d];
    for (int q = 0; q < an; q++) s += actf[MI(dof_act)[aa + q]];
    fs[d] = s;
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------ solver.py
struct LSPoint { float alpha, cost, d0, d1; };

// cost = 0.5 sum_active D Jaref^2 + 0.5 (Ma - fs).(a - as); group-uniform result
template <int G> __device__ float solver_cost(const Ctx& c, const float* a, const float* Ma, const float* Jaref, float* gauss_out) {
  const Layout& L = c.L;
  const float* D = WF(D); const float* fs = WF(fs); const float* as = WF(as);
  float s = 0.f, g = 0.f;
  for (int r = c.lane; r < L.nefc; r += G) {
    const float x = Jaref[r];
    if (r < L.ne || x < 0.f) s += D[r] * x * x;
  }
  for (int d = c.lane; d < L.nv; d += G) g += (Ma[d] - fs[d]) * (a[d] - as[d]);
  s = gsum<G>(s);
  g = gsum<G>(g);
  *gauss_out = 0.5f * g;
  return 0.5f * s + 0.5f * g;
}

// efc_force + qfrc_constraint (update_constraint's outputs)
template <int G> __device__ void solver_forces(const Ctx& c, const float* Jaref) {
  const Layout& L = c.L;
  const float* D = WF(D); float* force = WF(force);
  for (int r = c.lane; r < L.nefc; r += G) {
    const float x = Jaref[r];
    force[r] = (r < L.ne || x < 0.f) ? -D[r] * x : 0.f;
  }
  __syncwarp();
  mul_jt<G>(c, force, WF(fc));
}

// H = M + J^T diag(D active) J, packed lower, into WF(H)
template <int G> __device__ void solver_hessian(const Ctx& c, const float* Jaref) {
  const Layout& L = c.L;
  const int nv = L.nv;
  const float* D = WF(D); const float* B = WF(B); float* WB = WF(WB);
  // per contact: 3x3 weight in the (n,t1,t2) basis, then WB = W B on the contact's dof chain
  const int ncd = L.sparse ? MI(cd_adr)[L.ncon] : L.ncon * nv;
#pragma unroll 1
  for (int it = c.lane; it < ncd; it += G) {
    int ci, j;
    if (L.sparse) { const int cd = MI(cd_dof)[it]; ci = cd >> 16; j = cd & 0xffff; }
    else { ci = it / nv; j = it - ci * nv; }  // dense fallback reads every column of WB
    const int r0 = MI(con_row)[ci];
    const float b0 = B[(3 * ci) * nv + j], b1 = B[(3 * ci + 1) * nv + j], b2 = B[(3 * ci + 2) * nv + j];
    float o0, o1, o2;
    if (MI(con_condim)[ci] == 3) {
      const float* prm = MF(con_prm) + kConPrm * MI(con_pair)[ci];
      const float mu1 = c.frs * prm[11], mu2 = c.frs * prm[12];
      const float w0 = (Jaref[r0] < 0.f) ? D[r0] : 0.f, w1 = (Jaref[r0 + 1] < 0.f) ? D[r0 + 1] : 0.f;
      const float w2 = (Jaref[r0 + 2] < 0.f) ? D[r0 + 2] : 0.f, w3 = (Jaref[r0 + 3] < 0.f) ? D[r0 + 3] : 0.f;
      const float W00 = w0 + w1 + w2 + w3, W01 = mu1 * (w0 - w1), W02 = mu2 * (w2 - w3);
      const float W11 = mu1 * mu1 * (w0 + w1), W22 = mu2 * mu2 * (w2 + w3);
      o0 = W00 * b0 + W01 * b1 + W02 * b2;
      o1 = W01 * b0 + W11 * b1;
      o2 = W02 * b0 + W22 * b2;
    } else {
      const float w0 = (Jaref[r0] < 0.f) ? D[r0] : 0.f;
      o0 = w0 * b0; o1 = 0.f; o2 = 0.f;
    }
    WB[(3 * ci) * nv + j] = o0; WB[(3 * ci + 1) * nv + j] = o1; WB[(3 * ci + 2) * nv + j] = o2;
  }
  __syncwarp();
  const float* M = WF(M); float* H = WF(H);
  const float* eqc = WF(eqc);
  if (L.sparse) {
#pragma unroll 1
    for (int e = c.lane; e < L.nnz; e += G) {
      const int ij = MI(mpair)[e];
      const int i = ij >> 16, j = ij & 0xffff;
      float s = M[e];
      if (i == j) {
        const int lr = MI(dof_limrow)[i];
        if (lr >= 0) { const int r = L.ne + lr; if (Jaref[r] < 0.f) s += D[r] * WF(lims)[lr] * WF(lims)[lr]; }
      }
      const int h1 = MI(he_adr)[e + 1];
#pragma unroll 1
      for (int q = MI(he_adr)[e]; q < h1; q++) {
        const int code = MI(he_eq)[q];
        const int eq = code >> 2, which = code & 3;
        const float cf = eqc[eq];
        s += D[eq] * (which == 0 ? 1.f : (which == 1 ? cf * cf : cf));
      }
      const int q1 = MI(hc_adr)[e + 1];
#pragma unroll 1
      for (int q = MI(hc_adr)[e]; q < q1; q++) {
        const int r3 = 3 * MI(hc_con)[q] * nv;
        s += B[r3 + i] * WB[r3 + j] + B[r3 + nv + i] * WB[r3 + nv + j] + B[r3 + 2 * nv + i] * WB[r3 + 2 * nv + j];
      }
      H[e] = s;
    }
    __syncwarp();
    return;
  }
#pragma unroll 1
  for (int k = c.lane; k < L.ntri; k += G) {
    const int ij = MI(tri)[k];
    const int i = ij >> 16, j = ij & 0xffff;
    float s = M[k];
    if (i == j) {
      const int lr = MI(dof_limrow)[i];
      if (lr >= 0) { const int r = L.ne + lr; if (Jaref[r] < 0.f) s += D[r] * WF(lims)[lr] * WF(lims)[lr]; }
    }
    for (int e = 0; e < L.ne; e++) {
      const int j1 = MI(eq_j1)[e], j2 = MI(eq_j2)[e];
      const int d1 = MI(jnt_dofadr)[j1], d2 = (j2 >= 0) ? MI(jnt_dofadr)[j2] : -1;
      const float vi = (i == d1) ? 1.f : ((i == d2) ? eqc[e] : 0.f);
      const float vj = (j == d1) ? 1.f : ((j == d2) ? eqc[e] : 0.f);
      s += D[e] * vi * vj;
    }
#pragma unroll 1
    for (int it = 0; it < 3 * L.ncon; it++) s += B[it * nv + i] * WB[it * nv + j];
    H[k] = s;
  }
  __syncwarp();
}

// one point of the exact line search: (cost, d0, d1) at alpha; group-uniform
template <int G> __device__ __forceinline__ LSPoint ls_eval(const Ctx& c, float alpha, const float* Jaref, const float* jv,
                                                            float qg0, float qg1, float qg2) {
  const Layout& L = c.L;
  const float* D = WF(D);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f;
  for (int r = c.lane; r < L.nefc; r += G) {
    const float ja = Jaref[r], v = jv[r];
    const float x = ja + alpha * v;
    if (r < L.ne || x < 0.f) {
      const float Dr = D[r];
      q0 += 0.5f * ja * ja * Dr; q1 += v * ja * Dr; q2 += 0.5f * v * v * Dr;
    }
  }
  q0 = gsum<G>(q0) + qg0; q1 = gsum<G>(q1) + qg1; q2 = gsum<G>(q2) + qg2;
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = 2.f * alpha * q2 + q1;
  p.d1 = 2.f * q2 + ((q2 == 0.f) ? kMinVal : 0.f);
  return p;
}
__device__ __forceinline__ float safe_div(float a, float b) { return a / (b + ((b == 0.f) ? kMinVal : 0.f)); }

// solver._linesearch; `live` = this group still iterates (commits are predicated on it so that
// groups sharing a warp can run different trip counts convergently).
template <int G> __device__ void solver_linesearch(const Ctx& c, float gauss, bool live) {
  const Layout& L = c.L;
  float* a = WF(a); float* Ma = WF(Ma); float* Jaref = WF(Jaref);
  const float* search = WF(search); float* mv = WF(mv); float* jv = WF(jv); const float* fs = WF(fs);
  la_mul_m<G>(c, search, mv);
  mul_j<G>(c, search, jv);  // syncs
  float sn = 0.f, sMa = 0.f, sq = 0.f, smv = 0.f;
  for (int d = c.lane; d < L.nv; d += G) {
    const float s = search[d];
    sn += s * s; sMa += s * Ma[d]; sq += s * fs[d]; smv += s * mv[d];
  }
  sn = gsum<G>(sn); sMa = gsum<G>(sMa); sq = gsum<G>(sq); smv = gsum<G>(smv);
  const float smag = sqrtf(sn) * L.meaninertia * (float)max(1, L.nv);
  const float gtol = L.tolerance * L.ls_tolerance * smag;
  const float qg0 = gauss, qg1 = sMa - sq, qg2 = 0.5f * smv;
  const LSPoint p0 = ls_eval<G>(c, 0.f, Jaref, jv, qg0, qg1, qg2);
  const LSPoint l0 = ls_eval<G>(c, -safe_div(p0.d0, p0.d1), Jaref, jv, qg0, qg1, qg2);
  const bool lesser = l0.d0 < p0.d0;
  LSPoint hi = lesser ? p0 : l0;
  LSPoint lo = lesser ? l0 : p0;
  bool swap = true;
  int it = 0;
  while (true) {
    bool done = it >= L.ls_iterations;
    done = done || !swap;
    done = done || ((lo.d0 < 0.f) && (lo.d0 > -gtol));
    done = done || ((hi.d0 > 0.f) && (hi.d0 < gtol));
    if (!__any_sync(ABR_FULL, !done)) break;
    const LSPoint lo_next = ls_eval<G>(c, lo.alpha - safe_div(lo.d0, lo.d1), Jaref, jv, qg0, qg1, qg2);
    const LSPoint hi_next = ls_eval<G>(c, hi.alpha - safe_div(hi.d0, hi.d1), Jaref, jv, qg0, qg1, qg2);
    const LSPoint mid = ls_eval<G>(c, 0.5f * (lo.alpha + hi.alpha), Jaref, jv, qg0, qg1, qg2);
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
  const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  const float alpha = (lo.cost < hi.cost) ? lo.alpha : hi.alpha;
  if (improved && live) {
    for (int d = c.lane; d < L.nv; d += G) { a[d] += search[d] * alpha; Ma[d] += mv[d] * alpha; }
    for (int r = c.lane; r < L.nefc; r += G) Jaref[r] += jv[r] * alpha;
  }
  __syncwarp();
}

// solver.solve: on exit WF(a) = qacc, WF(warm) = qacc, WF(force)/WF(fc) = efc_force /
// qfrc_constraint of the final point. The loop is arranged so that every helper has ONE call
// site: pass k first refreshes the constraint forces (MJX update_constraint of the previous
// body, or of Context.create for k = 0), stops if the iteration budget is spent, then forms the
// gradient / search direction (update_gradient) and runs the line search. Groups sharing a warp
// may converge at different passes: `live` predicates every commit.
template <int G> __device__ void stage_solve(const Ctx& c) {
  const Layout& L = c.L;
  const int nv = L.nv, nefc = L.nefc;
  float* a = WF(a); float* Ma = WF(Ma); float* Jaref = WF(Jaref);
  float* mv = WF(mv); float* jv = WF(jv); float* search = WF(search); float* grad = WF(grad);
  const float* as = WF(as); const float* aref = WF(aref); float* warm = WF(warm);
  const float* M = WF(M); const float* fs = WF(fs);
  float gauss = 0.f, cost = 0.f;
  // ---- warm start: cheaper of qacc_smooth (candidate 0) and qacc_warmstart (candidate 1)
  const int ncand = (L.disableflags & ABR_DSBL_WARMSTART) ? 1 : 2;
  bool use_warm = false;
  for (int cand = 0; cand < ncand; cand++) {
    const float* v = cand ? warm : as;
    float* oM = cand ? mv : Ma;
    float* oJ = cand ? jv : Jaref;
    la_mul_m<G>(c, v, oM);
    mul_j<G>(c, v, oJ);
    for (int r = c.lane; r < nefc; r += G) oJ[r] -= aref[r];
    __syncwarp();
    float g2;
    const float c2 = solver_cost<G>(c, v, oM, oJ, &g2);
    if (cand == 0) { cost = c2; gauss = g2; }
    else if (c2 < cost) { use_warm = true; cost = c2; gauss = g2; }
  }
  for (int d = c.lane; d < nv; d += G) { a[d] = use_warm ? warm[d] : as[d]; if (use_warm) Ma[d] = mv[d]; }
  if (use_warm) for (int r = c.lane; r < nefc; r += G) Jaref[r] = jv[r];
  __syncwarp();
  float prev_cost = INFINITY;
  const float scale = 1.f / (L.meaninertia * (float)max(1, nv));
  float* pgrad = c.W + L.w_rk;        // CG Polak-Ribiere history
  float* pmgrad = c.W + L.w_rk + nv;
  bool live = true;
  for (int niter = 0;; niter++) {
    solver_forces<G>(c, Jaref);
    if (niter >= L.iterations) break;
    // gradient, preconditioned gradient
    for (int d = c.lane; d < nv; d += G) grad[d] = Ma[d] - fs[d] - WF(fc)[d];
    __syncwarp();
    if (L.solver == ABR_SOLVER_NEWTON) {
      solver_hessian<G>(c, Jaref);
      la_factor<G>(c, WF(H));
    }
    float* mg = mv;  // free outside the line search
    la_solve<G>(c, WF(H), grad, mg);
    if (L.iterations != 1) {
      float gn = 0.f;
      for (int d = c.lane; d < nv; d += G) gn += grad[d] * grad[d];
      gn = gsum<G>(gn);
      bool done = scale * (prev_cost - cost) < L.tolerance;
      done = done || (scale * sqrtf(gn) < L.tolerance);
      live = live && !done;
      if (!__any_sync(ABR_FULL, live)) break;
    }
    // search direction
    if (L.solver == ABR_SOLVER_NEWTON || niter == 0) {
      if (live) for (int d = c.lane; d < nv; d += G) search[d] = -mg[d];
    } else {
      float num = 0.f, den = 0.f;
      for (int d = c.lane; d < nv; d += G) { num += grad[d] * (mg[d] - pmgrad[d]); den += pgrad[d] * pmgrad[d]; }
      num = gsum<G>(num); den = gsum<G>(den);
      const float beta = fmaxf(0.f, num / fmaxf(kMinVal, den));
      if (live) for (int d = c.lane; d < nv; d += G) search[d] = -mg[d] + beta * search[d];
    }
    if (L.solver == ABR_SOLVER_CG && live)
      for (int d = c.lane; d < nv; d += G) { pgrad[d] = grad[d]; pmgrad[d] = mg[d]; }
    __syncwarp();
    solver_linesearch<G>(c, gauss, live);
    if (L.iterations != 1) {
      float g2;
      const float c2 = solver_cost<G>(c, a, Ma, Jaref, &g2);
      if (live) { prev_cost = cost; cost = c2; gauss = g2; }
    }
  }
  for (int d = c.lane; d < nv; d += G) warm[d] = a[d];
  __syncwarp();
}

// ------------------------------------------------------------------------------ forward / step
// mjx.forward: on exit WF(a) = qacc, WF(warm) updated, WF(fs) = qfrc_smooth, WF(fc) = qfrc_constraint
template <int G> __device__ __forceinline__ void forward(const Ctx& c) {
  const Layout& L = c.L;
  stage_kinematics<G>(c);
  stage_com<G>(c);
  stage_collision<G>(c);
  stage_crb<G>(c);
  const int nent = L.sparse ? L.nnz : L.ntri;
#pragma unroll 1
  for (int k = c.lane; k < nent; k += G) WF(H)[k] = WF(M)[k];
  __syncwarp();
  la_factor<G>(c, WF(H));
  stage_velocity<G>(c);
  la_solve<G>(c, WF(H), WF(fs), WF(as));
  if (L.nefc == 0) {
    for (int d = c.lane; d < L.nv; d += G) { WF(a)[d] = WF(as)[d]; WF(fc)[d] = 0.f; }
    __syncwarp();
    return;
  }
  stage_rows<G>(c);
  stage_solve<G>(c);  // CG keeps L_M in WF(H)
}

// forward._integrate_pos for all joints: q <- q (+) dt * v
template <int G> __device__ void integrate_pos(const Ctx& c, float* q, const float* v, float dt) {
  const Layout& L = c.L;
  for (int j = c.lane; j < L.njnt; j += G) {
    const int a = MI(jnt_qposadr)[j], d = MI(jnt_dofadr)[j];
    if (MI(jnt_type)[j] == ABR_JNT_FREE) {
      q[a] += dt * v[d]; q[a + 1] += dt * v[d + 1]; q[a + 2] += dt * v[d + 2];
      float w[3] = {v[d + 3], v[d + 4], v[d + 5]};
      const float nrm = v_normalize(w, 3);
      float ql[4], qn[4];
      axis_angle_quat(w, dt * nrm, ql);
      q_mul(q + a + 3, ql, qn);
      v_normalize(qn, 4);
      q[a + 3] = qn[0]; q[a + 4] = qn[1]; q[a + 5] = qn[2]; q[a + 6] = qn[3];
    } else {
      q[a] += dt * v[d];
    }
  }
}

// What follows a forward evaluation. `stage` counts the forward evaluations of the current step:
// Euler has one (stage 0 -> integrate); RK4 has four (stage 0 saves state and sets up the first
// intermediate point, stages 1,2 accumulate and move on, stage 3 accumulates and advances).
// Returns true when the step is complete. Single call site for forward() in every kernel.
template <int G> __device__ bool post_forward(const Ctx& c, int stage) {
  const Layout& L = c.L;
  const int nq = L.nq, nv = L.nv;
  const float dt = L.timestep;
  float* a = WF(a); float* qpos = WF(qpos); float* qvel = WF(qvel); float* warm = WF(warm);
  if (L.integrator != ABR_INT_RK4) {
    // forward.euler (+ implicit joint damping unless EULERDAMP is disabled)
    if (!(L.disableflags & ABR_DSBL_EULERDAMP)) {
      float* H = WF(H); const float* M = WF(M);
      const int nent = L.sparse ? L.nnz : L.ntri;
      const int* tab = L.sparse ? MI(mpair) : MI(tri);
#pragma unroll 1
      for (int k = c.lane; k < nent; k += G) {
        const int ij = tab[k];
        H[k] = M[k] + (((ij >> 16) == (ij & 0xffff)) ? c.dps * MF(dof_damping)[ij >> 16] * dt : 0.f);
      }
      float* rhs = WF(grad);
      for (int d = c.lane; d < nv; d += G) rhs[d] = WF(fs)[d] + WF(fc)[d];
      __syncwarp();
      la_factor<G>(c, H);
      la_solve<G>(c, H, rhs, a);
    }
    for (int d = c.lane; d < nv; d += G) qvel[d] += a[d] * dt;
    __syncwarp();
    integrate_pos<G>(c, qpos, qvel, dt);
    __syncwarp();
    return true;
  }
  // forward.rungekutta4; save area sits after the CG history
  float* q0 = c.W + L.w_rk + 2 * nv; float* v0 = q0 + nq; float* w0 = v0 + nv;
  float* sv = w0 + nv; float* sa = sv + nv; float* kv = sa + nv; float* dv = kv + nv;
  const float A[4] = {0.5f, 0.5f, 1.0f, 0.f};
  const float Bc[4] = {1.f / 6.f, 1.f / 3.f, 1.f / 3.f, 1.f / 6.f};
  if (stage == 0) {
    for (int i = c.lane; i < nq; i += G) q0[i] = qpos[i];
    for (int d = c.lane; d < nv; d += G) { v0[d] = qvel[d]; w0[d] = warm[d]; kv[d] = qvel[d]; sv[d] = 0.f; sa[d] = 0.f; }
    __syncwarp();
  }
  for (int d = c.lane; d < nv; d += G) { sv[d] += Bc[stage] * kv[d]; sa[d] += Bc[stage] * a[d]; }
  __syncwarp();
  if (stage < 3) {
    for (int d = c.lane; d < nv; d += G) dv[d] = A[stage] * kv[d];
    for (int i = c.lane; i < nq; i += G) qpos[i] = q0[i];
    __syncwarp();
    integrate_pos<G>(c, qpos, dv, dt);
    for (int d = c.lane; d < nv; d += G) { kv[d] = v0[d] + A[stage] * a[d] * dt; qvel[d] = kv[d]; }
    __syncwarp();
    return false;
  }
  for (int i = c.lane; i < nq; i += G) qpos[i] = q0[i];
  for (int d = c.lane; d < nv; d += G) { qvel[d] = v0[d] + sa[d] * dt; warm[d] = w0[d]; }
  __syncwarp();
  integrate_pos<G>(c, qpos, sv, dt);
  __syncwarp();
  return true;
}

// make_data-like initialisation of the constant parts of a world region
template <int G> __device__ void init_world(const Ctx& c) {
  const Layout& L = c.L;
  for (int i = c.lane; i < L.world_stride; i += G) c.W[i] = 0.f;
  __syncwarp();
}

}  // namespace abr
#endif
