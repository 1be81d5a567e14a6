/* A plain C99 client of the engine's C ABI (include/abr.h): no Python, no torch, no C++.
 * Builds the flattened model of a damped pendulum (one hinge about y, a point mass 1 kg at 0.5 m below the pivot) by hand,
 * uploads it, rolls 8 worlds for 200 steps from different start angles through abr_rollout_host and checks the physics it must obey:
 * every world finite, the pendulum released at rest never rises above its release height, and two identical worlds agree bit for bit.
 * Exit codes: 0 = all checks passed (or, without a CUDA device, the library refused with ABR_ENODEVICE as documented), 1 = failure.
 * Built and run by tests/test_abi.py::test_plain_c_client. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "abr.h"

#define NB 2
#define W 8
#define N 200

int main(void) {
  AbrModelHost h;
  memset(&h, 0, sizeof(h));
  h.nq = 1; h.nv = 1; h.nu = 1; h.na = 0; h.nbody = NB; h.njnt = 1; h.ngeom = 0; h.neq = 0; h.npair = 0; h.nvert = 0;
  h.nface = 0; h.nfacevert = 0; h.nedge = 0;
  h.opt.timestep = 0.005f; h.opt.impratio = 1.f; h.opt.tolerance = 1e-8f; h.opt.ls_tolerance = 0.01f;
  h.opt.gravity[0] = 0.f; h.opt.gravity[1] = 0.f; h.opt.gravity[2] = -9.81f;
  h.opt.integrator = ABR_INT_EULER; h.opt.cone = 0; h.opt.jacobian = 0; h.opt.solver = ABR_SOLVER_NEWTON;
  h.opt.iterations = 1; h.opt.ls_iterations = 5; h.opt.disableflags = 0;
  const float m = 1.f, l = 0.5f, Icom = 0.001f;          /* point-like bob with a small inertia about its own centre */
  const float Iaxis = Icom + m * l * l;                   /* inertia about the hinge */
  h.opt.meaninertia = Iaxis;
  static const int body_parentid[NB] = {0, 0}, body_rootid[NB] = {0, 1}, body_jntnum[NB] = {0, 1}, body_jntadr[NB] = {-1, 0};
  static const int body_dofnum[NB] = {0, 1}, body_dofadr[NB] = {-1, 0};
  static const float body_pos[3 * NB] = {0, 0, 0, 0, 0, 1.f}, body_quat[4 * NB] = {1, 0, 0, 0, 1, 0, 0, 0};
  static const float body_ipos[3 * NB] = {0, 0, 0, 0, 0, -0.5f}, body_iquat[4 * NB] = {1, 0, 0, 0, 1, 0, 0, 0};
  static const float body_mass[NB] = {0, 1.f}, body_subtreemass[NB] = {1.f, 1.f}, body_inertia[3 * NB] = {0, 0, 0, 0.001f, 0.001f, 0.001f};
  float body_invweight0[2 * NB] = {0, 0, 0, 0};
  body_invweight0[2] = l * l / Iaxis / 3.f * 2.f; /* translational: two in-plane axes see lever^2 / I, averaged over three */
  body_invweight0[3] = 1.f / Iaxis / 3.f;         /* rotational: one of three axes */
  h.body_parentid = body_parentid; h.body_rootid = body_rootid; h.body_jntnum = body_jntnum; h.body_jntadr = body_jntadr;
  h.body_dofnum = body_dofnum; h.body_dofadr = body_dofadr; h.body_pos = body_pos; h.body_quat = body_quat; h.body_ipos = body_ipos;
  h.body_iquat = body_iquat; h.body_mass = body_mass; h.body_subtreemass = body_subtreemass; h.body_inertia = body_inertia;
  h.body_invweight0 = body_invweight0;
  static const int jnt_type[1] = {ABR_JNT_HINGE}, jnt_qposadr[1] = {0}, jnt_dofadr[1] = {0}, jnt_bodyid[1] = {1}, jnt_limited[1] = {0};
  static const float jnt_solref[2] = {0.02f, 1.f}, jnt_solimp[5] = {0.9f, 0.95f, 0.001f, 0.5f, 2.f}, jnt_pos[3] = {0, 0, 0}, jnt_axis[3] = {0, 1, 0};
  static const float jnt_stiffness[1] = {0}, jnt_range[2] = {0, 0}, jnt_margin[1] = {0};
  h.jnt_type = jnt_type; h.jnt_qposadr = jnt_qposadr; h.jnt_dofadr = jnt_dofadr; h.jnt_bodyid = jnt_bodyid; h.jnt_limited = jnt_limited;
  h.jnt_solref = jnt_solref; h.jnt_solimp = jnt_solimp; h.jnt_pos = jnt_pos; h.jnt_axis = jnt_axis; h.jnt_stiffness = jnt_stiffness;
  h.jnt_range = jnt_range; h.jnt_margin = jnt_margin;
  static const int dof_bodyid[1] = {1}, dof_jntid[1] = {0}, dof_parentid[1] = {-1};
  static const float dof_armature[1] = {0}, dof_damping[1] = {0.05f};
  float dof_invweight0[1];
  dof_invweight0[0] = 1.f / Iaxis;
  h.dof_bodyid = dof_bodyid; h.dof_jntid = dof_jntid; h.dof_parentid = dof_parentid; h.dof_armature = dof_armature;
  h.dof_damping = dof_damping; h.dof_invweight0 = dof_invweight0;
  static const int none_i[1] = {0};
  static const float none_f[16] = {0};
  h.geom_type = none_i; h.geom_bodyid = none_i; h.geom_size = none_f; h.geom_pos = none_f; h.geom_quat = none_f;
  h.geom_vertadr = none_i; h.geom_vertnum = none_i; h.vert = none_f;
  h.geom_faceadr = none_i; h.geom_facenum = none_i; h.face_vertadr = none_i; h.face_vertnum = none_i; h.face_vert = none_i; h.face_normal = none_f;
  h.geom_edgeadr = none_i; h.geom_edgenum = none_i; h.edge_vert = none_i;
  h.pair_geom1 = none_i; h.pair_geom2 = none_i; h.pair_kind = none_i; h.pair_condim = none_i; h.pair_friction = none_f;
  h.pair_solref = none_f; h.pair_solimp = none_f; h.pair_includemargin = none_f;
  h.eq_type = none_i; h.eq_obj1id = none_i; h.eq_obj2id = none_i; h.eq_active = none_i; h.eq_solref = none_f; h.eq_solimp = none_f; h.eq_data = none_f;
  static const int act_trnid[1] = {0}, act_gaintype[1] = {ABR_GAIN_FIXED}, act_biastype[1] = {ABR_BIAS_NONE}, act_ctrllimited[1] = {1}, act_forcelimited[1] = {0};
  static const float act_ctrlrange[2] = {-2.f, 2.f}, act_forcerange[2] = {0, 0}, act_gainprm[3] = {1.f, 0, 0}, act_biasprm[3] = {0, 0, 0}, act_gear[1] = {1.f};
  h.actuator_trnid = act_trnid; h.actuator_gaintype = act_gaintype; h.actuator_biastype = act_biastype; h.actuator_ctrllimited = act_ctrllimited;
  h.actuator_forcelimited = act_forcelimited; h.actuator_ctrlrange = act_ctrlrange; h.actuator_forcerange = act_forcerange;
  h.actuator_gainprm = act_gainprm; h.actuator_biasprm = act_biasprm; h.actuator_gear = act_gear;
  static const float qpos0[1] = {0}, qpos_spring[1] = {0};
  h.qpos0 = qpos0; h.qpos_spring = qpos_spring;

  AbrModel* model = NULL;
  int rc = abr_model_create(&h, 0, &model);
  if (rc == ABR_ENODEVICE) {
    printf("no CUDA device: abr_model_create refused with ABR_ENODEVICE (%s)\n", abr_last_error());
    return 0;
  }
  if (rc != ABR_OK) { fprintf(stderr, "abr_model_create: %d %s\n", rc, abr_last_error()); return 1; }
  float x0[W * 2], *us = (float*)calloc((size_t)W * N, sizeof(float)), *xs = (float*)malloc(sizeof(float) * W * (N + 1) * 2);
  for (int w = 0; w < W; w++) { x0[2 * w] = 0.2f + 0.3f * (float)(w % 4); x0[2 * w + 1] = 0.f; }  /* worlds w and w + 4 are identical */
  rc = abr_rollout_host(model, x0, 2, us, N, W, N, xs, NULL, NULL);
  if (rc != ABR_OK) { fprintf(stderr, "abr_rollout_host: %d %s\n", rc, abr_last_error()); return 1; }
  int bad = 0;
  for (int w = 0; w < W; w++) {
    const float* x = xs + (size_t)w * (N + 1) * 2;
    float maxabs = 0.f;
    for (int t = 0; t <= N; t++) {
      if (!isfinite(x[2 * t]) || !isfinite(x[2 * t + 1])) bad++;
      if (fabsf(x[2 * t]) > maxabs) maxabs = fabsf(x[2 * t]);
    }
    if (x[0] != x0[2 * w]) bad++;                      /* row 0 of xs is the caller's x0 verbatim */
    if (maxabs > x0[2 * w] * 1.001f) bad++;            /* released at rest, damped: never above the release angle */
    if (fabsf(x[2 * N]) > x0[2 * w]) bad++;
    if (w >= 4 && memcmp(x, xs + (size_t)(w - 4) * (N + 1) * 2, sizeof(float) * (N + 1) * 2) != 0) bad++;  /* identical worlds, identical bits */
    /* the first swing takes about half a period of the physical pendulum: 2 pi sqrt(I / (m g l)) = 1.42 s -> 0.71 s = 142 steps */
    int tmin = 0;
    for (int t = 1; t <= N; t++) if (x[2 * t] < x[2 * tmin]) tmin = t;
    if (tmin < 120 || tmin > 170) bad++;
  }
  printf("plain C client: %d worlds x %d steps, %d failed checks; world 0 angle %.4f -> %.4f\n", W, N, bad, xs[0], xs[2 * N]);
  abr_model_destroy(model);
  free(us); free(xs);
  return bad ? 1 : 0;
}
