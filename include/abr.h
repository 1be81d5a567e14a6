/*
 * abr.h — C ABI of the B200 batched rigid-body rollout engine ("abr").
 *
 * This is the drop-in boundary for the one hot path ambersim drives: thousands of
 * independent mjx.step rollouts. The reference has no FFI layer of its own; the seam this
 * ABI replaces is the Python call boundary into MJX (paths relative to the reference tree):
 *
 *   mjx.device_put(mj.MjModel) -> mjx.Model     ambersim/utils/io_utils.py:225, ambersim/rl/base.py:52
 *   mjx.make_data(mjx.Model)   -> mjx.Data      ambersim/utils/io_utils.py:226, ambersim/trajopt/shooting.py:34
 *   mjx.forward(Model, Data)   -> Data          ambersim/trajopt/shooting.py:36, ambersim/rl/base.py:85
 *   mjx.step(Model, Data)      -> Data          ambersim/trajopt/shooting.py:41, ambersim/rl/base.py:93
 *
 * and the ambersim-level loops built on it:
 *
 *   shoot(m, x0, us) -> xs                               ambersim/trajopt/shooting.py:22-48
 *   VanillaPredictiveSampler.optimize(params)            ambersim/trajopt/shooting.py:119-157
 *   StaticGoalQuadraticCost.cost(xs, us, params)         ambersim/trajopt/cost.py:62-85
 *   MjxEnv.pipeline_init / pipeline_step                 ambersim/rl/base.py:81-96
 *
 * Conventions
 *   - every function is extern "C", returns int (0 = ABR_OK, <0 = error code), never throws;
 *     the message of the last error on the calling thread is abr_last_error().
 *   - all arrays are row-major float32 / int32. Field names of AbrModelHost follow mjx.Model.
 *   - "_dev" entry points take DEVICE pointers and a cudaStream_t (passed as void*); they are
 *     stream-ordered and never synchronise. "_host" entry points take HOST pointers, stage the
 *     copies on the handle's own stream and return after the result is in the host buffers.
 *   - state layout x = [qpos(nq); qvel(nv)]  (shooting.py:35,42), nx = nq + nv.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns ABR_ENODEVICE.
 */
#ifndef ABR_H_
#define ABR_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABR_VERSION 100

/* error codes */
#define ABR_OK 0
#define ABR_EINVAL (-1)       /* bad argument (null pointer, negative size, shape mismatch)      */
#define ABR_EUNSUPPORTED (-2) /* model feature outside the engine (mirrors MJX NotImplementedError,
                                 io_utils.py:228-241): elliptic cone, ball joint, mesh pair, ...   */
#define ABR_ECUDA (-3)        /* CUDA runtime error, text in abr_last_error()                    */
#define ABR_ENODEVICE (-4)    /* no CUDA device: the engine has no CPU path                      */
#define ABR_ECAPACITY (-5)    /* model larger than the compiled kernel limits                    */

/* enums (values = MuJoCo's) */
enum { ABR_JNT_FREE = 0, ABR_JNT_BALL = 1, ABR_JNT_SLIDE = 2, ABR_JNT_HINGE = 3 };
enum { ABR_GEOM_PLANE = 0, ABR_GEOM_SPHERE = 2, ABR_GEOM_CAPSULE = 3, ABR_GEOM_BOX = 6, ABR_GEOM_MESH = 7 };
enum { ABR_INT_EULER = 0, ABR_INT_RK4 = 1 };
enum { ABR_SOLVER_CG = 1, ABR_SOLVER_NEWTON = 2 };
enum { ABR_EQ_JOINT = 2 };
enum { ABR_GAIN_FIXED = 0, ABR_GAIN_AFFINE = 1 };
enum { ABR_BIAS_NONE = 0, ABR_BIAS_AFFINE = 1 };
enum { ABR_PAIR_PLANE_SPHERE = 0, ABR_PAIR_PLANE_CAPSULE = 1, ABR_PAIR_SPHERE_SPHERE = 2,
       ABR_PAIR_SPHERE_CAPSULE = 3, ABR_PAIR_CAPSULE_CAPSULE = 4,
       ABR_PAIR_PLANE_CONVEX = 5,   /* plane vs box / mesh: 4 contacts (mjx collision_convex.plane_convex) */
       ABR_PAIR_SPHERE_CONVEX = 6,  /* sphere vs box / mesh: 1 contact (collision_convex.sphere_convex) */
       ABR_PAIR_CAPSULE_CONVEX = 7, /* capsule vs box / mesh: 2 contacts (collision_convex.capsule_convex) */
       ABR_PAIR_CONVEX_CONVEX = 8   /* box / mesh vs box / mesh: separating axes + face clipping, 4 contacts (collision_convex.convex_convex) */ };
enum { ABR_MAX_FACE_VERTS = 8, ABR_MAX_CONVEX_VERTS = 64 };
/* mjtDisableBit */
enum {
  ABR_DSBL_CONSTRAINT = 1, ABR_DSBL_EQUALITY = 2, ABR_DSBL_FRICTIONLOSS = 4, ABR_DSBL_LIMIT = 8,
  ABR_DSBL_CONTACT = 16, ABR_DSBL_PASSIVE = 32, ABR_DSBL_GRAVITY = 64, ABR_DSBL_CLAMPCTRL = 128,
  ABR_DSBL_WARMSTART = 256, ABR_DSBL_FILTERPARENT = 512, ABR_DSBL_ACTUATION = 1024,
  ABR_DSBL_REFSAFE = 2048, ABR_DSBL_SENSOR = 4096, ABR_DSBL_EULERDAMP = 16384
};

/* model.opt (+ stat.meaninertia). What `model.opt.replace(...)` changes in the reference test
 * (tests/trajopt/test_predictive_sampler.py:22-31). */
/* ABR_STRUCT_BEGIN AbrOpt */
typedef struct AbrOpt {
  float timestep;
  float impratio;
  float tolerance;
  float ls_tolerance;
  float gravity[3];
  float meaninertia;
  int integrator;
  int cone;
  int jacobian;
  int solver;
  int iterations;
  int ls_iterations;
  int disableflags;
  int reserved0;
} AbrOpt;
/* ABR_STRUCT_END */

/* Flattened MjModel, structure-of-arrays, host side. Produced by the loader
 * (ambersim_b200/utils/io_utils.py: mj_to_mjx_model_and_data, replacing io_utils.py:222-241). */
/* ABR_STRUCT_BEGIN AbrModelHost */
typedef struct AbrModelHost {
  int nq;
  int nv;
  int nu;
  int na;
  int nbody;
  int njnt;
  int ngeom;
  int neq;
  int npair;       /* statically enumerated colliding geom pairs (MJX enumerates at trace time) */
  int nvert;       /* vertices of all convex (box / mesh) geoms */
  int nface;       /* polygon faces of all convex geoms */
  int nfacevert;   /* polygon corners of all faces */
  int nedge;       /* unique hull edges of all convex geoms */
  AbrOpt opt;
  /* bodies [nbody] */
  const int* body_parentid;
  const int* body_rootid;
  const int* body_jntnum;
  const int* body_jntadr;
  const int* body_dofnum;
  const int* body_dofadr;
  const float* body_pos;        /* [nbody,3] */
  const float* body_quat;       /* [nbody,4] (w,x,y,z) */
  const float* body_ipos;       /* [nbody,3] */
  const float* body_iquat;      /* [nbody,4] */
  const float* body_mass;       /* [nbody]   */
  const float* body_subtreemass;/* [nbody]   */
  const float* body_inertia;    /* [nbody,3] */
  const float* body_invweight0; /* [nbody,2] */
  /* joints [njnt] */
  const int* jnt_type;
  const int* jnt_qposadr;
  const int* jnt_dofadr;
  const int* jnt_bodyid;
  const int* jnt_limited;
  const float* jnt_solref;      /* [njnt,2] */
  const float* jnt_solimp;      /* [njnt,5] */
  const float* jnt_pos;         /* [njnt,3] */
  const float* jnt_axis;        /* [njnt,3] */
  const float* jnt_stiffness;   /* [njnt]   */
  const float* jnt_range;       /* [njnt,2] */
  const float* jnt_margin;      /* [njnt]   */
  /* dofs [nv] */
  const int* dof_bodyid;
  const int* dof_jntid;
  const int* dof_parentid;
  const float* dof_armature;
  const float* dof_damping;
  const float* dof_invweight0;
  /* geoms [ngeom] */
  const int* geom_type;
  const int* geom_bodyid;
  const float* geom_size;       /* [ngeom,3] */
  const float* geom_pos;        /* [ngeom,3] */
  const float* geom_quat;       /* [ngeom,4] */
  const int* geom_vertadr;      /* [ngeom] first vertex of a box / mesh geom in `vert` */
  const int* geom_vertnum;      /* [ngeom] number of vertices (0 for non-convex-set geoms) */
  const float* vert;            /* [nvert,3] convex-hull vertices in the geom frame (boxes: the 8 corners) */
  /* hull topology of the convex geoms (sphere / capsule / convex - convex pairs): polygon faces, counter-clockwise from outside */
  const int* geom_faceadr;      /* [ngeom] first face of the geom in the face tables */
  const int* geom_facenum;      /* [ngeom] number of faces (0: no 3-D hull) */
  const int* face_vertadr;      /* [nface] first corner of the face in `face_vert` */
  const int* face_vertnum;      /* [nface] corners of the face (3 .. ABR_MAX_FACE_VERTS) */
  const int* face_vert;         /* [nfacevert] vertex ids LOCAL to the geom's vertex set */
  const float* face_normal;     /* [nface,3] outward unit normal, geom frame */
  const int* geom_edgeadr;      /* [ngeom] first edge of the geom in `edge_vert` */
  const int* geom_edgenum;      /* [ngeom] number of unique hull edges */
  const int* edge_vert;         /* [nedge,2] local vertex ids of an edge's end points */
  /* static contact pairs [npair]; mixing of friction/solref/solimp/margin done by the loader */
  const int* pair_geom1;
  const int* pair_geom2;
  const int* pair_kind;         /* ABR_PAIR_* */
  const int* pair_condim;       /* 1 or 3 */
  const float* pair_friction;   /* [npair,5] */
  const float* pair_solref;     /* [npair,2] */
  const float* pair_solimp;     /* [npair,5] */
  const float* pair_includemargin; /* [npair] margin - gap */
  /* equality [neq] */
  const int* eq_type;
  const int* eq_obj1id;
  const int* eq_obj2id;
  const int* eq_active;
  const float* eq_solref;       /* [neq,2]  */
  const float* eq_solimp;       /* [neq,5]  */
  const float* eq_data;         /* [neq,11] */
  /* actuators [nu] (joint transmission only) */
  const int* actuator_trnid;    /* [nu] joint id */
  const int* actuator_gaintype;
  const int* actuator_biastype;
  const int* actuator_ctrllimited;
  const int* actuator_forcelimited;
  const float* actuator_ctrlrange;  /* [nu,2] */
  const float* actuator_forcerange; /* [nu,2] */
  const float* actuator_gainprm;    /* [nu,3] */
  const float* actuator_biasprm;    /* [nu,3] */
  const float* actuator_gear;       /* [nu]   gear[0] */
  /* reference configuration */
  const float* qpos0;           /* [nq] */
  const float* qpos_spring;     /* [nq] */
} AbrModelHost;
/* ABR_STRUCT_END */

/* Dense quadratic tracking cost, StaticGoalQuadraticCost (ambersim/trajopt/cost.py:13-85):
 *   0.5 * [ sum_{t<N} (x_t-xg)' Q (x_t-xg) + (x_N-xg)' Qf (x_N-xg) + sum_{t<N} u_t' R u_t ]
 * HOST pointers; copied to the device by abr_cost_create. */
/* ABR_STRUCT_BEGIN AbrQuadCostHost */
typedef struct AbrQuadCostHost {
  int nx;
  int nu;
  const float* Q;   /* [nx,nx] */
  const float* Qf;  /* [nx,nx] */
  const float* R;   /* [nu,nu] */
  const float* xg;  /* [nx]    */
} AbrQuadCostHost;
/* ABR_STRUCT_END */

typedef struct AbrModel AbrModel; /* opaque device-resident model handle */
typedef struct AbrCost AbrCost;   /* opaque device-resident cost handle  */

const char* abr_last_error(void);
int abr_version(void);
size_t abr_sizeof_model_host(void);
size_t abr_sizeof_opt(void);
int abr_device_count(void);

/* replaces mjx.device_put(mj_model) (io_utils.py:225). Validates feature support. */
int abr_model_create(const AbrModelHost* host, int device, AbrModel** out);
int abr_model_destroy(AbrModel* m);
/* replaces model.replace(opt=model.opt.replace(...)) (test_predictive_sampler.py:22-31) */
int abr_model_set_opt(AbrModel* m, const AbrOpt* opt);
int abr_model_get_opt(const AbrModel* m, AbrOpt* opt);
/* static sizes derived at create: ncon, ne, nl, nefc, tree depth, lanes per world chosen */
int abr_model_info(const AbrModel* m, int* ncon, int* ne, int* nl, int* nefc, int* depth);
/* one line of text: which kernel family and variant serves this model with its current options (e.g.
 * "limb kernels <NL=3, NC=1, flat 4-lane pattern>, 4 lanes per world, fast (eulerdamp off) variant") */
int abr_model_describe(const AbrModel* m, char* buf, int cap);
/* kernel family: 0 = auto (limb kernels, one lane per root-to-leaf path, when the model is eligible:
 * floating base, one hinge/slide joint per other body, plane-sphere contacts, Newton + Euler, no
 * equalities; otherwise the generic kernels), 1 = limb kernels or ABR_EUNSUPPORTED,
 * 4/8/16/32 = generic kernels with that many lanes per world */
int abr_model_set_lanes(AbrModel* m, int lanes);
/* Host-only (no device needed): the limb-path plan the engine derives for a model.
 * info[8] = {eligible, log2(lanes per world), chain length NL of the serving kernel, contacts per path NC,
 *            sharing pattern (2 bits per chain position), nefc, ncon, lanes actually carrying a path};
 * lane_body[8*(NL+1)] = body id at (lane, chain position) or -1 (padding / dummy lane);
 * lane_own[8] / lane_lvl[8] = per-lane owner bits and sharing levels (2 bits per position).
 * Arrays may be NULL; cap = ints available in lane_body. */
int abr_limb_plan_host(const AbrModelHost* host, int* info, int* lane_body, int cap, int* lane_own, int* lane_lvl);

/* Memory: the handle owns its scratch (per-sample costs, kept sample trajectories, staging of the *_host calls, the
 * slice carry) and grows it on first use; steady-state calls allocate nothing. abr_model_reserve sizes it up front
 * for rollouts of nworld worlds x N steps and solves of B problems x S samples x N steps, so that no later call of
 * at most those sizes allocates (the role of a workspace query + caller-provided workspace). */
int abr_model_reserve(AbrModel* m, int nworld, int N, int B, int S);
/* Caller-provided workspace for the stream-ordered calls ("no hidden allocation on the hot path"): abr_workspace_bytes
 * returns the DEVICE bytes abr_predictive_sample_dev / abr_mpc_dev need for solves of up to B problems x S samples x N steps;
 * abr_model_set_workspace hands the handle such a buffer (caller-owned, 16-byte aligned, must outlive the calls). From then on
 * those calls never call cudaMalloc / cudaFree: a larger solve returns ABR_ECAPACITY instead (a sweep whose kept trajectories do
 * not fit re-rolls its winners). workspace == NULL returns to handle-owned scratch. The rollout / env / forward _dev calls
 * use no scratch at all; the *_host calls stage through handle-owned buffers on the handle's own stream. */
int abr_workspace_bytes(const AbrModel* m, int N, int B, int S, size_t* bytes);
int abr_model_set_workspace(AbrModel* m, void* workspace, size_t bytes, int N, int B, int S);

int abr_cost_create(const AbrQuadCostHost* host, int device, AbrCost** out);
int abr_cost_destroy(AbrCost* c);

/* ---- shoot (shooting.py:22-48) for nworld independent worlds, DEVICE pointers -----------
 * x0        [nworld,nx] if x0_stride==nx, or [nx] shared if x0_stride==0
 * us        [nworld,N,nu]  (us_stride = N*nu), or shared guess when us_stride==0
 * xs_out    nullable [nworld,N+1,nx]; row 0 = caller's x0 verbatim (shooting.py:47)
 * cost      nullable; costs_out nullable [nworld] (fused cost.py:62-85)
 * Does make_data + forward (ctrl = 0) to seed qacc_warmstart (shooting.py:34-36), then N steps. */
int abr_rollout_dev(AbrModel* m, const float* x0, int x0_stride, const float* us, int us_stride,
                    int nworld, int N, float* xs_out, const AbrCost* cost, float* costs_out,
                    void* stream);
/* same, HOST pointers (copies inside) */
int abr_rollout_host(AbrModel* m, const float* x0, int x0_stride, const float* us, int us_stride,
                     int nworld, int N, float* xs_out, const AbrCost* cost, float* costs_out);

/* ---- VanillaPredictiveSampler.optimize (shooting.py:119-157), batch of B problems --------
 * x0 [B,nx]; us_guess [B,N,nu]; noise: nullable [B,S-1,N,nu] standard normals supplied by the
 * caller (parity mode: pass jax.random.normal's values); NULL => counter-based on-device normals
 * keyed by (seed, b, global_sample, t, u) so results do not depend on the GPU count.
 * sample_offset/S_total: this rank evaluates global samples [sample_offset, sample_offset+S) of
 * S_total (sample 0 = un-noised guess, shooting.py:140-142). Clips to actuator_ctrlrange
 * unconditionally (shooting.py:146-148). argmin = first minimum, NaN counts as minimum (jnp.argmin).
 * outputs: xs_star [B,N+1,nx], us_star [B,N,nu], best_idx [B] (GLOBAL sample index),
 * best_cost [B]; costs_out nullable [B,S]. */
int abr_predictive_sample_dev(AbrModel* m, const AbrCost* cost, const float* x0,
                              const float* us_guess, const float* noise,
                              unsigned long long seed, int B, int S, int N, float stdev,
                              int sample_offset, int S_total, float* xs_star, float* us_star,
                              int* best_idx, float* best_cost, float* costs_out, void* stream);
int abr_predictive_sample_host(AbrModel* m, const AbrCost* cost, const float* x0,
                               const float* us_guess, const float* noise,
                               unsigned long long seed, int B, int S, int N, float stdev,
                               int sample_offset, int S_total, float* xs_star, float* us_star,
                               int* best_idx, float* best_cost, float* costs_out);

/* ---- MjxEnv.pipeline_init / pipeline_step (rl/base.py:81-96), E envs, in place, DEVICE ---
 * state SoA: qpos [E,nq], qvel [E,nv], qacc_warmstart [E,nv], time [E].
 * abr_forward_dev: mjx.forward with ctrl (nullable => 0): writes qacc [E,nv] (nullable) and
 *   qacc_warmstart. abr_env_step_dev: optional auto-reset prologue
 *   (brax AutoResetWrapper: where(done, first_state, state)), then nsubsteps x mjx.step with
 *   ctrl held (rl/base.py:92-95). */
int abr_forward_dev(AbrModel* m, float* qpos, float* qvel, const float* ctrl, float* qacc_warmstart,
                    float* qacc, int E, void* stream);
int abr_env_step_dev(AbrModel* m, float* qpos, float* qvel, float* qacc_warmstart, float* time,
                     const float* ctrl, int E, int nsubsteps, const unsigned char* reset_mask,
                     const float* first_qpos, const float* first_qvel,
                     const float* first_qacc_warmstart, void* stream);

/* ---- derived mjx.Data fields, batched (the fields a user env's compute_obs / compute_reward reads: rl/base.py:98-125) ----
 * Nullable DEVICE pointers, row-major [E, ...], float32, mjx.Data field names and layouts:
 * xpos [E,nbody,3] xquat [E,nbody,4] xipos [E,nbody,3] xanchor [E,njnt,3] xaxis [E,njnt,3] cinert [E,nbody,10]
 * cdof [E,nv,6] cvel [E,nbody,6] cdof_dot [E,nv,6] qfrc_smooth [E,nv] qacc_smooth [E,nv] qfrc_constraint [E,nv]
 * efc_force / efc_D / efc_aref [E,nefc] contact_dist [E,ncon] contact_pos [E,ncon,3] contact_frame [E,ncon,9]
 * (nefc, ncon: abr_model_info; rows in MJX's order equality, limit, contact). */
/* ABR_STRUCT_BEGIN AbrDataFields */
typedef struct AbrDataFields {
  float* xpos; float* xquat; float* xipos; float* xanchor; float* xaxis; float* cinert; float* cdof; float* cvel;
  float* cdof_dot; float* qfrc_smooth; float* qacc_smooth; float* qfrc_constraint; float* efc_force; float* efc_D;
  float* efc_aref; float* contact_dist; float* contact_pos; float* contact_frame;
} AbrDataFields;
/* ABR_STRUCT_END */
/* abr_forward_dev / abr_env_step_dev (no reset prologue) that also write the requested derived fields of E worlds in the same
 * launch. After a step they are the fields of the LAST forward pass, i.e. of the state before the final integration: exactly
 * what mjx.step leaves in mjx.Data. Served by the generic kernels (the register-resident limb kernels keep no body arrays). */
int abr_forward_fields_dev(AbrModel* m, float* qpos, float* qvel, const float* ctrl, float* qacc_warmstart, float* qacc,
                           int E, const AbrDataFields* fields, void* stream);
int abr_env_step_fields_dev(AbrModel* m, float* qpos, float* qvel, float* qacc_warmstart, float* time, const float* ctrl,
                            int E, int nsubsteps, const AbrDataFields* fields, void* stream);

/* ---- per-env domain randomisation of model parameters (ambersim/trajopt/base.py:49-58 rationale) ----
 * dr: DEVICE pointer [E,2] = {contact friction scale, actuator strength scale} per env, or NULL to switch it off.
 * The friction scale multiplies both tangential coefficients of every contact of the env (and rescales the
 * pyramid rows' invweight accordingly, as if pair_friction had been scaled before mjx.device_put); the actuator
 * scale multiplies gainprm and biasprm (the actuator force before its forcerange clamp). The pointer is kept,
 * not copied (caller-owned, must outlive the env calls); abr_forward_dev / abr_env_step_dev /
 * abr_env_task_step_dev then require the same E. */
int abr_env_set_randomization(AbrModel* m, const float* dr, int E);
/* The same with nparam = 2 or 4 scales per env: dr [E,nparam] = {contact friction, actuator strength, joint damping, joint armature}.
 * The damping scale multiplies dof_damping (passive force and the implicit-damping Euler term), the armature scale dof_armature
 * (the diagonal of the joint-space inertia); like `model.replace(dof_damping=..., dof_armature=...)` in MJX they leave the
 * mj_setConst constants (invweight0, meaninertia) of the base model untouched. */
int abr_env_set_randomization_ex(AbrModel* m, const float* dr, int E, int nparam);

/* ---- one whole training-env step in ONE launch: MjxEnv.step of a quadratic
 * tracking task wrapped in brax's EpisodeWrapper + AutoResetWrapper, E envs in place, DEVICE pointers.
 *   physics : nsubsteps x mjx.step with ctrl held (rl/base.py:88-96)
 *   reward  : -(0.5 (x-xg)' Q (x-xg) + 0.5 u' R u) on the stepped state x = (qpos, qvel) and the action as
 *             given (`reward`: an AbrCost with DIAGONAL Q and R; Qf is not used)
 *   done    : qpos[2] < z_min (floating-base height; pass -INFINITY to disable; a non-finite height also
 *             terminates) or, when max_steps > 0, steps[e] + 1 >= max_steps (EpisodeWrapper's episode_length;
 *             truncation[e] = 1 when only the length ended the episode)
 *   reset   : AutoResetWrapper: where(done, first_state, state) on qpos / qvel / qacc_warmstart / time and obs;
 *             reward and done keep the finished step's values; an env whose done[e] is set on entry starts
 *             counting from 0 again (done is in/out: clear it before the first step)
 *   obs     : nullable [E,nq+nv] = (qpos, qvel) of the state the env leaves the launch in
 * steps [E] int in/out episode counters; reward_out [E]; done [E] bytes; truncation nullable [E] bytes. */
int abr_env_task_step_dev(AbrModel* m, float* qpos, float* qvel, float* qacc_warmstart, float* time,
                          const float* ctrl, int E, int nsubsteps, const float* first_qpos,
                          const float* first_qvel, const float* first_qacc_warmstart,
                          const AbrCost* reward, float z_min, int max_steps, int* steps, float* obs,
                          float* reward_out, unsigned char* done, unsigned char* truncation,
                          void* stream);

/* ---- stage dump for parity tests (tests only): one world, mjx.forward, HOST pointers ------
 * name in {"xpos","xquat","xipos","ximat","subtree_com","cinert","cdof","qM","cvel","cdof_dot",
 * "qfrc_bias","qfrc_passive","qfrc_actuator","qfrc_smooth","qacc_smooth","efc_J","efc_D",
 * "efc_aref","qacc","qfrc_constraint","efc_force"}; out has capacity `cap` floats; *n = count. */
int abr_debug_forward_host(AbrModel* m, const float* qpos, const float* qvel, const float* ctrl,
                           const float* qacc_warmstart, const char* name, float* out, int cap,
                           int* n);

/* ---- receding-horizon loop on the device (how `optimize` is driven in practice) --------
 * nticks x { predictive-sampling solve (as abr_predictive_sample_dev with B = 1, seed + tick) from the
 * current state x; the plant takes ONE physics step under us*[0] (= row 1 of the winner's trajectory);
 * the next guess is us* shifted by one step with its last control repeated }. No host round trip
 * between ticks: every launch is stream-ordered.
 * x [nx] in/out, us_guess [N,nu] in/out; logs (nullable): xs_log [nticks+1,nx] (row 0 = initial x),
 * us_log [nticks,nu] applied controls, cost_log [nticks] winner costs, idx_log [nticks] winner samples. */
int abr_mpc_dev(AbrModel* m, const AbrCost* cost, float* x, float* us_guess, unsigned long long seed,
                int S, int N, float stdev, int nticks, float* xs_log, float* us_log, float* cost_log,
                int* idx_log, void* stream);

/* ---- the sampler's cross-GPU exchange over NVLink peer memory: one process per GPU of ONE box ----
 * Each rank creates an exchange buffer (device memory it owns) and gets a CUDA IPC handle for it; the host
 * framework hands every rank all R handles (e.g. torch.distributed.all_gather_object) for abr_xchg_connect.
 * abr_xchg_merge_best_dev is then collective and stream-ordered: every rank passes its local winners
 * {best_cost [B], best_idx [B] (global sample ids), xs_star [B,nxs], us_star [B,nus]} (DEVICE pointers) and ONE
 * launch stores them into all peers' buffers (P2P stores + system-scope release flags), waits for the R records of
 * this call and writes the global first minimum (NaN minimal, lowest sample id wins ties) to the outputs: identical
 * on every rank, no NCCL call, no host round trip. max_record_floats bounds B * (2 + nxs + nus).
 * A peer that fails to arrive within about 2 s is reported, not waited for: that call's outputs are a sentinel
 * (cost +inf, idx -1, zeros), abr_xchg_timed_out (which synchronises the device) returns 1 once and clears the
 * status word, and the handle refuses further exchanges (ABR_EINVAL): the epoch-parity double buffering no
 * longer protects the records, so every rank destroys and re-creates its exchange.
 * abr_xchg_connect returns ABR_EINVAL when two ranks sit on the same GPU (the kernel spin-waits across ranks).
 * None of these calls changes the caller's current CUDA device. */
#define ABR_XCHG_HANDLE_BYTES 80 /* cudaIpcMemHandle_t (64) + the GPU's UUID (16): abr_xchg_connect refuses two ranks on one GPU */
typedef struct AbrXchg AbrXchg;
int abr_xchg_create(int device, int nranks, int rank, size_t max_record_floats, AbrXchg** out,
                    unsigned char* handle_out /* [ABR_XCHG_HANDLE_BYTES] */);
int abr_xchg_connect(AbrXchg* x, const unsigned char* handles /* [nranks][ABR_XCHG_HANDLE_BYTES] */);
int abr_xchg_merge_best_dev(AbrXchg* x, const float* best_cost, const int* best_idx, const float* xs_star,
                            const float* us_star, int B, int nxs, int nus, float* xs_out, float* us_out,
                            int* idx_out, float* cost_out, void* stream);
int abr_xchg_timed_out(AbrXchg* x, int* timed_out);
int abr_xchg_destroy(AbrXchg* x);

/* FP32 FMA-pipe peak microbenchmark (roofline denominator): returns TFLOP/s */
int abr_ffma_peak(int device, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* ABR_H_ */
