// abr_engine.cu — kernels + C ABI (include/abr.h) of the B200 batched rollout engine.
//
// One group of G lanes per world; a CTA stages the model blob into shared memory once and then
// each group runs its world's whole horizon (shoot, shooting.py:22-48) out of shared memory:
// HBM traffic is controls in and states/costs out only.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "abr.h"
#include "abr_layout.h"
#include "abr_kernels.cuh"
#include "abr_limb.cuh"
#include "abr_hand.cuh"

using namespace abr;

// ================================================================================ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
namespace abr { int set_error(int code, const std::string& msg) { return fail(code, msg); } }  // for the other translation units
#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(ABR_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

// ================================================================================ handles
struct HostModel {  // deep copy of AbrModelHost
  int nq, nv, nu, na, nbody, njnt, ngeom, neq, npair;
  AbrOpt opt;
  std::vector<int> body_parentid, body_rootid, body_jntnum, body_jntadr, body_dofnum, body_dofadr;
  std::vector<float> body_pos, body_quat, body_ipos, body_iquat, body_mass, body_subtreemass, body_inertia, body_invweight0;
  std::vector<int> jnt_type, jnt_qposadr, jnt_dofadr, jnt_bodyid, jnt_limited;
  std::vector<float> jnt_solref, jnt_solimp, jnt_pos, jnt_axis, jnt_stiffness, jnt_range, jnt_margin;
  std::vector<int> dof_bodyid, dof_jntid, dof_parentid;
  std::vector<float> dof_armature, dof_damping, dof_invweight0;
  std::vector<int> geom_type, geom_bodyid;
  std::vector<float> geom_size, geom_pos, geom_quat;
  std::vector<int> geom_vertadr, geom_vertnum;  // convex vertex sets of box / mesh geoms
  std::vector<float> vert;
  int nface = 0, nfacevert = 0, nedge = 0;  // hull topology of the convex geoms
  std::vector<int> geom_faceadr, geom_facenum, face_vertadr, face_vertnum, face_vert, geom_edgeadr, geom_edgenum, edge_vert;
  std::vector<float> face_normal;
  std::vector<int> pair_geom1, pair_geom2, pair_kind, pair_condim;
  std::vector<float> pair_friction, pair_solref, pair_solimp, pair_includemargin;
  std::vector<int> eq_type, eq_obj1id, eq_obj2id, eq_active;
  std::vector<float> eq_solref, eq_solimp, eq_data;
  std::vector<int> actuator_trnid, actuator_gaintype, actuator_biastype, actuator_ctrllimited, actuator_forcelimited;
  std::vector<float> actuator_ctrlrange, actuator_forcerange, actuator_gainprm, actuator_biasprm, actuator_gear;
  std::vector<float> qpos0, qpos_spring;
};

// The entry points run on the model's device and leave the caller's current device as they found it (a single-process
// multi-GPU host framework keeps its own notion of "current").
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); err = cudaSetDevice(dev); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ABR_ON_DEVICE(dev)                                                                                                   \
  DeviceGuard device_guard_(dev);                                                                                           \
  if (device_guard_.err != cudaSuccess) return fail(ABR_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(device_guard_.err))

struct Scratch {
  void* p = nullptr;
  size_t cap = 0;
  bool external = false;  // carved from a caller-provided workspace (abr_model_set_workspace): never grown, never freed here
  int ensure(size_t bytes) {
    if (bytes <= cap) return ABR_OK;
    if (external) return fail(ABR_ECAPACITY, "the caller-provided workspace is too small for this call (abr_workspace_bytes sizes it)");
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(ABR_ECUDA, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
    cap = bytes;
    return ABR_OK;
  }
  void release() { if (p && !external) cudaFree(p); p = nullptr; cap = 0; external = false; }
  void adopt(void* ptr, size_t bytes) { release(); p = ptr; cap = bytes; external = true; }
};

struct AbrModel {
  int device = 0;
  HostModel hm;
  Layout lay;        // production layout (aliased per-world regions)
  Layout lay_dbg;    // debug layout (no aliasing: every intermediate survives the step)
  std::vector<float> mf;
  std::vector<int> mi;
  float* d_blob = nullptr;
  int lanes = 0;     // 0 = auto
  int num_sms = 148;
  int max_smem = 0;
  cudaStream_t stream = nullptr;  // for the *_host entry points
  cudaStream_t copy_stream = nullptr;  // host->device slices of a pipelined abr_rollout_host
  const float* dr = nullptr; int dr_E = 0, dr_n = 2;  // abr_env_set_randomization(_ex)
  std::vector<cudaEvent_t> ev;
  Scratch s_costs, s_in, s_out, s_dbg, s_traj, s_carry;
  Scratch s_mpc;  // abr_mpc_dev's winner buffers: its own, so that a *_host call on the handle's stream never shares scratch with it
};

struct AbrCost {
  int device = 0;
  int nx = 0, nu = 0;
  int diag = 0;      // Q, Qf and R are all diagonal
  float* d = nullptr;  // [Q nx*nx][Qf nx*nx][R nu*nu][xg nx][qd nx][qfd nx][rd nu]
};

// ================================================================================ blob build
template <class T> static std::vector<T> vcopy(const T* p, size_t n) { return p ? std::vector<T>(p, p + n) : std::vector<T>(n, T(0)); }

static void copy_host_model(const AbrModelHost* h, HostModel& m) {
  m.nq = h->nq; m.nv = h->nv; m.nu = h->nu; m.na = h->na; m.nbody = h->nbody; m.njnt = h->njnt;
  m.ngeom = h->ngeom; m.neq = h->neq; m.npair = h->npair; m.opt = h->opt;
  const int nb = m.nbody, nj = m.njnt, nv = m.nv, ng = m.ngeom, np = m.npair, ne = m.neq, nu = m.nu;
  m.body_parentid = vcopy(h->body_parentid, nb); m.body_rootid = vcopy(h->body_rootid, nb);
  m.body_jntnum = vcopy(h->body_jntnum, nb); m.body_jntadr = vcopy(h->body_jntadr, nb);
  m.body_dofnum = vcopy(h->body_dofnum, nb); m.body_dofadr = vcopy(h->body_dofadr, nb);
  m.body_pos = vcopy(h->body_pos, 3 * nb); m.body_quat = vcopy(h->body_quat, 4 * nb);
  m.body_ipos = vcopy(h->body_ipos, 3 * nb); m.body_iquat = vcopy(h->body_iquat, 4 * nb);
  m.body_mass = vcopy(h->body_mass, nb); m.body_subtreemass = vcopy(h->body_subtreemass, nb);
  m.body_inertia = vcopy(h->body_inertia, 3 * nb); m.body_invweight0 = vcopy(h->body_invweight0, 2 * nb);
  m.jnt_type = vcopy(h->jnt_type, nj); m.jnt_qposadr = vcopy(h->jnt_qposadr, nj);
  m.jnt_dofadr = vcopy(h->jnt_dofadr, nj); m.jnt_bodyid = vcopy(h->jnt_bodyid, nj);
  m.jnt_limited = vcopy(h->jnt_limited, nj);
  m.jnt_solref = vcopy(h->jnt_solref, 2 * nj); m.jnt_solimp = vcopy(h->jnt_solimp, 5 * nj);
  m.jnt_pos = vcopy(h->jnt_pos, 3 * nj); m.jnt_axis = vcopy(h->jnt_axis, 3 * nj);
  m.jnt_stiffness = vcopy(h->jnt_stiffness, nj); m.jnt_range = vcopy(h->jnt_range, 2 * nj);
  m.jnt_margin = vcopy(h->jnt_margin, nj);
  m.dof_bodyid = vcopy(h->dof_bodyid, nv); m.dof_jntid = vcopy(h->dof_jntid, nv); m.dof_parentid = vcopy(h->dof_parentid, nv);
  m.dof_armature = vcopy(h->dof_armature, nv); m.dof_damping = vcopy(h->dof_damping, nv);
  m.dof_invweight0 = vcopy(h->dof_invweight0, nv);
  m.geom_type = vcopy(h->geom_type, ng); m.geom_bodyid = vcopy(h->geom_bodyid, ng);
  m.geom_size = vcopy(h->geom_size, 3 * ng); m.geom_pos = vcopy(h->geom_pos, 3 * ng); m.geom_quat = vcopy(h->geom_quat, 4 * ng);
  m.geom_vertadr = vcopy(h->geom_vertadr, ng); m.geom_vertnum = vcopy(h->geom_vertnum, ng); m.vert = vcopy(h->vert, 3 * h->nvert);
  m.nface = h->nface; m.nfacevert = h->nfacevert; m.nedge = h->nedge;
  m.geom_faceadr = vcopy(h->geom_faceadr, ng); m.geom_facenum = vcopy(h->geom_facenum, ng);
  m.face_vertadr = vcopy(h->face_vertadr, h->nface); m.face_vertnum = vcopy(h->face_vertnum, h->nface);
  m.face_vert = vcopy(h->face_vert, h->nfacevert); m.face_normal = vcopy(h->face_normal, 3 * h->nface);
  m.geom_edgeadr = vcopy(h->geom_edgeadr, ng); m.geom_edgenum = vcopy(h->geom_edgenum, ng); m.edge_vert = vcopy(h->edge_vert, 2 * h->nedge);
  m.pair_geom1 = vcopy(h->pair_geom1, np); m.pair_geom2 = vcopy(h->pair_geom2, np);
  m.pair_kind = vcopy(h->pair_kind, np); m.pair_condim = vcopy(h->pair_condim, np);
  m.pair_friction = vcopy(h->pair_friction, 5 * np); m.pair_solref = vcopy(h->pair_solref, 2 * np);
  m.pair_solimp = vcopy(h->pair_solimp, 5 * np); m.pair_includemargin = vcopy(h->pair_includemargin, np);
  m.eq_type = vcopy(h->eq_type, ne); m.eq_obj1id = vcopy(h->eq_obj1id, ne); m.eq_obj2id = vcopy(h->eq_obj2id, ne);
  m.eq_active = vcopy(h->eq_active, ne);
  m.eq_solref = vcopy(h->eq_solref, 2 * ne); m.eq_solimp = vcopy(h->eq_solimp, 5 * ne); m.eq_data = vcopy(h->eq_data, 11 * ne);
  m.actuator_trnid = vcopy(h->actuator_trnid, nu); m.actuator_gaintype = vcopy(h->actuator_gaintype, nu);
  m.actuator_biastype = vcopy(h->actuator_biastype, nu); m.actuator_ctrllimited = vcopy(h->actuator_ctrllimited, nu);
  m.actuator_forcelimited = vcopy(h->actuator_forcelimited, nu);
  m.actuator_ctrlrange = vcopy(h->actuator_ctrlrange, 2 * nu); m.actuator_forcerange = vcopy(h->actuator_forcerange, 2 * nu);
  m.actuator_gainprm = vcopy(h->actuator_gainprm, 3 * nu); m.actuator_biasprm = vcopy(h->actuator_biasprm, 3 * nu);
  m.actuator_gear = vcopy(h->actuator_gear, nu);
  m.qpos0 = vcopy(h->qpos0, m.nq); m.qpos_spring = vcopy(h->qpos_spring, m.nq);
}

struct Pool {
  std::vector<float> f;
  std::vector<int> i;
  int addf(const std::vector<float>& v) { int o = (int)f.size(); f.insert(f.end(), v.begin(), v.end()); return o; }
  int addi(const std::vector<int>& v) { int o = (int)i.size(); i.insert(i.end(), v.begin(), v.end()); return o; }
};

// constraint._kbi constants of one row source (solref, solimp) under the current options
static void row_prm(const AbrOpt& opt, const float* solref, const float* solimp, float invweight, float* out) {
  double timeconst = solref[0], dampratio = solref[1];
  if (!(opt.disableflags & ABR_DSBL_REFSAFE)) timeconst = std::max(timeconst, 2.0 * (double)opt.timestep);
  auto clip = [](double x, double lo, double hi) { return std::min(std::max(x, lo), hi); };
  double dmin = clip(solimp[0], 1e-4, 0.9999), dmax = clip(solimp[1], 1e-4, 0.9999);
  double width = std::max(1e-15, (double)solimp[2]);
  double mid = clip(solimp[3], 1e-4, 0.9999), power = std::max(1.0, (double)solimp[4]);
  double k = 1.0 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  double b = 2.0 / (dmax * timeconst);
  if (solref[0] <= 0) k = -(double)solref[0] / (dmax * dmax);
  if (solref[1] <= 0) b = -(double)solref[1] / dmax;
  out[0] = (float)k; out[1] = (float)b; out[2] = (float)dmin; out[3] = (float)dmax;
  out[4] = (float)(1.0 / width); out[5] = (float)mid; out[6] = (float)power; out[7] = invweight;
  out[8] = (float)(1.0 / std::pow(mid, power - 1.0));
  out[9] = (float)(1.0 / std::pow(1.0 - mid, power - 1.0));
}

static int pair_ncon(int kind) {
  if (kind == ABR_PAIR_PLANE_CONVEX || kind == ABR_PAIR_CONVEX_CONVEX) return 4;
  return (kind == ABR_PAIR_PLANE_CAPSULE || kind == ABR_PAIR_CAPSULE_CONVEX) ? 2 : 1;
}

static int build_blob(const HostModel& m, bool alias, Layout& L, std::vector<float>& mf, std::vector<int>& mi) {
  memset(&L, 0, sizeof(L));
  const AbrOpt& opt = m.opt;
  const int nb = m.nbody, nj = m.njnt, nv = m.nv, nu = m.nu, nq = m.nq;
  // ---- validation (mirrors MJX device_put's NotImplementedError, io_utils.py:228-241)
  if (opt.cone != 0) return fail(ABR_EUNSUPPORTED, "elliptic friction cones are not supported (pyramidal only)");
  if (opt.integrator != ABR_INT_EULER && opt.integrator != ABR_INT_RK4)
    return fail(ABR_EUNSUPPORTED, "integrator must be Euler (0) or RK4 (1)");
  if (opt.solver != ABR_SOLVER_NEWTON && opt.solver != ABR_SOLVER_CG)
    return fail(ABR_EUNSUPPORTED, "solver must be CG (1) or Newton (2)");
  if (nv > 64 || nb > 4096) return fail(ABR_ECAPACITY, "model too large for the compiled kernels (nv <= 64)");
  if (m.na != 0) return fail(ABR_EUNSUPPORTED, "stateful actuators are not supported");
  for (int j = 0; j < nj; j++)
    if (m.jnt_type[j] == ABR_JNT_BALL) return fail(ABR_EUNSUPPORTED, "ball joints are not supported");
  for (int e = 0; e < m.neq; e++)
    if (m.eq_active[e] && m.eq_type[e] != ABR_EQ_JOINT) return fail(ABR_EUNSUPPORTED, "only joint equalities are supported");
  for (int b = 1; b < nb; b++)
    if (m.body_parentid[b] >= b) return fail(ABR_EINVAL, "bodies must be ordered parent before child");

  const int dis = opt.disableflags;
  const bool con_all = !(dis & ABR_DSBL_CONSTRAINT);
  Pool P;
  L.nq = nq; L.nv = nv; L.nu = nu; L.nbody = nb; L.njnt = nj; L.ngeom = m.ngeom; L.npair = m.npair; L.nx = nq + nv;
  L.integrator = opt.integrator; L.solver = opt.solver; L.iterations = opt.iterations;
  L.ls_iterations = opt.ls_iterations; L.disableflags = dis;
  L.timestep = opt.timestep; L.tolerance = opt.tolerance; L.ls_tolerance = opt.ls_tolerance;
  L.meaninertia = opt.meaninertia; L.impratio = opt.impratio;
  for (int i = 0; i < 3; i++) L.gravity[i] = opt.gravity[i];

  // ---- float pool: plain model arrays
  L.f_body_pos = P.addf(m.body_pos); L.f_body_quat = P.addf(m.body_quat); L.f_body_ipos = P.addf(m.body_ipos);
  L.f_body_iquat = P.addf(m.body_iquat); L.f_body_mass = P.addf(m.body_mass); L.f_body_inertia = P.addf(m.body_inertia);
  L.f_jnt_pos = P.addf(m.jnt_pos); L.f_jnt_axis = P.addf(m.jnt_axis); L.f_jnt_range = P.addf(m.jnt_range);
  L.f_jnt_margin = P.addf(m.jnt_margin); L.f_jnt_stiffness = P.addf(m.jnt_stiffness);
  L.f_dof_armature = P.addf(m.dof_armature); L.f_dof_damping = P.addf(m.dof_damping);
  L.f_qpos0 = P.addf(m.qpos0); L.f_qpos_spring = P.addf(m.qpos_spring);
  L.f_geom_size = P.addf(m.geom_size); L.f_geom_pos = P.addf(m.geom_pos); L.f_geom_quat = P.addf(m.geom_quat);
  L.f_vert = P.addf(m.vert);  // mesh / box vertices live in the model blob (geom frame)
  L.f_face_normal = P.addf(m.face_normal);

  // ---- tree tables
  std::vector<int> depth(nb, 0), rootslot(nb, 0), roots;
  int maxdepth = 0;
  for (int b = 1; b < nb; b++) {
    depth[b] = depth[m.body_parentid[b]] + 1;
    maxdepth = std::max(maxdepth, depth[b]);
    if (m.body_parentid[b] == 0) { rootslot[b] = (int)roots.size(); roots.push_back(b); }
    else rootslot[b] = rootslot[m.body_parentid[b]];
  }
  L.depth = maxdepth; L.nroot = (int)roots.size();
  std::vector<int> level_adr(maxdepth + 2, 0), level_body;
  for (int lev = 0; lev <= maxdepth; lev++) {
    level_adr[lev] = (int)level_body.size();
    for (int b = 0; b < nb; b++) if (depth[b] == lev) level_body.push_back(b);
  }
  level_adr[maxdepth + 1] = (int)level_body.size();
  std::vector<int> childadr(nb, 0), childnum(nb, 0), child;
  for (int p = 0; p < nb; p++) {
    childadr[p] = (int)child.size();
    for (int b = 1; b < nb; b++) if (m.body_parentid[b] == p && b != p) { child.push_back(b); childnum[p]++; }
  }
  L.i_body_parent = P.addi(m.body_parentid); L.i_body_jntadr = P.addi(m.body_jntadr); L.i_body_jntnum = P.addi(m.body_jntnum);
  L.i_body_dofadr = P.addi(m.body_dofadr); L.i_body_dofnum = P.addi(m.body_dofnum); L.i_body_rootslot = P.addi(rootslot);
  L.i_body_childadr = P.addi(childadr); L.i_body_childnum = P.addi(childnum); L.i_child = P.addi(child);
  L.i_level_adr = P.addi(level_adr); L.i_level_body = P.addi(level_body);
  L.i_jnt_type = P.addi(m.jnt_type); L.i_jnt_qposadr = P.addi(m.jnt_qposadr); L.i_jnt_dofadr = P.addi(m.jnt_dofadr);
  L.i_jnt_body = P.addi(m.jnt_bodyid);
  L.i_dof_body = P.addi(m.dof_bodyid); L.i_dof_jnt = P.addi(m.dof_jntid);
  L.i_root_body = P.addi(roots);
  L.i_geom_body = P.addi(m.geom_bodyid);
  L.i_geom_vertadr = P.addi(m.geom_vertadr); L.i_geom_vertnum = P.addi(m.geom_vertnum);
  L.i_geom_faceadr = P.addi(m.geom_faceadr); L.i_geom_facenum = P.addi(m.geom_facenum);
  L.i_face_vertadr = P.addi(m.face_vertadr); L.i_face_vertnum = P.addi(m.face_vertnum); L.i_face_vert = P.addi(m.face_vert);
  L.i_geom_edgeadr = P.addi(m.geom_edgeadr); L.i_geom_edgenum = P.addi(m.geom_edgenum); L.i_edge_vert = P.addi(m.edge_vert);
  L.i_pair_g1 = P.addi(m.pair_geom1); L.i_pair_g2 = P.addi(m.pair_geom2); L.i_pair_kind = P.addi(m.pair_kind);

  // M sparsity (ancestor pairs) and packed-index table
  std::vector<int> mpair, tri_tab;
  for (int i = 0; i < nv; i++)
    for (int j = i; j >= 0; j = m.dof_parentid[j]) mpair.push_back((i << 16) | j);
  for (int i = 0; i < nv; i++)
    for (int j = 0; j <= i; j++) tri_tab.push_back((i << 16) | j);
  L.nmpair = (int)mpair.size(); L.ntri = nv * (nv + 1) / 2;
  L.i_mpair = P.addi(mpair); L.i_tri = P.addi(tri_tab);

  // ---- constraint rows: equality, limit, contact (static sizes)
  std::vector<int> eq_j1, eq_j2, lim_jnt, dof_limrow(nv, -1), row_info;
  std::vector<float> eq_prm, eq_data, lim_prm, con_prm;
  if (con_all && !(dis & ABR_DSBL_EQUALITY)) {
    for (int e = 0; e < m.neq; e++) {
      if (!m.eq_active[e]) continue;
      int j1 = m.eq_obj1id[e], j2 = m.eq_obj2id[e];
      for (int j : {j1, j2})
        if (j >= 0 && m.jnt_type[j] != ABR_JNT_HINGE && m.jnt_type[j] != ABR_JNT_SLIDE)
          return fail(ABR_EUNSUPPORTED, "joint equality on a non-scalar joint");
      float inv = m.dof_invweight0[m.jnt_dofadr[j1]] + (j2 >= 0 ? m.dof_invweight0[m.jnt_dofadr[j2]] : 0.f);
      float prm[kRowPrm];
      row_prm(opt, &m.eq_solref[2 * e], &m.eq_solimp[5 * e], inv, prm);
      row_info.push_back(0 | ((int)eq_j1.size() << 2));
      eq_j1.push_back(j1); eq_j2.push_back(j2);
      eq_prm.insert(eq_prm.end(), prm, prm + kRowPrm);
      eq_data.insert(eq_data.end(), &m.eq_data[11 * e], &m.eq_data[11 * e] + 5);
    }
  }
  L.ne = (int)eq_j1.size();
  if (con_all && !(dis & ABR_DSBL_LIMIT)) {
    for (int j = 0; j < nj; j++) {
      if (!m.jnt_limited[j]) continue;
      if (m.jnt_type[j] != ABR_JNT_HINGE && m.jnt_type[j] != ABR_JNT_SLIDE) continue;
      float prm[kRowPrm];
      row_prm(opt, &m.jnt_solref[2 * j], &m.jnt_solimp[5 * j], m.dof_invweight0[m.jnt_dofadr[j]], prm);
      dof_limrow[m.jnt_dofadr[j]] = (int)lim_jnt.size();
      row_info.push_back(1 | ((int)lim_jnt.size() << 2));
      lim_jnt.push_back(j);
      lim_prm.insert(lim_prm.end(), prm, prm + kRowPrm);
    }
  }
  L.nl = (int)lim_jnt.size();
  std::vector<int> con_pair, con_sub, con_row, con_condim, con_dofmask;
  // contact parameter blocks are per pair (allocated for every pair so indices stay pair ids)
  for (int p = 0; p < m.npair; p++) {
    int b1 = m.geom_bodyid[m.pair_geom1[p]], b2 = m.geom_bodyid[m.pair_geom2[p]];
    float t = m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2];
    float prm[kConPrm];
    float mu1 = m.pair_friction[5 * p], mu2 = m.pair_friction[5 * p + 1];
    row_prm(opt, &m.pair_solref[2 * p], &m.pair_solimp[5 * p], t, prm);
    if (m.pair_condim[p] == 3) {
      prm[7] = (t + mu1 * mu1 * t) * 2.f * mu1 * mu1 / opt.impratio;
      prm[10] = (t + mu2 * mu2 * t) * 2.f * mu2 * mu2 / opt.impratio;
    } else {
      prm[10] = t;
    }
    prm[11] = mu1; prm[12] = mu2; prm[13] = m.pair_includemargin[p];
    con_prm.insert(con_prm.end(), prm, prm + kConPrm);
  }
  if (con_all && !(dis & ABR_DSBL_CONTACT)) {
    for (int p = 0; p < m.npair; p++) {
      if (m.pair_condim[p] != 1 && m.pair_condim[p] != 3) return fail(ABR_EUNSUPPORTED, "contact condim must be 1 or 3");
      int b1 = m.geom_bodyid[m.pair_geom1[p]], b2 = m.geom_bodyid[m.pair_geom2[p]];
      std::vector<char> anc1(nb, 0), anc2(nb, 0);
      for (int b = b1; b > 0; b = m.body_parentid[b]) anc1[b] = 1;
      for (int b = b2; b > 0; b = m.body_parentid[b]) anc2[b] = 1;
      for (int s = 0; s < pair_ncon(m.pair_kind[p]); s++) {
        int ci = (int)con_pair.size();
        con_pair.push_back(p); con_sub.push_back(s); con_condim.push_back(m.pair_condim[p]);
        con_row.push_back((int)row_info.size());
        int nrow = m.pair_condim[p] == 1 ? 1 : 4;
        for (int r = 0; r < nrow; r++) row_info.push_back(2 | (ci << 2) | (r << 20));
        for (int d = 0; d < nv; d++) {
          int bd = m.dof_bodyid[d];
          int mk = (anc2[bd] ? 1 : 0) | (anc1[bd] ? 2 : 0);
          con_dofmask.push_back(mk == 3 ? 0 : mk);  // common ancestors cancel exactly
        }
      }
    }
  }
  L.ncon = (int)con_pair.size();
  L.nefc = (int)row_info.size();
  L.f_eq_prm = P.addf(eq_prm); L.f_eq_data = P.addf(eq_data); L.f_lim_prm = P.addf(lim_prm); L.f_con_prm = P.addf(con_prm);
  L.i_dof_limrow = P.addi(dof_limrow); L.i_lim_jnt = P.addi(lim_jnt); L.i_eq_j1 = P.addi(eq_j1); L.i_eq_j2 = P.addi(eq_j2);
  L.i_con_pair = P.addi(con_pair); L.i_con_sub = P.addi(con_sub); L.i_con_row = P.addi(con_row);
  L.i_con_condim = P.addi(con_condim); L.i_con_dofmask = P.addi(con_dofmask); L.i_row_info = P.addi(row_info);

  // ---- actuators
  std::vector<float> act_prm;
  std::vector<int> act_jnt, act_flags, dof_actadr(nv, 0), dof_actnum(nv, 0), dof_act;
  for (int u = 0; u < nu; u++) {
    int j = m.actuator_trnid[u];
    if (j < 0 || j >= nj || (m.jnt_type[j] != ABR_JNT_HINGE && m.jnt_type[j] != ABR_JNT_SLIDE))
      return fail(ABR_EUNSUPPORTED, "actuators must drive hinge or slide joints");
    float prm[kActPrm] = {m.actuator_ctrlrange[2 * u], m.actuator_ctrlrange[2 * u + 1], m.actuator_forcerange[2 * u],
                          m.actuator_forcerange[2 * u + 1], m.actuator_gainprm[3 * u], m.actuator_gainprm[3 * u + 1],
                          m.actuator_gainprm[3 * u + 2], m.actuator_biasprm[3 * u], m.actuator_biasprm[3 * u + 1],
                          m.actuator_biasprm[3 * u + 2], m.actuator_gear[u]};
    act_prm.insert(act_prm.end(), prm, prm + kActPrm);
    act_jnt.push_back(j);
    act_flags.push_back((m.actuator_ctrllimited[u] ? 1 : 0) | (m.actuator_forcelimited[u] ? 2 : 0) |
                        (m.actuator_gaintype[u] == ABR_GAIN_AFFINE ? 4 : 0) | (m.actuator_biastype[u] == ABR_BIAS_AFFINE ? 8 : 0));
  }
  for (int d = 0; d < nv; d++) {
    dof_actadr[d] = (int)dof_act.size();
    for (int u = 0; u < nu; u++)
      if (m.jnt_dofadr[act_jnt[u]] == d) { dof_act.push_back(u); dof_actnum[d]++; }
  }
  L.f_act_prm = P.addf(act_prm);
  L.i_act_jnt = P.addi(act_jnt); L.i_act_flags = P.addi(act_flags);
  L.i_dof_actadr = P.addi(dof_actadr); L.i_dof_actnum = P.addi(dof_actnum); L.i_dof_act = P.addi(dof_act);

  // ---- tree-sparse linear-algebra tables (see abr_layout.h)
  {
    const std::vector<int>& par = m.dof_parentid;
    std::vector<int> sp_adr(nv + 1, 0);
    for (int i = 0, e = 0; i <= nv; i++) {
      sp_adr[i] = e;
      if (i < nv) for (int j = i; j >= 0; j = par[j]) e++;
    }
    auto is_anc = [&](int a, int d) { for (int j = d; j >= 0; j = par[j]) if (j == a) return true; return false; };
    auto ent = [&](int i, int j) { int e = sp_adr[i]; for (int q = i; q >= 0; q = par[q], e++) if (q == j) return e; return -1; };
    bool ok = getenv("ABR_DENSE") == nullptr;
    for (size_t e = 0; e < eq_j1.size() && ok; e++) {
      if (eq_j2[e] < 0) continue;
      int d1 = m.jnt_dofadr[eq_j1[e]], d2 = m.jnt_dofadr[eq_j2[e]];
      if (!is_anc(d1, d2) && !is_anc(d2, d1)) ok = false;
    }
    std::vector<int> cd_adr(L.ncon + 1, 0), cd_dof;
    for (int ci = 0; ci < L.ncon; ci++) {
      cd_adr[ci] = (int)cd_dof.size();
      for (int d = 0; d < nv; d++) if (con_dofmask[ci * nv + d]) cd_dof.push_back((ci << 16) | d);
      for (int a = cd_adr[ci]; a < (int)cd_dof.size() && ok; a++)
        for (int b = a + 1; b < (int)cd_dof.size(); b++)
          if (!is_anc(cd_dof[a] & 0xffff, cd_dof[b] & 0xffff) && !is_anc(cd_dof[b] & 0xffff, cd_dof[a] & 0xffff)) ok = false;
    }
    cd_adr[L.ncon] = (int)cd_dof.size();
    L.sparse = ok ? 1 : 0;
    L.nnz = sp_adr[nv];
    std::vector<int> height(nv, 0);
    for (int k = nv - 1; k >= 0; k--) if (par[k] >= 0) height[par[k]] = std::max(height[par[k]], height[k] + 1);
    int nstage = 0;
    for (int k = 0; k < nv; k++) nstage = std::max(nstage, height[k] + 1);
    L.nstage = nstage;
    std::vector<int> st_adr(nstage + 1, 0), st_dof;
    std::vector<int> fu_tadr(nstage + 1, 0), fu_tgt, fu_cadr(1, 0), fu_src, fu_k;
    std::vector<int> sb_tadr(nstage + 1, 0), sb_tgt, sb_cadr(1, 0), sb_src;
    for (int st = 0; st < nstage; st++) {
      st_adr[st] = (int)st_dof.size();
      std::vector<std::vector<std::pair<int, int>>> upd(L.nnz);   // target entry -> (src pair, pivot)
      std::vector<std::vector<int>> bs(nv);                       // target dof -> (entry << 16 | k)
      for (int k = 0; k < nv; k++) {
        if (height[k] != st) continue;
        st_dof.push_back(k);
        for (int i = par[k]; i >= 0; i = par[i]) {
          bs[i].push_back((ent(k, i) << 16) | k);
          for (int j = i; j >= 0; j = par[j]) upd[ent(i, j)].push_back({(ent(k, i) << 16) | ent(k, j), k});
        }
      }
      fu_tadr[st] = (int)fu_tgt.size();
      for (int e = 0; e < L.nnz; e++) {
        if (upd[e].empty()) continue;
        fu_tgt.push_back(e);
        for (auto& u : upd[e]) { fu_src.push_back(u.first); fu_k.push_back(u.second); }
        fu_cadr.push_back((int)fu_src.size());
      }
      sb_tadr[st] = (int)sb_tgt.size();
      for (int i = 0; i < nv; i++) {
        if (bs[i].empty()) continue;
        sb_tgt.push_back(i);
        for (int v : bs[i]) sb_src.push_back(v);
        sb_cadr.push_back((int)sb_src.size());
      }
    }
    st_adr[nstage] = (int)st_dof.size(); fu_tadr[nstage] = (int)fu_tgt.size(); sb_tadr[nstage] = (int)sb_tgt.size();
    std::vector<int> mm_adr(nv + 1, 0), mm_src;
    for (int i = 0; i < nv; i++) {
      mm_adr[i] = (int)mm_src.size();
      for (int j = i; j >= 0; j = par[j]) mm_src.push_back((ent(i, j) << 16) | j);
      for (int k = i + 1; k < nv; k++) if (is_anc(i, k)) mm_src.push_back((ent(k, i) << 16) | k);
    }
    mm_adr[nv] = (int)mm_src.size();
    std::vector<int> hc_adr(L.nnz + 1, 0), hc_con, he_adr(L.nnz + 1, 0), he_eq;
    for (int i = 0, e = 0; i < nv; i++)
      for (int j = i; j >= 0; j = par[j], e++) {
        hc_adr[e] = (int)hc_con.size();
        he_adr[e] = (int)he_eq.size();
        for (int ci = 0; ci < L.ncon; ci++) if (con_dofmask[ci * nv + i] && con_dofmask[ci * nv + j]) hc_con.push_back(ci);
        for (size_t q = 0; q < eq_j1.size(); q++) {
          int d1 = m.jnt_dofadr[eq_j1[q]], d2 = eq_j2[q] >= 0 ? m.jnt_dofadr[eq_j2[q]] : -1;
          if (i == d1 && j == d1) he_eq.push_back(((int)q << 2) | 0);
          else if (i == d2 && j == d2) he_eq.push_back(((int)q << 2) | 1);
          else if ((i == d1 && j == d2) || (i == d2 && j == d1)) he_eq.push_back(((int)q << 2) | 2);
        }
      }
    hc_adr[L.nnz] = (int)hc_con.size(); he_adr[L.nnz] = (int)he_eq.size();
    L.i_sp_adr = P.addi(sp_adr); L.i_st_adr = P.addi(st_adr); L.i_st_dof = P.addi(st_dof);
    L.i_fu_tadr = P.addi(fu_tadr); L.i_fu_tgt = P.addi(fu_tgt); L.i_fu_cadr = P.addi(fu_cadr); L.i_fu_src = P.addi(fu_src); L.i_fu_k = P.addi(fu_k);
    L.i_sb_tadr = P.addi(sb_tadr); L.i_sb_tgt = P.addi(sb_tgt); L.i_sb_cadr = P.addi(sb_cadr); L.i_sb_src = P.addi(sb_src);
    L.i_mm_adr = P.addi(mm_adr); L.i_mm_src = P.addi(mm_src);
    L.i_hc_adr = P.addi(hc_adr); L.i_hc_con = P.addi(hc_con); L.i_he_adr = P.addi(he_adr); L.i_he_eq = P.addi(he_eq);
    L.i_cd_adr = P.addi(cd_adr); L.i_cd_dof = P.addi(cd_dof);
  }

  // ---- per-lane table constants shared by the limb (abr_limb.cuh) and hand (abr_hand.cuh) kernels
  using SetF = std::function<void(int, int, float)>;
  using SetI = std::function<void(int, int, int)>;
  auto fold_body = [&](const SetF& setf, int sb, int g, int body, bool has_joint) {
    // constants of the body folded into its PARENT frame (double precision, once): see the kinematics block of abr_limb.cuh
    const double bq[4] = {m.body_quat[4 * body], m.body_quat[4 * body + 1], m.body_quat[4 * body + 2], m.body_quat[4 * body + 3]};
    const double iq[4] = {m.body_iquat[4 * body], m.body_iquat[4 * body + 1], m.body_iquat[4 * body + 2], m.body_iquat[4 * body + 3]};
    auto qmat = [](const double* q, double* R) {
      R[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3]; R[1] = 2 * (q[1] * q[2] - q[0] * q[3]); R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
      R[3] = 2 * (q[1] * q[2] + q[0] * q[3]); R[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3]; R[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
      R[6] = 2 * (q[1] * q[3] - q[0] * q[2]); R[7] = 2 * (q[2] * q[3] + q[0] * q[1]); R[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
    };
    double Rb[9], Ri[9];
    qmat(bq, Rb); qmat(iq, Ri);
    double jp[3] = {0, 0, 0}, jx[3] = {0, 0, 0};
    if (has_joint) {
      const int j = m.body_jntadr[body];
      for (int i = 0; i < 3; i++) { jp[i] = m.jnt_pos[3 * j + i]; jx[i] = m.jnt_axis[3 * j + i]; }
    }
    for (int i = 0; i < 3; i++) {
      setf(sb + i, g, (float)(m.body_pos[3 * body + i] + Rb[3 * i] * jp[0] + Rb[3 * i + 1] * jp[1] + Rb[3 * i + 2] * jp[2]));
      setf(sb + 21 + i, g, (float)(Rb[3 * i] * jx[0] + Rb[3 * i + 1] * jx[1] + Rb[3 * i + 2] * jx[2]));
      setf(sb + 7 + i, g, m.body_ipos[3 * body + i]);
    }
    for (int i = 0; i < 4; i++) setf(sb + 3 + i, g, (float)bq[i]);
    // body_quat o (0, axis)
    setf(sb + 10, g, (float)(-bq[1] * jx[0] - bq[2] * jx[1] - bq[3] * jx[2]));
    setf(sb + 11, g, (float)(bq[0] * jx[0] + bq[2] * jx[2] - bq[3] * jx[1]));
    setf(sb + 12, g, (float)(bq[0] * jx[1] - bq[1] * jx[2] + bq[3] * jx[0]));
    setf(sb + 13, g, (float)(bq[0] * jx[2] + bq[1] * jx[1] - bq[2] * jx[0]));
    setf(sb + 14, g, m.body_mass[body]);
    // inertia tensor in the body frame: Ri diag(inertia) Ri'
    const double in[3] = {m.body_inertia[3 * body], m.body_inertia[3 * body + 1], m.body_inertia[3 * body + 2]};
    const int ra[6] = {0, 1, 2, 0, 0, 1}, cb[6] = {0, 1, 2, 1, 2, 2};
    for (int e = 0; e < 6; e++) {
      double t = 0;
      for (int k = 0; k < 3; k++) t += Ri[3 * ra[e] + k] * in[k] * Ri[3 * cb[e] + k];
      setf(sb + 15 + e, g, (float)t);
    }
  };
  auto fold_joint = [&](const SetF& setf, const SetI& seti, int sj, int si, int g, int body) {
    const int j = m.body_jntadr[body], d = m.jnt_dofadr[j], qa = m.jnt_qposadr[j];
    for (int i = 0; i < 3; i++) { setf(sj + i, g, m.jnt_pos[3 * j + i]); setf(sj + 3 + i, g, m.jnt_axis[3 * j + i]); }
    setf(sj + 6, g, m.qpos0[qa]); setf(sj + 7, g, m.qpos_spring[qa]); setf(sj + 8, g, m.jnt_stiffness[j]);
    setf(sj + 9, g, m.dof_damping[d]); setf(sj + 10, g, m.dof_armature[d]);
    setf(sj + 11, g, m.jnt_range[2 * j]); setf(sj + 12, g, m.jnt_range[2 * j + 1]); setf(sj + 13, g, m.jnt_margin[j]);
    int flags = m.jnt_type[j] == ABR_JNT_HINGE ? limb::kJHinge : limb::kJSlide;
    if (dof_limrow[d] >= 0) {
      flags |= limb::kJLimited;
      for (int i = 0; i < kRowPrm; i++) setf(sj + 14 + i, g, lim_prm[(size_t)kRowPrm * dof_limrow[d] + i]);
    }
    int act = -1;
    if (dof_actnum[d] == 1) {
      act = dof_act[dof_actadr[d]];
      flags |= limb::kJAct | (act_flags[act] << limb::kJActShift);
      for (int i = 0; i < kActPrm; i++) setf(sj + 24 + i, g, act_prm[(size_t)kActPrm * act + i]);
    }
    seti(si, g, flags); seti(si + 1, g, d); seti(si + 2, g, qa); seti(si + 3, g, act);
  };

  // ---- limb (path) decomposition: tables of abr_limb.cuh
  L.limb_ok = 0;
  {
    bool ok = getenv("ABR_NO_LIMB") == nullptr && opt.solver == ABR_SOLVER_NEWTON && opt.integrator == ABR_INT_EULER &&
              L.ne == 0 && nb >= 2 && nv >= 6 && roots.size() == 1 && roots[0] == 1;
    if (ok) {
      const int j0 = m.body_jntadr[1];
      ok = m.body_jntnum[1] == 1 && m.jnt_type[j0] == ABR_JNT_FREE && m.jnt_dofadr[j0] == 0 && m.jnt_qposadr[j0] == 0;
    }
    for (int b = 2; b < nb && ok; b++) {
      const int j = m.body_jntadr[b];
      ok = m.body_jntnum[b] == 1 && (m.jnt_type[j] == ABR_JNT_HINGE || m.jnt_type[j] == ABR_JNT_SLIDE);
    }
    for (int d = 0; d < nv && ok; d++) ok = dof_actnum[d] <= 1;
    for (int ci = 0; ci < L.ncon && ok; ci++) {  // every contact: a plane on the world body against a sphere on the tree
      const int p = con_pair[ci];
      ok = m.pair_kind[p] == ABR_PAIR_PLANE_SPHERE && m.geom_bodyid[m.pair_geom1[p]] == 0 && m.geom_bodyid[m.pair_geom2[p]] >= 1;
    }
    // lane blocks: a leaf takes one lane; an inner body the next power of two of its children's blocks
    std::vector<int> bsize(nb, 1), lane0(nb, 0), bdepth(nb, 0);
    auto np2 = [](int x) { int r = 1; while (r < x) r *= 2; return r; };
    if (ok) {
      for (int b = nb - 1; b >= 1; b--) {
        if (!childnum[b]) continue;
        int sum = 0;
        for (int q = 0; q < childnum[b]; q++) sum += bsize[child[childadr[b] + q]];
        bsize[b] = np2(sum);
      }
      ok = bsize[1] <= limb::kStride;
    }
    int NLm = 0, NCm = 0, G = 1;
    std::vector<int> con_lane(L.ncon, 0), con_slot(L.ncon, 0);
    if (ok) {
      G = bsize[1];
      for (int b = 1; b < nb; b++) {
        if (b > 1) bdepth[b] = bdepth[m.body_parentid[b]] + 1;
        NLm = std::max(NLm, bdepth[b]);
        std::vector<int> kids(child.begin() + childadr[b], child.begin() + childadr[b] + childnum[b]);
        std::stable_sort(kids.begin(), kids.end(), [&](int x, int y) { return bsize[x] > bsize[y]; });
        int off = lane0[b];
        for (int c : kids) { lane0[c] = off; off += bsize[c]; }
      }
      std::vector<int> cnt(limb::kStride, 0);
      for (int ci = 0; ci < L.ncon; ci++) {
        const int l = lane0[m.geom_bodyid[m.pair_geom2[con_pair[ci]]]];
        con_lane[ci] = l; con_slot[ci] = cnt[l]++;
        NCm = std::max(NCm, cnt[l]);
      }
    }
    static const int kInst[][2] = {{3, 1}, {6, 4}};  // compiled <NL, NC> instantiations (abr_limb_*.cu)
    int NLi = -1, NCi = -1;
    if (ok) {
      for (auto& in : kInst) if (in[0] >= NLm && in[1] >= NCm) { NLi = in[0]; NCi = in[1]; break; }
      ok = NLi >= 0;
    }
    if (ok) {
      const limb::Map mp{NLi, NCi};
      std::vector<float> T((size_t)mp.total() * limb::kStride, 0.f);
      auto setf = [&](int slot, int g, float v) { T[(size_t)slot * limb::kStride + g] = v; };
      auto seti = [&](int slot, int g, int v) { float f; memcpy(&f, &v, 4); T[(size_t)slot * limb::kStride + g] = f; };
      const float benign[kConPrm] = {0.f, 0.f, 0.9f, 0.95f, 1000.f, 0.5f, 2.f, 1.f, 2.f, 2.f, 1.f, 0.f, 0.f, 0.f};
      int mxbits = 0;
      double mass = 0;
      for (int b = 1; b < nb; b++) mass += m.body_mass[b];
      for (int g = 0; g < limb::kStride; g++) {
        int ownbits = 0, lvlbits = 0;
        for (int p = 0; p <= NLi; p++) {
          int body = -1;
          if (g < G) for (int b = 1; b < nb; b++) if (bdepth[b] == p && g >= lane0[b] && g < lane0[b] + bsize[b]) body = b;
          const int sb = mp.body(p);
          if (body >= 0) {
            int lv = 0;
            while ((1 << lv) < bsize[body]) lv++;
            if (g == lane0[body]) ownbits |= 1 << p;
            lvlbits |= lv << (2 * p);
            if (lv > ((mxbits >> (2 * p)) & 3)) mxbits = (mxbits & ~(3 << (2 * p))) | (lv << (2 * p));
            fold_body(setf, sb, g, body, p > 0);
          } else {
            ownbits |= 1 << p;  // padding is private
            setf(sb + 3, g, 1.f);
          }
          if (p == 0) continue;
          const int sj = mp.jnt(p), si = mp.ijnt(p);
          for (int i = 0; i < kRowPrm; i++) setf(sj + 14 + i, g, benign[i]);
          if (body >= 0) {
            fold_joint(setf, seti, sj, si, g, body);
          } else {
            setf(sj + 10, g, 1.f);  // unit armature keeps the padded dof's pivot at 1
            seti(si, g, 0); seti(si + 1, g, -1); seti(si + 2, g, -1); seti(si + 3, g, -1);
          }
        }
        for (int i = 0; i < 6; i++) { setf(mp.trunk() + i, g, m.dof_damping[i]); setf(mp.trunk() + 6 + i, g, m.dof_armature[i]); }
        for (int c = 0; c < NCi; c++) {
          const int sc = mp.con(c);
          for (int i = 0; i < kConPrm; i++) setf(sc + 4 + i, g, benign[i]);
          setf(sc + 20, g, 1.f); setf(sc + 24, g, 1.f); setf(sc + 28, g, 1.f); setf(sc + 32, g, 1.f);
          seti(mp.icon(c), g, -1); seti(mp.icon(c) + 1, g, 3);
        }
        seti(mp.ish(), g, ownbits); seti(mp.ish() + 1, g, lvlbits);
      }
      for (int ci = 0; ci < L.ncon; ci++) {
        const int g = con_lane[ci], c = con_slot[ci], p = con_pair[ci];
        const int g1 = m.pair_geom1[p], g2 = m.pair_geom2[p];
        const int sc = mp.con(c);
        for (int i = 0; i < 3; i++) setf(sc + i, g, m.geom_pos[3 * g2 + i]);
        setf(sc + 3, g, m.geom_size[3 * g2]);
        for (int i = 0; i < kConPrm; i++) setf(sc + 4 + i, g, con_prm[(size_t)kConPrm * p + i]);
        // plane on the world body: normal = z axis of the geom frame, frame as in math.make_frame
        const float* q = &m.geom_quat[4 * g1];
        float a[3] = {2.f * (q[1] * q[3] + q[0] * q[2]), 2.f * (q[2] * q[3] - q[0] * q[1]), q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3]};
        float an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        for (int i = 0; i < 3; i++) { setf(sc + 18 + i, g, a[i]); setf(sc + 21 + i, g, m.geom_pos[3 * g1 + i]); }
        float fa[3] = {a[0] / an, a[1] / an, a[2] / an}, fb[3] = {0.f, 0.f, 0.f};
        if (-0.5f < fa[1] && fa[1] < 0.5f) fb[1] = 1.f; else fb[2] = 1.f;
        const float ab = fa[0] * fb[0] + fa[1] * fb[1] + fa[2] * fb[2];
        for (int i = 0; i < 3; i++) fb[i] -= fa[i] * ab;
        const float bn = std::sqrt(fb[0] * fb[0] + fb[1] * fb[1] + fb[2] * fb[2]);
        for (int i = 0; i < 3; i++) fb[i] /= bn;
        const float fcx[3] = {fa[1] * fb[2] - fa[2] * fb[1], fa[2] * fb[0] - fa[0] * fb[2], fa[0] * fb[1] - fa[1] * fb[0]};
        for (int i = 0; i < 3; i++) { setf(sc + 24 + i, g, fa[i]); setf(sc + 27 + i, g, fb[i]); setf(sc + 30 + i, g, fcx[i]); }
        seti(mp.icon(c), g, bdepth[m.geom_bodyid[g2]]); seti(mp.icon(c) + 1, g, m.pair_condim[p]);
      }
      // contact-body form: each lane's contacts all on its leaf body, against one plane
      L.l_cb = 1;
      {
        std::vector<int> lane_plane(limb::kStride, -1);
        for (int ci = 0; ci < L.ncon; ci++) {
          const int p = con_pair[ci], g1 = m.pair_geom1[p], b2 = m.geom_bodyid[m.pair_geom2[p]], g = con_lane[ci];
          if (childnum[b2] != 0) L.l_cb = 0;                                  // not a leaf
          if (lane_plane[g] >= 0 && lane_plane[g] != g1) L.l_cb = 0;          // a second plane
          lane_plane[g] = g1;
        }
      }
      while (P.f.size() % 4) P.f.push_back(0.f);
      L.limb_ok = 1; L.lNL = NLi; L.lNC = NCi; L.l_mx = mxbits; L.l_mass = (float)mass;
      L.l_pow2 = 1; L.l_hinge = 1;
      for (int j = 0; j < nj; j++) if (m.jnt_type[j] == ABR_JNT_SLIDE) L.l_hinge = 0;
      for (size_t i = 0; i < lim_prm.size(); i += kRowPrm) if (lim_prm[i + 6] != 2.f) L.l_pow2 = 0;
      for (size_t i = 0; i < con_prm.size(); i += kConPrm) if (con_prm[i + 6] != 2.f) L.l_pow2 = 0;
      L.lg2G = 0;
      while ((1 << L.lg2G) < G) L.lg2G++;
      L.f_ltab = P.addf(T);
    }
  }

  // ---- hand decomposition (abr_hand.cuh): fixed-base chains + joint equalities, one lane per chain off a static body
  L.hand_ok = 0;
  {
    bool ok = getenv("ABR_NO_HAND") == nullptr && !L.limb_ok && opt.solver == ABR_SOLVER_NEWTON && opt.integrator == ABR_INT_EULER &&
              L.ncon == 0 && L.ne <= hand::kNE && nv >= 1 && nb >= 2;
    std::vector<int> moving(nb, 0), pos_in_chain(nb, 0), lane_of(nb, -1), starts;
    for (int b = 1; b < nb && ok; b++) {
      const int par = m.body_parentid[b];
      if (m.body_jntnum[b] == 0) { ok = !moving[par]; continue; }  // a static body may only hang off static bodies
      const int j = m.body_jntadr[b];
      ok = m.body_jntnum[b] == 1 && (m.jnt_type[j] == ABR_JNT_HINGE || m.jnt_type[j] == ABR_JNT_SLIDE);
      moving[b] = 1;
      if (moving[par]) { pos_in_chain[b] = pos_in_chain[par] + 1; lane_of[b] = lane_of[par]; }
      else { pos_in_chain[b] = 1; lane_of[b] = (int)starts.size(); starts.push_back(b); }
    }
    for (int b = 1; b < nb && ok; b++) {  // chains do not fork: at most one moving child per moving body
      if (!moving[b]) continue;
      int kids = 0;
      for (int q = 0; q < childnum[b]; q++) kids += moving[child[childadr[b] + q]];
      ok = kids <= 1 && pos_in_chain[b] <= 3;
    }
    for (int d = 0; d < nv && ok; d++) ok = dof_actnum[d] <= 1;
    ok = ok && !starts.empty() && (int)starts.size() <= hand::kLanes;
    if (ok) {
      constexpr int NLh = 3;
      const hand::Map mp{NLh};
      const int G = hand::kLanes;  // always four lanes per world: chains padded with dummy lanes (the dense Newton system is of order 4 NL)
      std::vector<float> T((size_t)mp.total() * limb::kStride, 0.f);
      auto setf = [&](int slot, int g, float v) { T[(size_t)slot * limb::kStride + g] = v; };
      auto seti = [&](int slot, int g, int v) { float f; memcpy(&f, &v, 4); T[(size_t)slot * limb::kStride + g] = f; };
      const float benign[kRowPrm] = {0.f, 0.f, 0.9f, 0.95f, 1000.f, 0.5f, 2.f, 1.f, 2.f, 2.f};
      // world pose of every static body (double precision, once)
      std::vector<double> wp(3 * nb, 0.0), wq(4 * nb, 0.0);
      wq[0] = 1.0;
      for (int b = 1; b < nb; b++) {
        if (moving[b]) continue;
        const int par = m.body_parentid[b];
        const double* q = &wq[4 * par];
        const double R[9] = {q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3], 2 * (q[1] * q[2] - q[0] * q[3]), 2 * (q[1] * q[3] + q[0] * q[2]),
                             2 * (q[1] * q[2] + q[0] * q[3]), q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3], 2 * (q[2] * q[3] - q[0] * q[1]),
                             2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3]};
        for (int i = 0; i < 3; i++)
          wp[3 * b + i] = wp[3 * par + i] + R[3 * i] * m.body_pos[3 * b] + R[3 * i + 1] * m.body_pos[3 * b + 1] + R[3 * i + 2] * m.body_pos[3 * b + 2];
        const double a[4] = {q[0], q[1], q[2], q[3]}, c[4] = {m.body_quat[4 * b], m.body_quat[4 * b + 1], m.body_quat[4 * b + 2], m.body_quat[4 * b + 3]};
        wq[4 * b] = a[0] * c[0] - a[1] * c[1] - a[2] * c[2] - a[3] * c[3];
        wq[4 * b + 1] = a[0] * c[1] + a[1] * c[0] + a[2] * c[3] - a[3] * c[2];
        wq[4 * b + 2] = a[0] * c[2] - a[1] * c[3] + a[2] * c[0] + a[3] * c[1];
        wq[4 * b + 3] = a[0] * c[3] + a[1] * c[2] - a[2] * c[1] + a[3] * c[0];
      }
      for (int g = 0; g < limb::kStride; g++) {
        setf(mp.base() + 3, g, 1.f);
        if (g < (int)starts.size()) {
          const int par = m.body_parentid[starts[g]];
          for (int i = 0; i < 3; i++) setf(mp.base() + i, g, (float)wp[3 * par + i]);
          for (int i = 0; i < 4; i++) setf(mp.base() + 3 + i, g, (float)wq[4 * par + i]);
        }
        for (int p = 1; p <= NLh; p++) {
          int body = -1;
          for (int b = 1; b < nb; b++) if (moving[b] && lane_of[b] == g && pos_in_chain[b] == p) body = b;
          const int sb = mp.body(p), sj = mp.jnt(p), si = mp.ijnt(p);
          for (int i = 0; i < kRowPrm; i++) setf(sj + 14 + i, g, benign[i]);
          if (body >= 0) {
            fold_body(setf, sb, g, body, true);
            fold_joint(setf, seti, sj, si, g, body);
          } else {
            setf(sb + 3, g, 1.f);
            setf(sj + 10, g, 1.f);  // unit armature keeps the padded dof's pivot at 1
            seti(si, g, 0); seti(si + 1, g, -1); seti(si + 2, g, -1); seti(si + 3, g, -1);
          }
        }
        for (int r = 0; r < hand::kNE; r++) {
          const int se = mp.eq(r);
          for (int i = 0; i < kRowPrm; i++) setf(se + i, g, benign[i]);
          seti(mp.ieq(r), g, 0); seti(mp.ieq(r) + 1, g, 0);
          if (r >= L.ne) continue;
          for (int i = 0; i < kRowPrm; i++) setf(se + i, g, eq_prm[(size_t)kRowPrm * r + i]);
          for (int i = 0; i < 5; i++) setf(se + kRowPrm + i, g, eq_data[(size_t)5 * r + i]);
          setf(se + kRowPrm + 5, g, 1.f);
          const int b1 = m.jnt_bodyid[eq_j1[r]], b2 = eq_j2[r] >= 0 ? m.jnt_bodyid[eq_j2[r]] : -1;
          if (lane_of[b1] == g) seti(mp.ieq(r), g, pos_in_chain[b1]);
          if (b2 >= 0 && lane_of[b2] == g) seti(mp.ieq(r) + 1, g, pos_in_chain[b2]);
        }
      }
      while (P.f.size() % 4) P.f.push_back(0.f);
      L.hand_ok = 1; L.lNL = NLh; L.lNC = 0;
      L.lg2G = 0;
      while ((1 << L.lg2G) < G) L.lg2G++;
      L.f_htab = P.addf(T);
    }
  }

  while (P.f.size() % 4) P.f.push_back(0.f);
  while (P.i.size() % 4) P.i.push_back(0);
  L.n_mf = (int)P.f.size(); L.n_mi = (int)P.i.size();
  mf = P.f; mi = P.i;

  // ---- per-world shared-memory layout
  const int ne = L.nefc, nc = L.ncon, ntri = L.sparse ? L.nnz : L.ntri;
  int off = 0;
  auto take = [&](int n) { int o = off; off += n; return o; };
  // persistent across the step
  L.w_qpos = take(nq); L.w_qvel = take(nv);   // x = [qpos; qvel] contiguous
  L.w_warm = take(nv); L.w_ctrl = take(nu);
  L.w_M = take(ntri); L.w_H = take(ntri);
  L.w_eqc = take(L.ne); L.w_lims = take(L.nl); L.w_B = take(3 * nc * nv);
  L.w_D = take(ne); L.w_aref = take(ne);
  L.w_fs = take(nv); L.w_as = take(nv); L.w_a = take(nv); L.w_fc = take(nv); L.w_y = take(nv); L.w_invD = take(nv);
  L.w_cinert = take(10 * nb); L.w_cdof = take(6 * nv);
  L.w_actf = take(nu); L.w_bv = take(3 * nc);
  int rk = 2 * nv;  // CG Polak-Ribiere history
  if (opt.integrator == ABR_INT_RK4) rk += nq + 6 * nv;
  L.w_rk = take(rk);
  const int base = off;
  // region A (position phase): poses, contact geometry            | solver vectors
  L.w_xpos = take(3 * nb); L.w_xquat = take(4 * nb); L.w_xipos = take(3 * nb);
  L.w_xanchor = take(3 * nj); L.w_xaxis = take(3 * nj); L.w_rootcom = take(3 * L.nroot);
  L.w_cdist = take(nc); L.w_cpos = take(3 * nc); L.w_cframe = take(9 * nc);
  const int endA = off;
  if (alias) off = base;
  L.w_Ma = take(nv); L.w_grad = take(nv); L.w_search = take(nv); L.w_mv = take(nv);
  L.w_Jaref = take(ne); L.w_jv = take(ne); L.w_force = take(ne); L.w_Fc = take(3 * nc); L.w_WB = take(3 * nc * nv);
  const int endS = off;
  off = std::max(endA, endS);
  // region B: crb + buf (inertia phase)                            | cvel, cacc, cdof_dot (velocity phase)
  const int baseB = off;
  L.w_crb = take(10 * nb); L.w_buf = take(6 * nv);
  const int endB1 = off;
  if (alias) off = baseB;
  L.w_cvel = take(6 * nb); L.w_cacc = take(6 * nb); L.w_cdofdot = take(6 * nv);
  off = std::max(endB1, off);
  // NOTE: region A holds rootcom/xpos/xquat, which stage_collision (before crb) and stage_com use;
  // the solver vectors only become live in stage_solve, after every reader of region A.
  L.world_stride = (off + 3) & ~3;
  return ABR_OK;
}

// segmented first-minimum argmin, NaN counts as the minimum (jnp.argmin, shooting.py:154)
__device__ __forceinline__ bool better(float ca, int ia, float cb, int ib) {
  const bool na = isnan(ca), nb = isnan(cb);
  if (na != nb) return na;
  if (!na && ca != cb) return ca < cb;
  return ia < ib;
}
__global__ void k_argmin(const float* costs, int S, int sample_offset, int* best_idx, float* best_cost) {
  __shared__ float sc[256];
  __shared__ int si[256];
  const int b = blockIdx.x;
  const float* cs = costs + (size_t)b * S;
  float bc = 0.f; int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = cs[i];
    if (bi == 0x7fffffff || better(v, i, bc, bi)) { bc = v; bi = i; }
  }
  sc[threadIdx.x] = bc; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float c2 = sc[threadIdx.x + o]; const int i2 = si[threadIdx.x + o];
      if (i2 != 0x7fffffff && (si[threadIdx.x] == 0x7fffffff || better(c2, i2, sc[threadIdx.x], si[threadIdx.x]))) {
        sc[threadIdx.x] = c2; si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { best_idx[b] = sample_offset + si[0]; best_cost[b] = sc[0]; }
}

// copy each problem's winning trajectory / control sequence out of the per-sample scratch (shooting.py:155-156)
__global__ void k_gather_winner(const float* xs_all, const float* us_all, const int* best_idx, int sample_offset, int S, int nxs, int nus,
                                float* xs_star, float* us_star) {
  const int b = blockIdx.x;
  const size_t w = (size_t)b * S + (best_idx[b] - sample_offset);
  if (xs_star) for (int i = threadIdx.x; i < nxs; i += blockDim.x) xs_star[(size_t)b * nxs + i] = xs_all[w * nxs + i];
  if (us_star) for (int i = threadIdx.x; i < nus; i += blockDim.x) us_star[(size_t)b * nus + i] = us_all[w * nus + i];
}

// one MPC tick's bookkeeping: plant state <- winner's state after its first control, logs, shifted guess
__global__ void k_mpc_advance(const float* xs_star, const float* us_star, const int* best_idx, const float* best_cost, int N, int nx, int nu,
                              int tick, float* x, float* us_guess, float* xs_log, float* us_log, float* cost_log, int* idx_log) {
  const int i = threadIdx.x;
  if (xs_log && tick == 0) for (int k = i; k < nx; k += blockDim.x) xs_log[k] = x[k];
  for (int k = i; k < nx; k += blockDim.x) {
    const float v = xs_star[nx + k];
    x[k] = v;
    if (xs_log) xs_log[(size_t)(tick + 1) * nx + k] = v;
  }
  for (int k = i; k < N * nu; k += blockDim.x) {
    const int t = k / nu;
    us_guess[k] = us_star[(t + 1 < N ? k + nu : k)];
  }
  if (us_log) for (int k = i; k < nu; k += blockDim.x) us_log[(size_t)tick * nu + k] = us_star[k];
  if (i == 0) { if (cost_log) cost_log[tick] = best_cost[0]; if (idx_log) idx_log[tick] = best_idx[0]; }
}

// FP32 FMA-pipe peak: 8 independent FFMA chains per thread
__global__ void __launch_bounds__(1024) k_ffma(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678f) out[0] = s;
}

// ================================================================================ launch helpers
static int pick_lanes(const AbrModel* m, int nworld) {
  if (m->lanes > 1) return m->lanes;
  const char* e = getenv("ABR_LANES");
  if (e && atoi(e) > 1) return atoi(e);
  // small models (a hand, a pendulum) have too little work per world for 32 lanes: as soon as narrower groups still give
  // every SM sub-partition a warp, they win (bh280, 4096 x 32 solve: 2.07 ms with 8 lanes, 3.55 ms with 32)
  if (m->lay.nv <= 10) {
    const long subparts = 4L * m->num_sms;
    // many contact slots (convex pairs carry 2 - 4 each and every slot runs its pair's selection on its own lane): 16 lanes take them in
    // fewer rounds (blocks.xml, 19 slots: 3.6e6 / 4.0e6 world-steps/s at 2048 / 16384 worlds with 16 lanes, 2.8e6 / 3.1e6 with 8, 3.1e6 with 32)
    if (m->lay.ncon > 8 && (long)nworld * 16 / 32 >= subparts / 2) return 16;
    if ((long)nworld * 8 / 32 >= subparts / 2) return 8;
    if ((long)nworld * 16 / 32 >= subparts / 2) return 16;
    return 32;
  }
  // enough worlds to fill the machine with narrow groups -> better lane efficiency
  const long per_sm = (long)nworld / m->num_sms;
  if (per_sm >= 256) return 8;
  if (per_sm >= 64) return 16;
  return 32;
}

static int launch_result(int rc) {
  if (rc == 0) return ABR_OK;
  if (rc == -1000) return fail(ABR_ECAPACITY, "model does not fit in shared memory");
  return fail(ABR_ECUDA, std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
}
// the limb path serves eligible models unless a generic group size was pinned (lanes = 4..32);
// lanes = 1 pins the limb path. Dense (non-diagonal) cost matrices stay on the generic kernels.
static bool use_limb(const AbrModel* m, const Layout& L, bool dense_cost, bool debug) {
  if (!L.limb_ok || debug || dense_cost) return false;
  int lanes = m->lanes;
  if (!lanes) { const char* e = getenv("ABR_LANES"); if (e) lanes = atoi(e); }
  return lanes == 0 || lanes == 1;
}
// the hand path serves fixed-base chains with joint equalities (abr_hand.cuh) unless a generic group size was pinned
static bool use_hand(const AbrModel* m, const Layout& L, bool dense_cost) {
  if (!L.hand_ok || dense_cost) return false;
  int lanes = m->lanes;
  if (!lanes) { const char* e = getenv("ABR_LANES"); if (e) lanes = atoi(e); }
  return lanes == 0 || lanes == 1;
}
static int launch_rollout(const AbrModel* m, const Layout& L, const RolloutArgs& a, cudaStream_t st) {
  if (a.nworld <= 0) return ABR_OK;
  if (use_hand(m, L, a.cost.enabled && !a.cost.diag) && a.t_begin == 0 && a.t_end == a.N) return launch_result(launch_hand_rollout_3(L, a, st));
  if (use_limb(m, L, a.cost.enabled && !a.cost.diag, false)) {
    const bool spec = getenv("ABR_LIMB_GENERAL") == nullptr;  // compile-time sharing patterns: 2 = flat 4 lanes, 86 = biped
    // fast variants (limb::Spec): the common options (no disable flag other than eulerdamp, one Newton iteration, default impedance power, hinge joints), specialised on
    // the mode and on whether anything is written out;
    // every other combination of options runs the general variant of the same kernel (ABR_LIMB_NOSPEC forces it)
    const bool fast = (L.disableflags == ABR_DSBL_EULERDAMP || L.disableflags == 0) && L.iterations == 1 && L.l_pow2 && L.l_hinge &&
                      getenv("ABR_LIMB_NOSPEC") == nullptr;
    const int sv = !fast ? -1 : (L.disableflags == 0 ? 1 : 0) | ((a.xs_out || a.us_out) ? 4 : 0) | (a.mode == 1 ? 8 : 0);
#define ABR_PICK_ROLLOUT(NL, NC, TAG)                                                                      \
  (sv == 0 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s0(L, a, st) : sv == 4 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s4(L, a, st) \
   : sv == 8 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s8(L, a, st) : sv == 12 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s12(L, a, st) \
   : sv == 1 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s1(L, a, st) : sv == 5 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s5(L, a, st) \
   : sv == 9 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s9(L, a, st) : sv == 13 ? launch_limb_rollout_##NL##_##NC##_##TAG##_s13(L, a, st) \
   : launch_limb_rollout_##NL##_##NC##_##TAG##_sg(L, a, st))
    if (L.lNL == 3 && L.lNC == 1)
      return launch_result(spec && L.l_mx == 2 ? ABR_PICK_ROLLOUT(3, 1, f2) : spec && L.l_mx == 3 ? ABR_PICK_ROLLOUT(3, 1, f3) : launch_limb_rollout_3_1_g_sg(L, a, st));
    if (L.lNL == 6 && L.lNC == 4)
      return launch_result(spec && L.l_cb && L.l_mx == 86 ? ABR_PICK_ROLLOUT(6, 4, b) : spec && L.l_cb && L.l_mx == 1 ? ABR_PICK_ROLLOUT(6, 4, l2)
                           : spec && L.l_cb && L.l_mx == 2 ? ABR_PICK_ROLLOUT(6, 4, f2) : launch_limb_rollout_6_4_g_sg(L, a, st));
#undef ABR_PICK_ROLLOUT
    return fail(ABR_ECAPACITY, "no compiled limb kernel for this model");
  }
  LaunchCfg cfg{m->max_smem};
  switch (pick_lanes(m, a.nworld)) {
    case 4: return launch_result(launch_rollout_4(cfg, L, a, st));
    case 8: return launch_result(launch_rollout_8(cfg, L, a, st));
    case 16: return launch_result(launch_rollout_16(cfg, L, a, st));
    default: return launch_result(launch_rollout_32(cfg, L, a, st));
  }
}
static int launch_env(const AbrModel* m, const Layout& L, const EnvArgs& a_in, cudaStream_t st) {
  if (a_in.E <= 0) return ABR_OK;
  EnvArgs a = a_in;
  if (m->dr) {
    if (a.E != m->dr_E) return fail(ABR_EINVAL, "env call: batch size differs from the one given to abr_env_set_randomization");
    a.dr = m->dr; a.dr_n = m->dr_n;
  }
  if (use_hand(m, L, false) && !a.dbg && !a.fo_on && !a.t_steps && !a.dr) return launch_result(launch_hand_env_3(L, a, st));
  if (use_limb(m, L, false, a.dbg != nullptr || a.fo_on)) {
    const bool spec = getenv("ABR_LIMB_GENERAL") == nullptr;
    const bool fast = (L.disableflags == ABR_DSBL_EULERDAMP || L.disableflags == 0) && L.iterations == 1 && L.l_pow2 && L.l_hinge &&
                      getenv("ABR_LIMB_NOSPEC") == nullptr;
    const bool ed = L.disableflags == 0;
#define ABR_PICK_ENV(NL, NC, TAG)                                                                             \
  (fast ? (ed ? launch_limb_env_##NL##_##NC##_##TAG##_s1(L, a, st) : launch_limb_env_##NL##_##NC##_##TAG##_s0(L, a, st)) \
        : launch_limb_env_##NL##_##NC##_##TAG##_sg(L, a, st))
    if (L.lNL == 3 && L.lNC == 1)
      return launch_result(spec && L.l_mx == 2 ? ABR_PICK_ENV(3, 1, f2) : spec && L.l_mx == 3 ? ABR_PICK_ENV(3, 1, f3) : launch_limb_env_3_1_g_sg(L, a, st));
    if (L.lNL == 6 && L.lNC == 4)
      return launch_result(spec && L.l_cb && L.l_mx == 86 ? ABR_PICK_ENV(6, 4, b) : spec && L.l_cb && L.l_mx == 1 ? ABR_PICK_ENV(6, 4, l2)
                           : spec && L.l_cb && L.l_mx == 2 ? ABR_PICK_ENV(6, 4, f2) : launch_limb_env_6_4_g_sg(L, a, st));
#undef ABR_PICK_ENV
    return fail(ABR_ECAPACITY, "no compiled limb kernel for this model");
  }
  LaunchCfg cfg{m->max_smem};
  switch (pick_lanes(m, a.E)) {
    case 4: return launch_result(launch_env_4(cfg, L, a, st));
    case 8: return launch_result(launch_env_8(cfg, L, a, st));
    case 16: return launch_result(launch_env_16(cfg, L, a, st));
    default: return launch_result(launch_env_32(cfg, L, a, st));
  }
}

static CostView cost_view(const AbrCost* c) {
  CostView v;
  memset(&v, 0, sizeof(v));
  if (!c) return v;
  const int nx = c->nx, nu = c->nu;
  v.Q = c->d; v.Qf = v.Q + nx * nx; v.R = v.Qf + nx * nx; v.xg = v.R + nu * nu;
  v.qd = v.xg + nx; v.qfd = v.qd + nx; v.rd = v.qfd + nx;
  v.enabled = 1; v.diag = c->diag;
  return v;
}

static int rebuild(AbrModel* m) {
  int rc = build_blob(m->hm, getenv("ABR_NO_ALIAS") == nullptr, m->lay, m->mf, m->mi);
  if (rc) return rc;
  std::vector<float> f2; std::vector<int> i2;
  rc = build_blob(m->hm, false, m->lay_dbg, f2, i2);
  if (rc) return rc;
  ABR_ON_DEVICE(m->device);
  if (m->d_blob) { cudaFree(m->d_blob); m->d_blob = nullptr; }
  const size_t bytes = sizeof(float) * ((size_t)m->lay.n_mf + m->lay.n_mi);
  CK(cudaMalloc(&m->d_blob, bytes));
  std::vector<float> blob(m->lay.n_mf + m->lay.n_mi);
  memcpy(blob.data(), m->mf.data(), sizeof(float) * m->lay.n_mf);
  memcpy(blob.data() + m->lay.n_mf, m->mi.data(), sizeof(int) * m->lay.n_mi);
  CK(cudaMemcpy(m->d_blob, blob.data(), bytes, cudaMemcpyHostToDevice));
  return ABR_OK;
}

// ================================================================================ C ABI
extern "C" {

const char* abr_last_error(void) { return g_err.c_str(); }
int abr_version(void) { return ABR_VERSION; }
size_t abr_sizeof_model_host(void) { return sizeof(AbrModelHost); }
size_t abr_sizeof_opt(void) { return sizeof(AbrOpt); }
int abr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int abr_model_create(const AbrModelHost* host, int device, AbrModel** out) {
  if (!host || !out) return fail(ABR_EINVAL, "abr_model_create: null argument");
  *out = nullptr;
  if (abr_device_count() <= 0) return fail(ABR_ENODEVICE, "no CUDA device: the engine has no CPU path");
  if (host->nq < 0 || host->nv < 0 || host->nbody < 1) return fail(ABR_EINVAL, "abr_model_create: bad sizes");
  {  // the collision functions index per-lane arrays of fixed size with the hull tables: check them here, once
    const AbrModelHost* h = host;
    if (h->npair < 0 || h->ngeom < 0 || h->nvert < 0 || h->nface < 0 || h->nfacevert < 0 || h->nedge < 0) return fail(ABR_EINVAL, "abr_model_create: bad sizes");
    if (h->ngeom > 0 && (!h->geom_vertadr || !h->geom_vertnum || !h->geom_faceadr || !h->geom_facenum || !h->geom_edgeadr || !h->geom_edgenum))
      return fail(ABR_EINVAL, "abr_model_create: null hull table");
    if ((h->nface > 0 && (!h->face_vertadr || !h->face_vertnum || !h->face_normal)) || (h->nfacevert > 0 && !h->face_vert) || (h->nedge > 0 && !h->edge_vert) ||
        (h->npair > 0 && (!h->pair_kind || !h->pair_geom1 || !h->pair_geom2)))
      return fail(ABR_EINVAL, "abr_model_create: null hull / pair table");
    for (int g = 0; g < h->ngeom; g++) {
      const int va = h->geom_vertadr[g], vn = h->geom_vertnum[g], fa = h->geom_faceadr[g], fn = h->geom_facenum[g], ea = h->geom_edgeadr[g], en = h->geom_edgenum[g];
      if (vn < 0 || va < 0 || va + vn > h->nvert || fn < 0 || fa < 0 || fa + fn > h->nface || en < 0 || ea < 0 || ea + en > h->nedge)
        return fail(ABR_EINVAL, "abr_model_create: hull table of a geom out of range");
      for (int f = fa; f < fa + fn; f++) {
        const int ca = h->face_vertadr[f], cn = h->face_vertnum[f];
        if (cn < 3 || cn > ABR_MAX_FACE_VERTS || ca < 0 || ca + cn > h->nfacevert) return fail(ABR_EINVAL, "abr_model_create: a hull face needs 3 .. ABR_MAX_FACE_VERTS corners");
        for (int k = ca; k < ca + cn; k++) if (h->face_vert[k] < 0 || h->face_vert[k] >= vn) return fail(ABR_EINVAL, "abr_model_create: face corner out of the geom's vertex set");
      }
      for (int e = 2 * ea; e < 2 * (ea + en); e++) if (h->edge_vert[e] < 0 || h->edge_vert[e] >= vn) return fail(ABR_EINVAL, "abr_model_create: edge end point out of the geom's vertex set");
    }
    for (int p = 0; p < h->npair; p++) {
      const int k = h->pair_kind[p], g1 = h->pair_geom1[p], g2 = h->pair_geom2[p];
      if (k < 0 || k > ABR_PAIR_CONVEX_CONVEX || g1 < 0 || g1 >= h->ngeom || g2 < 0 || g2 >= h->ngeom) return fail(ABR_EINVAL, "abr_model_create: bad contact pair");
      const bool hull2 = k == ABR_PAIR_SPHERE_CONVEX || k == ABR_PAIR_CAPSULE_CONVEX || k == ABR_PAIR_CONVEX_CONVEX;
      if (hull2 && (h->geom_facenum[g2] < 1 || h->geom_vertnum[g2] > ABR_MAX_CONVEX_VERTS)) return fail(ABR_EINVAL, "abr_model_create: convex pair needs a hull with faces and at most ABR_MAX_CONVEX_VERTS vertices");
      if (k == ABR_PAIR_CONVEX_CONVEX && (h->geom_facenum[g1] < 1 || h->geom_vertnum[g1] > ABR_MAX_CONVEX_VERTS)) return fail(ABR_EINVAL, "abr_model_create: convex pair needs a hull with faces and at most ABR_MAX_CONVEX_VERTS vertices");
      if (k == ABR_PAIR_PLANE_CONVEX && h->geom_vertnum[g2] < 1) return fail(ABR_EINVAL, "abr_model_create: plane - convex pair without vertices");
    }
  }
  ABR_ON_DEVICE(device);  // the blob, the handle's streams and its scratch all live on `device`
  AbrModel* m = new AbrModel();
  m->device = device;
  copy_host_model(host, m->hm);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete m; return fail(ABR_ECUDA, "cudaGetDeviceProperties failed"); }
  m->num_sms = prop.multiProcessorCount;
  m->max_smem = (int)prop.sharedMemPerBlockOptin;
  int rc = rebuild(m);
  if (rc) { delete m; return rc; }
  if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { delete m; return fail(ABR_ECUDA, "cudaStreamCreate failed"); }
  *out = m;
  return ABR_OK;
}

int abr_model_destroy(AbrModel* m) {
  if (!m) return ABR_OK;
  DeviceGuard guard(m->device);
  if (m->d_blob) cudaFree(m->d_blob);
  m->s_costs.release(); m->s_in.release(); m->s_out.release(); m->s_dbg.release(); m->s_traj.release(); m->s_carry.release(); m->s_mpc.release();
  for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return ABR_OK;
}

int abr_model_set_opt(AbrModel* m, const AbrOpt* opt) {
  if (!m || !opt) return fail(ABR_EINVAL, "abr_model_set_opt: null argument");
  AbrOpt old = m->hm.opt;
  m->hm.opt = *opt;
  int rc = rebuild(m);
  if (rc) { m->hm.opt = old; rebuild(m); }
  return rc;
}

int abr_model_get_opt(const AbrModel* m, AbrOpt* opt) {
  if (!m || !opt) return fail(ABR_EINVAL, "abr_model_get_opt: null argument");
  *opt = m->hm.opt;
  return ABR_OK;
}

int abr_model_info(const AbrModel* m, int* ncon, int* ne, int* nl, int* nefc, int* depth) {
  if (!m) return fail(ABR_EINVAL, "abr_model_info: null model");
  if (ncon) *ncon = m->lay.ncon;
  if (ne) *ne = m->lay.ne;
  if (nl) *nl = m->lay.nl;
  if (nefc) *nefc = m->lay.nefc;
  if (depth) *depth = m->lay.depth;
  return ABR_OK;
}

int abr_model_describe(const AbrModel* m, char* buf, int cap) {
  if (!m || !buf || cap <= 0) return fail(ABR_EINVAL, "abr_model_describe: bad argument");
  const Layout& L = m->lay;
  char tmp[256];
  if (use_limb(m, L, false, false)) {
    const bool spec = getenv("ABR_LIMB_GENERAL") == nullptr;
    const bool k31 = L.lNL == 3 && L.lNC == 1, k64 = L.lNL == 6 && L.lNC == 4;
    const bool flat = spec && ((k31 && (L.l_mx == 2 || L.l_mx == 3)) || (k64 && L.l_cb && (L.l_mx == 1 || L.l_mx == 2))), bip = k64 && spec && L.l_mx == 86 && L.l_cb;
    const bool lite = k64 && !bip;  // these families build the eulerdamp-off fast variants only
    const bool fast = (flat || bip) && (L.disableflags == ABR_DSBL_EULERDAMP || (L.disableflags == 0 && !lite)) && L.iterations == 1 && L.l_pow2 && L.l_hinge &&
                      getenv("ABR_LIMB_NOSPEC") == nullptr;
    snprintf(tmp, sizeof(tmp), "limb kernels <NL=%d, NC=%d, %s>, %d lanes per world, %s variant", L.lNL, L.lNC,
             flat ? (L.l_mx == 1 ? "flat 2-lane pattern" : (L.l_mx == 3 ? "flat 8-lane pattern" : "flat 4-lane pattern"))
                  : (bip ? "biped pattern, contact-body form" : "sharing pattern from the table"), 1 << L.lg2G,
             fast ? (L.disableflags == 0 ? "fast (eulerdamp on)" : "fast (eulerdamp off)") : "general");
  } else if (use_hand(m, L, false)) {
    snprintf(tmp, sizeof(tmp), "hand kernels <NL=%d> (fixed-base chains, %d joint-equality rows), %d lanes per world",
             L.lNL, L.ne, 1 << L.lg2G);
  } else {
    snprintf(tmp, sizeof(tmp), "generic kernels, %s lanes per world%s", m->lanes > 1 ? std::to_string(m->lanes).c_str() : "8 / 16 / 32 (by batch size)",
             L.limb_ok ? " (limb kernels available: abr_model_set_lanes(m, 0 or 1))" : "");
  }
  snprintf(buf, (size_t)cap, "%s", tmp);
  return ABR_OK;
}

int abr_model_set_lanes(AbrModel* m, int lanes) {
  if (!m) return fail(ABR_EINVAL, "abr_model_set_lanes: null model");
  if (lanes != 0 && lanes != 1 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) return fail(ABR_EINVAL, "lanes must be 0, 1, 4, 8, 16 or 32");
  if (lanes == 1 && !m->lay.limb_ok && !m->lay.hand_ok) return fail(ABR_EUNSUPPORTED, "lanes = 1 pins the limb / hand path, which this model/options are not eligible for");
  m->lanes = lanes;
  return ABR_OK;
}

int abr_limb_plan_host(const AbrModelHost* host, int* info, int* lane_body, int cap, int* lane_own, int* lane_lvl) {
  if (!host || !info) return fail(ABR_EINVAL, "abr_limb_plan_host: null argument");
  HostModel hm;
  copy_host_model(host, hm);
  Layout L;
  std::vector<float> mf; std::vector<int> mi;
  int rc = build_blob(hm, true, L, mf, mi);
  if (rc) return rc;
  memset(info, 0, sizeof(int) * 8);
  info[0] = L.limb_ok; info[5] = L.nefc; info[6] = L.ncon;
  if (!L.limb_ok) return ABR_OK;
  info[1] = L.lg2G; info[2] = L.lNL; info[3] = L.lNC; info[4] = L.l_mx;
  const limb::Map mp{L.lNL, L.lNC};
  const float* T = mf.data() + L.f_ltab;
  auto geti = [&](int slot, int g) { int v; memcpy(&v, &T[(size_t)slot * limb::kStride + g], 4); return v; };
  int used = 0;
  for (int g = 0; g < limb::kStride; g++) {
    bool any = false;
    for (int p = 0; p <= L.lNL; p++) {
      int body = -1;
      if (p == 0) body = (g < (1 << L.lg2G)) ? 1 : -1;
      else { const int d = geti(mp.ijnt(p) + 1, g); body = d >= 0 ? hm.dof_bodyid[d] : -1; any = any || d >= 0; }
      if (lane_body && g * (L.lNL + 1) + p < cap) lane_body[g * (L.lNL + 1) + p] = body;
    }
    if (any) used++;
    if (lane_own) lane_own[g] = geti(mp.ish(), g);
    if (lane_lvl) lane_lvl[g] = geti(mp.ish() + 1, g);
  }
  info[7] = used;
  return ABR_OK;
}

int abr_cost_create(const AbrQuadCostHost* h, int device, AbrCost** out) {
  if (!h || !out || !h->Q || !h->Qf || !h->R || !h->xg) return fail(ABR_EINVAL, "abr_cost_create: null argument");
  *out = nullptr;
  if (abr_device_count() <= 0) return fail(ABR_ENODEVICE, "no CUDA device: the engine has no CPU path");
  const int nx = h->nx, nu = h->nu;
  std::vector<float> buf((size_t)2 * nx * nx + nu * nu + 3 * nx + nu);
  float* p = buf.data();
  memcpy(p, h->Q, sizeof(float) * nx * nx); p += nx * nx;
  memcpy(p, h->Qf, sizeof(float) * nx * nx); p += nx * nx;
  memcpy(p, h->R, sizeof(float) * nu * nu); p += nu * nu;
  memcpy(p, h->xg, sizeof(float) * nx); p += nx;
  bool diag = true;
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < nx; j++)
      if (i != j && (h->Q[i * nx + j] != 0.f || h->Qf[i * nx + j] != 0.f)) diag = false;
  for (int i = 0; i < nu; i++)
    for (int j = 0; j < nu; j++)
      if (i != j && h->R[i * nu + j] != 0.f) diag = false;
  for (int i = 0; i < nx; i++) p[i] = h->Q[i * nx + i];
  p += nx;
  for (int i = 0; i < nx; i++) p[i] = h->Qf[i * nx + i];
  p += nx;
  for (int i = 0; i < nu; i++) p[i] = h->R[i * nu + i];
  AbrCost* c = new AbrCost();
  c->device = device; c->nx = nx; c->nu = nu; c->diag = diag ? 1 : 0;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess || cudaMalloc(&c->d, sizeof(float) * buf.size()) != cudaSuccess ||
      cudaMemcpy(c->d, buf.data(), sizeof(float) * buf.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete c;
    return fail(ABR_ECUDA, std::string("abr_cost_create: ") + cudaGetErrorString(cudaGetLastError()));
  }
  *out = c;
  return ABR_OK;
}

int abr_cost_destroy(AbrCost* c) {
  if (!c) return ABR_OK;
  DeviceGuard guard(c->device);
  if (c->d) cudaFree(c->d);
  delete c;
  return ABR_OK;
}

int abr_model_reserve(AbrModel* m, int nworld, int N, int B, int S) {
  if (!m || nworld < 0 || N < 0 || B < 0 || S < 0) return fail(ABR_EINVAL, "abr_model_reserve: bad argument");
  ABR_ON_DEVICE(m->device);
  const size_t nx = m->lay.nx, nu = m->lay.nu, nxs = (size_t)(N + 1) * nx, nus = (size_t)N * nu;
  int rc = ABR_OK;
  if (B > 0 && S > 0) {  // sampler: per-sample costs, kept trajectories (when under the cap), host staging of one solve
    rc = m->s_costs.ensure(sizeof(float) * (size_t)B * S);
    size_t keep_cap = (size_t)256 << 20;
    if (const char* e = getenv("ABR_KEEP_TRAJ_MB")) keep_cap = (size_t)atol(e) << 20;
    const size_t traj_bytes = sizeof(float) * (size_t)B * S * (nxs + nus);
    if (!rc && traj_bytes <= keep_cap) rc = m->s_traj.ensure(traj_bytes + 16);
    if (!rc) rc = m->s_in.ensure(sizeof(float) * ((size_t)B * nx + (size_t)B * nus * S + 4));
    if (!rc) rc = m->s_out.ensure(sizeof(float) * ((size_t)B * (nxs + nus) + (size_t)B * S + 2 * (size_t)B + 8));
  }
  if (!rc && nworld > 0) {  // host rollouts: controls and states in, trajectories and costs out, the slice carry
    rc = m->s_in.ensure(sizeof(float) * ((size_t)nworld * (nx + nus) + 4));
    if (!rc) rc = m->s_out.ensure(sizeof(float) * ((size_t)nworld * (nxs + 1) + 4));
    if (!rc) rc = m->s_carry.ensure(sizeof(float) * (size_t)nworld * (m->lay.nq + 2 * m->lay.nv + 8));
  }
  return rc;
}

// ---- caller-provided workspace for the stream-ordered (*_dev) calls: SURVEY 8(b)'s "no hidden allocation on the hot path"
// Layout of the workspace: [per-sample costs B*S][kept trajectories B*S*((N+1) nx + N nu), when under the keep cap][mpc winner buffers]
static void workspace_plan(const AbrModel* m, int N, int B, int S, size_t out[3]) {
  const size_t nx = m->lay.nx, nu = m->lay.nu, nxs = (size_t)(N + 1) * nx, nus = (size_t)N * nu;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t keep_cap = (size_t)256 << 20;
  if (const char* e = getenv("ABR_KEEP_TRAJ_MB")) keep_cap = (size_t)atol(e) << 20;
  const size_t traj = sizeof(float) * (size_t)B * S * (nxs + nus);
  out[0] = up(sizeof(float) * (size_t)B * S);
  out[1] = (B > 0 && S > 0 && traj <= keep_cap) ? up(traj + 16) : 0;
  out[2] = up(sizeof(float) * (nxs + nus + 8));
}
int abr_workspace_bytes(const AbrModel* m, int N, int B, int S, size_t* bytes) {
  if (!m || !bytes || N < 0 || B < 0 || S < 0) return fail(ABR_EINVAL, "abr_workspace_bytes: bad argument");
  size_t part[3];
  workspace_plan(m, N, B, S, part);
  *bytes = part[0] + part[1] + part[2];
  return ABR_OK;
}
int abr_model_set_workspace(AbrModel* m, void* workspace, size_t bytes, int N, int B, int S) {
  if (!m || N < 0 || B < 0 || S < 0) return fail(ABR_EINVAL, "abr_model_set_workspace: bad argument");
  if (!workspace) {  // back to handle-owned scratch (grown on first use / by abr_model_reserve)
    m->s_costs.release(); m->s_traj.release(); m->s_mpc.release();
    return ABR_OK;
  }
  size_t part[3];
  workspace_plan(m, N, B, S, part);
  if (bytes < part[0] + part[1] + part[2]) return fail(ABR_ECAPACITY, "abr_model_set_workspace: workspace smaller than abr_workspace_bytes for these sizes");
  if (((uintptr_t)workspace & 15) != 0) return fail(ABR_EINVAL, "abr_model_set_workspace: workspace must be 16-byte aligned");
  char* w = (char*)workspace;
  m->s_costs.adopt(w, part[0]);
  m->s_traj.adopt(w + part[0], part[1]);
  m->s_mpc.adopt(w + part[0] + part[1], part[2]);
  return ABR_OK;
}

int abr_rollout_dev(AbrModel* m, const float* x0, int x0_stride, const float* us, int us_stride, int nworld, int N,
                    float* xs_out, const AbrCost* cost, float* costs_out, void* stream) {
  if (nworld < 0 || N < 0) return fail(ABR_EINVAL, "abr_rollout_dev: negative size");
  if (m && nworld == 0) return ABR_OK;  // an empty batch has no buffers to point at
  if (!m || !x0 || (!us && N > 0 && m && m->lay.nu > 0)) return fail(ABR_EINVAL, "abr_rollout_dev: null argument");
  if (cost && (cost->nx != m->lay.nx || cost->nu != m->lay.nu)) return fail(ABR_EINVAL, "abr_rollout_dev: cost dimensions do not match the model");
  if (costs_out && !cost) return fail(ABR_EINVAL, "abr_rollout_dev: costs_out without a cost");
  ABR_ON_DEVICE(m->device);
  RolloutArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.x0 = x0; a.x0_stride = x0_stride; a.us = us; a.us_stride = us_stride;
  a.mode = 0; a.nworld = nworld; a.N = N; a.xs_out = xs_out; a.costs_out = costs_out;
  a.cost = cost_view(costs_out ? cost : nullptr);
  a.t_begin = 0; a.t_end = N;
  return launch_rollout(m, m->lay, a, (cudaStream_t)stream);
}

int abr_rollout_host(AbrModel* m, const float* x0, int x0_stride, const float* us, int us_stride, int nworld, int N,
                     float* xs_out, const AbrCost* cost, float* costs_out) {
  if (nworld < 0 || N < 0) return fail(ABR_EINVAL, "abr_rollout_host: negative size");
  if (m && nworld == 0) return ABR_OK;
  if (!m || !x0 || (!us && N > 0 && m && m->lay.nu > 0)) return fail(ABR_EINVAL, "abr_rollout_host: null argument");
  if (cost && (cost->nx != m->lay.nx || cost->nu != m->lay.nu)) return fail(ABR_EINVAL, "abr_rollout_host: cost dimensions do not match the model");
  if (costs_out && !cost) return fail(ABR_EINVAL, "abr_rollout_host: costs_out without a cost");
  ABR_ON_DEVICE(m->device);
  const int nx = m->lay.nx, nu = m->lay.nu;
  const size_t n_x0 = x0_stride ? (size_t)nworld * nx : nx;
  const size_t n_us = us_stride ? (size_t)nworld * N * nu : (size_t)N * nu;
  const size_t n_xs = xs_out ? (size_t)nworld * (N + 1) * nx : 0;
  const size_t n_c = costs_out ? nworld : 0;
  int rc = m->s_in.ensure(sizeof(float) * (n_x0 + n_us + 4));
  if (rc) return rc;
  rc = m->s_out.ensure(sizeof(float) * (n_xs + n_c + 4));
  if (rc) return rc;
  float* d_x0 = (float*)m->s_in.p; float* d_us = d_x0 + n_x0;
  float* d_xs = (float*)m->s_out.p; float* d_c = d_xs + n_xs;
  CK(cudaMemcpyAsync(d_x0, x0, sizeof(float) * n_x0, cudaMemcpyHostToDevice, m->stream));
  // Pipelined form (limb kernels, long horizons): the controls go up in horizon slices on a copy stream
  // while the previous slice's steps run; state and running cost pass between the slice launches in `carry`.
  int nslice = 1;
  if (us_stride != 0 && N >= 64 && getenv("ABR_NO_PIPELINE") == nullptr && use_limb(m, m->lay, costs_out && cost && !cost->diag, false)) nslice = 16;
  if (const char* e = getenv("ABR_SLICES")) { const int v = atoi(e); if (v >= 1 && v <= 64 && nslice > 1) nslice = v; }
  if (nslice == 1) {
    if (n_us) CK(cudaMemcpyAsync(d_us, us, sizeof(float) * n_us, cudaMemcpyHostToDevice, m->stream));
    rc = abr_rollout_dev(m, d_x0, x0_stride, d_us, us_stride, nworld, N, xs_out ? d_xs : nullptr, cost, costs_out ? d_c : nullptr, m->stream);
    if (rc) return rc;
  } else {
    if (!m->copy_stream) CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    while ((int)m->ev.size() < nslice + 1) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); m->ev.push_back(e); }
    rc = m->s_carry.ensure(sizeof(float) * (size_t)nworld * (m->lay.nq + 2 * m->lay.nv + 8));
    if (rc) return rc;
    // the copy stream must not overwrite d_us while an earlier call's kernels may still read it
    CK(cudaEventRecord(m->ev[0], m->stream));
    CK(cudaStreamWaitEvent(m->copy_stream, m->ev[0], 0));
    const int per = (N + nslice - 1) / nslice;
    RolloutArgs a;
    memset(&a, 0, sizeof(a));
    a.blob = m->d_blob; a.x0 = d_x0; a.x0_stride = x0_stride; a.us = d_us; a.us_stride = us_stride;
    a.mode = 0; a.nworld = nworld; a.N = N; a.xs_out = xs_out ? d_xs : nullptr; a.costs_out = costs_out ? d_c : nullptr;
    a.cost = cost_view(costs_out ? cost : nullptr); a.carry = (float*)m->s_carry.p;
    const size_t pitch = sizeof(float) * (size_t)N * nu;
    // slice boundaries: the first copy is exposed (nothing to overlap it with), so the schedule starts with a short slice and grows
    // by 1.5x (the next slice's copy must fit under this slice's steps: host-to-device moves a step's controls about twice as fast as
    // the kernel consumes them) until it reaches N / nslice; ABR_SLICES asks for uniform slices
    std::vector<int> bounds{0};
    if (getenv("ABR_SLICES")) {
      for (int tb = per; tb < N; tb += per) bounds.push_back(tb);
    } else {
      for (double len = 8.0; bounds.back() < N;) {
        const int step = std::max(1, std::min(2 * per, (int)len));  // up to N / 8 steps per slice once the ramp is over
        bounds.push_back(std::min(N, bounds.back() + step));
        len *= 1.5;
      }
      bounds.pop_back();
    }
    bounds.push_back(N);
    const int nb = (int)bounds.size() - 1;
    while ((int)m->ev.size() < nb + 2) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); m->ev.push_back(e); }
    for (int k = 0; k < nb; k++) {
      const int tb = bounds[k], te = bounds[k + 1];
      if (tb >= te) break;
      CK(cudaMemcpy2DAsync(d_us + (size_t)tb * nu, pitch, us + (size_t)tb * nu, pitch, sizeof(float) * (size_t)(te - tb) * nu, nworld,
                           cudaMemcpyHostToDevice, m->copy_stream));
      CK(cudaEventRecord(m->ev[k + 1], m->copy_stream));
      CK(cudaStreamWaitEvent(m->stream, m->ev[k + 1], 0));
      a.t_begin = tb; a.t_end = te;
      rc = launch_rollout(m, m->lay, a, m->stream);
      if (rc) return rc;
    }
  }
  if (n_xs) CK(cudaMemcpyAsync(xs_out, d_xs, sizeof(float) * n_xs, cudaMemcpyDeviceToHost, m->stream));
  if (n_c) CK(cudaMemcpyAsync(costs_out, d_c, sizeof(float) * n_c, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return ABR_OK;
}

int abr_predictive_sample_dev(AbrModel* m, const AbrCost* cost, const float* x0, const float* us_guess, const float* noise,
                              unsigned long long seed, int B, int S, int N, float stdev, int sample_offset, int S_total,
                              float* xs_star, float* us_star, int* best_idx, float* best_cost, float* costs_out, void* stream) {
  if (!m || !cost || !x0 || !us_guess || !best_idx || !best_cost) return fail(ABR_EINVAL, "abr_predictive_sample_dev: null argument");
  if (B <= 0 || S <= 0 || N < 0 || sample_offset < 0 || sample_offset + S > S_total) return fail(ABR_EINVAL, "abr_predictive_sample_dev: bad sizes");
  if (cost->nx != m->lay.nx || cost->nu != m->lay.nu) return fail(ABR_EINVAL, "abr_predictive_sample_dev: cost dimensions do not match the model");
  ABR_ON_DEVICE(m->device);
  cudaStream_t st = (cudaStream_t)stream;
  float* d_costs = costs_out;
  if (!d_costs) {
    int rc = m->s_costs.ensure(sizeof(float) * (size_t)B * S);
    if (rc) return rc;
    d_costs = (float*)m->s_costs.p;
  }
  RolloutArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.x0 = x0; a.x0_stride = m->lay.nx; a.us = us_guess; a.us_stride = N * m->lay.nu;
  a.noise = noise; a.seed = seed; a.stdev = stdev; a.mode = 1; a.S = S; a.S_total = S_total; a.sample_offset = sample_offset;
  a.nworld = B * S; a.N = N; a.costs_out = d_costs; a.cost = cost_view(cost);
  a.t_begin = 0; a.t_end = N;
  // Winner trajectories: when every sample's states fit a modest scratch they are kept (a few MB at the
  // 4096 x 32 solve) and the winner is gathered, so the solve is ONE rollout launch deep; larger sweeps
  // re-roll the B winners instead and never materialise the S x (N+1) x nx tensor (shooting.py:152).
  const size_t nxs = (size_t)(N + 1) * m->lay.nx, nus = (size_t)N * m->lay.nu;
  const size_t traj_bytes = sizeof(float) * (size_t)B * S * (nxs + nus);
  size_t keep_cap = (size_t)256 << 20;
  if (const char* e = getenv("ABR_KEEP_TRAJ_MB")) keep_cap = (size_t)atol(e) << 20;
  bool keep = (xs_star || us_star) && traj_bytes <= keep_cap;
  if (keep && m->s_traj.external && m->s_traj.cap < traj_bytes + 16) keep = false;  // caller's workspace has no room: re-roll the winners
  float* xs_all = nullptr; float* us_all = nullptr;
  if (keep) {
    int rc = m->s_traj.ensure(traj_bytes + 16);
    if (rc) return rc;
    xs_all = (float*)m->s_traj.p; us_all = xs_all + (size_t)B * S * nxs;
    a.xs_out = xs_all; a.us_out = us_all;
  }
  int rc = launch_rollout(m, m->lay, a, st);
  if (rc) return rc;
  k_argmin<<<B, 256, 0, st>>>(d_costs, S, sample_offset, best_idx, best_cost);
  CK(cudaGetLastError());
  if (keep) {
    k_gather_winner<<<B, 128, 0, st>>>(xs_all, us_all, best_idx, sample_offset, S, (int)nxs, (int)nus, xs_star, us_star);
    CK(cudaGetLastError());
  } else if (xs_star || us_star) {
    RolloutArgs w = a;
    w.sample_ids = best_idx; w.nworld = B; w.costs_out = nullptr; w.cost.enabled = 0;
    w.xs_out = xs_star; w.us_out = us_star;
    rc = launch_rollout(m, m->lay, w, st);
    if (rc) return rc;
  }
  return ABR_OK;
}

int abr_predictive_sample_host(AbrModel* m, const AbrCost* cost, const float* x0, const float* us_guess, const float* noise,
                               unsigned long long seed, int B, int S, int N, float stdev, int sample_offset, int S_total,
                               float* xs_star, float* us_star, int* best_idx, float* best_cost, float* costs_out) {
  if (!m || !cost || !x0 || !us_guess || !best_idx || !best_cost) return fail(ABR_EINVAL, "abr_predictive_sample_host: null argument");
  if (B <= 0 || S <= 0 || N < 0) return fail(ABR_EINVAL, "abr_predictive_sample_host: bad sizes");
  ABR_ON_DEVICE(m->device);
  const int nx = m->lay.nx, nu = m->lay.nu;
  const size_t n_x0 = (size_t)B * nx, n_ug = (size_t)B * N * nu;
  const size_t n_nz = noise ? (size_t)B * (S_total - 1) * N * nu : 0;
  const size_t n_xs = (size_t)B * (N + 1) * nx, n_us = (size_t)B * N * nu, n_cs = (size_t)B * S;
  int rc = m->s_in.ensure(sizeof(float) * (n_x0 + n_ug + n_nz + 4));
  if (rc) return rc;
  rc = m->s_out.ensure(sizeof(float) * (n_xs + n_us + n_cs + 2 * (size_t)B + 4));
  if (rc) return rc;
  float* d_x0 = (float*)m->s_in.p; float* d_ug = d_x0 + n_x0; float* d_nz = d_ug + n_ug;
  float* d_xs = (float*)m->s_out.p; float* d_us = d_xs + n_xs; float* d_cs = d_us + n_us;
  float* d_bc = d_cs + n_cs; int* d_bi = (int*)(d_bc + B);
  CK(cudaMemcpyAsync(d_x0, x0, sizeof(float) * n_x0, cudaMemcpyHostToDevice, m->stream));
  if (n_ug) CK(cudaMemcpyAsync(d_ug, us_guess, sizeof(float) * n_ug, cudaMemcpyHostToDevice, m->stream));
  if (n_nz) CK(cudaMemcpyAsync(d_nz, noise, sizeof(float) * n_nz, cudaMemcpyHostToDevice, m->stream));
  rc = abr_predictive_sample_dev(m, cost, d_x0, d_ug, noise ? d_nz : nullptr, seed, B, S, N, stdev, sample_offset, S_total,
                                 xs_star ? d_xs : nullptr, us_star ? d_us : nullptr, d_bi, d_bc, d_cs, m->stream);
  if (rc) return rc;
  if (xs_star) CK(cudaMemcpyAsync(xs_star, d_xs, sizeof(float) * n_xs, cudaMemcpyDeviceToHost, m->stream));
  if (us_star && n_us) CK(cudaMemcpyAsync(us_star, d_us, sizeof(float) * n_us, cudaMemcpyDeviceToHost, m->stream));
  if (costs_out) CK(cudaMemcpyAsync(costs_out, d_cs, sizeof(float) * n_cs, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(best_idx, d_bi, sizeof(int) * B, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(best_cost, d_bc, sizeof(float) * B, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return ABR_OK;
}

int abr_mpc_dev(AbrModel* m, const AbrCost* cost, float* x, float* us_guess, unsigned long long seed, int S, int N, float stdev, int nticks,
                float* xs_log, float* us_log, float* cost_log, int* idx_log, void* stream) {
  if (!m || !cost || !x || !us_guess) return fail(ABR_EINVAL, "abr_mpc_dev: null argument");
  if (S <= 0 || N <= 0 || nticks < 0) return fail(ABR_EINVAL, "abr_mpc_dev: bad sizes");
  ABR_ON_DEVICE(m->device);
  const int nx = m->lay.nx, nu = m->lay.nu;
  int rc = m->s_mpc.ensure(sizeof(float) * ((size_t)(N + 1) * nx + (size_t)N * nu + 8));
  if (rc) return rc;
  float* xs_star = (float*)m->s_mpc.p; float* us_star = xs_star + (size_t)(N + 1) * nx;
  float* best_cost = us_star + (size_t)N * nu; int* best_idx = (int*)(best_cost + 1);
  cudaStream_t st = (cudaStream_t)stream;
  for (int tick = 0; tick < nticks; tick++) {
    rc = abr_predictive_sample_dev(m, cost, x, us_guess, nullptr, seed + (unsigned long long)tick, 1, S, N, stdev, 0, S, xs_star, us_star,
                                   best_idx, best_cost, nullptr, stream);
    if (rc) return rc;
    k_mpc_advance<<<1, 128, 0, st>>>(xs_star, us_star, best_idx, best_cost, N, nx, nu, tick, x, us_guess, xs_log, us_log, cost_log, idx_log);
    CK(cudaGetLastError());
  }
  return ABR_OK;
}

int abr_forward_dev(AbrModel* m, float* qpos, float* qvel, const float* ctrl, float* qacc_warmstart, float* qacc, int E, void* stream) {
  if (E < 0) return fail(ABR_EINVAL, "abr_forward_dev: negative size");
  if (m && E == 0) return ABR_OK;
  if (!m || !qpos || !qvel) return fail(ABR_EINVAL, "abr_forward_dev: null argument");
  ABR_ON_DEVICE(m->device);
  EnvArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.qpos = qpos; a.qvel = qvel; a.warm = qacc_warmstart; a.qacc = qacc; a.ctrl = ctrl;
  a.E = E; a.nsubsteps = 0; a.forward_only = 1;
  return launch_env(m, m->lay, a, (cudaStream_t)stream);
}

// mjx.forward / mjx.step with the derived mjx.Data fields an env reads written in the same launch (rl/base.py:98-125)
static int fields_call(AbrModel* m, float* qpos, float* qvel, const float* ctrl, float* warm, float* qacc, float* time, int E, int nsubsteps,
                       bool forward_only, const AbrDataFields* fields, void* stream, const char* who) {
  if (E < 0 || nsubsteps < 0) return fail(ABR_EINVAL, std::string(who) + ": negative size");
  if (m && E == 0) return ABR_OK;
  if (!m || !qpos || !qvel || !fields || (!forward_only && !warm)) return fail(ABR_EINVAL, std::string(who) + ": null argument");
  ABR_ON_DEVICE(m->device);
  EnvArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.qpos = qpos; a.qvel = qvel; a.warm = warm; a.qacc = qacc; a.time = time; a.ctrl = ctrl;
  a.E = E; a.nsubsteps = nsubsteps; a.forward_only = forward_only ? 1 : 0;
  a.fo = *fields; a.fo_on = 1;
  return launch_env(m, m->lay_dbg, a, (cudaStream_t)stream);  // the un-aliased layout: every intermediate survives the step
}
int abr_forward_fields_dev(AbrModel* m, float* qpos, float* qvel, const float* ctrl, float* qacc_warmstart, float* qacc, int E,
                           const AbrDataFields* fields, void* stream) {
  return fields_call(m, qpos, qvel, ctrl, qacc_warmstart, qacc, nullptr, E, 0, true, fields, stream, "abr_forward_fields_dev");
}
int abr_env_step_fields_dev(AbrModel* m, float* qpos, float* qvel, float* qacc_warmstart, float* time, const float* ctrl, int E,
                            int nsubsteps, const AbrDataFields* fields, void* stream) {
  return fields_call(m, qpos, qvel, ctrl, qacc_warmstart, nullptr, time, E, nsubsteps, false, fields, stream, "abr_env_step_fields_dev");
}

int abr_env_step_dev(AbrModel* m, float* qpos, float* qvel, float* qacc_warmstart, float* time, const float* ctrl, int E,
                     int nsubsteps, const unsigned char* reset_mask, const float* first_qpos, const float* first_qvel,
                     const float* first_qacc_warmstart, void* stream) {
  if (E < 0 || nsubsteps < 0) return fail(ABR_EINVAL, "abr_env_step_dev: negative size");
  if (m && E == 0) return ABR_OK;
  if (!m || !qpos || !qvel || !qacc_warmstart) return fail(ABR_EINVAL, "abr_env_step_dev: null argument");
  if (reset_mask && (!first_qpos || !first_qvel)) return fail(ABR_EINVAL, "abr_env_step_dev: reset_mask without first state");
  ABR_ON_DEVICE(m->device);
  EnvArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.qpos = qpos; a.qvel = qvel; a.warm = qacc_warmstart; a.time = time; a.ctrl = ctrl;
  a.reset_mask = reset_mask; a.first_qpos = first_qpos; a.first_qvel = first_qvel; a.first_warm = first_qacc_warmstart;
  a.E = E; a.nsubsteps = nsubsteps; a.forward_only = 0;
  return launch_env(m, m->lay, a, (cudaStream_t)stream);
}

int abr_env_set_randomization_ex(AbrModel* m, const float* dr, int E, int nparam) {
  if (!m || E < 0 || (dr && E == 0) || (nparam != 2 && nparam != 4)) return fail(ABR_EINVAL, "abr_env_set_randomization: bad argument (nparam is 2 or 4)");
  m->dr = dr; m->dr_E = dr ? E : 0; m->dr_n = nparam;
  return ABR_OK;
}
int abr_env_set_randomization(AbrModel* m, const float* dr, int E) { return abr_env_set_randomization_ex(m, dr, E, 2); }

int abr_env_task_step_dev(AbrModel* m, float* qpos, float* qvel, float* qacc_warmstart, float* time, const float* ctrl, int E, int nsubsteps,
                          const float* first_qpos, const float* first_qvel, const float* first_qacc_warmstart, const AbrCost* reward,
                          float z_min, int max_steps, int* steps, float* obs, float* reward_out, unsigned char* done,
                          unsigned char* truncation, void* stream) {
  if (E < 0 || nsubsteps < 1 || max_steps < 0) return fail(ABR_EINVAL, "abr_env_task_step_dev: bad sizes");
  if (m && E == 0) return ABR_OK;
  if (!m || !qpos || !qvel || !qacc_warmstart || !ctrl || !first_qpos || !first_qvel || !reward || !steps || !reward_out || !done)
    return fail(ABR_EINVAL, "abr_env_task_step_dev: null argument");
  if (reward->nx != m->lay.nx || reward->nu != m->lay.nu) return fail(ABR_EINVAL, "abr_env_task_step_dev: reward dimensions do not match the model");
  if (!reward->diag) return fail(ABR_EUNSUPPORTED, "abr_env_task_step_dev: the fused reward takes diagonal Q and R");
  ABR_ON_DEVICE(m->device);
  const CostView cv = cost_view(reward);
  EnvArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.qpos = qpos; a.qvel = qvel; a.warm = qacc_warmstart; a.time = time; a.ctrl = ctrl;
  a.first_qpos = first_qpos; a.first_qvel = first_qvel; a.first_warm = first_qacc_warmstart;
  a.E = E; a.nsubsteps = nsubsteps; a.forward_only = 0;
  a.t_steps = steps; a.t_obs = obs; a.t_reward = reward_out; a.t_done = done; a.t_trunc = truncation;
  a.t_qd = cv.qd; a.t_rd = cv.rd; a.t_xg = cv.xg; a.t_zmin = z_min; a.t_max_steps = max_steps;
  return launch_env(m, m->lay, a, (cudaStream_t)stream);
}

int abr_debug_forward_host(AbrModel* m, const float* qpos, const float* qvel, const float* ctrl, const float* qacc_warmstart,
                           const char* name, float* out, int cap, int* n) {
  if (!m || !qpos || !qvel || !name || !out || !n) return fail(ABR_EINVAL, "abr_debug_forward_host: null argument");
  ABR_ON_DEVICE(m->device);
  const Layout& L = m->lay_dbg;
  const int nq = L.nq, nv = L.nv, nu = L.nu;
  int rc = m->s_dbg.ensure(sizeof(float) * ((size_t)nq + 3 * nv + nu + L.world_stride + 8));
  if (rc) return rc;
  float* d_q = (float*)m->s_dbg.p; float* d_v = d_q + nq; float* d_w = d_v + nv; float* d_a = d_w + nv;
  float* d_c = d_a + nv; float* d_dbg = d_c + nu;
  std::vector<float> zero(std::max(nv, nu) + 1, 0.f);
  CK(cudaMemcpy(d_q, qpos, sizeof(float) * nq, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_v, qvel, sizeof(float) * nv, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, qacc_warmstart ? qacc_warmstart : zero.data(), sizeof(float) * nv, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_c, ctrl ? ctrl : zero.data(), sizeof(float) * nu, cudaMemcpyHostToDevice));
  EnvArgs a;
  memset(&a, 0, sizeof(a));
  a.blob = m->d_blob; a.qpos = d_q; a.qvel = d_v; a.warm = d_w; a.qacc = d_a; a.ctrl = d_c;
  a.E = 1; a.forward_only = 1; a.dbg = d_dbg;
  rc = launch_env(m, L, a, m->stream);
  if (rc) return rc;
  CK(cudaStreamSynchronize(m->stream));
  std::vector<float> W(L.world_stride);
  CK(cudaMemcpy(W.data(), d_dbg, sizeof(float) * L.world_stride, cudaMemcpyDeviceToHost));
  std::vector<float> res;
  const std::string s(name);
  auto span = [&](int off, int cnt) { res.assign(W.begin() + off, W.begin() + off + cnt); };
  auto unpack = [&](int off) {  // packed lower -> dense symmetric
    res.assign((size_t)nv * nv, 0.f);
    for (int i = 0; i < nv; i++) for (int j = 0; j <= i; j++) { res[i * nv + j] = W[off + i * (i + 1) / 2 + j]; res[j * nv + i] = res[i * nv + j]; }
  };
  const int nb = L.nbody;
  if (s == "qpos") span(L.w_qpos, nq);
  else if (s == "xpos") span(L.w_xpos, 3 * nb);
  else if (s == "xquat") span(L.w_xquat, 4 * nb);
  else if (s == "xipos") span(L.w_xipos, 3 * nb);
  else if (s == "xanchor") span(L.w_xanchor, 3 * L.njnt);
  else if (s == "xaxis") span(L.w_xaxis, 3 * L.njnt);
  else if (s == "subtree_com") span(L.w_rootcom, 3 * L.nroot);  // per kinematic root
  else if (s == "cinert") span(L.w_cinert, 10 * nb);
  else if (s == "cdof") span(L.w_cdof, 6 * nv);
  else if (s == "crb") span(L.w_crb, 10 * nb);
  else if (s == "qM") {
    if (L.sparse) {
      res.assign((size_t)nv * nv, 0.f);
      for (int k = 0; k < L.nnz; k++) {
        const int ij = m->mi[L.i_mpair + k];
        res[(ij >> 16) * nv + (ij & 0xffff)] = W[L.w_M + k];
        res[(ij & 0xffff) * nv + (ij >> 16)] = W[L.w_M + k];
      }
    } else {
      unpack(L.w_M);
    }
  }
  else if (s == "cvel") span(L.w_cvel, 6 * nb);
  else if (s == "cdof_dot") span(L.w_cdofdot, 6 * nv);
  else if (s == "contact_dist") span(L.w_cdist, L.ncon);
  else if (s == "contact_pos") span(L.w_cpos, 3 * L.ncon);
  else if (s == "contact_frame") span(L.w_cframe, 9 * L.ncon);
  else if (s == "qfrc_smooth") span(L.w_fs, nv);
  else if (s == "qacc_smooth") span(L.w_as, nv);
  else if (s == "qacc") span(L.w_a, nv);
  else if (s == "qacc_warmstart") span(L.w_warm, nv);
  else if (s == "qfrc_constraint") span(L.w_fc, nv);
  else if (s == "efc_force") span(L.w_force, L.nefc);
  else if (s == "efc_D") span(L.w_D, L.nefc);
  else if (s == "efc_aref") span(L.w_aref, L.nefc);
  else if (s == "efc_J") {
    res.assign((size_t)L.nefc * nv, 0.f);
    const std::vector<int>& mi = m->mi;  // same tables for both layouts
    const std::vector<float>& mf = m->mf;
    for (int r = 0; r < L.nefc; r++) {
      const int info = mi[L.i_row_info + r];
      const int kind = info & 3, idx = (info >> 2) & 0x3ffff, sub = info >> 20;
      if (kind == 0) {
        const int j1 = mi[L.i_eq_j1 + idx], j2 = mi[L.i_eq_j2 + idx];
        if (j2 >= 0) res[r * nv + mi[L.i_jnt_dofadr + j2]] = W[L.w_eqc + idx];
        res[r * nv + mi[L.i_jnt_dofadr + j1]] = 1.f;
      } else if (kind == 1) {
        res[r * nv + mi[L.i_jnt_dofadr + mi[L.i_lim_jnt + idx]]] = W[L.w_lims + idx];
      } else {
        const bool active = W[L.w_D + r] != 0.f;  // MJX zeroes inactive rows
        const float* prm = &mf[L.f_con_prm + kConPrm * mi[L.i_con_pair + idx]];
        for (int d = 0; d < nv && active; d++) {
          float v = W[L.w_B + (3 * idx) * nv + d];
          if (mi[L.i_con_condim + idx] == 3) {
            const float mu = prm[11 + (sub >> 1)];
            v += ((sub & 1) ? -mu : mu) * W[L.w_B + (3 * idx + 1 + (sub >> 1)) * nv + d];
          }
          res[r * nv + d] = v;
        }
      }
    }
  } else {
    return fail(ABR_EINVAL, "abr_debug_forward_host: unknown field " + s);
  }
  *n = (int)res.size();
  if (*n > cap) return fail(ABR_EINVAL, "abr_debug_forward_host: output buffer too small");
  memcpy(out, res.data(), sizeof(float) * res.size());
  return ABR_OK;
}

int abr_ffma_peak(int device, double* tflops, double* ms_out) {
  if (!tflops) return fail(ABR_EINVAL, "abr_ffma_peak: null argument");
  if (abr_device_count() <= 0) return fail(ABR_ENODEVICE, "no CUDA device");
  ABR_ON_DEVICE(device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  float* d = nullptr;
  CK(cudaMalloc(&d, 4));
  const int grid = prop.multiProcessorCount * 2, tpb = 1024, iters = 4096;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; i++) k_ffma<<<grid, tpb>>>(d, iters, 1.0000001f, 1e-9f);
  CK(cudaEventRecord(e0));
  k_ffma<<<grid, tpb>>>(d, iters, 1.0000001f, 1e-9f);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double flops = 2.0 * 8 * 16 * (double)iters * (double)grid * tpb;
  *tflops = flops / (ms * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  return ABR_OK;
}

}  // extern "C"
