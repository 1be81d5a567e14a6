// abr_kernels.cuh — the fused rollout / env-step kernels, templated on the lanes-per-world G.
// Each G is instantiated in its own translation unit (abr_k4/8/16/32.cu) so the builds run in
// parallel; abr_engine.cu holds the host side and the small kernels.
#ifndef ABR_KERNELS_CUH_
#define ABR_KERNELS_CUH_

#include <stdint.h>

#include "abr_step.cuh"

#ifndef ABR_TPB
#define ABR_TPB 128
#endif
#ifndef ABR_MINB
#define ABR_MINB 1
#endif

namespace abr {

struct CostView {
  const float* Q; const float* Qf; const float* R; const float* xg;
  const float* qd; const float* qfd; const float* rd;
  int enabled; int diag;
};

struct RolloutArgs {
  const float* blob;
  const float* x0; int x0_stride;          // mode 0: per world; mode 1: per problem
  const float* us; int us_stride;          // mode 0: explicit controls [nworld,N,nu]
  const float* noise;                      // mode 1: nullable [B,S_total-1,N,nu]
  const int* sample_ids;                   // mode 1 winner pass: world = problem, sample = sample_ids[b]
  unsigned long long seed;
  float stdev;
  int mode;                                // 0 = shoot, 1 = predictive sampler
  int S, S_total, sample_offset;
  int nworld, N;
  float* xs_out;                           // nullable [nworld,N+1,nx]
  float* us_out;                           // nullable [nworld,N,nu] (controls actually applied)
  float* costs_out;                        // nullable [nworld]
  CostView cost;
  // horizon slice [t_begin, t_end) of a pipelined rollout (limb kernels): state and the running cost are
  // carried between the slices' launches in `carry` [nworld, nq + 2 nv + 8]; t_begin = 0 starts from x0
  int t_begin, t_end;
  float* carry;
};

__device__ __forceinline__ void stage_blob(const Layout& L, const float* blob, float* smem) {
  const int n = L.n_mf + L.n_mi;
  const float4* src = reinterpret_cast<const float4*>(blob);
  float4* dst = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < n / 4; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// Philox4x32-10 keyed by seed, counter = (sample, index/4, problem, tag); Box-Muller.
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t sample, uint32_t problem, uint32_t index) {
  uint32_t w[4];
  philox4x32(sample, index >> 2, problem, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const int pr = (index >> 1) & 1;
  const float u1 = ((float)w[2 * pr] + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)w[2 * pr + 1] + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.f * logf(u1));
  float s, cs;
  sincospif(2.f * u2, &s, &cs);
  return (index & 1) ? r * s : r * cs;
}

// quadratic cost terms (cost.py:62-85): returns this lane's share of (x-xg)' A (x-xg)
template <int G> __device__ __forceinline__ float quad_x(const Ctx& c, const CostView& cv, const float* x, bool final_) {
  const int nx = c.L.nx;
  float s = 0.f;
  if (cv.diag) {
    const float* qd = final_ ? cv.qfd : cv.qd;
    for (int i = c.lane; i < nx; i += G) { const float e = x[i] - __ldg(cv.xg + i); s += __ldg(qd + i) * e * e; }
  } else {
    const float* A = final_ ? cv.Qf : cv.Q;
    for (int i = c.lane; i < nx; i += G) {
      const float ei = x[i] - __ldg(cv.xg + i);
      float r = 0.f;
      for (int j = 0; j < nx; j++) r += __ldg(A + i * nx + j) * (x[j] - __ldg(cv.xg + j));
      s += ei * r;
    }
  }
  return s;
}
template <int G> __device__ __forceinline__ float quad_u(const Ctx& c, const CostView& cv, const float* u) {
  const int nu = c.L.nu;
  float s = 0.f;
  if (cv.diag) {
    for (int i = c.lane; i < nu; i += G) s += __ldg(cv.rd + i) * u[i] * u[i];
  } else {
    for (int i = c.lane; i < nu; i += G) {
      float r = 0.f;
      for (int j = 0; j < nu; j++) r += __ldg(cv.R + i * nu + j) * u[j];
      s += u[i] * r;
    }
  }
  return s;
}

template <int G>
__global__ void __launch_bounds__(ABR_TPB, ABR_MINB) k_rollout(const __grid_constant__ Layout L, const __grid_constant__ RolloutArgs A) {
  extern __shared__ __align__(16) float smem[];
  stage_blob(L, A.blob, smem);
  const int gpb = blockDim.x / G;
  const int grp = threadIdx.x / G;
  const int wraw = blockIdx.x * gpb + grp;
  const bool valid = wraw < A.nworld;
  const int w = valid ? wraw : A.nworld - 1;
  Ctx c{L, smem, reinterpret_cast<const int*>(smem + L.n_mf), smem + L.n_mf + L.n_mi + grp * L.world_stride, (int)(threadIdx.x % G)};
  const int nq = L.nq, nv = L.nv, nu = L.nu, nx = L.nx, N = A.N;
  init_world<G>(c);
  // problem / sample bookkeeping
  int prob = w, sample = 0;
  if (A.mode == 1) {
    if (A.sample_ids) { prob = w; sample = A.sample_ids[w]; }
    else { prob = w / A.S; sample = A.sample_offset + (w - prob * A.S); }
  }
  const float* x0 = A.x0 + (size_t)(A.mode == 1 ? prob : w) * A.x0_stride;
  float* x = c.W + L.w_qpos;
  for (int i = c.lane; i < nx; i += G) x[i] = x0[i];
  __syncwarp();
  float* xs = A.xs_out ? A.xs_out + (size_t)w * (N + 1) * nx : nullptr;
  if (xs && valid) for (int i = c.lane; i < nx; i += G) xs[i] = x[i];
  float cacc = 0.f;
  if (A.cost.enabled && N > 0) cacc += quad_x<G>(c, A.cost, x, false);
  else if (A.cost.enabled) cacc += quad_x<G>(c, A.cost, x, true);
  // t = -1 is mjx.forward with ctrl = 0, which seeds qacc_warmstart (shooting.py:36)
  float* ctrl = c.W + L.w_ctrl;
  for (int t = -1; t < N; t++) {
    if (t < 0) {
      // ctrl stays 0 (make_data)
    } else if (A.mode == 0) {
      const float* u = A.us + (size_t)w * A.us_stride + (size_t)t * nu;
      for (int i = c.lane; i < nu; i += G) ctrl[i] = u[i];
    } else {
      const float* ug = A.us + (size_t)prob * A.us_stride + (size_t)t * nu;
      for (int i = c.lane; i < nu; i += G) {
        float nz = 0.f;
        if (sample > 0) {
          if (A.noise) nz = A.noise[(((size_t)prob * (A.S_total - 1) + (sample - 1)) * N + t) * nu + i];
          else nz = philox_normal(A.seed, (uint32_t)sample, (uint32_t)prob, (uint32_t)(t * nu + i));
        }
        const float* prm = c.mf + L.f_act_prm + kActPrm * i;
        const float v = ug[i] + nz * A.stdev;
        ctrl[i] = fminf(fmaxf(v, prm[0]), prm[1]);  // clip to actuator_ctrlrange (shooting.py:146-148)
      }
    }
    __syncwarp();
    if (t >= 0) {
      if (A.us_out && valid) for (int i = c.lane; i < nu; i += G) A.us_out[((size_t)w * N + t) * nu + i] = ctrl[i];
      if (A.cost.enabled) cacc += quad_u<G>(c, A.cost, ctrl);
    }
    for (int stage = 0;; stage++) {  // one forward per Euler step, four per RK4 step
      forward<G>(c);
      if (t < 0 || post_forward<G>(c, stage)) break;
    }
    if (t >= 0) {
      if (xs && valid) for (int i = c.lane; i < nx; i += G) xs[(size_t)(t + 1) * nx + i] = x[i];
      if (A.cost.enabled) cacc += quad_x<G>(c, A.cost, x, t == N - 1);
    }
  }
  if (A.costs_out) {
    cacc = gsum<G>(cacc);
    if (valid && c.lane == 0) A.costs_out[w] = 0.5f * cacc;
  }
  (void)nq; (void)nv;
}

struct EnvArgs {
  const float* blob;
  float* qpos; float* qvel; float* warm; float* time; float* qacc;
  const float* ctrl;
  const unsigned char* reset_mask;
  const float* first_qpos; const float* first_qvel; const float* first_warm;
  int E, nsubsteps;
  int forward_only;
  float* dbg;  // nullable: world 0's whole shared-memory region after forward
  // fused task epilogue (abr_env_task_step_dev; t_steps == nullptr = off): obs = (qpos, qvel), reward =
  // -(0.5 (x-xg)' diag(qd) (x-xg) + 0.5 u' diag(rd) u), done = floating-base height below t_zmin or the
  // episode counter reaching t_max_steps; finished envs leave the launch already reset to their first state
  int* t_steps; float* t_obs; float* t_reward; unsigned char* t_done; unsigned char* t_trunc;
  const float* t_qd; const float* t_rd; const float* t_xg;
  float t_zmin; int t_max_steps;
  const float* dr;  // nullable [E,dr_n]: per-env scales of contact friction, actuator strength (, joint damping, joint armature)
  int dr_n;
  AbrDataFields fo; int fo_on;  // derived mjx.Data fields written per world (abr_*_fields_dev; needs the un-aliased layout)
};

template <int G>
__global__ void __launch_bounds__(ABR_TPB, ABR_MINB) k_env(const __grid_constant__ Layout L, const __grid_constant__ EnvArgs A) {
  extern __shared__ __align__(16) float smem[];
  stage_blob(L, A.blob, smem);
  const int gpb = blockDim.x / G;
  const int grp = threadIdx.x / G;
  const int wraw = blockIdx.x * gpb + grp;
  const bool valid = wraw < A.E;
  const int w = valid ? wraw : A.E - 1;
  Ctx c{L, smem, reinterpret_cast<const int*>(smem + L.n_mf), smem + L.n_mf + L.n_mi + grp * L.world_stride, (int)(threadIdx.x % G)};
  const int nq = L.nq, nv = L.nv, nu = L.nu;
  if (A.dr) {
    const float* r = A.dr + (size_t)A.dr_n * w;
    c.frs = r[0]; c.acs = r[1];
    if (A.dr_n >= 4) { c.dps = r[2]; c.ars = r[3]; }
  }
  init_world<G>(c);
  const bool reset = A.reset_mask && A.reset_mask[w];
  const bool was_done = A.t_steps && A.t_done[w];  // AutoResetWrapper zeroes the counter of an env that finished last step
  const float* sq = reset ? A.first_qpos : A.qpos;
  const float* sv = reset ? A.first_qvel : A.qvel;
  const float* sw = reset ? A.first_warm : A.warm;
  for (int i = c.lane; i < nq; i += G) c.W[L.w_qpos + i] = sq[(size_t)w * nq + i];
  for (int i = c.lane; i < nv; i += G) { c.W[L.w_qvel + i] = sv[(size_t)w * nv + i]; c.W[L.w_warm + i] = sw ? sw[(size_t)w * nv + i] : 0.f; }
  for (int i = c.lane; i < nu; i += G) c.W[L.w_ctrl + i] = A.ctrl ? A.ctrl[(size_t)w * nu + i] : 0.f;
  __syncwarp();
  const int nfw = A.forward_only ? 1 : A.nsubsteps;
  for (int s = 0; s < nfw; s++) {
    for (int stage = 0;; stage++) {
      forward<G>(c);
      if (A.forward_only || post_forward<G>(c, stage)) break;
    }
  }
  bool fin = false;
  if (A.t_steps) {  // EpisodeWrapper + AutoResetWrapper semantics around obs / reward / done of the stepped state
    __syncwarp();
    float r = 0.f;
    for (int i = c.lane; i < nq; i += G) { const float e = c.W[L.w_qpos + i] - A.t_xg[i]; r = fmaf(A.t_qd[i] * e, e, r); }
    for (int i = c.lane; i < nv; i += G) { const float e = c.W[L.w_qvel + i] - A.t_xg[nq + i]; r = fmaf(A.t_qd[nq + i] * e, e, r); }
    for (int i = c.lane; i < nu; i += G) { const float u = c.W[L.w_ctrl + i]; r = fmaf(A.t_rd[i] * u, u, r); }
    r = gsum<G>(r);
    const int st = ((reset || was_done) ? 0 : A.t_steps[w]) + 1;
    const bool term = !(c.W[L.w_qpos + 2] >= A.t_zmin);  // a non-finite height terminates too
    const bool trunc = A.t_max_steps > 0 && st >= A.t_max_steps;
    fin = term || trunc;
    __syncwarp();
    if (valid && c.lane == 0) {
      A.t_reward[w] = -0.5f * r; A.t_done[w] = fin ? 1 : 0; A.t_steps[w] = st;
      if (A.t_trunc) A.t_trunc[w] = (trunc && !term) ? 1 : 0;
    }
    if (fin) {
      for (int i = c.lane; i < nq; i += G) c.W[L.w_qpos + i] = A.first_qpos[(size_t)w * nq + i];
      for (int i = c.lane; i < nv; i += G) {
        c.W[L.w_qvel + i] = A.first_qvel[(size_t)w * nv + i];
        c.W[L.w_warm + i] = A.first_warm ? A.first_warm[(size_t)w * nv + i] : 0.f;
      }
    }
    __syncwarp();
    if (valid && A.t_obs) {
      for (int i = c.lane; i < nq; i += G) A.t_obs[(size_t)w * (nq + nv) + i] = c.W[L.w_qpos + i];
      for (int i = c.lane; i < nv; i += G) A.t_obs[(size_t)w * (nq + nv) + nq + i] = c.W[L.w_qvel + i];
    }
  }
  if (valid) {
    for (int i = c.lane; i < nq; i += G) A.qpos[(size_t)w * nq + i] = c.W[L.w_qpos + i];
    for (int i = c.lane; i < nv; i += G) {
      if (!A.forward_only) A.qvel[(size_t)w * nv + i] = c.W[L.w_qvel + i];
      if (A.warm) A.warm[(size_t)w * nv + i] = c.W[L.w_warm + i];
      if (A.qacc) A.qacc[(size_t)w * nv + i] = c.W[L.w_a + i];
    }
    if (A.time && c.lane == 0) {
      const float t0 = reset ? 0.f : A.time[w];
      A.time[w] = fin ? 0.f : (A.forward_only ? t0 : t0 + L.timestep * (float)A.nsubsteps);
    }
    if (A.dbg && wraw == 0) for (int i = c.lane; i < L.world_stride; i += G) A.dbg[i] = c.W[i];
    if (A.fo_on) {
      __syncwarp();
      auto put = [&](float* dst, int off, int n) {
        if (dst) for (int i = c.lane; i < n; i += G) dst[(size_t)w * n + i] = c.W[off + i];
      };
      const int nb = L.nbody;
      put(A.fo.xpos, L.w_xpos, 3 * nb); put(A.fo.xquat, L.w_xquat, 4 * nb); put(A.fo.xipos, L.w_xipos, 3 * nb);
      put(A.fo.xanchor, L.w_xanchor, 3 * L.njnt); put(A.fo.xaxis, L.w_xaxis, 3 * L.njnt); put(A.fo.cinert, L.w_cinert, 10 * nb);
      put(A.fo.cdof, L.w_cdof, 6 * nv); put(A.fo.cvel, L.w_cvel, 6 * nb); put(A.fo.cdof_dot, L.w_cdofdot, 6 * nv);
      put(A.fo.qfrc_smooth, L.w_fs, nv); put(A.fo.qacc_smooth, L.w_as, nv); put(A.fo.qfrc_constraint, L.w_fc, nv);
      put(A.fo.efc_force, L.w_force, L.nefc); put(A.fo.efc_D, L.w_D, L.nefc); put(A.fo.efc_aref, L.w_aref, L.nefc);
      put(A.fo.contact_dist, L.w_cdist, L.ncon); put(A.fo.contact_pos, L.w_cpos, 3 * L.ncon); put(A.fo.contact_frame, L.w_cframe, 9 * L.ncon);
    }
  }
}


// ---- launch plumbing shared by the per-G translation units
struct LaunchCfg { int max_smem; };

inline size_t smem_bytes(const Layout& L, int tpb, int G) {
  return sizeof(float) * ((size_t)L.n_mf + L.n_mi + (size_t)(tpb / G) * L.world_stride);
}

// returns cudaError_t as int (0 = ok), -1000 if the model does not fit in shared memory
template <class Args, class K> int launch_k(K kern, const LaunchCfg& cfg, const Layout& L, const Args& a, int nworld, int G, cudaStream_t st) {
  int tpb = ABR_TPB;
  while (tpb > 32 && tpb > G && smem_bytes(L, tpb, G) > (size_t)cfg.max_smem) tpb /= 2;
  const size_t sm = smem_bytes(L, tpb, G);
  if (sm > (size_t)cfg.max_smem) return -1000;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  const int gpb = tpb / G;
  const int grid = (nworld + gpb - 1) / gpb;
  kern<<<grid, tpb, sm, st>>>(L, a);
  return (int)cudaGetLastError();
}

#define ABR_DECLARE_LAUNCHERS(G)                                                                          \
  int launch_rollout_##G(const LaunchCfg&, const Layout&, const RolloutArgs&, cudaStream_t);              \
  int launch_env_##G(const LaunchCfg&, const Layout&, const EnvArgs&, cudaStream_t);
#define ABR_DEFINE_LAUNCHERS(G)                                                                           \
  int launch_rollout_##G(const LaunchCfg& cfg, const Layout& L, const RolloutArgs& a, cudaStream_t st) {  \
    return launch_k(k_rollout<G>, cfg, L, a, a.nworld, G, st);                                            \
  }                                                                                                       \
  int launch_env_##G(const LaunchCfg& cfg, const Layout& L, const EnvArgs& a, cudaStream_t st) {          \
    return launch_k(k_env<G>, cfg, L, a, a.E, G, st);                                                     \
  }
ABR_DECLARE_LAUNCHERS(4)
ABR_DECLARE_LAUNCHERS(8)
ABR_DECLARE_LAUNCHERS(16)
ABR_DECLARE_LAUNCHERS(32)

}  // namespace abr
#endif
