// abr_xchg.cu — the sampler's one cross-GPU exchange as a kernel over NVLink peer memory ("fused variant").
//
// VanillaPredictiveSampler.optimize sharded over R GPUs of one box (one process per GPU) ends with every rank
// holding its local winner {cost, global sample id, us*[N,nu], xs*[N+1,nx]} (a few KB). Instead of an NCCL
// all-gather followed by a dozen small selection kernels, ONE launch per rank
//   1. stores the local record into slot `rank` of every peer's exchange buffer (P2P st.global over NVLink /
//      NVSwitch, one CTA per destination so the R pushes overlap),
//   2. publishes it with a system-scope release store of the call's epoch into the peer's flag word,
//   3. waits (acquire loads, bounded spin) until all R records of this epoch have landed in its own buffer,
//   4. takes the first minimum (NaN counts as minimal, the lowest global sample id wins ties: jnp.argmin over
//      the concatenated costs, shooting.py:154) and copies the winner out.
// Every rank ends with the identical (xs*, us*, idx, cost). Records are double-buffered by epoch parity, so a
// rank that races ahead into the next solve cannot overwrite a record a slower peer is still reading (it cannot
// reach the solve after that one before the slow peer has published its next epoch).
// The buffers are plain cudaMalloc memory shared between the processes with CUDA IPC handles (exchanged by the
// host framework: torch.distributed all_gather_object in ambersim_b200/parallel.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "abr.h"

namespace abr {
int set_error(int code, const std::string& msg);  // abr_engine.cu (thread-local message behind abr_last_error)
}
using abr::set_error;

#define XCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return set_error(ABR_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace {
constexpr int kMaxRanks = 8;
constexpr int kFlagWords = 64;  // [2 parities][kMaxRanks] epochs + status, padded to 256 B

struct XchgArgs {
  float* peer[kMaxRanks];
  int R, rank, parity;
  unsigned epoch;
  int B, nxs, nus;
  size_t slot_floats;
  const float* cost; const int* idx; const float* xs; const float* us;
  float* xs_out; float* us_out; int* idx_out; float* cost_out;
  long long spin_cycles;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float* slot(float* base, int parity, int R, int r, size_t slot_floats) {
  return base + kFlagWords + ((size_t)parity * R + r) * slot_floats;
}

// grid = R CTAs: CTA d pushes this rank's record to rank d; CTA 0 then waits for all records and selects
__global__ void __launch_bounds__(256) k_xchg_merge(const XchgArgs A) {
  const int d = blockIdx.x;
  const int rec = 2 + A.nus + A.nxs;
  float* dst = slot(A.peer[d], A.parity, A.R, A.rank, A.slot_floats);
  for (int b = 0; b < A.B; b++) {
    float* o = dst + (size_t)b * rec;
    if (threadIdx.x == 0) { o[0] = A.cost[b]; o[1] = __int_as_float(A.idx[b]); }
    for (int i = threadIdx.x; i < A.nus; i += blockDim.x) o[2 + i] = A.us[(size_t)b * A.nus + i];
    for (int i = threadIdx.x; i < A.nxs; i += blockDim.x) o[2 + A.nus + i] = A.xs[(size_t)b * A.nxs + i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned*>(A.peer[d]) + A.parity * kMaxRanks + A.rank, A.epoch);
  if (d != 0) return;
  // ---- consumer: wait for the R records of this epoch in OUR buffer
  float* own = A.peer[A.rank];
  unsigned* flags = reinterpret_cast<unsigned*>(own) + A.parity * kMaxRanks;
  __shared__ int s_timeout;
  if (threadIdx.x == 0) s_timeout = 0;
  __syncthreads();
  if (threadIdx.x < A.R) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flags + threadIdx.x) - A.epoch) < 0) {
      if (clock64() - t0 > A.spin_cycles) { s_timeout = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (s_timeout) {  // a peer never arrived: report instead of hanging. The outputs carry a sentinel (cost +inf, id -1, zeros) so
                    // that a caller who does not poll abr_xchg_timed_out can never mistake stale memory for a winner
    if (threadIdx.x == 0) reinterpret_cast<unsigned*>(own)[2 * kMaxRanks] = A.epoch;
    for (int b = threadIdx.x; b < A.B; b += blockDim.x) { A.cost_out[b] = INFINITY; A.idx_out[b] = -1; }
    if (A.us_out) for (size_t i = threadIdx.x; i < (size_t)A.B * A.nus; i += blockDim.x) A.us_out[i] = 0.f;
    if (A.xs_out) for (size_t i = threadIdx.x; i < (size_t)A.B * A.nxs; i += blockDim.x) A.xs_out[i] = 0.f;
    return;
  }
  for (int b = 0; b < A.B; b++) {
    int win = 0; float wkey = 0.f; int widx = 0;
    for (int r = 0; r < A.R; r++) {  // every thread takes the same decision (R <= 8)
      const volatile float* rp = slot(own, A.parity, A.R, r, A.slot_floats) + (size_t)b * rec;
      const float c = rp[0];
      const float key = (c != c) ? -INFINITY : c;
      const int id = __float_as_int(rp[1]);
      if (r == 0 || key < wkey || (key == wkey && id < widx)) { win = r; wkey = key; widx = id; }
    }
    const volatile float* rp = slot(own, A.parity, A.R, win, A.slot_floats) + (size_t)b * rec;
    if (threadIdx.x == 0) { A.cost_out[b] = rp[0]; A.idx_out[b] = __float_as_int(rp[1]); }
    if (A.us_out) for (int i = threadIdx.x; i < A.nus; i += blockDim.x) A.us_out[(size_t)b * A.nus + i] = rp[2 + i];
    if (A.xs_out) for (int i = threadIdx.x; i < A.nxs; i += blockDim.x) A.xs_out[(size_t)b * A.nxs + i] = rp[2 + A.nus + i];
  }
}
}  // namespace

struct AbrXchg {
  int device = 0, R = 1, rank = 0;
  size_t slot_floats = 0, bytes = 0;
  float* own = nullptr;
  float* peer[kMaxRanks] = {};
  bool opened[kMaxRanks] = {};
  bool connected = false;
  unsigned epoch = 0;
  bool poisoned = false;  // a call timed out: the epoch-parity double buffering no longer protects the records
};

extern "C" {

int abr_xchg_create(int device, int nranks, int rank, size_t max_record_floats, AbrXchg** out, unsigned char* handle_out) {
  if (!out || !handle_out) return set_error(ABR_EINVAL, "abr_xchg_create: null argument");
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || max_record_floats == 0)
    return set_error(ABR_EINVAL, "abr_xchg_create: bad sizes (1..8 ranks of one box)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(ABR_ENODEVICE, "abr_xchg_create: no CUDA device");
  int dev_before = -1;
  cudaGetDevice(&dev_before);
  XCK(cudaSetDevice(device));
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{dev_before};
  AbrXchg* x = new AbrXchg();
  x->device = device; x->R = nranks; x->rank = rank;
  x->slot_floats = (max_record_floats + 31) / 32 * 32;
  x->bytes = sizeof(float) * (kFlagWords + 2 * (size_t)nranks * x->slot_floats);
  cudaError_t e = cudaMalloc(&x->own, x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->own, 0, x->bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, x->own);
  if (e != cudaSuccess) {
    if (x->own) cudaFree(x->own);
    delete x;
    return set_error(ABR_ECUDA, std::string("abr_xchg_create: ") + cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(cudaUUID_t) == ABR_XCHG_HANDLE_BYTES, "handle size");
  memcpy(handle_out, &h, sizeof(h));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    cudaFree(x->own);
    delete x;
    return set_error(ABR_ECUDA, std::string("abr_xchg_create: ") + cudaGetErrorString(e));
  }
  memcpy(handle_out + sizeof(h), &prop.uuid, sizeof(cudaUUID_t));  // lets abr_xchg_connect refuse two ranks on one GPU
  x->peer[rank] = x->own;
  *out = x;
  return ABR_OK;
}

int abr_xchg_connect(AbrXchg* x, const unsigned char* handles) {
  if (!x || !handles) return set_error(ABR_EINVAL, "abr_xchg_connect: null argument");
  // The kernel spin-waits for its peers' records: two ranks sharing one GPU could wait on each other forever (their kernels
  // need not be co-resident). One process per GPU is the contract; refuse anything else.
  for (int a = 0; a < x->R; a++)
    for (int b = a + 1; b < x->R; b++)
      if (memcmp(handles + (size_t)a * ABR_XCHG_HANDLE_BYTES + sizeof(cudaIpcMemHandle_t),
                 handles + (size_t)b * ABR_XCHG_HANDLE_BYTES + sizeof(cudaIpcMemHandle_t), sizeof(cudaUUID_t)) == 0)
        return set_error(ABR_EINVAL, "abr_xchg_connect: ranks " + std::to_string(a) + " and " + std::to_string(b) + " are on the same GPU (one rank per GPU)");
  int dev_before = -1;
  cudaGetDevice(&dev_before);
  XCK(cudaSetDevice(x->device));
  for (int r = 0; r < x->R; r++) {
    if (r == x->rank || x->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * ABR_XCHG_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    XCK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    x->peer[r] = (float*)p; x->opened[r] = true;
  }
  x->connected = true;
  if (dev_before >= 0) cudaSetDevice(dev_before);
  return ABR_OK;
}

int abr_xchg_merge_best_dev(AbrXchg* x, const float* best_cost, const int* best_idx, const float* xs_star, const float* us_star, int B,
                            int nxs, int nus, float* xs_out, float* us_out, int* idx_out, float* cost_out, void* stream) {
  if (!x || !best_cost || !best_idx || !idx_out || !cost_out) return set_error(ABR_EINVAL, "abr_xchg_merge_best_dev: null argument");
  if (B <= 0 || nxs < 0 || nus < 0 || (nxs > 0 && !xs_star) || (nus > 0 && !us_star)) return set_error(ABR_EINVAL, "abr_xchg_merge_best_dev: bad sizes");
  if (!x->connected && x->R > 1) return set_error(ABR_EINVAL, "abr_xchg_merge_best_dev: abr_xchg_connect has not been called");
  if ((size_t)B * (2 + nus + nxs) > x->slot_floats) return set_error(ABR_ECAPACITY, "abr_xchg_merge_best_dev: records exceed the capacity given at create");
  if (x->poisoned) return set_error(ABR_EINVAL, "abr_xchg_merge_best_dev: an earlier call timed out; destroy and re-create the exchange on every rank");
  int dev_before = -1;
  cudaGetDevice(&dev_before);
  XCK(cudaSetDevice(x->device));
  x->epoch += 1;
  XchgArgs a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < x->R; r++) a.peer[r] = x->peer[r];
  a.R = x->R; a.rank = x->rank; a.parity = (int)(x->epoch & 1u); a.epoch = x->epoch;
  a.B = B; a.nxs = nxs; a.nus = nus; a.slot_floats = x->slot_floats;
  a.cost = best_cost; a.idx = best_idx; a.xs = xs_star; a.us = us_star;
  a.xs_out = xs_out; a.us_out = us_out; a.idx_out = idx_out; a.cost_out = cost_out;
  a.spin_cycles = 4000000000LL;  // about 2 s at 1.9 GHz: a peer that never arrives is reported, not waited for forever
  k_xchg_merge<<<x->R, 256, 0, (cudaStream_t)stream>>>(a);
  const cudaError_t le = cudaGetLastError();
  if (dev_before >= 0) cudaSetDevice(dev_before);
  if (le != cudaSuccess) return set_error(ABR_ECUDA, std::string("k_xchg_merge: ") + cudaGetErrorString(le));
  return ABR_OK;
}

int abr_xchg_timed_out(AbrXchg* x, int* timed_out) {
  if (!x || !timed_out) return set_error(ABR_EINVAL, "abr_xchg_timed_out: null argument");
  int dev_before = -1;
  cudaGetDevice(&dev_before);
  XCK(cudaSetDevice(x->device));
  unsigned v = 0;
  XCK(cudaMemcpy(&v, reinterpret_cast<unsigned*>(x->own) + 2 * kMaxRanks, sizeof(v), cudaMemcpyDeviceToHost));
  if (v != 0) {  // read-and-clear: the status word reports each timeout once; the handle refuses further exchanges
    x->poisoned = true;
    XCK(cudaMemset(reinterpret_cast<unsigned*>(x->own) + 2 * kMaxRanks, 0, sizeof(unsigned)));
  }
  if (dev_before >= 0) cudaSetDevice(dev_before);
  *timed_out = (v != 0 || x->poisoned) ? 1 : 0;
  return ABR_OK;
}

int abr_xchg_destroy(AbrXchg* x) {
  if (!x) return ABR_OK;
  cudaSetDevice(x->device);
  for (int r = 0; r < x->R; r++)
    if (x->opened[r] && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->own) cudaFree(x->own);
  delete x;
  return ABR_OK;
}

}  // extern "C"
