// tscm_b200.cu — C-ABI implementation (include/tscm.h) and host orchestration
// of the device-side Levenberg-Marquardt loop.  The host never touches a
// residual: it uploads the problem, captures one LM iteration as a CUDA graph,
// replays it, and reads back the summary.  There is no CPU fallback.
//
// Replaces: ceres::Solve at /root/reference/TS.cpp:278 and multi_calib.cpp:216.

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tscm.h"
#include "tscm_kernels.cuh"
#include "tscm_eval5.cuh"
#include "tscm_p2p.cuh"
#include "tscm_schur_pairs.cuh"
#include "tscm_pair_lists.h"
#include "tscm_remap.cuh"

namespace {

thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

#define CUDA_TRY(expr)                                                                \
  do {                                                                                \
    cudaError_t e_ = (expr);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return TSCM_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

// ---------------------------------------------------------------------------
// NCCL through dlopen: the process that hosts us (PyTorch) already carries an
// NCCL; binding at run time avoids a second copy with clashing symbols.
// ---------------------------------------------------------------------------
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* Comm;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0, kNcclMax = 2;

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce &&
             api.GroupStart && api.GroupEnd;
  });
  return api;
}

template <typename T>
int upload(T** dst, const std::vector<T>& src) {
  const size_t bytes = std::max<size_t>(1, src.size()) * sizeof(T);
  CUDA_TRY(cudaMalloc((void**)dst, bytes));
  if (!src.empty()) CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TSCM_OK;
}

}  // namespace

using namespace tscm;

struct tscm_solver {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  tscm_options options{};
  LmOptions lm{};
  DeviceProblem P{};
  ParamSet ps[2]{};
  LmState* d_state = nullptr;
  LmState* h_state = nullptr;   // pinned
  Trace trace{};
  int trace_capacity = 0;
  // owned device buffers
  std::vector<void*> owned;
  double2* d_obs_in = nullptr;   // staging: reference layout [V][K]
  double2* d_obsT = nullptr;
  double* d_scale_e = nullptr;
  double* d_scale_c = nullptr;
  double* d_Spart = nullptr;
  double* d_rpart = nullptr;
  double* d_Sr = nullptr;
  double* d_yc = nullptr;
  double* d_mom = nullptr;       // [ntiles][188][32] moments of the evaluation in flight (k_eval5 -> k_view_blocks)
  double* d_fcg = nullptr;       // [ntiles][27][32] frame constants of the evaluation in flight
  double* d_bs_part = nullptr;
  double* d_gmax_part = nullptr;
  double* d_xn2_part = nullptr;
  unsigned int* d_ticket = nullptr;
  double* d_comm_stage = nullptr;   // NCCL fallback: staged evaluation record (+ gmax)
  double* d_dbg_lhs = nullptr;
  double* d_dbg_rhs = nullptr;
  SchurArgs schur{};
  Schur2Args schur2{};
  SchurSplitArgs split{};
  bool split_ok = false;
  size_t split_smem = 0;
  bool split_frames8 = false;    // dense rows produced by k_pair_frames + k_pair_blocks
  bool pairs_ok = false;         // sparse visibility: per-camera-pair Schur update (tscm_schur_pairs.cuh)
  PairArgs pairs{};
  bool schur2_ok = false;
  int schur2_nt = 0;
  size_t schur2_smem = 0;
  int schur_nblk = 0, schur_nt = 256, schur_ept = 20;
  size_t schur_smem = 0, solve_smem = 0, eval3_smem = 0, eval4_smem = 0, eval5_smem = 0;
  int eval_variant = 5;
  int want_err = 0;   // accumulate sum sqrt(s) (reprojection read-out only)
  int prof = 0;
  int bs_nblk = 0, fg_nblk = 0;
  // graph of one LM iteration
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_dirty = true;
  // comm
  NcclApi::Comm comm = nullptr;
  int rank = 0, num_ranks = 1;
  // NVLink peer-memory exchange (tscm_p2p.cuh): mailbox of this rank, IPC mappings of the peers
  bool p2p_on = false;
  double* d_mailbox = nullptr;
  unsigned long long* d_p2p_seq = nullptr;
  unsigned int* d_p2p_ticket = nullptr;
  int* d_p2p_err = nullptr;
  void* p2p_peer[kP2PMaxRanks] = {};
  P2PArgs p2p{};
  int64_t launches = 0;
  // host copies needed later
  int C = 0, F = 0, K = 0, V = 0;

  template <typename T>
  int alloc(T** p, size_t n) {
    CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(1, n) * sizeof(T)));
    CUDA_TRY(cudaMemsetAsync(*p, 0, std::max<size_t>(1, n) * sizeof(T), stream));
    owned.push_back(*p);
    return TSCM_OK;
  }
  template <typename T>
  int put(const T** p, const std::vector<T>& v) {
    T* d = nullptr;
    int rc = upload(&d, v);
    if (rc) return rc;
    owned.push_back(d);
    *p = d;
    return TSCM_OK;
  }
};

namespace {

void fill_lm_options(const tscm_options& o, LmOptions& lm) {
  lm.max_num_iterations = o.max_num_iterations;
  lm.function_tolerance = o.function_tolerance;
  lm.gradient_tolerance = o.gradient_tolerance;
  lm.parameter_tolerance = o.parameter_tolerance;
  lm.initial_radius = o.initial_trust_region_radius;
  lm.max_radius = o.max_trust_region_radius;
  lm.min_radius = o.min_trust_region_radius;
  lm.min_relative_decrease = o.min_relative_decrease;
  lm.min_lm_diagonal = o.min_lm_diagonal;
  lm.max_lm_diagonal = o.max_lm_diagonal;
  lm.max_num_consecutive_invalid_steps = o.max_num_consecutive_invalid_steps;
  lm.jacobi_scaling = o.jacobi_scaling;
  lm.loss_type = o.loss_type;
  lm.loss_scale = o.loss_scale;
  lm.ptol_needs_success = o.parameter_tolerance_needs_successful_step;
  lm.disable_tolerances = o.disable_tolerances;
}

int validate_problem(const tscm_problem* p) {
  if (!p) { set_error("problem is NULL"); return TSCM_ERR_INVALID_ARGUMENT; }
  if (p->num_cameras <= 0 || p->num_frames <= 0 || p->corners_per_board <= 0 || p->num_views <= 0) {
    set_error("empty problem: C=%d F=%d K=%d views=%d", p->num_cameras, p->num_frames,
              p->corners_per_board, p->num_views);
    return TSCM_ERR_INVALID_ARGUMENT;
  }
  if (!p->board_xy || !p->view_camera || !p->view_frame || !p->obs_xy) {
    set_error("problem has NULL arrays");
    return TSCM_ERR_INVALID_ARGUMENT;
  }
  if (p->fixed_camera < -1 || p->fixed_camera >= p->num_cameras) {
    set_error("fixed_camera %d out of range", p->fixed_camera);
    return TSCM_ERR_INVALID_ARGUMENT;
  }
  std::vector<char> seen(p->num_frames, 0);
  for (int v = 0; v < p->num_views; ++v) {
    const int m = p->view_camera[v], i = p->view_frame[v];
    if (m < 0 || m >= p->num_cameras || i < 0 || i >= p->num_frames) {
      set_error("view %d references camera %d / frame %d out of range", v, m, i);
      return TSCM_ERR_INVALID_ARGUMENT;
    }
    if (v > 0) {
      const int pm = p->view_camera[v - 1], pi = p->view_frame[v - 1];
      if (m < pm || (m == pm && i <= pi)) {
        set_error("views must be ordered camera-major with increasing frames (view %d)", v);
        return TSCM_ERR_INVALID_ARGUMENT;
      }
    }
    seen[i] = 1;
  }
  for (int i = 0; i < p->num_frames; ++i)
    if (!seen[i]) {
      set_error("frame %d is seen by no camera (drop it: multi_calib.cpp:102,167)", i);
      return TSCM_ERR_INVALID_ARGUMENT;
    }
  return TSCM_OK;
}

// The residual + Jacobian + normal-equation kernel: k_eval4 (moment form, default) or
// k_eval3 (rank-1 sweeps of full Jacobian rows; TSCM_EVAL_VARIANT=3, kept for A/B timing).
// part: 0 = the whole pass, 1 = k_eval5 only, 2 = k_view_blocks only (stage timing).
void launch_eval_kernel(tscm_solver* s, int which, int part = 0) {
  const DeviceProblem& P = s->P;
  if (s->eval_variant == 5) {
    const int ntiles = (P.V + 31) / 32;
    if (part != 2)
      k_eval5<<<std::min(ntiles, s->sm_count), kE5Threads, s->eval5_smem, s->stream>>>(
          P, s->ps[0], s->ps[1], s->d_state, which, s->lm, s->want_err, s->d_mom, s->d_fcg);
    if (part != 1)
      k_view_blocks<<<ntiles, kVbThreads, vb_smem_bytes(), s->stream>>>(P, s->ps[0], s->ps[1], s->d_state, which,
                                                                        s->d_mom, s->d_fcg);
    if (part == 0) s->launches += 1;
  } else if (part == 2) {
    return;
  } else if (s->eval_variant == 3)
    k_eval3<<<(P.V + 31) / 32, kE3Threads, s->eval3_smem, s->stream>>>(P, s->ps[0], s->ps[1], s->d_state,
                                                                      which, s->lm, s->prof);
  else
    k_eval4<<<(P.V + 31) / 32, kE3Threads, s->eval4_smem, s->stream>>>(P, s->ps[0], s->ps[1], s->d_state,
                                                                      which, s->lm, s->want_err);
}

// which: 0 = current, 1 = candidate (relative to st->cur); 2/3 = absolute set 0/1.
// prep: run k_prep_cams first (the iteration graph does not: k_solve already wrote the
// candidate's constants).  decide: fold the accept/reject decision into the tail block
// (single GPU; with several GPUs the records are all-reduced first and k_decide follows).
void launch_evaluation(tscm_solver* s, int which, int initial, bool prep = true, bool decide = false) {
  const DeviceProblem& P = s->P;
  cudaStream_t st = s->stream;
  if (prep) {
    k_prep_cams<<<(P.C + 31) / 32, 32, 0, st>>>(P, s->ps[0], s->ps[1], s->d_state, which);
    s->launches += 1;
  }
  launch_eval_kernel(s, which);
  PostArgs a;
  a.gmax_part = s->d_gmax_part; a.xn2_part = s->d_xn2_part;
  a.bs_part = s->d_bs_part; a.bs_nblk = s->bs_nblk; a.fg_nblk = s->fg_nblk;
  a.ticket = s->d_ticket; a.initial = initial; a.decide = decide ? 1 : 0;
  k_post_eval<<<P.C + s->fg_nblk, kPostThreads, 0, st>>>(P, s->ps[0], s->ps[1], s->d_state, which,
                                                            s->lm, s->trace, a);
  s->launches += 2;
}

// Global sums of the evaluation record.  The record lives in ps[sel].comm; the
// selection is only known on the device, so both candidates are reduced when
// `which` is relative (they are tiny: C*107+4 doubles).
int launch_eval_allreduce(tscm_solver* s, int which, bool decide = false) {
  if (s->num_ranks <= 1) return TSCM_OK;
  if (s->p2p_on) {
    k_xchg_eval<<<1, kXchgThreads, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, which, s->lm,
                                                    s->trace, s->p2p, decide ? 1 : 0);
    s->launches += 1;
    return TSCM_OK;
  }
  // The record of the set the device selects is staged through a fixed buffer: only that
  // record may be summed (the current point's is already global).
  NcclApi& n = nccl();
  const int cnt = s->P.C * kCamRec + kCommExtra;
  const int blocks = (cnt + 1 + 255) / 256;
  k_comm_stage<<<blocks, 256, 0, s->stream>>>(s->ps[0], s->ps[1], s->d_state, which, cnt, s->d_comm_stage, 0);
  n.GroupStart();
  int rc = n.AllReduce(s->d_comm_stage, s->d_comm_stage, (size_t)cnt, kNcclFloat64, kNcclSum, s->comm, s->stream);
  if (rc) { set_error("ncclAllReduce(sum) failed: %d", rc); n.GroupEnd(); return TSCM_ERR_COMM; }
  rc = n.AllReduce(s->d_comm_stage + cnt, s->d_comm_stage + cnt, 1, kNcclFloat64, kNcclMax, s->comm, s->stream);
  if (rc) { set_error("ncclAllReduce(max) failed: %d", rc); n.GroupEnd(); return TSCM_ERR_COMM; }
  rc = n.GroupEnd();
  if (rc) { set_error("ncclGroupEnd failed: %d", rc); return TSCM_ERR_COMM; }
  k_comm_stage<<<blocks, 256, 0, s->stream>>>(s->ps[0], s->ps[1], s->d_state, which, cnt, s->d_comm_stage, 1);
  s->launches += 2;
  return TSCM_OK;
}

void launch_schur(tscm_solver* s, double radius_override) {
  SchurArgs a = s->schur;
  a.radius_override = radius_override;
  if (s->split_ok) {
    SchurSplitArgs b = s->split;
    b.a = a;
    if (s->split_frames8) {
      // per-frame chain with 8 lanes per frame, then the rows with one thread per column
      k_pair_frames<<<(s->F + 31) / 32, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm, b);
      k_pair_blocks<<<(s->V * 16 + 255) / 256, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, b);
      s->launches += 1;
    } else {
      k_schur_frames<<<(s->F + 7) / 8, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm, b);
    }
    if (s->schur_ept == 1)
      k_schur_update<1><<<s->schur_nblk, s->schur_nt, s->split_smem, s->stream>>>(s->P, s->d_state, b);
    else
      k_schur_update<2><<<s->schur_nblk, s->schur_nt, s->split_smem, s->stream>>>(s->P, s->d_state, b);
    s->launches += 1;
  } else if (s->pairs_ok) {
    SchurSplitArgs b = s->split;
    b.a = a;
    k_pair_frames<<<(s->F + 31) / 32, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm, b);
    k_pair_blocks<<<(s->V * 16 + 255) / 256, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, b);
    k_schur_pairs2<<<s->sm_count, kPairWarps * 32, kPair2Smem, s->stream>>>(s->d_state, s->pairs);
    k_reduce_pairs<<<s->pairs.npairs, kPairReduceGroups * kPairPart, 0, s->stream>>>(s->P, s->d_state, s->pairs);
    s->launches += 3;
  } else if (s->schur2_ok) {
    Schur2Args b = s->schur2;
    b.a = a;
    k_schur2<<<s->schur_nblk, s->schur2_nt, s->schur2_smem, s->stream>>>(s->P, s->ps[0], s->ps[1],
                                                                        s->d_state, s->lm, b);
  } else if (s->schur_ept == 1)
    k_schur<1, 512><<<s->schur_nblk, s->schur_nt, s->schur_smem, s->stream>>>(s->P, s->ps[0], s->ps[1],
                                                                             s->d_state, s->lm, a);
  else
    k_schur<2, 768><<<s->schur_nblk, s->schur_nt, s->schur_smem, s->stream>>>(s->P, s->ps[0], s->ps[1],
                                                                             s->d_state, s->lm, a);
  const int n = s->P.Q + s->P.NL;
  const int nparts = s->pairs_ok ? 1 : s->schur_nblk;   // k_reduce_pairs leaves one complete partial
  if (s->p2p_on && s->num_ranks > 1)
    k_reduce_s_p2p<<<(n + 31) / 32, kReduceThreads, 0, s->stream>>>(s->P, s->d_state, s->d_Spart, s->d_rpart,
                                                                    nparts, s->d_Sr, s->p2p);
  else
    k_reduce_s<<<(n + 31) / 32, kReduceThreads, 0, s->stream>>>(s->P, s->d_state, s->d_Spart, s->d_rpart,
                                                                nparts, s->d_Sr);
  s->launches += 2;
}

int launch_schur_allreduce(tscm_solver* s) {
  if (s->num_ranks <= 1 || s->p2p_on) return TSCM_OK;   // P2P: exchanged inside k_reduce_s_p2p
  NcclApi& n = nccl();
  int rc = n.AllReduce(s->d_Sr, s->d_Sr, (size_t)s->P.Q + s->P.NL, kNcclFloat64, kNcclSum, s->comm, s->stream);
  if (rc) { set_error("ncclAllReduce(S) failed: %d", rc); return TSCM_ERR_COMM; }
  return TSCM_OK;
}

void launch_solve(tscm_solver* s, double radius_override, bool debug) {
  double* dl = debug ? s->d_dbg_lhs : nullptr;
  double* dr = debug ? s->d_dbg_rhs : nullptr;
  const int bmax = (s->P.NL + 1 + 31) / 32;
#define TSCM_SOLVE(B)                                                                          \
  k_solve<B><<<1, kSolveThreads, s->solve_smem, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, \
      s->lm, s->d_Sr, s->d_scale_c, s->d_yc, radius_override, dl, dr, s->prof)
  if (bmax <= 2) TSCM_SOLVE(2);
  else if (bmax <= 4) TSCM_SOLVE(4);
  else TSCM_SOLVE(7);
#undef TSCM_SOLVE
  s->launches += 1;
}

void launch_backsub(tscm_solver* s) {
  k_backsub<<<s->bs_nblk, kBacksubThreads, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state,
                                                          s->schur, s->d_yc, s->d_bs_part, s->bs_nblk,
                                                          s->split_ok ? s->split.Wg : nullptr);
  s->launches += 1;
}

int launch_iteration(tscm_solver* s) {
  launch_schur(s, 0.0);
  int rc = launch_schur_allreduce(s);
  if (rc) return rc;
  launch_solve(s, 0.0, false);
  launch_backsub(s);
  const bool single = s->num_ranks <= 1;
  launch_evaluation(s, 1, 0, /*prep=*/false, /*decide=*/single);
  if (!single) {
    rc = launch_eval_allreduce(s, 1, /*decide=*/true);
    if (rc) return rc;
    if (!s->p2p_on) {
      k_decide<<<1, 32, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm, s->trace);
      s->launches += 1;
    }
  }
  return TSCM_OK;
}

int launches_per_iteration(const tscm_solver* s) {
  return (s->num_ranks <= 1 ? 6 : (s->p2p_on ? 7 : 9)) + (s->split_ok ? (s->split_frames8 ? 2 : 1) : (s->pairs_ok ? 3 : 0)) + (s->eval_variant == 5 ? 1 : 0);
}

int ensure_graph(tscm_solver* s) {
  if (!s->graph_dirty && s->graph_exec) return TSCM_OK;
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
  const int64_t before = s->launches;
  CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
  int rc = launch_iteration(s);
  cudaError_t e = cudaStreamEndCapture(s->stream, &s->graph);
  s->launches = before;
  if (rc) return rc;
  if (e != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(e)); return TSCM_ERR_CUDA; }
  CUDA_TRY(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
  s->graph_dirty = false;
  return TSCM_OK;
}

// Iteration zero: evaluate the initial point held in parameter set 0.
int run_initial(tscm_solver* s) {
  CUDA_TRY(cudaMemsetAsync(s->d_state, 0, sizeof(LmState), s->stream));
  launch_evaluation(s, 2, 1);
  int rc = launch_eval_allreduce(s, 2);
  if (rc) return rc;
  const int n = s->P.F * 6 + s->P.C * 13;
  k_jacobi_scale<<<(n + 255) / 256, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm,
                                                        s->d_scale_e, s->d_scale_c);
  k_init<<<1, 32, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm, s->trace);
  s->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return TSCM_OK;
}

int fetch_state(tscm_solver* s) {
  CUDA_TRY(cudaMemcpyAsync(s->h_state, s->d_state, sizeof(LmState), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return TSCM_OK;
}

int ensure_trace(tscm_solver* s, int capacity) {
  if (capacity <= s->trace_capacity) return TSCM_OK;
  int rc;
  if ((rc = s->alloc(&s->trace.cost, capacity))) return rc;
  if ((rc = s->alloc(&s->trace.radius, capacity))) return rc;
  if ((rc = s->alloc(&s->trace.gmax, capacity))) return rc;
  if ((rc = s->alloc(&s->trace.step_norm, capacity))) return rc;
  if ((rc = s->alloc(&s->trace.flags, capacity))) return rc;
  s->trace.capacity = capacity;
  s->trace_capacity = capacity;
  s->graph_dirty = true;
  return TSCM_OK;
}

const char* termination_name(int t) {
  return t == TSCM_CONVERGENCE ? "CONVERGENCE" : t == TSCM_NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE";
}

}  // namespace

extern "C" {

const char* tscm_last_error(void) { return g_last_error.c_str(); }
const char* tscm_version(void) { return "tscm-b200 0.1.0 (sm_100a, fp64)"; }

void tscm_options_init(tscm_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->max_num_iterations = 50;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->loss_type = TSCM_LOSS_NONE;
  o->loss_scale = 1.0;
}

int tscm_solver_set_options(tscm_solver* s, const tscm_options* o) {
  if (!s || !o) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  if (o->max_num_iterations < 0 || !(o->initial_trust_region_radius > 0.0) ||
      (o->loss_type != TSCM_LOSS_NONE && !(o->loss_scale > 0.0)) || o->loss_type < 0 ||
      o->loss_type > TSCM_LOSS_CAUCHY) {
    set_error("invalid options");
    return TSCM_ERR_INVALID_ARGUMENT;
  }
  s->options = *o;
  fill_lm_options(*o, s->lm);
  s->graph_dirty = true;
  return TSCM_OK;
}

int tscm_solver_create(const tscm_problem* p, const tscm_options* o, int device,
                       tscm_solver** out) {
  if (!out) { set_error("out is NULL"); return TSCM_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  const auto t_create0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!getenv("TSCM_PROF")) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_create0).count();
    std::fprintf(stderr, "[tscm create] %-28s %8.2f ms\n", what, ms);
  };
  int rc = validate_problem(p);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: the calibration solve has no CPU fallback");
    return TSCM_ERR_NO_DEVICE;
  }
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return TSCM_ERR_NO_DEVICE;
  }
  lap("validate + device query");
  tscm_solver* s = new tscm_solver();
  s->device = device;
  s->sm_count = prop.multiProcessorCount;
  tscm_options defaults;
  tscm_options_init(&defaults);
  if ((rc = tscm_solver_set_options(s, o ? o : &defaults))) { delete s; return rc; }
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed"); delete s; return TSCM_ERR_CUDA;
  }
  const int C = p->num_cameras, F = p->num_frames, K = p->corners_per_board, V = p->num_views;
  s->C = C; s->F = F; s->K = K; s->V = V;
  DeviceProblem& P = s->P;
  P.C = C; P.F = F; P.K = K; P.V = V; P.fixed_camera = p->fixed_camera;
  P.Vpad = (V + 31) / 32 * 32;

  // ---- index tables -------------------------------------------------------
  std::vector<int> live_off(C + 1, 0);
  for (int c = 0; c < C; ++c) live_off[c + 1] = live_off[c] + (c == p->fixed_camera ? 7 : 13);
  const int NL = live_off[C];
  P.NL = NL; P.Q = NL * (NL + 1) / 2;
  std::vector<short> live_cam(NL), live_kk(NL);
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < live_off[c + 1] - live_off[c]; ++k) {
      live_cam[live_off[c] + k] = (short)c;
      live_kk[live_off[c] + k] = (short)(c == p->fixed_camera ? k + 6 : k);
    }
  std::vector<short> q_i(P.Q), q_j(P.Q);
  {
    int q = 0;
    for (int i = 0; i < NL; ++i) for (int j = i; j < NL; ++j) { q_i[q] = (short)i; q_j[q] = (short)j; ++q; }
  }
  std::vector<int> frame_ptr(F + 1, 0), frame_views(V), cam_view_begin(C + 1, 0);
  for (int v = 0; v < V; ++v) { frame_ptr[p->view_frame[v] + 1]++; cam_view_begin[p->view_camera[v] + 1]++; }
  for (int i = 0; i < F; ++i) frame_ptr[i + 1] += frame_ptr[i];
  for (int c = 0; c < C; ++c) cam_view_begin[c + 1] += cam_view_begin[c];
  {
    std::vector<int> fill(frame_ptr.begin(), frame_ptr.end() - 1);
    for (int v = 0; v < V; ++v) frame_views[fill[p->view_frame[v]]++] = v;   // camera order kept
  }
  // camera-partial slots: one per (evaluation CTA of 32 views, distinct camera in it)
  const int nblk_eval = (V + 31) / 32;
  std::vector<int> blk_slot(nblk_eval + 1, 0), cam_slot_begin(C + 1, 0);
  {
    std::vector<int> slot_cam;
    for (int b = 0; b < nblk_eval; ++b) {
      blk_slot[b] = (int)slot_cam.size();
      int last = -1;
      for (int v = 32 * b; v < std::min(V, 32 * b + 32); ++v)
        if (p->view_camera[v] != last) { last = p->view_camera[v]; slot_cam.push_back(last); }
    }
    blk_slot[nblk_eval] = (int)slot_cam.size();
    P.nslot = (int)slot_cam.size();
    for (int c : slot_cam) cam_slot_begin[c + 1]++;
    for (int c = 0; c < C; ++c) cam_slot_begin[c + 1] += cam_slot_begin[c];
    // camera-major views => slots are camera-major too (checked, not assumed)
    for (size_t i = 1; i < slot_cam.size(); ++i)
      if (slot_cam[i] < slot_cam[i - 1]) { set_error("internal: slot table"); delete s; return TSCM_ERR_INVALID_ARGUMENT; }
  }

  if (NL > 216) {
    set_error("reduced system of %d live parameters exceeds this build's limit (216, i.e. 17 cameras)", NL);
    delete s; return TSCM_ERR_UNSUPPORTED;
  }
  // 4x4 tiles of the upper block triangle, row-major
  const int nb = (NL + 3) / 4;
  std::vector<short> tile_bi, tile_bj;
  for (int bi = 0; bi < nb; ++bi) for (int bj = bi; bj < nb; ++bj) { tile_bi.push_back((short)bi); tile_bj.push_back((short)bj); }
  const int ntiles = (int)tile_bi.size();
  lap("index tables (host)");
#define TRY_RC(x) do { rc = (x); if (rc) { tscm_solver_destroy(s); return rc; } } while (0)
  std::vector<double> board(p->board_xy, p->board_xy + 2 * (size_t)K);
  std::vector<int> vcam(p->view_camera, p->view_camera + V), vfrm(p->view_frame, p->view_frame + V);
  TRY_RC(s->put(&P.board_xy, board));
  TRY_RC(s->put(&P.view_camera, vcam));
  TRY_RC(s->put(&P.view_frame, vfrm));
  TRY_RC(s->put(&P.frame_ptr, frame_ptr));
  TRY_RC(s->put(&P.frame_views, frame_views));
  TRY_RC(s->put(&P.cam_view_begin, cam_view_begin));
  TRY_RC(s->put(&P.live_off, live_off));
  TRY_RC(s->put(&P.live_cam, live_cam));
  TRY_RC(s->put(&P.live_kk, live_kk));
  TRY_RC(s->put(&P.q_i, q_i));
  TRY_RC(s->put(&P.q_j, q_j));
  TRY_RC(s->put(&P.blk_slot, blk_slot));
  TRY_RC(s->put(&P.cam_slot_begin, cam_slot_begin));

  lap("table upload");
  // ---- buffers --------------------------------------------------------------
  TRY_RC(s->alloc(&s->d_obs_in, (size_t)V * K));
  TRY_RC(s->alloc(&s->d_obsT, (size_t)P.Vpad * K));
  P.obsT = s->d_obsT;
  for (int k = 0; k < 2; ++k) {
    TRY_RC(s->alloc(&s->ps[k].intr, (size_t)C * 9));
    TRY_RC(s->alloc(&s->ps[k].cam_rt, (size_t)C * 6));
    TRY_RC(s->alloc(&s->ps[k].board_rt, (size_t)F * 6));
    TRY_RC(s->alloc(&s->ps[k].cam, (size_t)C));
    TRY_RC(s->alloc(&s->ps[k].G, (size_t)V * kViewStride));
    TRY_RC(s->alloc(&s->ps[k].cam_part, (size_t)P.nslot * kCamRec));
    TRY_RC(s->alloc(&s->ps[k].comm, (size_t)C * kCamRec + kCommExtra));
    TRY_RC(s->alloc(&s->ps[k].gmax, 1));
  }
  TRY_RC(s->alloc(&s->d_state, 1));
  if (cudaMallocHost((void**)&s->h_state, sizeof(LmState)) != cudaSuccess) {
    set_error("cudaMallocHost failed"); tscm_solver_destroy(s); return TSCM_ERR_CUDA;
  }
  TRY_RC(s->alloc(&s->d_scale_e, (size_t)F * 6));
  TRY_RC(s->alloc(&s->d_scale_c, (size_t)C * 13));
  // Schur configuration: one 4x4 tile per thread up to 512 tiles, two beyond
  s->schur_ept = ntiles <= 512 ? 1 : 2;
  s->schur_nt = std::max(256, ((ntiles + s->schur_ept - 1) / s->schur_ept + 31) / 32 * 32);
  s->schur.ntiles = ntiles;
  s->schur.NLp = nb * 4;
  TRY_RC(s->put(&s->schur.tile_bi, tile_bi));
  TRY_RC(s->put(&s->schur.tile_bj, tile_bj));
  s->schur_nblk = std::max(1, std::min(s->sm_count, (F + kSchurFB - 1) / kSchurFB));
  int fpb = (F + s->schur_nblk - 1) / s->schur_nblk;
  fpb = (fpb + kSchurFB - 1) / kSchurFB * kSchurFB;
  s->schur_nblk = (F + fpb - 1) / fpb;
  s->schur.frames_per_block = fpb;
  s->schur.Fpad = (F + 31) / 32 * 32;
  s->schur_smem = (size_t)(2 * kSchurFB * 6 * s->schur.NLp + kSchurFB * 6 + kSchurFB * 64) * sizeof(double) +
                  kSchurFB * 32 * sizeof(int);
  s->solve_smem = (size_t)((NL + 1) * (NL + 2) / 2 + 4 * NL + 4 + 4 * (NL + 1) + 2 * kSolveThreads + 8 +
                           C * kCamRec) * sizeof(double) +
                  (size_t)2 * NL * sizeof(short) + 16;
  if (const char* pv = getenv("TSCM_PROF")) s->prof = atoi(pv) >= 2 ? 1 : 0;   // 2: in-kernel cycle printf
  s->eval3_smem = (size_t)(2 * kE3Group * kE2Elems * 32 + 2 * K) * sizeof(double) +
                  (size_t)C * sizeof(CamConst);
  {
    const size_t main_d = (size_t)2 * kE3Group * kE4Elems * 32;
    const size_t epi_d = (size_t)kE4Mom * 33 + 32 * 108 + 32 * kViewStride + (size_t)kCamRec * 33;
    const size_t kpad = ((size_t)K + kE3Group - 1) / kE3Group * kE3Group;
    s->eval4_smem = (std::max(main_d, epi_d) + kFcElems * 32 + 5 * kpad + (kpad & 1)) * sizeof(double) +
                    (size_t)C * sizeof(CamConst);
  }
  // k_eval5 (persistent, mbarrier-pipelined) is the default; it needs ~214 KB of shared
  // memory plus the camera table, so very large rigs fall back to k_eval4.
  s->eval5_smem = e5_smem_bytes(K, C);
  if (const char* ev = getenv("TSCM_EVAL_VARIANT")) {
    const int v = atoi(ev);
    s->eval_variant = v == 3 ? 3 : (v == 4 ? 4 : 5);
  }
  if (s->eval_variant == 5 && s->eval5_smem > (size_t)prop.sharedMemPerBlockOptin) s->eval_variant = 4;
  TRY_RC(s->alloc(&s->schur.frame_rec, (size_t)kFrameRec * s->schur.Fpad));
  TRY_RC(s->alloc(&s->d_Spart, (size_t)s->schur_nblk * P.Q));
  TRY_RC(s->alloc(&s->d_rpart, (size_t)s->schur_nblk * NL));
  s->schur.Spart = s->d_Spart; s->schur.rpart = s->d_rpart;
  s->schur.scale_e = s->d_scale_e; s->schur.scale_c = s->d_scale_c;
  TRY_RC(s->alloc(&s->d_Sr, (size_t)P.Q + NL));
  {
    // per-frame column descriptors for the pipelined Schur kernel
    std::vector<int> col_ptr(F + 1, 0), col_src;
    std::vector<short> col_g, col_sidx;
    for (int f = 0; f < F; ++f) {
      for (int q = frame_ptr[f]; q < frame_ptr[f + 1]; ++q) {
        const int v = frame_views[q], m = p->view_camera[v];
        const int n = live_off[m + 1] - live_off[m];
        for (int k = 0; k < n; ++k) {
          col_src.push_back(v * 16 + (n == 13 ? k : k + 6));
          col_g.push_back((short)(live_off[m] + k));
          col_sidx.push_back((short)(m * 13 + (n == 13 ? k : k + 6)));
        }
      }
      col_ptr[f + 1] = (int)col_src.size();
    }
    TRY_RC(s->put(&s->schur2.col_ptr, col_ptr));
    TRY_RC(s->put(&s->schur2.col_src, col_src));
    TRY_RC(s->put(&s->schur2.col_g, col_g));
    TRY_RC(s->put(&s->split.col_sidx, col_sidx));
    const int cons = (ntiles + 31) / 32 * 32;
    s->schur2_nt = kSchurFB * 32 + cons;
    s->schur2_smem = (size_t)(2 * (2 * kSchurFB * 6 * s->schur.NLp + kSchurFB * 6) + kSchurFB * 64) * sizeof(double);
    // split form: materialise W_s, Y, z per frame (dense rows) when the rig's visibility is
    // dense enough for that to be cheaper than staging inside one CTA
    s->split.col_ptr = s->schur2.col_ptr; s->split.col_src = s->schur2.col_src; s->split.col_g = s->schur2.col_g;
    s->split_smem = (size_t)(2 * (2 * kSchurFB * 6 * s->schur.NLp + kSchurFB * 6)) * sizeof(double);
    const double fill = (double)col_src.size() / ((double)F * NL);
    s->split_ok = fill >= 0.4 && s->split_smem <= (size_t)prop.sharedMemPerBlockOptin && V < (1 << 27) &&
                  !getenv("TSCM_SCHUR_FUSED");
    if (s->split_ok) {
      TRY_RC(s->alloc(&s->split.Wg, (size_t)F * 6 * s->schur.NLp));
      TRY_RC(s->alloc(&s->split.Yg, (size_t)F * 6 * s->schur.NLp));
      TRY_RC(s->alloc(&s->split.zg, ((size_t)F + 8) * 6));
      TRY_RC(s->alloc(&s->split.fact, (size_t)F * 32));
      const char* f8 = getenv("TSCM_SPLIT_FRAMES8");
      s->split_frames8 = f8 ? atoi(f8) != 0 : false;   // measured on config 3: -6 us of kernel time, +1 launch: no net gain
    }
    s->schur2_ok = s->schur2_nt <= 640 && s->schur2_smem <= (size_t)prop.sharedMemPerBlockOptin &&
                   V < (1 << 27) && !getenv("TSCM_SCHUR_V1");
    // Sparse visibility with a reduced system too wide for the fused kernel (BASELINE config 4):
    // per-camera-pair update on per-view blocks.  TSCM_SCHUR_PAIRS=1 forces it (parity tests),
    // =0 disables it.
    const char* pe = getenv("TSCM_SCHUR_PAIRS");
    const bool want_pairs = pe ? atoi(pe) != 0 : (!s->split_ok && !s->schur2_ok);
    if (want_pairs && V < (1 << 27) && kPair2Smem <= (size_t)prop.sharedMemPerBlockOptin) {
      static_assert(sizeof(PairEntry) == sizeof(int2) && sizeof(PairRange) == sizeof(int2), "int2 layout");
      const PairLists lists = build_pair_lists(C, F, V, p->view_camera, p->view_frame, kPairChunk);
      std::vector<short> loff(C + 1);
      for (int m = 0; m <= C; ++m) loff[m] = (short)live_off[m];
      const int nitems = (int)lists.item_range.size();
      s->pairs.nitems = nitems;
      s->pairs.npairs = (int)lists.pair_a.size();
      TRY_RC(s->put(reinterpret_cast<const PairEntry**>(&s->pairs.ent), lists.ent));
      TRY_RC(s->put(reinterpret_cast<const PairRange**>(&s->pairs.item_range), lists.item_range));
      TRY_RC(s->put(&s->pairs.pair_item, lists.pair_item));
      TRY_RC(s->put(&s->pairs.pair_items, lists.pair_items));
      TRY_RC(s->put(&s->pairs.pair_a, lists.pair_a));
      TRY_RC(s->put(&s->pairs.pair_b, lists.pair_b));
      TRY_RC(s->put(&s->pairs.live_off, loff));
      TRY_RC(s->alloc(&s->pairs.part, (size_t)std::max(1, nitems) * kPairPart));
      TRY_RC(s->alloc(&s->split.Wv, (size_t)V * 96));
      TRY_RC(s->alloc(&s->split.Yv, (size_t)V * 96));
      TRY_RC(s->alloc(&s->split.fact, (size_t)F * 32));
      s->pairs.Wv = s->split.Wv; s->pairs.Yv = s->split.Yv;
      s->pairs.Sout = s->d_Spart; s->pairs.rout = s->d_rpart;
      s->split_ok = false; s->schur2_ok = false;
      s->pairs_ok = true;
    }
  }
  TRY_RC(s->alloc(&s->d_yc, (size_t)NL));
  s->bs_nblk = (F + kBacksubThreads / 32 - 1) / (kBacksubThreads / 32);
  s->fg_nblk = (F * 6 + kPostThreads - 1) / kPostThreads;
  TRY_RC(s->alloc(&s->d_ticket, 1));
  {
    // mailbox of the peer-memory exchange (tiny; allocated always so that its IPC handle can
    // be exported right after creation)
    const int nwords = P.Q + NL;
    s->p2p.nA = (nwords + 31) / 32 * 32;
    s->p2p.nctaA = (nwords + 31) / 32;
    s->p2p.nB = (C * kCamRec + kCommExtra + 1 + 31) / 32 * 32;
    TRY_RC(s->alloc(&s->d_mailbox, p2p_mailbox_words(s->p2p.nA, s->p2p.nctaA, s->p2p.nB)));
    TRY_RC(s->alloc(&s->d_p2p_seq, 2));
    TRY_RC(s->alloc(&s->d_p2p_ticket, 1));
    TRY_RC(s->alloc(&s->d_p2p_err, 1));
    s->p2p.seq = s->d_p2p_seq; s->p2p.ticket = s->d_p2p_ticket; s->p2p.err = s->d_p2p_err;
  }
  TRY_RC(s->alloc(&s->d_comm_stage, (size_t)C * kCamRec + kCommExtra + 8));
  TRY_RC(s->alloc(&s->d_bs_part, (size_t)4 * s->bs_nblk));
  TRY_RC(s->alloc(&s->d_gmax_part, (size_t)s->fg_nblk));
  TRY_RC(s->alloc(&s->d_xn2_part, (size_t)s->fg_nblk));
  TRY_RC(s->alloc(&s->d_dbg_lhs, (size_t)NL * NL));
  TRY_RC(s->alloc(&s->d_dbg_rhs, (size_t)NL));
  TRY_RC(ensure_trace(s, std::min(s->options.max_num_iterations, 1 << 16) + 2));

  lap("buffers");
  auto set_smem = [&](const void* fn, size_t bytes) -> int {
    if (bytes > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return TSCM_OK;
  };
  if (s->schur_smem > (size_t)prop.sharedMemPerBlockOptin || s->solve_smem > (size_t)prop.sharedMemPerBlockOptin) {
    set_error("problem needs %zu / %zu bytes of shared memory per CTA (limit %zu)", s->schur_smem,
              s->solve_smem, (size_t)prop.sharedMemPerBlockOptin);
    tscm_solver_destroy(s); return TSCM_ERR_UNSUPPORTED;
  }
  if (s->schur2_ok) TRY_RC(set_smem((const void*)k_schur2, s->schur2_smem));
  if (s->pairs_ok) TRY_RC(set_smem((const void*)k_schur_pairs2, kPair2Smem));
  if (s->split_ok) {
    TRY_RC(set_smem((const void*)k_schur_update<1>, s->split_smem));
    TRY_RC(set_smem((const void*)k_schur_update<2>, s->split_smem));
  }
  TRY_RC(set_smem((const void*)k_schur<1, 512>, s->schur_smem));
  TRY_RC(set_smem((const void*)k_schur<2, 768>, s->schur_smem));
  TRY_RC(set_smem((const void*)k_solve<2>, s->solve_smem));
  TRY_RC(set_smem((const void*)k_solve<4>, s->solve_smem));
  TRY_RC(set_smem((const void*)k_solve<7>, s->solve_smem));
  TRY_RC(set_smem((const void*)k_eval3, s->eval3_smem));
  TRY_RC(set_smem((const void*)k_eval4, s->eval4_smem));
  if (s->eval_variant == 5) {
    TRY_RC(set_smem((const void*)k_eval5, s->eval5_smem));
    TRY_RC(set_smem((const void*)k_view_blocks, vb_smem_bytes()));
    const size_t ntiles = ((size_t)V + 31) / 32;
    TRY_RC(s->alloc(&s->d_mom, ntiles * kE5MomEntries * 32));
    TRY_RC(s->alloc(&s->d_fcg, ntiles * kFcElems * 32));
  }

  lap("kernel attributes");
  TRY_RC(tscm_solver_set_observations(s, p->obs_xy));
  lap("observations H2D + transpose");
#undef TRY_RC
  *out = s;
  return TSCM_OK;
}

void tscm_solver_destroy(tscm_solver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->graph) cudaGraphDestroy(s->graph);
  if (s->comm && nccl().ok) nccl().CommDestroy(s->comm);
  for (int r = 0; r < kP2PMaxRanks; ++r)
    if (s->p2p_peer[r]) cudaIpcCloseMemHandle(s->p2p_peer[r]);
  for (void* p : s->owned) cudaFree(p);
  if (s->h_state) cudaFreeHost(s->h_state);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int tscm_solver_set_observations(tscm_solver* s, const double* obs_xy) {
  if (!s || !obs_xy) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaMemcpyAsync(s->d_obs_in, obs_xy, (size_t)s->V * s->K * sizeof(double2),
                           cudaMemcpyHostToDevice, s->stream));
  dim3 grid((s->V + 31) / 32, (s->K + 31) / 32), block(32, 8);
  k_transpose_obs<<<grid, block, 0, s->stream>>>(s->d_obs_in, s->d_obsT, s->V, s->K, s->P.Vpad);
  s->launches += 1;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return TSCM_OK;
}

int tscm_solver_set_parameters(tscm_solver* s, const double* intr, const double* cam_rt,
                               const double* board_rt) {
  if (!s || !intr || !cam_rt || !board_rt) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  // The initial point always goes to parameter set 0; run() resets cur = 0.
  CUDA_TRY(cudaMemcpyAsync(s->ps[0].intr, intr, (size_t)s->C * 9 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(s->ps[0].cam_rt, cam_rt, (size_t)s->C * 6 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(s->ps[0].board_rt, board_rt, (size_t)s->F * 6 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  // the constant block and the b, c intrinsics must also exist in set 1
  CUDA_TRY(cudaMemcpyAsync(s->ps[1].intr, intr, (size_t)s->C * 9 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(s->ps[1].cam_rt, cam_rt, (size_t)s->C * 6 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemsetAsync(s->d_state, 0, sizeof(LmState), s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return TSCM_OK;
}

int tscm_solver_get_parameters(tscm_solver* s, double* intr, double* cam_rt, double* board_rt) {
  if (!s) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = fetch_state(s);
  if (rc) return rc;
  const ParamSet& ps = s->ps[s->h_state->cur & 1];
  if (intr) CUDA_TRY(cudaMemcpyAsync(intr, ps.intr, (size_t)s->C * 9 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (cam_rt) CUDA_TRY(cudaMemcpyAsync(cam_rt, ps.cam_rt, (size_t)s->C * 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (board_rt) CUDA_TRY(cudaMemcpyAsync(board_rt, ps.board_rt, (size_t)s->F * 6 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return TSCM_OK;
}

int tscm_solver_run(tscm_solver* s, tscm_summary* summary) {
  if (!s) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = ensure_trace(s, std::min(s->options.max_num_iterations, 1 << 16) + 2);
  if (rc) return rc;
  // If a previous run left x in set 1, move it to set 0 (the initial point).
  if ((rc = fetch_state(s))) return rc;
  if (s->h_state->cur & 1) {
    CUDA_TRY(cudaMemcpyAsync(s->ps[0].intr, s->ps[1].intr, (size_t)s->C * 9 * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->ps[0].cam_rt, s->ps[1].cam_rt, (size_t)s->C * 6 * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(s->ps[0].board_rt, s->ps[1].board_rt, (size_t)s->F * 6 * sizeof(double), cudaMemcpyDeviceToDevice, s->stream));
  }
  if ((rc = ensure_graph(s))) return rc;
  if ((rc = run_initial(s))) return rc;
  // Replay the iteration graph; poll the device-side `done` flag every few
  // iterations (kernels of a finished solve are no-ops).
  const int batch = 8;
  int launched = 0;
  while (launched < s->options.max_num_iterations) {
    const int n = std::min(batch, s->options.max_num_iterations - launched);
    for (int k = 0; k < n; ++k) CUDA_TRY(cudaGraphLaunch(s->graph_exec, s->stream));
    s->launches += (int64_t)n * launches_per_iteration(s);
    launched += n;
    if ((rc = fetch_state(s))) return rc;
    if (s->h_state->done) break;
  }
  if ((rc = fetch_state(s))) return rc;
  CUDA_TRY(cudaGetLastError());
  if (s->p2p_on) {
    int perr = 0;
    CUDA_TRY(cudaMemcpy(&perr, s->d_p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (perr) { set_error("peer-memory exchange timed out waiting for another rank"); return TSCM_ERR_COMM; }
  }
  const LmState& st = *s->h_state;
  if (summary) {
    summary->termination_type = st.done ? st.termination : TSCM_NO_CONVERGENCE;
    summary->num_iterations = st.recorded;
    summary->num_successful_steps = st.num_successful;
    summary->num_unsuccessful_steps = st.num_unsuccessful;
    summary->initial_cost = st.initial_cost;
    summary->final_cost = st.final_cost;
    summary->final_radius = st.radius;
    const int n = std::min({st.recorded, summary->trace_capacity, s->trace.capacity});
    if (n > 0) {
      if (summary->trace_cost) CUDA_TRY(cudaMemcpy(summary->trace_cost, s->trace.cost, n * sizeof(double), cudaMemcpyDeviceToHost));
      if (summary->trace_radius) CUDA_TRY(cudaMemcpy(summary->trace_radius, s->trace.radius, n * sizeof(double), cudaMemcpyDeviceToHost));
      if (summary->trace_gradient_max_norm) CUDA_TRY(cudaMemcpy(summary->trace_gradient_max_norm, s->trace.gmax, n * sizeof(double), cudaMemcpyDeviceToHost));
      if (summary->trace_step_norm) CUDA_TRY(cudaMemcpy(summary->trace_step_norm, s->trace.step_norm, n * sizeof(double), cudaMemcpyDeviceToHost));
      if (summary->trace_step_flags) CUDA_TRY(cudaMemcpy(summary->trace_step_flags, s->trace.flags, n * sizeof(int), cudaMemcpyDeviceToHost));
    }
  }
  if (s->options.verbose && s->rank == 0) {
    // summary.BriefReport() of TS.cpp:280 / multi_calib.cpp:218
    std::printf("Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s\n",
                st.num_successful + st.num_unsuccessful, st.initial_cost, st.final_cost,
                termination_name(st.done ? st.termination : TSCM_NO_CONVERGENCE));
  }
  return TSCM_OK;
}

int tscm_solve(const tscm_problem* problem, const tscm_options* options, double* intrinsics,
               double* cam_rt, double* board_rt, tscm_summary* summary, int device) {
  const bool prof = getenv("TSCM_PROF") != nullptr;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = now();
  tscm_solver* s = nullptr;
  int rc = tscm_solver_create(problem, options, device, &s);
  if (rc) return rc;
  const auto t1 = now();
  rc = tscm_solver_set_parameters(s, intrinsics, cam_rt, board_rt);
  const auto t2 = now();
  if (!rc) rc = tscm_solver_run(s, summary);
  const auto t3 = now();
  if (!rc) rc = tscm_solver_get_parameters(s, intrinsics, cam_rt, board_rt);
  const auto t4 = now();
  tscm_solver_destroy(s);
  if (prof)
    std::fprintf(stderr, "[tscm_solve] create %.2f  set %.2f  run %.2f  get %.2f  destroy %.2f ms\n",
                 ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()));
  return rc;
}

int tscm_comm_unique_id(void* unique_id_128) {
  NcclApi& n = nccl();
  if (!n.ok) { set_error("libnccl.so.2 could not be loaded"); return TSCM_ERR_COMM; }
  NcclApi::UniqueId id;
  int rc = n.GetUniqueId(&id);
  if (rc) { set_error("ncclGetUniqueId failed: %d", rc); return TSCM_ERR_COMM; }
  std::memcpy(unique_id_128, &id, 128);
  return TSCM_OK;
}

int tscm_solver_attach_comm(tscm_solver* s, int rank, int num_ranks, const void* unique_id_128) {
  if (!s || !unique_id_128 || rank < 0 || rank >= num_ranks) { set_error("bad comm arguments"); return TSCM_ERR_INVALID_ARGUMENT; }
  NcclApi& n = nccl();
  if (!n.ok) { set_error("libnccl.so.2 could not be loaded"); return TSCM_ERR_COMM; }
  CUDA_TRY(cudaSetDevice(s->device));
  NcclApi::UniqueId id;
  std::memcpy(&id, unique_id_128, 128);
  int rc = n.CommInitRank(&s->comm, num_ranks, id, rank);
  if (rc) { set_error("ncclCommInitRank failed: %d", rc); return TSCM_ERR_COMM; }
  s->rank = rank; s->num_ranks = num_ranks;
  s->graph_dirty = true;
  return TSCM_OK;
}

int tscm_solver_p2p_export(tscm_solver* s, void* handle_64) {
  if (!s || !handle_64) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));      // the mailbox is zeroed before anyone maps it
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, s->d_mailbox));
  std::memcpy(handle_64, &h, 64);
  return TSCM_OK;
}

int tscm_solver_p2p_attach(tscm_solver* s, int rank, int num_ranks, const void* handles) {
  if (!s || !handles || rank < 0 || rank >= num_ranks) { set_error("bad p2p arguments"); return TSCM_ERR_INVALID_ARGUMENT; }
  if (num_ranks > kP2PMaxRanks) { set_error("peer-memory exchange supports at most %d ranks", kP2PMaxRanks); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  for (int r = 0; r < num_ranks; ++r) {
    if (r == rank) { s->p2p.mb[r] = s->d_mailbox; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * r, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      return TSCM_ERR_COMM;
    }
    s->p2p_peer[r] = ptr;
    s->p2p.mb[r] = static_cast<double*>(ptr);
  }
  s->p2p.rank = rank; s->p2p.world = num_ranks;
  s->rank = rank; s->num_ranks = num_ranks;
  s->p2p_on = true;
  s->graph_dirty = true;
  return TSCM_OK;
}

int tscm_solver_eval_jacobian(tscm_solver* s, double* residuals, double* jacobian, double* cost) {
  if (!s) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = fetch_state(s);
  if (rc) return rc;
  const int cur = s->h_state->cur & 1;
  const size_t N = (size_t)s->V * s->K;
  double *d_r = nullptr, *d_J = nullptr;
  if (residuals) CUDA_TRY(cudaMalloc((void**)&d_r, N * 2 * sizeof(double)));
  if (jacobian) CUDA_TRY(cudaMalloc((void**)&d_J, N * 42 * sizeof(double)));
  k_prep_cams<<<(s->C + 31) / 32, 32, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, 2 + cur);
  k_eval_rows<<<(unsigned)((N + 127) / 128), 128, 0, s->stream>>>(s->P, s->ps[cur], s->lm, d_r, d_J);
  s->launches += 2;
  if (cost) {
    launch_evaluation(s, 2 + cur, 1);
    rc = launch_eval_allreduce(s, 2 + cur);
    if (rc) return rc;
  }
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaGetLastError());
  if (residuals) { CUDA_TRY(cudaMemcpy(residuals, d_r, N * 2 * sizeof(double), cudaMemcpyDeviceToHost)); cudaFree(d_r); }
  if (jacobian) { CUDA_TRY(cudaMemcpy(jacobian, d_J, N * 42 * sizeof(double), cudaMemcpyDeviceToHost)); cudaFree(d_J); }
  if (cost) {
    std::vector<double> comm((size_t)s->C * kCamRec);
    CUDA_TRY(cudaMemcpy(comm.data(), s->ps[cur].comm, comm.size() * sizeof(double), cudaMemcpyDeviceToHost));
    double c = 0.0;
    for (int m = 0; m < s->C; ++m) c += comm[(size_t)m * kCamRec + kCamCost];
    *cost = c;
  }
  return TSCM_OK;
}

int tscm_solver_reduced_size(const tscm_solver* s) { return s ? s->P.NL : -1; }

int tscm_solver_reduced_system(tscm_solver* s, double radius, double* lhs, double* rhs) {
  if (!s || !(radius > 0.0)) { set_error("bad arguments"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  // Evaluate the point in set `cur`, compute Jacobi scaling there (as iteration 0 does).
  int rc = fetch_state(s);
  if (rc) return rc;
  const int cur = s->h_state->cur & 1;
  CUDA_TRY(cudaMemsetAsync(s->d_state, 0, sizeof(LmState), s->stream));
  if (cur) {
    LmState tmp{}; tmp.cur = 1;
    CUDA_TRY(cudaMemcpyAsync(s->d_state, &tmp, sizeof(LmState), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
  }
  launch_evaluation(s, 2 + cur, 1);
  if ((rc = launch_eval_allreduce(s, 2 + cur))) return rc;
  const int n = s->P.F * 6 + s->P.C * 13;
  k_jacobi_scale<<<(n + 255) / 256, 256, 0, s->stream>>>(s->P, s->ps[0], s->ps[1], s->d_state, s->lm,
                                                        s->d_scale_e, s->d_scale_c);
  launch_schur(s, radius);
  if ((rc = launch_schur_allreduce(s))) return rc;
  launch_solve(s, radius, true);
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaGetLastError());
  const int NL = s->P.NL;
  if (lhs) CUDA_TRY(cudaMemcpy(lhs, s->d_dbg_lhs, (size_t)NL * NL * sizeof(double), cudaMemcpyDeviceToHost));
  if (rhs) CUDA_TRY(cudaMemcpy(rhs, s->d_dbg_rhs, (size_t)NL * sizeof(double), cudaMemcpyDeviceToHost));
  return TSCM_OK;
}

int tscm_solver_reprojection_error(tscm_solver* s, double* per_camera, double* overall, double* rms) {
  if (!s) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  int rc = fetch_state(s);
  if (rc) return rc;
  const int cur = s->h_state->cur & 1;
  // Read-out uses the plain (loss-free) residuals, like multi_calib.cpp:235-283.
  LmOptions saved = s->lm;
  s->lm.loss_type = 0;
  s->want_err = 1;
  launch_evaluation(s, 2 + cur, 1);
  rc = launch_eval_allreduce(s, 2 + cur);
  s->lm = saved;
  s->want_err = 0;
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  CUDA_TRY(cudaGetLastError());
  std::vector<double> comm((size_t)s->C * kCamRec);
  CUDA_TRY(cudaMemcpy(comm.data(), s->ps[cur].comm, comm.size() * sizeof(double), cudaMemcpyDeviceToHost));
  // per-camera observation counts (global when sharded: counts are summed like the records)
  std::vector<int> cvb(s->C + 1);
  CUDA_TRY(cudaMemcpy(cvb.data(), s->P.cam_view_begin, cvb.size() * sizeof(int), cudaMemcpyDeviceToHost));
  double sum = 0.0, cost = 0.0, total = 0.0;
  for (int m = 0; m < s->C; ++m) {
    const double e = comm[(size_t)m * kCamRec + kCamErr];
    const double n = (double)(cvb[m + 1] - cvb[m]) * s->K;
    if (per_camera) per_camera[m] = n > 0 ? e / n : 0.0;
    sum += e; total += n;
    cost += comm[(size_t)m * kCamRec + kCamCost];
  }
  if (overall) *overall = total > 0 ? sum / total : 0.0;
  if (rms) *rms = total > 0 ? std::sqrt(2.0 * cost / total) : 0.0;
  // restore the records of the configured loss for a subsequent run()
  if (saved.loss_type) { launch_evaluation(s, 2 + cur, 1); CUDA_TRY(cudaStreamSynchronize(s->stream)); }
  return TSCM_OK;
}

int tscm_solver_time_stage(tscm_solver* s, int stage, int repeats, double* ms_per_launch) {
  if (!s || repeats <= 0 || !ms_per_launch) { set_error("bad arguments"); return TSCM_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(s->device));
  int rc;
  if ((rc = ensure_graph(s))) return rc;
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  if (stage != 4) {
    // make the device state valid and not `done`
    LmOptions saved = s->lm;
    s->lm.disable_tolerances = 1; s->lm.max_num_iterations = 1 << 30;
    rc = run_initial(s);
    if (!rc) { launch_schur(s, 0.0); rc = launch_schur_allreduce(s); }
    if (!rc) { launch_solve(s, 0.0, false); launch_backsub(s); launch_evaluation(s, 3, 0); }
    if (rc) { s->lm = saved; return rc; }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaEventRecord(e0, s->stream));
    const DeviceProblem& P = s->P;
    for (int r = 0; r < repeats; ++r) {
      switch (stage) {
        case 0: launch_eval_kernel(s, 3); s->launches += 1; break;
        case 1: launch_schur(s, 0.0); break;
        case 2: launch_solve(s, 0.0, false); break;
        case 3: launch_backsub(s); break;
        case 5: launch_evaluation(s, 3, 0); break;
        case 6: launch_eval_kernel(s, 3, 1); s->launches += 1; break;
        case 7: launch_eval_kernel(s, 3, 2); s->launches += 1; break;
        default: s->lm = saved; set_error("unknown stage %d", stage); return TSCM_ERR_INVALID_ARGUMENT;
      }
    }
    CUDA_TRY(cudaEventRecord(e1, s->stream));
    s->lm = saved;
  } else {
    // whole LM iterations: iteration zero, then replay the graph.  The options
    // baked into the graph are the solver's (use disable_tolerances = 1 and
    // max_num_iterations >= repeats for a fixed-iteration measurement).
    if ((rc = run_initial(s))) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaEventRecord(e0, s->stream));
    for (int r = 0; r < repeats; ++r) CUDA_TRY(cudaGraphLaunch(s->graph_exec, s->stream));
    s->launches += (int64_t)repeats * launches_per_iteration(s);
    CUDA_TRY(cudaEventRecord(e1, s->stream));
  }
  CUDA_TRY(cudaEventSynchronize(e1));
  CUDA_TRY(cudaGetLastError());
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_launch = (double)ms / repeats;
  if (stage == 4) {
    // a solve that terminated early turns the remaining launches into no-ops:
    // such a measurement is not a measurement
    if ((rc = fetch_state(s))) return rc;
    if (s->h_state->done || s->h_state->iteration != repeats) {
      set_error("LM loop stopped after %d of %d timed iterations (termination %d)",
                s->h_state->iteration, repeats, s->h_state->termination);
      return TSCM_ERR_UNSUPPORTED;
    }
  }
  return TSCM_OK;
}

// Remap tables: TS.cpp:284-330, rectify.cpp:86-199 as one batched kernel.
int tscm_remap_tables(const tscm_remap_job* jobs, int32_t num_jobs, int32_t map_width,
                      int32_t map_height, float* mapx, float* mapy, int device, double* kernel_ms) {
  if (!jobs || !mapx || !mapy || num_jobs <= 0 || map_width <= 0 || map_height <= 0) {
    set_error("bad remap arguments"); return TSCM_ERR_INVALID_ARGUMENT;
  }
  if (num_jobs > kRemapMaxJobs) {
    set_error("%d remap jobs exceed the per-call limit of %d", num_jobs, kRemapMaxJobs);
    return TSCM_ERR_INVALID_ARGUMENT;
  }
  std::vector<RemapBatch> hb(1);
  RemapBatch& B = hb[0];
  std::memset(&B, 0, sizeof(B));
  int64_t total = 0;
  for (int k = 0; k < num_jobs; ++k) {
    const tscm_remap_job& j = jobs[k];
    if (j.width <= 0 || j.height <= 0 || j.row0 < 0 || j.col0 < 0 ||
        (int64_t)j.row0 + j.height > map_height || (int64_t)j.col0 + j.width > map_width) {
      set_error("remap job %d: block %dx%d at (%d,%d) does not fit the %dx%d maps", k, j.width, j.height,
                j.col0, j.row0, map_width, map_height);
      return TSCM_ERR_INVALID_ARGUMENT;
    }
    if (j.ray_fx == 0.0 || j.ray_fy == 0.0) { set_error("remap job %d: zero ray focal length", k); return TSCM_ERR_INVALID_ARGUMENT; }
    RemapJobDev& d = B.job[k];
    std::memcpy(d.intr, j.intrinsics, sizeof(d.intr));
    std::memcpy(d.M, j.matrix, sizeof(d.M));
    d.ray_fx = j.ray_fx; d.ray_fy = j.ray_fy; d.ray_cx = j.ray_cx; d.ray_cy = j.ray_cy;
    d.offset_x = j.offset_x; d.offset_y = j.offset_y; d.cutoff_w2 = j.cutoff_w2;
    d.width = j.width; d.height = j.height; d.row0 = j.row0; d.col0 = j.col0;
    d.first_pixel = total;
    total += (int64_t)j.width * j.height;
  }
  B.num_jobs = num_jobs; B.map_width = map_width; B.total_pixels = total;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: remap tables have no CPU fallback");
    return TSCM_ERR_NO_DEVICE;
  }
  if (device >= 0) CUDA_TRY(cudaSetDevice(device));
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    return TSCM_ERR_NO_DEVICE;
  }
  const size_t map_bytes = (size_t)map_width * map_height * sizeof(float);
  RemapBatch* d_batch = nullptr;
  float *d_x = nullptr, *d_y = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = TSCM_OK;
  auto cleanup = [&]() {
    cudaFree(d_batch); cudaFree(d_x); cudaFree(d_y);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  };
#define REMAP_TRY(expr)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      cleanup(); return TSCM_ERR_CUDA;                                                    \
    }                                                                                     \
  } while (0)
  REMAP_TRY(cudaMalloc((void**)&d_batch, sizeof(RemapBatch)));
  REMAP_TRY(cudaMalloc((void**)&d_x, map_bytes));
  REMAP_TRY(cudaMalloc((void**)&d_y, map_bytes));
  REMAP_TRY(cudaMemcpy(d_batch, &B, sizeof(RemapBatch), cudaMemcpyHostToDevice));
  REMAP_TRY(cudaEventCreate(&e0));
  REMAP_TRY(cudaEventCreate(&e1));
  // one wave of resident CTAs (8 x 256 threads per SM), grid-stride over the pixels
  const int64_t want = (total + 255) / 256;
  const int grid = (int)std::min<int64_t>(want, (int64_t)prop.multiProcessorCount * 8);
  REMAP_TRY(cudaEventRecord(e0, 0));
  k_remap_tables<<<grid, 256>>>(d_batch, d_x, d_y);
  REMAP_TRY(cudaEventRecord(e1, 0));
  REMAP_TRY(cudaGetLastError());
  for (int k = 0; k < num_jobs && rc == TSCM_OK; ++k) {   // only the blocks that were written
    const tscm_remap_job& j = jobs[k];
    const size_t o = (size_t)j.row0 * map_width + j.col0, pitch = (size_t)map_width * sizeof(float);
    REMAP_TRY(cudaMemcpy2D(mapx + o, pitch, d_x + o, pitch, (size_t)j.width * sizeof(float), j.height, cudaMemcpyDeviceToHost));
    REMAP_TRY(cudaMemcpy2D(mapy + o, pitch, d_y + o, pitch, (size_t)j.width * sizeof(float), j.height, cudaMemcpyDeviceToHost));
  }
  REMAP_TRY(cudaEventSynchronize(e1));
  if (kernel_ms) {
    float ms = 0.f;
    REMAP_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *kernel_ms = ms;
  }
#undef REMAP_TRY
  cleanup();
  return rc;
}

int64_t tscm_solver_launch_count(const tscm_solver* s) { return s ? s->launches : 0; }

int tscm_device_fp64_peak(int device, double* tflops) {
  if (!tflops) { set_error("NULL argument"); return TSCM_ERR_INVALID_ARGUMENT; }
  if (device >= 0) CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  double* d_out = nullptr;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  CUDA_TRY(cudaMalloc((void**)&d_out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    k_dfma_peak<<<blocks, threads>>>(d_out, iters, 1.0000001);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
  CUDA_TRY(cudaGetLastError());
  *tflops = best;
  return TSCM_OK;
}

}  // extern "C"
