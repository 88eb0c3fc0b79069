// tscm_p2p.cuh — the two exchange steps of a multi-GPU LM iteration as kernels
// of this library over NVLink peer memory, fused with the reductions that precede them.
//
// With NCCL each exchange is a separate collective kernel (~25-30 us for a 40 KB message
// inside the iteration graph; measured 66 us per iteration at 4 GPUs).  Here every rank owns
// a small mailbox in device memory that its peers map through CUDA IPC.  An exchange is a
// one-shot all-to-all PUSH + local sum in RANK ORDER, so every rank obtains bit-identical
// totals (the LM decision stays in lock-step without a broadcast) and no atomics are
// involved:
//   writer:  payload -> slot [seq & 1][my rank] of EVERY peer's mailbox (posted NVLink
//            stores); __threadfence_system(); flag [seq & 1][my rank] of the peer = seq
//   reader:  spin on its OWN (local) flags (>= seq), then read its own mailbox
// One fence (a round trip) plus a one-way flag store per exchange; a pull design (poll the
// peer's flag, then load the peer's slot) measured ~10 us more per exchange.  The Schur
// exchange (155 CTAs) goes one step further and carries the sequence number inside every
// 16-byte entry (p2p_put16 / p2p_get16): no fence, no separate flag.
// Slots are double-buffered by sequence parity: a rank reaches exchange seq + 2 only after
// every peer has posted seq + 1, which each peer does only after it has finished reading seq.
//   k_reduce_s_p2p  sum of the local Schur partials (as k_reduce_s) + exchange of [S | rhs];
//                   every CTA exchanges its own 32 entries (one 256-byte row per peer, pushed
//                   and received by the warp with the peer's index), so the transfer is
//                   spread over 155 CTAs
//   k_xchg_eval     exchange of the evaluation record (C*107 + 4 sums and one max), followed
//                   by the accept/reject decision (replaces ncclAllReduce x2 + k_decide)
// A wait that does not complete within P2PArgs::timeout_ns (wall clock, %globaltimer) sets
// *err; k_xchg_eval then ends the solve with FAILURE instead of deciding on stale sums and the
// host reports TSCM_ERR_COMM.  Nothing spins forever; the flag is cleared at the start of a run.
#pragma once

#include "tscm_kernels.cuh"

namespace tscm {

constexpr int kP2PMaxRanks = 8;

struct P2PArgs {
  int rank, world;
  double* mb[kP2PMaxRanks];     // mailbox of every rank (mb[rank] is local memory)
  unsigned long long* seq;      // [2] local: completed Schur / evaluation exchanges
  unsigned int* ticket;         // local: CTA ticket of k_reduce_s_p2p
  int* err;                     // local: a wait timed out
  int nA;                       // doubles per Schur slot
  int nctaA;                    // CTAs of k_reduce_s_p2p (one flag each)
  int nB;                       // doubles per evaluation slot (C*kCamRec + kCommExtra + 1)
  unsigned long long timeout_ns;   // a wait gives up after this long (tscm_solver_set_exchange_timeout)
};
// mailbox layout (8-byte words), W = kP2PMaxRanks source ranks:
//   A[2][W][nA] of {value, sequence} pairs (16 bytes) | B[2][W][nB] | flagB[2][W]
__host__ __device__ inline size_t p2p_mailbox_words(int nA, int nctaA, int nB) {
  (void)nctaA;
  return (size_t)2 * kP2PMaxRanks * ((size_t)2 * nA + nB + 1);
}
// slot of source rank `src` in the mailbox of rank `r`
__device__ __forceinline__ double2* p2p_A(const P2PArgs& x, int r, int par, int src) {
  return reinterpret_cast<double2*>(x.mb[r]) + ((size_t)par * kP2PMaxRanks + src) * x.nA;
}
__device__ __forceinline__ double* p2p_B(const P2PArgs& x, int r, int par, int src) {
  return x.mb[r] + (size_t)2 * kP2PMaxRanks * ((size_t)2 * x.nA) + ((size_t)par * kP2PMaxRanks + src) * x.nB;
}
__device__ __forceinline__ unsigned long long* p2p_flagB(const P2PArgs& x, int r, int par, int src) {
  return reinterpret_cast<unsigned long long*>(x.mb[r] + (size_t)2 * kP2PMaxRanks * ((size_t)2 * x.nA + x.nB)) +
         (size_t)par * kP2PMaxRanks + src;
}
// The Schur exchange carries its flag INSIDE the payload: every entry travels as one aligned
// 16-byte store {value, sequence number}, which arrives as a unit (the granule NCCL's LL128
// protocol builds on over NVLink), so the receiver polls the entry itself and no
// __threadfence_system() — a round trip per CTA and peer, 155 CTAs at once: 14 us per exchange
// at 2 GPUs — stands between data and flag.  Sequence numbers never repeat (the counter lives
// as long as the solver, the mailbox starts zeroed), so a stale entry cannot match.
__device__ __forceinline__ void p2p_put16(double2* dst, double v, unsigned long long s) {
  asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(dst), "l"(__double_as_longlong(v)), "l"(s) : "memory");
}
__device__ __forceinline__ bool p2p_get16(const double2* src, unsigned long long s, const P2PArgs& x, double* v) {
  unsigned long long t0 = 0;
  for (unsigned i = 0;; ++i) {
    long long a;
    unsigned long long b;
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
    if (b == s) { *v = __longlong_as_double(a); return true; }
    if (i > 64) {
      __nanosleep(100);
      if ((i & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > x.timeout_ns) break;
      }
    }
  }
  *x.err = 1;
  *v = 0.0;
  return false;
}
__device__ __forceinline__ void p2p_post(unsigned long long* flag, unsigned long long s) {
  *reinterpret_cast<volatile unsigned long long*>(flag) = s;
}
__device__ __forceinline__ unsigned long long p2p_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Spin on a LOCAL flag until a peer has posted `s`.  Gives up after x.timeout_ns of wall time
// (a peer that never entered the solve): sets *x.err, which k_xchg_eval turns into a FAILURE
// termination instead of a decision on stale data, and which the host reports as TSCM_ERR_COMM.
__device__ __forceinline__ bool p2p_wait(const unsigned long long* flag, unsigned long long s, const P2PArgs& x) {
  const volatile unsigned long long* f = flag;
  unsigned long long t0 = 0;
  for (unsigned i = 0;; ++i) {
    if (*f >= s) { __threadfence_system(); return true; }
    if (i > 64) {
      __nanosleep(100);
      if ((i & 1023u) == 0) {
        const unsigned long long now = p2p_now_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > x.timeout_ns) break;
      }
    }
  }
  *x.err = 1;
  return false;
}

__global__ void __launch_bounds__(kReduceThreads)
k_reduce_s_p2p(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, LmOptions opt,
               const double* __restrict__ Spart, const double* __restrict__ rpart, int nblk,
               double* __restrict__ out, AssembleArgs A, P2PArgs x) {
  pdl_entry();
  if (st->done) return;
  __shared__ double s_part[8][33];
  __shared__ double s_mine[32];
  const int e = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  const bool ok = i < P.Q + P.NL;
  const unsigned long long s = x.seq[0] + 1;
  const int par = (int)(s & 1ull);
  // camera term of the finishing thread (added once, after the cross-rank sum)
  int pos = 0;
  double cam = 0.0;
  if (part == 0 && ok) {
    int r, c;
    pos = assemble_position(P, A, i, &r, &c);
    cam = assemble_cam_term(P, st->cur ? ps1 : ps0, opt, A, A.radius_override > 0.0 ? A.radius_override : st->radius, r, c);
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (ok) {
    const double* src = i < P.Q ? Spart + i : rpart + (i - P.Q);
    const size_t stride = i < P.Q ? P.Q : P.NL;
    int b = part;
    for (; b + 24 < nblk; b += 32) {
      s0 += src[(size_t)b * stride];
      s1 += src[(size_t)(b + 8) * stride];
      s2 += src[(size_t)(b + 16) * stride];
      s3 += src[(size_t)(b + 24) * stride];
    }
    for (; b < nblk; b += 8) s0 += src[(size_t)b * stride];
  }
  s_part[part][e] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (part == 0) {
    double t = s_part[0][e];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += s_part[q][e];
    s_mine[e] = ok ? t : 0.0;
  }
  __syncthreads();
  // warp p pushes this CTA's 32 sums to rank p and receives rank p's 32 sums: every lane sends
  // and polls its own {value, sequence} entry
  double v = 0.0;
  if (part < x.world) {
    if (part != x.rank) {
      p2p_put16(p2p_A(x, part, par, x.rank) + blockIdx.x * 32 + e, s_mine[e], s);
      p2p_get16(p2p_A(x, x.rank, par, part) + blockIdx.x * 32 + e, s, x, &v);
    } else {
      v = s_mine[e];
    }
  }
  s_part[part][e] = v;
  __syncthreads();
  if (part == 0 && ok) {
    double t = s_part[0][e];
    for (int q = 1; q < x.world; ++q) t += s_part[q][e];
    out[pos] = t + cam;
  }
  // the last CTA to finish closes the exchange (every CTA has read seq by then)
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int n = atomicAdd(x.ticket, 1u);
    if (n == gridDim.x - 1) { *x.ticket = 0u; x.seq[0] = s; }
  }
}

constexpr int kXchgThreads = 512;
__global__ void __launch_bounds__(kXchgThreads)
k_xchg_eval(DeviceProblem P, ParamSet ps0, ParamSet ps1, LmState* st, int which, LmOptions opt, Trace tr,
            P2PArgs x, int decide) {
  pdl_entry();
  if (which < 2 && st->done) return;
  const int sel = which >= 2 ? which - 2 : (st->cur ^ which);
  const ParamSet& ps = sel ? ps1 : ps0;
  const int tid = threadIdx.x;
  const unsigned long long s = x.seq[1] + 1;
  const int par = (int)(s & 1ull);
  const int n = P.C * kCamRec + kCommExtra;         // sums; entry n is the gradient max-norm
  // push the local record (sums | max) to every rank's mailbox, own included
  for (int i = tid; i <= n; i += kXchgThreads) {
    const double val = i < n ? ps.comm[i] : ps.gmax[0];
#pragma unroll
    for (int p = 0; p < kP2PMaxRanks; ++p)
      if (p < x.world) p2p_B(x, p, par, x.rank)[i] = val;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < x.world) {
    if (tid != x.rank) p2p_post(p2p_flagB(x, tid, par, x.rank), s);
  }
  if (tid < x.world && tid != x.rank) p2p_wait(p2p_flagB(x, x.rank, par, tid), s, x);
  __syncthreads();
  for (int i = tid; i <= n; i += kXchgThreads) {
    double acc = *reinterpret_cast<const volatile double*>(p2p_B(x, x.rank, par, 0) + i);
    for (int p = 1; p < x.world; ++p) {
      const double v = *reinterpret_cast<const volatile double*>(p2p_B(x, x.rank, par, p) + i);
      acc = i < n ? acc + v : fmax(acc, v);
    }
    if (i < n) ps.comm[i] = acc; else ps.gmax[0] = acc;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) x.seq[1] = s;
  const bool failed = *reinterpret_cast<volatile int*>(x.err) != 0;
  if (failed) {
    // a peer did not show up in this or the preceding Schur exchange: the sums are not global
    if (tid == 0) { st->termination = 2; st->done = 1; }
  } else if (decide && !st->done) {
    __shared__ double s_red[40];
    const CameraSummary cs = camera_summary(P, ps, s_red);     // ps: the candidate set (which = 1)
    if (tid == 0) decide_step(P, ps0, ps1, st, opt, tr, cs);
  }
}

}  // namespace tscm
