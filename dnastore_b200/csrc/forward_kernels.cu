// Forward (sum-product) log-likelihood over the machine-state x DNA-position lattice: SURVEY.md 8a-12.
//
// The reference has no such computation (its only forward-backward is the pair-HMM of
// src/fwdback.cpp, see pairhmm_kernels.cu), so this kernel implements the SPECIFICATION BY ANALOGY
// written down in oracle/forward_oracle.c: ViterbiMatrix's fill (reference src/viterbi.cpp:62-176)
// with every `max` replaced by the reference's table-based log_sum_exp (src/logsumexp.h:19-74), the
// within-column closure solved by synchronous (Jacobi) sweeps that stop after the first sweep that
// changes no cell.  The sweeps and the operand order are part of that specification, so the result
// is bit-identical to the oracle's.  PARITY UNPINNED against the reference.
//
// Mapping: one CTA per read, reads handed out dynamically; the columns live in an L2-resident
// scratch (6+2k doubles per state and CTA), the transition lists are the destination-indexed CSR
// tables of include/dnab_tables.h in the reference's list order.  A first version: correct and
// batched, not yet tuned (no shared-memory columns, no frontier).
#include <cuda_runtime.h>

#include <cstdint>

#include "viterbi_kernels.h"

namespace dnab {
namespace {

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xFFF0000000000000LL); }

__device__ __forceinline__ double lseUnary(const double* __restrict__ table, double x) {  // logsumexp.h:52-74
  if (x >= 10 || isnan(x) || isinf(x)) return 0;
  const int n = (int)(x / .0001);
  const double dx = x - (n * .0001);
  const double f0 = __ldg(table + n), f1 = __ldg(table + n + 1);
  const double df = f1 - f0;
  return f0 + df * (dx / .0001);
}
__device__ __forceinline__ double lse(const double* __restrict__ table, double a, double b) {  // logsumexp.h:34-50
  double mx, diff;
  if (a == b) {
    mx = a;
    diff = 0;
  } else if (a < b) {
    mx = b;
    diff = b - a;
  } else {
    mx = a;
    diff = a - b;
  }
  return mx + lseUnary(table, diff);
}

}  // namespace

__global__ void __launch_bounds__(1024, 1) forwardKernel(const ForwardTables tb, const ForwardArgs args) {
  const uint32_t N = tb.nStates, k = tb.k;
  const uint32_t tid = threadIdx.x, nThreads = blockDim.x;
  const double NEG = ninf();
  const double* L2T = tb.lseTable;
  __shared__ unsigned long long nextSlot;
  __shared__ __align__(16) uint8_t seqS[4096];
  __shared__ double symScore[kMaxSyms];

  for (uint32_t s = tid; s < kMaxSyms; s += nThreads) symScore[s] = s < tb.nSyms ? tb.symScore[s] : NEG;
  double* base = args.scratch + (size_t)blockIdx.x * (6 + 2 * k) * N;
  double* Sprev = base;
  double* S0 = base + N;
  double* Sb[2] = {base + 2 * (size_t)N, base + 3 * (size_t)N};
  double* Db[2] = {base + 4 * (size_t)N, base + 5 * (size_t)N};
  double* Tb[2] = {base + 6 * (size_t)N, base + (6 + k) * (size_t)N};

  for (;;) {
    __syncthreads();
    if (tid == 0) nextSlot = atomicAdd(args.nextRead, 1ull);
    __syncthreads();
    const int64_t read = (int64_t)nextSlot;
    if (read >= args.nReads) break;
    const int32_t L = args.readLen[read];
    {
      const uint8_t* src = args.packed + args.byteOff[read];
      const uint32_t nVec = ((uint32_t)(L + 3) / 4 + 15) / 16;
      for (uint32_t v = tid; v < nVec; v += nThreads)
        reinterpret_cast<uint4*>(seqS)[v] = __ldg(reinterpret_cast<const uint4*>(src) + v);
    }
    for (uint32_t i = tid; i < N * k; i += nThreads) {
      Tb[0][i] = NEG;
      Tb[1][i] = NEG;
    }
    __syncthreads();
    uint32_t cur = 0, tc = 0;
    long long sweepsTotal = 0;
    int status = 0;
    for (int32_t pos = 0; pos <= L; ++pos) {
      const uint32_t tp = tc;
      tc ^= 1;
      const uint32_t x = pos > 0 ? (seqS[(pos - 1) >> 2] >> (2 * ((pos - 1) & 3))) & 3u : 0u;
      // phase 1
      for (uint32_t d = tid; d < N; d += nThreads) {
        double acc = NEG;
        for (uint32_t i = 0; i < k; ++i) Tb[tc][(size_t)i * N + d] = NEG;
        if (pos == 0)
          acc = (tb.local || d == 0) ? 0. : NEG;
        else {
          for (uint32_t e = __ldg(tb.emitOff + d); e < __ldg(tb.emitOff + d + 1); ++e) {
            const uint32_t meta = __ldg(tb.emitMeta + e);
            acc = lse(L2T, acc, ((Sprev[__ldg(tb.emitSrc + e)] + symScore[meta & 31]) + tb.noGap) + tb.sub[(meta >> 5) * 4 + x]);
          }
          const uint32_t mdl = __ldg(tb.mdl + d);
          if (mdl > 0) {
            acc = lse(L2T, acc, Tb[tp][d] + tb.sub[__ldg(tb.ctx + (size_t)d * k) * 4 + x]);
            for (uint32_t i = 0; i + 1 < mdl; ++i)
              Tb[tc][(size_t)i * N + d] = Tb[tp][(size_t)(i + 1) * N + d] + tb.sub[__ldg(tb.ctx + (size_t)d * k + i + 1) * 4 + x];
          }
        }
        S0[d] = acc;
        Sb[0][d] = acc;
        Db[0][d] = NEG;
      }
      __syncthreads();
      // closure: Jacobi sweeps until a sweep changes no cell
      cur = 0;
      for (int sweep = 0;; ++sweep) {
        if (sweep >= args.maxSweeps) {
          status = 1;
          break;
        }
        const double* So = Sb[cur];
        const double* Do = Db[cur];
        double* Sn = Sb[cur ^ 1];
        double* Dn = Db[cur ^ 1];
        int changed = 0;
        for (uint32_t d = tid; d < N; d += nThreads) {
          double nd = NEG, ns = S0[d];
          for (uint32_t e = __ldg(tb.emitOff + d); e < __ldg(tb.emitOff + d + 1); ++e) {
            const uint32_t s = __ldg(tb.emitSrc + e);
            nd = lse(L2T, nd, lse(L2T, Do[s] + tb.delExtend, So[s] + tb.delOpen) + symScore[__ldg(tb.emitMeta + e) & 31]);
          }
          for (uint32_t e = __ldg(tb.nullOff + d); e < __ldg(tb.nullOff + d + 1); ++e) {
            const uint32_t s = __ldg(tb.nullSrc + e);
            const double sc = symScore[__ldg(tb.nullSym + e)];
            nd = lse(L2T, nd, Do[s] + sc);
            ns = lse(L2T, ns, So[s] + sc);
          }
          ns = lse(L2T, ns, nd + tb.delEnd);
          Dn[d] = nd;
          Sn[d] = ns;
          if (__double_as_longlong(nd) != __double_as_longlong(Do[d]) || __double_as_longlong(ns) != __double_as_longlong(So[d]))
            changed = 1;
        }
        cur ^= 1;
        ++sweepsTotal;
        if (!__syncthreads_or(changed)) break;
      }
      // phase 3 and the column hand-over
      for (uint32_t d = tid; d < N; d += nThreads) {
        const double s = Sb[cur][d];
        if (pos > 0) {
          const uint32_t mdl = __ldg(tb.mdl + d);
          for (uint32_t i = 0; i < mdl; ++i)
            Tb[tc][(size_t)i * N + d] = lse(L2T, Tb[tc][(size_t)i * N + d], (s + tb.tanDup) + tb.len[i]);
        }
        Sprev[d] = s;
        if (args.cells && read == 0) {
          double* c = args.cells + ((size_t)pos * N + d) * (k + 2);
          c[0] = s;
          c[1] = Db[cur][d];
          for (uint32_t i = 0; i < k; ++i) c[2 + i] = Tb[tc][(size_t)i * N + d];
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      double ll;
      if (tb.local) {
        ll = NEG;
        for (uint32_t d = 0; d < N; ++d) ll = lse(L2T, ll, Sprev[d]);
      } else
        ll = Sprev[N - 1];
      args.loglike[read] = ll;
      args.sweeps[read] = sweepsTotal;
      args.status[read] = status;
    }
  }
}

cudaError_t launchForward(const ForwardTables& tb, const ForwardArgs& args, uint32_t nBlocks, uint32_t threads,
                          cudaStream_t stream) {
  forwardKernel<<<nBlocks, threads, 0, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
