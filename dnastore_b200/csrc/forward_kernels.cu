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
// With ForwardArgs::F set the forward cells of every column are kept, and the kernel goes on with the BACKWARD
// pass and the posterior expected counts of the error-model events (the machine-lattice analogue of
// FwdBackMatrix::counts, reference src/fwdback.cpp:154-188), again exactly as specified in
// oracle/forward_oracle.c (dnab_oracle_backward_counts).
//
// Mapping: one CTA per read, reads handed out dynamically; the columns live in an L2-resident
// scratch (6+2k doubles per state and CTA), the transition lists are the destination-indexed CSR
// tables of include/dnab_tables.h in the reference's list order.  A first version: correct and
// batched, not yet tuned (no shared-memory columns, no frontier).
#include <cuda_runtime.h>

#include <cstdint>

#include "lse_table.cuh"
#include "viterbi_kernels.h"

namespace dnab {
namespace {

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xFFF0000000000000LL); }

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
  double* base = args.scratch + (size_t)blockIdx.x * (9 + 2 * k) * N;
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
        if (args.F) {
          double* c = args.F + ((size_t)blockIdx.x * (args.maxLen + 1) + pos) * (size_t)N * (k + 2) + (size_t)d * (k + 2);
          c[0] = s;
          c[1] = Db[cur][d];
          for (uint32_t i = 0; i < k; ++i) c[2 + i] = Tb[tc][(size_t)i * N + d];
        }
        if (args.cells && read == 0) {
          double* c = args.cells + ((size_t)pos * N + d) * (k + 2);
          c[0] = s;
          c[1] = Db[cur][d];
          for (uint32_t i = 0; i < k; ++i) c[2 + i] = Tb[tc][(size_t)i * N + d];
        }
      }
      __syncthreads();
    }
    __shared__ double llShared;
    if (tid == 0) {
      double ll;
      if (tb.local) {
        ll = NEG;
        for (uint32_t d = 0; d < N; ++d) ll = lse(L2T, ll, Sprev[d]);
      } else
        ll = Sprev[N - 1];
      args.loglike[read] = ll;
      args.sweeps[read] = sweepsTotal;
      llShared = ll;
    }
    __syncthreads();
    if (args.F && args.counts) {
      // ---- backward pass + posterior counts (oracle/forward_oracle.c: dnab_oracle_backward_counts) ----
      const double ll = llShared;
      const uint32_t W = k + 2;
      const double* F = args.F + (size_t)blockIdx.x * (args.maxLen + 1) * (size_t)N * W;
      double* Bn = base;                      // column pos+1, [N][W]
      double* Bc = base + (size_t)N * W;      // column pos
      double* bbase = base + 2 * (size_t)N * W;
      double* S2[2] = {bbase + N, bbase + 2 * (size_t)N};
      double* D2[2] = {bbase + 3 * (size_t)N, bbase + 4 * (size_t)N};
      constexpr int kMaxCounts = 5 + kMaxK + 16;
      double cnt[kMaxCounts];
#pragma unroll
      for (int i = 0; i < kMaxCounts; ++i) cnt[i] = 0.;
      auto post = [&](double f, double w, double b) -> double {
        return (f == NEG || b == NEG || w == NEG) ? 0. : exp(f + w + b - ll);
      };
      long long sweepsBack = 0;
      for (int32_t pos = L; pos >= 0; --pos) {
        const uint32_t xn = pos < L ? (seqS[pos >> 2] >> (2 * (pos & 3))) & 3u : 0u;
        for (uint32_t s = tid; s < N; s += nThreads) {
          const uint32_t mdl = __ldg(tb.mdl + s);
          double* bc = Bc + (size_t)s * W;
          for (uint32_t i = 0; i < k; ++i) bc[2 + i] = NEG;
          double b = NEG;
          if (pos == L)
            b = (tb.local || s == N - 1) ? 0. : NEG;
          else {
            const double* bn = Bn + (size_t)s * W;
            if (mdl > 0) {
              bc[2] = tb.sub[__ldg(tb.ctx + (size_t)s * k) * 4 + xn] + bn[0];
              for (uint32_t i = 0; i + 1 < mdl; ++i) bc[2 + i + 1] = tb.sub[__ldg(tb.ctx + (size_t)s * k + i + 1) * 4 + xn] + bn[2 + i];
            }
            for (uint32_t p = __ldg(tb.outEmitOff + s); p < __ldg(tb.outEmitOff + s + 1); ++p) {
              const uint32_t meta = __ldg(tb.outEmitMeta + p);
              b = lse(L2T, b, ((symScore[meta & 31] + tb.noGap) + tb.sub[(meta >> 5) * 4 + xn]) + Bn[(size_t)__ldg(tb.outEmitDst + p) * W]);
            }
          }
          if (pos > 0)
            for (uint32_t i = 0; i < mdl; ++i) b = lse(L2T, b, (tb.tanDup + tb.len[i]) + bc[2 + i]);
          bbase[s] = b;
          S2[0][s] = b;
          D2[0][s] = NEG;
        }
        __syncthreads();
        uint32_t cb = 0;
        for (int sweep = 0;; ++sweep) {
          if (sweep >= args.maxSweeps) {
            status = 1;
            break;
          }
          const double* So = S2[cb];
          const double* Do = D2[cb];
          double* Sn = S2[cb ^ 1];
          double* Dn = D2[cb ^ 1];
          int changed = 0;
          for (uint32_t s = tid; s < N; s += nThreads) {
            double ns = bbase[s], nd = NEG;
            for (uint32_t p = __ldg(tb.outEmitOff + s); p < __ldg(tb.outEmitOff + s + 1); ++p) {
              const double sc = symScore[__ldg(tb.outEmitMeta + p) & 31], bd = Do[__ldg(tb.outEmitDst + p)];
              ns = lse(L2T, ns, (tb.delOpen + sc) + bd);
              nd = lse(L2T, nd, (tb.delExtend + sc) + bd);
            }
            for (uint32_t p = __ldg(tb.outNullOff + s); p < __ldg(tb.outNullOff + s + 1); ++p) {
              const double sc = symScore[__ldg(tb.outNullSym + p)];
              const uint32_t d = __ldg(tb.outNullDst + p);
              ns = lse(L2T, ns, sc + So[d]);
              nd = lse(L2T, nd, sc + Do[d]);
            }
            nd = lse(L2T, nd, tb.delEnd + ns);
            Sn[s] = ns;
            Dn[s] = nd;
            if (__double_as_longlong(ns) != __double_as_longlong(So[s]) || __double_as_longlong(nd) != __double_as_longlong(Do[s]))
              changed = 1;
          }
          cb ^= 1;
          ++sweepsBack;
          if (!__syncthreads_or(changed)) break;
        }
        for (uint32_t s = tid; s < N; s += nThreads) {
          Bc[(size_t)s * W] = S2[cb][s];
          Bc[(size_t)s * W + 1] = D2[cb][s];
        }
        __syncthreads();
        // posterior usage of the moves leaving column pos
        for (uint32_t s = tid; s < N; s += nThreads) {
          const double* fc = F + ((size_t)pos * N + s) * W;
          const double fS = fc[0], fD = fc[1];
          const uint32_t mdl = __ldg(tb.mdl + s);
          const double* bc = Bc + (size_t)s * W;
          for (uint32_t p = __ldg(tb.outEmitOff + s); p < __ldg(tb.outEmitOff + s + 1); ++p) {
            const uint32_t meta = __ldg(tb.outEmitMeta + p), d = __ldg(tb.outEmitDst + p);
            const double sc = symScore[meta & 31];
            cnt[0] += post(fS, tb.delOpen + sc, Bc[(size_t)d * W + 1]);
            cnt[3] += post(fD, tb.delExtend + sc, Bc[(size_t)d * W + 1]);
            if (pos < L) {
              const uint32_t bs = meta >> 5;
              const double u = post(fS, (sc + tb.noGap) + tb.sub[bs * 4 + xn], Bn[(size_t)d * W]);
              cnt[2] += u;
              cnt[5 + k + bs * 4 + xn] += u;
            }
          }
          cnt[4] += post(fD, tb.delEnd, bc[0]);
          if (pos > 0)
            for (uint32_t i = 0; i < mdl; ++i) {
              const double u = post(fS, tb.tanDup + tb.len[i], bc[2 + i]);
              cnt[1] += u;
              cnt[5 + i] += u;
            }
          if (pos < L && mdl > 0) {
            const double* bn = Bn + (size_t)s * W;
            const uint32_t c0 = __ldg(tb.ctx + (size_t)s * k);
            cnt[5 + k + c0 * 4 + xn] += post(fc[2], tb.sub[c0 * 4 + xn], bn[0]);
            for (uint32_t i = 0; i + 1 < mdl; ++i) {
              const uint32_t ci = __ldg(tb.ctx + (size_t)s * k + i + 1);
              cnt[5 + k + ci * 4 + xn] += post(fc[2 + i + 1], tb.sub[ci * 4 + xn], bn[2 + i]);
            }
          }
        }
        __syncthreads();
        double* tmp = Bn;
        Bn = Bc;
        Bc = tmp;
      }
      // block reduction of the counts (fixed order) and the backward log-likelihood
      __shared__ double red[1024];
      const uint32_t nc = 5 + k + 16;
      for (uint32_t ci = 0; ci < nc; ++ci) {
        red[tid] = cnt[ci];
        __syncthreads();
        for (uint32_t half = 512; half > 0; half >>= 1) {
          if (tid < half && tid + half < nThreads) red[tid] += red[tid + half];
          __syncthreads();
        }
        if (tid == 0) args.counts[(size_t)read * nc + ci] = red[0];
        __syncthreads();
      }
      if (tid == 0) {
        double lb;
        if (tb.local) {
          lb = NEG;
          for (uint32_t s = 0; s < N; ++s) lb = lse(L2T, lb, Bn[(size_t)s * W]);
        } else
          lb = Bn[0];
        args.loglikeBack[read] = lb;
        args.sweepsBack[read] = sweepsBack;
      }
    }
    if (tid == 0) args.status[read] = status;
  }
}

cudaError_t launchForward(const ForwardTables& tb, const ForwardArgs& args, uint32_t nBlocks, uint32_t threads,
                          cudaStream_t stream) {
  forwardKernel<<<nBlocks, threads, 0, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
