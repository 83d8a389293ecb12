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
// tables of include/dnab_tables.h in the reference's list order.
//
// Frontier.  A Jacobi sweep evaluates state d from the previous sweep's values of its in-neighbours; if
// none of them changed in the previous sweep, the result is bit for bit what the previous sweep computed
// for d.  So from the second sweep on only the states with a changed in-neighbour are evaluated (38-67 %
// of the state-sweeps on the BASELINE machines): pass A reads one "changed" bit per in-neighbour from a
// bitmap in shared memory and compacts the active states into a per-CTA queue (one shared-memory atomic
// per warp), pass B evaluates the queue with all lanes busy.  A state that is skipped but changed in the
// previous sweep is copied into the other ping-pong buffer.  The cells, the sweep counts and the
// termination test are those of the specification (every state evaluated in every sweep).
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
  __shared__ uint32_t qCount[2];
  extern __shared__ uint32_t chgBits[];  // two bitmaps of nWords words: states changed by the latest / the previous sweep
  const uint32_t nWords = (N + 31) / 32, lane = tid & 31;
  uint32_t* const chg[2] = {chgBits, chgBits + nWords};
  // the frontier pays for its scan and its extra barrier only when a sweep is long: on machines with fewer than
  // 4 states per thread every sweep evaluates every state (measured on l4c4, 384 states: 14.3k vs 10.3k reads/s)
  const bool useFrontier = N >= 4 * nThreads;

  for (uint32_t s = tid; s < kMaxSyms; s += nThreads) symScore[s] = s < tb.nSyms ? tb.symScore[s] : NEG;
  double* base = args.scratch + (size_t)blockIdx.x * (10 + 2 * k) * N;
  uint32_t* const queue = reinterpret_cast<uint32_t*>(base + (size_t)(9 + 2 * k) * N);  // active states of a sweep
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
      // closure: Jacobi sweeps until a sweep changes no cell (evaluated on the frontier, see the header)
      cur = 0;
      uint32_t cw = 0;
      for (int sweep = 0;; ++sweep) {
        if (sweep >= args.maxSweeps) {
          status = 1;
          break;
        }
        const double* So = Sb[cur];
        const double* Do = Db[cur];
        double* Sn = Sb[cur ^ 1];
        double* Dn = Db[cur ^ 1];
        const uint32_t* chgOld = chg[cw];
        uint32_t* chgNew = chg[cw ^ 1];
        int changed = 0;
        // two states per call, evaluated in lockstep with branch-free arithmetic: the loads and the table
        // look-ups of the two dependent chains overlap (the sweeps are latency-bound); lane B may be absent
        auto evaluate2 = [&](uint32_t dA, uint32_t dB, bool hasA, bool hasB, bool& cA, bool& cB) {
          if (!hasA) dA = 0;
          if (!hasB) dB = 0;
          uint32_t eA = __ldg(tb.emitOff + dA), eB = __ldg(tb.emitOff + dB);
          const uint32_t eA1 = hasA ? __ldg(tb.emitOff + dA + 1) : eA, eB1 = hasB ? __ldg(tb.emitOff + dB + 1) : eB;
          uint32_t uA = __ldg(tb.nullOff + dA), uB = __ldg(tb.nullOff + dB);
          const uint32_t uA1 = hasA ? __ldg(tb.nullOff + dA + 1) : uA, uB1 = hasB ? __ldg(tb.nullOff + dB + 1) : uB;
          double ndA = NEG, nsA = S0[dA], ndB = NEG, nsB = S0[dB];
          while (eA < eA1 || eB < eB1) {
            const bool vA = eA < eA1, vB = eB < eB1;
            const uint32_t sA = vA ? __ldg(tb.emitSrc + eA) : 0u, sB = vB ? __ldg(tb.emitSrc + eB) : 0u;
            const uint32_t mA = vA ? __ldg(tb.emitMeta + eA) : 0u, mB = vB ? __ldg(tb.emitMeta + eB) : 0u;
            const double iA = lseFlat(L2T, Do[sA] + tb.delExtend, So[sA] + tb.delOpen) + symScore[mA & 31];
            const double iB = lseFlat(L2T, Do[sB] + tb.delExtend, So[sB] + tb.delOpen) + symScore[mB & 31];
            const double tA = lseFlat(L2T, ndA, iA), tB = lseFlat(L2T, ndB, iB);
            ndA = vA ? tA : ndA;
            ndB = vB ? tB : ndB;
            eA += vA;
            eB += vB;
          }
          while (uA < uA1 || uB < uB1) {
            const bool vA = uA < uA1, vB = uB < uB1;
            const uint32_t sA = vA ? __ldg(tb.nullSrc + uA) : 0u, sB = vB ? __ldg(tb.nullSrc + uB) : 0u;
            const double scA = symScore[vA ? __ldg(tb.nullSym + uA) : 0u], scB = symScore[vB ? __ldg(tb.nullSym + uB) : 0u];
            const double tdA = lseFlat(L2T, ndA, Do[sA] + scA), tdB = lseFlat(L2T, ndB, Do[sB] + scB);
            const double tsA = lseFlat(L2T, nsA, So[sA] + scA), tsB = lseFlat(L2T, nsB, So[sB] + scB);
            ndA = vA ? tdA : ndA;
            nsA = vA ? tsA : nsA;
            ndB = vB ? tdB : ndB;
            nsB = vB ? tsB : nsB;
            uA += vA;
            uB += vB;
          }
          nsA = lseFlat(L2T, nsA, ndA + tb.delEnd);
          nsB = lseFlat(L2T, nsB, ndB + tb.delEnd);
          cA = cB = false;
          if (hasA) {
            Dn[dA] = ndA;
            Sn[dA] = nsA;
            cA = __double_as_longlong(ndA) != __double_as_longlong(Do[dA]) || __double_as_longlong(nsA) != __double_as_longlong(So[dA]);
          }
          if (hasB) {
            Dn[dB] = ndB;
            Sn[dB] = nsB;
            cB = __double_as_longlong(ndB) != __double_as_longlong(Do[dB]) || __double_as_longlong(nsB) != __double_as_longlong(So[dB]);
          }
        };
        if (sweep == 0 || !useFrontier) {
          if (tid == 0) qCount[1] = 0;
          for (uint32_t d0 = tid - lane; d0 < N; d0 += 2 * nThreads) {
            const uint32_t dA = d0 + lane, dB = dA + nThreads;
            bool cA, cB;
            evaluate2(dA, dB, dA < N, dB < N, cA, cB);
            const uint32_t maskA = __ballot_sync(0xFFFFFFFFu, cA), maskB = __ballot_sync(0xFFFFFFFFu, cB);
            if (lane == 0) {
              chgNew[d0 >> 5] = maskA;
              if (d0 + nThreads < N) chgNew[(d0 + nThreads) >> 5] = maskB;
            }
            changed |= cA | cB;
          }
        } else {
          // pass A: which states have an in-neighbour that changed in the previous sweep?
          for (uint32_t d0 = tid - lane; d0 < N; d0 += nThreads) {
            const uint32_t d = d0 + lane;
            bool act = false;
            if (d < N) {
              for (uint32_t e = __ldg(tb.emitOff + d); e < __ldg(tb.emitOff + d + 1); ++e) {
                const uint32_t s = __ldg(tb.emitSrc + e);
                act |= (chgOld[s >> 5] >> (s & 31)) & 1u;
              }
              for (uint32_t e = __ldg(tb.nullOff + d); e < __ldg(tb.nullOff + d + 1); ++e) {
                const uint32_t s = __ldg(tb.nullSrc + e);
                act |= (chgOld[s >> 5] >> (s & 31)) & 1u;
              }
              if (!act && ((chgOld[d >> 5] >> (d & 31)) & 1u)) {  // same value as last sweep: bring the other buffer up to date
                Sn[d] = So[d];
                Dn[d] = Do[d];
              }
            }
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, act);
            uint32_t at = 0;
            if (lane == 0) {
              chgNew[d0 >> 5] = 0;
              if (mask) at = atomicAdd(&qCount[sweep & 1], (uint32_t)__popc(mask));
            }
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            if (act) queue[at + __popc(mask & ((1u << lane) - 1u))] = d;
          }
          __syncthreads();
          const uint32_t nQueued = qCount[sweep & 1];
          if (tid == 0) qCount[(sweep + 1) & 1] = 0;
          // pass B: evaluate them
          for (uint32_t i = tid; i < nQueued; i += 2 * nThreads) {
            const bool hasB = i + nThreads < nQueued;
            const uint32_t dA = queue[i], dB = hasB ? queue[i + nThreads] : 0u;
            bool cA, cB;
            evaluate2(dA, dB, true, hasB, cA, cB);
            if (cA) atomicOr(&chgNew[dA >> 5], 1u << (dA & 31));
            if (cB) atomicOr(&chgNew[dB >> 5], 1u << (dB & 31));
            changed |= cA | cB;
          }
        }
        cur ^= 1;
        cw ^= 1;
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
          // written once, read once by the backward pass much later: streaming stores, so that the cell stream
          // does not evict the closure columns of the 148 CTAs (64 MB) from L2
          double* c = args.F + ((size_t)blockIdx.x * (args.maxLen + 1) + pos) * (size_t)N * (k + 2) + (size_t)d * (k + 2);
          __stcs(c, s);
          __stcs(c + 1, Db[cur][d]);
          for (uint32_t i = 0; i < k; ++i) __stcs(c + 2 + i, Tb[tc][(size_t)i * N + d]);
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
      // posterior classes of the move that emits a read base: input-symbol id 0 (no input symbol) .. nSyms-1, then
      // "tandem duplication" (dnab_posterior_batch); at most kPostClasses, accumulated in registers
      double pacc[kPostClasses];
#pragma unroll
      for (int c = 0; c < kPostClasses; ++c) pacc[c] = 0.;
      const uint32_t dupClass = tb.nSyms;
      __shared__ double redP[32 * kPostClasses];
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
        uint32_t cb = 0, cwb = 0;
        for (int sweep = 0;; ++sweep) {
          if (sweep >= args.maxSweeps) {
            status = 1;
            break;
          }
          const double* So = S2[cb];
          const double* Do = D2[cb];
          double* Sn = S2[cb ^ 1];
          double* Dn = D2[cb ^ 1];
          const uint32_t* chgOld = chg[cwb];
          uint32_t* chgNew = chg[cwb ^ 1];
          int changed = 0;
          auto evaluate2 = [&](uint32_t sA, uint32_t sB, bool hasA, bool hasB, bool& cA, bool& cB) {
            if (!hasA) sA = 0;
            if (!hasB) sB = 0;
            uint32_t pA = __ldg(tb.outEmitOff + sA), pB = __ldg(tb.outEmitOff + sB);
            const uint32_t pA1 = hasA ? __ldg(tb.outEmitOff + sA + 1) : pA, pB1 = hasB ? __ldg(tb.outEmitOff + sB + 1) : pB;
            uint32_t uA = __ldg(tb.outNullOff + sA), uB = __ldg(tb.outNullOff + sB);
            const uint32_t uA1 = hasA ? __ldg(tb.outNullOff + sA + 1) : uA, uB1 = hasB ? __ldg(tb.outNullOff + sB + 1) : uB;
            double nsA = bbase[sA], ndA = NEG, nsB = bbase[sB], ndB = NEG;
            while (pA < pA1 || pB < pB1) {
              const bool vA = pA < pA1, vB = pB < pB1;
              const double scA = symScore[(vA ? __ldg(tb.outEmitMeta + pA) : 0u) & 31], scB = symScore[(vB ? __ldg(tb.outEmitMeta + pB) : 0u) & 31];
              const double bdA = Do[vA ? __ldg(tb.outEmitDst + pA) : 0u], bdB = Do[vB ? __ldg(tb.outEmitDst + pB) : 0u];
              const double tsA = lseFlat(L2T, nsA, (tb.delOpen + scA) + bdA), tsB = lseFlat(L2T, nsB, (tb.delOpen + scB) + bdB);
              const double tdA = lseFlat(L2T, ndA, (tb.delExtend + scA) + bdA), tdB = lseFlat(L2T, ndB, (tb.delExtend + scB) + bdB);
              nsA = vA ? tsA : nsA;
              ndA = vA ? tdA : ndA;
              nsB = vB ? tsB : nsB;
              ndB = vB ? tdB : ndB;
              pA += vA;
              pB += vB;
            }
            while (uA < uA1 || uB < uB1) {
              const bool vA = uA < uA1, vB = uB < uB1;
              const double scA = symScore[vA ? __ldg(tb.outNullSym + uA) : 0u], scB = symScore[vB ? __ldg(tb.outNullSym + uB) : 0u];
              const uint32_t dA = vA ? __ldg(tb.outNullDst + uA) : 0u, dB = vB ? __ldg(tb.outNullDst + uB) : 0u;
              const double tsA = lseFlat(L2T, nsA, scA + So[dA]), tsB = lseFlat(L2T, nsB, scB + So[dB]);
              const double tdA = lseFlat(L2T, ndA, scA + Do[dA]), tdB = lseFlat(L2T, ndB, scB + Do[dB]);
              nsA = vA ? tsA : nsA;
              ndA = vA ? tdA : ndA;
              nsB = vB ? tsB : nsB;
              ndB = vB ? tdB : ndB;
              uA += vA;
              uB += vB;
            }
            ndA = lseFlat(L2T, ndA, tb.delEnd + nsA);
            ndB = lseFlat(L2T, ndB, tb.delEnd + nsB);
            cA = cB = false;
            if (hasA) {
              Sn[sA] = nsA;
              Dn[sA] = ndA;
              cA = __double_as_longlong(nsA) != __double_as_longlong(So[sA]) || __double_as_longlong(ndA) != __double_as_longlong(Do[sA]);
            }
            if (hasB) {
              Sn[sB] = nsB;
              Dn[sB] = ndB;
              cB = __double_as_longlong(nsB) != __double_as_longlong(So[sB]) || __double_as_longlong(ndB) != __double_as_longlong(Do[sB]);
            }
          };
          if (sweep == 0 || !useFrontier) {
            if (tid == 0) qCount[1] = 0;
            for (uint32_t s0 = tid - lane; s0 < N; s0 += 2 * nThreads) {
              const uint32_t sA = s0 + lane, sB = sA + nThreads;
              bool cA, cB;
              evaluate2(sA, sB, sA < N, sB < N, cA, cB);
              const uint32_t maskA = __ballot_sync(0xFFFFFFFFu, cA), maskB = __ballot_sync(0xFFFFFFFFu, cB);
              if (lane == 0) {
                chgNew[s0 >> 5] = maskA;
                if (s0 + nThreads < N) chgNew[(s0 + nThreads) >> 5] = maskB;
              }
              changed |= cA | cB;
            }
          } else {
            // pass A: states with an out-neighbour that changed in the previous sweep
            for (uint32_t s0 = tid - lane; s0 < N; s0 += nThreads) {
              const uint32_t s = s0 + lane;
              bool act = false;
              if (s < N) {
                for (uint32_t p = __ldg(tb.outEmitOff + s); p < __ldg(tb.outEmitOff + s + 1); ++p) {
                  const uint32_t d = __ldg(tb.outEmitDst + p);
                  act |= (chgOld[d >> 5] >> (d & 31)) & 1u;
                }
                for (uint32_t p = __ldg(tb.outNullOff + s); p < __ldg(tb.outNullOff + s + 1); ++p) {
                  const uint32_t d = __ldg(tb.outNullDst + p);
                  act |= (chgOld[d >> 5] >> (d & 31)) & 1u;
                }
                if (!act && ((chgOld[s >> 5] >> (s & 31)) & 1u)) {
                  Sn[s] = So[s];
                  Dn[s] = Do[s];
                }
              }
              const uint32_t mask = __ballot_sync(0xFFFFFFFFu, act);
              uint32_t at = 0;
              if (lane == 0) {
                chgNew[s0 >> 5] = 0;
                if (mask) at = atomicAdd(&qCount[sweep & 1], (uint32_t)__popc(mask));
              }
              at = __shfl_sync(0xFFFFFFFFu, at, 0);
              if (act) queue[at + __popc(mask & ((1u << lane) - 1u))] = s;
            }
            __syncthreads();
            const uint32_t nQueued = qCount[sweep & 1];
            if (tid == 0) qCount[(sweep + 1) & 1] = 0;
            for (uint32_t i = tid; i < nQueued; i += 2 * nThreads) {
              const bool hasB = i + nThreads < nQueued;
              const uint32_t sA = queue[i], sB = hasB ? queue[i + nThreads] : 0u;
              bool cA, cB;
              evaluate2(sA, sB, true, hasB, cA, cB);
              if (cA) atomicOr(&chgNew[sA >> 5], 1u << (sA & 31));
              if (cB) atomicOr(&chgNew[sB >> 5], 1u << (sB & 31));
              changed |= cA | cB;
            }
          }
          cb ^= 1;
          cwb ^= 1;
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
          const double fS = __ldcs(fc), fD = __ldcs(fc + 1);
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
              if (args.post) {
                const uint32_t sy = meta & 31u;
#pragma unroll
                for (int c = 0; c < kPostClasses; ++c) pacc[c] += ((uint32_t)c == sy) ? u : 0.;
              }
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
            double dupU = post(__ldcs(fc + 2), tb.sub[c0 * 4 + xn], bn[0]);
            cnt[5 + k + c0 * 4 + xn] += dupU;
            for (uint32_t i = 0; i + 1 < mdl; ++i) {
              const uint32_t ci = __ldg(tb.ctx + (size_t)s * k + i + 1);
              const double u = post(__ldcs(fc + 2 + i + 1), tb.sub[ci * 4 + xn], bn[2 + i]);
              cnt[5 + k + ci * 4 + xn] += u;
              dupU += u;
            }
            if (args.post) {
#pragma unroll
              for (int c = 0; c < kPostClasses; ++c) pacc[c] += ((uint32_t)c == dupClass) ? dupU : 0.;
            }
          }
        }
        __syncthreads();
        if (args.post) {
          // who emitted read base `pos`: block reduction in a fixed order (warp tree, then the warps in order)
#pragma unroll
          for (int c = 0; c < kPostClasses; ++c) {
            double v = pacc[c];
            for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, sh);
            if ((tid & 31u) == 0) redP[(tid >> 5) * kPostClasses + c] = v;
            pacc[c] = 0.;
          }
          __syncthreads();
          if (tid <= dupClass && pos < L) {
            double v = 0.;
            for (uint32_t w = 0; w < (nThreads + 31) / 32; ++w) v += redP[w * kPostClasses + tid];
            args.post[((size_t)args.postOff[read] + pos) * (dupClass + 1) + tid] = v;
          }
          __syncthreads();
        }
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
  const size_t smem = 2 * (size_t)((tb.nStates + 31) / 32) * sizeof(uint32_t);  // the two "changed" bitmaps
  if (smem > 200 * 1024) return cudaErrorInvalidValue;                            // > 819,200 states
  if (smem > 32 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(forwardKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  forwardKernel<<<nBlocks, threads, smem, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
