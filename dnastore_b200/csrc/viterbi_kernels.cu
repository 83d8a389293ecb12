// Hand-written sm_100a kernels for dnastore's Viterbi hot path.
//
// What they compute is the reference's ViterbiMatrix fill + traceback
// (reference src/viterbi.cpp:62-176 and :195-304) for a BATCH of reads; how they
// compute it is B200-native:
//
//  * One thread-block CLUSTER per read.  The state space is cut into C contiguous
//    slices; CTA `rank` keeps its slice of the three live fp64 columns -- S(pos-1),
//    S(pos), D(pos) -- and (when they fit) the k duplication columns T in its own
//    shared memory.  A transition whose source lives in another CTA is read through
//    distributed shared memory (mapa + ld.shared::cluster), never through HBM.
//  * Per column: (1) emission step from the previous column; (2) the within-column
//    closure over null transitions and deletions as a frontier-driven, pull-style
//    chaotic relaxation -- the system is monotone, so ANY schedule reaches the same
//    least fixed point bit for bit; races are benign (values only
//    grow towards the fixed point) and are made well-defined with relaxed
//    cluster-scope accesses; (3) one predecessor byte per DP cell, evaluated with
//    the TRACEBACK's own floating-point association and candidate order
//    (src/viterbi.cpp:251-286) on the converged column, streamed to HBM with
//    coalesced byte-plane stores; (4) duplication opens.
//  * A second kernel walks the predecessor bytes on the device, one thread per
//    read, and emits the decoded input-symbol string, log-likelihood and status.
//
// Only fp64 add / compare / select are used on the device; every score is computed
// on the host with the reference's libm calls (include/dnab_tables.h).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "viterbi_kernels.h"

namespace cg = cooperative_groups;

namespace dnab {

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t mapToRank(uint32_t localAddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(localAddr), "r"(rank));
  return r;
}
// plain (weak) DSMEM load: used where the column being read is final
__device__ __forceinline__ double ldClusterF64(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// relaxed cluster-scope accesses: used inside the closure where other CTAs may be
// raising the same cells concurrently
__device__ __forceinline__ double ldRelaxedF64(uint32_t addr) {
  double v;
  asm volatile("ld.relaxed.cluster.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void stRelaxedF64(uint32_t addr, double v) {
  asm volatile("st.relaxed.cluster.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// flag / control stores into a peer: made visible by the next cluster barrier
__device__ __forceinline__ void stClusterU8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void stClusterU32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// std::max(a,b) of the reference: keeps a on ties, no NaN handling needed
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }

__device__ __forceinline__ double negInf() { return __longlong_as_double(0xFFF0000000000000LL); }

// ---------------------------------------------------------------------------
// shared-memory carve-up (identical in every CTA of a cluster, which is what lets
// a local offset be mapped into a peer with mapa)
// ---------------------------------------------------------------------------
uint32_t fillSmemBytes(uint32_t M, uint32_t k, uint32_t tInSmem, uint32_t maxLen) {
  return makeLayout(M, k, tInSmem, maxLen).total;
}

// ---------------------------------------------------------------------------
// fill kernel
// ---------------------------------------------------------------------------
struct Cta {
  const DevTables* tb;
  unsigned char* smem;
  uint32_t smemBase;  // shared-window address of smem[0]
  const SmemLayout* layp;  // lives in the kernel parameter (constant) space
  uint32_t rank, M, tid, nThreads;
  const double* symScore;
  const uint32_t* boff;
};

// The first eight words of a state block in registers (one 32-byte sector).
struct BlockRegs {
  uint32_t w0, w1, e0, e1, e2, e3, e4, e5;
  const uint32_t* p;
};
__device__ __forceinline__ BlockRegs loadBlock(const Cta& c, uint32_t i) {
  BlockRegs b;
  b.p = c.tb->blocks + c.boff[i];
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(b.p));
  const uint4 d = __ldg(reinterpret_cast<const uint4*>(b.p) + 1);
  b.w0 = a.x; b.w1 = a.y; b.e0 = a.z; b.e1 = a.w;
  b.e2 = d.x; b.e3 = d.y; b.e4 = d.z; b.e5 = d.w;
  return b;
}
// Edge words [base, base+4) of a block: from registers for the first chunk, else from L1/L2.
__device__ __forceinline__ void chunkWords(const BlockRegs& b, uint32_t base, uint32_t n, uint32_t (&w)[4]) {
  if (base == 0) {
    w[0] = b.e0; w[1] = b.e1; w[2] = b.e2; w[3] = b.e3;
  } else {
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) w[j] = (base + j < n) ? __ldg(b.p + 2 + base + j) : 0u;
  }
}

// Wake every successor of local state i (its cells grew). Returns true if a peer CTA was marked.
// bit 0 of the result: a state of this CTA was woken; bit 1: a peer CTA was woken.
__device__ __forceinline__ uint32_t wakeSuccessors(const Cta& c, const BlockRegs& b, uint32_t nIn, uint32_t nOut,
                                                   uint32_t localBuf, uint32_t remoteBuf) {
  uint32_t sent = 0;
  for (uint32_t j = 0; j < nOut; ++j) {
    const uint32_t idx = 2 + nIn + j;
    const uint32_t w = idx == 2 ? b.e0 : idx == 3 ? b.e1 : idx == 4 ? b.e2 : idx == 5 ? b.e3 : idx == 6 ? b.e4
                     : idx == 7 ? b.e5 : __ldg(b.p + idx);
    const uint32_t r = edgeRank(w), l = edgeLocal(w);
    if (r == c.rank) {
      c.smem[c.layp->flagLocal[localBuf] + l] = 1;
      sent |= 1u;
    } else {
      stClusterU8(mapToRank(c.smemBase + c.layp->flagRemote[remoteBuf] + l, r), 1u);
      sent |= 2u;
    }
  }
  return sent;
}

// One relaxation of local state i (pull form of src/viterbi.cpp:118-158):
//   D(d) = max( D(d), max_emit-in ( max(D(s)+delExtend, S(s)+delOpen) + score ), max_null-in ( D(s)+score ) )
//   S(d) = max( S(d), max_null-in ( S(s)+score ), D(d)+delEnd )
// where the stored S(s) already contains D(s)+delEnd from s's own last relaxation.
// When a cell grew, every successor is woken (result: see wakeSuccessors).
__device__ __forceinline__ uint32_t relaxState(const Cta& c, uint32_t i, uint32_t sCurOff, uint32_t localBuf,
                                               uint32_t remoteBuf) {
  const DevTables& tb = *c.tb;
  const BlockRegs b = loadBlock(c, i);
  const uint32_t nE = hdrNEmit(b.w0), nN = hdrNNull(b.w0), nIn = nE + nN;
  const uint32_t myS = c.smemBase + sCurOff + i * 8, myD = c.smemBase + c.layp->dBuf + i * 8;
  const double oldS = ldRelaxedF64(myS), oldD = ldRelaxedF64(myD);
  double newS = oldS, newD = oldD;
  const uint32_t dMinusS = c.layp->dBuf - sCurOff;
  for (uint32_t base = 0; base < nIn; base += 4) {
    uint32_t w[4];
    chunkWords(b, base, nIn, w);
    double ss[4], ds[4];
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j)
      if (base + j < nIn) {  // issue every load of the chunk before the first use
        const uint32_t aS = mapToRank(c.smemBase + sCurOff + edgeLocal(w[j]) * 8, edgeRank(w[j]));
        ss[j] = ldRelaxedF64(aS);
        ds[j] = ldRelaxedF64(aS + dMinusS);
      }
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j)
      if (base + j < nIn) {
        const double sc = c.symScore[edgeSym(w[j])];
        if (base + j < nE) {
          newD = dmax(newD, dmax(ds[j] + tb.delExtend, ss[j] + tb.delOpen) + sc);
        } else {
          newD = dmax(newD, ds[j] + sc);
          newS = dmax(newS, ss[j] + sc);
        }
      }
  }
  newS = dmax(newS, newD + tb.delEnd);
  uint32_t sent = 0;
  if ((newD > oldD) || (newS > oldS)) {
    if (newD > oldD) stRelaxedF64(myD, newD);
    if (newS > oldS) stRelaxedF64(myS, newS);
    sent = wakeSuccessors(c, b, nIn, hdrNOut(b.w0), localBuf, remoteBuf);
  }
  return sent;
}

template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
    viterbiFillKernel(const __grid_constant__ DevTables tb, const __grid_constant__ FillArgs args) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t C = tb.C, M = tb.M, k = tb.k;
  const uint32_t rank = C > 1 ? cluster.block_rank() : 0;
  const uint32_t clusterId = blockIdx.x / C;
  const uint32_t nClusters = gridDim.x / C;
  const uint32_t tid = threadIdx.x, nThreads = blockDim.x;
  const uint32_t Np = C * M;
  const double NEG = negInf();

  Cta c;
  c.tb = &tb;
  c.smem = smem;
  c.smemBase = smemAddr(smem);
  c.layp = &args.lay;
  c.rank = rank;
  c.M = M;
  c.tid = tid;
  c.nThreads = nThreads;
  const SmemLayout& lay = args.lay;

  double* symScore = reinterpret_cast<double*>(smem + lay.symScore);
  double* tsE = reinterpret_cast<double*>(smem + lay.tsE);
  double* tsDext = reinterpret_cast<double*>(smem + lay.tsDext);
  double* tsDopen = reinterpret_cast<double*>(smem + lay.tsDopen);
  double* subS = reinterpret_cast<double*>(smem + lay.sub);
  double* tsT = reinterpret_cast<double*>(smem + lay.tsT);
  double* lenS = reinterpret_cast<double*>(smem + lay.len);
  volatile uint32_t* ctl = reinterpret_cast<volatile uint32_t*>(smem + lay.ctl);
  uint32_t* boff = reinterpret_cast<uint32_t*>(smem + lay.boff);
  uint8_t* seqS = smem + lay.seq;
  double* dCol = reinterpret_cast<double*>(smem + lay.dBuf);
  double* tCol = tb.tInSmem ? reinterpret_cast<double*>(smem + lay.tBuf)
                            : args.tScratch + ((size_t)clusterId * C + rank) * (size_t)k * M;
  c.symScore = symScore;
  c.boff = boff;

  // read-independent tables; the traceback-association sums are formed here once
  for (uint32_t s = tid; s < kMaxSyms; s += nThreads) {
    const double sc = s < tb.nSyms ? tb.symScore[s] : NEG;
    symScore[s] = sc;
    tsDext[s] = sc + tb.delExtend;
    tsDopen[s] = sc + tb.delOpen;
  }
  for (uint32_t j = tid; j < 16; j += nThreads) subS[j] = tb.sub[j];
  for (uint32_t j = tid; j < 8; j += nThreads) {
    lenS[j] = j < k ? tb.len[j] : NEG;
    tsT[j] = j < k ? tb.tanDup + tb.len[j] : NEG;
  }
  auto buildTsE = [&]() {
    for (uint32_t j = tid; j < tb.nSyms * 16; j += nThreads) {
      const uint32_t s = j >> 4, bx = j & 15;
      tsE[j] = (tb.symScore[s] + tb.noGap) + tb.sub[bx];
    }
  };
  buildTsE();
  for (uint32_t i = tid; i < M; i += nThreads) boff[i] = __ldg(&tb.blockOff[rank * M + i]);
  for (uint32_t j = tid; j < 64; j += nThreads) ctl[j] = 0;
  {
    uint32_t* fl = reinterpret_cast<uint32_t*>(smem + lay.flagLocal[0]);
    const uint32_t nFlagWords = (lay.seq - lay.flagLocal[0]) / 4;  // the four flag arrays are contiguous
    for (uint32_t j = tid; j < nFlagWords; j += nThreads) fl[j] = 0;
  }
  __syncthreads();

  auto clusterBarrier = [&]() {
    if (C > 1)
      cluster.sync();
    else
      __syncthreads();
  };
  clusterBarrier();  // every CTA's flags are clear before a peer can set them

  unsigned long long dbgDense = 0, dbgSyncWait = 0, dbgCompact = 0, dbgProc = 0, dbgClusterWait = 0;
  unsigned long long dbgRounds = 0, dbgIters = 0, dbgT1 = 0, dbgT2 = 0, dbgT3 = 0, dbgCols = 0, dbgWork = 0;
  const bool dbgOn = args.dbg != nullptr;

  for (int64_t read = clusterId; read < args.nReads; read += nClusters) {
    const int32_t L = args.readLen[read];
    {  // stage the packed read
      const uint8_t* src = args.packed + args.byteOff[read];
      const uint32_t nVec = ((uint32_t)(L + 3) / 4 + 15) / 16;
      for (uint32_t v = tid; v < nVec; v += nThreads)
        reinterpret_cast<uint4*>(seqS)[v] = __ldg(reinterpret_cast<const uint4*>(src) + v);
    }
    uint8_t* predRead = args.pred + (size_t)read * (size_t)(args.maxLen + 1) * (k + 2) * Np;
    __syncthreads();

    for (int32_t pos = 0; pos <= L; ++pos) {
      const uint32_t cur = pos & 1, prev = cur ^ 1;
      const uint32_t sCurOff = lay.sBuf[cur], sPrevOff = lay.sBuf[prev];
      double* sCur = reinterpret_cast<double*>(smem + sCurOff);
      const uint32_t x = pos > 0 ? (seqS[(pos - 1) >> 2] >> (2 * ((pos - 1) & 3))) & 3u : 0u;

      long long tc0 = dbgOn ? clock64() : 0;
      // ---- (1) emission step: S0 from the previous column, T shift (src/viterbi.cpp:92-106) ----
      for (uint32_t i = tid; i < M; i += nThreads) {
        double s = NEG;
        if (pos == 0) {
          const uint32_t g = rank * M + i;
          const bool real = __ldg(&tb.origId[g]) != 0xFFFFFFFFu;
          s = (real && (tb.local || g == tb.startG)) ? 0.0 : NEG;  // src/viterbi.cpp:75-79
          for (uint32_t j = 0; j < k; ++j) tCol[j * M + i] = NEG;
        } else {
          const BlockRegs b = loadBlock(c, i);
          const uint32_t nE = hdrNEmit(b.w0), mdl = hdrMdl(b.w0);
          double t0 = NEG;
          if (mdl > 0) t0 = tCol[i];
          for (uint32_t base = 0; base < nE; base += 4) {
            uint32_t w[4];
            chunkWords(b, base, nE, w);
            double v[4];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
              if (base + j < nE) v[j] = ldClusterF64(mapToRank(c.smemBase + sPrevOff + edgeLocal(w[j]) * 8, edgeRank(w[j])));
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
              if (base + j < nE)
                s = dmax(s, ((v[j] + symScore[edgeSym(w[j])]) + tb.noGap) + subS[edgeBase(w[j]) * 4 + x]);
          }
          if (mdl > 0) {
            const double t2s = t0 + subS[hdrCtx(b.w1, 0) * 4 + x];
            s = dmax(s, t2s);
            for (uint32_t j = 0; j + 1 < mdl; ++j) tCol[j * M + i] = tCol[(j + 1) * M + i] + subS[hdrCtx(b.w1, j + 1) * 4 + x];
            tCol[(mdl - 1) * M + i] = t2s;  // slot mdl-1 is free until step (4): park the T->S candidate there
          }
        }
        sCur[i] = s;
        dCol[i] = NEG;
      }
      clusterBarrier();
      long long tc1 = dbgOn ? clock64() : 0;

      // ---- (2) closure: null transitions + deletions (src/viterbi.cpp:110-159) ----
      // Round 0 relaxes every state once.  Then: each CTA iterates over the states woken by its
      // OWN states until none is left (CTA barriers only); states woken by a peer wait in the
      // remote flag buffer of the current round and are picked up after the next cluster
      // barrier.  The cluster is done when a whole round woke nobody across CTAs.
      {
        // Every thread owns the 4-state flag words tid, tid+nThreads, ...: it relaxes the states whose
        // flag is set in the buffers being consumed and sets flags in the buffers being filled.
        const uint32_t nWords = (M + 3) / 4;
        uint32_t round = 0, it = 0, sent = 0;
        for (uint32_t i = tid; i < M; i += nThreads) sent |= relaxState(c, i, sCurOff, 1, 0);
        if (dbgOn && tid == 0) dbgDense += clock64() - tc1;
        bool pickRemote = false;
        for (;;) {
          // local fixed point: one CTA barrier per iteration
          while (__syncthreads_or((int)(sent & 1u)) || pickRemote) {
            ++it;
            sent &= 2u;
            uint32_t* fl = reinterpret_cast<uint32_t*>(smem + lay.flagLocal[it & 1]);
            uint32_t* fr = reinterpret_cast<uint32_t*>(smem + lay.flagRemote[(round & 1) ^ 1]);
            for (uint32_t w = tid; w < nWords; w += nThreads) {
              uint32_t f = fl[w];
              if (f) fl[w] = 0;
              if (pickRemote) {
                const uint32_t g = fr[w];
                if (g) fr[w] = 0;
                f |= g;
              }
              if (f) {
                if (dbgOn) dbgWork += __popc(f & 0x01010101u);
#pragma unroll 1
                for (uint32_t q = 0; q < 4; ++q)
                  if (f & (0xFFu << (8 * q))) sent |= relaxState(c, 4 * w + q, sCurOff, (it & 1) ^ 1, round & 1);
              }
            }
            pickRemote = false;
            if (dbgOn && tid == 0) dbgIters++;
          }
          if (C == 1) break;
          long long tcb = dbgOn ? clock64() : 0;
          const uint32_t anySent = (uint32_t)__syncthreads_or((int)(sent & 2u));
          if (tid < C) stClusterU32(mapToRank(c.smemBase + lay.ctl + (16 + (round & 1) * kMaxCluster + rank) * 4, tid), anySent);
          cluster.sync();
          if (dbgOn && tid == 0) dbgClusterWait += clock64() - tcb;
          uint32_t tot = 0;
          for (uint32_t r = 0; r < C; ++r) tot |= ctl[16 + (round & 1) * kMaxCluster + r];
          if (!tot) break;
          if (dbgOn && tid == 0) dbgRounds++;
          ++round;  // peers now fill the other remote buffer; the one just completed is picked up next
          pickRemote = true;
          sent = 0;
        }
      }

      long long tc2 = dbgOn ? clock64() : 0;
      // ---- (3) predecessor records with the traceback's arithmetic (src/viterbi.cpp:251-286)
      //      (4) duplication opens (src/viterbi.cpp:161-168) ----
      uint8_t* predCol = predRead + (size_t)pos * (k + 2) * Np;
      for (uint32_t i = tid; i < M; i += nThreads) {
        const uint32_t g = rank * M + i;
        const BlockRegs b = loadBlock(c, i);
        const uint32_t nE = hdrNEmit(b.w0), nN = hdrNNull(b.w0), mdl = hdrMdl(b.w0), nIn = nE + nN;
        const double sHere = sCur[i], dHere = dCol[i];
        const uint32_t dMinusS = lay.dBuf - sCurOff;
        const double parked = (mdl > 0 && pos > 0) ? tCol[(mdl - 1) * M + i] : NEG;  // T(state,pos-1,0)+sub

        double best = NEG, bestD = NEG;
        uint32_t idx = kNoPred, idxD = kNoPred;
        for (uint32_t base = 0; base < nIn; base += 2) {
          uint32_t w[2];
          if (base == 0) {
            w[0] = b.e0;
            w[1] = b.e1;
          } else if (base == 2) {
            w[0] = b.e2;
            w[1] = b.e3;
          } else if (base == 4) {
            w[0] = b.e4;
            w[1] = b.e5;
          } else {
            w[0] = __ldg(b.p + 2 + base);
            w[1] = base + 1 < nIn ? __ldg(b.p + 3 + base) : 0u;
          }
          double vp[2], vs[2], vd[2];
#pragma unroll
          for (uint32_t j = 0; j < 2; ++j)
            if (base + j < nIn) {
              const uint32_t aS = mapToRank(c.smemBase + sCurOff + edgeLocal(w[j]) * 8, edgeRank(w[j]));
              vs[j] = ldClusterF64(aS);
              vd[j] = ldClusterF64(aS + dMinusS);
              if (base + j < nE && pos > 0) vp[j] = ldClusterF64(aS + (sPrevOff - sCurOff));
            }
#pragma unroll
          for (uint32_t j = 0; j < 2; ++j)
            if (base + j < nIn) {
              const uint32_t e = base + j, sym = edgeSym(w[j]);
              if (e < nE) {
                if (pos > 0) {
                  const double v = vp[j] + tsE[sym * 16 + edgeBase(w[j]) * 4 + x];
                  if (v > best) {
                    best = v;
                    idx = e;
                  }
                }
                const double ve = vd[j] + tsDext[sym];
                if (ve > bestD) {
                  bestD = ve;
                  idxD = 2 * e;
                }
                const double vo = vs[j] + tsDopen[sym];
                if (vo > bestD) {
                  bestD = vo;
                  idxD = 2 * e + 1;
                }
              } else {
                const double sc = symScore[sym];
                const double v = vs[j] + sc;
                if (v > best) {
                  best = v;
                  idx = e;
                }
                const double vn = vd[j] + sc;
                if (vn > bestD) {
                  bestD = vn;
                  idxD = nE + e;  // = 2*nE + (e - nE)
                }
              }
            }
        }
        {
          const double v = dHere + tb.delEnd;
          if (v > best) {
            best = v;
            idx = nIn;
          }
        }
        if (mdl > 0 && pos > 0 && parked > best) {
          best = parked;
          idx = nIn + 1;
        }
        if (tb.local && pos == 0) {
          const double v = ldClusterF64(mapToRank(c.smemBase + sCurOff + (tb.startG % M) * 8, tb.startG / M)) + 0.0;
          if (v > best) {
            best = v;
            idx = nIn + 2;
          }
        }
        const bool real = __ldg(&tb.origId[g]) != 0xFFFFFFFFu;
        predCol[g] = (uint8_t)(real ? idx : kNoPred);
        predCol[Np + g] = (uint8_t)(real ? idxD : kNoPred);
        for (uint32_t j = 0; j < k; ++j) {
          uint32_t idxT = kNoPred;
          if (pos > 0 && j < mdl) {
            const double shifted = (j + 1 < mdl) ? tCol[j * M + i] : NEG;
            if (j + 1 < mdl && shifted > NEG) idxT = 0;
            if (sHere + tsT[j] > shifted) idxT = 1;
            tCol[j * M + i] = dmax(shifted, (sHere + tb.tanDup) + lenS[j]);  // (4)
          }
          predCol[(size_t)(2 + j) * Np + g] = (uint8_t)idxT;
        }
        if (args.cells && read == 0 && real) {
          double* cell = args.cells + ((size_t)pos * tb.nStates + __ldg(&tb.origId[g])) * (k + 2);
          cell[0] = sHere;
          cell[1] = dHere;
          for (uint32_t j = 0; j < k; ++j) cell[2 + j] = (pos > 0 && j < mdl) ? tCol[j * M + i] : NEG;
        }
      }
      clusterBarrier();
      if (dbgOn && tid == 0) {
        const long long tc3 = clock64();
        dbgT1 += tc1 - tc0;
        dbgT2 += tc2 - tc1;
        dbgT3 += tc3 - tc2;
        dbgCols++;
      }
    }

    // ---- end of read: log-likelihood and traceback start (src/viterbi.cpp:171-173, 239-245) ----
    {
      const double* sLast = reinterpret_cast<const double*>(smem + lay.sBuf[L & 1]);
      if (!tb.local) {
        if (rank == tb.endG / M && tid == 0) {
          args.loglike[read] = sLast[tb.endG % M];
          args.startState[read] = tb.endG;
        }
      } else {
        // first strict maximum in REFERENCE state order: max value, then smallest original index
        double bv = NEG;
        uint32_t bo = 0xFFFFFFFFu, bg = 0;
        for (uint32_t i = tid; i < M; i += nThreads) {
          const uint32_t g = rank * M + i;
          const uint32_t o = __ldg(&tb.origId[g]);
          if (o == 0xFFFFFFFFu) continue;
          const double v = sLast[i];
          if (v > bv || (v == bv && o < bo)) {
            bv = v;
            bo = o;
            bg = g;
          }
        }
        for (int off = 16; off > 0; off >>= 1) {
          const double ov = __shfl_down_sync(0xFFFFFFFFu, bv, off);
          const uint32_t oo = __shfl_down_sync(0xFFFFFFFFu, bo, off);
          const uint32_t og = __shfl_down_sync(0xFFFFFFFFu, bg, off);
          if (ov > bv || (ov == bv && oo < bo)) {
            bv = ov;
            bo = oo;
            bg = og;
          }
        }
        __syncthreads();
        // reduction scratch: the tsE table (rebuilt below) is free between reads
        double* rv = tsE;
        uint32_t* ro = reinterpret_cast<uint32_t*>(tsE + 32);
        if ((tid & 31) == 0) {
          rv[tid >> 5] = bv;
          ro[2 * (tid >> 5)] = bo;
          ro[2 * (tid >> 5) + 1] = bg;
        }
        __syncthreads();
        if (tid == 0) {
          const uint32_t nw = (nThreads + 31) / 32;
          for (uint32_t w = 1; w < nw; ++w)
            if (rv[w] > bv || (rv[w] == bv && ro[2 * w] < bo)) {
              bv = rv[w];
              bo = ro[2 * w];
              bg = ro[2 * w + 1];
            }
          args.partVal[read * C + rank] = bv;
          args.partOrig[read * C + rank] = bo;
          args.partG[read * C + rank] = bg;
        }
        __syncthreads();
        buildTsE();
        __syncthreads();
      }
    }
  }
  if (dbgOn && tid == 0 && rank == 0) {
    atomicAdd(&args.dbg[0], dbgCols);
    atomicAdd(&args.dbg[1], dbgIters);
    atomicAdd(&args.dbg[2], dbgWork);
    atomicAdd(&args.dbg[3], dbgT1);
    atomicAdd(&args.dbg[4], dbgT2);
    atomicAdd(&args.dbg[5], dbgT3);
    atomicAdd(&args.dbg[6], dbgRounds);
    atomicAdd(&args.dbg[7], dbgDense);
    atomicAdd(&args.dbg[8], dbgSyncWait);
    atomicAdd(&args.dbg[9], dbgCompact);
    atomicAdd(&args.dbg[10], dbgProc);
    atomicAdd(&args.dbg[11], dbgClusterWait);
  }
  clusterBarrier();  // no CTA may exit while a peer can still read its shared memory
}

// ---------------------------------------------------------------------------
// traceback kernel: one thread per read follows the predecessor bytes
// (reference src/viterbi.cpp:195-304: loop :247, emitted symbols :299-300)
// ---------------------------------------------------------------------------
__global__ void viterbiTracebackKernel(const DevTables tb, const TracebackArgs args) {
  const int64_t read = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (read >= args.nReads) return;
  const uint32_t C = tb.C, M = tb.M, k = tb.k, Np = C * M;
  const int32_t L = args.readLen[read];
  const double NEG = negInf();

  uint32_t g;
  double ll;
  if (!tb.local) {
    g = args.startState[read];
    ll = args.loglike[read];
  } else {
    double bv = NEG;
    uint32_t bo = 0xFFFFFFFFu, bg = 0;
    for (uint32_t r = 0; r < C; ++r) {
      const double v = args.partVal[read * C + r];
      const uint32_t o = args.partOrig[read * C + r];
      if (o == 0xFFFFFFFFu) continue;
      if (v > bv || (v == bv && o < bo)) {
        bv = v;
        bo = o;
        bg = args.partG[read * C + r];
      }
    }
    g = bg;
    ll = bv;
    args.loglike[read] = ll;
  }

  char* out = args.decoded + (size_t)read * args.decodedStride;
  int32_t* path = args.path ? args.path + (size_t)read * 3 * args.pathStride : nullptr;
  int32_t nOut = 0, nPath = 0;
  int32_t status = DNAB_READ_OK_;
  if (!(ll > NEG)) {
    args.decodedLen[read] = 0;
    args.status[read] = DNAB_READ_NO_DECODING_;
    if (args.pathLen) args.pathLen[read] = 0;
    return;
  }

  const uint8_t* predRead = args.pred + (size_t)read * (size_t)(args.maxLen + 1) * (k + 2) * Np;
  int32_t pos = L;
  uint32_t mut = 0;
  // symbols are produced last-to-first: fill the caller's slot from its END, then shift
  const int32_t cap = args.decodedStride;
  while (pos >= 0 && g != tb.startG) {
    if (path) {
      if (nPath < args.pathStride) {
        path[3 * nPath] = (int32_t)tb.origId[g];
        path[3 * nPath + 1] = pos;
        path[3 * nPath + 2] = (int32_t)mut;
      } else
        status = DNAB_READ_OVERFLOW_;
    }
    ++nPath;
    const uint32_t p = predRead[((size_t)pos * (k + 2) + mut) * Np + g];
    if (p == kNoPred) {
      status = DNAB_READ_TRACEBACK_FAILED_;
      break;
    }
    const uint32_t* blk = tb.blocks + tb.blockOff[g];
    const uint32_t nE = hdrNEmit(blk[0]), nN = hdrNNull(blk[0]);
    const uint32_t* edges = blk + 2;
    uint32_t sym = 0;
    if (mut == 0) {
      if (p < nE) {
        const uint32_t w = edges[p];
        sym = edgeSym(w);
        g = edgeRank(w) * M + edgeLocal(w);
        --pos;
      } else if (p < nE + nN) {
        const uint32_t w = edges[p];
        sym = edgeSym(w);
        g = edgeRank(w) * M + edgeLocal(w);
      } else if (p == nE + nN) {
        mut = 1;
      } else if (p == nE + nN + 1) {
        mut = 2;
        --pos;
      } else {
        g = tb.startG;  // local mode, pos == 0: jump to (0,0,S)
      }
    } else if (mut == 1) {
      if (p < 2 * nE) {
        const uint32_t w = edges[p >> 1];
        sym = edgeSym(w);
        g = edgeRank(w) * M + edgeLocal(w);
        mut = (p & 1) ? 0 : 1;
      } else {
        const uint32_t w = edges[nE + (p - 2 * nE)];
        sym = edgeSym(w);
        g = edgeRank(w) * M + edgeLocal(w);
      }
    } else {
      if (p == 0) {
        mut += 1;
        --pos;
      } else
        mut = 0;
    }
    if (sym) {
      if (nOut < cap)
        out[cap - 1 - nOut] = (char)tb.symChar[sym];
      else
        status = DNAB_READ_OVERFLOW_;
      ++nOut;
    }
  }
  const int32_t kept = nOut < cap ? nOut : cap;
  for (int32_t i = 0; i < kept; ++i) out[i] = out[cap - kept + i];
  args.decodedLen[read] = kept;
  args.status[read] = status;
  if (args.pathLen) args.pathLen[read] = nPath;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
typedef void (*FillKernelPtr)(const DevTables, const FillArgs);
static FillKernelPtr pickFillKernel(uint32_t threads) {
  if (threads > 512) return viterbiFillKernel<1024, 1>;  // 64 registers/thread
  if (threads > 256) return viterbiFillKernel<512, 1>;   // 128
  if (threads > 128) return viterbiFillKernel<256, 2>;   // 128
  return viterbiFillKernel<128, 4>;                      // 128
}

static cudaError_t prepFill(FillKernelPtr kern, const DevTables& tb, uint32_t smemBytes) {
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return err;
  if (tb.C > 8) err = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return err;
}

cudaError_t launchFill(const DevTables& tb, const FillArgs& args, uint32_t nClusters, uint32_t threads,
                       uint32_t smemBytes, cudaStream_t stream) {
  FillKernelPtr kern = pickFillKernel(threads);
  cudaError_t err = prepFill(kern, tb, smemBytes);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nClusters * tb.C);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tb.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tb, args);
}

cudaError_t queryMaxClusters(const DevTables& tb, uint32_t threads, uint32_t smemBytes, int* nClusters) {
  FillKernelPtr kern = pickFillKernel(threads);
  cudaError_t err = prepFill(kern, tb, smemBytes);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tb.C);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tb.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaOccupancyMaxActiveClusters(nClusters, kern, &cfg);
}

cudaError_t launchTraceback(const DevTables& tb, const TracebackArgs& args, cudaStream_t stream) {
  const int threads = 64;
  const int blocks = (int)((args.nReads + threads - 1) / threads);
  viterbiTracebackKernel<<<blocks, threads, 0, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
