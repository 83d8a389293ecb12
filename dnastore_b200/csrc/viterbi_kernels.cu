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
  asm volatile("ld.relaxed.cluster.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void stRelaxedF64(uint32_t addr, double v) {
  asm volatile("st.relaxed.cluster.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void stRelaxedU8(uint32_t addr, uint32_t v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void stRelaxedU32(uint32_t addr, uint32_t v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// std::max(a,b) of the reference: keeps a on ties, no NaN handling needed
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }

__device__ __forceinline__ double negInf() { return __longlong_as_double(0xFFF0000000000000LL); }

// ---------------------------------------------------------------------------
// shared-memory carve-up (identical in every CTA of a cluster, which is what lets
// a local offset be mapped into a peer with mapa)
// ---------------------------------------------------------------------------
struct SmemLayout {
  uint32_t sBuf[2];   // byte offsets of the two S columns
  uint32_t dBuf;
  uint32_t tBuf;      // k*M doubles, only if tInSmem
  uint32_t tsE;       // [nSyms*16] (score+noGap)+sub  -- traceback association, src/viterbi.cpp:255
  uint32_t symScore;  // [kMaxSyms]
  uint32_t tsDext;    // [kMaxSyms] score+delExtend    -- src/viterbi.cpp:272
  uint32_t tsDopen;   // [kMaxSyms] score+delOpen      -- src/viterbi.cpp:273
  uint32_t sub;       // [16]
  uint32_t tsT;       // [kMaxK] tanDup+len[i]         -- src/viterbi.cpp:286
  uint32_t len;       // [kMaxK]
  uint32_t chg;       // [2][kMaxCluster] u32
  uint32_t count;     // u32 (+pad)
  uint32_t work;      // [M] u32
  uint32_t dirty[2];  // [Mpad] u8 each
  uint32_t seq;       // packed read
  uint32_t total;
};

__host__ __device__ inline SmemLayout makeLayout(uint32_t M, uint32_t k, uint32_t tInSmem, uint32_t maxLen) {
  SmemLayout L;
  uint32_t at = 0;
  auto take = [&](uint32_t bytes) {
    const uint32_t here = at;
    at += (bytes + 15u) & ~15u;
    return here;
  };
  L.sBuf[0] = take(M * 8);
  L.sBuf[1] = take(M * 8);
  L.dBuf = take(M * 8);
  L.tBuf = tInSmem ? take(k * M * 8) : 0;
  L.tsE = take(kMaxSyms * 16 * 8);
  L.symScore = take(kMaxSyms * 8);
  L.tsDext = take(kMaxSyms * 8);
  L.tsDopen = take(kMaxSyms * 8);
  L.sub = take(16 * 8);
  L.tsT = take(8 * 8);
  L.len = take(8 * 8);
  L.chg = take(2 * kMaxCluster * 4);
  L.count = take(16);
  L.work = take(M * 4);
  L.dirty[0] = take(M);
  L.dirty[1] = take(M);
  L.seq = take((maxLen + 3) / 4 + 16);
  L.total = at;
  return L;
}

uint32_t fillSmemBytes(uint32_t M, uint32_t k, uint32_t tInSmem, uint32_t maxLen) {
  return makeLayout(M, k, tInSmem, maxLen).total;
}

// ---------------------------------------------------------------------------
// fill kernel
// ---------------------------------------------------------------------------
struct ColumnCtx {
  // per-CTA constants
  const DevTables* tb;
  unsigned char* smem;
  uint32_t smemBase;  // shared-window address of smem[0]
  SmemLayout lay;
  uint32_t rank, M, C, k;
};

// One relaxation of state `i` of this CTA's slice (pull form of src/viterbi.cpp:118-158):
//   D(d) = max( D(d), max_emit-in ( max(D(s)+delExtend, S(s)+delOpen) + score ), max_null-in ( D(s)+score ) )
//   S(d) = max( S(d), max_null-in ( S(s)+score ), D(d)+delEnd )
// where the stored S(s) already contains D(s)+delEnd from s's own last relaxation.
// Returns true (and marks every successor dirty in `markBuf`) when a cell grew.
__device__ __forceinline__ bool relaxState(const ColumnCtx& c, uint32_t i, uint32_t sCurOff, uint32_t markBuf) {
  const DevTables& tb = *c.tb;
  const uint32_t g = c.rank * c.M + i;
  const uint2 rec = __ldg(&tb.stateRec[g]);
  const uint32_t nE = recNEmit(rec.y), nN = recNNull(rec.y);
  const double* symScore = reinterpret_cast<const double*>(c.smem + c.lay.symScore);
  const uint32_t myS = c.smemBase + sCurOff + i * 8, myD = c.smemBase + c.lay.dBuf + i * 8;
  const double oldS = ldRelaxedF64(myS), oldD = ldRelaxedF64(myD);
  double newS = oldS, newD = oldD;
  const uint32_t dMinusS = c.lay.dBuf - sCurOff;
  const uint32_t* edges = tb.inEdges + rec.x;
  for (uint32_t e = 0; e < nE; ++e) {
    const uint32_t w = __ldg(&edges[e]);
    const uint32_t aS = mapToRank(c.smemBase + sCurOff + edgeLocal(w) * 8, edgeRank(w));
    const double ss = ldRelaxedF64(aS), ds = ldRelaxedF64(aS + dMinusS);
    const double cand = dmax(ds + tb.delExtend, ss + tb.delOpen) + symScore[edgeSym(w)];
    newD = dmax(newD, cand);
  }
  for (uint32_t e = 0; e < nN; ++e) {
    const uint32_t w = __ldg(&edges[nE + e]);
    const uint32_t aS = mapToRank(c.smemBase + sCurOff + edgeLocal(w) * 8, edgeRank(w));
    const double ss = ldRelaxedF64(aS), ds = ldRelaxedF64(aS + dMinusS);
    const double sc = symScore[edgeSym(w)];
    newD = dmax(newD, ds + sc);
    newS = dmax(newS, ss + sc);
  }
  newS = dmax(newS, newD + tb.delEnd);
  const bool grew = (newD > oldD) || (newS > oldS);
  if (grew) {
    if (newD > oldD) stRelaxedF64(myD, newD);
    if (newS > oldS) stRelaxedF64(myS, newS);
    const uint32_t o0 = __ldg(&tb.outOff[g]), o1 = __ldg(&tb.outOff[g + 1]);
    for (uint32_t o = o0; o < o1; ++o) {
      const uint32_t w = __ldg(&tb.outEdges[o]);
      stRelaxedU8(mapToRank(c.smemBase + c.lay.dirty[markBuf] + edgeLocal(w), edgeRank(w)), 1u);
    }
  }
  return grew;
}

__global__ void __launch_bounds__(1024, 1) viterbiFillKernel(const __grid_constant__ DevTables tb, const FillArgs args) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t C = tb.C, M = tb.M, k = tb.k;
  const uint32_t rank = C > 1 ? cluster.block_rank() : 0;
  const uint32_t clusterId = blockIdx.x / C;
  const uint32_t nClusters = gridDim.x / C;
  const uint32_t tid = threadIdx.x, nThreads = blockDim.x;
  const uint32_t Np = C * M;
  const double NEG = negInf();

  ColumnCtx c;
  c.tb = &tb;
  c.smem = smem;
  c.smemBase = smemAddr(smem);
  c.lay = makeLayout(M, k, tb.tInSmem, args.maxLen);
  c.rank = rank;
  c.M = M;
  c.C = C;
  c.k = k;
  const SmemLayout& lay = c.lay;

  double* symScore = reinterpret_cast<double*>(smem + lay.symScore);
  double* tsE = reinterpret_cast<double*>(smem + lay.tsE);
  double* tsDext = reinterpret_cast<double*>(smem + lay.tsDext);
  double* tsDopen = reinterpret_cast<double*>(smem + lay.tsDopen);
  double* subS = reinterpret_cast<double*>(smem + lay.sub);
  double* tsT = reinterpret_cast<double*>(smem + lay.tsT);
  double* lenS = reinterpret_cast<double*>(smem + lay.len);
  volatile uint32_t* chg = reinterpret_cast<volatile uint32_t*>(smem + lay.chg);
  uint32_t* count = reinterpret_cast<uint32_t*>(smem + lay.count);
  uint32_t* work = reinterpret_cast<uint32_t*>(smem + lay.work);
  uint8_t* seqS = smem + lay.seq;
  double* dCol = reinterpret_cast<double*>(smem + lay.dBuf);
  double* tCol = tb.tInSmem ? reinterpret_cast<double*>(smem + lay.tBuf)
                            : args.tScratch + ((size_t)clusterId * C + rank) * (size_t)k * M;

  // read-independent score tables; the traceback-association sums are formed here once
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
  for (uint32_t j = tid; j < tb.nSyms * 16; j += nThreads) {
    const uint32_t s = j >> 4, bx = j & 15;
    tsE[j] = (tb.symScore[s] + tb.noGap) + tb.sub[bx];
  }
  __syncthreads();

  auto clusterBarrier = [&]() {
    if (C > 1)
      cluster.sync();
    else
      __syncthreads();
  };

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

      // ---- (1) emission step: S0 from the previous column, T shift (src/viterbi.cpp:92-106) ----
      for (uint32_t i = tid; i < M; i += nThreads) {
        const uint32_t g = rank * M + i;
        const uint2 rec = __ldg(&tb.stateRec[g]);
        const uint32_t nE = recNEmit(rec.y), mdl = recMdl(rec.y);
        double s = NEG;
        if (pos == 0) {
          const bool real = __ldg(&tb.origId[g]) != 0xFFFFFFFFu;
          s = (real && (tb.local || g == tb.startG)) ? 0.0 : NEG;  // src/viterbi.cpp:75-79
          for (uint32_t j = 0; j < k; ++j) tCol[j * M + i] = NEG;
        } else {
          const uint32_t* edges = tb.inEdges + rec.x;
          for (uint32_t e = 0; e < nE; ++e) {
            const uint32_t w = __ldg(&edges[e]);
            const double v = ldClusterF64(mapToRank(c.smemBase + sPrevOff + edgeLocal(w) * 8, edgeRank(w)));
            const double cand = ((v + symScore[edgeSym(w)]) + tb.noGap) + subS[edgeBase(w) * 4 + x];
            s = dmax(s, cand);
          }
          if (mdl > 0) {
            const double t2s = tCol[i] + subS[recCtx(rec.y, 0) * 4 + x];
            s = dmax(s, t2s);
            for (uint32_t j = 0; j + 1 < mdl; ++j) tCol[j * M + i] = tCol[(j + 1) * M + i] + subS[recCtx(rec.y, j + 1) * 4 + x];
            tCol[(mdl - 1) * M + i] = t2s;  // slot mdl-1 is free until step (4): park the T->S candidate there
          }
        }
        sCur[i] = s;
        dCol[i] = NEG;
        smem[lay.dirty[0] + i] = 0;
        smem[lay.dirty[1] + i] = 0;
      }
      clusterBarrier();

      // ---- (2) closure: null transitions + deletions (src/viterbi.cpp:110-159) ----
      bool grew = false;
      for (uint32_t i = tid; i < M; i += nThreads) grew |= relaxState(c, i, sCurOff, 1);  // sweep 1: every state
      for (uint32_t sweep = 1;; ++sweep) {
        const uint32_t par = sweep & 1;
        const int anyLocal = __syncthreads_or(grew ? 1 : 0);
        uint32_t any = (uint32_t)anyLocal;
        if (C > 1) {
          if (tid < C) stRelaxedU32(mapToRank(c.smemBase + lay.chg + (par * kMaxCluster + rank) * 4, tid), any);
          cluster.sync();
          any = 0;
          for (uint32_t r = 0; r < C; ++r) any |= chg[par * kMaxCluster + r];
        }
        if (!any) break;
        // compact this CTA's dirty flags into a dense worklist
        if (tid == 0) *count = 0;
        __syncthreads();
        uint8_t* flags = smem + lay.dirty[par];
        for (uint32_t base = 0; base < M; base += nThreads) {
          const uint32_t i = base + tid;
          const bool set = i < M && flags[i];
          if (set) flags[i] = 0;
          const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, set);
          if (ballot) {
            const uint32_t lane = tid & 31;
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(count, __popc(ballot));
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            if (set) work[at + __popc(ballot & ((1u << lane) - 1))] = i;
          }
        }
        __syncthreads();
        const uint32_t n = *count;
        grew = false;
        for (uint32_t j = tid; j < n; j += nThreads) grew |= relaxState(c, work[j], sCurOff, par ^ 1);
      }

      // ---- (3) predecessor records with the traceback's arithmetic (src/viterbi.cpp:251-286)
      //      (4) duplication opens (src/viterbi.cpp:161-168) ----
      uint8_t* predCol = predRead + (size_t)pos * (k + 2) * Np;
      for (uint32_t i = tid; i < M; i += nThreads) {
        const uint32_t g = rank * M + i;
        const uint2 rec = __ldg(&tb.stateRec[g]);
        const uint32_t nE = recNEmit(rec.y), nN = recNNull(rec.y), mdl = recMdl(rec.y);
        const uint32_t* edges = tb.inEdges + rec.x;
        const double sHere = sCur[i], dHere = dCol[i];
        const uint32_t dMinusS = lay.dBuf - sCurOff;

        double best = NEG;
        uint32_t idx = kNoPred;
        double bestD = NEG;
        uint32_t idxD = kNoPred;
        for (uint32_t e = 0; e < nE; ++e) {
          const uint32_t w = __ldg(&edges[e]);
          const uint32_t sym = edgeSym(w);
          if (pos > 0) {
            const double v = ldClusterF64(mapToRank(c.smemBase + sPrevOff + edgeLocal(w) * 8, edgeRank(w))) +
                             tsE[sym * 16 + edgeBase(w) * 4 + x];
            if (v > best) {
              best = v;
              idx = e;
            }
          }
          const uint32_t aS = mapToRank(c.smemBase + sCurOff + edgeLocal(w) * 8, edgeRank(w));
          const double vd = ldClusterF64(aS + dMinusS) + tsDext[sym];
          if (vd > bestD) {
            bestD = vd;
            idxD = 2 * e;
          }
          const double vs = ldClusterF64(aS) + tsDopen[sym];
          if (vs > bestD) {
            bestD = vs;
            idxD = 2 * e + 1;
          }
        }
        for (uint32_t e = 0; e < nN; ++e) {
          const uint32_t w = __ldg(&edges[nE + e]);
          const double sc = symScore[edgeSym(w)];
          const uint32_t aS = mapToRank(c.smemBase + sCurOff + edgeLocal(w) * 8, edgeRank(w));
          const double v = ldClusterF64(aS) + sc;
          if (v > best) {
            best = v;
            idx = nE + e;
          }
          const double vd = ldClusterF64(aS + dMinusS) + sc;
          if (vd > bestD) {
            bestD = vd;
            idxD = 2 * nE + e;
          }
        }
        {
          const double v = dHere + tb.delEnd;
          if (v > best) {
            best = v;
            idx = nE + nN;
          }
        }
        if (mdl > 0 && pos > 0) {
          const double v = tCol[(mdl - 1) * M + i];  // parked T(state,pos-1,0)+sub
          if (v > best) {
            best = v;
            idx = nE + nN + 1;
          }
        }
        if (tb.local && pos == 0) {
          const double v = ldClusterF64(mapToRank(c.smemBase + sCurOff + (tb.startG % M) * 8, tb.startG / M)) + 0.0;
          if (v > best) {
            best = v;
            idx = nE + nN + 2;
          }
        }
        const bool real = __ldg(&tb.origId[g]) != 0xFFFFFFFFu;
        predCol[g] = (uint8_t)(real ? idx : kNoPred);
        predCol[Np + g] = (uint8_t)(real ? idxD : kNoPred);
        for (uint32_t j = 0; j < k; ++j) {
          uint32_t idxT = kNoPred;
          if (pos > 0 && j < mdl) {
            double bestT = NEG;
            const double shifted = (j + 1 < mdl) ? tCol[j * M + i] : NEG;
            if (j + 1 < mdl && shifted > bestT) {
              bestT = shifted;
              idxT = 0;
            }
            const double open = sHere + tsT[j];
            if (open > bestT) idxT = 1;
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
        // reuse the worklist area as reduction scratch: [warp] -> (value, orig, g)
        double* rv = reinterpret_cast<double*>(smem + lay.tsE);  // >= 32 doubles, free between reads
        uint32_t* ro = work;
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
        // tsE was clobbered: rebuild it before the next read
        for (uint32_t j = tid; j < tb.nSyms * 16; j += nThreads) {
          const uint32_t s = j >> 4, bx = j & 15;
          tsE[j] = (tb.symScore[s] + tb.noGap) + tb.sub[bx];
        }
        __syncthreads();
      }
    }
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
    const uint2 rec = tb.stateRec[g];
    const uint32_t nE = recNEmit(rec.y), nN = recNNull(rec.y);
    const uint32_t* edges = tb.inEdges + rec.x;
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
cudaError_t launchFill(const DevTables& tb, const FillArgs& args, uint32_t nClusters, uint32_t threads,
                       uint32_t smemBytes, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(viterbiFillKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return err;
  if (tb.C > 8) {
    err = cudaFuncSetAttribute(viterbiFillKernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return err;
  }
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
  return cudaLaunchKernelEx(&cfg, viterbiFillKernel, tb, args);
}

cudaError_t queryMaxClusters(const DevTables& tb, uint32_t threads, uint32_t smemBytes, int* nClusters) {
  cudaError_t err = cudaFuncSetAttribute(viterbiFillKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return err;
  if (tb.C > 8) {
    err = cudaFuncSetAttribute(viterbiFillKernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (err != cudaSuccess) return err;
  }
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
  return cudaOccupancyMaxActiveClusters(nClusters, viterbiFillKernel, &cfg);
}

cudaError_t launchTraceback(const DevTables& tb, const TracebackArgs& args, cudaStream_t stream) {
  const int threads = 64;
  const int blocks = (int)((args.nReads + threads - 1) / threads);
  viterbiTracebackKernel<<<blocks, threads, 0, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
