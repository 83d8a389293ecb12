// Hand-written sm_100a kernels for dnastore's Viterbi hot path.
//
// What they compute is the reference's ViterbiMatrix fill + traceback
// (reference src/viterbi.cpp:62-176 and :195-304) for a BATCH of reads; how they
// compute it is B200-native:
//
//  * One thread-block CLUSTER per read.  The state space is cut into C slices (the host
//    picks a locality-preserving partition); CTA `rank` keeps its slice of the three live
//    fp64 columns -- S(pos-1), S(pos), D(pos) -- and, when they fit, the k duplication
//    columns T and its slice of the transition table in its own shared memory.  A
//    transition whose source lives in another CTA is read through distributed shared
//    memory (mapa + ld.shared::cluster), never through HBM.
//  * Per column: (1) emission step from the previous column; (2) the within-column
//    closure over null transitions and deletions as a frontier-driven, pull-style
//    chaotic relaxation -- the system is monotone, so ANY schedule reaches the same
//    least fixed point bit for bit; races are benign (values only grow
//    towards the fixed point) and use volatile / relaxed cluster-scope accesses;
//    (3) one predecessor byte per DP cell, evaluated with the TRACEBACK's own
//    floating-point association and candidate order (src/viterbi.cpp:251-286) on the
//    converged column, streamed to HBM with coalesced byte-plane stores;
//    (4) duplication opens.
//  * A second kernel walks the predecessor bytes on the device, one thread per read,
//    and emits the decoded input-symbol string, log-likelihood and status.
//
// Only fp64 add / compare / select are used on the device; every score is computed on
// the host with the reference's libm calls (include/dnab_tables.h).
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "viterbi_kernels.h"

namespace cg = cooperative_groups;

namespace dnab {

// ---------------------------------------------------------------------------
// PTX helpers.  Shared memory is addressed with 32-bit shared-window addresses so
// that local accesses are plain LDS/STS; peers are reached through mapa.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// score tables: constant after set-up and always addressed through data loaded after the set-up
// barrier, so these may be plain (reorderable, CSE-able) loads.  Never use for DP cells.
__device__ __forceinline__ double ldsTab(uint32_t a) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 ldsV2(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
// DP cells of this CTA: volatile so that every relaxation re-reads them
__device__ __forceinline__ double ldsCell(uint32_t a) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void stsCell(uint32_t a, double v) {
  asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void stsU8(uint32_t a, uint32_t v) {
  asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldsVolatile32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t atomOrShared(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atomExchShared(uint32_t a, uint32_t v) {
  uint32_t old;
  asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void redAddShared(uint32_t a, int32_t v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// DP cells of a peer CTA
__device__ __forceinline__ uint32_t mapToRank(uint32_t localAddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(localAddr), "r"(rank));
  return r;
}
__device__ __forceinline__ double ldPeerFinal(uint32_t localAddr, uint32_t rank) {  // column is final
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(mapToRank(localAddr, rank)));
  return v;
}
__device__ __forceinline__ double ldPeerRacing(uint32_t localAddr, uint32_t rank) {  // inside the closure
  double v;
  asm volatile("ld.relaxed.cluster.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(mapToRank(localAddr, rank)));
  return v;
}
// flag / control stores into a peer: made visible by the next cluster barrier
__device__ __forceinline__ void stPeerU8(uint32_t localAddr, uint32_t rank, uint32_t v) {
  asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(mapToRank(localAddr, rank)), "r"(v) : "memory");
}
__device__ __forceinline__ void stPeerU32(uint32_t localAddr, uint32_t rank, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapToRank(localAddr, rank)), "r"(v) : "memory");
}

// std::max(a,b) of the reference: keeps a on ties, no NaN handling needed
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }

__device__ __forceinline__ double negInf() { return __longlong_as_double(0xFFF0000000000000LL); }

// ---------------------------------------------------------------------------
// per-CTA context: shared-window addresses of everything the inner loops touch
// ---------------------------------------------------------------------------
struct Cta {
  uint32_t aD, aBoff, aSym, aExt, aOpen, aSub, aTsE, aFlag, aPending;
  uint32_t rank, M, tid, nThreads;
  double noGap, delOpen, delExtend, delEnd;
};

// The CTA's slice of the state blocks: in shared memory (kBS) or in global memory behind L1/L2.
template <bool kBS>
struct Blocks {
  uint32_t aBase;         // shared address of the slice (kBS)
  const uint32_t* gBase;  // global address of the slice (!kBS)
  __device__ __forceinline__ uint2 ld2(uint32_t wordOff) const {
    if (kBS) return ldsV2(aBase + wordOff * 4);
    return __ldg(reinterpret_cast<const uint2*>(gBase + wordOff));
  }
  __device__ __forceinline__ uint32_t ld1(uint32_t wordOff) const {
    if (kBS) return lds32(aBase + wordOff * 4);
    return __ldg(gBase + wordOff);
  }
  // start fetching a block that will be needed soon (global-memory tables only)
  __device__ __forceinline__ void prefetch(uint32_t wordOff) const {
    if (!kBS) asm volatile("prefetch.global.L1 [%0];" ::"l"(gBase + wordOff));
  }
};

// A cell of column `aCol` (shared address of the column in THIS CTA) addressed by an edge word.
__device__ __forceinline__ double ldEdgeFinal(uint32_t aCol, uint32_t ea) {
  const uint32_t a = aCol + edgeOff(ea);
  return (ea & kEdgeRemote) ? ldPeerFinal(a, edgeRank(ea)) : ldsCell(a);
}
__device__ __forceinline__ double ldEdgeRacing(uint32_t aCol, uint32_t ea) {
  const uint32_t a = aCol + edgeOff(ea);
  return (ea & kEdgeRemote) ? ldPeerRacing(a, edgeRank(ea)) : ldsCell(a);
}

// One relaxation of local state i (pull form of src/viterbi.cpp:118-158):
//   D(d) = max( D(d), max_emit-in ( max(D(s)+delExtend, S(s)+delOpen) + score ), max_null-in ( D(s)+score ) )
//   S(d) = max( S(d), max_null-in ( S(s)+score ), D(d)+delEnd )
// where the stored S(s) already contains D(s)+delEnd from s's own last relaxation.
// When a cell grew every successor is woken: a successor in this CTA gets its flag word
// bit set atomically in the CTA's bitmap (and the CTA's pending count incremented if it was clear); a successor in a peer
// gets a byte in that peer's remote buffer at aFlagRemote.  Returns true if a peer was woken.
template <bool kBS>
__device__ __forceinline__ bool relaxState(const Cta& c, const Blocks<kBS>& blk, uint32_t i, uint32_t aSc,
                                           uint32_t aFlagRemote) {
  const uint32_t off = lds32(c.aBoff + 4 * i);
  const uint2 h = blk.ld2(off);
  const uint32_t nE = hdrNEmit(h.x), nIn = hdrNIn(h.x);
  const uint32_t myS = aSc + 8 * i, myD = c.aD + 8 * i;
  const double oldS = ldsCell(myS), oldD = ldsCell(myD);
  double newS = oldS, newD = oldD;
  uint32_t e = off + 2;
  for (uint32_t j = 0; j < nE; ++j, e += 2) {
    const uint2 w = blk.ld2(e);
    const double ss = ldEdgeRacing(aSc, w.x), ds = ldEdgeRacing(c.aD, w.x);
    newD = dmax(newD, dmax(ds + c.delExtend, ss + c.delOpen) + ldsTab(c.aSym + edgeSymOff(w.y)));
  }
  for (uint32_t j = nE; j < nIn; ++j, e += 2) {
    const uint2 w = blk.ld2(e);
    const double ss = ldEdgeRacing(aSc, w.x), ds = ldEdgeRacing(c.aD, w.x);
    const double sc = ldsTab(c.aSym + edgeSymOff(w.y));
    newD = dmax(newD, ds + sc);
    newS = dmax(newS, ss + sc);
  }
  newS = dmax(newS, newD + c.delEnd);
  bool sentRemote = false;
  if ((newD > oldD) || (newS > oldS)) {
    if (newD > oldD) stsCell(myD, newD);
    if (newS > oldS) stsCell(myS, newS);
    __threadfence_block();  // the new cells are visible in this CTA before any successor is woken
    const uint32_t nOut = hdrNOut(h.x);
    for (uint32_t j = 0; j < nOut; ++j, ++e) {
      const uint32_t w = blk.ld1(e);
      if (w & kEdgeRemote) {
        stPeerU8(aFlagRemote + outLocal(w), edgeRank(w), 1u);
        sentRemote = true;
      } else {
        // count first, publish second: `pending` may transiently over-count but never under-count,
        // so no warp can see 0 while a bit is set or about to be set
        const uint32_t l = outLocal(w), bit = 1u << (l & 31);
        redAddShared(c.aPending, 1);
        if (atomOrShared(c.aFlag + 4 * (l >> 5), bit) & bit) redAddShared(c.aPending, -1);
      }
    }
  }
  return sentRemote;
}

template <int kMaxThreads, int kMinBlocks, bool kBS>
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
  const SmemLayout& lay = args.lay;
  const uint32_t sm = smemAddr(smem);

  Cta c;
  c.aD = sm + lay.dBuf;
  c.aBoff = sm + lay.boff;
  c.aSym = sm + lay.symScore;
  c.aExt = sm + lay.tsDext;
  c.aOpen = sm + lay.tsDopen;
  c.aSub = sm + lay.sub;
  c.aTsE = sm + lay.tsE;
  c.aFlag = sm + lay.flag;
  c.aPending = sm + lay.ctl;
  c.rank = rank;
  c.M = M;
  c.tid = tid;
  c.nThreads = nThreads;
  c.noGap = tb.noGap;
  c.delOpen = tb.delOpen;
  c.delExtend = tb.delExtend;
  c.delEnd = tb.delEnd;

  const uint32_t sliceBegin = __ldg(&tb.sliceOff[rank]), sliceEnd = __ldg(&tb.sliceOff[rank + 1]);
  Blocks<kBS> blk;
  blk.aBase = sm + lay.blocks;
  blk.gBase = tb.blocks + sliceBegin;

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
  double* sG = tb.sPrevGlobal ? args.sScratch + (size_t)clusterId * 2 * Np : nullptr;
  double* tCol = tb.tInSmem ? reinterpret_cast<double*>(smem + lay.tBuf)
                            : args.tScratch + ((size_t)clusterId * C + rank) * (size_t)k * M;

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
  for (uint32_t i = tid; i < M; i += nThreads) boff[i] = __ldg(&tb.blockOff[rank * M + i]) - sliceBegin;
  if (kBS) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + lay.blocks);
    for (uint32_t j = tid; j < sliceEnd - sliceBegin; j += nThreads) dst[j] = __ldg(tb.blocks + sliceBegin + j);
  }
  for (uint32_t j = tid; j < 64; j += nThreads) ctl[j] = 0;
  {
    uint32_t* fl = reinterpret_cast<uint32_t*>(smem + lay.flag);
    const uint32_t nFlagWords = (lay.seq - lay.flag) / 4;  // the flag arrays are contiguous
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

  unsigned long long dbgDense = 0, dbgClusterWait = 0;
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
      const uint32_t aScur = sm + (cur ? lay.sBuf[1] : lay.sBuf[0]), aSprev = sm + (prev ? lay.sBuf[1] : lay.sBuf[0]);
      // when S(pos-1) lives in global scratch: the cluster's two S columns, indexed by padded state
      const double* sPrevG = sG + (size_t)prev * Np;
      double* sCurG = sG + (size_t)cur * Np;
      const uint32_t M8 = M * 8;
      const uint32_t x = pos > 0 ? (seqS[(pos - 1) >> 2] >> (2 * ((pos - 1) & 3))) & 3u : 0u;
      const uint32_t x8 = x * 8;

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
          const uint32_t off = lds32(c.aBoff + 4 * i);
          if (i + nThreads < M) blk.prefetch(lds32(c.aBoff + 4 * (i + nThreads)));
          const uint2 h = blk.ld2(off);
          const uint32_t nE = hdrNEmit(h.x), mdl = hdrMdl(h.x);
          double t0 = NEG;
          if (mdl > 0) t0 = tCol[i];
          for (uint32_t j = 0; j < nE; j += 2) {
            const bool two = j + 1 < nE;
            const uint2 e0 = blk.ld2(off + 2 + 2 * j);
            const uint2 e1 = two ? blk.ld2(off + 4 + 2 * j) : e0;
            double v0, v1 = NEG;
            if (tb.sPrevGlobal) {  // plain coherent loads: the column was published before the last cluster barrier
              v0 = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(sPrevG) + edgeRank(e0.x) * M8 + edgeOff(e0.x));
              if (two) v1 = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(sPrevG) + edgeRank(e1.x) * M8 + edgeOff(e1.x));
            } else {
              v0 = ldEdgeFinal(aSprev, e0.x);
              if (two) v1 = ldEdgeFinal(aSprev, e1.x);
            }
            const double c0 = ((v0 + ldsTab(c.aSym + edgeSymOff(e0.y))) + c.noGap) + ldsTab(c.aSub + edgeSubOff(e0.y) + x8);
            const double c1 = ((v1 + ldsTab(c.aSym + edgeSymOff(e1.y))) + c.noGap) + ldsTab(c.aSub + edgeSubOff(e1.y) + x8);
            s = dmax(s, c0);
            s = dmax(s, c1);  // v1 = -inf when the pair is incomplete
          }
          if (mdl > 0) {
            const double t2s = t0 + subS[hdrCtx(h.y, 0) * 4 + x];
            s = dmax(s, t2s);
            for (uint32_t j = 0; j + 1 < mdl; ++j) tCol[j * M + i] = tCol[(j + 1) * M + i] + subS[hdrCtx(h.y, j + 1) * 4 + x];
            tCol[(mdl - 1) * M + i] = t2s;  // slot mdl-1 is free until step (4): park the T->S candidate there
          }
        }
        stsCell(aScur + 8 * i, s);
        stsCell(c.aD + 8 * i, NEG);
      }
      if (tid == 0) ctl[0] = nThreads >> 5;  // closure: one token per warp, returned after its first sweep
      clusterBarrier();
      long long tc1 = dbgOn ? clock64() : 0;

      // ---- (2) closure: null transitions + deletions (src/viterbi.cpp:110-159) ----
      // Inside a CTA the relaxation is ASYNCHRONOUS.  Warp w owns the 32-state groups w, w+nWarps,
      // ... and one word of the CTA's dirty bitmap per group.  It first relaxes all its states once,
      // then keeps sweeping: it takes the set bits of its words (atomic exchange), compacts the
      // dirty states into a warp-private list and relaxes them 32 at a time, one state per lane.
      // A state that grows sets the bits of its successors (atomic OR); `pending` counts bits that
      // are set or being served, so pending == 0 means the CTA is at a fixed point -- no CTA
      // barrier separates the hops of a propagation chain.  Successors owned by a peer CTA are
      // recorded in that peer's remote buffer of the current round; the cluster meets at a
      // barrier when every CTA is locally quiet, the recorded states are flagged, and the next
      // round starts.  The cluster is done when a whole round woke nobody across CTAs.
      {
        const uint32_t warp = tid >> 5, lane = tid & 31, nWarps = nThreads >> 5;
        const uint32_t nGroups = (M + 31) / 32;
        const uint32_t myGroups = nGroups > warp ? (nGroups - warp + nWarps - 1) / nWarps : 0;  // <= 32 * 32 states
        uint16_t* myList = reinterpret_cast<uint16_t*>(smem + lay.list) + warp * ((nGroups + nWarps - 1) / nWarps) * 32;
        uint32_t round = 0;
        bool sent = false;
        // round 0, first sweep: every state once
        for (uint32_t g = warp; g < nGroups; g += nWarps) {
          const uint32_t i = 32 * g + lane;
          if (i + 32 * nWarps < M) blk.prefetch(lds32(c.aBoff + 4 * (i + 32 * nWarps)));
          if (i < M) sent |= relaxState<kBS>(c, blk, i, aScur, sm + lay.flagRemote[0]);
        }
        __syncwarp();
        if (lane == 0) redAddShared(c.aPending, -1);  // pending == 0 now also means every warp swept once
        if (dbgOn && tid == 0) dbgDense += clock64() - tc1;
        for (;;) {
          const uint32_t aRemoteOut = sm + ((round & 1) ? lay.flagRemote[1] : lay.flagRemote[0]);
          // `pending` is exact only when nobody is relaxing: a warp that reads 0 parks at the CTA barrier
          // below, where the count is re-read once every warp has arrived (all atomics completed)
          uint32_t spins = 0;
          do {
            for (;;) {
              // the exit decision must be warp-uniform: the sweep below uses warp collectives
              __syncwarp();
              uint32_t p = lane == 0 ? ldsVolatile32(c.aPending) : 0u;
              p = __shfl_sync(0xFFFFFFFFu, p, 0);
              if (p == 0u) break;
              if (++spins > (1u << 22)) __trap();  // never hang the GPU: a stuck closure is a bug
              // take this warp's dirty bits: lane j looks after the warp's j-th group (myGroups <= 32)
              uint32_t taken = 0;
              bool didWork = false;
              for (uint32_t base = 0; base < myGroups; base += 32) {
                const uint32_t j = base + lane;
                const uint32_t aWord = c.aFlag + 4 * (warp + j * nWarps);
                taken = (j < myGroups && ldsVolatile32(aWord)) ? atomExchShared(aWord, 0u) : 0u;
                const uint32_t cnt = __popc(taken);
                uint32_t incl = cnt;  // inclusive prefix sum over the lanes
  #pragma unroll
                for (uint32_t d = 1; d < 32; d <<= 1) {
                  const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                  if (lane >= d) incl += up;
                }
                const uint32_t n = __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (n == 0) continue;
                didWork = true;
                uint32_t at = incl - cnt;
                const uint32_t stateBase = 32 * (warp + j * nWarps);
                while (taken) {
                  const uint32_t b = __ffs(taken) - 1;
                  taken &= taken - 1;
                  myList[at++] = (uint16_t)(stateBase + b);
                }
                __syncwarp();
                __threadfence_block();  // cells written before these bits were set are visible from here on
                for (uint32_t q = lane + 32; q < n; q += 32) blk.prefetch(lds32(c.aBoff + 4 * myList[q]));
              for (uint32_t q = lane; q < n; q += 32) sent |= relaxState<kBS>(c, blk, myList[q], aScur, aRemoteOut);
                __syncwarp();
                if (lane == 0) redAddShared(c.aPending, -(int32_t)n);  // only now: the successors they woke are counted
                if (dbgOn && tid == 0) dbgWork += n;
              }
              if (dbgOn && tid == 0) dbgIters++;
              // an idle warp backs off: spinning warps would steal issue slots from the warps that relax
              if (!didWork) __nanosleep(args.idleSleepNs);
            }
          } while (__syncthreads_or(ldsVolatile32(c.aPending) != 0u));
          if (C == 1) break;
          long long tcb = dbgOn ? clock64() : 0;
          const uint32_t anySent = (uint32_t)__syncthreads_or(sent ? 1 : 0);
          if (tid < C) stPeerU32(sm + lay.ctl + (16 + (round & 1) * kMaxCluster + rank) * 4, tid, anySent);
          cluster.sync();
          if (dbgOn && tid == 0) dbgClusterWait += clock64() - tcb;
          uint32_t tot = 0;
          for (uint32_t r = 0; r < C; ++r) tot |= ctl[16 + (round & 1) * kMaxCluster + r];
          if (!tot) break;
          if (dbgOn && tid == 0) dbgRounds++;
          // flag the states peers woke during the round just completed; peers now fill the other buffer
          uint32_t* fr = reinterpret_cast<uint32_t*>(smem + ((round & 1) ? lay.flagRemote[1] : lay.flagRemote[0]));
          for (uint32_t w = tid; w < (M + 3) / 4; w += nThreads) {
            const uint32_t f = fr[w];
            if (f) {
              fr[w] = 0;
#pragma unroll
              for (uint32_t q = 0; q < 4; ++q)
                if ((f >> (8 * q)) & 0xFFu) {
                  const uint32_t l = 4 * w + q, bit = 1u << (l & 31);
                  redAddShared(c.aPending, 1);
                  if (atomOrShared(c.aFlag + 4 * (l >> 5), bit) & bit) redAddShared(c.aPending, -1);
                }
            }
          }
          ++round;
          sent = false;
          __syncthreads();
        }
        __syncthreads();  // every warp has left the sweep before the columns are read as final
      }

      long long tc2 = dbgOn ? clock64() : 0;
      // ---- (3) predecessor records with the traceback's arithmetic (src/viterbi.cpp:251-286)
      //      (4) duplication opens (src/viterbi.cpp:161-168) ----
      uint8_t* predCol = predRead + (size_t)pos * (k + 2) * Np;
      for (uint32_t i = tid; i < M; i += nThreads) {
        const uint32_t g = rank * M + i;
        const uint32_t off = lds32(c.aBoff + 4 * i);
        if (i + nThreads < M) blk.prefetch(lds32(c.aBoff + 4 * (i + nThreads)));
        const uint2 h = blk.ld2(off);
        const uint32_t nE = hdrNEmit(h.x), nIn = hdrNIn(h.x), mdl = hdrMdl(h.x);
        const double sHere = ldsCell(aScur + 8 * i), dHere = ldsCell(c.aD + 8 * i);
        if (tb.sPrevGlobal) sCurG[g] = sHere;  // publish the converged column for the next position
        const double parked = (mdl > 0 && pos > 0) ? tCol[(mdl - 1) * M + i] : NEG;  // T(state,pos-1,0)+sub

        double best = NEG, bestD = NEG;
        uint32_t idx = kNoPred, idxD = kNoPred;
        for (uint32_t e = 0; e < nIn; ++e) {
          const uint2 ew = blk.ld2(off + 2 + 2 * e);
          const double vs = ldEdgeRacing(aScur, ew.x), vd = ldEdgeRacing(c.aD, ew.x);
          if (e < nE) {
            if (pos > 0) {
              const double vp = tb.sPrevGlobal
                                    ? *reinterpret_cast<const double*>(reinterpret_cast<const char*>(sPrevG) + edgeRank(ew.x) * M8 + edgeOff(ew.x))
                                    : ldEdgeFinal(aSprev, ew.x);
              const double v = vp + ldsTab(c.aTsE + edgeTsEOff(ew.y) + x8);
              if (v > best) {
                best = v;
                idx = e;
              }
            }
            const double ve = vd + ldsTab(c.aExt + edgeSymOff(ew.y));
            if (ve > bestD) {
              bestD = ve;
              idxD = 2 * e;
            }
            const double vo = vs + ldsTab(c.aOpen + edgeSymOff(ew.y));
            if (vo > bestD) {
              bestD = vo;
              idxD = 2 * e + 1;
            }
          } else {
            const double sc = ldsTab(c.aSym + edgeSymOff(ew.y));
            const double v = vs + sc;
            if (v > best) {
              best = v;
              idx = e;
            }
            const double vn = vd + sc;
            if (vn > bestD) {
              bestD = vn;
              idxD = nE + e;  // = 2*nE + (e - nE)
            }
          }
        }
        {
          const double v = dHere + c.delEnd;
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
          const uint32_t a = aScur + (tb.startG % M) * 8, r = tb.startG / M;
          const double v = (r == rank ? ldsCell(a) : ldPeerRacing(a, r)) + 0.0;
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
      const double* sLast = reinterpret_cast<const double*>(smem + ((L & 1) ? lay.sBuf[1] : lay.sBuf[0]));
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
        for (int sh = 16; sh > 0; sh >>= 1) {
          const double ov = __shfl_down_sync(0xFFFFFFFFu, bv, sh);
          const uint32_t oo = __shfl_down_sync(0xFFFFFFFFu, bo, sh);
          const uint32_t og = __shfl_down_sync(0xFFFFFFFFu, bg, sh);
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
    const uint32_t* blkp = tb.blocks + tb.blockOff[g];
    const uint32_t nE = hdrNEmit(blkp[0]), nIn = hdrNIn(blkp[0]);
    const uint2* edges = reinterpret_cast<const uint2*>(blkp + 2);
    auto srcOf = [&](uint2 e) { return edgeRank(e.x) * M + edgeOff(e.x) / 8; };
    uint32_t sym = 0;
    if (mut == 0) {
      if (p < nIn) {  // incoming emit edge (previous column) or null edge (same column)
        const uint2 e = edges[p];
        sym = edgeSymOff(e.y) / 8;
        g = srcOf(e);
        if (p < nE) --pos;
      } else if (p == nIn) {
        mut = 1;
      } else if (p == nIn + 1) {
        mut = 2;
        --pos;
      } else {
        g = tb.startG;  // local mode, pos == 0: jump to (0,0,S)
      }
    } else if (mut == 1) {
      if (p < 2 * nE) {
        const uint2 e = edges[p >> 1];
        sym = edgeSymOff(e.y) / 8;
        g = srcOf(e);
        mut = (p & 1) ? 0 : 1;
      } else {
        const uint2 e = edges[p - nE];  // null edge number p-2*nE sits after the nE emit edges
        sym = edgeSymOff(e.y) / 8;
        g = srcOf(e);
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
static FillKernelPtr pickFillKernel(uint32_t threads, bool blocksInSmem) {
  if (blocksInSmem) {
    if (threads > 512) return viterbiFillKernel<1024, 1, true>;  // 64 registers/thread
    if (threads > 256) return viterbiFillKernel<512, 1, true>;   // 128
    if (threads > 128) return viterbiFillKernel<256, 2, true>;   // 128
    return viterbiFillKernel<128, 4, true>;                      // 128
  }
  if (threads > 512) return viterbiFillKernel<1024, 1, false>;
  if (threads > 256) return viterbiFillKernel<512, 1, false>;
  if (threads > 128) return viterbiFillKernel<256, 2, false>;
  return viterbiFillKernel<128, 4, false>;
}

static cudaError_t prepFill(FillKernelPtr kern, const DevTables& tb, uint32_t smemBytes) {
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return err;
  if (tb.C > 8) err = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return err;
}

static void clusterConfig(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, const DevTables& tb, uint32_t nClusters,
                          uint32_t threads, uint32_t smemBytes, cudaStream_t stream) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(nClusters * tb.C);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smemBytes;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tb.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

cudaError_t launchFill(const DevTables& tb, const FillArgs& args, uint32_t nClusters, uint32_t threads,
                       uint32_t smemBytes, cudaStream_t stream) {
  FillKernelPtr kern = pickFillKernel(threads, tb.blocksInSmem != 0);
  cudaError_t err = prepFill(kern, tb, smemBytes);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  clusterConfig(cfg, attr, tb, nClusters, threads, smemBytes, stream);
  return cudaLaunchKernelEx(&cfg, kern, tb, args);
}

cudaError_t queryMaxClusters(const DevTables& tb, uint32_t threads, uint32_t smemBytes, int* nClusters) {
  FillKernelPtr kern = pickFillKernel(threads, tb.blocksInSmem != 0);
  cudaError_t err = prepFill(kern, tb, smemBytes);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  clusterConfig(cfg, attr, tb, 1, threads, smemBytes, nullptr);
  return cudaOccupancyMaxActiveClusters(nClusters, kern, &cfg);
}

cudaError_t launchTraceback(const DevTables& tb, const TracebackArgs& args, cudaStream_t stream) {
  const int threads = 64;
  const int blocks = (int)((args.nReads + threads - 1) / threads);
  viterbiTracebackKernel<<<blocks, threads, 0, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
