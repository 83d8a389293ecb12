// viterbiFillBatchKernel: the ViterbiMatrix fill (reference src/viterbi.cpp:62-176) with the READS AS THE
// SIMD LANES.  A group of 32 reads walks the machine together: lane r of every warp works on read r of the
// group and a warp works on one state at a time.  What that buys over one-read-per-cluster
// (viterbi_fill_push.cu, 45 warp-instructions per state-column per read, < 6 % of them fp64):
//   * a transition word, its score and its addresses are decoded ONCE per warp and serve 32 reads;
//   * a gather of S(src), D(src) is one 512-byte row of shared memory, conflict free (one ld.shared.v2.f64);
//   * the closure's control flow (which states are dirty, which transitions to relax) is warp-uniform:
//     no divergence, no per-lane queues; lanes differ only in the fp64 values they add and compare;
//   * records leave as full 32-byte sectors: pred[group][pos][state][cell kind][32 reads].
// The closure is the same least fixed point as the reference's worklist (src/viterbi.cpp:110-159; any fair
// schedule reaches the same bits), organised as OWNER-COMPUTES edge relaxation: every state
// has a 32-bit work mask, bit j = "in-transition j's source grew (for some read of the group)".  The warp
// that owns a state relaxes exactly the flagged transitions (reading the source's row, writing only its own
// row: no atomics on DP cells, no lost updates), and if any lane grew it flags the corresponding bit of
// each successor's mask.  Levels are breadth first (level 0 relaxes every transition of every state), one
// CTA barrier per level.  tools/union_frontier.py measured what sharing one frontier among 32 reads costs:
// 2.9-7.6 state visits per column instead of 2.0-2.9 per read, i.e. 0.09-0.24 warp-level visits per read.
//
// A TEAM of T CTAs holds the (S,D) columns of one group in shared memory, M = ceil(N/T) states each
// (M*512 bytes); T = 1 for dnastore-l4, the whole GPU for the 46,670-state BASELINE machine.  Nothing
// crosses CTAs through shared memory: a state with successors in other CTAs PUBLISHES its (S,D) row to an
// L2-resident array whenever it grows and then flags the successor's bit in the owner's INBOX mask; the
// owner drains its inbox when it runs out of local work, and the team meets at a counter barrier in L2
// that also tells whether anybody flagged a remote bit since the last meeting.  Nothing is limited by the
// 16-CTA cluster size, so machines that do not fit a cluster use the same code path.
//
// After the closure one dense pass per column (a) evaluates the predecessor records with the TRACEBACK's
// own association and candidate order (src/viterbi.cpp:251-286), first strict maximum, one byte per DP
// cell; (b) opens duplications (src/viterbi.cpp:161-168); (c) performs the emission step of the NEXT column
// (src/viterbi.cpp:92-106) from the same gathered S(pos)[src] rows.
#include <cuda_runtime.h>

#include <cstdint>

#include "viterbi_batch.h"
#include "viterbi_kernels.h"

namespace dnab {
namespace {

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double negInf() { return __longlong_as_double(0xFFF0000000000000LL); }
// std::max(a,b) of the reference: keeps a on ties
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }

// (S,D) of one read of one state: 16 bytes; a warp reads a 512-byte row.  volatile: other warps write rows.
__device__ __forceinline__ double2 ldsRow(uint32_t a) {
  double2 v;
  asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void stsRow(uint32_t a, double s, double d) {
  asm volatile("st.volatile.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(s), "d"(d) : "memory");
}
// rows published through L2 for readers on other SMs: never through the reader's L1
__device__ __forceinline__ double2 ldPub2(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stPub2(double2* p, double s, double d) {
  asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(s), "d"(d) : "memory");
}
__device__ __forceinline__ double ldCg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stCg(double* p, double v) { asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
// predecessor records are written once and read by another kernel: streaming stores
__device__ __forceinline__ void stRecord(uint8_t* p, uint32_t v) {
  asm volatile("st.global.cs.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldVolatileGlobal32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ldAcquire64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

enum : uint32_t { kCtlFlag = 0, kCtlAny = 4, kCtlSent = 5 };

}  // namespace

template <int W, bool kDebug>
__global__ void __launch_bounds__(W * 32, 1)
    viterbiFillBatchKernel(const __grid_constant__ BatchTables tb, const __grid_constant__ BatchArgs args) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr uint32_t nThreads = W * 32;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t T = tb.T, M = tb.M, k = tb.k, Np = T * M, N = tb.nStates, K2 = k + 2;
  const uint32_t team = blockIdx.x / T, rank = blockIdx.x - team * T;
  const uint32_t nSlots = (M + W - 1) / W;
  const double NEG = negInf();
  const BatchLayout lay = makeBatchLayout(M, tb.maxIn, tb.maxOut, W);

  const uint32_t aSD = smemAddr(smem + lay.sd);
  uint32_t* maskCur = reinterpret_cast<uint32_t*>(smem + lay.maskA);
  uint32_t* maskNext = reinterpret_cast<uint32_t*>(smem + lay.maskB);
  uint4* hdrS = reinterpret_cast<uint4*>(smem + lay.hdr);
  uint2* inS = reinterpret_cast<uint2*>(smem + lay.inE);
  uint32_t* outS = reinterpret_cast<uint32_t*>(smem + lay.outE);
  double* tsE = reinterpret_cast<double*>(smem + lay.tsE);
  double* subS = reinterpret_cast<double*>(smem + lay.sub);
  volatile uint32_t* ctl = reinterpret_cast<volatile uint32_t*>(smem + lay.ctl);
  double* redV = reinterpret_cast<double*>(smem + lay.red);
  uint32_t* redO = reinterpret_cast<uint32_t*>(smem + lay.red + W * 32 * 8);

  // ---- the CTA's slice of the tables, resident for the whole launch ----
  {
    for (uint32_t i = tid; i < M; i += nThreads) {
      hdrS[i] = tb.hdr[rank * M + i];
      maskCur[i] = 0;
      maskNext[i] = 0;
    }
    const uint32_t inBase = tb.rankInOff[rank], nIn = tb.rankInOff[rank + 1] - inBase;
    for (uint32_t i = tid; i < nIn; i += nThreads) inS[i] = tb.inEdges[inBase + i];
    const uint32_t outBase = tb.rankOutOff[rank], nOut = tb.rankOutOff[rank + 1] - outBase;
    for (uint32_t i = tid; i < nOut; i += nThreads) outS[i] = tb.outEdges[outBase + i];
    for (uint32_t i = tid; i < 512; i += nThreads) tsE[i] = tb.tsE[i];
    if (tid < 16) subS[tid] = tb.sub[tid];
    if (tid < 64) ctl[tid] = 0;
  }
  __syncthreads();

  double* const sPubT = args.sPub + (size_t)team * 2 * Np * 32;
  double* const s0T = args.s0Next + (size_t)team * Np * 32;
  double* const tParkT = args.tPark + (size_t)team * k * Np * 32;
  double2* const sdPubT = args.sdPub + (size_t)team * 2 * Np * 32;  // two parities of the column
  uint32_t* const inboxT = args.inbox + (size_t)team * T * (W * 32);
  unsigned long long* const bar = args.barrier + (size_t)team * 2;
  uint32_t barGen = 0, barHigh0 = 0, barHigh1 = 0;
  bool sentRemote = false;

  // Team barrier (T > 1): CTA barrier, one arrival per CTA on an L2 counter (two alternating counters; the high
  // half counts the CTAs that flagged a remote bit since the last meeting), CTA barrier.  Returns "anybody sent".
  auto teamBarrier = [&]() -> bool {
    if (T == 1) {
      __syncthreads();
      return false;
    }
    const int mine = __syncthreads_or(sentRemote ? 1 : 0);
    sentRemote = false;
    if (tid == 0) {
      unsigned long long* c = bar + (barGen & 1u);
      __threadfence();
      atomicAdd(c, 1ull + (mine ? (1ull << 32) : 0ull));
      const uint32_t target = (barGen / 2 + 1) * T;
      unsigned long long v;
      while ((uint32_t)((v = ldAcquire64(c)) & 0xFFFFFFFFull) < target) {
      }
      const uint32_t high = (uint32_t)(v >> 32);
      const uint32_t prev = (barGen & 1u) ? barHigh1 : barHigh0;
      ctl[kCtlAny] = high != prev ? 1u : 0u;
      ctl[kCtlSent] = high;
    }
    __syncthreads();
    const uint32_t high = ctl[kCtlSent];
    if (barGen & 1u)
      barHigh1 = high;
    else
      barHigh0 = high;
    ++barGen;
    const bool any = ctl[kCtlAny] != 0;
    __syncthreads();  // ctl is rewritten by the next meeting
    return any;
  };

  unsigned long long dbgLevels = 0, dbgVisits = 0, dbgEdges = 0, dbgRounds = 0, dbgClosure = 0, dbgRecord = 0;

  for (int64_t group = team; group < args.nGroups; group += args.nTeams) {
    const int64_t slotIdx = group * 32 + lane;
    int64_t r = -1;
    if (args.order)
      r = args.order[slotIdx];
    else if (args.readBase + slotIdx < args.nReads)
      r = args.readBase + slotIdx;
    const int32_t L = r >= 0 ? args.readLen[r] : -1;
    const uint8_t* seq = r >= 0 ? args.packed + args.byteOff[r] : nullptr;
    int32_t Lmax = L;
    for (int sh = 16; sh > 0; sh >>= 1) Lmax = max(Lmax, __shfl_xor_sync(0xFFFFFFFFu, Lmax, sh));
    uint8_t* const predG = args.pred + (size_t)group * (size_t)(args.maxLen + 1) * N * K2 * 32;
    unsigned long long win = 0;
    uint32_t x = 0, xNext = 0;

    for (int32_t pos = 0; pos <= Lmax; ++pos) {
      const bool act = pos <= L;
      x = xNext;  // observed base pos-1
      if (pos < L) {
        if ((pos & 31) == 0) win = *reinterpret_cast<const unsigned long long*>(seq + (size_t)(pos >> 5) * 8);
        xNext = (uint32_t)(win >> (2 * (pos & 31))) & 3u;
      }
      const uint32_t par = (uint32_t)pos & 1u;
      double2* const sdPubCol = sdPubT + (size_t)par * Np * 32;

      // ---- (1) the column starts from the emission step's result (fused into the previous column's
      //          record pass): S = S0, D = -inf (src/viterbi.cpp:66,75-79,92-106) ----
      for (uint32_t sl = 0; sl < nSlots; ++sl) {
        const uint32_t d = sl * W + warp;
        if (d >= M) break;
        const uint4 h = hdrS[d];
        const uint32_t g = rank * M + d;
        double s0;
        if (pos == 0)
          s0 = (!bhPad(h) && (tb.local || (rank == tb.startRank && d == tb.startLocal))) ? 0.0 : NEG;
        else
          s0 = s0T[(size_t)g * 32 + lane];
        stsRow(aSD + (d * 32 + lane) * 16, s0, NEG);
        if (T > 1 && bhRemoteOut(h)) stPub2(sdPubCol + (size_t)g * 32 + lane, s0, NEG);
      }
      if (tid < 3) ctl[kCtlFlag + tid] = 0;
      teamBarrier();
      unsigned long long stamp = 0;
      if (kDebug) stamp = clock64();

      // ---- (2) closure (src/viterbi.cpp:97-99,110-159), owner-computes edge relaxation ----
      // relaxes the flagged in-transitions of state d; when a lane grew: stores the row, publishes it if some
      // successor lives in another CTA, flags the successors' masks.  Returns "grew" (warp-uniform).
      auto relax = [&](uint32_t d, uint32_t m) -> bool {
        const uint4 h = hdrS[d];
        const uint32_t inOff = bhInOff(h), nE = bhNEmit(h), nIn = nE + bhNNull(h);
        const uint32_t aOwn = aSD + (d * 32 + lane) * 16;
        const double2 own = ldsRow(aOwn);
        double s = own.x, dd = own.y;
        auto edge = [&](uint32_t j) {
          const uint2 e = inS[inOff + j];
          const double2 v = (T > 1 && beRemote(e)) ? ldPub2(sdPubCol + (size_t)e.x * 32 + lane)
                                                   : ldsRow(aSD + ((e.x - rank * M) * 32 + lane) * 16);
          const double sc = tb.symScore[beSym(e)];
          if (j < nE) {
            dd = dmax(dd, dmax(v.y + tb.delExtend, v.x + tb.delOpen) + sc);  // :124-125
          } else {
            dd = dmax(dd, v.y + sc);  // :140
            s = dmax(s, v.x + sc);    // :147 (:98-99)
          }
          if (kDebug) ++dbgEdges;
        };
        if (m == kBatchAllEdges) {
          for (uint32_t j = 0; j < nIn; ++j) edge(j);
        } else {
          while (m) {
            const uint32_t j = (uint32_t)__ffs((int)m) - 1u;
            m &= m - 1u;
            if (j == 31u)
              for (uint32_t jj = 31; jj < nIn; ++jj) edge(jj);
            else
              edge(j);
          }
        }
        s = dmax(s, dd + tb.delEnd);  // :119-121
        const bool grew = act && ((s > own.x) || (dd > own.y));
        if (!__any_sync(0xFFFFFFFFu, grew)) return false;
        stsRow(aOwn, s, dd);
        const uint32_t nOut = bhNOut(h), outOff = bhOutOff(h);
        if (T > 1 && bhRemoteOut(h)) {
          stPub2(sdPubCol + (size_t)(rank * M + d) * 32 + lane, s, dd);
          __threadfence();
          __syncwarp();
        }
        for (uint32_t o = lane; o < nOut; o += 32) {
          const uint32_t w = outS[outOff + o];
          if (T > 1 && boRemote(w)) {
            const uint32_t i2 = boLocal(w);  // the owner's inbox is laid out [warp][slot]
            atomicOr(inboxT + boRank(w) * (W * 32) + (i2 % W) * 32 + i2 / W, 1u << boBit(w));
            sentRemote = true;
          } else
            atomicOr(maskNext + boLocal(w), 1u << boBit(w));
        }
        return true;
      };

      {
        uint32_t lvl = 0;
        bool activated = false;
        // level 0: every transition of every state
        for (uint32_t sl = 0; sl < nSlots; ++sl) {
          const uint32_t d = sl * W + warp;
          if (d >= M) break;
          activated |= relax(d, kBatchAllEdges);
        }
        for (;;) {
          // local levels until this CTA is quiet
          for (;;) {
            if (activated && lane == 0) ctl[kCtlFlag + (lvl + 1) % 3] = 1u;
            if (tid == 0) ctl[kCtlFlag + (lvl + 2) % 3] = 0u;
            __syncthreads();
            {
              uint32_t* t = maskCur;
              maskCur = maskNext;
              maskNext = t;
            }
            ++lvl;
            if (!ctl[kCtlFlag + lvl % 3]) break;
            if (kDebug) ++dbgLevels;
            activated = false;
            uint32_t mym = 0;
            {
              const uint32_t i = lane * W + warp;
              if (lane < nSlots && i < M) {
                mym = maskCur[i];
                if (mym) maskCur[i] = 0;
              }
            }
            uint32_t work = __ballot_sync(0xFFFFFFFFu, mym != 0);
            while (work) {
              const uint32_t sl = (uint32_t)__ffs((int)work) - 1u;
              work &= work - 1u;
              const uint32_t m = __shfl_sync(0xFFFFFFFFu, mym, sl);
              if (kDebug) ++dbgVisits;
              activated |= relax(sl * W + warp, m);
            }
          }
          if (T == 1) break;
          // quiet: drain the inbox (one coalesced load per warp: word warp*32+slot), else meet the team
          bool drained = false;
          {
            const uint32_t i = lane * W + warp;
            if (lane < nSlots && i < M) {
              uint32_t* p = inboxT + rank * (W * 32) + warp * 32 + lane;
              if (ldVolatileGlobal32(p)) {
                const uint32_t m = atomicExch(p, 0u);
                if (m) {
                  atomicOr(maskNext + i, m);
                  drained = true;
                }
              }
            }
          }
          activated = __any_sync(0xFFFFFFFFu, drained);
          if (activated) __threadfence();  // the rows published before those bits were set are read after this point
          if (__syncthreads_or(activated ? 1 : 0)) continue;  // new local work: back to the levels
          if (kDebug) ++dbgRounds;
          if (!teamBarrier()) break;  // nobody flagged a remote bit since the last meeting: fixed point
          // somebody did: its inbox bits were set before it arrived, so the drain above sees them now
          activated = false;
        }
      }
      if (kDebug) {
        const unsigned long long t = clock64();
        dbgClosure += t - stamp;
        stamp = t;
      }

      // ---- (3) predecessor records with the traceback's arithmetic (src/viterbi.cpp:251-286), (4) duplication
      //      opens (:161-168), (5) emission step of column pos+1 (:92-106) ----
      {
        const double* const sPrevCol = sPubT + (size_t)(par ^ 1u) * Np * 32;
        double* const sCurCol = sPubT + (size_t)par * Np * 32;
        double bv = NEG;  // local mode: first maximum of S(.,L) in reference state order
        uint32_t bo = 0xFFFFFFFFu;
        for (uint32_t sl = 0; sl < nSlots; ++sl) {
          const uint32_t d = sl * W + warp;
          if (d >= M) break;
          const uint4 h = hdrS[d];
          if (bhPad(h)) continue;
          const uint32_t inOff = bhInOff(h), nE = bhNEmit(h), nIn = nE + bhNNull(h), mdl = bhMdl(h), orig = h.w;
          const uint32_t g = rank * M + d;
          const double2 own = ldsRow(aSD + (d * 32 + lane) * 16);
          const double sH = own.x, dH = own.y;
          double best = NEG, bestD = NEG, s0n = NEG;
          uint32_t idx = kNoPred, idxD = kNoPred;
          for (uint32_t j = 0; j < nIn; ++j) {
            const uint2 e = inS[inOff + j];
            const uint32_t sym = beSym(e);
            const double2 v = (T > 1 && beRemote(e)) ? ldPub2(sdPubCol + (size_t)e.x * 32 + lane)
                                                     : ldsRow(aSD + ((e.x - rank * M) * 32 + lane) * 16);
            if (j < nE) {
              const uint32_t sb = sym * 4 + beBase(e);
              if (pos > 0) {
                const double c = ldCg(sPrevCol + (size_t)e.x * 32 + lane) + tsE[sb * 4 + x];  // :255
                if (c > best) {
                  best = c;
                  idx = j;
                }
              }
              double c = v.y + tb.tsDext[sym];  // :272
              if (c > bestD) {
                bestD = c;
                idxD = 2 * j;
              }
              c = v.x + tb.tsDopen[sym];  // :273
              if (c > bestD) {
                bestD = c;
                idxD = 2 * j + 1;
              }
              s0n = dmax(s0n, ((v.x + tb.symScore[sym]) + tb.noGap) + subS[beBase(e) * 4 + xNext]);  // :94-95 of pos+1
            } else {
              const double sc = tb.symScore[sym];
              double c = v.x + sc;  // :257
              if (c > best) {
                best = c;
                idx = j;
              }
              c = v.y + sc;  // :276
              if (c > bestD) {
                bestD = c;
                idxD = nE + j;
              }
            }
          }
          {
            const double c = dH + tb.delEnd;  // :258
            if (c > best) {
              best = c;
              idx = nIn;
            }
          }
          if (mdl > 0 && pos > 0) {
            const double parked = tParkT[((size_t)(mdl - 1) * Np + g) * 32 + lane];  // T(state,pos-1,0)+sub, :261
            if (parked > best) {
              best = parked;
              idx = nIn + 1;
            }
          }
          if (tb.local && pos == 0) {  // :263-264
            const double2 v0 = (T > 1 && tb.startRank != rank)
                                   ? ldPub2(sdPubCol + (size_t)(tb.startRank * M + tb.startLocal) * 32 + lane)
                                   : ldsRow(aSD + (tb.startLocal * 32 + lane) * 16);
            if (v0.x + 0.0 > best) {
              best = v0.x + 0.0;
              idx = nIn + 2;
            }
          }
          uint8_t* const pr = predG + ((size_t)pos * N + orig) * K2 * 32 + lane;
          if (act) {
            stRecord(pr, idx);
            stRecord(pr + 32, idxD);
          }
          double tNow[kMaxK];
#pragma unroll
          for (uint32_t t = 0; t < (uint32_t)kMaxK; ++t) {
            tNow[t] = NEG;
            if (t < k) {
              uint32_t idxT = kNoPred;
              if (pos > 0 && t < mdl) {
                double shifted = NEG;
                if (t + 1 < mdl) shifted = tParkT[((size_t)t * Np + g) * 32 + lane];  // T(state,pos-1,t+1)+sub, :285
                if (t + 1 < mdl && shifted > NEG) idxT = 0;
                if (sH + tb.tsT[t] > shifted) idxT = 1;                      // :286
                tNow[t] = dmax(shifted, (sH + tb.tanDup) + tb.len[t]);       // :166-167
              }
              if (act) stRecord(pr + (2 + t) * 32, idxT);
            }
          }
          if (kDebug && args.cells && group == 0 && lane == 0 && act) {
            double* cell = args.cells + ((size_t)pos * N + orig) * K2;
            cell[0] = sH;
            cell[1] = dH;
#pragma unroll
            for (uint32_t t = 0; t < (uint32_t)kMaxK; ++t)
              if (t < k) cell[2 + t] = tNow[t];
          }
          // column pos+1: T -> S candidate and the T shift (:102-106), parked for the next column
          if (mdl > 0) {
            const double t2s = tNow[0] + subS[bhCtx(h, 0) * 4 + xNext];
            s0n = dmax(s0n, t2s);
#pragma unroll
            for (uint32_t t = 0; t + 1 < (uint32_t)kMaxK; ++t)
              if (t + 1 < mdl) tParkT[((size_t)t * Np + g) * 32 + lane] = tNow[t + 1] + subS[bhCtx(h, t + 1) * 4 + xNext];
            tParkT[((size_t)(mdl - 1) * Np + g) * 32 + lane] = t2s;
          }
          s0T[(size_t)g * 32 + lane] = s0n;
          stCg(sCurCol + (size_t)g * 32 + lane, sH);  // the converged column, gathered by the next position's records
          if (!tb.local) {
            if (rank == tb.endRank && d == tb.endLocal && pos == L && r >= 0) args.loglike[r] = sH;  // viterbi.h:102
          } else if (sH > bv || (sH == bv && orig < bo)) {
            bv = sH;
            bo = orig;
          }
        }
        if (tb.local && __any_sync(0xFFFFFFFFu, pos == L)) {  // :171-173, :240-242 (uniform over the CTA: L is per lane)
          redV[warp * 32 + lane] = bv;
          redO[warp * 32 + lane] = bo;
          __syncthreads();
          if (warp == 0) {
            for (uint32_t w = 1; w < (uint32_t)W; ++w) {
              const double v = redV[w * 32 + lane];
              const uint32_t o = redO[w * 32 + lane];
              if (v > bv || (v == bv && o < bo)) {
                bv = v;
                bo = o;
              }
            }
            if (pos == L) {
              args.partVal[((size_t)group * T + rank) * 32 + lane] = bv;
              args.partOrig[((size_t)group * T + rank) * 32 + lane] = bo;
            }
          }
        }
      }
      __syncthreads();  // rows of this column are dead: the next column's S0 may overwrite them
      if (kDebug) dbgRecord += clock64() - stamp;
    }
    // lanes of an empty slot never reach pos == L; nothing to write for them
  }
  if (kDebug && args.dbg && lane == 0) {
    // summed over warps of rank 0 of every team (levels are CTA-uniform: count them once per CTA)
    if (rank == 0) {
      if (warp == 0) atomicAdd(&args.dbg[1], dbgLevels);
      atomicAdd(&args.dbg[2], dbgVisits);
      atomicAdd(&args.dbg[3], dbgEdges);
      if (warp == 0) atomicAdd(&args.dbg[4], dbgClosure);
      if (warp == 0) atomicAdd(&args.dbg[5], dbgRecord);
      if (warp == 0) atomicAdd(&args.dbg[6], dbgRounds);
    }
  }
}

// ---------------------------------------------------------------------------
// traceback over batch-layout records: one thread per read follows the predecessor bytes
// (reference src/viterbi.cpp:195-304: loop :247, emitted symbols :299-300), decoding each byte against the
// state's transition lists in REFERENCE order (the records are indexed by reference state)
// ---------------------------------------------------------------------------
__global__ void viterbiTracebackBatchKernel(const BatchTraceTables tb, const BatchTraceArgs args) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= args.nSlots) return;
  int64_t read = -1;
  if (args.order)
    read = args.order[slot];
  else if (args.readBase + slot < args.nReads)
    read = args.readBase + slot;
  if (read < 0) return;
  const int64_t group = slot >> 5;
  const uint32_t lane = (uint32_t)(slot & 31);
  const uint32_t N = tb.nStates, k = tb.k, K2 = k + 2;
  const int32_t L = args.readLen[read];
  const double NEG = negInf();

  uint32_t state;
  double ll;
  if (!tb.local) {
    state = N - 1;
    ll = args.loglike[read];
  } else {
    double bv = NEG;
    uint32_t bo = 0xFFFFFFFFu;
    for (uint32_t r = 0; r < tb.T; ++r) {
      const double v = args.partVal[((size_t)group * tb.T + r) * 32 + lane];
      const uint32_t o = args.partOrig[((size_t)group * tb.T + r) * 32 + lane];
      if (o == 0xFFFFFFFFu) continue;
      if (v > bv || (v == bv && o < bo)) {
        bv = v;
        bo = o;
      }
    }
    state = bo == 0xFFFFFFFFu ? 0u : bo;
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
  const uint8_t* predG = args.pred + (size_t)group * (size_t)(args.maxLen + 1) * N * K2 * 32 + lane;
  int32_t pos = L;
  uint32_t mut = 0;
  const int32_t cap = args.decodedStride;
  while (pos >= 0 && state != 0) {
    if (path) {
      if (nPath < args.pathStride) {
        path[3 * nPath] = (int32_t)state;
        path[3 * nPath + 1] = pos;
        path[3 * nPath + 2] = (int32_t)mut;
      } else
        status = DNAB_READ_OVERFLOW_;
    }
    ++nPath;
    const uint32_t p = predG[(((size_t)pos * N + state) * K2 + mut) * 32];
    if (p == kNoPred) {
      status = DNAB_READ_TRACEBACK_FAILED_;
      break;
    }
    const uint32_t e0 = tb.emitOff[state], nE = tb.emitOff[state + 1] - e0;
    const uint32_t n0 = tb.nullOff[state], nIn = nE + (tb.nullOff[state + 1] - n0);
    uint32_t sym = 0;
    if (mut == 0) {
      if (p < nE) {
        sym = tb.emitSym[e0 + p];
        state = tb.emitSrc[e0 + p];
        --pos;
      } else if (p < nIn) {
        sym = tb.nullSym[n0 + (p - nE)];
        state = tb.nullSrc[n0 + (p - nE)];
      } else if (p == nIn) {
        mut = 1;
      } else if (p == nIn + 1) {
        mut = 2;
        --pos;
      } else {
        state = 0;  // local mode, pos == 0: jump to (0,0,S)
      }
    } else if (mut == 1) {
      if (p < 2 * nE) {
        sym = tb.emitSym[e0 + (p >> 1)];
        state = tb.emitSrc[e0 + (p >> 1)];
        mut = (p & 1) ? 0 : 1;
      } else {
        sym = tb.nullSym[n0 + (p - 2 * nE)];
        state = tb.nullSrc[n0 + (p - 2 * nE)];
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
typedef void (*BatchKernelPtr)(const BatchTables, const BatchArgs);
static BatchKernelPtr pickBatchKernel(uint32_t warps, bool debug) {
  if (warps > 16) return debug ? viterbiFillBatchKernel<32, true> : viterbiFillBatchKernel<32, false>;
  if (warps > 8) return debug ? viterbiFillBatchKernel<16, true> : viterbiFillBatchKernel<16, false>;
  return debug ? viterbiFillBatchKernel<8, true> : viterbiFillBatchKernel<8, false>;
}

cudaError_t queryBatchTeams(const BatchTables& tb, uint32_t warps, uint32_t smemBytes, int* ctasPerSm) {
  BatchKernelPtr kern = pickBatchKernel(warps, false);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return err;
  const uint32_t w = warps > 16 ? 32 : warps > 8 ? 16 : 8;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctasPerSm, kern, (int)(w * 32), smemBytes);
}

cudaError_t launchFillBatch(const BatchTables& tb, const BatchArgs& args, uint32_t warps, uint32_t smemBytes,
                            cudaStream_t stream) {
  const bool debug = args.dbg != nullptr || args.cells != nullptr;
  BatchKernelPtr kern = pickBatchKernel(warps, debug);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
  if (err != cudaSuccess) return err;
  const uint32_t w = warps > 16 ? 32 : warps > 8 ? 16 : 8;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(args.nTeams * tb.T);
  cfg.blockDim = dim3(w * 32);
  cfg.dynamicSmemBytes = smemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // the CTAs of a team meet at barriers in L2: all must be resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tb.T > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, tb, args);
}

cudaError_t launchTracebackBatch(const BatchTraceTables& tb, const BatchTraceArgs& args, cudaStream_t stream) {
  const int threads = 64;
  const int blocks = (int)((args.nSlots + threads - 1) / threads);
  viterbiTracebackBatchKernel<<<blocks, threads, 0, stream>>>(tb, args);
  return cudaGetLastError();
}

}  // namespace dnab
