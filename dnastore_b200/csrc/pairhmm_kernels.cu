// Pair-HMM forward / backward / expected counts on the GPU.
//
// What it computes is the reference's ForwardMatrix, BackwardMatrix and FwdBackMatrix::counts
// (reference src/fwdback.cpp:43-78, :80-116, :154-188 with the posterior formulas of
// src/fwdback.h:92-113) over the (original DNA x observed DNA) lattice of a 2-row alignment,
// restricted to the guide-alignment envelope (src/alignpath.h:48-53), with the reference's
// table-based log_sum_exp (src/logsumexp.h:34-74): same table, same interpolation, same
// operand order, so forward and backward log-likelihoods are bit-identical to the reference's.
//
// Mapping: alignments are independent -> one thread per alignment, a batch fills the GPU.
// Only envelope cells are stored: row ip keeps the contiguous run op in [lo[ip], hi[ip]]
// (the envelope coordinate b[op] is non-decreasing), 2+k doubles per cell; the forward cell is
// written once and read once by the counts pass (the 16 B/cell of SURVEY.md 8d).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "capi_error.h"
#include "host/pairhmm.h"
#include "lse_table.cuh"

namespace dnab {

constexpr int kMaxDup = 16;

struct PairScores {
  double delOpen, tanDup, noGap, delExtend, delEnd, sub[16], len[kMaxDup];
  int k;
};

// Layout of a batch.  32 consecutive alignments form a warp (lane = alignment); everything a warp touches in one step is
// lane-interleaved, so the lanes read and write whole lines:
//   * the warp walks the UNION of its lanes' envelopes: row ip covers op in [loW[ip], hiW[ip]] (min / max over the lanes), a
//     lane takes part in a cell when the cell is inside its own envelope [lo, hi];
//   * component m of the warp's cell c (counted over the union rows) of lane l sits at ((cellBase + c) * (2+k) + m) * 32 + l;
//   * tokens, per-lane lo / hi: [warp base + position][32].
struct PairBatch {
  int64_t n;
  int64_t firstWarp;        // this launch covers warps [firstWarp, firstWarp + gridDim-many): a chunk that fits F and B
  int64_t cellOrigin;       // cellBase of firstWarp: F and B hold the chunk's cells only
  const uint8_t* in;        // [inBase[w] + p][32] tokens of the original strands
  const uint8_t* out;       // [outBase[w] + p][32]
  const int32_t* inLen;     // [n]
  const int32_t* outLen;    // [n]
  const int32_t* lo;        // [rowBase[w] + ip][32] the lane's own envelope row (lo > hi: no cell)
  const int32_t* hi;
  const int32_t* loW;       // [rowBase[w] + ip] union over the warp
  const int32_t* hiW;
  const int64_t* rowOff;    // [rowBase[w] + ip] first cell of the union row, relative to cellBase[w]
  const int64_t* rowBase;   // [nWarps + 1]
  const int64_t* cellBase;  // [nWarps]
  const int64_t* inBase;    // [nWarps]
  const int64_t* outBase;   // [nWarps]
  double* F;
  double* B;
  const double* lseTable;
  double* fwdLL;
  double* backLL;
  double* counts;           // [n][5 + k + 16]
};

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xFFF0000000000000LL); }

// K = compile-time bound on the duplication depth k (the cell's components live in registers); every cell is computed in
// registers and written ONCE (no -inf fill of the two matrices), and the neighbour in the same row -- F(ip,op-1) on the way
// forward, B(ip,op+1) on the way back -- is the cell the lane has just computed, so only the other row is read from memory.
// Operand order of every log_sum_exp is the reference's (src/fwdback.cpp:46-76, :84-114, :154-188); every lane visits its
// own cells in the reference's order, so the expected counts are accumulated in the reference's order too.
// PHASE 0: blocks [0, gridDim/2) fill F, blocks [gridDim/2, gridDim) fill B (the two recursions are independent);
// PHASE 1: the expected counts from both.  Separate launches keep each phase's register count -- and with it the number of
// resident warps that hide the latency of the log_sum_exp table look-ups (800 KB: L2, not L1) -- small.
template <int K, int PHASE>
__global__ void __launch_bounds__(64) pairHmmFwdBackKernel(const PairScores sc, const PairBatch pb) {
  const int halfGrid = PHASE == 0 ? (int)(gridDim.x >> 1) : 0;
  const bool backward = PHASE == 0 && (int)blockIdx.x >= halfGrid;
  const int64_t i = pb.firstWarp * 32 + (int64_t)(blockIdx.x - (backward ? halfGrid : 0)) * blockDim.x + threadIdx.x;
  const int64_t w = i >> 5;
  if (w * 32 >= pb.n) return;  // whole warp past the end (a partial last warp keeps its idle lanes: the loops are warp-uniform)
  const bool live = i < pb.n;
  const int lane = threadIdx.x & 31;
  const int k = sc.k, W = 2 + k;
  const int inLen = live ? pb.inLen[i] : -1, outLen = live ? pb.outLen[i] : -1;
  const int64_t rowBase = pb.rowBase[w];
  const int rows = (int)(pb.rowBase[w + 1] - rowBase);
  const uint8_t* in = pb.in + pb.inBase[w] * 32 + lane;
  const uint8_t* out = pb.out + pb.outBase[w] * 32 + lane;
  const int32_t* loL = pb.lo + rowBase * 32 + lane;
  const int32_t* hiL = pb.hi + rowBase * 32 + lane;
  const int32_t* loW = pb.loW + rowBase;
  const int32_t* hiW = pb.hiW + rowBase;
  const int64_t* rowOff = pb.rowOff + rowBase;
  double* F = pb.F + (pb.cellBase[w] - pb.cellOrigin) * W * 32 + lane;
  double* B = pb.B + (pb.cellBase[w] - pb.cellOrigin) * W * 32 + lane;
  const double* T = pb.lseTable;
  const double NEG = ninf();
  const int64_t WS = (int64_t)W * 32;  // doubles from one cell to the next
  // Everything that depends on the row only is fetched once per row: the lane's own range, the range of the neighbouring
  // row (rows are walked in order, so it is the previous iteration's), the row's first cell (cell (ip, op) of matrix M is
  // rowPtr(M, ip) + op * WS) and the input tokens the row's transitions look at.
  auto ownLo = [&](int ip) { return ip <= inLen ? loL[ip * 32] : 1; };
  auto ownHi = [&](int ip) { return ip <= inLen ? hiL[ip * 32] : 0; };
  auto rowPtr = [&](double* M, int ip) { return M + (rowOff[ip] - loW[ip]) * WS; };
  auto mdl = [&](int ip) { return k < ip ? k : ip; };
  auto outTok = [&](int p) { return (int)out[p * 32]; };
  auto rowTokens = [&](int ip, int (&tok)[K]) {  // tok[d] = in[ip-1-d]: the token under the transition that copies (d = 0) or re-reads it
#pragma unroll
    for (int d = 0; d < K; ++d) tok[d] = d < ip ? (int)in[(ip - 1 - d) * 32] : 0;
  };

  if (PHASE == 0 && !backward) {
    // ---- forward (src/fwdback.cpp:46-76)
    int pLo = 1, pHi = 0;  // the lane's range in row ip-1
    for (int ip = 0; ip < rows; ++ip) {
      const int m = mdl(ip), myLo = ownLo(ip), myHi = ownHi(ip), uLo = loW[ip], uHi = hiW[ip];
      double* const cur = rowPtr(F, ip);
      const double* const prv = rowPtr(F, ip > 0 ? ip - 1 : 0);
      int tok[K];
      rowTokens(ip, tok);
      double prev[K + 2];  // F(ip, op-1)
      for (int op = uLo; op <= uHi; ++op) {
        if (op < myLo || op > myHi) continue;
        double c[K + 2];
#pragma unroll
        for (int d = 0; d < K + 2; ++d) c[d] = NEG;
        if (ip == 0 && op == 0) c[0] = 0;
        if (ip > 0 && op > 0) {
          const int ot = outTok(op - 1);
          if (op - 1 >= pLo && op - 1 <= pHi) c[0] = prv[(op - 1) * WS] + sc.noGap + sc.sub[tok[0] * 4 + ot];
          if (op > myLo) {
#pragma unroll
            for (int d = 0; d < K - 1; ++d)
              if (d < m - 1) c[2 + d] = prev[2 + d + 1] + sc.sub[tok[d + 1] * 4 + ot];
            c[0] = lse(T, c[0], prev[2] + sc.sub[tok[0] * 4 + ot]);
          }
        }
        if (ip > 0 && op >= pLo && op <= pHi) {
          const double* del = prv + op * WS;
          c[1] = lse(T, del[0] + sc.delOpen, del[32] + sc.delExtend);
        }
        c[0] = lse(T, c[0], c[1] + sc.delEnd);
#pragma unroll
        for (int d = 0; d < K; ++d)
          if (d < m) c[2 + d] = lse(T, c[2 + d], c[0] + sc.tanDup + sc.len[d]);
        double* cell = cur + op * WS;
#pragma unroll
        for (int d = 0; d < K + 2; ++d) {
          if (d < W) cell[d * 32] = c[d];
          prev[d] = c[d];
        }
      }
      pLo = myLo;
      pHi = myHi;
    }
    if (live) {
      const int eLo = ownLo(inLen), eHi = ownHi(inLen);
      pb.fwdLL[i] = (outLen >= eLo && outLen <= eHi) ? rowPtr(F, inLen)[outLen * WS] : NEG;
    }
  }

  if (PHASE == 0 && backward) {
    // ---- backward (src/fwdback.cpp:84-114)
    int nLo = 1, nHi = 0;  // the lane's range in row ip+1
    for (int ip = rows - 1; ip >= 0; --ip) {
      const int m = mdl(ip), myLo = ownLo(ip), myHi = ownHi(ip), uLo = loW[ip], uHi = hiW[ip];
      double* const cur = rowPtr(B, ip);
      const double* const nxt = rowPtr(B, ip + 1 < rows ? ip + 1 : ip);
      int tok[K];
      rowTokens(ip, tok);
      const int tokNext = ip < inLen ? (int)in[ip * 32] : 0;  // in[ip]: the token row ip+1 copies
      double prev[K + 2];  // B(ip, op+1)
      for (int op = uHi; op >= uLo; --op) {
        if (op < myLo || op > myHi) continue;
        double c[K + 2];
#pragma unroll
        for (int d = 0; d < K + 2; ++d) c[d] = NEG;
        if (ip == inLen && op == outLen) c[0] = 0;
        if (op < outLen) {
          const int ot = outTok(op);  // out[op]: the token under every move into column op+1
          if (ip < inLen && op + 1 >= nLo && op + 1 <= nHi) c[0] = sc.noGap + sc.sub[tokNext * 4 + ot] + nxt[(op + 1) * WS];
          if (ip > 0 && op < myHi) {
#pragma unroll
            for (int d = 1; d < K; ++d)
              if (d < m) c[2 + d] = sc.sub[tok[d] * 4 + ot] + prev[2 + d - 1];
            c[2] = sc.sub[tok[0] * 4 + ot] + prev[0];
          }
        }
        if (ip < inLen && op >= nLo && op <= nHi) {
          const double* del = nxt + op * WS;
          c[0] = lse(T, c[0], sc.delOpen + del[32]);
          c[1] = sc.delExtend + del[32];
        }
#pragma unroll
        for (int d = 0; d < K; ++d)
          if (d < m) c[0] = lse(T, c[0], c[2 + d] + sc.tanDup + sc.len[d]);
        c[1] = lse(T, c[1], c[0] + sc.delEnd);
        double* cell = cur + op * WS;
#pragma unroll
        for (int d = 0; d < K + 2; ++d) {
          if (d < W) cell[d * 32] = c[d];
          prev[d] = c[d];
        }
      }
      nLo = myLo;
      nHi = myHi;
    }
    if (live) pb.backLL[i] = (0 >= ownLo(0) && 0 <= ownHi(0)) ? rowPtr(B, 0)[0] : NEG;
  }

  if (PHASE != 1) return;
  // ---- expected counts (src/fwdback.cpp:154-188); cells outside the envelope read as -inf
  const double ll = live ? pb.fwdLL[i] : NEG;
  // the 16 substitution counts live in shared memory (one column per thread): a dynamically indexed register array would
  // cost a 16-way select per update
  __shared__ double nSubS[16][64];
  double acc[5], nLen[K];
#pragma unroll
  for (int c = 0; c < 5; ++c) acc[c] = 0;
#pragma unroll
  for (int c = 0; c < K; ++c) nLen[c] = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) nSubS[c][threadIdx.x] = 0;
  auto addSub = [&](int idx, double v) { nSubS[idx][threadIdx.x] += v; };
  int pLo = 1, pHi = 0;
  for (int ip = 0; ip < rows; ++ip) {
    const int m = mdl(ip), myLo = ownLo(ip), myHi = ownHi(ip), uLo = loW[ip], uHi = hiW[ip];
    const double* const fcur = rowPtr(F, ip);
    const double* const fprv = rowPtr(F, ip > 0 ? ip - 1 : 0);
    const double* const bcur = rowPtr(B, ip);
    int tok[K];
    rowTokens(ip, tok);
    for (int op = uLo; op <= uHi; ++op) {
      if (op < myLo || op > myHi) continue;
      const double* bc = bcur + op * WS;
      const double* fc = fcur + op * WS;
      if (ip > 0 && op > 0) {
        const int ot = outTok(op - 1);
        const bool diag = op - 1 >= pLo && op - 1 <= pHi, left = op > myLo;
        const double c = exp((diag ? fprv[(op - 1) * WS] : NEG) + sc.noGap + sc.sub[tok[0] * 4 + ot] + bc[0] - ll);
        acc[2] += c;
        addSub(tok[0] * 4 + ot, c);
#pragma unroll
        for (int d = 0; d < K - 1; ++d)
          if (d < m - 1) {
            const double ci = exp((left ? fc[(2 + d + 1) * 32 - WS] : NEG) + sc.sub[tok[d + 1] * 4 + ot] + bc[(2 + d) * 32] - ll);
            addSub(tok[d + 1] * 4 + ot, ci);
          }
        const double c0 = exp((left ? fc[2 * 32 - WS] : NEG) + sc.sub[tok[0] * 4 + ot] + bc[0] - ll);
        addSub(tok[0] * 4 + ot, c0);
      }
      if (ip > 0) {
        const bool up = op >= pLo && op <= pHi;
        acc[0] += exp((up ? fprv[op * WS] : NEG) + sc.delOpen + bc[32] - ll);
        acc[3] += exp((up ? fprv[op * WS + 32] : NEG) + sc.delExtend + bc[32] - ll);
      }
      acc[4] += exp(fc[32] + sc.delEnd + bc[0] - ll);
#pragma unroll
      for (int d = 0; d < K; ++d)
        if (d < m) {
          const double c = exp(fc[0] + sc.tanDup + sc.len[d] + bc[(2 + d) * 32] - ll);
          acc[1] += c;
          nLen[d] += c;
        }
    }
    pLo = myLo;
    pHi = myHi;
  }
  if (!live) return;
  double* cnt = pb.counts + i * (5 + k + 16);
#pragma unroll
  for (int c = 0; c < 5; ++c) cnt[c] = acc[c];
#pragma unroll
  for (int c = 0; c < K; ++c)
    if (c < k) cnt[5 + c] = nLen[c];
#pragma unroll
  for (int c = 0; c < 16; ++c) cnt[5 + k + c] = nSubS[c][threadIdx.x];
}

typedef void (*PairKernelPtr)(const PairScores, const PairBatch);
template <int PHASE>
static PairKernelPtr pickPairKernel(int k) {
  if (k <= 2) return pairHmmFwdBackKernel<2, PHASE>;
  if (k <= 4) return pairHmmFwdBackKernel<4, PHASE>;
  if (k <= 6) return pairHmmFwdBackKernel<6, PHASE>;
  if (k <= 8) return pairHmmFwdBackKernel<8, PHASE>;
  return pairHmmFwdBackKernel<kMaxDup, PHASE>;
}

// ---------------------------------------------------------------------------------------------
template <class T>
static bool upload(T*& d, const std::vector<T>& h) {
  const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  if (cudaMalloc(&d, bytes) != cudaSuccess) return false;
  return h.empty() || cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

static std::atomic<int64_t> gPairChunkCells{0};
void setPairHmmChunkCells(int64_t cells) { gPairChunkCells.store(cells < 0 ? 0 : cells); }

bool pairHmmFwdBackBatch(int device, const MutatorParams& params, bool strict, const std::vector<PairAlignment>& aligns,
                         std::vector<double>& fwdLL, std::vector<double>& backLL, std::vector<MutatorCounts>& counts,
                         double* kernelMs) {
  const int64_t n = (int64_t)aligns.size();
  const int k = (int)params.maxDupLen();
  fwdLL.assign(n, 0.);
  backLL.assign(n, 0.);
  counts.assign(n, MutatorCounts(k));
  if (n == 0) return true;
  if (k > kMaxDup) {
    setLastError("pair-HMM: maxDupLen > 16 is not supported");
    return false;
  }
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || device < 0 || device >= nDev) {
    setLastError("pair-HMM forward/backward: no usable CUDA device; this library has no CPU fallback");
    return false;
  }
  cudaSetDevice(device);

  // scores exactly as MutatorScores (src/mutator.cpp:56-75)
  PairScores sc{};
  sc.k = k;
  sc.delOpen = std::log(params.pDelOpen);
  sc.tanDup = std::log(params.pTanDup);
  sc.noGap = std::log(params.pNoGap());
  sc.delExtend = std::log(params.pDelExtend);
  sc.delEnd = std::log(params.pDelEnd());
  const double nullScore = std::log(1. / 4.);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      sc.sub[i * 4 + j] = (i == j ? std::log(params.pMatch())
                                  : ((i != j && (i & 1) == (j & 1)) ? std::log(params.pTransition)
                                                                    : std::log(params.pTransversion / 2))) -
                          nullScore;
  for (int l = 0; l < k; ++l) sc.len[l] = std::log(params.pLen[l]);

  // envelope rows: inRange(ip,op) <=> |a[ip]-b[op]| <= maxDistance, a contiguous run of op per ip.
  // Alignments are independent, so they are dealt to the warps in order of length (slot s = rank in that order): the lanes
  // of a warp then walk envelopes of nearly the same shape and the union rows hold few idle cells.
  const int maxDist = strict ? 0 : k;
  std::vector<int64_t> order((size_t)n);
  for (int64_t i = 0; i < n; ++i) order[(size_t)i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
    return aligns[(size_t)x].a.size() != aligns[(size_t)y].a.size() ? aligns[(size_t)x].a.size() < aligns[(size_t)y].a.size()
                                                                     : aligns[(size_t)x].out.size() < aligns[(size_t)y].out.size();
  });
  const int64_t nWarps = (n + 31) / 32;
  std::vector<uint8_t> in, out;
  std::vector<int32_t> inLen((size_t)n), outLen((size_t)n), lo, hi, loW, hiW;
  std::vector<int64_t> rowOff, rowBase{0}, cellBase, inBase, outBase;
  int64_t totalCells = 0;
  for (int64_t w = 0; w < nWarps; ++w) {
    const int64_t first = w * 32, lanes = std::min<int64_t>(32, n - first);
    size_t rows = 0, maxIn = 0, maxOut = 0;
    for (int64_t l = 0; l < lanes; ++l) {
      const PairAlignment& al = aligns[(size_t)order[(size_t)(first + l)]];
      rows = std::max(rows, al.a.size());
      maxIn = std::max(maxIn, al.in.size());
      maxOut = std::max(maxOut, al.out.size());
      inLen[(size_t)(first + l)] = (int32_t)al.in.size();
      outLen[(size_t)(first + l)] = (int32_t)al.out.size();
    }
    inBase.push_back((int64_t)in.size() / 32);
    outBase.push_back((int64_t)out.size() / 32);
    in.resize(in.size() + maxIn * 32, 0);
    out.resize(out.size() + maxOut * 32, 0);
    const size_t r0 = loW.size();
    lo.resize((r0 + rows) * 32, 1);  // lo > hi: the lane has no cell in this row
    hi.resize((r0 + rows) * 32, 0);
    loW.resize(r0 + rows, INT32_MAX);
    hiW.resize(r0 + rows, -1);
    for (int64_t l = 0; l < lanes; ++l) {
      const PairAlignment& al = aligns[(size_t)order[(size_t)(first + l)]];
      for (size_t p = 0; p < al.in.size(); ++p) in[((size_t)inBase.back() + p) * 32 + (size_t)l] = al.in[p];
      for (size_t p = 0; p < al.out.size(); ++p) out[((size_t)outBase.back() + p) * 32 + (size_t)l] = al.out[p];
      const int oLen = (int)al.out.size();
      int from = 0;  // b[] is non-decreasing: the run of row ip starts no earlier than the run of row ip-1
      for (size_t ip = 0; ip < al.a.size(); ++ip) {
        if (ip > 0 && al.a[ip] < al.a[ip - 1]) from = 0;
        while (from <= oLen && al.b[(size_t)from] < al.a[ip] - maxDist) ++from;
        int to = from;
        while (to <= oLen && al.b[(size_t)to] <= al.a[ip] + maxDist) ++to;
        const int lcur = from, hcur = to - 1;
        if (hcur >= lcur) {
          lo[(r0 + ip) * 32 + (size_t)l] = lcur;
          hi[(r0 + ip) * 32 + (size_t)l] = hcur;
          loW[r0 + ip] = std::min(loW[r0 + ip], lcur);
          hiW[r0 + ip] = std::max(hiW[r0 + ip], hcur);
        }
      }
    }
    cellBase.push_back(totalCells);
    int64_t cells = 0;
    for (size_t ip = 0; ip < rows; ++ip) {
      if (hiW[r0 + ip] < 0) loW[r0 + ip] = 0;  // nobody has a cell here: an empty union row
      rowOff.push_back(cells);
      if (hiW[r0 + ip] >= loW[r0 + ip]) cells += hiW[r0 + ip] - loW[r0 + ip] + 1;
    }
    totalCells += cells;
    rowBase.push_back((int64_t)loW.size());
  }
  const int W = 2 + k, nc = 5 + k + 16;
  PairBatch pb{};
  pb.n = n;
  uint8_t *dIn = nullptr, *dOut = nullptr;
  int64_t *dRowOff = nullptr, *dRowBase = nullptr, *dCellBase = nullptr, *dInBase = nullptr, *dOutBase = nullptr;
  int32_t *dLo = nullptr, *dHi = nullptr, *dLoW = nullptr, *dHiW = nullptr, *dInLen = nullptr, *dOutLen = nullptr;
  double *dF = nullptr, *dB = nullptr, *dTable = nullptr, *dFwd = nullptr, *dBack = nullptr, *dCounts = nullptr;
  // F and B of the whole batch may not fit the device (131,072 alignments of 200 nt at k = 6: 43 GB): the warps are run in
  // chunks of consecutive warps (an even number, a block holds two) whose cells fit a budget of 40 % of the free memory
  // each for F and B
  cellBase.push_back(totalCells);
  size_t freeB = 0, totalB = 0;
  cudaMemGetInfo(&freeB, &totalB);
  int64_t budgetCells = std::max<int64_t>((int64_t)((double)freeB * 0.4 / ((double)W * 32 * sizeof(double))), 1);
  if (gPairChunkCells.load() > 0) budgetCells = std::min(budgetCells, gPairChunkCells.load());
  std::vector<int64_t> chunkFirst{0};
  int64_t chunkCells = 0;
  for (int64_t w = 0; w < nWarps; w += 2) {
    const int64_t pairCells = cellBase[(size_t)std::min(w + 2, nWarps)] - cellBase[(size_t)w];
    if (chunkCells > 0 && chunkCells + pairCells > budgetCells) {
      chunkFirst.push_back(w);
      chunkCells = 0;
    }
    chunkCells += pairCells;
  }
  chunkFirst.push_back(nWarps);
  int64_t maxChunkCells = 1;
  for (size_t c = 0; c + 1 < chunkFirst.size(); ++c)
    maxChunkCells = std::max(maxChunkCells, cellBase[(size_t)chunkFirst[c + 1]] - cellBase[(size_t)chunkFirst[c]]);
  const size_t cellDoubles = (size_t)maxChunkCells * W * 32;
  if (in.empty()) in.resize(32, 0);
  if (out.empty()) out.resize(32, 0);
  bool ok = upload(dIn, in) && upload(dOut, out) && upload(dInLen, inLen) && upload(dOutLen, outLen) &&
            upload(dRowOff, rowOff) && upload(dRowBase, rowBase) && upload(dCellBase, cellBase) && upload(dInBase, inBase) &&
            upload(dOutBase, outBase) && upload(dLo, lo) && upload(dHi, hi) && upload(dLoW, loW) && upload(dHiW, hiW) &&
            upload(dTable, logSumExpLookupTable()) && cudaMalloc(&dF, cellDoubles * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&dB, cellDoubles * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&dFwd, n * sizeof(double)) == cudaSuccess && cudaMalloc(&dBack, n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&dCounts, (size_t)n * nc * sizeof(double)) == cudaSuccess;
  std::vector<double> hc((size_t)n * nc), hf((size_t)n), hb((size_t)n);
  if (ok) {
    pb.in = dIn; pb.out = dOut; pb.inLen = dInLen; pb.outLen = dOutLen; pb.lo = dLo; pb.hi = dHi; pb.loW = dLoW; pb.hiW = dHiW;
    pb.rowOff = dRowOff; pb.rowBase = dRowBase; pb.cellBase = dCellBase; pb.inBase = dInBase; pb.outBase = dOutBase;
    pb.F = dF; pb.B = dB; pb.lseTable = dTable; pb.fwdLL = dFwd; pb.backLL = dBack; pb.counts = dCounts;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = 64;
    cudaEventRecord(e0);
    for (size_t c = 0; c + 1 < chunkFirst.size(); ++c) {
      pb.firstWarp = chunkFirst[c];
      pb.cellOrigin = cellBase[(size_t)chunkFirst[c]];
      const int blocks = (int)((chunkFirst[c + 1] - chunkFirst[c] + 1) / 2);
      if (blocks <= 0) continue;
      pickPairKernel<0>(k)<<<2 * blocks, threads>>>(sc, pb);
      pickPairKernel<1>(k)<<<blocks, threads>>>(sc, pb);
    }
    cudaEventRecord(e1);
    ok = cudaGetLastError() == cudaSuccess &&
         cudaMemcpy(hf.data(), dFwd, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(hb.data(), dBack, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(hc.data(), dCounts, hc.size() * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    float ms = 0;
    if (ok && kernelMs && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *kernelMs = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (!ok) setLastError(std::string("pair-HMM forward/backward: CUDA error: ") + cudaGetErrorString(cudaGetLastError()));
  for (void* p : {(void*)dIn, (void*)dOut, (void*)dInLen, (void*)dOutLen, (void*)dRowOff, (void*)dRowBase, (void*)dCellBase,
                  (void*)dInBase, (void*)dOutBase, (void*)dLo, (void*)dHi, (void*)dLoW, (void*)dHiW, (void*)dF, (void*)dB,
                  (void*)dTable, (void*)dFwd, (void*)dBack, (void*)dCounts})
    if (p) cudaFree(p);
  if (!ok) return false;
  for (int64_t slot = 0; slot < n; ++slot) {
    const size_t i = (size_t)order[(size_t)slot];
    fwdLL[i] = hf[(size_t)slot];
    backLL[i] = hb[(size_t)slot];
    const double* c = hc.data() + (size_t)slot * nc;
    MutatorCounts& m = counts[i];
    m.nDelOpen = c[0];
    m.nTanDup = c[1];
    m.nNoGap = c[2];
    m.nDelExtend = c[3];
    m.nDelEnd = c[4];
    for (int l = 0; l < k; ++l) m.nLen[l] = c[5 + l];
    for (int j = 0; j < 16; ++j) m.nSub[j] = c[5 + k + j];
  }
  return true;
}

}  // namespace dnab
