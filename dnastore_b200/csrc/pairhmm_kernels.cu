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

struct PairBatch {
  int64_t n;
  const uint8_t* in;        // concatenated tokens
  const uint8_t* out;
  const int64_t* inOff;     // [n+1]
  const int64_t* outOff;    // [n+1]
  const int32_t* lo;        // concatenated per-row ranges, row r of alignment i at rowBase[i]+r
  const int32_t* hi;
  const int64_t* rowOff;    // cell offset of each row (same indexing as lo/hi), relative to cellBase[i]
  const int64_t* rowBase;   // [n+1]
  const int64_t* cellBase;  // [n+1] first cell of each alignment in F / B
  double* F;
  double* B;
  const double* lseTable;
  double* fwdLL;
  double* backLL;
  double* counts;           // [n][5 + k + 16]
};

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xFFF0000000000000LL); }

__global__ void pairHmmFwdBackKernel(const PairScores sc, const PairBatch pb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pb.n) return;
  const int k = sc.k, W = 2 + k;
  const uint8_t* in = pb.in + pb.inOff[i];
  const uint8_t* out = pb.out + pb.outOff[i];
  const int inLen = (int)(pb.inOff[i + 1] - pb.inOff[i]), outLen = (int)(pb.outOff[i + 1] - pb.outOff[i]);
  const int32_t* lo = pb.lo + pb.rowBase[i];
  const int32_t* hi = pb.hi + pb.rowBase[i];
  const int64_t* rowOff = pb.rowOff + pb.rowBase[i];
  double* F = pb.F + pb.cellBase[i] * W;
  double* B = pb.B + pb.cellBase[i] * W;
  const double* T = pb.lseTable;
  const double NEG = ninf();
  const int64_t nCells = pb.cellBase[i + 1] - pb.cellBase[i];
  for (int64_t c = 0; c < nCells * W; ++c) {
    F[c] = NEG;
    B[c] = NEG;
  }
  auto inr = [&](int ip, int op) { return ip >= 0 && ip <= inLen && op >= lo[ip] && op <= hi[ip]; };
  auto at = [&](double* M, int ip, int op) { return M + (rowOff[ip] + (op - lo[ip])) * W; };
  auto mdl = [&](int ip) { return k < ip ? k : ip; };
  auto sub = [&](int ip, int op) { return sc.sub[in[ip - 1] * 4 + out[op - 1]]; };
  auto tsub = [&](int ip, int op, int d) { return sc.sub[in[ip - 1 - d] * 4 + out[op - 1]]; };

  // ---- forward (src/fwdback.cpp:46-76)
  if (inr(0, 0)) at(F, 0, 0)[0] = 0;
  for (int ip = 0; ip <= inLen; ++ip)
    for (int op = lo[ip]; op <= hi[ip]; ++op) {
      double* cell = at(F, ip, op);
      if (ip > 0 && op > 0) {
        if (inr(ip - 1, op - 1)) cell[0] = at(F, ip - 1, op - 1)[0] + sc.noGap + sub(ip, op);
        if (inr(ip, op - 1)) {
          const double* ins = at(F, ip, op - 1);
          for (int d = 0; d < mdl(ip) - 1; ++d) cell[2 + d] = ins[2 + d + 1] + tsub(ip, op, d + 1);
          cell[0] = lse(T, cell[0], ins[2] + tsub(ip, op, 0));
        }
      }
      if (ip > 0 && inr(ip - 1, op)) {
        const double* del = at(F, ip - 1, op);
        cell[1] = lse(T, del[0] + sc.delOpen, del[1] + sc.delExtend);
      }
      cell[0] = lse(T, cell[0], cell[1] + sc.delEnd);
      for (int d = 0; d < mdl(ip); ++d) cell[2 + d] = lse(T, cell[2 + d], cell[0] + sc.tanDup + sc.len[d]);
    }
  const double ll = inr(inLen, outLen) ? at(F, inLen, outLen)[0] : NEG;
  pb.fwdLL[i] = ll;

  // ---- backward (src/fwdback.cpp:84-114)
  if (inr(inLen, outLen)) at(B, inLen, outLen)[0] = 0;
  for (int ip = inLen; ip >= 0; --ip)
    for (int op = hi[ip]; op >= lo[ip]; --op) {
      double* cell = at(B, ip, op);
      if (op < outLen) {
        if (ip < inLen && inr(ip + 1, op + 1)) cell[0] = sc.noGap + sub(ip + 1, op + 1) + at(B, ip + 1, op + 1)[0];
        if (ip > 0 && inr(ip, op + 1)) {
          const double* ins = at(B, ip, op + 1);
          for (int d = 1; d < mdl(ip); ++d) cell[2 + d] = tsub(ip, op + 1, d) + ins[2 + d - 1];
          cell[2] = tsub(ip, op + 1, 0) + ins[0];
        }
      }
      if (ip < inLen && inr(ip + 1, op)) {
        const double* del = at(B, ip + 1, op);
        cell[0] = lse(T, cell[0], sc.delOpen + del[1]);
        cell[1] = sc.delExtend + del[1];
      }
      for (int d = 0; d < mdl(ip); ++d) cell[0] = lse(T, cell[0], cell[2 + d] + sc.tanDup + sc.len[d]);
      cell[1] = lse(T, cell[1], cell[0] + sc.delEnd);
    }
  pb.backLL[i] = inr(0, 0) ? at(B, 0, 0)[0] : NEG;

  // ---- expected counts (src/fwdback.cpp:154-188); cells outside the envelope read as -inf
  double* cnt = pb.counts + i * (5 + k + 16);
  for (int c = 0; c < 5 + k + 16; ++c) cnt[c] = 0;
  double *nLen = cnt + 5, *nSub = cnt + 5 + k;
  auto fget = [&](int ip, int op, int m) { return inr(ip, op) ? at(F, ip, op)[m] : NEG; };
  for (int ip = 0; ip <= inLen; ++ip)
    for (int op = lo[ip]; op <= hi[ip]; ++op) {
      const double* bc = at(B, ip, op);
      if (ip > 0 && op > 0) {
        const double c = exp(fget(ip - 1, op - 1, 0) + sc.noGap + sub(ip, op) + bc[0] - ll);
        cnt[2] += c;
        nSub[in[ip - 1] * 4 + out[op - 1]] += c;
        for (int d = 0; d < mdl(ip) - 1; ++d) {
          const double ci = exp(fget(ip, op - 1, 2 + d + 1) + tsub(ip, op, d + 1) + bc[2 + d] - ll);
          nSub[in[ip - 1 - (d + 1)] * 4 + out[op - 1]] += ci;
        }
        const double c0 = exp(fget(ip, op - 1, 2) + tsub(ip, op, 0) + bc[0] - ll);
        nSub[in[ip - 1] * 4 + out[op - 1]] += c0;
      }
      if (ip > 0) {
        cnt[0] += exp(fget(ip - 1, op, 0) + sc.delOpen + bc[1] - ll);
        cnt[3] += exp(fget(ip - 1, op, 1) + sc.delExtend + bc[1] - ll);
      }
      const double* fc = at(F, ip, op);
      cnt[4] += exp(fc[1] + sc.delEnd + bc[0] - ll);
      for (int d = 0; d < mdl(ip); ++d) {
        const double c = exp(fc[0] + sc.tanDup + sc.len[d] + bc[2 + d] - ll);
        cnt[1] += c;
        nLen[d] += c;
      }
    }
}

// ---------------------------------------------------------------------------------------------
template <class T>
static bool upload(T*& d, const std::vector<T>& h) {
  const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  if (cudaMalloc(&d, bytes) != cudaSuccess) return false;
  return h.empty() || cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

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

  // envelope rows: inRange(ip,op) <=> |a[ip]-b[op]| <= maxDistance, a contiguous run of op per ip
  const int maxDist = strict ? 0 : k;
  std::vector<uint8_t> in, out;
  std::vector<int64_t> inOff{0}, outOff{0}, rowOff, rowBase{0}, cellBase{0};
  std::vector<int32_t> lo, hi;
  for (const auto& al : aligns) {
    in.insert(in.end(), al.in.begin(), al.in.end());
    out.insert(out.end(), al.out.begin(), al.out.end());
    inOff.push_back((int64_t)in.size());
    outOff.push_back((int64_t)out.size());
    int64_t cells = 0;
    const int outLen = (int)al.out.size();
    for (size_t ip = 0; ip < al.a.size(); ++ip) {
      int l = outLen + 1, h = -1;
      for (int op = 0; op <= outLen; ++op)
        if (std::abs(al.a[ip] - al.b[op]) <= maxDist) {
          l = std::min(l, op);
          h = std::max(h, op);
        }
      lo.push_back(l);
      hi.push_back(h);
      rowOff.push_back(cells);
      if (h >= l) cells += h - l + 1;
    }
    rowBase.push_back((int64_t)lo.size());
    cellBase.push_back(cellBase.back() + cells);
  }
  const int W = 2 + k, nc = 5 + k + 16;
  PairBatch pb{};
  pb.n = n;
  uint8_t *dIn = nullptr, *dOut = nullptr;
  int64_t *dInOff = nullptr, *dOutOff = nullptr, *dRowOff = nullptr, *dRowBase = nullptr, *dCellBase = nullptr;
  int32_t *dLo = nullptr, *dHi = nullptr;
  double *dF = nullptr, *dB = nullptr, *dTable = nullptr, *dFwd = nullptr, *dBack = nullptr, *dCounts = nullptr;
  const size_t cellDoubles = (size_t)std::max<int64_t>(cellBase.back(), 1) * W;
  bool ok = upload(dIn, in) && upload(dOut, out) && upload(dInOff, inOff) && upload(dOutOff, outOff) &&
            upload(dRowOff, rowOff) && upload(dRowBase, rowBase) && upload(dCellBase, cellBase) && upload(dLo, lo) &&
            upload(dHi, hi) && upload(dTable, logSumExpLookupTable()) &&
            cudaMalloc(&dF, cellDoubles * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&dB, cellDoubles * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&dFwd, n * sizeof(double)) == cudaSuccess && cudaMalloc(&dBack, n * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&dCounts, (size_t)n * nc * sizeof(double)) == cudaSuccess;
  std::vector<double> hc((size_t)n * nc);
  if (ok) {
    pb.in = dIn; pb.out = dOut; pb.inOff = dInOff; pb.outOff = dOutOff; pb.lo = dLo; pb.hi = dHi;
    pb.rowOff = dRowOff; pb.rowBase = dRowBase; pb.cellBase = dCellBase; pb.F = dF; pb.B = dB;
    pb.lseTable = dTable; pb.fwdLL = dFwd; pb.backLL = dBack; pb.counts = dCounts;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = 64, blocks = (int)((n + threads - 1) / threads);
    cudaEventRecord(e0);
    pairHmmFwdBackKernel<<<blocks, threads>>>(sc, pb);
    cudaEventRecord(e1);
    ok = cudaGetLastError() == cudaSuccess &&
         cudaMemcpy(fwdLL.data(), dFwd, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(backLL.data(), dBack, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(hc.data(), dCounts, hc.size() * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
    float ms = 0;
    if (ok && kernelMs && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) *kernelMs = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (!ok) setLastError(std::string("pair-HMM forward/backward: CUDA error: ") + cudaGetErrorString(cudaGetLastError()));
  for (void* p : {(void*)dIn, (void*)dOut, (void*)dInOff, (void*)dOutOff, (void*)dRowOff, (void*)dRowBase, (void*)dCellBase,
                  (void*)dLo, (void*)dHi, (void*)dF, (void*)dB, (void*)dTable, (void*)dFwd, (void*)dBack, (void*)dCounts})
    if (p) cudaFree(p);
  if (!ok) return false;
  for (int64_t i = 0; i < n; ++i) {
    const double* c = hc.data() + (size_t)i * nc;
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
