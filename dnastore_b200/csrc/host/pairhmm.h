// Host side of the pair-HMM forward/backward path: Stockholm
// alignments, the guide-alignment envelope, expected-count bookkeeping and the Baum-Welch
// driver.  Mirrors the reference's interfaces:
//   readStockholmDatabase            reference src/stockholm.cpp:36-72,154-167
//   Alignment / GuideAlignmentEnvelope   src/alignpath.cpp:189-204,237-265, src/alignpath.h:48-53
//   MutatorCounts                    src/mutator.h:43-67, src/mutator.cpp:77-234
//   expectedCounts / baumWelchParams src/fwdback.cpp:190-230
// The lattice fills themselves run on the GPU (pairhmm_kernels.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tables.h"

namespace dnab {

struct StockholmAlignment {
  std::vector<std::string> names, gapped;  // rows in first-seen order
};
std::vector<StockholmAlignment> readStockholmDatabase(const std::string& filename);

// One 2-row alignment prepared for the lattice: row 0 = original DNA ("in"), row 1 = observed ("out").
struct PairAlignment {
  std::vector<uint8_t> in, out;  // tokens 0..3
  std::vector<int32_t> a, b;     // envelope coordinates: inRange(ip,op) <=> |a[ip]-b[op]| <= maxDistance
};
PairAlignment makePairAlignment(const StockholmAlignment& s);

struct MutatorCounts {
  double nDelOpen = 0, nTanDup = 0, nNoGap = 0, nDelExtend = 0, nDelEnd = 0;
  double nSub[16] = {0};
  std::vector<double> nLen;
  explicit MutatorCounts(size_t maxDupLen = 0) : nLen(maxDupLen, 0.) {}
  MutatorCounts& initLaplace(double n = 1);
  MutatorCounts& operator+=(const MutatorCounts& c);
  double nMatch() const;
  double nTransition() const;
  double nTransversion() const;
  MutatorParams mlParams() const;
  MutatorParams mlParams(const MutatorCounts& prior) const;
  double logPrior(const MutatorParams& params) const;
  std::string asJSON() const;
};

// The reference's log(1+exp(-x)) lookup table (src/logsumexp.h:19-53, logsumexp.cpp:5-15):
// 100,001 entries, step 1e-4, built with libm on the host and uploaded to the device.
const std::vector<double>& logSumExpLookupTable();

// Runs forward, backward and counts for a batch of alignments on `device`.
// Returns false (and sets the error string) on a CUDA failure.
bool pairHmmFwdBackBatch(int device, const MutatorParams& params, bool strictAlignments,
                         const std::vector<PairAlignment>& aligns, std::vector<double>& fwdLL,
                         std::vector<double>& backLL, std::vector<MutatorCounts>& counts, double* kernelMs = nullptr);

// Upper bound on the union-envelope cells (per warp of 32 alignments) one launch keeps in F and in B; 0 = as many as fit
// 40 % of the free device memory each.  Results do not depend on it.
void setPairHmmChunkCells(int64_t cells);

// expectedCounts (src/fwdback.cpp:190-209) and baumWelchParams (:211-230) over a database.
bool expectedCounts(int device, const MutatorParams& params, const std::vector<PairAlignment>& db, bool strict,
                    MutatorCounts& total, double& loglike);
bool baumWelchParams(int device, const MutatorParams& init, const MutatorCounts& prior,
                     const std::vector<PairAlignment>& db, bool strict, MutatorParams& fitted, int* iterations = nullptr);

}  // namespace dnab
