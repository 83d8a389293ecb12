#include "pairhmm.h"

#include <cmath>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>

#include "fasta.h"

namespace dnab {

// ---- Stockholm: "name  gapped-sequence" rows, "#..." annotation, "//" ends an alignment; rows that
// recur in later blocks are concatenated (reference src/stockholm.cpp:36-72)
std::vector<StockholmAlignment> readStockholmDatabase(const std::string& filename) {
  std::ifstream in(filename);
  if (!in) throw std::runtime_error("File " + filename + " not found");
  std::vector<StockholmAlignment> db;
  StockholmAlignment cur;
  std::map<std::string, size_t> row;
  auto flush = [&]() {
    if (!cur.names.empty()) db.push_back(cur);
    cur = StockholmAlignment();
    row.clear();
  };
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string first, second, extra;
    if (!(ls >> first)) continue;
    if (first[0] == '#') continue;
    if (first == "//") {
      flush();
      continue;
    }
    if (!(ls >> second) || (ls >> extra)) continue;  // not a "name sequence" row
    auto it = row.find(first);
    if (it == row.end()) {
      row[first] = cur.names.size();
      cur.names.push_back(first);
      cur.gapped.push_back(second);
    } else
      cur.gapped[it->second] += second;
  }
  flush();
  return db;
}

PairAlignment makePairAlignment(const StockholmAlignment& s) {
  if (s.gapped.size() != 2)
    throw std::runtime_error("Training mutator model requires a 2-row alignment; this alignment has " +
                             std::to_string(s.gapped.size()) + " rows");
  auto isGap = [](char c) { return c == '-' || c == '.'; };
  const std::string &r1 = s.gapped[0], &r2 = s.gapped[1];
  const size_t cols = std::max(r1.size(), r2.size());
  PairAlignment p;
  // cumulativeMatches[col], row?PosToCol[pos] collapsed into a[ip], b[op] (alignpath.cpp:237-265)
  std::vector<int32_t> cum(cols + 1, 0);
  std::vector<size_t> pos1{0}, pos2{0};
  int32_t matches = 0;
  for (size_t col = 0; col < cols; ++col) {
    const bool in1 = col < r1.size() && !isGap(r1[col]), in2 = col < r2.size() && !isGap(r2[col]);
    if (in1) {
      const int t = baseToken(r1[col]);
      if (t < 0) throw std::runtime_error(std::string("Unknown symbol ") + r1[col] + " in sequence " + s.names[0] + " (alphabet is ACGT)");
      p.in.push_back((uint8_t)t);
      pos1.push_back(col + 1);
    }
    if (in2) {
      const int t = baseToken(r2[col]);
      if (t < 0) throw std::runtime_error(std::string("Unknown symbol ") + r2[col] + " in sequence " + s.names[1] + " (alphabet is ACGT)");
      p.out.push_back((uint8_t)t);
      pos2.push_back(col + 1);
    }
    if (in1 && in2) ++matches;
    cum[col + 1] = matches;
  }
  for (size_t c : pos1) p.a.push_back(cum[c]);
  for (size_t c : pos2) p.b.push_back(cum[c]);
  return p;
}

// ---- counts (reference src/mutator.cpp:77-234)
static bool isTransitionTok(int x, int y) { return x != y && (x & 1) == (y & 1); }

MutatorCounts& MutatorCounts::initLaplace(double n) {
  nDelOpen = nTanDup = nNoGap = nDelExtend = nDelEnd = n;
  for (double& v : nSub) v = n;
  for (double& v : nLen) v = n;
  return *this;
}
MutatorCounts& MutatorCounts::operator+=(const MutatorCounts& c) {
  if (nLen.size() != c.nLen.size()) throw std::runtime_error("Length mismatch");
  nDelOpen += c.nDelOpen;
  nTanDup += c.nTanDup;
  nNoGap += c.nNoGap;
  nDelExtend += c.nDelExtend;
  nDelEnd += c.nDelEnd;
  for (int i = 0; i < 16; ++i) nSub[i] += c.nSub[i];
  for (size_t l = 0; l < nLen.size(); ++l) nLen[l] += c.nLen[l];
  return *this;
}
double MutatorCounts::nMatch() const {
  double n = 0;
  for (int i = 0; i < 4; ++i) n += nSub[i * 4 + i];
  return n;
}
double MutatorCounts::nTransition() const {
  double n = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (isTransitionTok(i, j)) n += nSub[i * 4 + j];
  return n;
}
double MutatorCounts::nTransversion() const {
  double n = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (i != j && !isTransitionTok(i, j)) n += nSub[i * 4 + j];
  return n;
}
MutatorParams MutatorCounts::mlParams() const {
  MutatorParams p;
  p.pLen.assign(nLen.size(), 1. / (double)nLen.size());  // the length distribution is not re-estimated
  p.pDelOpen = nDelOpen / (nDelOpen + nTanDup + nNoGap);
  p.pTanDup = nTanDup / (nDelOpen + nTanDup + nNoGap);
  p.pDelExtend = nDelExtend / (nDelExtend + nDelEnd);
  const double ni = nTransition(), nv = nTransversion(), nm = nMatch();
  p.pTransition = ni / (ni + nv + nm);
  p.pTransversion = nv / (ni + nv + nm);
  return p;
}
MutatorParams MutatorCounts::mlParams(const MutatorCounts& prior) const {
  MutatorCounts c = *this;
  c += prior;
  return c.mlParams();
}
static double logBetaPdf(double prob, double alpha, double beta) {
  return std::lgamma(alpha + beta) - std::lgamma(alpha) - std::lgamma(beta) + (alpha - 1) * std::log(prob) +
         (beta - 1) * std::log(1 - prob);
}
static double logDirichletPdfCounts(const std::vector<double>& prob, const std::vector<double>& count) {
  std::vector<double> alpha(count);
  for (double& c : alpha) ++c;
  double ld = std::lgamma(std::accumulate(alpha.begin(), alpha.end(), 0.));
  for (size_t n = 0; n < prob.size(); ++n) ld += (alpha[n] - 1) * std::log(prob[n]) - std::lgamma(alpha[n]);
  return ld;
}
double MutatorCounts::logPrior(const MutatorParams& params) const {
  const std::vector<double> pGap = {params.pDelOpen, params.pTanDup, params.pNoGap()};
  const std::vector<double> nGap = {nDelOpen, nTanDup, nNoGap};
  const std::vector<double> pS = {params.pTransition, params.pTransversion, params.pMatch()};
  const std::vector<double> nS = {nTransition(), nTransversion(), nMatch()};
  return logBetaPdf(params.pDelExtend, nDelExtend + 1, nDelEnd + 1) + logDirichletPdfCounts(pGap, nGap) +
         logDirichletPdfCounts(pS, nS);
}
std::string MutatorCounts::asJSON() const {
  std::ostringstream out;
  out << "{\n";
  out << " \"nDelOpen\": " << nDelOpen << ",\n";
  out << " \"nTanDup\": " << nTanDup << ",\n";
  out << " \"nNoGap\": " << nNoGap << ",\n";
  out << " \"nDelExtend\": " << nDelExtend << ",\n";
  out << " \"nDelEnd\": " << nDelEnd << ",\n";
  out << " \"nLen\": [ ";
  for (size_t i = 0; i < nLen.size(); ++i) out << (i ? ", " : "") << nLen[i];
  out << " ],\n";
  out << " \"nSub\": [ ";
  for (int i = 0; i < 4; ++i) {
    out << (i > 0 ? ", " : "") << "[";
    for (int j = 0; j < 4; ++j) out << (j ? "," : "") << nSub[i * 4 + j];
    out << "]";
  }
  out << " ],\n";
  out << " \"nMatch\": " << nMatch() << ",\n";
  out << " \"nTransition\": " << nTransition() << ",\n";
  out << " \"nTransversion\": " << nTransversion() << "\n";
  out << "}\n";
  return out.str();
}

const std::vector<double>& logSumExpLookupTable() {
  // a function-local static is initialised exactly once even when several host threads (one per device) make
  // their first forward / pair-HMM call together
  static const std::vector<double> table = []() {
    const int entries = ((int)(10 / .0001)) + 1;
    std::vector<double> t((size_t)entries);
    for (int n = 0; n < entries; ++n) {
      const double x = n * .0001;
      t[(size_t)n] = std::log(1. + std::exp(-x));
    }
    return t;
  }();
  return table;
}

bool expectedCounts(int device, const MutatorParams& params, const std::vector<PairAlignment>& db, bool strict,
                    MutatorCounts& total, double& loglike) {
  std::vector<double> f, b;
  std::vector<MutatorCounts> per;
  if (!pairHmmFwdBackBatch(device, params, strict, db, f, b, per)) return false;
  total = MutatorCounts(params.maxDupLen());
  loglike = 0;
  for (size_t i = 0; i < db.size(); ++i) {  // summed in database order, like the reference
    total += per[i];
    loglike += f[i];
  }
  return true;
}

bool baumWelchParams(int device, const MutatorParams& init, const MutatorCounts& prior,
                     const std::vector<PairAlignment>& db, bool strict, MutatorParams& fitted, int* iterations) {
  MutatorParams current = init;
  double best = -INFINITY;
  int iter = 0;
  for (; iter < 100; ++iter) {  // BaumWelchMaxIter
    MutatorCounts counts;
    double ll = 0;
    if (!expectedCounts(device, current, db, strict, counts, ll)) return false;
    ll += prior.logPrior(current);
    if ((ll - best) / std::fabs(best) < .001) break;  // BaumWelchMinFracInc
    best = ll;
    current = counts.mlParams(prior);
    current.local = init.local;
  }
  fitted = current;
  if (iterations) *iterations = iter;
  return true;
}

}  // namespace dnab
