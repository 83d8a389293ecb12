#include "tables.h"

#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "json_lenient.h"

namespace dnab {

static const char kDna[] = "ACGT";

static int baseOf(char c) {
  const char* p = std::strchr(kDna, std::toupper((unsigned char)c));
  if (!p || !c) throw std::runtime_error(std::string(1, c) + " is not a nucleotide character");
  return (int)(p - kDna);
}

// A<->G and C<->T are transitions: same low bit of the base code (reference src/kmer.h:85-87).
static bool isTransition(int x, int y) { return x != y && (x & 1) == (y & 1); }

MutatorParams MutatorParams::fromFlags(int len, double subProb, double ivRatio, double dupProb, double delOpen,
                                       double delExt, bool global) {
  MutatorParams p;
  const size_t k = (size_t)(len / 2);
  p.pLen.assign(k, 1. / (double)k);
  p.pTanDup = dupProb;
  p.pDelOpen = delOpen;
  p.pDelExtend = delExt;
  p.pTransition = subProb * ivRatio / (1 + ivRatio);
  p.pTransversion = subProb / (1 + ivRatio);
  p.local = !global;
  return p;
}

MutatorParams MutatorParams::fromJSONText(const std::string& text) {
  JsonLenientParser parser(text);
  const JsonNode root = parser.parse();
  MutatorParams p;
  p.pDelOpen = root.at("pDelOpen").num;
  p.pDelExtend = root.at("pDelExtend").num;
  p.pTanDup = root.at("pTanDup").num;
  p.pTransition = root.at("pTransition").num;
  p.pTransversion = root.at("pTransversion").num;
  p.local = root.at("local").b;
  for (const JsonNode& v : root.at("pLen").items) p.pLen.push_back(v.num);
  return p;
}

MutatorParams MutatorParams::fromFile(const std::string& filename) {
  std::ifstream in(filename);
  if (!in) throw std::runtime_error("File not found: " + filename);
  std::stringstream buf;
  buf << in.rdbuf();
  return fromJSONText(buf.str());
}

std::string MutatorParams::asJSON() const {
  std::ostringstream out;
  out << "{\n";
  out << " \"pDelOpen\": " << pDelOpen << ",\n";
  out << " \"pDelExtend\": " << pDelExtend << ",\n";
  out << " \"pTanDup\": " << pTanDup << ",\n";
  out << " \"pTransition\": " << pTransition << ",\n";
  out << " \"pTransversion\": " << pTransversion << ",\n";
  out << " \"pLen\": [ ";
  for (size_t i = 0; i < pLen.size(); ++i) out << (i ? ", " : "") << pLen[i];
  out << " ],\n";
  out << " \"local\": " << (local ? "true" : "false") << "\n";
  out << "}\n";
  return out.str();
}

InputModel::InputModel(const std::string& alphabet, double symWeight, double controlWeight)
    : inputAlphabet(alphabet) {
  // the normaliser is accumulated in alphabet (= sorted character) order
  double norm = 0;
  for (char c : inputAlphabet) norm += (symProb[c] = isControlSym(c) ? controlWeight : symWeight);
  for (auto& sp : symProb) sp.second /= norm;
}

InputModel InputModel::forDecoder(const Machine& machine, const MutatorParams& params) {
  const std::string alph = machine.inputAlphabet(RelaxedInput | ControlInput | SEOFInput);
  return InputModel(alph, 1., std::pow(4., -(double)(4 * params.maxDupLen())));
}

void CompiledTables::bind() {
  t.emit_off = emit_off.data();
  t.emit_src = emit_src.data();
  t.emit_score = emit_score.data();
  t.emit_base = emit_base.data();
  t.emit_in = emit_in.data();
  t.null_off = null_off.data();
  t.null_src = null_src.data();
  t.null_score = null_score.data();
  t.null_in = null_in.data();
  t.ctx = ctx.data();
  t.mdl = mdl.data();
  t.len = len.data();
  t.n_emit = (uint32_t)emit_src.size();
  t.n_null = (uint32_t)null_src.size();
}

void compileTables(const Machine& machine, const MutatorParams& params, CompiledTables& out) {
  machine.verifyContexts();
  for (char c : machine.outputAlphabet())
    if (!std::strchr(kDna, std::toupper((unsigned char)c))) throw std::runtime_error("Not a DNA-outputting machine");

  const InputModel inmod = InputModel::forDecoder(machine, params);
  if (machine.decoderNullGraphIsCyclic(inmod.inputAlphabet))
    throw std::domain_error("Transducer is cyclic, can't toposort");

  const size_t n = machine.nStates();
  const uint32_t k = (uint32_t)std::min(machine.maxLeftContext(), params.maxDupLen());
  out = CompiledTables();
  out.t.n_states = (uint32_t)n;
  out.t.k = k;
  out.t.local = params.local ? 1 : 0;

  // per-symbol scores: one libm log per distinct symbol (same value the reference
  // recomputes per transition, src/viterbi.cpp:41)
  std::map<char, double> symScore;
  for (const auto& sp : inmod.symProb) symScore[sp.first] = std::log(sp.second);
  auto kept = [&](const MachineTransition& t) { return t.in == kNullSym || t.in == kEOF || symScore.count(t.in); };
  auto scoreOf = [&](const MachineTransition& t) { return symScore.count(t.in) ? symScore.at(t.in) : 0.; };

  // counting pass, then a stable fill in (source, transition) order
  out.emit_off.assign(n + 1, 0);
  out.null_off.assign(n + 1, 0);
  for (const auto& ms : machine.state)
    for (const auto& t : ms.trans)
      if (kept(t)) ++(t.out == kNullSym ? out.null_off : out.emit_off)[t.dest + 1];
  for (size_t s = 0; s < n; ++s) {
    out.emit_off[s + 1] += out.emit_off[s];
    out.null_off[s + 1] += out.null_off[s];
  }
  const size_t nEmit = out.emit_off[n], nNull = out.null_off[n];
  out.emit_src.resize(nEmit);
  out.emit_score.resize(nEmit);
  out.emit_base.resize(nEmit);
  out.emit_in.resize(nEmit);
  out.null_src.resize(nNull);
  out.null_score.resize(nNull);
  out.null_in.resize(nNull);
  std::vector<uint32_t> emitFill(out.emit_off.begin(), out.emit_off.end() - 1);
  std::vector<uint32_t> nullFill(out.null_off.begin(), out.null_off.end() - 1);
  for (size_t s = 0; s < n; ++s)
    for (const auto& t : machine.state[s].trans) {
      if (!kept(t)) continue;
      if (t.out == kNullSym) {
        const uint32_t e = nullFill[t.dest]++;
        out.null_src[e] = (uint32_t)s;
        out.null_score[e] = scoreOf(t);
        out.null_in[e] = (uint8_t)t.in;
      } else {
        const uint32_t e = emitFill[t.dest]++;
        out.emit_src[e] = (uint32_t)s;
        out.emit_score[e] = scoreOf(t);
        out.emit_base[e] = (uint8_t)baseOf(t.out);
        out.emit_in[e] = (uint8_t)t.in;
      }
    }

  // duplication contexts: non-wildcard left-context characters, most recent first
  out.ctx.assign(n * (size_t)k + 1, 0);
  out.mdl.assign(n, 0);
  for (size_t s = 0; s < n; ++s) {
    std::vector<int> bases;
    for (char lc : machine.state[s].leftContext)
      if (lc != kWildContext) bases.push_back(baseOf(lc));
    const size_t m = std::min((size_t)k, bases.size());
    out.mdl[s] = (uint8_t)m;
    for (size_t i = 0; i < m; ++i) out.ctx[s * k + i] = (uint8_t)bases[bases.size() - 1 - i];
  }

  // MutatorScores (reference src/mutator.cpp:56-75)
  out.t.delOpen = std::log(params.pDelOpen);
  out.t.tanDup = std::log(params.pTanDup);
  out.t.noGap = std::log(params.pNoGap());
  out.t.delExtend = std::log(params.pDelExtend);
  out.t.delEnd = std::log(params.pDelEnd());
  const double nullScore = std::log(1. / 4.);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      out.t.sub[i * 4 + j] = (i == j ? std::log(params.pMatch())
                                     : (isTransition(i, j) ? std::log(params.pTransition)
                                                           : std::log(params.pTransversion / 2))) -
                             nullScore;
  out.len.assign(std::max<size_t>(params.maxDupLen(), 1), 0.);
  for (size_t l = 0; l < params.maxDupLen(); ++l) out.len[l] = std::log(params.pLen[l]);
  out.bind();
}

}  // namespace dnab
