// The "machine compiler": Machine + error model -> flat dnab_tables (include/dnab_tables.h).
//
// Host-side equivalents of what the reference rebuilds for every single read:
//   MutatorParams / MutatorScores   reference src/mutator.h:9-41, src/mutator.cpp:6-75
//   InputModel                      reference src/viterbi.cpp:6-14, :309-310
//   MachineScores                   reference src/viterbi.cpp:23-60
// Every double is produced with the same libm call and the same summation order
// as the reference so that the tables are bit-identical to the ones it would use.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../../include/dnab_tables.h"
#include "machine.h"

namespace dnab {

struct MutatorParams {
  double pDelOpen = .001, pDelExtend = .01, pTanDup = .001, pTransition = 0, pTransversion = 0;
  std::vector<double> pLen;
  bool local = true;

  size_t maxDupLen() const { return pLen.size(); }
  double pMatch() const { return 1. - pTransition - pTransversion; }
  double pNoGap() const { return 1. - pDelOpen - pTanDup; }
  double pDelEnd() const { return 1. - pDelExtend; }

  // The CLI's construction from flags (reference t/dnastore.cpp:119-129).
  static MutatorParams fromFlags(int len, double subProb, double ivRatio, double dupProb, double delOpen,
                                 double delExt, bool global);
  // -F/--error-file (reference src/mutator.cpp:18-30)
  static MutatorParams fromJSONText(const std::string& text);
  static MutatorParams fromFile(const std::string& filename);
  std::string asJSON() const;  // reference src/mutator.cpp:6-16
};

struct InputModel {
  std::string inputAlphabet;
  std::map<char, double> symProb;
  InputModel(const std::string& alphabet, double symWeight, double controlWeight);
  // The decoder's model: alphabet = relaxed | control | SOF/EOF symbols of the machine,
  // control weight 4^(-4*maxDupLen) (reference src/viterbi.cpp:309-310).
  static InputModel forDecoder(const Machine& machine, const MutatorParams& params);
};

// Owns the arrays a dnab_tables points into.
struct CompiledTables {
  dnab_tables t{};
  std::vector<uint32_t> emit_off, emit_src, null_off, null_src;
  std::vector<double> emit_score, null_score, len;
  std::vector<uint8_t> emit_base, emit_in, null_in, ctx, mdl;
  void bind();  // point t at the vectors
};

// Throws std::runtime_error("Not a DNA-outputting machine") like the reference's
// Assert (src/viterbi.cpp:27-28) and std::domain_error on a null cycle
// (src/trans.cpp:631-632).
void compileTables(const Machine& machine, const MutatorParams& params, CompiledTables& out);

}  // namespace dnab
