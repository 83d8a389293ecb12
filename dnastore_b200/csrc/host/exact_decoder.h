// Error-free ("exact") decoding on the host: what dnastore's -d/--decode-file, --decode-string and
// --decode-bits do (reference src/decoder.h:7-190 `Decoder<Writer>`, :193-240 `BinaryWriter`;
// call sites t/dnastore.cpp:185-211).  This is the O(L) companion of the Viterbi path, kept so that
// BASELINE config 1 ("exact decode (-d) and Viterbi (-V) of data/hello.fa") and the reference's
// testdecode goldens (Makefile:142-144,153,168,176,183) run through this host; it is inherently
// sequential and cheap, so it stays on the CPU and never touches the GPU library state.
//
// The tracker follows every machine state that is consistent with the DNA read so far, each with
// the input symbols that are implied but not yet certain (its "pending" string).  Symbols are
// released as soon as every hypothesis agrees on them, exactly when the reference releases them.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "machine.h"

namespace dnab {

class ExactDecoder {
 public:
  explicit ExactDecoder(const Machine& machine);

  // Consumes one DNA base (case-insensitive) / a string of them; resolved input symbols are appended
  // to symbols().  Throws std::runtime_error("Can't decode 'X'") when no hypothesis can emit the base
  // and "Decoder error: state ... has two possible input queues" when the machine is ambiguous
  // (the reference Asserts, i.e. aborts, in both cases: decoder.h:149,137-141,71-75).
  void decodeSymbol(char base);
  void decodeString(const std::string& bases);

  // End of input (reference Decoder::close, decoder.h:28-47): releases the pending string of the
  // unique end state, or records the reference's "Decoder unresolved" warnings.
  void close();

  const std::string& symbols() const { return released_; }
  std::string takeSymbols() {
    std::string s;
    s.swap(released_);
    return s;
  }
  const std::vector<std::string>& warnings() const { return warnings_; }
  size_t hypotheses() const { return live_.size(); }

 private:
  // one hypothesis: machine state + input symbols implied since the last release; kept sorted by state
  // (the reference iterates a std::map<State, deque>, and "last writer wins" depends on that order)
  using Hypothesis = std::pair<uint64_t, std::string>;
  using HypothesisSet = std::vector<Hypothesis>;
  static Hypothesis* find(HypothesisSet& set, uint64_t state);
  static void put(HypothesisSet& set, uint64_t state, const std::string& pending);
  static bool usable(const MachineTransition& t);

  void followSilentTransitions();  // decoder.h:54-103 `expand`
  void releaseAgreedPrefix();      // decoder.h:160-184 `shiftResolvedSymbols`
  [[noreturn]] void ambiguous(uint64_t state, const std::string& a, const std::string& b) const;

  const Machine& machine_;
  HypothesisSet live_;
  std::string released_;
  std::vector<std::string> warnings_;
};

// BinaryWriter (reference src/decoder.h:193-240): '0'/'1' symbols are packed into bytes, first bit =
// least significant; '^', '$' are skipped silently, control symbols and anything else with a warning.
// Bits of an incomplete last byte are returned in `leftoverBits` in the order the reference prints them
// in its "N bits (...) remaining on output" warning (most significant first).
struct PackedBits {
  std::string bytes;
  std::string leftoverBits;
  std::vector<std::string> warnings;
};
PackedBits packDecodedSymbols(const std::string& symbols);

}  // namespace dnab
