// Host-side transducer: dnastore's Machine / JSON machine format, kept so that
// --load-machine / --compose-machine / --save-machine behave as in the reference
// (reference src/trans.h:53-126, src/trans.cpp:402-469,505-602,636-670).
// The GPU never sees this structure: compile_tables() (tables.h) flattens it once.
#pragma once
#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

namespace dnab {

// Symbol conventions (reference src/trans.h:14-40).
constexpr char kNullSym = '\0';
constexpr char kSOF = '^';
constexpr char kEOF = '$';
constexpr char kFlush = '.';
constexpr char kWildContext = '*';
inline bool isControlSym(char c) { return c >= 'A' && c <= 'Z'; }
inline bool isRelaxedSym(char c) { return c == '0' || c == '1'; }
inline bool isStrictSym(char c) {
  return c == 'i' || c == 'j' || c == 'x' || c == 'y' || c == 'z' || c == 'p' || c == 'q' || c == 'r' || c == 's';
}

enum InputFlags : int {
  StrictInput = 1,
  RelaxedInput = 2,
  FlushInput = 4,
  ControlInput = 8,
  SEOFInput = 16,
};

struct MachineTransition {
  char in = kNullSym;   // input symbol, 0 = none
  char out = kNullSym;  // output symbol (a DNA base for decodable machines), 0 = none
  uint64_t dest = 0;
  MachineTransition() {}
  MachineTransition(char i, char o, uint64_t d) : in(i), out(o), dest(d) {}
  bool isNull() const { return in == kNullSym && out == kNullSym; }
};

struct MachineState {
  std::string name, leftContext, rightContext;
  std::vector<MachineTransition> trans;
  bool isEnd() const { return trans.empty(); }
  bool exitsWithInput() const;
  bool exitsWithoutInput() const;
  bool emitsOutput() const;
  bool isWait() const { return exitsWithInput() && !exitsWithoutInput(); }
  bool isNonWait() const { return !exitsWithInput() && exitsWithoutInput(); }
  const MachineTransition* transFor(char in) const;
};

struct Machine {
  std::vector<MachineState> state;

  size_t nStates() const { return state.size(); }
  size_t maxLeftContext() const;

  // JSON machine format {"state":[{"n":i,"id":..,"l":..,"r":..,"trans":[{"in":c,"out":c,"to":j}]}]}
  // (reference src/trans.cpp:402-469).  Reading is lenient about commas; writing
  // reproduces the reference's text byte for byte.
  static Machine fromJSONText(const std::string& text);
  static Machine fromFile(const std::string& filename);
  void writeJSON(std::ostream& out) const;
  std::string toJSON() const;

  // Throws std::runtime_error with the reference's message when an emitted
  // character contradicts a context (reference src/trans.cpp:484-496).
  void verifyContexts() const;

  bool isWaitingMachine() const;
  Machine waitingMachine() const;                                  // src/trans.cpp:636-670
  static Machine compose(const Machine& first, const Machine& second);  // src/trans.cpp:505-602

  std::string inputAlphabet(int flags) const;   // sorted, src/trans.cpp:280-292
  std::string outputAlphabet() const;

  // True when the transitions without DNA output that the decoder keeps (null
  // input or input in `alphabet`) contain a cycle; the reference throws
  // std::domain_error in that case (src/trans.cpp:604-634).
  bool decoderNullGraphIsCyclic(const std::string& alphabet) const;
};

}  // namespace dnab
