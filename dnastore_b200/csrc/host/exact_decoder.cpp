// See exact_decoder.h.  Behaviour restated from reference src/decoder.h (cited per function); the data
// structure (a sorted flat vector of hypotheses instead of std::map<State, deque<char>>) and the control
// flow are this project's own.
#include "exact_decoder.h"

#include <algorithm>
#include <cctype>
#include <stdexcept>

namespace dnab {

namespace {
std::string shown(const std::string& pending) { return pending.empty() ? std::string("empty") : pending; }
std::string pluralOf(size_t n, const char* noun) { return std::to_string(n) + " " + noun + (n == 1 ? "" : "s"); }
}  // namespace

ExactDecoder::Hypothesis* ExactDecoder::find(HypothesisSet& set, uint64_t state) {
  auto it = std::lower_bound(set.begin(), set.end(), state, [](const Hypothesis& h, uint64_t s) { return h.first < s; });
  return (it != set.end() && it->first == state) ? &*it : nullptr;
}

void ExactDecoder::put(HypothesisSet& set, uint64_t state, const std::string& pending) {
  auto it = std::lower_bound(set.begin(), set.end(), state, [](const Hypothesis& h, uint64_t s) { return h.first < s; });
  if (it != set.end() && it->first == state)
    it->second = pending;  // std::map operator[] assignment: the later writer wins
  else
    set.insert(it, Hypothesis(state, pending));
}

// decoder.h:123-128: the decoder follows transitions without input, with a bit, with ^ / $ or with a control symbol
bool ExactDecoder::usable(const MachineTransition& t) {
  return t.in == kNullSym || t.in == '0' || t.in == '1' || t.in == kEOF || t.in == kSOF || isControlSym(t.in);
}

void ExactDecoder::ambiguous(uint64_t state, const std::string& a, const std::string& b) const {
  throw std::runtime_error("Assertion Failed: Decoder error: state " + machine_.state[state].name +
                           " has two possible input queues (" + a + ", " + b + ")");
}

ExactDecoder::ExactDecoder(const Machine& machine) : machine_(machine) {  // decoder.h:16-22
  if (machine_.state.empty()) throw std::runtime_error("ExactDecoder: empty machine");
  live_.push_back(Hypothesis(0, std::string()));
  followSilentTransitions();
}

// decoder.h:54-103.  Rounds: every hypothesis that sits in an end state or in a state that can emit a
// base survives; every transition without output is followed once, extending the pending string by its
// input symbol.  A state reached again (in this or an earlier round) must carry the same pending string.
// Rounds repeat until no new state appears.
void ExactDecoder::followSilentTransitions() {
  HypothesisSet visited, round;
  bool grew;
  do {
    grew = false;
    for (const auto& h : live_)
      if (!find(visited, h.first)) put(visited, h.first, h.second);
    round.clear();
    for (const auto& h : live_) {
      const MachineState& ms = machine_.state[h.first];
      if (ms.isEnd() || ms.emitsOutput()) put(round, h.first, h.second);
    }
    for (const auto& h : live_)
      for (const auto& t : machine_.state[h.first].trans) {
        if (!usable(t) || t.out != kNullSym) continue;
        std::string pending = h.second;
        if (t.in != kNullSym) pending.push_back(t.in);
        if (const Hypothesis* old = find(visited, t.dest)) {
          if (old->second != pending) ambiguous(t.dest, old->second, pending);
        } else {
          put(round, t.dest, pending);
          grew = true;
        }
      }
    live_.swap(round);
  } while (grew);
}

// decoder.h:160-184: while every hypothesis has a non-empty pending string starting with the same
// symbol, that symbol is certain: release it.
void ExactDecoder::releaseAgreedPrefix() {
  for (;;) {
    if (live_.empty() || live_.front().second.empty()) return;
    const char first = live_.front().second[0];
    for (const auto& h : live_)
      if (h.second.empty() || h.second[0] != first) return;
    released_.push_back(first);
    for (auto& h : live_) h.second.erase(h.second.begin());
  }
}

void ExactDecoder::decodeSymbol(char base) {  // decoder.h:130-158
  const char out = (char)std::toupper((unsigned char)base);
  HypothesisSet next;
  for (const auto& h : live_)
    for (const auto& t : machine_.state[h.first].trans) {
      if (!usable(t) || t.out != out) continue;
      std::string pending = h.second;
      if (t.in != kNullSym) pending.push_back(t.in);
      if (const Hypothesis* old = find(next, t.dest))
        if (old->second != pending) ambiguous(t.dest, old->second, pending);
      put(next, t.dest, pending);
    }
  if (next.empty()) throw std::runtime_error(std::string("Assertion Failed: Can't decode '") + out + "'");
  live_.swap(next);
  followSilentTransitions();
  if (live_.size() == 1) {
    // a single hypothesis waiting for input has nothing left to disambiguate (decoder.h:152-156)
    if (machine_.state[live_.front().first].exitsWithInput()) {
      released_ += live_.front().second;
      live_.front().second.clear();
    }
  } else
    releaseAgreedPrefix();
}

void ExactDecoder::decodeString(const std::string& bases) {  // decoder.h:186-189
  for (char c : bases) decodeSymbol(c);
}

void ExactDecoder::close() {  // decoder.h:28-47
  if (live_.empty()) return;
  followSilentTransitions();
  std::vector<const Hypothesis*> ends;
  for (const auto& h : live_)
    if (machine_.state[h.first].isEnd()) ends.push_back(&h);
  if (ends.size() == 1)
    released_ += ends.front()->second;
  else if (ends.size() > 1) {
    warnings_.push_back("Decoder unresolved: " + std::to_string(ends.size()) + " possible end states");
    for (const Hypothesis* h : ends)
      warnings_.push_back("State " + machine_.state[h->first].name + ": input queue " + shown(h->second));
  } else if (live_.size() > 1) {
    warnings_.push_back("Decoder unresolved: " + std::to_string(live_.size()) + " possible states");
    for (const auto& h : live_)
      warnings_.push_back("State " + machine_.state[h.first].name + ": input queue " + shown(h.second));
  }
  live_.clear();
}

PackedBits packDecodedSymbols(const std::string& symbols) {  // decoder.h:193-240 (msb0 == false)
  PackedBits out;
  unsigned byte = 0, nBits = 0;
  for (char c : symbols) {
    if (c == '0' || c == '1') {
      if (c == '1') byte |= 1u << nBits;
      if (++nBits == 8) {
        out.bytes.push_back((char)byte);
        byte = 0;
        nBits = 0;
      }
    } else if (isControlSym(c))
      out.warnings.push_back("Ignoring control character #" + std::to_string(c - 'A') + " ('" + std::string(1, c) + "') in decoder");
    else if (c != kSOF && c != kEOF) {
      static const char* hex = "0123456789abcdef";
      const unsigned char u = (unsigned char)c;
      out.warnings.push_back("Ignoring unknown character '" + std::string(1, c) + "' (\\x" + hex[u >> 4] + hex[u & 15] + ") in decoder");
    }
  }
  if (nBits) {
    // the reference prints the unfinished byte most significant bit first (it reverses its LSB-first buffer)
    for (unsigned b = nBits; b-- > 0;) out.leftoverBits.push_back(((byte >> b) & 1u) ? '1' : '0');
    out.warnings.push_back(pluralOf(nBits, "bit") + " (" + out.leftoverBits + ") remaining on output");
  }
  return out;
}

}  // namespace dnab
