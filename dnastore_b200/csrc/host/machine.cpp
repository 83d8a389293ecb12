#include "machine.h"

#include <zlib.h>

#include <algorithm>
#include <deque>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

#include "json_lenient.h"

namespace dnab {

bool MachineState::exitsWithInput() const {
  return std::any_of(trans.begin(), trans.end(), [](const MachineTransition& t) { return t.in != kNullSym; });
}
bool MachineState::exitsWithoutInput() const {
  return std::any_of(trans.begin(), trans.end(), [](const MachineTransition& t) { return t.in == kNullSym; });
}
bool MachineState::emitsOutput() const {
  return std::any_of(trans.begin(), trans.end(), [](const MachineTransition& t) { return t.out != kNullSym; });
}
const MachineTransition* MachineState::transFor(char in) const {
  for (const auto& t : trans)
    if (t.in == in) return &t;
  return nullptr;
}

size_t Machine::maxLeftContext() const {
  size_t w = 0;
  for (const auto& ms : state) w = std::max(w, ms.leftContext.size());
  return w;
}

static char singleChar(const JsonNode& n, const char* what) {
  if (n.kind != JsonNode::String || n.str.size() != 1)
    throw std::runtime_error(std::string("Invalid ") + what + " character: " + n.str);
  return n.str[0];
}

Machine Machine::fromJSONText(const std::string& text) {
  JsonLenientParser parser(text);
  const JsonNode root = parser.parse();
  const JsonNode& jstate = root.at("state");
  if (jstate.kind != JsonNode::Array) throw std::runtime_error("machine JSON: \"state\" is not an array");
  Machine m;
  m.state.reserve(jstate.items.size());
  for (const JsonNode& js : jstate.items) {
    MachineState ms;
    if (const JsonNode* n = js.find("n"))
      if ((size_t)n->num != m.state.size())
        throw std::runtime_error("State n=" + std::to_string((size_t)n->num) + " out of sequence");
    if (const JsonNode* id = js.find("id")) ms.name = id->str;
    if (const JsonNode* l = js.find("l")) ms.leftContext = l->str;
    if (const JsonNode* r = js.find("r")) ms.rightContext = r->str;
    const JsonNode& jtrans = js.at("trans");
    for (const JsonNode& jt : jtrans.items) {
      MachineTransition t;
      t.dest = (uint64_t)jt.at("to").num;
      if (const JsonNode* in = jt.find("in")) t.in = singleChar(*in, "input");
      if (const JsonNode* out = jt.find("out")) t.out = singleChar(*out, "output");
      ms.trans.push_back(t);
    }
    m.state.push_back(std::move(ms));
  }
  for (const auto& ms : m.state)
    for (const auto& t : ms.trans)
      if (t.dest >= m.state.size()) throw std::runtime_error("machine JSON: transition target out of range");
  m.verifyContexts();
  return m;
}

// Plain or gzip-compressed JSON (zlib reads both transparently).
Machine Machine::fromFile(const std::string& filename) {
  gzFile fp = gzopen(filename.c_str(), "rb");
  if (!fp) throw std::runtime_error("File not found: " + filename);
  std::string text;
  char buf[1 << 16];
  int got;
  while ((got = gzread(fp, buf, sizeof buf)) > 0) text.append(buf, (size_t)got);
  gzclose(fp);
  if (got < 0) throw std::runtime_error("Error reading " + filename);
  return fromJSONText(text);
}

void Machine::writeJSON(std::ostream& out) const {
  out << "{\"state\": [\n";
  for (size_t s = 0; s < state.size(); ++s) {
    const MachineState& ms = state[s];
    out << " {\"n\":" << s << ",";
    if (!ms.name.empty()) out << "\"id\":\"" << ms.name << "\",";
    if (!ms.leftContext.empty()) out << "\"l\":\"" << ms.leftContext << "\",";
    if (!ms.rightContext.empty()) out << "\"r\":\"" << ms.rightContext << "\",";
    out << "\"trans\":[";
    bool first = true;
    for (const auto& t : ms.trans) {
      if (!first) out << ",";
      first = false;
      out << "{";
      if (t.in) out << "\"in\":\"" << t.in << "\",";
      if (t.out) out << "\"out\":\"" << t.out << "\",";
      out << "\"to\":" << t.dest << "}";
    }
    out << "]}";
    if (s + 1 < state.size()) out << ",";
    out << "\n";
  }
  out << "]}\n";
}

std::string Machine::toJSON() const {
  std::ostringstream o;
  writeJSON(o);
  return o.str();
}

void Machine::verifyContexts() const {
  for (const auto& ms : state)
    for (const auto& t : ms.trans) {
      if (!t.out) continue;
      const MachineState& md = state[t.dest];
      if (!ms.rightContext.empty() && t.out != ms.rightContext[0])
        throw std::runtime_error("In transition from " + ms.name + " to " + md.name + ": emitted character (" + t.out +
                                 ") does not match source's right context (" + ms.rightContext + ")");
      if (!md.leftContext.empty() && t.out != md.leftContext.back())
        throw std::runtime_error("In transition from " + ms.name + " to " + md.name + ": emitted character (" + t.out +
                                 ") does not match destination's left context (" + md.leftContext + ")");
    }
}

bool Machine::isWaitingMachine() const {
  for (const auto& ms : state)
    if (!ms.isWait() && !ms.isNonWait() && !ms.isEnd()) return false;
  return true;
}

// A state that is neither purely waiting nor purely non-waiting (this includes an
// end state, which has no transitions at all) is split into "<name>;n" holding its
// input-free transitions plus a null hop to "<name>;w", which holds the
// input-consuming ones and is numbered directly after it.
Machine Machine::waitingMachine() const {
  const size_t n = state.size();
  std::vector<char> split(n);
  std::vector<uint64_t> newIndex(n);
  uint64_t next = 0;
  for (size_t s = 0; s < n; ++s) {
    split[s] = !state[s].isWait() && !state[s].isNonWait();
    newIndex[s] = next;
    next += split[s] ? 2 : 1;
  }
  Machine wm;
  wm.state.reserve(next);
  for (size_t s = 0; s < n; ++s) {
    const MachineState& ms = state[s];
    if (!split[s]) {
      MachineState copy = ms;
      for (auto& t : copy.trans) t.dest = newIndex[t.dest];
      wm.state.push_back(std::move(copy));
      continue;
    }
    MachineState nw, w;
    nw.name = ms.name + ";n";
    w.name = ms.name + ";w";
    nw.leftContext = w.leftContext = ms.leftContext;
    nw.rightContext = w.rightContext = ms.rightContext;
    for (const auto& t : ms.trans) {
      MachineTransition moved(t.in, t.out, newIndex[t.dest]);
      (t.in == kNullSym ? nw : w).trans.push_back(moved);
    }
    nw.trans.push_back(MachineTransition(kNullSym, kNullSym, newIndex[s] + 1));
    wm.state.push_back(std::move(nw));
    wm.state.push_back(std::move(w));
  }
  return wm;
}

// Product construction first x second (first's output feeds second's input).
// The numbering contract that makes composed machines reproducible (the
// reference's goldens data/mr2l4c4.json, h74l4c4.json, s16*l4c4.json):
//  * product state (i,j) has provisional index i*|second|+j; survivors are
//    renumbered in ascending provisional order;
//  * survivors = reachable from (0,0) and co-reachable from (last,last);
//  * a survivor whose only transition is a pure null hop is merged into the end
//    of its null chain;
//  * transitions of survivors that lead to a non-survivor are KEPT and point at
//    state 0 (the reference's remap table is zero-initialised, trans.cpp:583-594).
Machine Machine::compose(const Machine& first, const Machine& origSecond) {
  const Machine second = origSecond.isWaitingMachine() ? origSecond : origSecond.waitingMachine();
  if (!second.isWaitingMachine())
    throw std::runtime_error("Attempt to compose transducers A*B where B is not a waiting machine");
  if (first.state.empty() || second.state.empty() || !first.state.back().isEnd() || !second.state.back().isEnd())
    throw std::runtime_error("Last state must be end state");

  const uint64_t n1 = first.nStates(), n2 = second.nStates();
  const uint64_t total = n1 * n2;
  auto pid = [n2](uint64_t i, uint64_t j) { return i * n2 + j; };

  // Transitions of product state (i,j), generated on demand.
  auto expand = [&](uint64_t c, std::vector<MachineTransition>& out) {
    out.clear();
    const uint64_t i = c / n2, j = c % n2;
    const MachineState& msi = first.state[i];
    const MachineState& msj = second.state[j];
    if (msj.isWait() || msj.isEnd()) {
      for (const auto& it : msi.trans) {
        if (it.out == kNullSym)
          out.emplace_back(it.in, kNullSym, pid(it.dest, j));
        else
          for (const auto& jt : msj.trans)
            if (it.out == jt.in) out.emplace_back(it.in, jt.out, pid(it.dest, jt.dest));
      }
    } else
      for (const auto& jt : msj.trans) out.emplace_back(kNullSym, jt.out, pid(i, jt.dest));
  };

  // forward sweep from the start, remembering each visited state's transitions
  std::vector<int32_t> slot(total, -1);  // provisional index -> position in `found`
  std::vector<uint64_t> found;
  std::vector<std::vector<MachineTransition>> foundTrans;
  {
    std::vector<MachineTransition> tmp;
    slot[pid(0, 0)] = 0;
    found.push_back(pid(0, 0));
    for (size_t head = 0; head < found.size(); ++head) {
      expand(found[head], tmp);
      for (const auto& t : tmp)
        if (slot[t.dest] < 0) {
          slot[t.dest] = (int32_t)found.size();
          found.push_back(t.dest);
        }
      foundTrans.push_back(tmp);
    }
  }

  // backward sweep from (last,last) inside the visited set
  std::vector<char> alive(found.size(), 0);
  {
    std::vector<std::vector<int32_t>> preds(found.size());
    for (size_t f = 0; f < found.size(); ++f)
      for (const auto& t : foundTrans[f]) preds[slot[t.dest]].push_back((int32_t)f);
    const int32_t endSlot = slot[pid(n1 - 1, n2 - 1)];
    if (endSlot >= 0) {
      std::deque<int32_t> q{endSlot};
      alive[endSlot] = 1;
      while (!q.empty()) {
        const int32_t c = q.front();
        q.pop_front();
        for (int32_t p : preds[c])
          if (!alive[p]) {
            alive[p] = 1;
            q.push_back(p);
          }
      }
    }
  }

  // survivors in ascending provisional order
  std::vector<int32_t> order;
  for (size_t f = 0; f < found.size(); ++f)
    if (alive[f]) order.push_back((int32_t)f);
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return found[a] < found[b]; });

  auto chainEnd = [&](int32_t f) {
    int32_t d = f;
    while (foundTrans[d].size() == 1 && foundTrans[d].front().isNull()) d = slot[foundTrans[d].front().dest];
    return d;
  };
  std::vector<int64_t> finalIndex(found.size(), -1);
  std::vector<int32_t> merged(found.size(), -1);
  uint64_t nKept = 0;
  for (int32_t f : order) {
    const int32_t d = chainEnd(f);
    if (d != f)
      merged[f] = d;
    else
      finalIndex[f] = (int64_t)nKept++;
  }
  for (int32_t f : order)
    if (merged[f] >= 0) finalIndex[f] = finalIndex[merged[f]];

  Machine out;
  out.state.reserve(nKept);
  for (int32_t f : order) {
    if (merged[f] >= 0) continue;
    const uint64_t i = found[f] / n2, j = found[f] % n2;
    MachineState ms;
    ms.name = "(" + first.state[i].name + "," + second.state[j].name + ")";
    ms.leftContext = second.state[j].leftContext;
    ms.rightContext = second.state[j].rightContext;
    ms.trans = foundTrans[f];
    for (auto& t : ms.trans) {
      const int32_t ds = slot[t.dest];
      t.dest = (ds >= 0 && alive[ds]) ? (uint64_t)finalIndex[ds] : 0;
    }
    out.state.push_back(std::move(ms));
  }
  return out;
}

std::string Machine::inputAlphabet(int flags) const {
  std::set<char> alph;
  for (const auto& ms : state)
    for (const auto& t : ms.trans) {
      if (t.in == kNullSym) continue;
      const bool keep = ((t.in == kEOF || t.in == kSOF) && (flags & SEOFInput)) ||
                        (isControlSym(t.in) && (flags & ControlInput)) || (t.in == kFlush && (flags & FlushInput)) ||
                        (isRelaxedSym(t.in) && (flags & RelaxedInput)) || (isStrictSym(t.in) && (flags & StrictInput));
      if (keep) alph.insert(t.in);
    }
  return std::string(alph.begin(), alph.end());
}

std::string Machine::outputAlphabet() const {
  std::set<char> alph;
  for (const auto& ms : state)
    for (const auto& t : ms.trans)
      if (t.out != kNullSym) alph.insert(t.out);
  return std::string(alph.begin(), alph.end());
}

bool Machine::decoderNullGraphIsCyclic(const std::string& alphabet) const {
  const size_t n = state.size();
  std::vector<int> pending(n, 0);
  size_t edges = 0;
  auto kept = [&](const MachineTransition& t) {
    return t.out == kNullSym && (t.in == kNullSym || alphabet.find(t.in) != std::string::npos);
  };
  for (const auto& ms : state)
    for (const auto& t : ms.trans)
      if (kept(t)) {
        ++pending[t.dest];
        ++edges;
      }
  std::vector<uint64_t> ready;
  for (size_t s = 0; s < n; ++s)
    if (!pending[s]) ready.push_back(s);
  while (!ready.empty()) {
    const uint64_t v = ready.back();
    ready.pop_back();
    for (const auto& t : state[v].trans)
      if (kept(t)) {
        --edges;
        if (--pending[t.dest] == 0) ready.push_back(t.dest);
      }
  }
  return edges > 0;
}

}  // namespace dnab
