// Lenient JSON reader for dnastore machine / error-model files.
//
// dnastore's files in the wild are not strict JSON: data/sync16.json and
// data/water*.json have NO commas between array elements, data/flusher.json has a
// trailing comma (the reference's vendored parser accepts both, SURVEY.md 2 row 6,
// reference src/gason.cpp:295-299).  This reader therefore treats ',' as optional
// whitespace between elements.  It keeps the format, not the parser.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dnab {

struct JsonNode {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<JsonNode> items;                             // Array
  std::vector<std::pair<std::string, JsonNode>> members;   // Object, file order

  const JsonNode* find(const std::string& key) const {
    for (const auto& m : members)
      if (m.first == key) return &m.second;
    return nullptr;
  }
  const JsonNode& at(const std::string& key) const {
    const JsonNode* n = find(key);
    if (!n) throw std::runtime_error("JSON: missing key \"" + key + "\"");
    return *n;
  }
};

class JsonLenientParser {
 public:
  explicit JsonLenientParser(const std::string& text) : s(text), p(0) {}
  JsonNode parse() {
    JsonNode v = value();
    skip();
    return v;
  }

 private:
  const std::string& s;
  size_t p;

  void skip() {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r' || s[p] == ',')) ++p;
  }
  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error(std::string("JSON parse error: ") + what + " at offset " + std::to_string(p));
  }
  JsonNode value() {
    skip();
    if (p >= s.size()) fail("unexpected end of input");
    const char c = s[p];
    JsonNode n;
    if (c == '{') {
      ++p;
      n.kind = JsonNode::Object;
      for (;;) {
        skip();
        if (p >= s.size()) fail("unterminated object");
        if (s[p] == '}') {
          ++p;
          break;
        }
        if (s[p] != '"') fail("expected member name");
        std::string key = string();
        skip();
        if (p >= s.size() || s[p] != ':') fail("expected ':'");
        ++p;
        n.members.emplace_back(std::move(key), value());
      }
    } else if (c == '[') {
      ++p;
      n.kind = JsonNode::Array;
      for (;;) {
        skip();
        if (p >= s.size()) fail("unterminated array");
        if (s[p] == ']') {
          ++p;
          break;
        }
        n.items.push_back(value());
      }
    } else if (c == '"') {
      n.kind = JsonNode::String;
      n.str = string();
    } else if (s.compare(p, 4, "true") == 0) {
      n.kind = JsonNode::Bool;
      n.b = true;
      p += 4;
    } else if (s.compare(p, 5, "false") == 0) {
      n.kind = JsonNode::Bool;
      n.b = false;
      p += 5;
    } else if (s.compare(p, 4, "null") == 0) {
      p += 4;
    } else {
      size_t used = 0;
      try {
        n.num = std::stod(s.substr(p, 64), &used);
      } catch (...) {
        fail("bad token");
      }
      n.kind = JsonNode::Number;
      p += used;
    }
    return n;
  }
  std::string string() {
    std::string out;
    ++p;  // opening quote
    while (p < s.size() && s[p] != '"') {
      char c = s[p++];
      if (c == '\\' && p < s.size()) {
        const char e = s[p++];
        switch (e) {
          case 'n': c = '\n'; break;
          case 't': c = '\t'; break;
          case 'r': c = '\r'; break;
          case 'b': c = '\b'; break;
          case 'f': c = '\f'; break;
          case 'u': {
            if (p + 4 > s.size()) fail("bad \\u escape");
            c = (char)std::stoi(s.substr(p, 4), nullptr, 16);
            p += 4;
            break;
          }
          default: c = e;
        }
      }
      out.push_back(c);
    }
    if (p >= s.size()) fail("unterminated string");
    ++p;  // closing quote
    return out;
  }
};

}  // namespace dnab
