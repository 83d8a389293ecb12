// TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for the subset of
// boost::program_options that the reference's t/dnastore.cpp (:41-86) and
// src/logger.{h,cpp} (:15,42 / :46-54) use.  Boost headers are absent from this
// image; this header lets the UNMODIFIED reference sources compile into
// oracle/_ref/.  It is not part of the product and is never linked into it.
//
// Supported: options_description(caption).add_options()(name,desc)
//            (name, value<T>()->default_value(x), desc), value<int|double|string|
//            vector<string>>, variables_map::count/at().as<T>(),
//            parse_command_line, store, notify, operator<<.
// Parsing: "--long value", "--long=value", "-s value", "-svalue" (e.g. -v0, -l6).
#ifndef DNAB_ORACLE_PROGRAM_OPTIONS_SHIM
#define DNAB_ORACLE_PROGRAM_OPTIONS_SHIM

#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct value_base {
  bool has_default = false;
  virtual ~value_base() {}
  virtual std::shared_ptr<value_base> fresh() const = 0;
  virtual void parse(const std::string& s) = 0;
  virtual bool multi() const { return false; }
};

template <class T> struct scalar_parse {
  static T go(const std::string& s) {
    std::istringstream in(s);
    T v;
    in >> v;
    if (in.fail()) throw std::runtime_error("bad option value: " + s);
    return v;
  }
};
template <> struct scalar_parse<std::string> {
  static std::string go(const std::string& s) { return s; }
};

template <class T> struct typed_value : value_base {
  T v{};
  typed_value* default_value(const T& d) {
    v = d;
    has_default = true;
    return this;
  }
  std::shared_ptr<value_base> fresh() const override {
    auto p = std::make_shared<typed_value<T>>();
    p->v = v;
    p->has_default = has_default;
    return p;
  }
  void parse(const std::string& s) override { v = scalar_parse<T>::go(s); }
};

template <class E> struct typed_value<std::vector<E>> : value_base {
  std::vector<E> v;
  std::shared_ptr<value_base> fresh() const override {
    return std::make_shared<typed_value<std::vector<E>>>();
  }
  void parse(const std::string& s) override { v.push_back(scalar_parse<E>::go(s)); }
  bool multi() const override { return true; }
};

template <class T> typed_value<T>* value() { return new typed_value<T>(); }

struct variable_value {
  std::shared_ptr<value_base> held;
  template <class T> const T& as() const {
    auto* p = dynamic_cast<typed_value<T>*>(held.get());
    if (!p) throw std::runtime_error("option type mismatch");
    return p->v;
  }
};

struct variables_map : std::map<std::string, variable_value> {
  size_t count(const std::string& k) const {
    return std::map<std::string, variable_value>::count(k);
  }
  const variable_value& at(const std::string& k) const {
    auto it = find(k);
    if (it == end()) throw std::out_of_range("no such option: " + k);
    return it->second;
  }
};

struct option_spec {
  std::string longName;
  char shortName = 0;
  std::shared_ptr<value_base> proto;  // null => flag
  std::string desc;
};

class options_description;

class options_adder {
  options_description* owner;

 public:
  explicit options_adder(options_description* o) : owner(o) {}
  options_adder& operator()(const char* name, const char* desc);
  template <class T>
  options_adder& operator()(const char* name, typed_value<T>* v, const char* desc);
  void add(const char* name, std::shared_ptr<value_base> v, const char* desc);
};

class options_description {
 public:
  std::string caption;
  std::vector<option_spec> opts;
  explicit options_description(const std::string& c = "") : caption(c) {}
  options_adder add_options() { return options_adder(this); }
  const option_spec* findLong(const std::string& n) const {
    for (auto& o : opts)
      if (o.longName == n) return &o;
    return nullptr;
  }
  const option_spec* findShort(char c) const {
    for (auto& o : opts)
      if (o.shortName == c) return &o;
    return nullptr;
  }
};

inline void options_adder::add(const char* name, std::shared_ptr<value_base> v, const char* desc) {
  option_spec s;
  std::string n(name);
  auto comma = n.find(',');
  if (comma != std::string::npos) {
    s.longName = n.substr(0, comma);
    s.shortName = n[comma + 1];
  } else
    s.longName = n;
  s.proto = v;
  s.desc = desc;
  owner->opts.push_back(s);
}
inline options_adder& options_adder::operator()(const char* name, const char* desc) {
  add(name, nullptr, desc);
  return *this;
}
template <class T>
options_adder& options_adder::operator()(const char* name, typed_value<T>* v, const char* desc) {
  add(name, std::shared_ptr<value_base>(v), desc);
  return *this;
}

inline std::ostream& operator<<(std::ostream& out, const options_description& d) {
  out << d.caption << ":\n";
  for (auto& o : d.opts) {
    out << "  ";
    if (o.shortName) out << "-" << o.shortName << " [ --" << o.longName << " ]";
    else out << "--" << o.longName;
    if (o.proto) out << " arg";
    out << "  " << o.desc << "\n";
  }
  return out;
}

struct parsed_options {
  const options_description* desc;
  std::vector<std::pair<const option_spec*, std::string>> items;
};

inline parsed_options parse_command_line(int argc, char** argv, const options_description& desc) {
  parsed_options po;
  po.desc = &desc;
  for (int i = 1; i < argc; ++i) {
    std::string a(argv[i]);
    const option_spec* spec = nullptr;
    std::string val;
    bool haveVal = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      std::string n = a.substr(2);
      auto eq = n.find('=');
      if (eq != std::string::npos) {
        val = n.substr(eq + 1);
        n = n.substr(0, eq);
        haveVal = true;
      }
      spec = desc.findLong(n);
      if (!spec) throw std::runtime_error("unrecognised option '" + a + "'");
    } else if (a.size() >= 2 && a[0] == '-') {
      spec = desc.findShort(a[1]);
      if (!spec) throw std::runtime_error("unrecognised option '" + a + "'");
      if (a.size() > 2) {
        val = a.substr(2);
        haveVal = true;
      }
    } else
      throw std::runtime_error("too many positional options have been specified on the command line");
    if (spec->proto) {
      if (!haveVal) {
        if (i + 1 >= argc)
          throw std::runtime_error("the required argument for option '--" + spec->longName + "' is missing");
        val = argv[++i];
      }
      po.items.push_back({spec, val});
    } else
      po.items.push_back({spec, std::string()});
  }
  return po;
}

inline void store(const parsed_options& po, variables_map& vm) {
  for (auto& it : po.items) {
    const option_spec* s = it.first;
    auto found = vm.find(s->longName);
    if (!s->proto) {
      if (found == vm.end()) {
        variable_value vv;
        vv.held = std::make_shared<typed_value<bool>>();
        vm[s->longName] = vv;
      }
      continue;
    }
    if (found == vm.end() || !s->proto->multi()) {
      variable_value vv;
      vv.held = s->proto->fresh();
      vv.held->parse(it.second);
      vm[s->longName] = vv;
    } else
      found->second.held->parse(it.second);
  }
  for (auto& o : po.desc->opts)
    if (o.proto && o.proto->has_default && !vm.count(o.longName)) {
      variable_value vv;
      vv.held = o.proto->fresh();
      vm[o.longName] = vv;
    }
}

inline void notify(variables_map&) {}

}  // namespace program_options
}  // namespace boost

#endif
