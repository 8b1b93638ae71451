// Minimal stand-in for Boost.Program_options (TEST INFRASTRUCTURE ONLY): just what the reference's command-line functions
// (src/delly.h:199-400, src/tegua.h:209-440) use — option groups with "long,s" names, typed values bound to variables with
// default values, one positional list, count(). Behaviour matched: long options `--name value` / `--name=value`, short `-s value` /
// `-svalue`, everything else positional.
#pragma once
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <boost/filesystem.hpp>
namespace boost { namespace program_options {
struct value_base {
  virtual ~value_base() {}
  virtual void parse(std::string const&) = 0;
  virtual bool has_default() const = 0;
  virtual void apply_default() = 0;
  virtual std::string default_text() const { return ""; }
};
template <typename T> inline void po_convert(std::string const& s, T& out) { std::istringstream is(s); is >> out; if (is.fail()) throw std::runtime_error("invalid option value: " + s); }
inline void po_convert(std::string const& s, std::string& out) { out = s; }
inline void po_convert(std::string const& s, boost::filesystem::path& out) { out = boost::filesystem::path(s); }
template <typename T> struct typed_value : value_base {
  T* dst; bool hasDef = false; T def{};
  explicit typed_value(T* d) : dst(d) {}
  typed_value* default_value(T const& v) { hasDef = true; def = v; return this; }
  template <typename U> typed_value* default_value(U const& v) { hasDef = true; def = (T) v; return this; }
  void parse(std::string const& s) override { po_convert(s, *dst); }
  bool has_default() const override { return hasDef; }
  void apply_default() override { *dst = def; }
};
template <typename T> struct typed_value<std::vector<T> > : value_base {
  std::vector<T>* dst;
  explicit typed_value(std::vector<T>* d) : dst(d) {}
  void parse(std::string const& s) override { T v; po_convert(s, v); dst->push_back(v); }
  bool has_default() const override { return false; }
  void apply_default() override {}
};
template <typename T> inline typed_value<T>* value(T* d) { return new typed_value<T>(d); }
struct option_description {
  std::string longName; char shortName = 0; std::shared_ptr<value_base> val; std::string desc;
};
class options_description;
struct options_adder {
  options_description* owner;
  options_adder& operator()(const char* name, const char* desc);
  options_adder& operator()(const char* name, value_base* v, const char* desc);
};
class options_description {
 public:
  std::string caption;
  std::vector<option_description> opts;
  options_description() {}
  explicit options_description(std::string const& c) : caption(c) {}
  options_adder add_options() { return options_adder{this}; }
  options_description& add(options_description const& o) { for (auto const& x : o.opts) opts.push_back(x); return *this; }
  option_description const* find_long(std::string const& n) const { for (auto const& o : opts) if (o.longName == n) return &o; return nullptr; }
  option_description const* find_short(char c) const { for (auto const& o : opts) if (o.shortName == c) return &o; return nullptr; }
};
inline void po_add(options_description* d, const char* name, value_base* v, const char* desc) {
  option_description o; std::string n(name); std::size_t k = n.find(',');
  if (k == std::string::npos) o.longName = n; else { o.longName = n.substr(0, k); o.shortName = n[k + 1]; }
  o.val.reset(v); o.desc = desc; d->opts.push_back(o);
}
inline options_adder& options_adder::operator()(const char* name, const char* desc) { po_add(owner, name, nullptr, desc); return *this; }
inline options_adder& options_adder::operator()(const char* name, value_base* v, const char* desc) { po_add(owner, name, v, desc); return *this; }
inline std::ostream& operator<<(std::ostream& o, options_description const& d) {
  for (auto const& x : d.opts) { o << "  "; if (x.shortName) o << "-" << x.shortName << " [ --" << x.longName << " ]"; else o << "--" << x.longName; o << "  " << x.desc << "\n"; }
  return o;
}
struct positional_options_description { std::string name; positional_options_description& add(const char* n, int) { name = n; return *this; } };
struct parsed_options { std::vector<std::pair<option_description const*, std::string> > items; options_description const* desc = nullptr; };
class command_line_parser {
 public:
  command_line_parser(int argc, char** argv) { for (int i = 1; i < argc; ++i) args.push_back(argv[i]); }
  command_line_parser& options(options_description const& d) { desc = &d; return *this; }
  command_line_parser& positional(positional_options_description const& p) { pos = &p; return *this; }
  parsed_options run() {
    parsed_options out; out.desc = desc;
    for (std::size_t i = 0; i < args.size(); ++i) {
      std::string const& a = args[i];
      option_description const* o = nullptr; std::string val; bool haveVal = false;
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        std::string n = a.substr(2); std::size_t e = n.find('=');
        if (e != std::string::npos) { val = n.substr(e + 1); n = n.substr(0, e); haveVal = true; }
        o = desc->find_long(n);
        if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
      } else if (a.size() >= 2 && a[0] == '-' && a != "-") {
        o = desc->find_short(a[1]);
        if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
        if (a.size() > 2) { val = a.substr(2); haveVal = true; }
      } else {
        o = pos ? desc->find_long(pos->name) : nullptr;
        if (!o) throw std::runtime_error("too many positional options");
        out.items.push_back(std::make_pair(o, a));
        continue;
      }
      if (o->val) {
        if (!haveVal) { if (i + 1 >= args.size()) throw std::runtime_error("missing value for " + a); val = args[++i]; }
        out.items.push_back(std::make_pair(o, val));
      } else out.items.push_back(std::make_pair(o, std::string()));
    }
    return out;
  }
 private:
  std::vector<std::string> args; options_description const* desc = nullptr; positional_options_description const* pos = nullptr;
};
class variables_map {
 public:
  std::map<std::string, int> seen;
  std::size_t count(std::string const& n) const { auto it = seen.find(n); return it == seen.end() ? 0 : 1; }
};
inline void store(parsed_options const& p, variables_map& vm) {
  for (auto const& it : p.items) { if (it.first->val) it.first->val->parse(it.second); vm.seen[it.first->longName] = 1; }
  for (auto const& o : p.desc->opts) if (o.val && o.val->has_default() && !vm.count(o.longName)) { o.val->apply_default(); vm.seen[o.longName] = 1; }
}
inline void notify(variables_map&) {}
}}  // namespace boost::program_options
