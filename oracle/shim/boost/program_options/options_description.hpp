// Container-only stand-in for the slice of boost::program_options the reference's subcommand entry points use
// (src/consensus.h:330-385, src/sage.h:60-112, src/assemble.h:60-102): long/short names, typed values bound to a
// variable, default values, one positional list, count(). A plain left-to-right argv walk. TEST INFRASTRUCTURE ONLY:
// it lets the oracle run the reference's own `int consensus(argc, argv)` / `align` / `assemble` files-in -> files-out.
#pragma once
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "tracy_boost_stubs.hpp"
namespace boost { namespace program_options {
struct value_semantic {
  virtual ~value_semantic() {}
  virtual void assign(std::string const& text) = 0;
  virtual bool apply_default() = 0;
};
namespace detail {
template <typename T> inline void from_text(std::string const& s, T& out) { std::stringstream ss(s); ss >> out; if (ss.fail()) throw std::runtime_error("bad option value: " + s); }
inline void from_text(std::string const& s, std::string& out) { out = s; }
inline void from_text(std::string const& s, boost::filesystem::path& out) { out = boost::filesystem::path(s); }
template <typename T> inline void from_text(std::string const& s, std::vector<T>& out) { T v; from_text(s, v); out.push_back(v); }
}
template <typename T> struct typed_value : value_semantic {
  T* dst; bool has_default; T def;
  explicit typed_value(T* d) : dst(d), has_default(false), def() {}
  template <typename U> typed_value* default_value(U const& v) { def = static_cast<T>(v); has_default = true; return this; }
  typed_value* default_value(const char* v) { detail::from_text(std::string(v), def); has_default = true; return this; }
  void assign(std::string const& text) override { if (dst) detail::from_text(text, *dst); }
  bool apply_default() override { if (has_default && dst) *dst = def; return has_default; }
};
template <typename T> inline typed_value<T>* value(T* dst) { return new typed_value<T>(dst); }
struct option_entry { std::string longname; char shortname; std::shared_ptr<value_semantic> sem; };
class options_description {
 public:
  std::vector<std::shared_ptr<option_entry> > opts;
  options_description() {}
  explicit options_description(std::string const&) {}
  struct easy_init {
    options_description* owner;
    easy_init& add(const char* names, value_semantic* s) {
      std::string n(names); auto e = std::make_shared<option_entry>(); e->shortname = 0;
      auto k = n.find(',');
      if (k == std::string::npos) e->longname = n; else { e->longname = n.substr(0, k); if (k + 1 < n.size()) e->shortname = n[k + 1]; }
      e->sem.reset(s); owner->opts.push_back(e); return *this;
    }
    easy_init& operator()(const char* names, const char*) { return add(names, nullptr); }
    easy_init& operator()(const char* names, value_semantic* s, const char*) { return add(names, s); }
    easy_init& operator()(const char* names, value_semantic* s) { return add(names, s); }
  };
  easy_init add_options() { easy_init e; e.owner = this; return e; }
  options_description& add(options_description const& o) { opts.insert(opts.end(), o.opts.begin(), o.opts.end()); return *this; }
};
inline std::ostream& operator<<(std::ostream& os, options_description const& d) {
  for (auto const& e : d.opts) os << "  --" << e->longname << "\n";
  return os;
}
class positional_options_description {
 public:
  std::string name;
  positional_options_description& add(const char* n, int) { name = n; return *this; }
};
struct parsed_options { std::vector<std::pair<std::shared_ptr<option_entry>, std::string> > items; std::vector<std::shared_ptr<option_entry> > all; };
class command_line_parser {
  std::vector<std::string> args_; options_description const* desc_ = nullptr; positional_options_description const* pos_ = nullptr;
 public:
  command_line_parser(int argc, char** argv) { for (int i = 1; i < argc; ++i) args_.push_back(argv[i]); }
  command_line_parser& options(options_description const& d) { desc_ = &d; return *this; }
  command_line_parser& positional(positional_options_description const& p) { pos_ = &p; return *this; }
  parsed_options run() const {
    parsed_options out; out.all = desc_->opts;
    auto by_long = [&](std::string const& n) { for (auto const& e : desc_->opts) if (e->longname == n) return e; throw std::runtime_error("unknown option --" + n); };
    auto by_short = [&](char c) { for (auto const& e : desc_->opts) if (e->shortname == c) return e; throw std::runtime_error(std::string("unknown option -") + c); };
    for (size_t i = 0; i < args_.size(); ++i) {
      std::string const& a = args_[i];
      std::shared_ptr<option_entry> e; std::string val; bool have = false;
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        auto k = a.find('=');
        e = by_long(a.substr(2, k == std::string::npos ? k : k - 2));
        if (k != std::string::npos) { val = a.substr(k + 1); have = true; }
      } else if (a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
        e = by_short(a[1]);
        if (a.size() > 2) { val = a.substr(2); have = true; }
      } else {
        if (!pos_) throw std::runtime_error("unexpected positional argument " + a);
        out.items.push_back(std::make_pair(by_long(pos_->name), a));
        continue;
      }
      if (e->sem && !have) { if (i + 1 >= args_.size()) throw std::runtime_error("missing value for " + a); val = args_[++i]; }
      out.items.push_back(std::make_pair(e, val));
    }
    return out;
  }
};
struct variable_value { bool is_default = false; bool defaulted() const { return is_default; } };
class variables_map : public std::map<std::string, variable_value> {
 public:
  size_t count(std::string const& k) const { return std::map<std::string, variable_value>::count(k); }
};
inline void store(parsed_options const& p, variables_map& vm) {
  for (auto const& e : p.all) if (e->sem && e->sem->apply_default()) vm[e->longname].is_default = true;
  for (auto const& it : p.items) { if (it.first->sem) it.first->sem->assign(it.second); vm[it.first->longname].is_default = false; }
}
inline void notify(variables_map&) {}
}}  // namespace boost::program_options
