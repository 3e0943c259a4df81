// Container-only stand-in for boost::tokenizer<boost::char_separator<char>> as src/web.h:44-50 uses it: split at any of the
// separator characters, empty tokens dropped. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <vector>
namespace boost {
template <typename C> struct char_separator { std::basic_string<C> seps; explicit char_separator(const C* s) : seps(s) {} };
template <typename Sep> class tokenizer {
  std::vector<std::string> tok_;
 public:
  typedef std::vector<std::string>::const_iterator iterator;
  tokenizer(std::string const& s, Sep const& sep) {
    std::string cur;
    for (char ch : s) {
      if (sep.seps.find(ch) != std::string::npos) { if (!cur.empty()) tok_.push_back(cur); cur.clear(); }
      else cur.push_back(ch);
    }
    if (!cur.empty()) tok_.push_back(cur);
  }
  iterator begin() const { return tok_.begin(); }
  iterator end() const { return tok_.end(); }
};
}  // namespace boost
