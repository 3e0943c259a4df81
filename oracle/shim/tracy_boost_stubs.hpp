// Container-only stand-ins for the few Boost facilities src/fasta.h, src/scf.h and src/fmindex.h name outside the DP
// (string helpers, a path type, a timestamp for log lines). No algorithmic content. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <sstream>
#include <string>
namespace boost {
inline void erase_all(std::string& s, const std::string& what) {
  if (what.empty()) return;
  for (std::size_t p = s.find(what); p != std::string::npos; p = s.find(what, p)) s.erase(p, what.size());
}
inline std::string to_lower_copy(const std::string& s) { std::string r(s); for (auto& ch : r) ch = (char)std::tolower((unsigned char)ch); return r; }
inline std::string to_upper_copy(const std::string& s) { std::string r(s); for (auto& ch : r) ch = (char)std::toupper((unsigned char)ch); return r; }
template <typename T, typename S> inline T lexical_cast(const S& x) { std::stringstream ss; ss << x; T v; ss >> v; return v; }
namespace filesystem {
class path {
  std::string p_;
 public:
  path() {}
  path(const std::string& s) : p_(s) {}
  path(const char* s) : p_(s) {}
  const std::string& string() const { return p_; }
  const char* c_str() const { return p_.c_str(); }
  path parent_path() const { auto k = p_.find_last_of('/'); return k == std::string::npos ? path("") : path(p_.substr(0, k)); }
  path filename() const { auto k = p_.find_last_of('/'); return k == std::string::npos ? *this : path(p_.substr(k + 1)); }
  path stem() const { std::string f = filename().string(); auto k = f.find_last_of('.'); return (k == std::string::npos || k == 0) ? path(f) : path(f.substr(0, k)); }
  path extension() const { std::string f = filename().string(); auto k = f.find_last_of('.'); return (k == std::string::npos || k == 0) ? path("") : path(f.substr(k)); }
  friend path operator/(const path& a, const path& b) { return a.p_.empty() ? b : path(a.p_ + "/" + b.p_); }
};
inline bool exists(const path& p) { FILE* f = fopen(p.string().c_str(), "rb"); if (f) fclose(f); return f != nullptr; }
inline bool is_regular_file(const path& p) { return exists(p); }
inline std::size_t file_size(const path& p) { FILE* f = fopen(p.string().c_str(), "rb"); if (!f) return 0; fseek(f, 0, SEEK_END); long n = ftell(f); fclose(f); return n > 0 ? (std::size_t)n : 0; }
inline bool remove(const path& p) { return ::remove(p.string().c_str()) == 0; }
}  // namespace filesystem
namespace gregorian {
struct date {};
inline std::string to_iso_string(const date&) { return "00000000"; }
}  // namespace gregorian
namespace posix_time {
struct ptime { gregorian::date date() const { return gregorian::date(); } };
struct second_clock { static ptime local_time() { return ptime(); } };
inline std::string to_simple_string(const ptime&) { return "0000-00-00 00:00:00"; }
}  // namespace posix_time
}  // namespace boost
