// Container-only stand-in for the slice of boost::asio the reference's Ensembl client names (src/web.h:60-135): enough for
// `variantsInRegion` to COMPILE. There is no network here and the oracle never asks for annotation: connect() throws, which the
// reference catches and reports. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <istream>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
namespace system {
struct error_code { int v = 0; bool operator!=(int o) const { return v != o; } bool operator==(int o) const { return v == o; } };
struct system_error : std::runtime_error { explicit system_error(error_code const&) : std::runtime_error("boost::system::system_error (stand-in)") {} };
}  // namespace system
namespace asio {
namespace error { enum { eof = 2 }; }
struct io_context {};
class streambuf : public std::stringbuf {
 public:
  std::size_t size() const { return (std::size_t)const_cast<streambuf*>(this)->in_avail(); }
};
struct transfer_at_least_t { std::size_t n; };
inline transfer_at_least_t transfer_at_least(std::size_t n) { return transfer_at_least_t{n}; }
namespace ip {
struct tcp {
  struct endpoints_t {};
  struct resolver {
    explicit resolver(io_context&) {}
    endpoints_t resolve(std::string const&, std::string const&) { return endpoints_t(); }
  };
  struct socket { explicit socket(io_context&) {} };
};
}  // namespace ip
inline void connect(ip::tcp::socket&, ip::tcp::endpoints_t const&) { throw std::runtime_error("no network in the oracle build (boost::asio stand-in)"); }
inline std::size_t write(ip::tcp::socket&, streambuf&) { return 0; }
inline std::size_t read_until(ip::tcp::socket&, streambuf&, std::string const&) { return 0; }
inline std::size_t read(ip::tcp::socket&, streambuf&, transfer_at_least_t, boost::system::error_code& ec) { ec.v = error::eof; return 0; }
}  // namespace asio
}  // namespace boost
