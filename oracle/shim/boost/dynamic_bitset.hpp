// Minimal stand-in for boost::dynamic_bitset<> (test-oracle infrastructure only; see multi_array.hpp).
// tracy's gotoh.h uses: ctor (nbits, bool), operator[] read and `bits[i] = true` write.
#ifndef TRACY_B200_SHIM_DYNAMIC_BITSET_HPP
#define TRACY_B200_SHIM_DYNAMIC_BITSET_HPP
#include <cstddef>
#include <cstdint>
#include <vector>
namespace boost {
template <typename Block = unsigned long> class dynamic_bitset {
 public:
  class reference {
   public:
    reference(Block& b, Block m) : b_(b), m_(m) {}
    operator bool() const { return (b_ & m_) != 0; }
    reference& operator=(bool v) { if (v) b_ |= m_; else b_ &= ~m_; return *this; }
   private:
    Block& b_; Block m_;
  };
  dynamic_bitset() : n_(0) {}
  dynamic_bitset(std::size_t n, bool v) : w_((n + B - 1) / B, v ? ~Block(0) : Block(0)), n_(n) {}
  std::size_t size() const { return n_; }
  bool operator[](std::size_t i) const { return (w_[i / B] >> (i % B)) & 1; }
  reference operator[](std::size_t i) { return reference(w_[i / B], Block(1) << (i % B)); }
 private:
  static const std::size_t B = sizeof(Block) * 8;
  std::vector<Block> w_;
  std::size_t n_;
};
}  // namespace boost
#endif
