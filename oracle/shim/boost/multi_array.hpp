// Minimal stand-in for boost::multi_array, written for this repository's test oracle only.
// TEST INFRASTRUCTURE: lets the *unmodified* tracy headers under /root/reference/src compile
// in a container without Boost.  Containers only -- no arithmetic lives here.
// Covers what tracy's hot-path headers use: 2-D (and 1-/3-D) arrays, extents[a][b],
// resize(extents[..]) (element-preserving, like Boost), shape(), operator[][] and ::index.
#ifndef TRACY_B200_SHIM_MULTI_ARRAY_HPP
#define TRACY_B200_SHIM_MULTI_ARRAY_HPP
#include <cstddef>
#include <memory>
#include <algorithm>

namespace boost {
namespace shim_detail {
template <std::size_t N> struct extent_gen {
  std::size_t e[N ? N : 1];
  extent_gen<N + 1> operator[](std::size_t v) const {
    extent_gen<N + 1> r;
    for (std::size_t i = 0; i < N; ++i) r.e[i] = e[i];
    r.e[N] = v;
    return r;
  }
};
}  // namespace shim_detail
static const shim_detail::extent_gen<0> extents = shim_detail::extent_gen<0>();

template <typename T, std::size_t N> class multi_array;

// ---- 2-D --------------------------------------------------------------------------------
template <typename T> class multi_array<T, 2> {
 public:
  typedef std::ptrdiff_t index;
  typedef std::size_t size_type;
  typedef T element;
  struct row_ref { T* p; T& operator[](index j) const { return p[j]; } };
  struct const_row_ref { const T* p; const T& operator[](index j) const { return p[j]; } };

  multi_array() : d_(), sh_{0, 0} {}
  explicit multi_array(shim_detail::extent_gen<2> const& x) : d_(), sh_{0, 0} { alloc(x.e[0], x.e[1]); }
  multi_array(multi_array const& o) : d_(), sh_{0, 0} { copy_from(o); }
  multi_array& operator=(multi_array const& o) { if (this != &o) copy_from(o); return *this; }

  void resize(shim_detail::extent_gen<2> const& x) {
    std::size_t r = x.e[0], c = x.e[1];
    std::unique_ptr<T[]> nd(new T[r * c + 1]());
    std::size_t rr = std::min(r, sh_[0]), cc = std::min(c, sh_[1]);
    for (std::size_t i = 0; i < rr; ++i)
      for (std::size_t j = 0; j < cc; ++j) nd[i * c + j] = d_[i * sh_[1] + j];
    d_.swap(nd);
    sh_[0] = r; sh_[1] = c;
  }
  const size_type* shape() const { return sh_; }
  size_type num_elements() const { return sh_[0] * sh_[1]; }
  T* data() { return d_.get(); }
  const T* data() const { return d_.get(); }
  row_ref operator[](index i) { return row_ref{d_.get() + i * (index)sh_[1]}; }
  const_row_ref operator[](index i) const { return const_row_ref{d_.get() + i * (index)sh_[1]}; }

 private:
  void alloc(std::size_t r, std::size_t c) { d_.reset(new T[r * c + 1]()); sh_[0] = r; sh_[1] = c; }
  void copy_from(multi_array const& o) {
    alloc(o.sh_[0], o.sh_[1]);
    std::copy(o.d_.get(), o.d_.get() + o.sh_[0] * o.sh_[1], d_.get());
  }
  std::unique_ptr<T[]> d_;
  size_type sh_[2];
};

// ---- 1-D (used by a few non-hot-path helpers) --------------------------------------------
template <typename T> class multi_array<T, 1> {
 public:
  typedef std::ptrdiff_t index;
  typedef std::size_t size_type;
  multi_array() : d_(), sh_{0} {}
  explicit multi_array(shim_detail::extent_gen<1> const& x) : d_(new T[x.e[0] + 1]()), sh_{x.e[0]} {}
  void resize(shim_detail::extent_gen<1> const& x) {
    std::unique_ptr<T[]> nd(new T[x.e[0] + 1]());
    for (std::size_t i = 0; i < std::min(x.e[0], sh_[0]); ++i) nd[i] = d_[i];
    d_.swap(nd); sh_[0] = x.e[0];
  }
  const size_type* shape() const { return sh_; }
  T& operator[](index i) { return d_[i]; }
  const T& operator[](index i) const { return d_[i]; }
 private:
  std::unique_ptr<T[]> d_;
  size_type sh_[1];
};
}  // namespace boost
#endif
