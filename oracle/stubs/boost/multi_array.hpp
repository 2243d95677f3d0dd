// Stand-in for <boost/multi_array.hpp> (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// Boost is not in this image.  The reference's POMDP translation units
// (fast_informed_bound_cuda.cu, point_based_value_iteration_cuda.cu,
// search_tree_cuda.cu) use a small part of Boost.MultiArray: row-major
// N-dimensional views over memory they own elsewhere, element access,
// sub-arrays, `indices[...]` slices, 1-D iterators and element-wise
// assignment.  This header is an independent implementation of exactly that
// subset, written from the documented interface, so that those files compile
// UNMODIFIED for the reference checker (oracle/_ref).
//
// The only property the reference's arithmetic depends on is the element
// order of 1-D views (ascending index, stride taken from the row-major
// layout) -- that is what fixes the float summation order of the
// std::accumulate / std::inner_product / std::partial_sum calls made on them.
// Semantics kept from Boost: construction of a view is shallow, assignment to
// a view or array is an element-wise copy, multi_array value-initialises its
// storage.
#pragma once
// (the real header pulls in <algorithm>, <numeric> and <functional>, which the
// reference relies on)
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <functional>
#include <numeric>
#include <iterator>
#include <type_traits>
#include <vector>

namespace boost {

namespace multi_array_types {
typedef std::ptrdiff_t index;
typedef std::size_t size_type;
// Only the default ("whole dimension") and [start, finish) forms are needed.
struct index_range {
  bool whole;
  index start_, finish_;
  index_range() : whole(true), start_(0), finish_(0) {}
  index_range(index s, index f) : whole(false), start_(s), finish_(f) {}
};
}  // namespace multi_array_types

namespace detail {
namespace pp2d_stub {

using multi_array_types::index;
using multi_array_types::index_range;

template <std::size_t N>
struct extent_gen {
  std::size_t e[N ? N : 1];
  extent_gen<N + 1> operator[](std::size_t n) const {
    extent_gen<N + 1> r;
    for (std::size_t i = 0; i < N; ++i) r.e[i] = e[i];
    r.e[N] = n;
    return r;
  }
};

// indices[...]: NR entries seen so far, ND of them ranges.
template <std::size_t NR, std::size_t ND>
struct index_gen {
  bool is_range[NR ? NR : 1];
  index_range range[NR ? NR : 1];
  index at[NR ? NR : 1];
  template <std::size_t ND2>
  index_gen<NR + 1, ND2> grow() const {
    index_gen<NR + 1, ND2> r;
    for (std::size_t i = 0; i < NR; ++i) {
      r.is_range[i] = is_range[i];
      r.range[i] = range[i];
      r.at[i] = at[i];
    }
    return r;
  }
  index_gen<NR + 1, ND + 1> operator[](const index_range& rg) const {
    index_gen<NR + 1, ND + 1> r = grow<ND + 1>();
    r.is_range[NR] = true;
    r.range[NR] = rg;
    r.at[NR] = 0;
    return r;
  }
  index_gen<NR + 1, ND> operator[](index i) const {
    index_gen<NR + 1, ND> r = grow<ND>();
    r.is_range[NR] = false;
    r.at[NR] = i;
    return r;
  }
};

template <typename T>
class strided_iterator {
 public:
  typedef std::random_access_iterator_tag iterator_category;
  typedef typename std::remove_const<T>::type value_type;
  typedef std::ptrdiff_t difference_type;
  typedef T* pointer;
  typedef T& reference;
  strided_iterator() : p_(nullptr), s_(1) {}
  strided_iterator(T* p, std::ptrdiff_t s) : p_(p), s_(s) {}
  reference operator*() const { return *p_; }
  reference operator[](difference_type n) const { return p_[n * s_]; }
  strided_iterator& operator++() { p_ += s_; return *this; }
  strided_iterator operator++(int) { strided_iterator t = *this; p_ += s_; return t; }
  strided_iterator& operator--() { p_ -= s_; return *this; }
  strided_iterator operator--(int) { strided_iterator t = *this; p_ -= s_; return t; }
  strided_iterator& operator+=(difference_type n) { p_ += n * s_; return *this; }
  strided_iterator& operator-=(difference_type n) { p_ -= n * s_; return *this; }
  strided_iterator operator+(difference_type n) const { return strided_iterator(p_ + n * s_, s_); }
  strided_iterator operator-(difference_type n) const { return strided_iterator(p_ - n * s_, s_); }
  difference_type operator-(const strided_iterator& o) const { return (p_ - o.p_) / s_; }
  bool operator==(const strided_iterator& o) const { return p_ == o.p_; }
  bool operator!=(const strided_iterator& o) const { return p_ != o.p_; }
  bool operator<(const strided_iterator& o) const { return (o - *this) > 0; }
  bool operator>(const strided_iterator& o) const { return (*this - o) > 0; }
  bool operator<=(const strided_iterator& o) const { return !(*this > o); }
  bool operator>=(const strided_iterator& o) const { return !(*this < o); }

 private:
  T* p_;
  std::ptrdiff_t s_;
};

template <typename T, std::size_t N>
class view;

// What operator[](index) returns: an element reference for N == 1, an
// (N-1)-dimensional view otherwise.
template <typename T, std::size_t N>
struct sub_of {
  typedef view<T, N - 1> type;
  static type make(T* base, const std::size_t* shape, const std::ptrdiff_t* stride) {
    return type(base, shape + 1, stride + 1);
  }
};
template <typename T>
struct sub_of<T, 1> {
  typedef T& type;
  static type make(T* base, const std::size_t*, const std::ptrdiff_t*) { return *base; }
};

template <typename T, std::size_t N>
class view {
 public:
  typedef typename std::remove_const<T>::type element;
  typedef typename std::remove_const<T>::type value_type;
  typedef std::size_t size_type;
  typedef multi_array_types::index index;
  typedef strided_iterator<T> iterator;
  typedef strided_iterator<const T> const_iterator;
  template <std::size_t K> struct subarray { typedef view<T, K> type; };
  template <std::size_t K> struct const_subarray { typedef view<const T, K> type; };
  template <std::size_t K> struct array_view { typedef view<T, K> type; };
  template <std::size_t K> struct const_array_view { typedef view<const T, K> type; };

  view() : base_(nullptr) {
    for (std::size_t i = 0; i < N; ++i) { shape_[i] = 0; stride_[i] = 0; }
  }
  view(T* base, const std::size_t* shape, const std::ptrdiff_t* stride) : base_(base) {
    for (std::size_t i = 0; i < N; ++i) { shape_[i] = shape[i]; stride_[i] = stride[i]; }
  }
  view(T* base, const extent_gen<N>& ext) : base_(base) { set_row_major(ext); }
  // shallow, like Boost's view / sub-array copy construction
  view(const view& o) : base_(o.base_) {
    for (std::size_t i = 0; i < N; ++i) { shape_[i] = o.shape_[i]; stride_[i] = o.stride_[i]; }
  }
  // view<float> -> view<const float>
  template <typename U, typename = typename std::enable_if<
                            std::is_same<const U, T>::value>::type>
  view(const view<U, N>& o) : base_(o.origin()) {
    for (std::size_t i = 0; i < N; ++i) { shape_[i] = o.shape()[i]; stride_[i] = o.strides()[i]; }
  }

  // element-wise, like Boost's assignment
  const view& operator=(const view& o) const { copy_from(o); return *this; }
  template <typename U>
  const view& operator=(const view<U, N>& o) const { copy_from(o); return *this; }

  typename sub_of<T, N>::type operator[](index i) const {
    return sub_of<T, N>::make(base_ + i * stride_[0], shape_, stride_);
  }
  template <std::size_t ND>
  view<T, ND> operator[](const index_gen<N, ND>& g) const {
    T* b = base_;
    std::size_t shp[ND ? ND : 1];
    std::ptrdiff_t str[ND ? ND : 1];
    std::size_t d = 0;
    for (std::size_t i = 0; i < N; ++i) {
      if (g.is_range[i]) {
        const index lo = g.range[i].whole ? 0 : g.range[i].start_;
        const index hi = g.range[i].whole ? (index)shape_[i] : g.range[i].finish_;
        b += lo * stride_[i];
        shp[d] = (std::size_t)(hi - lo);
        str[d] = stride_[i];
        ++d;
      } else {
        b += g.at[i] * stride_[i];
      }
    }
    return view<T, ND>(b, shp, str);
  }

  // iteration is provided for 1-D views only (all the reference needs)
  iterator begin() const {
    static_assert(N == 1, "stand-in: iterators exist for 1-D views only");
    return iterator(base_, stride_[0]);
  }
  iterator end() const {
    static_assert(N == 1, "stand-in: iterators exist for 1-D views only");
    return iterator(base_ + (std::ptrdiff_t)shape_[0] * stride_[0], stride_[0]);
  }

  T* origin() const { return base_; }
  T* data() const { return base_; }
  size_type size() const { return shape_[0]; }
  const std::size_t* shape() const { return shape_; }
  const std::ptrdiff_t* strides() const { return stride_; }
  size_type num_elements() const {
    size_type n = 1;
    for (std::size_t i = 0; i < N; ++i) n *= shape_[i];
    return n;
  }
  static size_type num_dimensions() { return N; }

 protected:
  void set_row_major(const extent_gen<N>& ext) {
    std::ptrdiff_t s = 1;
    for (std::size_t i = N; i-- > 0;) {
      shape_[i] = ext.e[i];
      stride_[i] = s;
      s *= (std::ptrdiff_t)ext.e[i];
    }
  }
  template <typename U>
  void copy_from(const view<U, N>& o) const {
    for (std::size_t i = 0; i < N; ++i) assert(shape_[i] == o.shape()[i]);
    copy_dim<U>(base_, o.origin(), o.strides(), 0);
  }
  template <typename U>
  void copy_dim(T* dst, U* src, const std::ptrdiff_t* sstride, std::size_t d) const {
    if (d + 1 == N) {
      for (std::size_t i = 0; i < shape_[d]; ++i)
        dst[(std::ptrdiff_t)i * stride_[d]] = src[(std::ptrdiff_t)i * sstride[d]];
    } else {
      for (std::size_t i = 0; i < shape_[d]; ++i)
        copy_dim<U>(dst + (std::ptrdiff_t)i * stride_[d], src + (std::ptrdiff_t)i * sstride[d],
                    sstride, d + 1);
    }
  }
  T* base_;
  std::size_t shape_[N];
  std::ptrdiff_t stride_[N];
};

}  // namespace pp2d_stub
}  // namespace detail

static const detail::pp2d_stub::extent_gen<0> extents = {};
static const detail::pp2d_stub::index_gen<0, 0> indices = {};

template <typename T, std::size_t N>
class const_multi_array_ref : public detail::pp2d_stub::view<const T, N> {
  typedef detail::pp2d_stub::view<const T, N> base;
 public:
  const_multi_array_ref(const T* p, const detail::pp2d_stub::extent_gen<N>& e) : base(p, e) {}
};

template <typename T, std::size_t N>
class multi_array_ref : public detail::pp2d_stub::view<T, N> {
  typedef detail::pp2d_stub::view<T, N> base;
 public:
  multi_array_ref(T* p, const detail::pp2d_stub::extent_gen<N>& e) : base(p, e) {}
  using base::operator=;
};

template <typename T, std::size_t N>
class multi_array : public detail::pp2d_stub::view<T, N> {
  typedef detail::pp2d_stub::view<T, N> base;
 public:
  explicit multi_array(const detail::pp2d_stub::extent_gen<N>& e) : base() {
    std::size_t n = 1;
    for (std::size_t i = 0; i < N; ++i) n *= e.e[i];
    store_.assign(n, T());
    this->base_ = store_.data();
    this->set_row_major(e);
  }
  multi_array(const multi_array& o) : base(), store_(o.store_) {
    this->base_ = store_.data();
    for (std::size_t i = 0; i < N; ++i) {
      this->shape_[i] = o.shape_[i];
      this->stride_[i] = o.stride_[i];
    }
  }
  multi_array& operator=(const multi_array& o) { this->copy_from(o); return *this; }
  template <typename U>
  multi_array& operator=(const detail::pp2d_stub::view<U, N>& o) { this->copy_from(o); return *this; }

 private:
  std::vector<T> store_;
};

}  // namespace boost
