/* oracle/shim — a MINIMAL stand-in for the part of Boost uBLAS that eturro/mmseq's
 * src/mmseq.cpp and src/uh.cpp use, so that the UNMODIFIED reference sources compile in an
 * image without Boost (oracle/Makefile, target ref_mmseq).  Test infrastructure only.
 * Semantics reproduced (row-major compressed_matrix):
 *   - iterator1 from begin1() visits every row index 0..size1-1; its begin()/end() give an
 *     iterator2 over the stored elements of that row in ascending column order;
 *   - iterator2 from begin2() visits every column index; its begin()/end() give an iterator1
 *     over the stored elements of that column in ascending row order;
 *   - M(i,j) = v stores an element (also a zero, as uBLAS does), M(i,j) reads 0 when absent;
 *   - resize(s1, s2, false) drops the contents. */
#pragma once
/* the real Boost headers pull these in transitively; the reference relies on that */
#include <stdint.h>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <string>

#include <algorithm>
#include <cstddef>
#include <iterator>
#include <utility>
#include <vector>

namespace boost { namespace numeric { namespace ublas {

template <class T>
class vector {
 public:
  vector() {}
  explicit vector(std::size_t n) : d_(n) {}
  std::size_t size() const { return d_.size(); }
  T& operator[](std::size_t i) { return d_[i]; }
  const T& operator[](std::size_t i) const { return d_[i]; }
  T& operator()(std::size_t i) { return d_[i]; }
  const T& operator()(std::size_t i) const { return d_[i]; }
 private:
  std::vector<T> d_;
};

template <class T>
class compressed_matrix {
  typedef std::vector<std::pair<std::size_t, T> > row_t;
 public:
  typedef T value_type;
  typedef std::size_t size_type;
  compressed_matrix() : s1_(0), s2_(0) {}
  compressed_matrix(size_type s1, size_type s2, size_type = 0) : s1_(s1), s2_(s2), rows_(s1) {}
  size_type size1() const { return s1_; }
  size_type size2() const { return s2_; }
  void resize(size_type s1, size_type s2, bool preserve = true) {
    if (!preserve) rows_.clear();
    rows_.resize(s1);
    if (preserve) for (size_type i = 0; i < rows_.size(); ++i) while (!rows_[i].empty() && rows_[i].back().first >= s2) rows_[i].pop_back();
    s1_ = s1; s2_ = s2;
  }
  /* element proxy */
  class reference {
   public:
    reference(compressed_matrix& m, size_type i, size_type j) : m_(m), i_(i), j_(j) {}
    reference& operator=(const T& v) { m_.set(i_, j_, v); return *this; }
    reference& operator=(const reference& o) { m_.set(i_, j_, (T)o); return *this; }
    operator T() const { return m_.get(i_, j_); }
   private:
    compressed_matrix& m_; size_type i_, j_;
  };
  reference operator()(size_type i, size_type j) { return reference(*this, i, j); }
  T operator()(size_type i, size_type j) const { return get(i, j); }
  T get(size_type i, size_type j) const {
    const row_t& r = rows_[i];
    typename row_t::const_iterator it = std::lower_bound(r.begin(), r.end(), j, cmp());
    return (it != r.end() && it->first == j) ? it->second : T();
  }
  void set(size_type i, size_type j, const T& v) {
    row_t& r = rows_[i];
    typename row_t::iterator it = std::lower_bound(r.begin(), r.end(), j, cmp());
    if (it != r.end() && it->first == j) it->second = v; else r.insert(it, std::make_pair(j, v));
  }
  size_type nnz() const { size_type c = 0; for (size_type i = 0; i < rows_.size(); ++i) c += rows_[i].size(); return c; }

  class const_iterator1;
  class const_iterator2;
  /* iterator along dimension 1 (rows).  dense == true: every row index at column j_ (from begin1());
   * dense == false: stored elements of column j_ (from an iterator2's begin()). */
  class const_iterator1 {
   public:
    typedef std::forward_iterator_tag iterator_category;
    typedef T value_type; typedef std::ptrdiff_t difference_type; typedef const T* pointer; typedef T reference;
    const_iterator1() : m_(0), i_(0), j_(0), dense_(true) {}
    const_iterator1(const compressed_matrix* m, size_type i, size_type j, bool dense) : m_(m), i_(i), j_(j), dense_(dense) { if (!dense_) skip(); }
    size_type index1() const { return i_; }
    size_type index2() const { return j_; }
    T operator*() const { return m_->get(i_, j_); }
    const_iterator1& operator++() { ++i_; if (!dense_) skip(); return *this; }
    const_iterator1 operator++(int) { const_iterator1 t(*this); ++*this; return t; }
    bool operator==(const const_iterator1& o) const { return i_ == o.i_; }
    bool operator!=(const const_iterator1& o) const { return i_ != o.i_; }
    const_iterator2 begin() const { return const_iterator2(m_, i_, 0, false); }
    const_iterator2 end() const { return const_iterator2(m_, i_, m_->s2_, false, true); }
   private:
    void skip() { while (i_ < m_->s1_ && !m_->has(i_, j_)) ++i_; }
    const compressed_matrix* m_; size_type i_, j_; bool dense_;
  };
  /* iterator along dimension 2 (columns).  dense: every column index (from begin2());
   * otherwise the stored elements of row i_ (from an iterator1's begin()). */
  class const_iterator2 {
   public:
    typedef std::forward_iterator_tag iterator_category;
    typedef T value_type; typedef std::ptrdiff_t difference_type; typedef const T* pointer; typedef T reference;
    const_iterator2() : m_(0), i_(0), j_(0), pos_(0), dense_(true) {}
    const_iterator2(const compressed_matrix* m, size_type i, size_type j, bool dense, bool at_end = false) : m_(m), i_(i), j_(j), pos_(0), dense_(dense) {
      if (!dense_) { pos_ = at_end ? m_->rows_[i_].size() : 0; sync(); }
    }
    size_type index1() const { return i_; }
    size_type index2() const { return j_; }
    T operator*() const { return dense_ ? m_->get(i_, j_) : m_->rows_[i_][pos_].second; }
    const_iterator2& operator++() { if (dense_) ++j_; else { ++pos_; sync(); } return *this; }
    const_iterator2 operator++(int) { const_iterator2 t(*this); ++*this; return t; }
    bool operator==(const const_iterator2& o) const { return j_ == o.j_; }
    bool operator!=(const const_iterator2& o) const { return j_ != o.j_; }
    const_iterator1 begin() const { return const_iterator1(m_, 0, j_, false); }
    const_iterator1 end() const { return const_iterator1(m_, m_->s1_, j_, true); }
   private:
    void sync() { j_ = pos_ < m_->rows_[i_].size() ? m_->rows_[i_][pos_].first : m_->s2_; }
    const compressed_matrix* m_; size_type i_, j_, pos_; bool dense_;
  };
  typedef const_iterator1 iterator1;
  typedef const_iterator2 iterator2;
  const_iterator1 begin1() const { return const_iterator1(this, 0, 0, true); }
  const_iterator1 end1() const { return const_iterator1(this, s1_, 0, true); }
  const_iterator2 begin2() const { return const_iterator2(this, 0, 0, true); }
  const_iterator2 end2() const { return const_iterator2(this, 0, s2_, true); }

  bool has(size_type i, size_type j) const {
    const row_t& r = rows_[i];
    typename row_t::const_iterator it = std::lower_bound(r.begin(), r.end(), j, cmp());
    return it != r.end() && it->first == j;
  }
  const row_t& row(size_type i) const { return rows_[i]; }
 private:
  struct cmp { bool operator()(const std::pair<std::size_t, T>& a, std::size_t j) const { return a.first < j; } };
  size_type s1_, s2_;
  std::vector<row_t> rows_;
};

}}}  // namespace boost::numeric::ublas
