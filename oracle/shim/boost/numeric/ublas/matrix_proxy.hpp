/* oracle/shim — matrix_row / matrix_column of the uBLAS stand-in (see matrix_sparse.hpp). */
#pragma once
#include "matrix_sparse.hpp"
namespace boost { namespace numeric { namespace ublas {
template <class M>
class matrix_row {
 public:
  typedef typename M::value_type value_type;
  matrix_row(const M& m, std::size_t i) : m_(&m), i_(i) {}
  class const_iterator {
   public:
    const_iterator(const M* m, std::size_t i, std::size_t pos) : m_(m), i_(i), pos_(pos) {}
    std::size_t index() const { return m_->row(i_)[pos_].first; }
    value_type operator*() const { return m_->row(i_)[pos_].second; }
    const_iterator& operator++() { ++pos_; return *this; }
    const_iterator operator++(int) { const_iterator t(*this); ++pos_; return t; }
    bool operator==(const const_iterator& o) const { return pos_ == o.pos_; }
    bool operator!=(const const_iterator& o) const { return pos_ != o.pos_; }
   private:
    const M* m_; std::size_t i_, pos_;
  };
  typedef const_iterator iterator;
  const_iterator begin() const { return const_iterator(m_, i_, 0); }
  const_iterator end() const { return const_iterator(m_, i_, m_->row(i_).size()); }
  std::size_t size() const { return m_->size2(); }
 private:
  const M* m_; std::size_t i_;
};
template <class M> class matrix_column {};
}}}
