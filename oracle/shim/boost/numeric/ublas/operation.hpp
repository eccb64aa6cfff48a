/* oracle/shim — inner_prod, trans, prod of the uBLAS stand-in (see matrix_sparse.hpp). */
#pragma once
#include "matrix_proxy.hpp"
namespace boost { namespace numeric { namespace ublas {
/* sparse row (x) dense vector: stored elements in ascending column order */
template <class M, class V>
double inner_prod(const matrix_row<M>& r, const vector<V>& v) {
  double s = 0;
  for (typename matrix_row<M>::const_iterator it = r.begin(); it != r.end(); ++it) s += (*it) * v[it.index()];
  return s;
}
template <class T>
compressed_matrix<T> trans(const compressed_matrix<T>& m) {
  compressed_matrix<T> t(m.size2(), m.size1());
  for (std::size_t i = 0; i < m.size1(); ++i)
    for (std::size_t q = 0; q < m.row(i).size(); ++q) t.set(m.row(i)[q].first, i, m.row(i)[q].second);
  return t;
}
template <class A, class B> void prod(const A&, const B&) {}
}}}
