// TEST INFRASTRUCTURE ONLY: a serial stand-in for the three TBB facilities src/chaining.cpp uses, so that the file compiles
// VERBATIM without TBB (oracle/Makefile). Results do not depend on the execution order: the points collected by the nested
// parallel_for are sorted with a total order (comparePoint) before anything reads them.
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>
namespace tbb {
template <class I, class F>
void parallel_for(I begin, I end, const F& f) {
    for (I i = begin; i < end; ++i) f(i);
}
template <class It, class Cmp>
void parallel_sort(It a, It b, Cmp c) {
    std::sort(a, b, c);
}
template <class T>
class concurrent_vector : public std::vector<T> {
  public:
    using std::vector<T>::vector;
};
}  // namespace tbb
