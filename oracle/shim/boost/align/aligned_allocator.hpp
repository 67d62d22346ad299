// Stand-in for boost::alignment::aligned_allocator (boost is not installed): just enough for the reference's
// serialisation/Serialiser.hpp, which keeps its byte stream in a std::vector with a 16-byte aligned allocator.
// Test infrastructure only (oracle/_ref build).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <new>

namespace boost {
namespace alignment {

template <class T, std::size_t Alignment>
struct aligned_allocator {
  using value_type = T;
  template <class U> struct rebind { using other = aligned_allocator<U, Alignment>; };
  aligned_allocator() = default;
  template <class U> aligned_allocator(const aligned_allocator<U, Alignment>&) {}
  T* allocate(std::size_t n) {
    const std::size_t bytes = ((n * sizeof(T) + Alignment - 1) / Alignment) * Alignment;
    void* p = std::aligned_alloc(Alignment, bytes ? bytes : Alignment);
    if (!p) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, std::size_t) { std::free(p); }
  template <class U> bool operator==(const aligned_allocator<U, Alignment>&) const { return true; }
  template <class U> bool operator!=(const aligned_allocator<U, Alignment>&) const { return false; }
};

}  // namespace alignment
}  // namespace boost
