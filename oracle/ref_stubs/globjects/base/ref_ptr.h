// ORACLE BUILD STUB (test infrastructure): minimal globjects::ref_ptr so VolumeSampler's members compile. No GL.
#ifndef RR_REF_STUB_REF_PTR_H
#define RR_REF_STUB_REF_PTR_H
#include <memory>
namespace globjects {
template <typename T>
class ref_ptr {
 public:
  ref_ptr() = default;
  ref_ptr(T* p) : m_p(p) {}
  T* operator->() const { return m_p.get(); }
  T* get() const { return m_p.get(); }
  operator T*() const { return m_p.get(); }
 private:
  std::shared_ptr<T> m_p;
};
}  // namespace globjects
#endif
