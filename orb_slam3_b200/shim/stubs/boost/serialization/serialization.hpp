// SYNTAX-CHECK STAND-IN, not Boost.Serialization: the names the reference's headers use in their serialize() templates.
#pragma once
#include <cstddef>
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(T)
#define BOOST_CLASS_EXPORT_KEY(T)
#define BOOST_CLASS_EXPORT_KEY2(T, K)
#define BOOST_CLASS_EXPORT_IMPLEMENT(T)
#define BOOST_CLASS_EXPORT(T)
#define BOOST_CLASS_EXPORT_GUID(T, K)
#define BOOST_CLASS_VERSION(T, N)
namespace boost {
namespace serialization {
struct access {};
template <class Base, class Derived> Base& base_object(Derived& d) { return d; }
template <class T> struct array_wrapper { T* p; std::size_t n; };
template <class T> array_wrapper<T> make_array(T* p, std::size_t n) { return array_wrapper<T>{p, n}; }
template <class T> T& make_nvp(const char*, T& t) { return t; }
template <class Archive, class T> void split_free(Archive&, T&, const unsigned int) {}
template <class Archive, class T> void split_member(Archive&, T&, const unsigned int) {}
}  // namespace serialization
namespace archive { struct binary_iarchive; struct binary_oarchive; struct text_iarchive; struct text_oarchive; }
}  // namespace boost
