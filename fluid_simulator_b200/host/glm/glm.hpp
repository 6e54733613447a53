// Minimal glm subset (vec<2|3,T>, elementwise arithmetic, dot/length/normalize) for building the facade and its tests where
// g-truc/glm (header-only, the reference's only third-party dependency on this path, CMakeLists.txt:25) is not installed.
// A real glm on the include path takes precedence: put it before fluid_simulator_b200/host in -I order.
#pragma once
#include <cmath>
#include <cstddef>
#include <type_traits>

namespace glm {

template <int N, typename T> struct vec;

template <typename T> struct vec<2, T> {
    T x, y;
    constexpr vec() : x(T(0)), y(T(0)) {}
    constexpr explicit vec(T s) : x(s), y(s) {}
    template <typename A, typename B> constexpr vec(A a, B b) : x(T(a)), y(T(b)) {}
    template <typename U> constexpr vec(const vec<2, U>& o) : x(T(o.x)), y(T(o.y)) {}
    constexpr T& operator[](int i) { return i == 0 ? x : y; }
    constexpr const T& operator[](int i) const { return i == 0 ? x : y; }
};

template <typename T> struct vec<3, T> {
    T x, y, z;
    constexpr vec() : x(T(0)), y(T(0)), z(T(0)) {}
    constexpr explicit vec(T s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C> constexpr vec(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
    // glm allows implicit conversion between component types (dvec3 -> ivec3 truncates toward zero).
    template <typename U> constexpr vec(const vec<3, U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    constexpr T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

using dvec3 = vec<3, double>;
using vec3 = vec<3, float>;
using ivec3 = vec<3, int>;
using dvec2 = vec<2, double>;
using vec2 = vec<2, float>;
using ivec2 = vec<2, int>;

#define GLM_SHIM_BINOP(op)                                                                          \
    template <typename T> constexpr vec<3, T> operator op(const vec<3, T>& a, const vec<3, T>& b) { \
        return vec<3, T>(a.x op b.x, a.y op b.y, a.z op b.z);                                       \
    }                                                                                               \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>         \
    constexpr vec<3, T> operator op(const vec<3, T>& a, S s) {                                      \
        return vec<3, T>(a.x op T(s), a.y op T(s), a.z op T(s));                                    \
    }                                                                                               \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>         \
    constexpr vec<3, T> operator op(S s, const vec<3, T>& a) {                                      \
        return vec<3, T>(T(s) op a.x, T(s) op a.y, T(s) op a.z);                                    \
    }                                                                                               \
    template <typename T> constexpr vec<2, T> operator op(const vec<2, T>& a, const vec<2, T>& b) { \
        return vec<2, T>(a.x op b.x, a.y op b.y);                                                   \
    }                                                                                               \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>         \
    constexpr vec<2, T> operator op(const vec<2, T>& a, S s) {                                      \
        return vec<2, T>(a.x op T(s), a.y op T(s));                                                 \
    }                                                                                               \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>         \
    constexpr vec<2, T> operator op(S s, const vec<2, T>& a) {                                      \
        return vec<2, T>(T(s) op a.x, T(s) op a.y);                                                 \
    }
GLM_SHIM_BINOP(+)
GLM_SHIM_BINOP(-)
GLM_SHIM_BINOP(*)
GLM_SHIM_BINOP(/)
#undef GLM_SHIM_BINOP

template <typename T> constexpr vec<3, T> operator-(const vec<3, T>& a) { return vec<3, T>(-a.x, -a.y, -a.z); }
template <typename T> constexpr vec<2, T> operator-(const vec<2, T>& a) { return vec<2, T>(-a.x, -a.y); }

#define GLM_SHIM_ASSIGN(op)                                                                   \
    template <typename T> constexpr vec<3, T>& operator op##=(vec<3, T>& a, const vec<3, T>& b) { \
        a.x op## = b.x; a.y op## = b.y; a.z op## = b.z; return a;                             \
    }                                                                                         \
    template <typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>   \
    constexpr vec<3, T>& operator op##=(vec<3, T>& a, S s) {                                  \
        a.x op## = T(s); a.y op## = T(s); a.z op## = T(s); return a;                          \
    }
GLM_SHIM_ASSIGN(+)
GLM_SHIM_ASSIGN(-)
GLM_SHIM_ASSIGN(*)
GLM_SHIM_ASSIGN(/)
#undef GLM_SHIM_ASSIGN

template <typename T> constexpr bool operator==(const vec<3, T>& a, const vec<3, T>& b) {
    return a.x == b.x && a.y == b.y && a.z == b.z;
}
template <typename T> constexpr bool operator!=(const vec<3, T>& a, const vec<3, T>& b) { return !(a == b); }

template <typename T> constexpr T dot(const vec<3, T>& a, const vec<3, T>& b) {
    // glm's compute_dot<vec3>: tmp = a*b; return tmp.x + tmp.y + tmp.z
    return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
template <typename T> constexpr T dot(const vec<2, T>& a, const vec<2, T>& b) { return a.x * b.x + a.y * b.y; }
template <typename T> inline T length(const vec<3, T>& a) { return std::sqrt(dot(a, a)); }
template <typename T> inline T length(const vec<2, T>& a) { return std::sqrt(dot(a, a)); }
template <typename T> inline vec<3, T> normalize(const vec<3, T>& a) {
    // glm: v * inversesqrt(dot(v, v)) with inversesqrt(x) = 1 / sqrt(x)
    return a * (T(1) / std::sqrt(dot(a, a)));
}

}  // namespace glm
