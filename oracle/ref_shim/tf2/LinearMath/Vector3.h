// tf2::Vector3 stand-in (oracle/ref_shim/README.md): double precision, the members the reference uses.
#pragma once
#include <cmath>
typedef double tf2Scalar;
// tf2Scalar.h clamps the argument of the inverse trigonometric functions to [-1, 1]
inline tf2Scalar tf2Acos(tf2Scalar x) { return std::acos(x < tf2Scalar(-1) ? tf2Scalar(-1) : (x > tf2Scalar(1) ? tf2Scalar(1) : x)); }
inline tf2Scalar tf2Asin(tf2Scalar x) { return std::asin(x < tf2Scalar(-1) ? tf2Scalar(-1) : (x > tf2Scalar(1) ? tf2Scalar(1) : x)); }
namespace tf2 {
class Vector3 {
   public:
    tf2Scalar m_floats[4];
    Vector3() { m_floats[0] = m_floats[1] = m_floats[2] = m_floats[3] = 0.0; }
    Vector3(tf2Scalar x, tf2Scalar y, tf2Scalar z) { setValue(x, y, z); }
    void setValue(tf2Scalar x, tf2Scalar y, tf2Scalar z) {
        m_floats[0] = x;
        m_floats[1] = y;
        m_floats[2] = z;
        m_floats[3] = 0.0;
    }
    const tf2Scalar& x() const { return m_floats[0]; }
    const tf2Scalar& y() const { return m_floats[1]; }
    const tf2Scalar& z() const { return m_floats[2]; }
    const tf2Scalar& getX() const { return m_floats[0]; }
    const tf2Scalar& getY() const { return m_floats[1]; }
    const tf2Scalar& getZ() const { return m_floats[2]; }
    void setX(tf2Scalar v) { m_floats[0] = v; }
    void setY(tf2Scalar v) { m_floats[1] = v; }
    void setZ(tf2Scalar v) { m_floats[2] = v; }
    tf2Scalar& operator[](int i) { return m_floats[i]; }
    const tf2Scalar& operator[](int i) const { return m_floats[i]; }
    Vector3& operator+=(const Vector3& v) {
        m_floats[0] += v.m_floats[0];
        m_floats[1] += v.m_floats[1];
        m_floats[2] += v.m_floats[2];
        return *this;
    }
    Vector3& operator-=(const Vector3& v) {
        m_floats[0] -= v.m_floats[0];
        m_floats[1] -= v.m_floats[1];
        m_floats[2] -= v.m_floats[2];
        return *this;
    }
    Vector3& operator*=(const tf2Scalar& s) {
        m_floats[0] *= s;
        m_floats[1] *= s;
        m_floats[2] *= s;
        return *this;
    }
    tf2Scalar dot(const Vector3& v) const { return m_floats[0] * v.m_floats[0] + m_floats[1] * v.m_floats[1] + m_floats[2] * v.m_floats[2]; }
    tf2Scalar length2() const { return dot(*this); }
    tf2Scalar length() const { return std::sqrt(length2()); }
};
inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.m_floats[0] + b.m_floats[0], a.m_floats[1] + b.m_floats[1], a.m_floats[2] + b.m_floats[2]); }
inline Vector3 operator-(const Vector3& a, const Vector3& b) { return Vector3(a.m_floats[0] - b.m_floats[0], a.m_floats[1] - b.m_floats[1], a.m_floats[2] - b.m_floats[2]); }
inline Vector3 operator-(const Vector3& v) { return Vector3(-v.m_floats[0], -v.m_floats[1], -v.m_floats[2]); }
inline Vector3 operator*(const Vector3& v, const tf2Scalar& s) { return Vector3(v.m_floats[0] * s, v.m_floats[1] * s, v.m_floats[2] * s); }
inline Vector3 operator*(const tf2Scalar& s, const Vector3& v) { return v * s; }
}  // namespace tf2
