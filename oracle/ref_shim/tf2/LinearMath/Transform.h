// tf2::Transform stand-in (oracle/ref_shim/README.md): basis + origin, composition and inverse in tf2's order of operations.
#pragma once
#include "Matrix3x3.h"
namespace tf2 {
class Transform {
   public:
    Matrix3x3 m_basis;
    Vector3 m_origin;
    Transform() {}
    explicit Transform(const Quaternion& q, const Vector3& c = Vector3(0, 0, 0)) : m_basis(q), m_origin(c) {}
    explicit Transform(const Matrix3x3& b, const Vector3& c = Vector3(0, 0, 0)) : m_basis(b), m_origin(c) {}
    Vector3 operator()(const Vector3& x) const { return Vector3(m_basis[0].dot(x) + m_origin.x(), m_basis[1].dot(x) + m_origin.y(), m_basis[2].dot(x) + m_origin.z()); }
    Vector3 operator*(const Vector3& x) const { return (*this)(x); }
    Matrix3x3& getBasis() { return m_basis; }
    const Matrix3x3& getBasis() const { return m_basis; }
    Vector3& getOrigin() { return m_origin; }
    const Vector3& getOrigin() const { return m_origin; }
    Quaternion getRotation() const {
        Quaternion q;
        m_basis.getRotation(q);
        return q;
    }
    void setOrigin(const Vector3& o) { m_origin = o; }
    void setBasis(const Matrix3x3& b) { m_basis = b; }
    void setRotation(const Quaternion& q) { m_basis.setRotation(q); }
    void setIdentity() {
        m_basis.setIdentity();
        m_origin.setValue(0, 0, 0);
    }
    Transform& operator*=(const Transform& t) {
        m_origin += m_basis * t.m_origin;
        m_basis *= t.m_basis;
        return *this;
    }
    Transform inverse() const {
        const Matrix3x3 inv = m_basis.transpose();
        return Transform(inv, inv * -m_origin);
    }
    Transform operator*(const Transform& t) const { return Transform(m_basis * t.m_basis, (*this)(t.m_origin)); }
};
}  // namespace tf2
