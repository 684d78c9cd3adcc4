// tf2::Quaternion stand-in (oracle/ref_shim/README.md).  slerp, normalisation and setRPY follow tf2's published semantics.
#pragma once
#include "Vector3.h"
namespace tf2 {
class Quaternion {
   public:
    tf2Scalar m_floats[4];
    Quaternion() { m_floats[0] = m_floats[1] = m_floats[2] = m_floats[3] = 0.0; }
    Quaternion(tf2Scalar x, tf2Scalar y, tf2Scalar z, tf2Scalar w) { setValue(x, y, z, w); }
    void setValue(tf2Scalar x, tf2Scalar y, tf2Scalar z, tf2Scalar w) {
        m_floats[0] = x;
        m_floats[1] = y;
        m_floats[2] = z;
        m_floats[3] = w;
    }
    const tf2Scalar& x() const { return m_floats[0]; }
    const tf2Scalar& y() const { return m_floats[1]; }
    const tf2Scalar& z() const { return m_floats[2]; }
    const tf2Scalar& w() const { return m_floats[3]; }
    const tf2Scalar& getX() const { return m_floats[0]; }
    const tf2Scalar& getY() const { return m_floats[1]; }
    const tf2Scalar& getZ() const { return m_floats[2]; }
    const tf2Scalar& getW() const { return m_floats[3]; }
    tf2Scalar dot(const Quaternion& q) const { return m_floats[0] * q.x() + m_floats[1] * q.y() + m_floats[2] * q.z() + m_floats[3] * q.m_floats[3]; }
    tf2Scalar length2() const { return dot(*this); }
    tf2Scalar length() const { return std::sqrt(length2()); }
    Quaternion operator-() const { return Quaternion(-m_floats[0], -m_floats[1], -m_floats[2], -m_floats[3]); }
    Quaternion& operator/=(const tf2Scalar& s) {
        const tf2Scalar r = tf2Scalar(1.0) / s;
        m_floats[0] *= r;
        m_floats[1] *= r;
        m_floats[2] *= r;
        m_floats[3] *= r;
        return *this;
    }
    Quaternion operator/(const tf2Scalar& s) const {
        Quaternion q = *this;
        q /= s;
        return q;
    }
    Quaternion& normalize() { return *this /= length(); }
    Quaternion normalized() const { return *this / length(); }
    // half-angle products in tf2's order
    void setRPY(const tf2Scalar& roll, const tf2Scalar& pitch, const tf2Scalar& yaw) {
        const tf2Scalar halfYaw = yaw * 0.5, halfPitch = pitch * 0.5, halfRoll = roll * 0.5;
        const tf2Scalar cosYaw = std::cos(halfYaw), sinYaw = std::sin(halfYaw);
        const tf2Scalar cosPitch = std::cos(halfPitch), sinPitch = std::sin(halfPitch);
        const tf2Scalar cosRoll = std::cos(halfRoll), sinRoll = std::sin(halfRoll);
        setValue(sinRoll * cosPitch * cosYaw - cosRoll * sinPitch * sinYaw,  // x
                 cosRoll * sinPitch * cosYaw + sinRoll * cosPitch * sinYaw,  // y
                 cosRoll * cosPitch * sinYaw - sinRoll * sinPitch * cosYaw,  // z
                 cosRoll * cosPitch * cosYaw + sinRoll * sinPitch * sinYaw); // w
    }
    tf2Scalar angleShortestPath(const Quaternion& q) const {
        const tf2Scalar s = std::sqrt(length2() * q.length2());
        if (dot(q) < 0) return tf2Acos(dot(-q) / s) * tf2Scalar(2.0);
        return tf2Acos(dot(q) / s) * tf2Scalar(2.0);
    }
    Quaternion slerp(const Quaternion& q, const tf2Scalar& t) const {
        const tf2Scalar theta = angleShortestPath(q) / tf2Scalar(2.0);
        if (theta != tf2Scalar(0.0)) {
            const tf2Scalar d = tf2Scalar(1.0) / std::sin(theta);
            const tf2Scalar s0 = std::sin((tf2Scalar(1.0) - t) * theta);
            const tf2Scalar s1 = std::sin(t * theta);
            if (dot(q) < 0)
                return Quaternion((m_floats[0] * s0 + -q.x() * s1) * d, (m_floats[1] * s0 + -q.y() * s1) * d, (m_floats[2] * s0 + -q.z() * s1) * d,
                                  (m_floats[3] * s0 + -q.m_floats[3] * s1) * d);
            return Quaternion((m_floats[0] * s0 + q.x() * s1) * d, (m_floats[1] * s0 + q.y() * s1) * d, (m_floats[2] * s0 + q.z() * s1) * d,
                              (m_floats[3] * s0 + q.m_floats[3] * s1) * d);
        }
        return *this;
    }
};
}  // namespace tf2
