// tf2::Matrix3x3 stand-in (oracle/ref_shim/README.md): rows as Vector3; rotation <-> quaternion, Euler YPR in tf2's conventions.
#pragma once
#include "Quaternion.h"
namespace tf2 {
class Matrix3x3 {
   public:
    Vector3 m_el[3];
    Matrix3x3() {}
    explicit Matrix3x3(const Quaternion& q) { setRotation(q); }
    Matrix3x3(tf2Scalar xx, tf2Scalar xy, tf2Scalar xz, tf2Scalar yx, tf2Scalar yy, tf2Scalar yz, tf2Scalar zx, tf2Scalar zy, tf2Scalar zz) { setValue(xx, xy, xz, yx, yy, yz, zx, zy, zz); }
    void setValue(tf2Scalar xx, tf2Scalar xy, tf2Scalar xz, tf2Scalar yx, tf2Scalar yy, tf2Scalar yz, tf2Scalar zx, tf2Scalar zy, tf2Scalar zz) {
        m_el[0].setValue(xx, xy, xz);
        m_el[1].setValue(yx, yy, yz);
        m_el[2].setValue(zx, zy, zz);
    }
    void setIdentity() { setValue(1, 0, 0, 0, 1, 0, 0, 0, 1); }
    Vector3& operator[](int i) { return m_el[i]; }
    const Vector3& operator[](int i) const { return m_el[i]; }
    const Vector3& getRow(int i) const { return m_el[i]; }
    void setRotation(const Quaternion& q) {
        const tf2Scalar d = q.length2();
        const tf2Scalar s = tf2Scalar(2.0) / d;
        const tf2Scalar xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
        const tf2Scalar wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
        const tf2Scalar xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
        const tf2Scalar yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
        setValue(tf2Scalar(1.0) - (yy + zz), xy - wz, xz + wy, xy + wz, tf2Scalar(1.0) - (xx + zz), yz - wx, xz - wy, yz + wx, tf2Scalar(1.0) - (xx + yy));
    }
    void getRotation(Quaternion& q) const {
        const tf2Scalar trace = m_el[0].x() + m_el[1].y() + m_el[2].z();
        tf2Scalar temp[4];
        if (trace > tf2Scalar(0.0)) {
            tf2Scalar s = std::sqrt(trace + tf2Scalar(1.0));
            temp[3] = s * tf2Scalar(0.5);
            s = tf2Scalar(0.5) / s;
            temp[0] = (m_el[2].y() - m_el[1].z()) * s;
            temp[1] = (m_el[0].z() - m_el[2].x()) * s;
            temp[2] = (m_el[1].x() - m_el[0].y()) * s;
        } else {
            const int i = m_el[0].x() < m_el[1].y() ? (m_el[1].y() < m_el[2].z() ? 2 : 1) : (m_el[0].x() < m_el[2].z() ? 2 : 0);
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            tf2Scalar s = std::sqrt(m_el[i][i] - m_el[j][j] - m_el[k][k] + tf2Scalar(1.0));
            temp[i] = s * tf2Scalar(0.5);
            s = tf2Scalar(0.5) / s;
            temp[3] = (m_el[k][j] - m_el[j][k]) * s;
            temp[j] = (m_el[j][i] + m_el[i][j]) * s;
            temp[k] = (m_el[k][i] + m_el[i][k]) * s;
        }
        q.setValue(temp[0], temp[1], temp[2], temp[3]);
    }
    void setEulerYPR(tf2Scalar eulerZ, tf2Scalar eulerY, tf2Scalar eulerX) {
        const tf2Scalar ci = std::cos(eulerX), cj = std::cos(eulerY), ch = std::cos(eulerZ);
        const tf2Scalar si = std::sin(eulerX), sj = std::sin(eulerY), sh = std::sin(eulerZ);
        const tf2Scalar cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
        setValue(cj * ch, sj * sc - cs, sj * cc + ss, cj * sh, sj * ss + cc, sj * cs - sc, -sj, cj * si, cj * ci);
    }
    void setRPY(tf2Scalar roll, tf2Scalar pitch, tf2Scalar yaw) { setEulerYPR(yaw, pitch, roll); }
    void getEulerYPR(tf2Scalar& yaw, tf2Scalar& pitch, tf2Scalar& roll, unsigned int solution_number = 1) const {
        struct Euler {
            tf2Scalar yaw, pitch, roll;
        } e1, e2;
        const tf2Scalar PI = 3.14159265358979323846;
        if (std::fabs(m_el[2].x()) >= 1) {  // gimbal lock
            e1.yaw = 0;
            e2.yaw = 0;
            const tf2Scalar delta = std::atan2(m_el[2].y(), m_el[2].z());
            if (m_el[2].x() < 0) {
                e1.pitch = PI / tf2Scalar(2.0);
                e2.pitch = PI / tf2Scalar(2.0);
                e1.roll = delta;
                e2.roll = delta;
            } else {
                e1.pitch = -PI / tf2Scalar(2.0);
                e2.pitch = -PI / tf2Scalar(2.0);
                e1.roll = delta;
                e2.roll = delta;
            }
        } else {
            e1.pitch = -tf2Asin(m_el[2].x());
            e2.pitch = PI - e1.pitch;
            e1.roll = std::atan2(m_el[2].y() / std::cos(e1.pitch), m_el[2].z() / std::cos(e1.pitch));
            e2.roll = std::atan2(m_el[2].y() / std::cos(e2.pitch), m_el[2].z() / std::cos(e2.pitch));
            e1.yaw = std::atan2(m_el[1].x() / std::cos(e1.pitch), m_el[0].x() / std::cos(e1.pitch));
            e2.yaw = std::atan2(m_el[1].x() / std::cos(e2.pitch), m_el[0].x() / std::cos(e2.pitch));
        }
        const Euler& e = solution_number == 1 ? e1 : e2;
        yaw = e.yaw;
        pitch = e.pitch;
        roll = e.roll;
    }
    void getRPY(tf2Scalar& roll, tf2Scalar& pitch, tf2Scalar& yaw, unsigned int solution_number = 1) const { getEulerYPR(yaw, pitch, roll, solution_number); }
    Matrix3x3 transpose() const { return Matrix3x3(m_el[0].x(), m_el[1].x(), m_el[2].x(), m_el[0].y(), m_el[1].y(), m_el[2].y(), m_el[0].z(), m_el[1].z(), m_el[2].z()); }
    tf2Scalar tdotx(const Vector3& v) const { return m_el[0].x() * v.x() + m_el[1].x() * v.y() + m_el[2].x() * v.z(); }
    tf2Scalar tdoty(const Vector3& v) const { return m_el[0].y() * v.x() + m_el[1].y() * v.y() + m_el[2].y() * v.z(); }
    tf2Scalar tdotz(const Vector3& v) const { return m_el[0].z() * v.x() + m_el[1].z() * v.y() + m_el[2].z() * v.z(); }
    Matrix3x3& operator*=(const Matrix3x3& m) {
        setValue(m.tdotx(m_el[0]), m.tdoty(m_el[0]), m.tdotz(m_el[0]), m.tdotx(m_el[1]), m.tdoty(m_el[1]), m.tdotz(m_el[1]), m.tdotx(m_el[2]), m.tdoty(m_el[2]), m.tdotz(m_el[2]));
        return *this;
    }
};
inline Vector3 operator*(const Matrix3x3& m, const Vector3& v) { return Vector3(m[0].dot(v), m[1].dot(v), m[2].dot(v)); }
inline Matrix3x3 operator*(const Matrix3x3& m1, const Matrix3x3& m2) {
    return Matrix3x3(m2.tdotx(m1[0]), m2.tdoty(m1[0]), m2.tdotz(m1[0]), m2.tdotx(m1[1]), m2.tdoty(m1[1]), m2.tdotz(m1[1]), m2.tdotx(m1[2]), m2.tdoty(m1[2]), m2.tdotz(m1[2]));
}
}  // namespace tf2
