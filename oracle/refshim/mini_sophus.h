// mini_sophus.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for the part of Sophus (README.md:46 pins commit a0fe89a; not in the image) that
// the reference's wheel / plane factor code and utility/sophus_utils.hpp touch: SO3 with exp / log / hat / matrix / unit_quaternion /
// composition and the two epsilon constants.  Written for this repository from Sophus' published formulas (same small-angle branches and
// thresholds as SURVEY Appendix A / oracle/vo_math.h); SE3, Sim3 and RxSO3 are only declared -- sophus_utils.hpp names them inside
// templates that the factor code never instantiates.
#pragma once
#include "mini_eigen.h"

#define EIGEN_STATIC_ASSERT_FIXED_SIZE(T) static_assert((int)T::RowsAtCompileTime != Eigen::Dynamic && (int)T::ColsAtCompileTime != Eigen::Dynamic, "fixed size expected")
#define EIGEN_STATIC_ASSERT_VECTOR_SPECIFIC_SIZE(T, N) static_assert((int)T::RowsAtCompileTime * (int)T::ColsAtCompileTime == (N), "vector size")
#define EIGEN_STATIC_ASSERT_MATRIX_SPECIFIC_SIZE(T, R, C) static_assert((int)T::RowsAtCompileTime == (R) && (int)T::ColsAtCompileTime == (C), "matrix size")

namespace Sophus {

template <typename Scalar> struct Constants {
    static Scalar epsilon() { return Scalar(1e-10); }
    static Scalar epsilonSqrt() { return std::sqrt(epsilon()); }
    static Scalar pi() { return Scalar(3.141592653589793238462643383279502884); }
};
template <typename Scalar> class SE3;
template <typename Scalar> class Sim3;
template <typename Scalar> class RxSO3;

template <typename Scalar_> class SO3 {
  public:
    typedef Scalar_ Scalar;
    typedef Eigen::Matrix<Scalar, 3, 1> Tangent;
    typedef Eigen::Matrix<Scalar, 3, 1> Point;
    typedef Eigen::Matrix<Scalar, 3, 3> Transformation;
    SO3() : q_(Eigen::Quaternion<Scalar>::Identity()) {}
    template <typename D> explicit SO3(const Eigen::QuaternionBase<D> &q) : q_(q) { q_.normalize(); }       // the constructor normalises
    explicit SO3(const Transformation &R) : q_(R) {}
    static SO3 exp(const Tangent &omega) {
        const Scalar theta_sq = omega.squaredNorm(), theta = std::sqrt(theta_sq), half_theta = Scalar(0.5) * theta;
        Scalar imag_factor, real_factor;
        if (theta < Constants<Scalar>::epsilon()) {
            const Scalar theta_po4 = theta_sq * theta_sq;
            imag_factor = Scalar(0.5) - Scalar(1.0 / 48.0) * theta_sq + Scalar(1.0 / 3840.0) * theta_po4;
            real_factor = Scalar(1) - Scalar(1.0 / 8.0) * theta_sq + Scalar(1.0 / 384.0) * theta_po4;
        } else {
            imag_factor = std::sin(half_theta) / theta;
            real_factor = std::cos(half_theta);
        }
        SO3 r; r.q_ = Eigen::Quaternion<Scalar>(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z());
        return r;
    }
    template <typename D> static SO3 exp(const Eigen::MatrixBase<D> &omega) { return exp(Tangent(omega)); }
    Tangent log() const {
        const Scalar squared_n = q_.x() * q_.x() + q_.y() * q_.y() + q_.z() * q_.z(), n = std::sqrt(squared_n), w = q_.w();
        Scalar two_atan_nbyw_by_n;
        if (n < Constants<Scalar>::epsilon()) {
            const Scalar squared_w = w * w;
            two_atan_nbyw_by_n = Scalar(2) / w - Scalar(2) * squared_n / (w * squared_w);
        } else if (std::abs(w) < Constants<Scalar>::epsilon()) {
            two_atan_nbyw_by_n = (w > Scalar(0) ? Constants<Scalar>::pi() : -Constants<Scalar>::pi()) / n;
        } else {
            two_atan_nbyw_by_n = Scalar(2) * std::atan(n / w) / n;
        }
        return Tangent(two_atan_nbyw_by_n * q_.x(), two_atan_nbyw_by_n * q_.y(), two_atan_nbyw_by_n * q_.z());
    }
    template <typename D> static Transformation hat(const Eigen::MatrixBase<D> &omega) {
        Transformation O;
        O << Scalar(0), -omega(2), omega(1), omega(2), Scalar(0), -omega(0), -omega(1), omega(0), Scalar(0);
        return O;
    }
    Transformation matrix() const { return q_.toRotationMatrix(); }
    const Eigen::Quaternion<Scalar> &unit_quaternion() const { return q_; }
    SO3 inverse() const { SO3 r; r.q_ = q_.conjugate(); return r; }
    SO3 operator*(const SO3 &o) const {                              // group product, re-normalised only when the norm has drifted
        SO3 r; r.q_ = q_ * o.q_;
        const Scalar sn = r.q_.squaredNorm();
        if (sn != Scalar(1)) { const Scalar s = Scalar(2.0) / (Scalar(1.0) + sn); r.q_ = Eigen::Quaternion<Scalar>(r.q_.w() * s, r.q_.x() * s, r.q_.y() * s, r.q_.z() * s); }
        return r;
    }
    template <typename D> Point operator*(const Eigen::MatrixBase<D> &p) const { return q_ * p; }
  private:
    Eigen::Quaternion<Scalar> q_;
};
typedef SO3<double> SO3d;
typedef SO3<float> SO3f;

}  // namespace Sophus
