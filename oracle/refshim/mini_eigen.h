// mini_eigen.h -- TEST INFRASTRUCTURE ONLY.  A small, eager, fixed-size stand-in for the subset of Eigen 3 that the reference's factor
// sources use (vins_estimator/src/factor/*.{h,cpp}, utility/utility.h), written from scratch for this repository so that those
// sources compile UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/libviw_ref.so (oracle/Makefile: `make ref`).
// The image has no Eigen; this is not Eigen and shares no code with it.  Every operation returns a concrete Matrix (no expression
// templates); sums run left to right over the inner index, which is also the order Eigen's fixed-size products reduce to without
// vectorisation -- agreement with the restated oracle is checked to ~1e-12 relative, not bit for bit (tests/test_reference_factors.py).
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <algorithm>
#include <numeric>
#include <functional>
#include <vector>
#include <map>
#include <cstring>
#include <limits>
#include <string>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };
enum { ComputeThinU = 4, ComputeThinV = 8, ComputeFullU = 16, ComputeFullV = 32 };
typedef std::ptrdiff_t Index;

template <typename T, int R, int C, int Opt = ColMajor> class Matrix;
template <typename XprType, int BR, int BC> class Block;
template <typename PlainType> class Map;
template <typename T> class Quaternion;
template <typename MatrixType> class JacobiSVD;
template <typename MatrixType> class SelfAdjointEigenSolver;
template <typename MatrixType> class LDLT;
template <typename D> struct traits;
template <typename XprType> class DynBlock;
template <typename T> class ArrayX;

template <typename T, int R, int C, int Opt> struct traits<Matrix<T, R, C, Opt>> { typedef T Scalar; enum { Rows = R, Cols = C }; };
template <typename X, int BR, int BC> struct traits<Block<X, BR, BC>> { typedef typename traits<X>::Scalar Scalar; enum { Rows = BR, Cols = BC }; };
template <typename X> struct traits<DynBlock<X>> { typedef typename traits<X>::Scalar Scalar; enum { Rows = Dynamic, Cols = Dynamic }; };
template <typename P> struct traits<Map<P>> { typedef typename traits<typename std::remove_const<P>::type>::Scalar Scalar; enum { Rows = traits<typename std::remove_const<P>::type>::Rows, Cols = traits<typename std::remove_const<P>::type>::Cols }; };

template <typename Derived> class MatrixBase;
// the concrete result type of an operation: fixed when both extents are known at compile time, else run-time sized
template <typename T, int R, int C> struct plain_type { typedef Matrix<T, (R == Dynamic || C == Dynamic) ? Dynamic : R, (R == Dynamic || C == Dynamic) ? Dynamic : C> type; };
struct SizeTag {};
template <typename P> P make_plain(int r, int c) { return P(SizeTag(), r, c); }
inline constexpr int pick_dim(int a, int b) { return a != Dynamic ? a : b; }

// comma initialiser: m << a, b, c, ... in row-major reading order
template <typename Derived> class CommaInit {
    Derived &m; int k;
  public:
    typedef typename traits<Derived>::Scalar Scalar;
    CommaInit(Derived &m_, Scalar first) : m(m_), k(0) { put(first); }
    void put(Scalar v) { const int cols = m.cols(); assert(k < m.rows() * cols); m.coeffRef(k / cols, k % cols) = v; k++; }
    CommaInit &operator,(Scalar v) { put(v); return *this; }
};


template <typename Derived> class MatrixBase {
  public:
    typedef typename traits<Derived>::Scalar Scalar;
    enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols, SizeAtCompileTime = RowsAtCompileTime * ColsAtCompileTime };
    typedef typename plain_type<Scalar, RowsAtCompileTime, ColsAtCompileTime>::type PlainObject;
    typedef typename plain_type<Scalar, ColsAtCompileTime, RowsAtCompileTime>::type TransposeReturn;
    Derived &derived() { return *static_cast<Derived *>(this); }
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    int rows() const { return derived().rowsImpl(); }
    int cols() const { return derived().colsImpl(); }
    int size() const { return rows() * cols(); }
    PlainObject plain() const { return make_plain<PlainObject>(rows(), cols()); }
    Scalar coeff(int i, int j) const { return derived().coeff(i, j); }
    Scalar &coeffRef(int i, int j) { return derived().coeffRef(i, j); }
    Scalar operator()(int i, int j) const { return coeff(i, j); }
    Scalar &operator()(int i, int j) { return coeffRef(i, j); }
    // vector access (row or column vectors)
    Scalar operator()(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
    Scalar &operator()(int i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
    Scalar operator[](int i) const { return (*this)(i); }
    Scalar &operator[](int i) { return (*this)(i); }
    Scalar x() const { return (*this)(0); } Scalar &x() { return (*this)(0); }
    Scalar y() const { return (*this)(1); } Scalar &y() { return (*this)(1); }
    Scalar z() const { return (*this)(2); } Scalar &z() { return (*this)(2); }
    Scalar w() const { return (*this)(3); } Scalar &w() { return (*this)(3); }

    PlainObject eval() const { PlainObject r = plain(); for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r.coeffRef(i, j) = coeff(i, j); return r; }
    TransposeReturn transpose() const { TransposeReturn r = make_plain<TransposeReturn>(cols(), rows()); for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r.coeffRef(j, i) = coeff(i, j); return r; }

    // ---- fixed-size sub-blocks (views that can be read and assigned)
    template <int BR, int BC> Block<Derived, BR, BC> block(int i, int j) { return Block<Derived, BR, BC>(derived(), i, j); }
    template <int BR, int BC> const Block<Derived, BR, BC> block(int i, int j) const { return Block<Derived, BR, BC>(const_cast<Derived &>(derived()), i, j); }
    template <int N> Block<Derived, N, ColsAtCompileTime> topRows() { return Block<Derived, N, ColsAtCompileTime>(derived(), 0, 0, N, cols()); }
    template <int N> const Block<Derived, N, ColsAtCompileTime> topRows() const { return Block<Derived, N, ColsAtCompileTime>(const_cast<Derived &>(derived()), 0, 0, N, cols()); }
    template <int N> Block<Derived, N, ColsAtCompileTime> bottomRows() { return Block<Derived, N, ColsAtCompileTime>(derived(), rows() - N, 0, N, cols()); }
    template <int N> const Block<Derived, N, ColsAtCompileTime> bottomRows() const { return Block<Derived, N, ColsAtCompileTime>(const_cast<Derived &>(derived()), rows() - N, 0, N, cols()); }
    template <int N> Block<Derived, RowsAtCompileTime, N> leftCols() { return Block<Derived, RowsAtCompileTime, N>(derived(), 0, 0, rows(), N); }
    template <int N> const Block<Derived, RowsAtCompileTime, N> leftCols() const { return Block<Derived, RowsAtCompileTime, N>(const_cast<Derived &>(derived()), 0, 0, rows(), N); }
    template <int N> Block<Derived, RowsAtCompileTime, N> rightCols() { return Block<Derived, RowsAtCompileTime, N>(derived(), 0, cols() - N, rows(), N); }
    template <int N> const Block<Derived, RowsAtCompileTime, N> rightCols() const { return Block<Derived, RowsAtCompileTime, N>(const_cast<Derived &>(derived()), 0, cols() - N, rows(), N); }
    template <int BR, int BC> Block<Derived, BR, BC> bottomRightCorner() { return Block<Derived, BR, BC>(derived(), rows() - BR, cols() - BC); }
    template <int BR, int BC> const Block<Derived, BR, BC> bottomRightCorner() const { return Block<Derived, BR, BC>(const_cast<Derived &>(derived()), rows() - BR, cols() - BC); }
    template <int BR, int BC> Block<Derived, BR, BC> topRightCorner() { return Block<Derived, BR, BC>(derived(), 0, cols() - BC); }
    template <int BR, int BC> const Block<Derived, BR, BC> topRightCorner() const { return Block<Derived, BR, BC>(const_cast<Derived &>(derived()), 0, cols() - BC); }
    template <int BR, int BC> Block<Derived, BR, BC> bottomLeftCorner() { return Block<Derived, BR, BC>(derived(), rows() - BR, 0); }
    template <int BR, int BC> const Block<Derived, BR, BC> bottomLeftCorner() const { return Block<Derived, BR, BC>(const_cast<Derived &>(derived()), rows() - BR, 0); }
    template <int BR, int BC> Block<Derived, BR, BC> topLeftCorner() { return Block<Derived, BR, BC>(derived(), 0, 0); }
    template <int BR, int BC> const Block<Derived, BR, BC> topLeftCorner() const { return Block<Derived, BR, BC>(const_cast<Derived &>(derived()), 0, 0); }
    Block<Derived, RowsAtCompileTime, 1> col(int j) { return Block<Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
    const Block<Derived, RowsAtCompileTime, 1> col(int j) const { return Block<Derived, RowsAtCompileTime, 1>(const_cast<Derived &>(derived()), 0, j, rows(), 1); }
    Block<Derived, 1, ColsAtCompileTime> row(int i) { return Block<Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
    const Block<Derived, 1, ColsAtCompileTime> row(int i) const { return Block<Derived, 1, ColsAtCompileTime>(const_cast<Derived &>(derived()), i, 0, 1, cols()); }
    // vector segments (column vectors)
    template <int N> Block<Derived, N, 1> head() { return Block<Derived, N, 1>(derived(), 0, 0); }
    template <int N> const Block<Derived, N, 1> head() const { return Block<Derived, N, 1>(const_cast<Derived &>(derived()), 0, 0); }
    template <int N> Block<Derived, N, 1> tail() { return Block<Derived, N, 1>(derived(), rows() - N, 0); }
    template <int N> const Block<Derived, N, 1> tail() const { return Block<Derived, N, 1>(const_cast<Derived &>(derived()), rows() - N, 0); }
    template <int N> Block<Derived, N, 1> segment(int i) { return Block<Derived, N, 1>(derived(), i, 0); }
    template <int N> const Block<Derived, N, 1> segment(int i) const { return Block<Derived, N, 1>(const_cast<Derived &>(derived()), i, 0); }
    // ---- run-time sized views
    DynBlock<Derived> block(int i, int j, int r, int c) { return DynBlock<Derived>(derived(), i, j, r, c); }
    const DynBlock<Derived> block(int i, int j, int r, int c) const { return DynBlock<Derived>(const_cast<Derived &>(derived()), i, j, r, c); }
    DynBlock<Derived> segment(int i, int n) { return cols() == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
    const DynBlock<Derived> segment(int i, int n) const { return cols() == 1 ? block(i, 0, n, 1) : block(0, i, 1, n); }
    DynBlock<Derived> head(int n) { return segment(0, n); }
    const DynBlock<Derived> head(int n) const { return segment(0, n); }
    DynBlock<Derived> tail(int n) { return segment(size() - n, n); }
    const DynBlock<Derived> tail(int n) const { return segment(size() - n, n); }
    DynBlock<Derived> leftCols(int n) { return block(0, 0, rows(), n); }
    const DynBlock<Derived> leftCols(int n) const { return block(0, 0, rows(), n); }
    DynBlock<Derived> rightCols(int n) { return block(0, cols() - n, rows(), n); }
    const DynBlock<Derived> rightCols(int n) const { return block(0, cols() - n, rows(), n); }
    DynBlock<Derived> middleCols(int j, int n) { return block(0, j, rows(), n); }
    const DynBlock<Derived> middleCols(int j, int n) const { return block(0, j, rows(), n); }
    DynBlock<Derived> topRows(int n) { return block(0, 0, n, cols()); }
    const DynBlock<Derived> topRows(int n) const { return block(0, 0, n, cols()); }
    DynBlock<Derived> bottomRows(int n) { return block(rows() - n, 0, n, cols()); }
    const DynBlock<Derived> bottomRows(int n) const { return block(rows() - n, 0, n, cols()); }
    ArrayX<Scalar> array() const;
    template <typename NewT> typename plain_type<NewT, RowsAtCompileTime, ColsAtCompileTime>::type cast() const {
        typedef typename plain_type<NewT, RowsAtCompileTime, ColsAtCompileTime>::type R; R r = make_plain<R>(rows(), cols());
        for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r.coeffRef(i, j) = NewT(coeff(i, j)); return r;
    }
    JacobiSVD<PlainObject> jacobiSvd(unsigned options = 0) const;
    PlainObject cwiseSqrt() const { PlainObject r = plain(); for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r.coeffRef(i, j) = std::sqrt(coeff(i, j)); return r; }

    // ---- assignment from any expression of the same size
    template <typename O> Derived &assign(const MatrixBase<O> &o) {
        static_assert(((int)traits<O>::Rows == Dynamic || (int)RowsAtCompileTime == Dynamic || (int)traits<O>::Rows == (int)RowsAtCompileTime) &&
                      ((int)traits<O>::Cols == Dynamic || (int)ColsAtCompileTime == Dynamic || (int)traits<O>::Cols == (int)ColsAtCompileTime), "mini_eigen: size mismatch in assignment");
        derived().resizeLike(o.rows(), o.cols());
        PlainObject t = plain(); for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) t.coeffRef(i, j) = o.coeff(i, j);    // via a temporary: aliasing-safe
        for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) = t.coeff(i, j);
        return derived();
    }
    template <typename O> Derived &operator+=(const MatrixBase<O> &o) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) += o.coeff(i, j); return derived(); }
    template <typename O> Derived &operator-=(const MatrixBase<O> &o) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) -= o.coeff(i, j); return derived(); }
    template <typename O> Derived &operator*=(const MatrixBase<O> &o) { PlainObject t = (*this) * o; return assign(t); }
    template <typename O> void swap(MatrixBase<O> &o) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) { const Scalar t = coeff(i, j); coeffRef(i, j) = o.coeff(i, j); o.coeffRef(i, j) = t; } }
    Derived &operator*=(Scalar s) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) *= s; return derived(); }
    Derived &operator/=(Scalar s) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) /= s; return derived(); }
    CommaInit<Derived> operator<<(Scalar first) { return CommaInit<Derived>(derived(), first); }
    Derived &setZero() { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) = Scalar(0); return derived(); }
    Derived &setIdentity() { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) = i == j ? Scalar(1) : Scalar(0); return derived(); }
    Derived &setConstant(Scalar v) { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) coeffRef(i, j) = v; return derived(); }
    Derived &noalias() { return derived(); }

    // ---- reductions
    Scalar squaredNorm() const { Scalar s = 0; for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j); return s; }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    PlainObject normalized() const { PlainObject r = eval(); const Scalar n2 = squaredNorm(); if (n2 > Scalar(0)) r /= std::sqrt(n2); return r; }
    void normalize() { const Scalar n2 = squaredNorm(); if (n2 > Scalar(0)) *this /= std::sqrt(n2); }
    Scalar sum() const { Scalar s = 0; for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) s += coeff(i, j); return s; }
    Scalar trace() const { Scalar s = 0; for (int i = 0; i < rows(); i++) s += coeff(i, i); return s; }
    Scalar maxCoeff() const { Scalar m = coeff(0, 0); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) if (coeff(i, j) > m) m = coeff(i, j); return m; }
    Scalar minCoeff() const { Scalar m = coeff(0, 0); for (int j = 0; j < cols(); j++) for (int i = 0; i < rows(); i++) if (coeff(i, j) < m) m = coeff(i, j); return m; }
    template <typename O> Scalar dot(const MatrixBase<O> &o) const { Scalar s = 0; for (int i = 0; i < size(); i++) s += (*this)(i) * o(i); return s; }
    template <typename O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O> &o) const {
        Matrix<Scalar, 3, 1> r; const MatrixBase &a = *this;
        r(0) = a(1) * o(2) - a(2) * o(1); r(1) = a(2) * o(0) - a(0) * o(2); r(2) = a(0) * o(1) - a(1) * o(0); return r;
    }
    template <typename O> bool operator==(const MatrixBase<O> &o) const { for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) if (coeff(i, j) != o.coeff(i, j)) return false; return true; }
    // vector -> diagonal matrix
    typename plain_type<Scalar, ((int)RowsAtCompileTime == Dynamic || (int)ColsAtCompileTime == Dynamic) ? Dynamic : (int)SizeAtCompileTime,
                        ((int)RowsAtCompileTime == Dynamic || (int)ColsAtCompileTime == Dynamic) ? Dynamic : (int)SizeAtCompileTime>::type asDiagonal() const {
        typedef typename plain_type<Scalar, ((int)RowsAtCompileTime == Dynamic || (int)ColsAtCompileTime == Dynamic) ? Dynamic : (int)SizeAtCompileTime,
                                    ((int)RowsAtCompileTime == Dynamic || (int)ColsAtCompileTime == Dynamic) ? Dynamic : (int)SizeAtCompileTime>::type D;
        D r = make_plain<D>(size(), size()); r.setZero(); for (int i = 0; i < size(); i++) r.coeffRef(i, i) = (*this)(i); return r;
    }
    PlainObject cwiseProduct(const PlainObject &o) const { PlainObject r = plain(); for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r.coeffRef(i, j) = coeff(i, j) * o.coeff(i, j); return r; }
    PlainObject cwiseAbs() const { PlainObject r = plain(); for (int i = 0; i < rows(); i++) for (int j = 0; j < cols(); j++) r.coeffRef(i, j) = std::abs(coeff(i, j)); return r; }
    // general inverse: Gauss-Jordan with partial pivoting (Eigen uses cofactors up to 4x4 and partial-pivot LU above)
    LDLT<PlainObject> ldlt() const;
    PlainObject inverse() const {
        assert(rows() == cols());
        const int n = rows();
        PlainObject a = eval(), inv = plain(); inv.setIdentity();
        for (int c = 0; c < n; c++) {
            int p = c; for (int r = c + 1; r < n; r++) if (std::abs(a.coeff(r, c)) > std::abs(a.coeff(p, c))) p = r;
            if (p != c) for (int k = 0; k < n; k++) { std::swap(a.coeffRef(p, k), a.coeffRef(c, k)); std::swap(inv.coeffRef(p, k), inv.coeffRef(c, k)); }
            const Scalar d = a.coeff(c, c);
            for (int k = 0; k < n; k++) { a.coeffRef(c, k) /= d; inv.coeffRef(c, k) /= d; }
            for (int r = 0; r < n; r++) if (r != c) { const Scalar f = a.coeff(r, c); if (f != Scalar(0)) for (int k = 0; k < n; k++) { a.coeffRef(r, k) -= f * a.coeff(c, k); inv.coeffRef(r, k) -= f * inv.coeff(c, k); } }
        }
        return inv;
    }
    Scalar determinant() const {
        static_assert((int)RowsAtCompileTime == 3 && (int)ColsAtCompileTime == 3, "mini_eigen: determinant only for 3x3");
        return coeff(0, 0) * (coeff(1, 1) * coeff(2, 2) - coeff(1, 2) * coeff(2, 1)) - coeff(0, 1) * (coeff(1, 0) * coeff(2, 2) - coeff(1, 2) * coeff(2, 0)) +
               coeff(0, 2) * (coeff(1, 0) * coeff(2, 1) - coeff(1, 1) * coeff(2, 0));
    }
    static PlainObject Zero() { PlainObject r; r.setZero(); return r; }
    static PlainObject Identity() { PlainObject r; r.setIdentity(); return r; }
    static PlainObject Ones() { PlainObject r; r.setConstant(Scalar(1)); return r; }
    static PlainObject Constant(Scalar v) { PlainObject r; r.setConstant(v); return r; }
    static PlainObject Zero(int n_) { PlainObject r = make_plain<PlainObject>(n_, 1); r.setZero(); return r; }
    static PlainObject Zero(int r_, int c_) { PlainObject r = make_plain<PlainObject>(r_, c_); r.setZero(); return r; }
    static PlainObject Identity(int r_, int c_) { PlainObject r = make_plain<PlainObject>(r_, c_); r.setIdentity(); return r; }
};

// ------------------------------------------------------------------------------------------------ Matrix
template <typename T, int R, int C, int Opt> class Matrix : public MatrixBase<Matrix<T, R, C, Opt>> {
    static_assert(R > 0 && C > 0, "mini_eigen: fixed sizes here; Matrix<T, Dynamic, Dynamic> is specialised below");
    T d[R * C];
  public:
    typedef MatrixBase<Matrix> Base;
    typedef T Scalar;
    Matrix() {}
    explicit Matrix(int) {}                                           // Eigen::Vector3d ypr(3): a size, ignored for fixed sizes
    Matrix(SizeTag, int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
    static int rowsImpl() { return R; }
    static int colsImpl() { return C; }
    static void resizeLike(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
    template <typename A, typename B, typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
    Matrix(A a, B b) { if constexpr (R * C == 2) { d[0] = T(a); d[1] = T(b); } else { assert((int)a == R && (int)b == C); } }   // two coefficients, or (rows, cols) of a fixed-size matrix
    template <typename A, typename B, typename Cc, typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<Cc>::value>::type>
    Matrix(A a, B b, Cc c) { static_assert(R * C == 3, "3 coefficients"); d[0] = T(a); d[1] = T(b); d[2] = T(c); }
    Matrix(T a, T b, T c, T e) { static_assert(R * C == 4, "4 coefficients"); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    Matrix(const Matrix &o) { for (int i = 0; i < R * C; i++) d[i] = o.d[i]; }
    template <typename O> Matrix(const MatrixBase<O> &o) { Base::assign(o); }
    Matrix &operator=(const Matrix &o) { for (int i = 0; i < R * C; i++) d[i] = o.d[i]; return *this; }
    template <typename O> Matrix &operator=(const MatrixBase<O> &o) { return Base::assign(o); }
    T coeff(int i, int j) const { assert(i >= 0 && i < R && j >= 0 && j < C); return Opt == RowMajor ? d[i * C + j] : d[j * R + i]; }
    T &coeffRef(int i, int j) { assert(i >= 0 && i < R && j >= 0 && j < C); return Opt == RowMajor ? d[i * C + j] : d[j * R + i]; }
    T *data() { return d; }
    const T *data() const { return d; }
};

// run-time sized matrix (MatrixXd F = MatrixXd::Zero(15, 15) in integration_base.h; the Jacobian buffers of ResidualBlockInfo are RowMajor)
template <typename T, int Opt> class Matrix<T, Dynamic, Dynamic, Opt> : public MatrixBase<Matrix<T, Dynamic, Dynamic, Opt>> {
    std::vector<T> d; int r_, c_;
  public:
    typedef MatrixBase<Matrix> Base;
    typedef T Scalar;
    Matrix() : r_(0), c_(0) {}
    Matrix(SizeTag, int r, int c) : d((size_t)r * c), r_(r), c_(c) {}
    Matrix(int r, int c) : d((size_t)r * c), r_(r), c_(c) {}
    Matrix(const Matrix &o) : d(o.d), r_(o.r_), c_(o.c_) {}
    template <typename O> Matrix(const MatrixBase<O> &o) : r_(0), c_(0) { Base::assign(o); }
    Matrix &operator=(const Matrix &o) { d = o.d; r_ = o.r_; c_ = o.c_; return *this; }
    template <typename O> Matrix &operator=(const MatrixBase<O> &o) { return Base::assign(o); }
    int rowsImpl() const { return r_; }
    int colsImpl() const { return c_; }
    void resizeLike(int r, int c) { if (r != r_ || c != c_) { d.assign((size_t)r * c, T(0)); r_ = r; c_ = c; } }
    void resize(int r, int c) { resizeLike(r, c); }
    T coeff(int i, int j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return Opt == RowMajor ? d[(size_t)i * c_ + j] : d[(size_t)j * r_ + i]; }
    T &coeffRef(int i, int j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return Opt == RowMajor ? d[(size_t)i * c_ + j] : d[(size_t)j * r_ + i]; }
    T *data() { return d.data(); }
    const T *data() const { return d.data(); }
};
// run-time sized column vector (VectorXd)
template <typename T, int Opt> struct traits<Matrix<T, Dynamic, 1, Opt>> { typedef T Scalar; enum { Rows = Dynamic, Cols = 1 }; };
template <typename T, int Opt> class Matrix<T, Dynamic, 1, Opt> : public MatrixBase<Matrix<T, Dynamic, 1, Opt>> {
    std::vector<T> d;
  public:
    typedef MatrixBase<Matrix> Base;
    typedef T Scalar;
    Matrix() {}
    explicit Matrix(int n) : d((size_t)n) {}
    Matrix(SizeTag, int r, int c) : d((size_t)r) { assert(c == 1); (void)c; }
    Matrix(const Matrix &o) : d(o.d) {}
    template <typename O> Matrix(const MatrixBase<O> &o) { Base::assign(o); }
    Matrix(const ArrayX<T> &a);
    Matrix &operator=(const Matrix &o) { d = o.d; return *this; }
    template <typename O> Matrix &operator=(const MatrixBase<O> &o) { return Base::assign(o); }
    int rowsImpl() const { return (int)d.size(); }
    static int colsImpl() { return 1; }
    void resizeLike(int r, int c) { assert(c == 1); (void)c; if (r != (int)d.size()) d.assign((size_t)r, T(0)); }
    void resize(int n) { resizeLike(n, 1); }
    T coeff(int i, int j) const { assert(i >= 0 && i < (int)d.size() && j == 0); (void)j; return d[i]; }
    T &coeffRef(int i, int j) { assert(i >= 0 && i < (int)d.size() && j == 0); (void)j; return d[i]; }
    T *data() { return d.data(); }
    const T *data() const { return d.data(); }
};
// run-time sized view
template <typename XprType> class DynBlock : public MatrixBase<DynBlock<XprType>> {
    XprType *x; int i0, j0, r_, c_;
  public:
    typedef MatrixBase<DynBlock> Base;
    typedef typename traits<XprType>::Scalar Scalar;
    DynBlock(XprType &x_, int i, int j, int r, int c) : x(&x_), i0(i), j0(j), r_(r), c_(c) { assert(i >= 0 && j >= 0 && r >= 0 && c >= 0 && i + r <= x_.rows() && j + c <= x_.cols()); }
    int rowsImpl() const { return r_; }
    int colsImpl() const { return c_; }
    void resizeLike(int r, int c) const { assert(r == r_ && c == c_); (void)r; (void)c; }
    Scalar coeff(int i, int j) const { return x->coeff(i0 + i, j0 + j); }
    Scalar &coeffRef(int i, int j) const { return x->coeffRef(i0 + i, j0 + j); }
    DynBlock &operator=(const DynBlock &o) { Base::assign(o); return *this; }
    template <typename O> DynBlock &operator=(const MatrixBase<O> &o) { Base::assign(o); return *this; }
    template <typename O> const DynBlock &operator=(const MatrixBase<O> &o) const { const_cast<DynBlock *>(this)->assign(o); return *this; }
};
// coefficient-wise view of a vector: (v.array() > eps).select(v.array().inverse(), 0)
template <typename T> struct BoolArrayX {
    std::vector<char> b;
    ArrayX<T> select(const ArrayX<T> &then, T otherwise) const;
};
template <typename T> class ArrayX {
  public:
    std::vector<T> v;
    ArrayX() {}
    explicit ArrayX(size_t n) : v(n) {}
    BoolArrayX<T> operator>(T s) const { BoolArrayX<T> r; r.b.resize(v.size()); for (size_t i = 0; i < v.size(); i++) r.b[i] = v[i] > s; return r; }
    ArrayX inverse() const { ArrayX r(v.size()); for (size_t i = 0; i < v.size(); i++) r.v[i] = T(1) / v[i]; return r; }
    ArrayX sqrt() const { ArrayX r(v.size()); for (size_t i = 0; i < v.size(); i++) r.v[i] = std::sqrt(v[i]); return r; }
};
template <typename T> ArrayX<T> BoolArrayX<T>::select(const ArrayX<T> &then, T otherwise) const { ArrayX<T> r(b.size()); for (size_t i = 0; i < b.size(); i++) r.v[i] = b[i] ? then.v[i] : otherwise; return r; }
template <typename T, int Opt> Matrix<T, Dynamic, 1, Opt>::Matrix(const ArrayX<T> &a) : d(a.v) {}
template <typename Derived> ArrayX<typename MatrixBase<Derived>::Scalar> MatrixBase<Derived>::array() const { ArrayX<Scalar> r((size_t)size()); for (int i = 0; i < size(); i++) r.v[i] = (*this)(i); return r; }

// ------------------------------------------------------------------------------------------------ Block (a view)
template <typename XprType, int BR, int BC> class Block : public MatrixBase<Block<XprType, BR, BC>> {
    XprType *x; int i0, j0, r_, c_;                                   // r_, c_ only matter where BR / BC is Dynamic (a fixed-count slice of a run-time sized parent)
  public:
    typedef MatrixBase<Block> Base;
    typedef typename traits<XprType>::Scalar Scalar;
    Block(XprType &x_, int i, int j, int r = BR, int c = BC) : x(&x_), i0(i), j0(j), r_(r), c_(c) {
        assert((BR == Dynamic || r == BR) && (BC == Dynamic || c == BC));
        assert(i >= 0 && j >= 0 && r >= 0 && c >= 0 && i + r <= x_.rows() && j + c <= x_.cols());
    }
    int rowsImpl() const { return BR == Dynamic ? r_ : BR; }
    int colsImpl() const { return BC == Dynamic ? c_ : BC; }
    void resizeLike(int r, int c) const { assert(r == rowsImpl() && c == colsImpl()); (void)r; (void)c; }
    Scalar coeff(int i, int j) const { return x->coeff(i0 + i, j0 + j); }
    Scalar &coeffRef(int i, int j) const { return x->coeffRef(i0 + i, j0 + j); }
    Block &operator=(const Block &o) { Base::assign(o); return *this; }
    template <typename O> Block &operator=(const MatrixBase<O> &o) { Base::assign(o); return *this; }
    template <typename O> const Block &operator=(const MatrixBase<O> &o) const { const_cast<Block *>(this)->assign(o); return *this; }
};

// ------------------------------------------------------------------------------------------------ Map of a plain matrix
template <typename T, int R, int C, int Opt> class Map<Matrix<T, R, C, Opt>> : public MatrixBase<Map<Matrix<T, R, C, Opt>>> {
    T *p;
  public:
    typedef MatrixBase<Map> Base;
    typedef T Scalar;
    explicit Map(T *p_) : p(p_) {}
    static int rowsImpl() { return R; }
    static int colsImpl() { return C; }
    static void resizeLike(int r, int c) { assert(r == R && c == C); (void)r; (void)c; }
    T coeff(int i, int j) const { return Opt == RowMajor ? p[i * C + j] : p[j * R + i]; }
    T &coeffRef(int i, int j) const { return Opt == RowMajor ? p[i * C + j] : p[j * R + i]; }
    Map &operator=(const Map &o) { Base::assign(o); return *this; }
    template <typename O> Map &operator=(const MatrixBase<O> &o) { Base::assign(o); return *this; }
};
template <typename T, int R, int C, int Opt> class Map<const Matrix<T, R, C, Opt>> : public MatrixBase<Map<const Matrix<T, R, C, Opt>>> {
    const T *p;
  public:
    typedef T Scalar;
    explicit Map(const T *p_) : p(p_) {}
    static int rowsImpl() { return R; }
    static int colsImpl() { return C; }
    static void resizeLike(int, int) {}
    T coeff(int i, int j) const { return Opt == RowMajor ? p[i * C + j] : p[j * R + i]; }
    T &coeffRef(int i, int j) const { return const_cast<T *>(p)[Opt == RowMajor ? i * C + j : j * R + i]; }
};

// run-time sized maps: Map<VectorXd>(p, n), Map<const VectorXd>(p, n), Map<Matrix<T, Dynamic, Dynamic, RowMajor>>(p, r, c)
template <typename T, int Opt> class Map<Matrix<T, Dynamic, 1, Opt>> : public MatrixBase<Map<Matrix<T, Dynamic, 1, Opt>>> {
    T *p; int n_;
  public:
    typedef MatrixBase<Map> Base;
    typedef T Scalar;
    Map(T *p_, int n) : p(p_), n_(n) {}
    int rowsImpl() const { return n_; }
    static int colsImpl() { return 1; }
    void resizeLike(int r, int c) const { assert(r == n_ && c == 1); (void)r; (void)c; }
    T coeff(int i, int) const { return p[i]; }
    T &coeffRef(int i, int) const { return p[i]; }
    Map &operator=(const Map &o) { Base::assign(o); return *this; }
    template <typename O> Map &operator=(const MatrixBase<O> &o) { Base::assign(o); return *this; }
};
template <typename T, int Opt> class Map<const Matrix<T, Dynamic, 1, Opt>> : public MatrixBase<Map<const Matrix<T, Dynamic, 1, Opt>>> {
    const T *p; int n_;
  public:
    typedef T Scalar;
    Map(const T *p_, int n) : p(p_), n_(n) {}
    int rowsImpl() const { return n_; }
    static int colsImpl() { return 1; }
    void resizeLike(int, int) const {}
    T coeff(int i, int) const { return p[i]; }
    T &coeffRef(int i, int) const { return const_cast<T *>(p)[i]; }
};
template <typename T, int Opt> class Map<Matrix<T, Dynamic, Dynamic, Opt>> : public MatrixBase<Map<Matrix<T, Dynamic, Dynamic, Opt>>> {
    T *p; int r_, c_;
  public:
    typedef MatrixBase<Map> Base;
    typedef T Scalar;
    Map(T *p_, int r, int c) : p(p_), r_(r), c_(c) {}
    int rowsImpl() const { return r_; }
    int colsImpl() const { return c_; }
    void resizeLike(int r, int c) const { assert(r == r_ && c == c_); (void)r; (void)c; }
    T coeff(int i, int j) const { return Opt == RowMajor ? p[(size_t)i * c_ + j] : p[(size_t)j * r_ + i]; }
    T &coeffRef(int i, int j) const { return Opt == RowMajor ? p[(size_t)i * c_ + j] : p[(size_t)j * r_ + i]; }
    Map &operator=(const Map &o) { Base::assign(o); return *this; }
    template <typename O> Map &operator=(const MatrixBase<O> &o) { Base::assign(o); return *this; }
};

// ------------------------------------------------------------------------------------------------ arithmetic (eager)
#define MINI_EIGEN_SAME(A, B) static_assert(((int)traits<A>::Rows == Dynamic || (int)traits<B>::Rows == Dynamic || (int)traits<A>::Rows == (int)traits<B>::Rows) && \
    ((int)traits<A>::Cols == Dynamic || (int)traits<B>::Cols == Dynamic || (int)traits<A>::Cols == (int)traits<B>::Cols), "mini_eigen: size mismatch")
template <typename A, typename B> struct sum_type { typedef typename plain_type<typename traits<A>::Scalar, pick_dim(traits<A>::Rows, traits<B>::Rows), pick_dim(traits<A>::Cols, traits<B>::Cols)>::type type; };
template <typename A, typename B> typename sum_type<A, B>::type operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) {
    MINI_EIGEN_SAME(A, B); assert(a.rows() == b.rows() && a.cols() == b.cols());
    typename sum_type<A, B>::type r = make_plain<typename sum_type<A, B>::type>(a.rows(), a.cols());
    for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r.coeffRef(i, j) = a.coeff(i, j) + b.coeff(i, j); return r;
}
template <typename A, typename B> typename sum_type<A, B>::type operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) {
    MINI_EIGEN_SAME(A, B); assert(a.rows() == b.rows() && a.cols() == b.cols());
    typename sum_type<A, B>::type r = make_plain<typename sum_type<A, B>::type>(a.rows(), a.cols());
    for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r.coeffRef(i, j) = a.coeff(i, j) - b.coeff(i, j); return r;
}
template <typename A> typename A::PlainObject operator-(const MatrixBase<A> &a) {
    typename A::PlainObject r = a.plain(); for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r.coeffRef(i, j) = -a.coeff(i, j); return r;
}
template <typename A, typename B> struct prod_type { typedef typename plain_type<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols>::type type; };
template <typename A, typename B> typename prod_type<A, B>::type operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
    static_assert((int)traits<A>::Cols == Dynamic || (int)traits<B>::Rows == Dynamic || (int)traits<A>::Cols == (int)traits<B>::Rows, "mini_eigen: inner sizes differ in *");
    assert(a.cols() == b.rows());
    typename prod_type<A, B>::type r = make_plain<typename prod_type<A, B>::type>(a.rows(), b.cols());
    for (int i = 0; i < a.rows(); i++) for (int j = 0; j < b.cols(); j++) { typename traits<A>::Scalar s = a.coeff(i, 0) * b.coeff(0, j); for (int k = 1; k < a.cols(); k++) s += a.coeff(i, k) * b.coeff(k, j); r.coeffRef(i, j) = s; }
    return r;
}
template <typename A, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> typename A::PlainObject operator*(const MatrixBase<A> &a, S s) {
    typename A::PlainObject r = a.plain(); for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r.coeffRef(i, j) = a.coeff(i, j) * typename traits<A>::Scalar(s); return r;
}
template <typename A, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> typename A::PlainObject operator*(S s, const MatrixBase<A> &a) {
    typename A::PlainObject r = a.plain(); for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r.coeffRef(i, j) = typename traits<A>::Scalar(s) * a.coeff(i, j); return r;
}
template <typename A, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type> typename A::PlainObject operator/(const MatrixBase<A> &a, S s) {
    typename A::PlainObject r = a.plain(); for (int i = 0; i < a.rows(); i++) for (int j = 0; j < a.cols(); j++) r.coeffRef(i, j) = a.coeff(i, j) / typename traits<A>::Scalar(s); return r;
}
// a 1 x 1 product used as a scalar: var += v.transpose() * v
template <typename S, typename A> typename std::enable_if<std::is_arithmetic<S>::value && (int)traits<A>::Rows == 1 && (int)traits<A>::Cols == 1, S &>::type operator+=(S &s, const MatrixBase<A> &a) { s += a.coeff(0, 0); return s; }
template <typename A> std::ostream &operator<<(std::ostream &os, const MatrixBase<A> &a) {
    for (int i = 0; i < a.rows(); i++) { for (int j = 0; j < a.cols(); j++) os << (j ? " " : "") << a.coeff(i, j); if (i + 1 < a.rows()) os << "\n"; }
    return os;
}

// ------------------------------------------------------------------------------------------------ Cholesky (LLT), lower factor
template <typename MatrixType> class LLT {
    typename MatrixType::PlainObject L; bool ok;
  public:
    template <typename O> explicit LLT(const MatrixBase<O> &a) : ok(true) {
        const int n = a.rows(); L = make_plain<typename MatrixType::PlainObject>(n, n); L.setZero();
        for (int j = 0; j < n; j++) {
            typename MatrixType::Scalar s = a.coeff(j, j); for (int k = 0; k < j; k++) s -= L.coeff(j, k) * L.coeff(j, k);
            if (!(s > 0)) ok = false;
            const typename MatrixType::Scalar ljj = std::sqrt(s); L.coeffRef(j, j) = ljj;
            for (int i = j + 1; i < n; i++) { typename MatrixType::Scalar t = a.coeff(i, j); for (int k = 0; k < j; k++) t -= L.coeff(i, k) * L.coeff(j, k); L.coeffRef(i, j) = t / ljj; }
        }
    }
    const typename MatrixType::PlainObject &matrixL() const { return L; }
    typename MatrixType::PlainObject matrixU() const { return L.transpose(); }
    bool success() const { return ok; }
};

// ------------------------------------------------------------------------------------------------ LDL^T with diagonal pivoting
// (the pivot rule of Eigen's LDLT: the largest remaining |diagonal|; solve() applies the pseudo-inverse of D)
template <typename MatrixType> class LDLT {
    typedef typename traits<MatrixType>::Scalar T;
    std::vector<T> a; std::vector<int> perm; int n;
  public:
    template <typename O> explicit LDLT(const MatrixBase<O> &m) : n(m.rows()) {
        a.resize((size_t)n * n); perm.resize(n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[(size_t)i * n + j] = i >= j ? m.coeff(i, j) : m.coeff(j, i);
        auto A = [&](int i, int j) -> T & { return a[(size_t)i * n + j]; };
        for (int k = 0; k < n; k++) {
            int p = k; T best = std::fabs(A(k, k));
            for (int i = k + 1; i < n; i++) if (std::fabs(A(i, i)) > best) { best = std::fabs(A(i, i)); p = i; }
            perm[k] = p;
            if (p != k) {
                for (int j = 0; j < k; j++) std::swap(A(k, j), A(p, j));
                for (int i = p + 1; i < n; i++) std::swap(A(i, k), A(i, p));
                for (int i = k + 1; i < p; i++) std::swap(A(i, k), A(p, i));
                std::swap(A(k, k), A(p, p));
            }
            const T d = A(k, k), inv = std::fabs(d) > std::numeric_limits<T>::min() ? T(1) / d : T(0);
            for (int i = k + 1; i < n; i++) { const T lik = A(i, k) * inv; for (int j = k + 1; j <= i; j++) A(i, j) -= lik * A(j, k); }
            for (int i = k + 1; i < n; i++) A(i, k) *= inv;
        }
    }
    template <typename O> typename O::PlainObject solve(const MatrixBase<O> &b) const {
        typename O::PlainObject x = b.eval();
        for (int k = 0; k < n; k++) if (perm[k] != k) std::swap(x.coeffRef(k, 0), x.coeffRef(perm[k], 0));
        for (int i = 0; i < n; i++) { T s = x.coeff(i, 0); for (int j = 0; j < i; j++) s -= a[(size_t)i * n + j] * x.coeff(j, 0); x.coeffRef(i, 0) = s; }
        for (int i = 0; i < n; i++) { const T d = a[(size_t)i * n + i]; x.coeffRef(i, 0) = std::fabs(d) > std::numeric_limits<T>::min() ? x.coeff(i, 0) / d : T(0); }
        for (int i = n - 1; i >= 0; i--) { T s = x.coeff(i, 0); for (int j = i + 1; j < n; j++) s -= a[(size_t)j * n + i] * x.coeff(j, 0); x.coeffRef(i, 0) = s; }
        for (int k = n - 1; k >= 0; k--) if (perm[k] != k) std::swap(x.coeffRef(k, 0), x.coeffRef(perm[k], 0));
        return x;
    }
};
template <typename Derived> LDLT<typename MatrixBase<Derived>::PlainObject> MatrixBase<Derived>::ldlt() const { return LDLT<PlainObject>(*this); }

// ------------------------------------------------------------------------------------------------ symmetric eigen decomposition
// Cyclic Jacobi rotations (Eigen tridiagonalises and runs implicit QR; eigenvalues agree to round-off, eigenvectors up to sign and, for
// repeated eigenvalues, basis).  eigenvalues() ascending, eigenvectors() in the matching columns.
template <typename MatrixType> class SelfAdjointEigenSolver {
    typedef typename traits<MatrixType>::Scalar T;
    Matrix<T, Dynamic, 1> w; Matrix<T, Dynamic, Dynamic> V;
  public:
    template <typename O> explicit SelfAdjointEigenSolver(const MatrixBase<O> &m) {
        const int n = m.rows(); assert(m.cols() == n);
        std::vector<T> a((size_t)n * n), v((size_t)n * n, T(0));
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[(size_t)i * n + j] = i >= j ? m.coeff(i, j) : m.coeff(j, i);       // lower triangle is read
        for (int i = 0; i < n; i++) v[(size_t)i * n + i] = T(1);
        for (int sweep = 0; sweep < 60; sweep++) {
            T off = 0, diag = 0;
            for (int i = 0; i < n; i++) { diag += a[(size_t)i * n + i] * a[(size_t)i * n + i]; for (int j = 0; j < i; j++) off += a[(size_t)i * n + j] * a[(size_t)i * n + j]; }
            if (off <= T(1e-32) * (diag + off) || off == T(0)) break;
            for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
                const T apq = a[(size_t)p * n + q];
                if (apq == T(0)) continue;
                const T theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (T(2) * apq);
                const T t = (theta >= 0 ? T(1) : T(-1)) / (std::abs(theta) + std::sqrt(theta * theta + T(1)));
                const T c = T(1) / std::sqrt(t * t + T(1)), sn = t * c;
                for (int k = 0; k < n; k++) { const T akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q]; a[(size_t)k * n + p] = c * akp - sn * akq; a[(size_t)k * n + q] = sn * akp + c * akq; }
                for (int k = 0; k < n; k++) { const T apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k]; a[(size_t)p * n + k] = c * apk - sn * aqk; a[(size_t)q * n + k] = sn * apk + c * aqk; }
                for (int k = 0; k < n; k++) { const T vkp = v[(size_t)k * n + p], vkq = v[(size_t)k * n + q]; v[(size_t)k * n + p] = c * vkp - sn * vkq; v[(size_t)k * n + q] = sn * vkp + c * vkq; }
            }
        }
        std::vector<int> order(n); for (int i = 0; i < n; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return a[(size_t)x * n + x] < a[(size_t)y * n + y]; });
        w.resize(n); V.resize(n, n);
        for (int k = 0; k < n; k++) { w.coeffRef(k, 0) = a[(size_t)order[k] * n + order[k]]; for (int i = 0; i < n; i++) V.coeffRef(i, k) = v[(size_t)i * n + order[k]]; }
    }
    const Matrix<T, Dynamic, 1> &eigenvalues() const { return w; }
    const Matrix<T, Dynamic, Dynamic> &eigenvectors() const { return V; }
};

// ------------------------------------------------------------------------------------------------ singular value decomposition
// One-sided (Hestenes) Jacobi on the columns: singular values descending, right singular vectors in the matching columns of matrixV().
// (Eigen's JacobiSVD is two-sided with a QR preconditioner; the factors agree up to the sign of each singular vector pair.)
template <typename MatrixType> class JacobiSVD {
    typedef typename traits<MatrixType>::Scalar T;
    typedef typename plain_type<T, traits<MatrixType>::Cols, traits<MatrixType>::Cols>::type VType;
    typedef typename plain_type<T, traits<MatrixType>::Cols, 1>::type SType;
    VType V_; SType s_;
  public:
    template <typename O> explicit JacobiSVD(const MatrixBase<O> &A, unsigned = 0) {
        const int m = A.rows(), n = A.cols();
        std::vector<T> u((size_t)m * n), v((size_t)n * n, T(0));
        for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) u[(size_t)j * m + i] = A.coeff(i, j);                 // column-major
        for (int j = 0; j < n; j++) v[(size_t)j * n + j] = T(1);
        for (int sweep = 0; sweep < 60; sweep++) {
            bool rotated = false;
            for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
                T alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < m; i++) { const T a = u[(size_t)p * m + i], b = u[(size_t)q * m + i]; alpha += a * a; beta += b * b; gamma += a * b; }
                if (gamma == T(0) || std::abs(gamma) <= T(1e-16) * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const T zeta = (beta - alpha) / (T(2) * gamma);
                const T t = (zeta >= 0 ? T(1) : T(-1)) / (std::abs(zeta) + std::sqrt(T(1) + zeta * zeta));
                const T c = T(1) / std::sqrt(T(1) + t * t), sn = c * t;
                for (int i = 0; i < m; i++) { const T a = u[(size_t)p * m + i], b = u[(size_t)q * m + i]; u[(size_t)p * m + i] = c * a - sn * b; u[(size_t)q * m + i] = sn * a + c * b; }
                for (int i = 0; i < n; i++) { const T a = v[(size_t)p * n + i], b = v[(size_t)q * n + i]; v[(size_t)p * n + i] = c * a - sn * b; v[(size_t)q * n + i] = sn * a + c * b; }
            }
            if (!rotated) break;
        }
        std::vector<T> sv(n); std::vector<int> order(n);
        for (int j = 0; j < n; j++) { T a = 0; for (int i = 0; i < m; i++) a += u[(size_t)j * m + i] * u[(size_t)j * m + i]; sv[j] = std::sqrt(a); order[j] = j; }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
        V_ = make_plain<VType>(n, n); s_ = make_plain<SType>(n, 1);
        for (int k = 0; k < n; k++) { s_.coeffRef(k, 0) = sv[order[k]]; for (int i = 0; i < n; i++) V_.coeffRef(i, k) = v[(size_t)order[k] * n + i]; }
    }
    const VType &matrixV() const { return V_; }
    typename plain_type<T, Dynamic, Dynamic>::type matrixU() const { assert(!"mini_eigen: JacobiSVD::matrixU is not provided"); return typename plain_type<T, Dynamic, Dynamic>::type(); }
    const SType &singularValues() const { return s_; }
};
template <typename Derived> JacobiSVD<typename MatrixBase<Derived>::PlainObject> MatrixBase<Derived>::jacobiSvd(unsigned options) const { return JacobiSVD<PlainObject>(*this, options); }

// ------------------------------------------------------------------------------------------------ quaternions
template <typename Derived> struct qtraits;
template <typename QDerived> class QuatVec;
template <typename QDerived> struct traits<QuatVec<QDerived>> { typedef typename qtraits<QDerived>::Scalar Scalar; enum { Rows = 3, Cols = 1 }; };
template <typename QDerived> class QuatVec : public MatrixBase<QuatVec<QDerived>> {
    QDerived *q;
  public:
    typedef typename qtraits<QDerived>::Scalar Scalar;
    explicit QuatVec(QDerived *q_) : q(q_) {}
    static int rowsImpl() { return 3; }
    static int colsImpl() { return 1; }
    static void resizeLike(int, int) {}
    Scalar coeff(int i, int) const { return q->c(i); }
    Scalar &coeffRef(int i, int) const { return q->c(i); }
    template <typename O> QuatVec &operator=(const MatrixBase<O> &o) { for (int i = 0; i < 3; i++) q->c(i) = o(i); return *this; }
};
template <typename Derived> class QuaternionBase {
  public:
    typedef typename qtraits<Derived>::Scalar Scalar;
    typedef Matrix<Scalar, 3, 1> Vector3;
    typedef Matrix<Scalar, 3, 3> Matrix3;
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    Derived &derived() { return *static_cast<Derived *>(this); }
    // storage order x, y, z, w
    Scalar x() const { return derived().c(0); } Scalar y() const { return derived().c(1); } Scalar z() const { return derived().c(2); } Scalar w() const { return derived().c(3); }
    Scalar &x() { return derived().c(0); } Scalar &y() { return derived().c(1); } Scalar &z() { return derived().c(2); } Scalar &w() { return derived().c(3); }
    typedef QuatVec<Derived> VecView;                             // q.vec(): the imaginary part, readable and assignable
    VecView vec() { return VecView(&derived()); }
    const VecView vec() const { return VecView(const_cast<Derived *>(&derived())); }
    Matrix<Scalar, 4, 1> coeffs() const { Matrix<Scalar, 4, 1> r; for (int i = 0; i < 4; i++) r(i) = derived().c(i); return r; }
    Scalar squaredNorm() const { return x() * x() + y() * y() + z() * z() + w() * w(); }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const Scalar n = norm(); for (int i = 0; i < 4; i++) derived().c(i) /= n; }
    Quaternion<Scalar> normalized() const { const Scalar n = norm(); return Quaternion<Scalar>(w() / n, x() / n, y() / n, z() / n); }
    Quaternion<Scalar> conjugate() const { return Quaternion<Scalar>(w(), -x(), -y(), -z()); }
    Quaternion<Scalar> inverse() const {                           // conjugate / squared norm
        const Scalar n2 = squaredNorm();
        if (n2 > Scalar(0)) return Quaternion<Scalar>(w() / n2, -x() / n2, -y() / n2, -z() / n2);
        return Quaternion<Scalar>(0, 0, 0, 0);
    }
    template <typename O> Quaternion<Scalar> operator*(const QuaternionBase<O> &b) const {
        const QuaternionBase &a = *this;
        return Quaternion<Scalar>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                                  a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                  a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                                  a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    // rotation of a vector: v + w * 2(u x v) + u x 2(u x v)
    template <typename O> typename std::enable_if<(int)traits<O>::Rows == 3 && (int)traits<O>::Cols == 3, Matrix3>::type operator*(const MatrixBase<O> &m) const {
        return toRotationMatrix() * m;                              // a rotation times a 3x3 matrix: the rotation matrix times it (RotationBase)
    }
    template <typename O> typename std::enable_if<(int)traits<O>::Rows == 3 && (int)traits<O>::Cols == 1, Vector3>::type operator*(const MatrixBase<O> &v) const {
        Vector3 u; u(0) = x(); u(1) = y(); u(2) = z();
        Vector3 vv = v.eval();
        Vector3 uv = u.cross(vv); uv += uv;
        return vv + w() * uv + u.cross(uv);
    }
    Matrix3 toRotationMatrix() const {
        Matrix3 r;
        const Scalar tx = Scalar(2) * x(), ty = Scalar(2) * y(), tz = Scalar(2) * z();
        const Scalar twx = tx * w(), twy = ty * w(), twz = tz * w(), txx = tx * x(), txy = ty * x(), txz = tz * x(), tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        r(0, 0) = Scalar(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = Scalar(1) - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = Scalar(1) - (txx + tyy);
        return r;
    }
    Matrix3 matrix() const { return toRotationMatrix(); }
    template <typename O> Scalar dot(const QuaternionBase<O> &o) const { return x() * o.x() + y() * o.y() + z() * o.z() + w() * o.w(); }
    template <typename O> Scalar angularDistance(const QuaternionBase<O> &o) const { const Quaternion<Scalar> d = (*this) * o.conjugate(); Scalar n = std::sqrt(d.x() * d.x() + d.y() * d.y() + d.z() * d.z()); return Scalar(2) * std::atan2(n, std::abs(d.w())); }
};

template <typename T> struct qtraits<Quaternion<T>> { typedef T Scalar; };
template <typename T> class Quaternion : public QuaternionBase<Quaternion<T>> {
    T q[4];
  public:
    typedef T Scalar;
    Quaternion() {}
    Quaternion(T w, T x, T y, T z) { q[0] = x; q[1] = y; q[2] = z; q[3] = w; }
    explicit Quaternion(const T *p) { for (int i = 0; i < 4; i++) q[i] = p[i]; }
    template <typename O> Quaternion(const QuaternionBase<O> &o) { q[0] = o.x(); q[1] = o.y(); q[2] = o.z(); q[3] = o.w(); }
    template <typename O> Quaternion &operator=(const QuaternionBase<O> &o) { const T a = o.x(), b = o.y(), c_ = o.z(), d = o.w(); q[0] = a; q[1] = b; q[2] = c_; q[3] = d; return *this; }
    // rotation matrix -> quaternion (Shepperd's branches on the trace)
    template <typename O, typename = typename std::enable_if<(int)traits<O>::Rows == 3 && (int)traits<O>::Cols == 3>::type> explicit Quaternion(const MatrixBase<O> &m) { *this = m; }
    template <typename O> typename std::enable_if<(int)traits<O>::Rows == 3 && (int)traits<O>::Cols == 3, Quaternion &>::type operator=(const MatrixBase<O> &m) {
        T t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
        if (t > T(0)) {
            t = std::sqrt(t + T(1.0)); q[3] = T(0.5) * t; t = T(0.5) / t;
            q[0] = (m.coeff(2, 1) - m.coeff(1, 2)) * t; q[1] = (m.coeff(0, 2) - m.coeff(2, 0)) * t; q[2] = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
        } else {
            int i = 0; if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1; if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + T(1.0)); q[i] = T(0.5) * t; t = T(0.5) / t;
            q[3] = (m.coeff(k, j) - m.coeff(j, k)) * t; q[j] = (m.coeff(j, i) + m.coeff(i, j)) * t; q[k] = (m.coeff(k, i) + m.coeff(i, k)) * t;
        }
        return *this;
    }
    T c(int i) const { return q[i]; }
    T &c(int i) { return q[i]; }
    static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
    // the rotation that takes the direction of a onto the direction of b (Utility::g2R); the antiparallel case picks some orthogonal axis
    template <typename A, typename B> static Quaternion FromTwoVectors(const MatrixBase<A> &a, const MatrixBase<B> &b) {
        const Matrix<T, 3, 1> v0 = a.normalized(), v1 = b.normalized();
        const T c = v1.dot(v0);
        if (c < T(-1) + T(1e-12)) { Matrix<T, 3, 1> ax = v0.cross(Matrix<T, 3, 1>(T(1), T(0), T(0))); if (ax.norm() < T(1e-6)) ax = v0.cross(Matrix<T, 3, 1>(T(0), T(1), T(0))); ax.normalize(); return Quaternion(T(0), ax.x(), ax.y(), ax.z()); }
        const Matrix<T, 3, 1> axis = v0.cross(v1);
        const T s = std::sqrt((T(1) + c) * T(2)), invs = T(1) / s;
        return Quaternion(s * T(0.5), axis.x() * invs, axis.y() * invs, axis.z() * invs);
    }
    Quaternion &setIdentity() { q[0] = q[1] = q[2] = T(0); q[3] = T(1); return *this; }
};
template <typename T> struct qtraits<Map<Quaternion<T>>> { typedef T Scalar; };
template <typename T> struct qtraits<Map<const Quaternion<T>>> { typedef T Scalar; };
template <typename T> class Map<Quaternion<T>> : public QuaternionBase<Map<Quaternion<T>>> {
    T *p;
  public:
    typedef T Scalar;
    explicit Map(T *p_) : p(p_) {}
    T c(int i) const { return p[i]; }
    T &c(int i) { return p[i]; }
    template <typename O> Map &operator=(const QuaternionBase<O> &o) { const T a = o.x(), b = o.y(), c_ = o.z(), d = o.w(); p[0] = a; p[1] = b; p[2] = c_; p[3] = d; return *this; }
    Map &operator=(const Map &o) { for (int i = 0; i < 4; i++) p[i] = o.p[i]; return *this; }
};
template <typename T> class Map<const Quaternion<T>> : public QuaternionBase<Map<const Quaternion<T>>> {
    const T *p;
  public:
    typedef T Scalar;
    explicit Map(const T *p_) : p(p_) {}
    T c(int i) const { return p[i]; }
    T &c(int i) { return const_cast<T *>(p)[i]; }
};

template <typename T> class AngleAxis {
    T a; Matrix<T, 3, 1> ax;
  public:
    template <typename O> AngleAxis(T angle, const MatrixBase<O> &axis) : a(angle), ax(axis) {}
    template <typename O, typename = typename std::enable_if<(int)traits<O>::Rows == 3 && (int)traits<O>::Cols == 3>::type> explicit AngleAxis(const MatrixBase<O> &R) {
        const Quaternion<T> q(R); T n = std::sqrt(q.x() * q.x() + q.y() * q.y() + q.z() * q.z());
        if (n != T(0)) { a = T(2) * std::atan2(n, std::abs(q.w())); if (q.w() < T(0)) n = -n; ax = Matrix<T, 3, 1>(q.x() / n, q.y() / n, q.z() / n); }
        else { a = T(0); ax = Matrix<T, 3, 1>(T(1), T(0), T(0)); }
    }
    T angle() const { return a; }
    const Matrix<T, 3, 1> &axis() const { return ax; }
    Matrix<T, 3, 3> toRotationMatrix() const {
        Matrix<T, 3, 3> r; const T s = std::sin(a), c = std::cos(a);
        const Matrix<T, 3, 1> cc = ax * (T(1) - c);
        T tmp;
        tmp = cc.x() * ax.y(); r(0, 1) = tmp - s * ax.z(); r(1, 0) = tmp + s * ax.z();
        tmp = cc.x() * ax.z(); r(0, 2) = tmp + s * ax.y(); r(2, 0) = tmp - s * ax.y();
        tmp = cc.y() * ax.z(); r(1, 2) = tmp - s * ax.x(); r(2, 1) = tmp + s * ax.x();
        r(0, 0) = cc.x() * ax.x() + c; r(1, 1) = cc.y() * ax.y() + c; r(2, 2) = cc.z() * ax.z() + c;
        return r;
    }
};

typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d; typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, 3, 1> Vector3f; typedef Matrix<float, 3, 3> Matrix3f;
typedef Quaternion<double> Quaterniond; typedef Quaternion<float> Quaternionf;
typedef AngleAxis<double> AngleAxisd;

}  // namespace Eigen
