// arcanefem_b200/FemTypes.h -- the small fixed-size value types of ArcaneFEM's femutils, for user code that keeps
// using them next to the B200 assembly path (element lambdas written against them keep compiling; custom kernels
// over the device views can use them on the device).
//
// Names, template parameters and method meaning follow (paths relative to the ArcaneFEM source root):
//   Real4           femutils/FemUtils.h:35-76
//   RealMatrix<N,M> femutils/FemUtils.h:89-250   (+ operator^, matrixAddition, matrixMultiplication,
//                                                  matrixTranspose, massMatrix: :258-358)
//   RealVector<N>   femutils/FemUtils.h:362-545  (+ vector * matrix, operator^, massMatrix: :552-598)
// Written from the interface; storage is a plain row-major array in every type.  Header-only, no dependency on the
// C ABI.  Under nvcc every member is __host__ __device__.
#pragma once

#include <cstddef>
#include <cstdint>
#include <initializer_list>

#if defined(__CUDACC__)
#define AFB_HOST_DEVICE __host__ __device__
#else
#define AFB_HOST_DEVICE
#endif

namespace arcanefem_b200 {

using Real = double;

/*---------------------------------------------------------------------------*/
//! Four reals (the per-node values of a tetrahedron, e.g. dPhi/dx of its four shape functions).
struct Real4 {
  Real data[4];

  AFB_HOST_DEVICE Real& operator[](std::size_t i) { return data[i]; }
  AFB_HOST_DEVICE const Real& operator[](std::size_t i) const { return data[i]; }
  AFB_HOST_DEVICE Real4 operator+(const Real4& o) const { return { { data[0] + o.data[0], data[1] + o.data[1], data[2] + o.data[2], data[3] + o.data[3] } }; }
  AFB_HOST_DEVICE Real4 operator-(const Real4& o) const { return { { data[0] - o.data[0], data[1] - o.data[1], data[2] - o.data[2], data[3] - o.data[3] } }; }
  AFB_HOST_DEVICE Real4 operator*(Real s) const { return { { data[0] * s, data[1] * s, data[2] * s, data[3] * s } }; }
  friend AFB_HOST_DEVICE Real4 operator*(Real s, const Real4& v) { return v * s; }
};

/*---------------------------------------------------------------------------*/
//! Dense N x M matrix of reals, row-major.
template <int N, int M>
class RealMatrix {
 public:
  static constexpr int totalNbElement() { return N * M; }

  AFB_HOST_DEVICE RealMatrix()
  {
    for (int k = 0; k < N * M; ++k) m_v[k] = 0.0;
  }
  //! N*M values in row-major order (missing ones stay zero)
  AFB_HOST_DEVICE RealMatrix(std::initializer_list<Real> flat)
  {
    int k = 0;
    for (Real x : flat)
      if (k < N * M) m_v[k++] = x;
    for (; k < N * M; ++k) m_v[k] = 0.0;
  }
  //! one inner list per row
  AFB_HOST_DEVICE RealMatrix(std::initializer_list<std::initializer_list<Real>> rows)
  {
    for (int k = 0; k < N * M; ++k) m_v[k] = 0.0;
    int i = 0;
    for (const auto& row : rows) {
      if (i >= N) break;
      int j = 0;
      for (Real x : row)
        if (j < M) m_v[i * M + j++] = x;
      ++i;
    }
  }

  AFB_HOST_DEVICE void fill(Real value)
  {
    for (int k = 0; k < N * M; ++k) m_v[k] = value;
  }
  AFB_HOST_DEVICE Real& operator()(std::int32_t i, std::int32_t j) { return m_v[i * M + j]; }
  AFB_HOST_DEVICE Real operator()(std::int32_t i, std::int32_t j) const { return m_v[i * M + j]; }
  AFB_HOST_DEVICE void multInPlace(Real s)
  {
    for (int k = 0; k < N * M; ++k) m_v[k] *= s;
  }
  AFB_HOST_DEVICE RealMatrix operator+(const RealMatrix& o) const
  {
    RealMatrix r(*this);
    r += o;
    return r;
  }
  AFB_HOST_DEVICE RealMatrix& operator+=(const RealMatrix& o)
  {
    for (int k = 0; k < N * M; ++k) m_v[k] += o.m_v[k];
    return *this;
  }
  AFB_HOST_DEVICE RealMatrix operator-(const RealMatrix& o) const
  {
    RealMatrix r(*this);
    r -= o;
    return r;
  }
  AFB_HOST_DEVICE RealMatrix& operator-=(const RealMatrix& o)
  {
    for (int k = 0; k < N * M; ++k) m_v[k] -= o.m_v[k];
    return *this;
  }
  AFB_HOST_DEVICE RealMatrix operator-() const
  {
    RealMatrix r;
    for (int k = 0; k < N * M; ++k) r.m_v[k] = -m_v[k];
    return r;
  }
  AFB_HOST_DEVICE RealMatrix operator*(Real s) const
  {
    RealMatrix r(*this);
    r.multInPlace(s);
    return r;
  }
  AFB_HOST_DEVICE RealMatrix operator/(Real s) const
  {
    RealMatrix r;
    for (int k = 0; k < N * M; ++k) r.m_v[k] = m_v[k] / s;
    return r;
  }
  friend AFB_HOST_DEVICE RealMatrix operator*(Real s, const RealMatrix& a) { return a * s; }

 private:
  Real m_v[N * M];
};

//! a (N x M) + transpose-shaped b (M x N), element (i,j) = a(i,j) + b(i,j) over the leading N x N part -- the reference
//! uses it with N == M (femutils/FemUtils.h:286-300)
template <int N, int M>
AFB_HOST_DEVICE inline RealMatrix<N, N> matrixAddition(const RealMatrix<N, M>& a, const RealMatrix<M, N>& b)
{
  RealMatrix<N, N> r;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) r(i, j) = a(i, j) + b(i, j);
  return r;
}

//! (N x M) * (M x N) -> N x N
template <int N, int M>
AFB_HOST_DEVICE inline RealMatrix<N, N> matrixMultiplication(const RealMatrix<N, M>& a, const RealMatrix<M, N>& b)
{
  RealMatrix<N, N> r;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      Real s = 0.0;
      for (int k = 0; k < M; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}

template <int N, int M>
AFB_HOST_DEVICE inline RealMatrix<M, N> matrixTranspose(const RealMatrix<N, M>& a)
{
  RealMatrix<M, N> r;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < M; ++j) r(j, i) = a(i, j);
  return r;
}

//! outer product of the two rows, with the diagonal doubled: the P1 mass matrix pattern (1 + delta_ij) phi_i phi_j
template <int N>
AFB_HOST_DEVICE inline RealMatrix<N, N> massMatrix(const RealMatrix<1, N>& lhs, const RealMatrix<1, N>& rhs)
{
  RealMatrix<N, N> r;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) r(i, j) = lhs(0, i) * rhs(0, j) * (i == j ? 2.0 : 1.0);
  return r;
}

//! outer product
AFB_HOST_DEVICE inline RealMatrix<4, 4> operator^(const Real4& lhs, const Real4& rhs)
{
  RealMatrix<4, 4> r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r(i, j) = lhs[i] * rhs[j];
  return r;
}

/*---------------------------------------------------------------------------*/
//! N reals.
template <int N>
class RealVector {
 public:
  AFB_HOST_DEVICE RealVector()
  {
    for (int k = 0; k < N; ++k) m_v[k] = 0.0;
  }
  AFB_HOST_DEVICE RealVector(std::initializer_list<Real> values)
  {
    int k = 0;
    for (Real x : values)
      if (k < N) m_v[k++] = x;
    for (; k < N; ++k) m_v[k] = 0.0;
  }
  static constexpr int size() { return N; }
  AFB_HOST_DEVICE Real& operator()(std::int32_t i) { return m_v[i]; }
  AFB_HOST_DEVICE Real operator()(std::int32_t i) const { return m_v[i]; }
  AFB_HOST_DEVICE Real& operator[](std::int32_t i) { return m_v[i]; }
  AFB_HOST_DEVICE Real operator[](std::int32_t i) const { return m_v[i]; }
  AFB_HOST_DEVICE void fill(Real value)
  {
    for (int k = 0; k < N; ++k) m_v[k] = value;
  }
  AFB_HOST_DEVICE void multInPlace(Real s)
  {
    for (int k = 0; k < N; ++k) m_v[k] *= s;
  }
  AFB_HOST_DEVICE void addInPlace(Real s)
  {
    for (int k = 0; k < N; ++k) m_v[k] += s;
  }
  AFB_HOST_DEVICE void setEqualTo(const RealVector& b) { *this = b; }
  AFB_HOST_DEVICE void add(const RealVector& b) { *this += b; }
  AFB_HOST_DEVICE void sub(const RealVector& b) { *this -= b; }
  AFB_HOST_DEVICE RealVector& operator+=(const RealVector& o)
  {
    for (int k = 0; k < N; ++k) m_v[k] += o.m_v[k];
    return *this;
  }
  AFB_HOST_DEVICE RealVector& operator-=(const RealVector& o)
  {
    for (int k = 0; k < N; ++k) m_v[k] -= o.m_v[k];
    return *this;
  }
  AFB_HOST_DEVICE RealVector operator+(const RealVector& o) const
  {
    RealVector r(*this);
    r += o;
    return r;
  }
  AFB_HOST_DEVICE RealVector operator-(const RealVector& o) const
  {
    RealVector r(*this);
    r -= o;
    return r;
  }
  AFB_HOST_DEVICE RealVector operator-() const
  {
    RealVector r;
    for (int k = 0; k < N; ++k) r.m_v[k] = -m_v[k];
    return r;
  }
  AFB_HOST_DEVICE RealVector operator*(Real s) const
  {
    RealVector r(*this);
    r.multInPlace(s);
    return r;
  }
  AFB_HOST_DEVICE RealVector operator/(Real s) const
  {
    RealVector r;
    for (int k = 0; k < N; ++k) r.m_v[k] = m_v[k] / s;
    return r;
  }
  friend AFB_HOST_DEVICE RealVector operator*(Real s, const RealVector& v) { return v * s; }
  friend AFB_HOST_DEVICE Real dot(const RealVector& u, const RealVector& v)
  {
    Real s = 0.0;
    for (int k = 0; k < N; ++k) s += u.m_v[k] * v.m_v[k];
    return s;
  }

 private:
  Real m_v[N];
};

//! row vector times square matrix
template <int N>
AFB_HOST_DEVICE inline RealVector<N> operator*(const RealVector<N>& lhs, const RealMatrix<N, N>& rhs)
{
  RealVector<N> r;
  for (int j = 0; j < N; ++j) {
    Real s = 0.0;
    for (int i = 0; i < N; ++i) s += lhs(i) * rhs(i, j);
    r(j) = s;
  }
  return r;
}

//! outer product
template <int N>
AFB_HOST_DEVICE inline RealMatrix<N, N> operator^(const RealVector<N>& lhs, const RealVector<N>& rhs)
{
  RealMatrix<N, N> r;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) r(i, j) = lhs(i) * rhs(j);
  return r;
}

//! outer product with the diagonal doubled (see the RealMatrix<1,N> overload)
template <int N>
AFB_HOST_DEVICE inline RealMatrix<N, N> massMatrix(const RealVector<N>& lhs, const RealVector<N>& rhs)
{
  RealMatrix<N, N> r;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) r(i, j) = lhs(i) * rhs(j) * (i == j ? 2.0 : 1.0);
  return r;
}

} // namespace arcanefem_b200
