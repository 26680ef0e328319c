// CPU-only checks of the value types and CSR row iteration of the facade (no GPU call is made: the view is built over
// host arrays, which is how a unit test of user code written against the reference's interfaces would use it).
//   g++ -std=c++17 -Wall -Werror -Iinclude tests/cpp/types_driver.cpp -Larcanefem_b200 -lafb200
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "arcanefem_b200/FemUtils.h"

using namespace arcanefem_b200;

static int failures = 0;
#define EXPECT(...)                                                       \
  do {                                                                    \
    if (!(__VA_ARGS__)) {                                                        \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #__VA_ARGS__);       \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

static bool close(double a, double b) { return std::fabs(a - b) <= 1e-14 * (1.0 + std::fabs(b)); }

int main()
{
  // ---- Real4 ----
  Real4 a{ { 1.0, 2.0, 3.0, 4.0 } }, b{ { 0.5, -1.0, 2.0, 0.0 } };
  Real4 c = a + b * 2.0 - 0.5 * a;
  EXPECT(close(c[0], 1.5) && close(c[1], -1.0) && close(c[2], 5.5) && close(c[3], 2.0));
  RealMatrix<4, 4> outer = a ^ b;
  EXPECT(close(outer(2, 1), -3.0) && close(outer(3, 0), 2.0) && close(outer(0, 3), 0.0));

  // ---- RealMatrix ----
  RealMatrix<2, 3> m{ { 1.0, 2.0, 3.0 }, { 4.0, 5.0, 6.0 } };
  RealMatrix<2, 3> flat{ 1.0, 2.0, 3.0, 4.0, 5.0, 6.0 };
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) EXPECT(m(i, j) == flat(i, j));
  RealMatrix<3, 2> mt = matrixTranspose(m);
  EXPECT(mt(2, 0) == 3.0 && mt(0, 1) == 4.0);
  RealMatrix<2, 2> mm = matrixMultiplication(m, mt); // m m^T
  EXPECT(close(mm(0, 0), 14.0) && close(mm(0, 1), 32.0) && close(mm(1, 0), 32.0) && close(mm(1, 1), 77.0));
  RealMatrix<2, 2> id{ { 1.0, 0.0 }, { 0.0, 1.0 } };
  RealMatrix<2, 2> sum = matrixAddition(mm, id);
  EXPECT(close(sum(0, 0), 15.0) && close(sum(1, 1), 78.0) && close(sum(0, 1), 32.0));
  RealMatrix<2, 2> lin = (mm + id * 2.0 - id) / 2.0;
  EXPECT(close(lin(0, 0), 7.5) && close(lin(1, 0), 16.0));
  lin += id;
  lin.multInPlace(2.0);
  EXPECT(close(lin(0, 0), 17.0) && close((-lin)(1, 1), -80.0) && close((3.0 * id)(1, 1), 3.0));
  RealMatrix<1, 3> phi{ 1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0 };
  RealMatrix<3, 3> mass = massMatrix(phi, phi); // (1 + delta_ij) / 9: the P1 triangle mass pattern
  EXPECT(close(mass(0, 0), 2.0 / 9.0) && close(mass(0, 2), 1.0 / 9.0));
  RealMatrix<3, 3> z;
  EXPECT(z(1, 2) == 0.0);
  z.fill(4.0);
  EXPECT(z(2, 2) == 4.0 && RealMatrix<3, 3>::totalNbElement() == 9);

  // ---- RealVector ----
  RealVector<3> u{ 1.0, 2.0, 3.0 }, v{ -1.0, 0.5, 2.0 };
  EXPECT(close(dot(u, v), 6.0));
  RealVector<3> w = (u + v) * 2.0 - u / 2.0;
  EXPECT(close(w(0), -0.5) && close(w[1], 4.0) && close(w(2), 8.5));
  w.addInPlace(1.0);
  w.multInPlace(2.0);
  w.sub(u);
  w.add(v);
  EXPECT(close(w(0), -1.0) && close(w(1), 8.5) && close(w(2), 18.0));
  RealMatrix<3, 3> uv = u ^ v;
  EXPECT(close(uv(2, 0), -3.0) && close(uv(1, 2), 4.0));
  RealVector<3> row = u * uv; // u^T (u v^T) = |u|^2 v^T
  EXPECT(close(row(0), -14.0) && close(row(1), 7.0) && close(row(2), 28.0));
  RealMatrix<3, 3> vmass = massMatrix(u, u);
  EXPECT(close(vmass(1, 1), 8.0) && close(vmass(0, 2), 3.0));
  RealVector<3> copy;
  copy.setEqualTo(u);
  EXPECT(copy(2) == 3.0 && (-copy)(0) == -1.0 && (2.0 * copy)(1) == 4.0 && RealVector<3>::size() == 3);

  // ---- CSR view: rows without sentinel, iteration, search ----
  //  [ 4 -1  . ]
  //  [-1  4 -1 ]
  //  [ . -1  4 ]
  const Int32 rows[3] = { 0, 2, 5 }, nbcol[3] = { 2, 3, 2 }, cols[7] = { 0, 1, 0, 1, 2, 1, 2 };
  Real vals[7] = { 4.0, -1.0, -1.0, 4.0, -1.0, -1.0, 4.0 };
  CsrFormatMatrixView view(DeviceSpan<const Int32>{ rows, 3 }, DeviceSpan<const Int32>{ nbcol, 3 }, DeviceSpan<const Int32>{ cols, 7 }, DeviceSpan<Real>{ vals, 7 });
  EXPECT(view.nbRow() == 3 && view.nbColumn() == 7 && view.nbValue() == 7);
  EXPECT(view.row(1) == 2 && view.nbColumnForRow(1) == 3);
  int visited = 0;
  for (Int32 r = 0; r < view.nbRow(); ++r) {
    Real row_sum = 0.0;
    for (CsrRowColumnIndex rc : view.rowRange(r)) {
      row_sum += view.value(rc);
      ++visited;
    }
    EXPECT(view.rowRange(r).size() == view.nbColumnForRow(r));
    EXPECT(close(row_sum, r == 1 ? 2.0 : 3.0));
  }
  EXPECT(visited == 7);
  CsrRowColumnIndex hit = view.tryFindColumnInRow(2, 1), miss = view.tryFindColumnInRow(0, 2);
  EXPECT(!hit.isNull() && hit.value() == 5 && view.column(hit) == 1);
  EXPECT(miss.isNull() && miss.value() == -1);
  view.value(hit) += 0.25; // the reference's matrixAddValue on a view
  EXPECT(close(vals[5], -0.75));
  EXPECT(!CsrRowColumnIterator().isValid());

  if (failures) {
    std::printf("%d check(s) failed\n", failures);
    return 1;
  }
  std::printf("types ok\n");
  return 0;
}
