// ---------------------------------------------------------------------------
// erk_tables.hpp -- explicit Butcher tables of the native driver (host/euler3d_b200.cpp).
// Same tables and the same selection rule as driver.py (select_table); pinned against it and
// against the order conditions by tests/test_driver_cpu.py.
// ---------------------------------------------------------------------------
#pragma once
#include <cstring>
#include <initializer_list>

// Embedded explicit Runge-Kutta tables.  "order" selects ARKODE's default table of that order;
// order = 0 selects by ARKODE_ERKTableID ("etable", euler3D_main.cpp:207-213).  Provided ids:
// 0 Heun-Euler 2-1-2, 1 Bogacki-Shampine 4-2-3, 3 Zonneveld 5-3-4, 6 Cash-Karp 6-4-5,
// 7 Fehlberg 6-4-5, 8 Dormand-Prince 7-4-5, 12 Knoth-Wolke 3-3 (no embedding: fixed step only).
struct Table { int s, p, q; bool embedded; double A[7][7], b[7], bh[7]; };
static void set_row(double* dst, std::initializer_list<double> v) { int i = 0; for (double x : v) dst[i++] = x; }
bool make_table(int order, int etable, Table& T)
{
  memset(&T, 0, sizeof T);
  T.embedded = true;
  int id = etable;
  if (order == 2) id = 0; else if (order == 3) id = 1; else if (order == 4) id = 3; else if (order == 5) id = 6;
  else if (order != 0) return false;
  else if (etable < 0) id = 3;
  switch (id) {
  case 0: T.s = 2; T.p = 2; T.q = 1; set_row(T.A[1], {1.0}); set_row(T.b, {0.5, 0.5}); set_row(T.bh, {1.0, 0.0}); break;
  case 1: T.s = 4; T.p = 3; T.q = 2;
    set_row(T.A[1], {0.5}); set_row(T.A[2], {0.0, 0.75}); set_row(T.A[3], {2.0 / 9, 1.0 / 3, 4.0 / 9});
    set_row(T.b, {2.0 / 9, 1.0 / 3, 4.0 / 9, 0.0}); set_row(T.bh, {7.0 / 24, 0.25, 1.0 / 3, 0.125}); break;
  case 3: T.s = 5; T.p = 4; T.q = 3;
    set_row(T.A[1], {0.5}); set_row(T.A[2], {0.0, 0.5}); set_row(T.A[3], {0.0, 0.0, 1.0});
    set_row(T.A[4], {5.0 / 32, 7.0 / 32, 13.0 / 32, -1.0 / 32});
    set_row(T.b, {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6, 0.0});
    set_row(T.bh, {-0.5, 7.0 / 3, 7.0 / 3, 13.0 / 6, -16.0 / 3}); break;
  case 6: T.s = 6; T.p = 5; T.q = 4;
    set_row(T.A[1], {1.0 / 5}); set_row(T.A[2], {3.0 / 40, 9.0 / 40}); set_row(T.A[3], {3.0 / 10, -9.0 / 10, 6.0 / 5});
    set_row(T.A[4], {-11.0 / 54, 5.0 / 2, -70.0 / 27, 35.0 / 27});
    set_row(T.A[5], {1631.0 / 55296, 175.0 / 512, 575.0 / 13824, 44275.0 / 110592, 253.0 / 4096});
    set_row(T.b, {37.0 / 378, 0.0, 250.0 / 621, 125.0 / 594, 0.0, 512.0 / 1771});
    set_row(T.bh, {2825.0 / 27648, 0.0, 18575.0 / 48384, 13525.0 / 55296, 277.0 / 14336, 0.25}); break;
  case 7: T.s = 6; T.p = 5; T.q = 4;
    set_row(T.A[1], {0.25}); set_row(T.A[2], {3.0 / 32, 9.0 / 32});
    set_row(T.A[3], {1932.0 / 2197, -7200.0 / 2197, 7296.0 / 2197});
    set_row(T.A[4], {439.0 / 216, -8.0, 3680.0 / 513, -845.0 / 4104});
    set_row(T.A[5], {-8.0 / 27, 2.0, -3544.0 / 2565, 1859.0 / 4104, -11.0 / 40});
    set_row(T.b, {16.0 / 135, 0.0, 6656.0 / 12825, 28561.0 / 56430, -9.0 / 50, 2.0 / 55});
    set_row(T.bh, {25.0 / 216, 0.0, 1408.0 / 2565, 2197.0 / 4104, -0.2, 0.0}); break;
  case 8: T.s = 7; T.p = 5; T.q = 4;
    set_row(T.A[1], {0.2}); set_row(T.A[2], {3.0 / 40, 9.0 / 40}); set_row(T.A[3], {44.0 / 45, -56.0 / 15, 32.0 / 9});
    set_row(T.A[4], {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729});
    set_row(T.A[5], {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656});
    set_row(T.A[6], {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84});
    set_row(T.b, {35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84, 0.0});
    set_row(T.bh, {5179.0 / 57600, 0.0, 7571.0 / 16695, 393.0 / 640, -92097.0 / 339200, 187.0 / 2100, 1.0 / 40}); break;
  case 12: T.s = 3; T.p = 3; T.q = 0; T.embedded = false;
    set_row(T.A[1], {1.0 / 3}); set_row(T.A[2], {-3.0 / 16, 15.0 / 16}); set_row(T.b, {1.0 / 6, 3.0 / 10, 8.0 / 15}); break;
  default: return false;
  }
  return true;
}

