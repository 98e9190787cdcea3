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
// 7 Fehlberg 6-4-5, 8 Dormand-Prince 7-4-5, 10 Verner 8-5-6, 11 Fehlberg 13-7-8, 12 Knoth-Wolke 3-3
// (no embedding: fixed step only).
struct Table { int s, p, q; bool embedded; double A[13][13], b[13], bh[13]; };
static void set_row(double* dst, std::initializer_list<double> v) { int i = 0; for (double x : v) dst[i++] = x; }
bool make_table(int order, int etable, Table& T)
{
  memset(&T, 0, sizeof T);
  T.embedded = true;
  int id = etable;
  if (order == 2) id = 0; else if (order == 3) id = 1; else if (order == 4) id = 3; else if (order == 5) id = 6;
  else if (order == 6) id = 10; else if (order == 8) id = 11;
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
  case 10: T.s = 8; T.p = 6; T.q = 5;
    set_row(T.A[1], {1.0 / 6}); set_row(T.A[2], {4.0 / 75, 16.0 / 75}); set_row(T.A[3], {5.0 / 6, -8.0 / 3, 5.0 / 2});
    set_row(T.A[4], {-165.0 / 64, 55.0 / 6, -425.0 / 64, 85.0 / 96});
    set_row(T.A[5], {12.0 / 5, -8.0, 4015.0 / 612, -11.0 / 36, 88.0 / 255});
    set_row(T.A[6], {-8263.0 / 15000, 124.0 / 75, -643.0 / 680, -81.0 / 250, 2484.0 / 10625, 0.0});
    set_row(T.A[7], {3501.0 / 1720, -300.0 / 43, 297275.0 / 52632, -319.0 / 2322, 24068.0 / 84065, 0.0, 3850.0 / 26703});
    set_row(T.b, {3.0 / 40, 0.0, 875.0 / 2244, 23.0 / 72, 264.0 / 1955, 0.0, 125.0 / 11592, 43.0 / 616});
    set_row(T.bh, {13.0 / 160, 0.0, 2375.0 / 5984, 5.0 / 16, 12.0 / 85, 3.0 / 44, 0.0, 0.0}); break;
  case 11: T.s = 13; T.p = 8; T.q = 7;
    set_row(T.A[1], {2.0 / 27}); set_row(T.A[2], {1.0 / 36, 1.0 / 12}); set_row(T.A[3], {1.0 / 24, 0.0, 1.0 / 8});
    set_row(T.A[4], {5.0 / 12, 0.0, -25.0 / 16, 25.0 / 16}); set_row(T.A[5], {1.0 / 20, 0.0, 0.0, 1.0 / 4, 1.0 / 5});
    set_row(T.A[6], {-25.0 / 108, 0.0, 0.0, 125.0 / 108, -65.0 / 27, 125.0 / 54});
    set_row(T.A[7], {31.0 / 300, 0.0, 0.0, 0.0, 61.0 / 225, -2.0 / 9, 13.0 / 900});
    set_row(T.A[8], {2.0, 0.0, 0.0, -53.0 / 6, 704.0 / 45, -107.0 / 9, 67.0 / 90, 3.0});
    set_row(T.A[9], {-91.0 / 108, 0.0, 0.0, 23.0 / 108, -976.0 / 135, 311.0 / 54, -19.0 / 60, 17.0 / 6, -1.0 / 12});
    set_row(T.A[10], {2383.0 / 4100, 0.0, 0.0, -341.0 / 164, 4496.0 / 1025, -301.0 / 82, 2133.0 / 4100, 45.0 / 82, 45.0 / 164, 18.0 / 41});
    set_row(T.A[11], {3.0 / 205, 0.0, 0.0, 0.0, 0.0, -6.0 / 41, -3.0 / 205, -3.0 / 41, 3.0 / 41, 6.0 / 41, 0.0});
    set_row(T.A[12], {-1777.0 / 4100, 0.0, 0.0, -341.0 / 164, 4496.0 / 1025, -289.0 / 82, 2193.0 / 4100, 51.0 / 82, 33.0 / 164, 12.0 / 41, 0.0, 1.0});
    set_row(T.b, {0.0, 0.0, 0.0, 0.0, 0.0, 34.0 / 105, 9.0 / 35, 9.0 / 35, 9.0 / 280, 9.0 / 280, 0.0, 41.0 / 840, 41.0 / 840});
    set_row(T.bh, {41.0 / 840, 0.0, 0.0, 0.0, 0.0, 34.0 / 105, 9.0 / 35, 9.0 / 35, 9.0 / 280, 9.0 / 280, 41.0 / 840, 0.0, 0.0}); break;
  case 12: T.s = 3; T.p = 3; T.q = 0; T.embedded = false;
    set_row(T.A[1], {1.0 / 3}); set_row(T.A[2], {-3.0 / 16, 15.0 / 16}); set_row(T.b, {1.0 / 6, 3.0 / 10, 8.0 / 15}); break;
  default: return false;
  }
  return true;
}

