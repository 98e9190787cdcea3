// ---------------------------------------------------------------------------
// erk_stepper.hpp -- the embedded explicit Runge-Kutta loop of the native driver
// (host/euler3d_b200.cpp), independent of where the vectors live.  Same algorithm as driver.py
// (ERKStep): Butcher table from erk_tables.hpp, WRMS error test with scalar tolerances, PID step
// controller with ARKODE's default constants, error-test failures, optional fixed step, the CFL
// hook, stop times.  It stands in for what euler3D_main.cpp:191-417 asks of ARKODE's ARKStep; it
// is not ARKODE.
//
// Ops supplies:  typedef Vec;  void lincomb(Vec& out, int n, const double* c, Vec* const* v)
// (out = sum c[q] * *v[q], n <= 14);  double wrms(const Vec& x, const Vec& y, rtol, atol)
// (N_VWrmsNorm of x with weights 1/(rtol |y| + atol));  int rhs(t, Vec& y, Vec& ydot);
// int stability(Vec& w, t, cfl, double* dt).
// ---------------------------------------------------------------------------
#pragma once
#include <algorithm>
#include <cmath>
#include "erk_tables.hpp"

template <class Ops>
struct ErkStepper {
  typedef typename Ops::Vec Vec;
  Ops ops;
  Table T;
  Vec w, ytmp, yerr, k[13];
  double t = 0, h = 0, rtol = 1e-8, atol = 1e-12, hmin = 0, hmax = 0, h0 = 0, cfl = 0;
  int fixedstep = 0, mxsteps = 5000, maxnef = 7;
  double safety = 0.96, bias = 1.5, growth = 20.0, k1 = 0.58, k2 = 0.21, k3 = 0.1, etamx1 = 1e4, etamxf = 0.3;
  double e2 = 1.0, e3 = 1.0;
  long nst = 0, nst_a = 0, nfe = 0, netf = 0;
  bool failed = false;      // a right-hand side or stability evaluation returned an error

  void lincomb(Vec& out, int n, const double* c, Vec* const* v) { ops.lincomb(out, n, c, v); }
  double wrms(const Vec& x, const Vec& y) { return ops.wrms(x, y, rtol, atol); }
  void f(double tt, Vec& y, Vec& out)
  {
    nfe++;
    if (ops.rhs(tt, y, out) != 0) failed = true;
  }
  double initial_step(double tout)
  {
    if (fixedstep) return hmax;
    if (h0 > 0) return h0;
    f(t, w, k[0]);
    const double d0 = wrms(w, w), d1 = wrms(k[0], w);
    double hh = (d0 > 1e-5 && d1 > 1e-5) ? 0.01 * d0 / d1 : 1e-6;
    hh = std::min(hh, fabs(tout - t));
    { const double c[2] = {1.0, hh}; Vec* v[2] = {&w, &k[0]}; lincomb(ytmp, 2, c, v); }
    f(t + hh, ytmp, k[1]);
    { const double c[2] = {1.0, -1.0}; Vec* v[2] = {&k[1], &k[0]}; lincomb(yerr, 2, c, v); }
    const double d2 = wrms(yerr, w) / hh, dm = std::max(d1, d2);
    const double h1 = dm > 1e-15 ? pow(0.01 / dm, 1.0 / (T.p + 1)) : std::max(1e-6, 1e-3 * hh);
    return std::min(std::min(100.0 * hh, h1), fabs(tout - t));
  }
  double attempt(double hh)
  {
    for (int i = 0; i < T.s; i++) {
      if (i == 0) { f(t, w, k[0]); continue; }
      double c[16]; Vec* v[16]; int n = 0; double ci = 0;
      c[n] = 1.0; v[n++] = &w;
      for (int j = 0; j < i; j++) { ci += T.A[i][j]; if (T.A[i][j] != 0.0) { c[n] = hh * T.A[i][j]; v[n++] = &k[j]; } }
      lincomb(ytmp, n, c, v);
      f(t + ci * hh, ytmp, k[i]);
    }
    { double c[16]; Vec* v[16]; int n = 0; c[n] = 1.0; v[n++] = &w;
      for (int j = 0; j < T.s; j++) if (T.b[j] != 0.0) { c[n] = hh * T.b[j]; v[n++] = &k[j]; }
      lincomb(ytmp, n, c, v); }
    if (fixedstep) return 0.0;
    { double c[16]; Vec* v[16]; int n = 0;
      for (int j = 0; j < T.s; j++) if (T.b[j] != T.bh[j]) { c[n] = hh * (T.b[j] - T.bh[j]); v[n++] = &k[j]; }
      lincomb(yerr, n, c, v); }
    return bias * wrms(yerr, w);
  }
  double eta_pid(double dsm) const
  {
    const double e1 = std::max(dsm, 1e-10), kk = T.q + 1;
    return safety * pow(e1, -k1 / kk) * pow(e2, k2 / kk) * pow(e3, -k3 / kk);
  }
  // ARKStepEvolve(.., tout, .., ARK_NORMAL) with the stop time at tout: 0, or -1 on failure
  int evolve(double tout)
  {
    if (h == 0.0) h = initial_step(tout);
    long steps = 0;
    while (t < tout * (1 - 1e-14) - 1e-300) {
      if (steps >= mxsteps || failed) return -1;
      double hh = h;
      if (hmax > 0 && !fixedstep) hh = std::min(hh, hmax);
      if (cfl > 0 && !fixedstep) {
        double dt = 0;
        if (ops.stability(w, t, cfl, &dt) != 0) return -1;
        hh = std::min(hh, dt);
      }
      hh = std::min(hh, tout - t);
      int nef = 0;
      double dsm = 0;
      for (;;) {
        nst_a++;
        dsm = attempt(hh);
        if (failed) return -1;
        if (fixedstep || dsm <= 1.0) break;
        netf++; nef++;
        if (nef >= maxnef || hh <= std::max(hmin, 1e-14 * std::max(fabs(t), 1.0))) return -1;
        hh *= std::min(nef >= 2 ? etamxf : 1.0, std::max(0.1, eta_pid(dsm)));
      }
      std::swap(w, ytmp);
      t += hh; nst++; steps++;
      if (!fixedstep) {
        double eta = std::min(eta_pid(dsm), nst == 1 ? etamx1 : growth);
        if (eta > 1.0 && eta < 1.5) eta = 1.0;
        e3 = e2; e2 = std::max(dsm, 1e-10);
        h = std::max(hh * eta, hmin);
      }
    }
    // snap onto tout only across round-off: a tout at or behind the clock leaves it alone
    // (ARKStepEvolve in ARK_NORMAL mode never rewinds the integrator)
    if (fabs(t - tout) <= 1e-14 * std::max(fabs(tout), 1.0) + 1e-300) t = tout;
    return failed ? -1 : 0;
  }
};
