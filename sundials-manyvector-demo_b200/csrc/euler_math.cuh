// ---------------------------------------------------------------------------
// euler_math.cuh -- face-flux arithmetic of the fluid RHS, written for the FP64 pipe.
//
// What the reference computes per face (face_flux, /root/reference/src/utilities.cpp:270-479)
// is kept; how it is computed is reorganised so that one face costs ~750 FP64-pipe
// instructions for the five fluid fields and ~100 per tracer instead of the
// reference's ~1550 / ~150:
//   * the 0.5 of the Lax-Friedrichs split (:388,:431) is folded out of the WENO input:
//     WENO(0.5 g; eps) == 0.5 WENO(g; 4 eps) exactly (power-of-two scaling), so the
//     split is g = F +- alpha*w and one 0.5 is applied to the sum f+ + f- at the end;
//   * LV/RV (:312-364) are never formed: their structural zeros are skipped and the
//     rows that share sub-expressions are evaluated together (13 ops per projection);
//   * tracers: F +- alpha*w = (u +- alpha)*c  (identity projection, :395,:439,:473);
//   * the three nonlinear weights are combined over a common denominator, so one
//     reciprocal per reconstruction replaces the reference's four divisions (:408-422);
//   * f- is f+ on the mirrored five points (:443-467 is :399-423 reflected).
// Every one of these is an exact identity in real arithmetic; in FP64 they change
// rounding at the 1e-16 level (the same size as letting the compiler contract FMAs in
// the reference itself).  Quirks that matter for parity are
// kept: c^2 without the square root in the eigenvectors (:309), the face-local 6-point
// alpha (:368-380), SUNRsqrt's "non-positive -> 0" (:298-299,:378), epsilon added before
// squaring (:408-410).
//
// The file is host/device so that tests can run exactly this arithmetic on the CPU
// (tests/emu) against the oracle without a GPU.
// ---------------------------------------------------------------------------
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EB_HD __host__ __device__ __forceinline__
#else
#define EB_HD inline
#endif

#if defined(EB_WENO_ONE_NEWTON)
#define EB_WENO_RCP fast_rcp1
#elif defined(EB_RCP_TWO_NEWTON)
#define EB_WENO_RCP fast_rcp
#else
#define EB_WENO_RCP fast_rcp3
#endif

// EB_FOLD_HALF: the face functions return TWICE the reference's flux (f+ + f- without the 0.5 of the
// Lax-Friedrichs split) and the divergence multiplies by 0.5/dx instead of 1/dx -- the same bits (scaling by
// a power of two commutes with every rounding on the way, under- and overflow aside), one multiplication
// less per flux.  Host code forms the inverse spacings with EB_RD_SCALE.  Off in the builds that divide.
#if defined(EB_STRICT) || defined(EB_TRUE_DIVISION) || defined(EB_NO_FOLD_HALF)
#define EB_FOLD_HALF 0
#define EB_RD_SCALE 1.0
#else
#define EB_FOLD_HALF 1
#define EB_RD_SCALE 0.5
#endif

namespace eb {

// SUNRsqrt of SUNDIALS 6.2 (sundials_math.h): x <= 0 gives 0.
EB_HD double sun_sqrt(double x) { return (x <= 0.0) ? 0.0 : sqrt(x); }

// Reciprocals on the FP64 pipe from the hardware seed (MUFU.RCP64H: reads the upper 32 bits of b, |1 - b r| <
// 2^-19); the full IEEE division sequence costs roughly twice as many FP64 issue slots.
//   fast_rcp   two Newton steps (4 FMAs): correctly rounded in every case tried (tools note in DESIGN 3.1);
//              used where the reciprocal scales whole fluxes (1/rho, the Roe average, 1/c^2).
//   fast_rcp3  one third-order step, 1/b = r (1 + e + e^2 + O(e^3)), e = 1 - b r (3 dependent FMAs): the
//              truncated tail e^3 < 2^-57 lies below the last bit, the result is within 1 ulp; used in weno5,
//              where the reciprocal multiplies the nonlinear correction term only.
//   fast_rcp1  one Newton step (relative error ~1e-14), -DEB_WENO_ONE_NEWTON: measured, not used.
EB_HD double fast_rcp1(double b)
{
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / b;
#endif
}

EB_HD double fast_rcp(double b)
{
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / b;
#endif
}

EB_HD double fast_rcp3(double b)
{
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
#elif defined(EB_CUDA_EMU)
  // CPU tier (tests/emu): a single-precision reciprocal as the seed (about as good as the hardware's), then the
  // same three FMAs, so that the emulated kernel carries the last-bit behaviour of this form
  const double r = (double)(1.0f / (float)b);
  if (!(r > 0.0) || !(r < 1.7e38) || !(b > 1e-37)) return 1.0 / b;
  const double e = fma(-b, r, 1.0);
  const double t = fma(e, e, e);
  return fma(r, t, r);
#else
  return 1.0 / b;
#endif
}

// Fifth-order WENO value at the face from five samples, left-biased ("f+") form of
// utilities.cpp:399-423 applied to UNHALVED split fluxes (hence 4*epsilon).  Call with
// the samples reversed for the right-biased ("f-") form.
//   result = q2 + [ a1 (q1-q2) + a3 (q3-q2) ] / (a1+a2+a3),  a_k = d_k / (eps+beta_k)^2
// with q1-q2 = -(D1-D2)/6 and q3-q2 = (D3-D2)/3 (D_k the second differences).
// Everything is built from the four first differences d_j = v_{j+1} - v_j (40 FP64-pipe
// instructions with two Newton steps in the reciprocal):
//   D1 = d3-d2, D2 = d2-d1, D3 = d1-d0;  E1 = d3-3 d2, E2 = -(d1+d2), E3 = 3 d1-d0;
//   q2 = v2 + (2 d2 + d1)/6;  numerator and denominator scaled by 10.
#ifndef EB_WENO_CLASSIC
// the reconstruction from the four first differences and the centre sample (35 instructions)
EB_HD double weno5_d(double d0, double d1, double d2, double d3, double v2)
{
  // beta_k/bc = D_k^2 + (0.25/bc) E_k^2; the common factor 1/bc cancels in the weight ratios
  const double c2 = 0.25 / (13.0 / 12.0);
  const double epsb = (4.0 * 1e-6) / (13.0 / 12.0);
  const double D1 = d3 - d2, D2 = d2 - d1, D3 = d1 - d0;
  const double E1 = fma(-3.0, d2, d3);
  const double E2 = d1 + d2;
  const double E3 = fma(3.0, d1, -d0);
  const double b1 = fma(D1, D1, fma(c2 * E1, E1, epsb));
  const double b2 = fma(D2, D2, fma(c2 * E2, E2, epsb));
  const double b3 = fma(D3, D3, fma(c2 * E3, E3, epsb));
  const double s1 = b1 * b1, s2 = b2 * b2, s3 = b3 * b3;
  const double P1 = s2 * s3, P2 = s1 * s3, P3 = s1 * s2;
  const double den = fma(3.0, P1, fma(6.0, P2, P3));
  const double t3 = P3 * (D3 - D2), t1 = P1 * (D1 - D2);
  const double num = fma(1.0 / 3.0, t3, -0.5 * t1);
  const double q2 = fma(1.0 / 6.0, fma(2.0, d2, d1), v2);
  return fma(num, EB_WENO_RCP(den), q2);
}
EB_HD double weno5(double v0, double v1, double v2, double v3, double v4)
{
  return weno5_d(v1 - v0, v2 - v1, v3 - v2, v4 - v3, v2);
}
#else
// the first formulation (43 instructions), kept for A/B runs: -DEB_WENO_CLASSIC
EB_HD double weno5(double v0, double v1, double v2, double v3, double v4)
{
  const double c2 = 0.25 / (13.0 / 12.0);
  const double epsb = (4.0 * 1e-6) / (13.0 / 12.0);
  const double D1 = fma(-2.0, v3, v2) + v4;
  const double D2 = fma(-2.0, v2, v1) + v3;
  const double D3 = fma(-2.0, v1, v0) + v2;
  const double E1 = fma(3.0, v2, fma(-4.0, v3, v4));
  const double E2 = v1 - v3;
  const double E3 = fma(3.0, v2, fma(-4.0, v1, v0));
  const double b1 = fma(D1, D1, fma(c2 * E1, E1, epsb));
  const double b2 = fma(D2, D2, fma(c2 * E2, E2, epsb));
  const double b3 = fma(D3, D3, fma(c2 * E3, E3, epsb));
  const double s1 = b1 * b1, s2 = b2 * b2, s3 = b3 * b3;
  const double P1 = s2 * s3, P2 = s1 * s3, P3 = s1 * s2;
  const double den = fma(0.3, P1, fma(0.6, P2, 0.1 * P3));
  const double num = fma((0.1 / 3.0) * P3, D3 - D2, (-0.3 / 6.0) * P1 * (D1 - D2));
  const double q2 = fma(1.0 / 3.0, v3, fma(5.0 / 6.0, v2, (-1.0 / 6.0) * v1));
  return fma(num, EB_WENO_RCP(den), q2);
}
#endif

// Roe-averaged face state and the projection coefficients derived from it.
struct Eigen {
  double u, v, w, H, q, cs;   // cs = (gamma-1)(H - q/2): c^2, NOT c (utilities.cpp:309)
  double gc, hgc, hc;         // (gamma-1)/cs, half of it, 0.5/cs
};

// Characteristic projection y = LV x (utilities.cpp:342-364 with the zeros skipped).
EB_HD void project(const Eigen& E, double x0, double x1, double x2, double x3, double x4,
                   double& y0, double& y1, double& y2, double& y3, double& y4)
{
  const double S = fma(-E.w, x3, fma(-E.v, x2, fma(-E.u, x1, x4)));
  const double C = E.hgc * fma(0.5 * E.q, x0, S);
  const double D = E.hc * fma(E.u, x0, -x1);
  y0 = C + D;
  y4 = C - D;
  y1 = fma(-E.v, x0, x2);
  y2 = fma(-E.w, x0, x3);
  y3 = -E.gc * fma(E.q - E.H, x0, S);
}

// Per-cell quantities every face that touches the cell needs (each cell sits in 18
// stencils per RHS): 1/rho, pressure (euler3D.hpp:1383-1388), sound speed
// sqrt(gamma p / rho) and sqrt(rho), the last two with SUNRsqrt semantics.  Computed once
// per cell by aux_kernel into four arrays; on the fly for ghost/halo points.
struct CellAux { double rinv, p, c, sr; };
EB_HD CellAux cell_aux(double gamma, double r, double mx, double my, double mz, double e)
{
  CellAux a;
  a.rinv = fast_rcp(r);
  const double m2sum = fma(mz, mz, fma(my, my, mx * mx));
  a.p = (gamma - 1.0) * fma(-0.5 * m2sum, a.rinv, e);
  a.c = sun_sqrt(gamma * a.p * a.rinv);
  a.sr = sun_sqrt(r);
  return a;
}

// Six-point stencil of the five fluid fields in sweep-aligned order: mn is the
// momentum normal to the face, m1/m2 the tangential ones in the order the reference's
// swap leaves them (utilities.cpp:283-285: x:(mx,my,mz)  y:(my,mx,mz)  z:(mz,my,mx)),
// plus the per-cell derived values (sr only for the two cells adjacent to the face).
struct FluidStencil {
  double r[6], mn[6], m1[6], m2[6], e[6];
  double rinv[6], p[6], c[6];
  double srL, srR;
};

// Fluid part of one face.  Returns through `f` the five face fluxes in sweep-aligned
// order (rho, normal, tan1, tan2, energy), and through alpha / u[6] what the tracers
// of the same face need (face-local max wave speed, normal velocity per point).
EB_HD void fluid_face(const FluidStencil& s, double gamma, double f[5], double& alpha_out, double u[6])
{
  const double gm1 = gamma - 1.0;
  const double* p = s.p;
  double alpha = 0.0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    u[j] = s.mn[j] * s.rinv[j];
    const double a = fabs(u[j]) + s.c[j];
    alpha = (alpha < a) ? a : alpha;
  }
  alpha_out = alpha;

  // Roe average of the two cells adjacent to the face (utilities.cpp:298-304):
  // 0.5*(a/sL + b/sR)/(0.5*(sL+sR)) = (a/sL + b/sR)/(sL+sR), and 1/sqrt(rho) = sqrt(rho)/rho.
  // A non-positive density gives sL = 0 in the reference and a division by it: keep that
  // non-finite (the Dirichlet-ghost corner).
  Eigen E;
  {
    const double isL = (s.srL > 0.0) ? s.srL * s.rinv[2] : nan("");
    const double isR = (s.srR > 0.0) ? s.srR * s.rinv[3] : nan("");
    const double iS = fast_rcp(s.srL + s.srR);
    E.u = fma(s.mn[2], isL, s.mn[3] * isR) * iS;
    E.v = fma(s.m1[2], isL, s.m1[3] * isR) * iS;
    E.w = fma(s.m2[2], isL, s.m2[3] * isR) * iS;
    E.H = fma(p[2] + s.e[2], isL, (p[3] + s.e[3]) * isR) * iS;
    E.q = fma(E.w, E.w, fma(E.v, E.v, E.u * E.u));
    E.cs = gm1 * fma(-0.5, E.q, E.H);
    const double cinv = fast_rcp(E.cs);
    E.gc = gm1 * cinv;
    E.hgc = 0.5 * E.gc;
    E.hc = 0.5 * cinv;
  }

  // Split fluxes g+ (points 0..4) and g- (points 1..5), projected (utilities.cpp:386-396,429-440)
  double gp[5][5], gm[5][5];   // [point][characteristic]
#ifndef EB_PROJECT_SHARED
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const double F0 = s.mn[j];
    const double F1 = fma(u[j], s.mn[j], p[j]);
    const double F2 = u[j] * s.m1[j];
    const double F3 = u[j] * s.m2[j];
    const double F4 = u[j] * (s.e[j] + p[j]);
    if (j < 5)
      project(E, fma(alpha, s.r[j], F0), fma(alpha, s.mn[j], F1), fma(alpha, s.m1[j], F2),
              fma(alpha, s.m2[j], F3), fma(alpha, s.e[j], F4),
              gp[j][0], gp[j][1], gp[j][2], gp[j][3], gp[j][4]);
    if (j > 0)
      project(E, fma(-alpha, s.r[j], F0), fma(-alpha, s.mn[j], F1), fma(-alpha, s.m1[j], F2),
              fma(-alpha, s.m2[j], F3), fma(-alpha, s.e[j], F4),
              gm[j - 1][0], gm[j - 1][1], gm[j - 1][2], gm[j - 1][3], gm[j - 1][4]);
  }
#else
  // -DEB_PROJECT_SHARED (measured, not the default: 512^3 NVAR=5 27.7 -> 25.7 ms, NVAR=15 58.0 -> 57.8 ms, but
  // on smooth states, where the right-hand side is rounding noise of the flux terms, its distance from the
  // reference is that of an independent rounding sequence -- 1.2e-12 in e_t on the advection golden case where
  // the direct form, which rounds the quantities the reference rounds, stays at 0.9e-12).
  // The flux is F_j = u_j w_j + p_j (0,1,0,0,u_j), so the two projected splits of a point share one
  // projection of its state:
  //   LV (F_j +- alpha w_j) = (u_j +- alpha) A_j + p_j B_j ,   A_j = LV w_j ,
  //   B_j = LV (0,1,0,0,u_j) = (hgc s - hc, 0, 0, -gc s, hgc s + hc) ,  s = u_j - u_roe ,  gc = 2 hgc :
  // 30 instructions for the four points that need both splits instead of 41 (five to form F_j, then
  // 2 x (5 + 13)); the two end points, which need one split each, are cheaper the direct way (23).
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j == 0 || j == 5) {
      const double sa = (j == 0) ? alpha : -alpha;
      const double x0 = fma(sa, s.r[j], s.mn[j]);
      const double x1 = fma(sa, s.mn[j], fma(u[j], s.mn[j], p[j]));
      const double x2 = fma(sa, s.m1[j], u[j] * s.m1[j]);
      const double x3 = fma(sa, s.m2[j], u[j] * s.m2[j]);
      const double x4 = fma(sa, s.e[j], u[j] * (s.e[j] + p[j]));
      if (j == 0) project(E, x0, x1, x2, x3, x4, gp[0][0], gp[0][1], gp[0][2], gp[0][3], gp[0][4]);
      else project(E, x0, x1, x2, x3, x4, gm[4][0], gm[4][1], gm[4][2], gm[4][3], gm[4][4]);
    } else {
      double a0, a1, a2, a3, a4;
      project(E, s.r[j], s.mn[j], s.m1[j], s.m2[j], s.e[j], a0, a1, a2, a3, a4);
      const double pc = p[j] * (E.hgc * (u[j] - E.u));
      const double ph = p[j] * E.hc;
      const double b0 = pc - ph, b4 = pc + ph, b3 = -2.0 * pc;
      const double vp = u[j] + alpha, vm = u[j] - alpha;
      gp[j][0] = fma(vp, a0, b0); gp[j][1] = vp * a1; gp[j][2] = vp * a2; gp[j][3] = fma(vp, a3, b3); gp[j][4] = fma(vp, a4, b4);
      gm[j - 1][0] = fma(vm, a0, b0); gm[j - 1][1] = vm * a1; gm[j - 1][2] = vm * a2; gm[j - 1][3] = fma(vm, a3, b3); gm[j - 1][4] = fma(vm, a4, b4);
    }
  }
#endif

  // WENO per characteristic field; 0.5 of the split applied once to f+ + f- (EB_FOLD_HALF: by the divergence)
  double ff[5];
#pragma unroll
  for (int c = 0; c < 5; c++) {
    const double f2 = weno5(gp[0][c], gp[1][c], gp[2][c], gp[3][c], gp[4][c]) +
                      weno5(gm[4][c], gm[3][c], gm[2][c], gm[1][c], gm[0][c]);
    ff[c] = EB_FOLD_HALF ? f2 : 0.5 * f2;
  }

  // Back to conserved variables: RV ff (utilities.cpp:318-340,470-472)
  const double f0 = ff[0] + ff[3] + ff[4];
  const double dl = ff[4] - ff[0];
  f[0] = f0;
  f[1] = fma(E.u, f0, E.cs * dl);
  f[2] = fma(E.v, f0, ff[1]);
  f[3] = fma(E.w, f0, ff[2]);
  f[4] = fma(E.H, ff[0] + ff[4], fma(E.u * E.cs, dl, fma(E.v, ff[1], fma(E.w, ff[2], 0.5 * E.q * ff[3]))));
}

// One tracer on one face: c[6] are its stencil values, up[j] = u_j + alpha (j=0..4 used),
// um[j] = u_j - alpha (j=1..5 used).  (utilities.cpp:376-377,388,395,431,439,473)
EB_HD double tracer_face(const double c[6], const double up[6], const double um[6])
{
#ifdef EB_WENO_CLASSIC
  const double fp = weno5(up[0] * c[0], up[1] * c[1], up[2] * c[2], up[3] * c[3], up[4] * c[4]);
  const double fm = weno5(um[5] * c[5], um[4] * c[4], um[3] * c[3], um[2] * c[2], um[1] * c[1]);
#else
  // the two outermost samples of each reconstruction enter only through a first difference: one FMA forms
  // that difference from the product's factors (a single rounding where the reference has two), so
  // a reconstruction takes three products and four differences instead of five and four
  const double p1 = up[1] * c[1], p2 = up[2] * c[2], p3 = up[3] * c[3];
  const double fp = weno5_d(fma(-up[0], c[0], p1), p2 - p1, p3 - p2, fma(up[4], c[4], -p3), p2);
  const double q4 = um[4] * c[4], q3 = um[3] * c[3], q2 = um[2] * c[2];
  const double fm = weno5_d(fma(-um[5], c[5], q4), q3 - q4, q2 - q3, fma(um[1], c[1], -q2), q3);
#endif
  return EB_FOLD_HALF ? fp + fm : 0.5 * (fp + fm);
}

}  // namespace eb
