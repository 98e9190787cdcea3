// ---------------------------------------------------------------------------
// strict_face.cuh -- the reference's face_flux arithmetic, operation for operation
// (/root/reference/src/utilities.cpp:270-479), for the STRICT build of the library
// (-DEB_STRICT -fmad=false: libeulerb200_strict.so).
//
// The fast kernel's arithmetic (euler_math.cuh) is re-associated and FMA-contracted: equal to the
// reference to ~1e-15 normwise, but not bit for bit.  This variant exists to separate the two
// questions a parity failure could hide behind each other: with the same operation order, IEEE
// division / square root and no contraction, the CUDA path (stencil resolution, ghost maps, halo
// slabs, tile seams, shared-memory exchange, divergence) reproduces the reference BIT FOR BIT
// (tests/test_gpu_strict.py), so every difference of the fast build is rounding of the re-derived
// arithmetic and nothing else.  About 3x slower than the fast kernel; never the default.
// ---------------------------------------------------------------------------
#pragma once
#include "euler_math.cuh"

namespace eb {
namespace strict {

enum { MAXVAR = 64 };

EB_HD double eos(double gamma, double rho, double mx, double my, double mz, double et)
{
  return (gamma - 1.0) * (et - (mx * mx + my * my + mz * mz) * 0.5 / rho);     // euler3D.hpp:1383-1388
}

// utilities.cpp:399-423 (plus) / :443-467 (minus)
EB_HD double weno5(const double q[5], bool plus)
{
  const double bc = 1.083333333333333333333333333333333333333;
  const double eps = 1e-6;
  const double c13 = 0.3333333333333333333333333333333333333333;
  const double c56 = 0.8333333333333333333333333333333333333333;
  const double c16 = 0.1666666666666666666666666666666666666667;
  const double c76 = 1.166666666666666666666666666666666666667;
  const double c116 = 1.833333333333333333333333333333333333333;
  double t, b1, b2, b3, w1, w2, w3, f1, f2, f3;
  t = q[2] - 2.0 * q[3] + q[4];
  b1 = bc * (t * t);
  t = 3.0 * q[2] - 4.0 * q[3] + q[4];
  b1 = b1 + 0.25 * (t * t);
  t = q[1] - 2.0 * q[2] + q[3];
  b2 = bc * (t * t);
  t = q[1] - q[3];
  b2 = b2 + 0.25 * (t * t);
  t = q[0] - 2.0 * q[1] + q[2];
  b3 = bc * (t * t);
  t = q[0] - 4.0 * q[1] + 3.0 * q[2];
  b3 = b3 + 0.25 * (t * t);
  if (plus) {
    w1 = 0.3 / ((eps + b1) * (eps + b1));
    w2 = 0.6 / ((eps + b2) * (eps + b2));
    w3 = 0.1 / ((eps + b3) * (eps + b3));
    f1 = c13 * q[2] + c56 * q[3] - c16 * q[4];
    f2 = -c16 * q[1] + c56 * q[2] + c13 * q[3];
    f3 = c13 * q[0] - c76 * q[1] + c116 * q[2];
  } else {
    w1 = 0.1 / ((eps + b1) * (eps + b1));
    w2 = 0.6 / ((eps + b2) * (eps + b2));
    w3 = 0.3 / ((eps + b3) * (eps + b3));
    f1 = c116 * q[2] - c76 * q[3] + c13 * q[4];
    f2 = c13 * q[1] + c56 * q[2] - c16 * q[3];
    f3 = -c16 * q[0] + c56 * q[1] + c13 * q[2];
  }
  return (f1 * w1 + f2 * w2 + f3 * w3) / (w1 + w2 + w3);
}

// s[6][nvar]: cells i-3 .. i+2 about the face, reference field order; modified in place by the
// momentum swap exactly as the reference does.  f_face[nvar] out.
EB_HD void face_flux(double (*s)[MAXVAR], int nvar, int idir, double gamma, double* f_face)
{
  double p[6], flux[6][MAXVAR], fs[5][MAXVAR], ff[MAXVAR];
  double RV[5][5], LV[5][5];
  double tmp;
  if (idir > 0)
    for (int j = 0; j < 6; j++) { tmp = s[j][1]; s[j][1] = s[j][1 + idir]; s[j][1 + idir] = tmp; }
  for (int j = 0; j < 6; j++) p[j] = eos(gamma, s[j][0], s[j][1], s[j][2], s[j][3], s[j][4]);
  const double sqL = sun_sqrt(s[2][0]);
  const double sqR = sun_sqrt(s[3][0]);
  const double sqbar = 0.5 * (sqL + sqR);
  const double u = 0.5 * (s[2][1] / sqL + s[3][1] / sqR) / sqbar;
  const double v = 0.5 * (s[2][2] / sqL + s[3][2] / sqR) / sqbar;
  const double w = 0.5 * (s[2][3] / sqL + s[3][3] / sqR) / sqbar;
  const double H = 0.5 * ((p[2] + s[2][4]) / sqL + (p[3] + s[3][4]) / sqR) / sqbar;
  const double qsq = u * u + v * v + w * w;
  const double gamm = gamma - 1.0;
  const double csnd = gamm * (H - 0.5 * qsq);
  const double cinv = 1.0 / csnd;
  for (int a = 0; a < 5; a++)
    for (int b = 0; b < 5; b++) { RV[a][b] = 0.0; LV[a][b] = 0.0; }
  RV[0][0] = 1.0;            RV[0][3] = 1.0;        RV[0][4] = 1.0;
  RV[1][0] = u - csnd;       RV[1][3] = u;          RV[1][4] = u + csnd;
  RV[2][0] = v; RV[2][1] = 1.0; RV[2][3] = v;       RV[2][4] = v;
  RV[3][0] = w; RV[3][2] = 1.0; RV[3][3] = w;       RV[3][4] = w;
  RV[4][0] = H - u * csnd; RV[4][1] = v; RV[4][2] = w; RV[4][3] = 0.5 * qsq; RV[4][4] = H + u * csnd;
  LV[0][0] = 0.5 * cinv * (u + 0.5 * gamm * qsq);
  LV[0][1] = -0.5 * cinv * (gamm * u + 1.0);
  LV[0][2] = -0.5 * v * gamm * cinv;
  LV[0][3] = -0.5 * w * gamm * cinv;
  LV[0][4] = 0.5 * gamm * cinv;
  LV[1][0] = -v;  LV[1][2] = 1.0;
  LV[2][0] = -w;  LV[2][3] = 1.0;
  LV[3][0] = -gamm * cinv * (qsq - H);
  LV[3][1] = u * gamm * cinv;
  LV[3][2] = v * gamm * cinv;
  LV[3][3] = w * gamm * cinv;
  LV[3][4] = -gamm * cinv;
  LV[4][0] = -0.5 * cinv * (u - 0.5 * gamm * qsq);
  LV[4][1] = -0.5 * cinv * (gamm * u - 1.0);
  LV[4][2] = -0.5 * v * gamm * cinv;
  LV[4][3] = -0.5 * w * gamm * cinv;
  LV[4][4] = 0.5 * gamm * cinv;

  double alpha = 0.0;
  for (int j = 0; j < 6; j++) {
    const double uj = s[j][1] / s[j][0];
    flux[j][0] = s[j][1];
    flux[j][1] = uj * s[j][1] + p[j];
    flux[j][2] = uj * s[j][2];
    flux[j][3] = uj * s[j][3];
    flux[j][4] = uj * (s[j][4] + p[j]);
    for (int c = 5; c < nvar; c++) flux[j][c] = uj * s[j][c];
    const double cj = sun_sqrt(gamma * p[j] / s[j][0]);
    tmp = fabs(uj) + cj;
    alpha = (alpha < tmp) ? tmp : alpha;
  }
  for (int pass = 0; pass < 2; pass++) {       // f+ on points 0..4, f- on points 1..5
    for (int j = 0; j < 5; j++)
      for (int c = 0; c < nvar; c++)
        fs[j][c] = pass == 0 ? 0.5 * (flux[j][c] + alpha * s[j][c]) : 0.5 * (flux[j + 1][c] - alpha * s[j + 1][c]);
    for (int c = 0; c < nvar; c++) {
      double q[5];
      for (int j = 0; j < 5; j++)
        q[j] = (c < 5) ? LV[c][0] * fs[j][0] + LV[c][1] * fs[j][1] + LV[c][2] * fs[j][2] + LV[c][3] * fs[j][3] + LV[c][4] * fs[j][4]
                       : fs[j][c];
      if (pass == 0) ff[c] = weno5(q, true);
      else ff[c] += weno5(q, false);
    }
  }
  for (int c = 0; c < 5; c++)
    f_face[c] = RV[c][0] * ff[0] + RV[c][1] * ff[1] + RV[c][2] * ff[2] + RV[c][3] * ff[3] + RV[c][4] * ff[4];
  for (int c = 5; c < nvar; c++) f_face[c] = ff[c];
  if (idir > 0) { tmp = f_face[1]; f_face[1] = f_face[1 + idir]; f_face[1 + idir] = tmp; }
}

}  // namespace strict
}  // namespace eb
