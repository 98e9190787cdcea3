/* ---------------------------------------------------------------------------
 * euler_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the fluid right-hand side of sundials-manyvector-demo
 * (the fEuler ARKRhsFn) used ONLY as the parity checker for the CUDA path:
 * it may be imported/linked from tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, never from the product.
 *
 * Parity status: PINNED.  This file is checked (tests/test_oracle.py) against
 *   (a) the unmodified reference compiled into oracle/_ref/ (bit-for-bit on every
 *       boundary-condition type, every NVAR the reference builds, x/y/z variants), and
 *   (b) golden vectors generated from (a) and committed under tests/golden/.
 * The operation order below deliberately follows the reference expression by
 * expression so that, without FMA contraction, results are bit-identical.
 *
 * Reference lines restated (all under /root/reference/src/):
 *   fEuler      utilities.cpp:17-253        face_flux  utilities.cpp:270-479
 *   stability   utilities.cpp:483-528       eos / legal_state  euler3D.hpp:1383-1414
 *   halo pack   euler3D.hpp:644-786         BC ghost fill      euler3D.hpp:797-1166
 *   stencil gather (pack1D_*_bdry)          euler3D.hpp:1250-1378
 * ------------------------------------------------------------------------- */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORA_MAXVAR 32

enum { ORA_BC_PERIODIC = 0, ORA_BC_NEUMANN = 1, ORA_BC_DIRICHLET = 2, ORA_BC_REFLECTING = 3,
       ORA_BC_EXTERNAL = -1 /* ghost layers supplied by the caller (a neighbouring rank) */ };

typedef struct {
  long nxl, nyl, nzl;      /* local extents (euler3D.hpp:198-200) */
  int nchem;               /* NVAR - 5 (euler3D.hpp:297) */
  double dx, dy, dz;       /* euler3D.hpp:209-211 */
  double gamma;            /* euler3D.hpp:221 */
  int bc[6];               /* W,E,S,N,B,F: BC code, or ORA_BC_EXTERNAL */
  double forcing[5];       /* constant forcing assigned into wdot (external_forces hook) */
} oracle_cfg;

/* SUNRsqrt semantics of SUNDIALS 6.2: non-positive -> 0 (used at utilities.cpp:298-299,378,511) */
static double rsqrt_sun(double x) { return (x <= 0.0) ? 0.0 : sqrt(x); }

/* euler3D.hpp:1383-1388 */
static double eos(double gamma, double rho, double mx, double my, double mz, double et)
{
  return (gamma - 1.0) * (et - (mx * mx + my * my + mz * mz) * 0.5 / rho);
}

/* euler3D.hpp:1405-1414 */
static int legal_state(double gamma, double rho, double mx, double my, double mz, double et)
{
  int d = (rho > 0.0) ? 0 : 1;
  int e = (et > 0.0) ? 0 : 2;
  int p = (eos(gamma, rho, mx, my, mz, et) > 0.0) ? 0 : 4;
  return d + e + p;
}

/* One fifth-order WENO reconstruction at the face from five projected values
 * q[0..4]; `plus` selects the f+ (left-biased, utilities.cpp:399-423) or f-
 * (right-biased, utilities.cpp:443-467) weights and stencil coefficients. */
static double weno5(const double q[5], int plus)
{
  const double bc = 1.083333333333333333333333333333333333333;   /* 13/12 */
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

/* utilities.cpp:270-479.  s is the 6-point stencil [6][nvar] (cells i-3..i+2 about
 * face i-1/2); it is modified in place exactly as the reference does (momentum swap). */
void oracle_face_flux(double* s, int nvar, int idir, double gamma, double* f_face)
{
  double p[6], flux[6][ORA_MAXVAR], fs[5][ORA_MAXVAR], fp[5][ORA_MAXVAR], ff[ORA_MAXVAR];
  double RV[5][5], LV[5][5];
  double sqL, sqR, sqbar, u, v, w, H, qsq, gamm, csnd, cinv, alpha, tmp;
  int j, c;
#define S(j, c) s[(j) * nvar + (c)]

  /* :283-285 rotate so that the sweep direction is "x" */
  if (idir > 0)
    for (j = 0; j < 6; j++) { tmp = S(j, 1); S(j, 1) = S(j, 1 + idir); S(j, 1 + idir) = tmp; }

  /* :288-289 */
  for (j = 0; j < 6; j++) p[j] = eos(gamma, S(j, 0), S(j, 1), S(j, 2), S(j, 3), S(j, 4));

  /* :298-304 Roe average from the two cells adjacent to the face */
  sqL = rsqrt_sun(S(2, 0));
  sqR = rsqrt_sun(S(3, 0));
  sqbar = 0.5 * (sqL + sqR);
  u = 0.5 * (S(2, 1) / sqL + S(3, 1) / sqR) / sqbar;
  v = 0.5 * (S(2, 2) / sqL + S(3, 2) / sqR) / sqbar;
  w = 0.5 * (S(2, 3) / sqL + S(3, 3) / sqR) / sqbar;
  H = 0.5 * ((p[2] + S(2, 4)) / sqL + (p[3] + S(3, 4)) / sqR) / sqbar;

  /* :307-364 eigenvector matrices; note csnd here is c^2 (no square root), as in the reference */
  qsq = u * u + v * v + w * w;
  gamm = gamma - 1.0;
  csnd = gamm * (H - 0.5 * qsq);
  cinv = 1.0 / csnd;
  memset(RV, 0, sizeof RV);
  memset(LV, 0, sizeof LV);
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

  /* :368-380 physical fluxes and the face-local (6-point) maximum wave speed */
  alpha = 0.0;
  for (j = 0; j < 6; j++) {
    double uj = S(j, 1) / S(j, 0);
    double cj;
    flux[j][0] = S(j, 1);
    flux[j][1] = uj * S(j, 1) + p[j];
    flux[j][2] = uj * S(j, 2);
    flux[j][3] = uj * S(j, 3);
    flux[j][4] = uj * (S(j, 4) + p[j]);
    for (c = 5; c < nvar; c++) flux[j][c] = uj * S(j, c);
    cj = rsqrt_sun(gamma * p[j] / S(j, 0));
    tmp = fabs(uj) + cj;
    alpha = (alpha < tmp) ? tmp : alpha;      /* std::max(alpha, tmp) */
  }

  /* :386-423 f+ : Lax-Friedrichs split on points 0..4, project fluid rows, WENO */
  for (j = 0; j < 5; j++)
    for (c = 0; c < nvar; c++) fs[j][c] = 0.5 * (flux[j][c] + alpha * S(j, c));
  for (j = 0; j < 5; j++) {
    for (c = 0; c < 5; c++)
      fp[j][c] = LV[c][0] * fs[j][0] + LV[c][1] * fs[j][1] + LV[c][2] * fs[j][2]
               + LV[c][3] * fs[j][3] + LV[c][4] * fs[j][4];
    for (c = 5; c < nvar; c++) fp[j][c] = fs[j][c];      /* tracers: identity projection (:395) */
  }
  for (c = 0; c < nvar; c++) {
    double q[5];
    for (j = 0; j < 5; j++) q[j] = fp[j][c];
    ff[c] = weno5(q, 1);
  }

  /* :429-467 f- : split on points 1..5 */
  for (j = 0; j < 5; j++)
    for (c = 0; c < nvar; c++) fs[j][c] = 0.5 * (flux[j + 1][c] - alpha * S(j + 1, c));
  for (j = 0; j < 5; j++) {
    for (c = 0; c < 5; c++)
      fp[j][c] = LV[c][0] * fs[j][0] + LV[c][1] * fs[j][1] + LV[c][2] * fs[j][2]
               + LV[c][3] * fs[j][3] + LV[c][4] * fs[j][4];
    for (c = 5; c < nvar; c++) fp[j][c] = fs[j][c];
  }
  for (c = 0; c < nvar; c++) {
    double q[5];
    for (j = 0; j < 5; j++) q[j] = fp[j][c];
    ff[c] += weno5(q, 0);
  }

  /* :470-473 back to conserved variables */
  for (c = 0; c < 5; c++)
    f_face[c] = RV[c][0] * ff[0] + RV[c][1] * ff[1] + RV[c][2] * ff[2] + RV[c][3] * ff[3] + RV[c][4] * ff[4];
  for (c = 5; c < nvar; c++) f_face[c] = ff[c];

  /* :476-477 */
  if (idir > 0) { tmp = f_face[1]; f_face[1] = f_face[1 + idir]; f_face[1 + idir] = tmp; }
#undef S
}

/* ------------------------------ ghost layers ------------------------------ */
/* All six ghost buffers use the reference's receive-buffer layout
 * (euler3D.hpp:648,696,744): value (v, d, a, b) at v + nvar*(d + 3*(a + na*b)) with
 *   W/E: a=j, b=k, na=nyl;   S/N: a=i, b=k, na=nxl;   B/F: a=i, b=j, na=nxl. */

static long face_len(const oracle_cfg* c, int f)
{
  long nv = 5 + c->nchem;
  if (f < 2) return nv * 3 * c->nyl * c->nzl;
  if (f < 4) return nv * 3 * c->nxl * c->nzl;
  return nv * 3 * c->nxl * c->nyl;
}
long oracle_face_len(const oracle_cfg* c, int f) { return face_len(c, f); }

static double cell_value(const oracle_cfg* c, const double* const* w, int v, long i, long j, long k)
{
  long cell = i + c->nxl * (j + c->nyl * k);
  return (v < 5) ? w[v][cell] : w[5][(v - 5) + c->nchem * cell];
}

/* Map (face f, layer d, tangential a, b) to the owned cell (i,j,k) that layer `src`
 * along the face normal refers to. */
static void face_cell(const oracle_cfg* c, int f, long src, long a, long b, long* i, long* j, long* k)
{
  (void)c;
  if (f < 2)      { *i = src; *j = a; *k = b; }
  else if (f < 4) { *i = a; *j = src; *k = b; }
  else            { *i = a; *j = b; *k = src; }
}

/* What a rank SENDS through face f (euler3D.hpp:644-786): its three layers nearest
 * that face, in increasing index order. */
void oracle_pack_send(const oracle_cfg* c, const double* const* w, int f, double* buf)
{
  long nv = 5 + c->nchem, na, nb, n, a, b, i, j, k;
  int d, v;
  if (f < 2) { na = c->nyl; nb = c->nzl; n = c->nxl; }
  else if (f < 4) { na = c->nxl; nb = c->nzl; n = c->nyl; }
  else { na = c->nxl; nb = c->nyl; n = c->nzl; }
  for (b = 0; b < nb; b++)
    for (a = 0; a < na; a++)
      for (d = 0; d < 3; d++) {
        long src = (f % 2 == 0) ? d : n - 3 + d;
        face_cell(c, f, src, a, b, &i, &j, &k);
        for (v = 0; v < nv; v++) buf[v + nv * (d + 3 * (a + na * b))] = cell_value(c, w, v, i, j, k);
      }
}

/* Ghost layers of face f for a physical boundary (euler3D.hpp:797-1166), or, for a
 * periodic face on a single rank, the wrap-around copy.  Low side: ghost[d] = own[2-d]
 * (mirror).  High side: ghost[d] = own[n-3+d] (plain copy -- as the reference does).
 * Reflecting negates the face-normal momentum; Dirichlet negates everything. */
void oracle_fill_ghost(const oracle_cfg* c, const double* const* w, int f, double* buf)
{
  long nv = 5 + c->nchem, na, nb, n, a, b, i, j, k;
  int d, v, bc = c->bc[f], normal = 1 + f / 2;
  if (f < 2) { na = c->nyl; nb = c->nzl; n = c->nxl; }
  else if (f < 4) { na = c->nxl; nb = c->nzl; n = c->nyl; }
  else { na = c->nxl; nb = c->nyl; n = c->nzl; }
  for (b = 0; b < nb; b++)
    for (a = 0; a < na; a++)
      for (d = 0; d < 3; d++) {
        long src;
        if (bc == ORA_BC_PERIODIC) src = (f % 2 == 0) ? n - 3 + d : d;
        else src = (f % 2 == 0) ? 2 - d : n - 3 + d;
        face_cell(c, f, src, a, b, &i, &j, &k);
        for (v = 0; v < nv; v++) {
          double x = cell_value(c, w, v, i, j, k);
          if (bc == ORA_BC_DIRICHLET) x = -x;
          else if (bc == ORA_BC_REFLECTING && v == normal) x = -x;
          buf[v + nv * (d + 3 * (a + na * b))] = x;
        }
      }
}

/* pack1D_{x,y,z}_bdry (euler3D.hpp:1250-1378): gather cells fidx-3..fidx+2 along `dir`
 * for the face whose index along dir is fidx in [0, n]; entries outside the owned range
 * come from the low/high ghost buffer of that direction. */
static void gather_stencil(const oracle_cfg* c, const double* const* w, double* const* ghost,
                           int dir, long i, long j, long k, double* s)
{
  long nv = 5 + c->nchem, n, fidx, a, b, na, l;
  int v;
  if (dir == 0) { n = c->nxl; fidx = i; a = j; b = k; na = c->nyl; }
  else if (dir == 1) { n = c->nyl; fidx = j; a = i; b = k; na = c->nxl; }
  else { n = c->nzl; fidx = k; a = i; b = j; na = c->nxl; }
  for (l = 0; l < 6; l++) {
    long pos = fidx - 3 + l;
    if (pos < 0) {
      const double* g = ghost[2 * dir];
      for (v = 0; v < nv; v++) s[l * nv + v] = g[v + nv * ((pos + 3) + 3 * (a + na * b))];
    } else if (pos >= n) {
      const double* g = ghost[2 * dir + 1];
      for (v = 0; v < nv; v++) s[l * nv + v] = g[v + nv * ((pos - n) + 3 * (a + na * b))];
    } else {
      long ii = i, jj = j, kk = k;
      if (dir == 0) ii = pos; else if (dir == 1) jj = pos; else kk = pos;
      for (v = 0; v < nv; v++) s[l * nv + v] = cell_value(c, w, v, ii, jj, kk);
    }
  }
}

/* utilities.cpp:17-253.  w[0..4] fluid SoA, w[5] chem AoS (or NULL); same for wdot.
 * ext[f] supplies the ghost layers of faces whose bc[f] == ORA_BC_EXTERNAL (what the
 * reference would have received from the neighbouring rank); other entries are ignored.
 * Returns 0, or -1 if any owned cell fails legal_state (utilities.cpp:83-84,132-133);
 * *state_mask (may be NULL) receives the OR of the 1/2/4 failure bits. */
int oracle_feuler(const oracle_cfg* c, const double* const* w, double* const* wdot,
                  const double* const* ext, int* state_mask)
{
  const long nx = c->nxl, ny = c->nyl, nz = c->nzl, N = nx * ny * nz;
  const int nv = 5 + c->nchem;
  double* ghost[6];
  double *xf, *yf, *zf, s[6 * ORA_MAXVAR];
  long i, j, k, cell;
  int f, v, mask = 0;

  for (f = 0; f < 6; f++) {
    ghost[f] = (double*)malloc(sizeof(double) * face_len(c, f));
    if (c->bc[f] == ORA_BC_EXTERNAL) memcpy(ghost[f], ext[f], sizeof(double) * face_len(c, f));
    else oracle_fill_ghost(c, w, f, ghost[f]);
  }

  /* :28 zero, :65 forcing assigned into wdot */
  for (v = 0; v < 5; v++) for (cell = 0; cell < N; cell++) wdot[v][cell] = c->forcing[v];
  if (c->nchem > 0) for (cell = 0; cell < N * c->nchem; cell++) wdot[5][cell] = 0.0;

  /* legal_state on every owned cell */
  for (cell = 0; cell < N; cell++)
    mask |= legal_state(c->gamma, w[0][cell], w[1][cell], w[2][cell], w[3][cell], w[4][cell]);
  if (state_mask) *state_mask = mask;
  if (mask) { for (f = 0; f < 6; f++) free(ghost[f]); return -1; }

  /* :76-195 face fluxes: lower face of every cell, plus the upper face of the last cell */
  xf = (double*)malloc(sizeof(double) * nv * (nx + 1) * ny * nz);
  yf = (double*)malloc(sizeof(double) * nv * nx * (ny + 1) * nz);
  zf = (double*)malloc(sizeof(double) * nv * nx * ny * (nz + 1));
  for (k = 0; k < nz; k++)
    for (j = 0; j < ny; j++)
      for (i = 0; i <= nx; i++) {
        gather_stencil(c, w, ghost, 0, i, j, k, s);
        oracle_face_flux(s, nv, 0, c->gamma, &xf[nv * (i + (nx + 1) * (j + ny * k))]);
      }
  for (k = 0; k < nz; k++)
    for (j = 0; j <= ny; j++)
      for (i = 0; i < nx; i++) {
        gather_stencil(c, w, ghost, 1, i, j, k, s);
        oracle_face_flux(s, nv, 1, c->gamma, &yf[nv * (i + nx * (j + (ny + 1) * k))]);
      }
  for (k = 0; k <= nz; k++)
    for (j = 0; j < ny; j++)
      for (i = 0; i < nx; i++) {
        gather_stencil(c, w, ghost, 2, i, j, k, s);
        oracle_face_flux(s, nv, 2, c->gamma, &zf[nv * (i + nx * (j + ny * k))]);
      }

  /* :198-245 flux divergence */
  for (k = 0; k < nz; k++)
    for (j = 0; j < ny; j++)
      for (i = 0; i < nx; i++) {
        cell = i + nx * (j + ny * k);
        for (v = 0; v < nv; v++) {
          double div = (xf[v + nv * ((i + 1) + (nx + 1) * (j + ny * k))] - xf[v + nv * (i + (nx + 1) * (j + ny * k))]) / c->dx
                     + (yf[v + nv * (i + nx * ((j + 1) + (ny + 1) * k))] - yf[v + nv * (i + nx * (j + (ny + 1) * k))]) / c->dy
                     + (zf[v + nv * (i + nx * (j + ny * (k + 1)))] - zf[v + nv * (i + nx * (j + ny * k))]) / c->dz;
          if (v < 5) wdot[v][cell] -= div;
          else wdot[5][(v - 5) + c->nchem * cell] -= div;
        }
      }

  free(xf); free(yf); free(zf);
  for (f = 0; f < 6; f++) free(ghost[f]);
  return 0;
}

/* Local part of utilities.cpp:505-513: max over owned cells of | |mx/rho| + c |.
 * (The reference takes the max of |mx/rho| three times; my, mz never enter.) */
double oracle_max_wavespeed(const oracle_cfg* c, const double* const* w)
{
  const long N = c->nxl * c->nyl * c->nzl;
  double alpha = 0.0;
  long i;
  for (i = 0; i < N; i++) {
    double u = fabs(w[1][i] / w[0][i]);
    double p = eos(c->gamma, w[0][i], w[1][i], w[2][i], w[3][i], w[4][i]);
    double cs = rsqrt_sun(c->gamma * p / w[0][i]);
    double a = fabs(u + cs);
    alpha = (alpha < a) ? a : alpha;
  }
  return alpha;
}

/* utilities.cpp:520 given the (globally reduced) alpha */
double oracle_dt_stab(const oracle_cfg* c, double cfl, double alpha)
{
  double h = (c->dx < c->dy) ? c->dx : c->dy;
  h = (h < c->dz) ? h : c->dz;
  return cfl * h / alpha;
}
