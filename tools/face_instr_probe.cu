#include "euler_math.cuh"
using namespace eb;
__global__ void ffk(const double* in, double* out, double gamma)
{
  FluidStencil s;
  const double* p = in + threadIdx.x * 50;
  for (int l = 0; l < 6; l++) { s.r[l] = p[l]; s.mn[l] = p[6 + l]; s.m1[l] = p[12 + l]; s.m2[l] = p[18 + l]; s.e[l] = p[24 + l]; s.rinv[l] = p[30 + l]; s.p[l] = p[36 + l]; s.c[l] = p[42 + l]; }
  s.srL = p[48]; s.srR = p[49];
  double f[5], alpha, u[6];
  fluid_face(s, gamma, f, alpha, u);
  double* o = out + threadIdx.x * 12;
  for (int v = 0; v < 5; v++) o[v] = f[v];
  o[5] = alpha;
  for (int l = 0; l < 6; l++) o[6 + l] = u[l];
}
__global__ void tfk(const double* in, double* out)
{
  const double* p = in + threadIdx.x * 18;
  double c[6], up[6], um[6];
  for (int l = 0; l < 6; l++) { c[l] = p[l]; up[l] = p[6 + l]; um[l] = p[12 + l]; }
  out[threadIdx.x] = tracer_face(c, up, um);
}
