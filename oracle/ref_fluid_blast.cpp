/* ---------------------------------------------------------------------------
 * ref_fluid_blast.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * C-callable harness around the UNMODIFIED initial_conditions() of the reference's
 * fluid_blast.cpp (compiled from /root/reference/src with its link-time hooks renamed on the
 * command line, see oracle/Makefile), used once by tests/golden/make_golden.py to produce the
 * golden initial state that pins problems.py's restatement (clump positions come from
 * std::mt19937_64 seeded with the rank count, fluid_blast.cpp:102-128).
 * ------------------------------------------------------------------------- */
#include <euler3D.hpp>

void shim_set_world(int nprocs);
void shim_set_rank(int rank);

extern "C" int refdrv_fluid_blast_ic(const long* n, double mass_units, double length_units, double time_units,
                                     double gamma, double* const* w)
{
  shim_set_world(1); shim_set_rank(0);
  EulerData u;
  u.nx = n[0]; u.ny = n[1]; u.nz = n[2];
  u.xlbc = u.xrbc = u.ylbc = u.yrbc = u.zlbc = u.zrbc = BC_REFLECTING;
  u.gamma = gamma;
  u.MassUnits = mass_units; u.LengthUnits = length_units; u.TimeUnits = time_units;
  if (u.UpdateUnits() != 0) return -1;
  if (u.SetupDecomp() != 0) return -2;
  const long N = u.nxl * u.nyl * u.nzl;
  N_Vector sub[5];
  for (int f = 0; f < 5; f++) sub[f] = N_VMake_Serial(N, w[f], u.ctx);
  N_Vector wv = N_VMake_MPIManyVector(u.comm, 5, sub, u.ctx);
  const realtype t0 = 0.0;
  const int ret = initial_conditions(t0, wv, u);
  N_VDestroy(wv);
  for (int f = 0; f < 5; f++) N_VDestroy(sub[f]);
  return ret;
}

/* Stand-alone form (the reference prints through std::cout, which does not survive being
 * driven from inside the Python process): writes n[0]*n[1]*n[2]*5 doubles to argv[4]. */
#ifdef FB_MAIN
#include <cstdio>
int main(int argc, char** argv)
{
  if (argc < 5) return 2;
  long n[3] = {atol(argv[1]), atol(argv[2]), atol(argv[3])};
  const long N = n[0] * n[1] * n[2];
  std::vector<double> buf(5 * N);
  double* w[5];
  for (int f = 0; f < 5; f++) w[f] = buf.data() + f * N;
  const int ret = refdrv_fluid_blast_ic(n, 3.0e70, 3.0857e30, 1.0e12, 5.0 / 3.0, w);
  FILE* fp = fopen(argv[4], "wb");
  if (!fp) return 3;
  fwrite(buf.data(), sizeof(double), buf.size(), fp);
  fclose(fp);
  return ret;
}
#endif
