// Test harness (not product code): builds the initial state of a problem with the native driver's
// host/problems.hpp, writes it as a solution file, reads it back and writes the copy, so that the
// python tests can compare both files with the golden fixtures and with problems.py.
//   native_problems_check <problem> <nx> <ny> <nz> <nchem> <out1> <out2>
#include "../sundials-manyvector-demo_b200/host/problems.hpp"

int main(int argc, char** argv)
{
  if (argc < 8) return 2;
  eb_problems::Problem P;
  P.name = argv[1];
  P.nx = atol(argv[2]); P.ny = atol(argv[3]); P.nz = atol(argv[4]);
  P.nchem = atoi(argv[5]);
  P.xl = P.yl = P.zl = 0.0; P.xr = P.yr = P.zr = 1.0; P.gamma = 1.4;
  if (P.is_blast()) {          // tests/fluid_blast/input_fluid_blast.txt, tests/primordial_blast/input_*
    P.gamma = 5.0 / 3.0; P.MassUnits = 3.0e70; P.LengthUnits = 3.0857e30;
    P.TimeUnits = P.name == "fluid_blast" ? 1.0e12 : 1.0e11;
  } else if (P.name.compare(0, 9, "hurricane") == 0) {
    P.xl = P.yl = P.zl = -1.0; P.gamma = 2.0;
  } else if (P.name == "rayleigh_taylor") {
    P.xl = -0.25; P.xr = 0.25; P.yl = -0.75; P.yr = 0.75;
  }
  const size_t N = (size_t)(P.nx * P.ny * P.nz);
  std::vector<std::vector<double>> a(5, std::vector<double>(N)), b(5, std::vector<double>(N));
  std::vector<double> ca(N * P.nchem), cb(N * P.nchem);
  double* const fa[5] = {a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data()};
  double* const fb[5] = {b[0].data(), b[1].data(), b[2].data(), b[3].data(), b[4].data()};
  bool analytic = false;
  if (eb_problems::initial_conditions(P, 0.125, fa, ca.data(), &analytic) != 0) return 1;
  if (eb_problems::write_solution(argv[6], P, 0.125, fa, ca.data()) != 0) return 1;
  double t = -1.0;
  if (eb_problems::read_solution(argv[6], P, &t, fb, cb.data()) != 0 || t != 0.125) return 1;
  if (eb_problems::write_solution(argv[7], P, t, fb, cb.data()) != 0) return 1;
  printf("analytic=%d\n", analytic ? 1 : 0);
  return 0;
}
