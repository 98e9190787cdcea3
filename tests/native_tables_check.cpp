// Test infrastructure: prints the native driver's Butcher tables (host/erk_tables.hpp) so that
// tests/test_driver_cpu.py can compare them with driver.py's.  Usage: native_tables_check <order> <etable>
#include <cstdio>
#include <cstdlib>
#include "erk_tables.hpp"
int main(int argc, char** argv)
{
  Table T;
  if (argc < 3 || !make_table(atoi(argv[1]), atoi(argv[2]), T)) { printf("none\n"); return 0; }
  printf("%d %d %d %d\n", T.s, T.p, T.q, T.embedded ? 1 : 0);
  for (int i = 0; i < T.s; i++) { for (int j = 0; j < T.s; j++) printf("%.17g ", T.A[i][j]); printf("\n"); }
  for (int j = 0; j < T.s; j++) printf("%.17g ", T.b[j]);
  printf("\n");
  for (int j = 0; j < T.s; j++) printf("%.17g ", T.bh[j]);
  printf("\n");
  return 0;
}
