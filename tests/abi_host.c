/* A plain-C99 host of the C ABI (tests/test_abi_from_c.py compiles this with gcc -std=c99 -pedantic and runs it):
 * include/amc_b200.h must be valid C, and the library must be callable with plain pointers and PODs -- no torch, no C++.
 * Only host-side entry points are exercised (no GPU needed): ABI version, parameter layout, error reporting. */
#include <stdio.h>
#include <string.h>
#include "amc_b200.h"

int main(void) {
  AmcDesc d;
  AmcParamLayout L;
  memset(&d, 0, sizeof d);
  printf("abi %d\n", amc_abi_version());
  d.kind = AMC_KIND_RAWIQ; d.dtype = AMC_BF16; d.B = 4; d.d = 128; d.h = 8; d.F = 1024; d.C = 11; d.n_layers = 6;
  d.in_ch = 2; d.seq_len = 1024; d.seg = 16; d.has_cls = 1; d.head_ln = 1;
  if (amc_param_layout(&d, &L) != 0) { printf("error: %s\n", amc_last_error()); return 1; }
  printf("T %d total %lld\n", (int)L.T, (long long)L.total);
  d.h = 7;
  if (amc_param_layout(&d, &L) == 0) return 2;
  printf("expected error: %s\n", amc_last_error());
  return 0;
}
