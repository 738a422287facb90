/* Minimal C client of libpds_b200 through the host layer of include/pdsb.h — what a cgo / JNI / Rust-FFI binding of the
 * reference's `pl_lr` would do: hand over Arrow-style column buffers, get coefficients back.
 *
 *   gcc -std=c99 -Iinclude examples/c_client.c polars_ds_extension_b200/_polars_ds_b200.so \
 *       -Wl,-rpath,$PWD/polars_ds_extension_b200 -o /tmp/c_client && /tmp/c_client
 *
 * Prints the fitted coefficients of y = 2 x1 - 0.5 x2 + 1 (needs a CUDA device; without one it prints the library's
 * error string and exits 2 — there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pdsb.h"

int main(void) {
  enum { N = 1000 };
  double *x1 = malloc(sizeof(double) * N), *x2 = malloc(sizeof(double) * N), *y = malloc(sizeof(double) * N);
  unsigned s = 12345u;
  for (int i = 0; i < N; ++i) {
    s = s * 1664525u + 1013904223u; x1[i] = (double)(s >> 8) / (1u << 24);
    s = s * 1664525u + 1013904223u; x2[i] = (double)(s >> 8) / (1u << 24);
    y[i] = 2.0 * x1[i] - 0.5 * x2[i] + 1.0;
  }
  pdsb_chunk cy = {y, NULL, 0, N}, c1 = {x1, NULL, 0, N}, c2 = {x2, NULL, 0, N};
  pdsb_column cols[3] = {{"y", PDSB_F64, 1, &cy, 0}, {"x1", PDSB_F64, 1, &c1, 0}, {"x2", PDSB_F64, 1, &c2, 0}};
  pdsb_lr_kwargs kw;
  memset(&kw, 0, sizeof kw);
  kw.bias = 1; kw.null_policy = "raise"; kw.solver = "qr"; kw.tol = 1e-5; kw.std_err = "se";
  kw.max_iter = 200; kw.singular_x_tol = 1e-12;
  pdsb_host_result out;
  memset(&out, 0, sizeof out);
  int rc = pdsb_host_lin_reg(cols, 3, &kw, /*f32*/ 0, /*n_targets*/ 1, /*want_pred*/ 0, /*w_rcond*/ 0, &out);
  if (rc != 0) {
    fprintf(stderr, "pdsb_host_lin_reg failed: %s\n", pdsb_last_error());
    return 2;
  }
  const double* b = (const double*)out.coeffs;
  printf("coeffs = [%.9f, %.9f, %.9f] (gated = %d)\n", b[0], b[1], b[2], out.gated);
  int ok = out.n_coef == 3 && !out.gated;
  const double want[3] = {2.0, -0.5, 1.0};
  for (int i = 0; i < 3 && ok; ++i) ok = (b[i] - want[i] < 1e-8) && (want[i] - b[i] < 1e-8);
  pdsb_host_result_free(&out);
  free(x1); free(x2); free(y);
  return ok ? 0 : 1;
}
