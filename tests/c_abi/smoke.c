/* Plain-C consumer of include/mpvss_b200.h (what a cgo / Rust-FFI / JNI binding sees): no Python, no
 * torch.  Build: gcc -I include tests/c_abi/smoke.c -L mpvss_rs_b200 -lmpvss_b200 -o smoke
 * Checks on device 0: 2^5 and 4^3 through the fixed-base tables, a batch product, X_i = prod C_j^(i^j) for
 * C = (2, 3), positions 1..3 (2*3, 2*3^2, 2*3^3), and the scalar polynomial 7 + 5x at x = 1, 2. */
#include <stdio.h>
#include <string.h>
#include "mpvss_b200.h"

static int low64_is(const uint8_t* le256, uint64_t v) {
  for (int i = 0; i < 256; ++i)
    if (le256[i] != (i < 8 ? (uint8_t)(v >> (8 * i)) : 0)) return 0;
  return 1;
}
static void put(uint8_t* le256, uint64_t v) {
  memset(le256, 0, 256);
  for (int i = 0; i < 8; ++i) le256[i] = (uint8_t)(v >> (8 * i));
}
#define CHECK(x) do { if (!(x)) { printf("FAILED: %s (%s)\n", #x, ctx ? mpvss_last_error(ctx) : "-"); return 1; } } while (0)

int main(void) {
  mpvss_ctx* ctx = NULL;
  CHECK(mpvss_ctx_create(MPVSS_GROUP_MODP, 0, &ctx) == MPVSS_OK);
  CHECK(mpvss_element_bytes(ctx) == 256 && mpvss_scalar_bytes(ctx) == 256);
  uint8_t s[2 * 256], out[3 * 256], a[2 * 256], b[2 * 256];
  put(s, 5);
  CHECK(mpvss_fixed_base_exp(ctx, MPVSS_GEN_MAIN, s, 1, out) == MPVSS_OK && low64_is(out, 32));
  put(s, 3);
  CHECK(mpvss_fixed_base_exp(ctx, MPVSS_GEN_SUBGROUP, s, 1, out) == MPVSS_OK && low64_is(out, 64));
  put(a, 6); put(a + 256, 1000003); put(b, 7); put(b + 256, 999983);
  CHECK(mpvss_batch_mul(ctx, a, b, 2, out) == MPVSS_OK && low64_is(out, 42) && low64_is(out + 256, 1000003ull * 999983ull));
  put(a, 2); put(a + 256, 3);                       /* commitments C_0 = 2, C_1 = 3 */
  CHECK(mpvss_poly_eval_exp(ctx, a, 2, NULL, 3, out) == MPVSS_OK);
  CHECK(low64_is(out, 6) && low64_is(out + 256, 18) && low64_is(out + 512, 54));
  put(a, 7); put(a + 256, 5);                       /* P(x) = 7 + 5x */
  CHECK(mpvss_scalar_poly_eval(ctx, a, 2, NULL, 2, out) == MPVSS_OK && low64_is(out, 12) && low64_is(out + 256, 17));
  CHECK(mpvss_batch_mul(ctx, NULL, b, 2, out) == MPVSS_ERR_ARG);   /* status codes, no crash */
  CHECK(mpvss_comm_size(ctx) == 1 && mpvss_comm_rank(ctx) == 0);
  mpvss_ctx_destroy(ctx);
  printf("c abi smoke ok\n");
  return 0;
}
