/* C99 client of the C ABI (no Python, no C++): what a cgo / JNI / Rust FFI binding sees.  Built and run by
 * tests/test_host_logic.py::test_c_client_links_and_runs_host_entry_points.  Exercises only host entry points; the
 * device ones must fail loudly without a GPU (no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/zkcert_cuda.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  int expect_gpu = argc > 1 && strcmp(argv[1], "--gpu") == 0;
  CHECK(strstr(zkc_version(), "sm_100a") != NULL);
  /* G2 arithmetic and the pairing entry point */
  zkc_g2_affine g2, g2x2;
  CHECK(zkc_g2_generator(&g2) == ZKC_OK);
  uint8_t seed[32] = {0};
  zkc_fr s[1];
  CHECK(zkc_rng_fr_random(seed, 0, 0, s, 1) == ZKC_OK);                 /* the gen_srs secret: Fr::random of ChaCha20Rng([0; 32]) */
  CHECK(zkc_g2_mul(&g2, &s[0], &g2x2) == ZKC_OK);                       /* s_g2 = [s]G2 */
  CHECK(memcmp(&g2, &g2x2, sizeof g2) != 0);
  int one = -1;
  zkc_g1_affine id;
  memset(&id, 0, sizeof id);
  CHECK(zkc_pairing_check(&id, &g2, 1, &one) == ZKC_OK && one == 1);    /* e(O, G2) = 1 */
  /* params file codec: size law and rejection of a truncated file */
  CHECK(zkc_params_size(3) == 4 + 2 * 64 * 8 + 256);
  uint32_t k = 99;
  uint8_t junk[16] = {3, 0, 0, 0};
  CHECK(zkc_params_read(junk, sizeof junk, 1, &k, NULL, NULL, NULL, NULL) != ZKC_OK);
  /* hashes */
  uint8_t dig[64];
  CHECK(zkc_host_hash(1, NULL, (const uint8_t*)"", 0, dig) == 0 && dig[0] == 0xc5 && dig[1] == 0xd2);
  /* partition arithmetic */
  uint64_t lo, hi;
  CHECK(zkc_team_shard_range(10, 3, 2, &lo, &hi) == ZKC_OK && lo == 7 && hi == 10);
  /* device entry points: a context only exists with a GPU */
  zkc_ctx* ctx = NULL;
  int st = zkc_ctx_create(0, &ctx);
  if (expect_gpu) { CHECK(st == ZKC_OK && ctx != NULL); zkc_ctx_destroy(ctx); }
  else { CHECK(st != ZKC_OK && ctx == NULL); CHECK(strlen(zkc_last_error(NULL)) > 0); }
  printf("cabi_host ok\n");
  return 0;
}
