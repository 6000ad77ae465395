/* Minimal C caller of the drop-in boundary (include/hades_cuda.h): the README example of the reference
 * (README.md:50-65: permute [1; WIDTH] and check that it changed), batched.
 * Build: gcc examples/perm_batch.c -Iinclude -Lhades252_b200/lib -lhades_b200 -Wl,-rpath,$PWD/hades252_b200/lib -o perm_batch */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hades_constants.h"
#include "hades_cuda.h"

int main(void) {
    /* BlsScalar::from(1u64) in memory: R mod p */
    const uint64_t one[4] = {0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};
    const size_t n = 1 << 16;
    uint64_t* states = malloc(n * HADES_WIDTH * 4 * sizeof(uint64_t));
    for (size_t i = 0; i < n * HADES_WIDTH; i++) memcpy(states + 4 * i, one, sizeof one);

    hades_ctx* ctx = NULL;
    int dev = 0;
    int rc = hades_init(&ctx, &dev, 1, HADES_WIDTH, &HADES_ROUND_CONSTANTS[0][0], HADES_N_ROUND_CONSTANTS, &HADES_MDS_MATRIX_5[0][0]);
    if (rc) { fprintf(stderr, "hades_init: %s\n", hades_last_error(NULL)); return rc; }
    rc = hades_perm_batch(ctx, states, n);
    if (rc) { fprintf(stderr, "hades_perm_batch: %s\n", hades_last_error(ctx)); return rc; }

    int changed = memcmp(states, one, sizeof one) != 0;
    int all_equal = 1;
    for (size_t i = 1; i < n; i++) all_equal &= !memcmp(states, states + i * HADES_WIDTH * 4, HADES_WIDTH * 32);
    printf("perm([1;5])[0] limbs: %016llx %016llx %016llx %016llx  changed=%d deterministic=%d\n",
           (unsigned long long)states[0], (unsigned long long)states[1], (unsigned long long)states[2],
           (unsigned long long)states[3], changed, all_equal);
    hades_destroy(ctx);
    free(states);
    return !(changed && all_equal);
}
