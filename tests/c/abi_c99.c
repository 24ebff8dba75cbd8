/* include/bliss_b200.h is a C header: this C99 program (gcc -std=c99 -pedantic) links against libbliss_b200.so and
 * uses the entry points that need no device; on a box without a GPU bliss_b200_init must fail with
 * BLISS_B200_E_NO_DEVICE -- there is no CPU fallback.  Built and run by tests/test_host_abi.py. */
#include <stdio.h>
#include <string.h>

#include "bliss_b200.h"

int main(void) {
    float w[23 * 23];
    int i, rc;
    if (bliss_b200_feature_count(2) != 23 || bliss_b200_feature_count(1) != 20 || bliss_b200_feature_count(9) != 0) return 2;
    if (bliss_b200_feature_weights(2, w) != BLISS_B200_OK) return 3;
    if (w[0] != 0.25f || w[24] != 1.0f || w[1] != 0.0f) return 4; /* VERSION2_WEIGHTS, src/lib.rs:209-234 */
    for (i = 10; i < 23; i++)
        if (w[i * 23 + i] != 3.0f / 13.0f) return 5;
    if (strcmp(bliss_b200_strerror(BLISS_B200_SONG_TOO_SHORT), "empty or too short song.") != 0) return 6; /* src/song/mod.rs:426-430 */
    rc = bliss_b200_init(0);
    if (rc == BLISS_B200_OK) {
        float out[23];
        static float pcm[22050];
        rc = bliss_b200_analyze(pcm, 4000, 2, out); /* too short: a per-song status, not a call failure */
        if (rc != BLISS_B200_SONG_TOO_SHORT) return 7;
        puts("DEVICE_OK");
        return 0;
    }
    if (rc != BLISS_B200_E_NO_DEVICE) return 8;
    printf("NO_DEVICE: %s\n", bliss_b200_last_error());
    return 0;
}
