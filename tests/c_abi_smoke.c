/* Plain-C client of include/rv_b200.h: no Python, no torch -- the drop-in boundary is a C ABI.
 * Runs the chain through rv_chain_u8 on host buffers and checks it bit for bit against the CPU oracle (rv_oracle.c).
 *   gcc -O1 -o c_abi_smoke tests/c_abi_smoke.c -Iinclude -Lroad-vision-system_b200/csrc -lrv_b200 -Loracle -lrv_oracle -lm
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rv_b200.h"

int rvo_chain(const uint8_t *bgr, int H, int W, int space, double clip_limit, int grid, int ksize, uint8_t *out);

int main(void)
{
    const int n = 3, h = 211, w = 389;
    const size_t fb = (size_t)3 * w * h;
    uint8_t *in = malloc(fb * n), *out = malloc(fb * n), *want = malloc(fb);
    uint32_t s = 12345u;
    for (size_t i = 0; i < fb * n; ++i) { s = s * 1664525u + 1013904223u; in[i] = (uint8_t)(150 + ((s >> 24) & 63)); }

    rv_ctx *ctx = NULL;
    int rc = rv_create(0, &ctx);
    if (rc != RV_OK) { fprintf(stderr, "rv_create failed: %d (no sm_100 GPU?)\n", rc); return 2; }
    printf("%s\n", rv_version());

    int bad = 0;
    const int cases[4][3] = {{RV_SPACE_YCRCB, 8, 3}, {RV_SPACE_LAB, 8, 5}, {RV_SPACE_YCRCB, 7, 7}, {RV_SPACE_LAB, 16, 0}};
    for (int c = 0; c < 4; ++c) {
        rv_params p;
        memset(&p, 0, sizeof p);
        p.space = cases[c][0]; p.grid = cases[c][1]; p.ksize = cases[c][2]; p.clahe = 1; p.clip_limit = 2.0;
        rc = rv_chain_u8(ctx, in, out, n, h, w, (size_t)3 * w, (size_t)3 * w, &p, RV_MEM_HOST, NULL, NULL);
        if (rc != RV_OK) { fprintf(stderr, "rv_chain_u8: %d %s\n", rc, rv_last_error(ctx)); return 3; }
        for (int f = 0; f < n; ++f) {
            if (rvo_chain(in + f * fb, h, w, p.space, p.clip_limit, p.grid, p.ksize, want) != 0) return 4;
            if (memcmp(want, out + f * fb, fb) != 0) { fprintf(stderr, "case %d frame %d differs\n", c, f); ++bad; }
        }
    }
    /* error path: bad ksize must be refused with a message, nothing computed */
    rv_params p;
    memset(&p, 0, sizeof p);
    p.grid = 8; p.ksize = 4; p.clahe = 1; p.clip_limit = 2.0;
    rc = rv_chain_u8(ctx, in, out, 1, h, w, (size_t)3 * w, (size_t)3 * w, &p, RV_MEM_HOST, NULL, NULL);
    if (rc != RV_ERR_ARG || strlen(rv_last_error(ctx)) == 0) { fprintf(stderr, "bad ksize was not rejected (%d)\n", rc); ++bad; }
    printf("launches %ld, mismatches %d\n", rv_launch_count(ctx), bad);
    rv_destroy(ctx);
    free(in); free(out); free(want);
    return bad ? 1 : 0;
}
