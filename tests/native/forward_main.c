/* Native caller of the whole-forward entry points: no Python, no torch, no CUDA headers -- only include/bflow_b200.h.
 *   forward_main <plan file> <voxel.f32 | -> <image0.f32 | -> <image1.f32 | -> <low_out.f32> <up_out.f32> [repeats]
 * Inputs / outputs are raw little-endian float32 files in the reference's NCHW layouts. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bflow_b200.h"

static float* read_f32(const char* path, size_t n) {
    if (strcmp(path, "-") == 0) return NULL;
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    float* p = (float*)malloc(n * sizeof(float));
    if (fread(p, sizeof(float), n, f) != n) { fprintf(stderr, "%s: short read\n", path); exit(2); }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s plan voxel img0 img1 low_out up_out [repeats]\n", argv[0]); return 2; }
    bflow_forward* fw = NULL;
    if (bflow_forward_load(argv[1], &fw) != BFLOW_OK) { fprintf(stderr, "load failed: %s\n", bflow_last_error()); return 1; }
    int m[16];
    bflow_forward_info(fw, m);
    const size_t B = m[0], Cv = m[1], H = m[2], W = m[3], h = m[4], w = m[5], c2 = m[6];
    float* vox = m[8] ? read_f32(argv[2], B * Cv * H * W) : NULL;
    float* im0 = m[9] ? read_f32(argv[3], B * 3 * H * W) : NULL;
    float* im1 = m[9] ? read_f32(argv[4], B * 3 * H * W) : NULL;
    float* low = (float*)malloc(B * c2 * h * w * sizeof(float));
    float* up = (float*)malloc(B * c2 * H * W * sizeof(float));
    const int reps = argc > 7 ? atoi(argv[7]) : 1;
    for (int r = 0; r < reps; ++r) {
        if (bflow_forward_run(fw, vox, im0, im1, NULL, low, up, NULL) != BFLOW_OK) { fprintf(stderr, "run failed: %s\n", bflow_last_error()); return 1; }
    }
    bflow_forward_destroy(fw);   /* synchronises: the host results are complete */
    FILE* f = fopen(argv[5], "wb"); fwrite(low, sizeof(float), B * c2 * h * w, f); fclose(f);
    f = fopen(argv[6], "wb"); fwrite(up, sizeof(float), B * c2 * H * W, f); fclose(f);
    printf("forward %zux%zux%zu, %d iterations, %d run(s): ok\n", B, H, W, m[7], reps);
    return 0;
}
