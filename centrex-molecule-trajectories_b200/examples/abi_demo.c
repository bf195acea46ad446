/*
 * abi_demo.c -- the C ABI of libcmt_b200.so used from plain C: no Python, no torch, no CUDA headers.
 *
 *   gcc -O2 -I include centrex-molecule-trajectories_b200/examples/abi_demo.c -ldl -lm -o abi_demo
 *   ./abi_demo centrex-molecule-trajectories_b200/lib/libcmt_b200.so 1000000 7
 *
 * Builds the apertures + lens + field plates + detection aperture beamline of
 * examples/lens_simulation_beamline.py (reference) with a harmonic stand-in for the lens table
 * (a_r = -k r), propagates N molecules drawn on the device from the CeNTREX source and prints
 * the per-fate counts as one JSON line.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cmt.h"

#define LOAD(name) \
    *(void **)(&name##_) = dlsym(lib, #name); \
    if (!name##_) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s libcmt_b200.so [N] [seed]\n", argv[0]); return 2; }
    const long long n = argc > 2 ? atoll(argv[2]) : 1000000;
    const unsigned long long seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }

    int (*cmt_beamline_create_)(const cmt_element_t *, int, const cmt_table_t *, int, int, int, double, int, cmt_beamline_t **);
    void (*cmt_beamline_destroy_)(cmt_beamline_t *);
    int (*cmt_run_host_philox_)(const cmt_beamline_t *, const cmt_source_t *, uint64_t, int64_t, int64_t, int64_t *, int64_t *);
    const char *(*cmt_last_error_)(void);
    LOAD(cmt_beamline_create) LOAD(cmt_beamline_destroy) LOAD(cmt_run_host_philox) LOAD(cmt_last_error)

    const double in = 0.0254;
    const char *fates[] = { "4K shield", "40K shield", "BB exit", "Lens entrance", "Inside lens",
                            "Field plates", "DR aperture", "Detected" };
    cmt_element_t el[6];
    memset(el, 0, sizeof(el));
    el[0].type = CMT_CIRCULAR; el[0].fate = 0; el[0].z0 = 1.7 * in;               el[0].z1 = el[0].z0 + 0.25 * in; el[0].R = 0.5 * in;
    el[1].type = CMT_CIRCULAR; el[1].fate = 1; el[1].z0 = el[0].z1 + 1.25 * in;   el[1].z1 = el[1].z0 + 0.25 * in; el[1].R = 0.5 * in;
    el[2].type = CMT_CIRCULAR; el[2].fate = 2; el[2].z0 = el[1].z1 + 2.5 * in;    el[2].z1 = el[2].z0 + 0.75 * in; el[2].R = 2.0 * in;
    el[3].type = CMT_LENS;     el[3].fate = 3; el[3].fate2 = 4; el[3].table = 0;
    el[3].z0 = el[2].z1 + 33 * in; el[3].z1 = el[3].z0 + 0.6; el[3].R = 1.75 * in / 2; el[3].dz = 1e-3;
    el[3].n_steps = (int)rint(0.6 / 1e-3);
    el[4].type = CMT_FIELDPLATES; el[4].fate = 5; el[4].z0 = 2.43; el[4].z1 = 5.43; el[4].x1 = -0.01; el[4].x2 = 0.01;
    el[5].type = CMT_RECTANGULAR; el[5].fate = 6; el[5].z0 = el[4].z1 + 39.9 * in; el[5].z1 = el[5].z0 + 0.25 * in;
    el[5].x1 = -0.009; el[5].x2 = 0.009; el[5].y1 = -0.015; el[5].y2 = 0.015;

    enum { NT = 222 };
    static double r[NT], a[NT];
    for (int i = 0; i < NT; ++i) { r[i] = 1.01 * el[3].R * i / (NT - 1); a[i] = -2.0e4 * r[i]; }
    cmt_table_t tab = { r, a, NT, 0 };

    cmt_source_t src;
    memset(&src, 0, sizeof(src));
    src.pos_kind = CMT_POS_DISC; src.p0 = 0.01; src.z = 0.25 * in;
    src.vmean[2] = 184.0; src.vsigma[0] = 39.5; src.vsigma[1] = 39.5; src.vsigma[2] = 16.0;

    cmt_beamline_t *bl = NULL;
    if (cmt_beamline_create_(el, 6, &tab, 1, 8, 7, 9.80665, 0, &bl)) { fprintf(stderr, "create: %s\n", cmt_last_error_()); return 1; }
    int64_t counters[8] = {0}, work[CMT_WORK_SLOTS] = {0};
    if (cmt_run_host_philox_(bl, &src, seed, 0, n, counters, work)) { fprintf(stderr, "run: %s\n", cmt_last_error_()); return 1; }
    cmt_beamline_destroy_(bl);

    printf("{");
    for (int f = 0; f < 8; ++f) printf("\"%s\": %lld, ", fates[f], (long long)counters[f]);
    printf("\"lens_rk_steps\": %lld}\n", (long long)work[1]);
    return 0;
}
