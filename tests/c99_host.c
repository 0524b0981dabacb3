/* c99_host.c -- a plain C99 caller of the C ABI (gcc -std=c99 -pedantic -Wall -Werror), the closest thing to compiling
 * the ISO_C_BINDING module this image allows (no Fortran compiler): it includes every header under include/, takes the
 * address of every entry point (so the prototypes and the exported symbols must agree), prints the layout of the structs
 * that cross the boundary, and -- on a GPU box -- drives a small case through the same call sequence the Fortran host of
 * INTEGRATION.md makes (main.f90:80-139), from a binary blob the test writes, so that Python never touches the handle.
 *
 *   c99_host layout                    sizeof / offsetof of swpc3d_grid, swpc3d_snap_cfg, swpcpsv_grid
 *   c99_host run <in.bin> <out.bin>    create, upload, nt steps, download
 * Test infrastructure only. */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "swpc3d_b200.h"
#include "swpc3d_host.h"
#include "swpcpsv_b200.h"
#include "swpcpsv_host.h"

#define OFF(T, f) printf(#T "." #f " %u\n", (unsigned)offsetof(T, f))

static int layout(void) {
    printf("sizeof swpc3d_grid %u\n", (unsigned)sizeof(swpc3d_grid));
    OFF(swpc3d_grid, nx); OFF(swpc3d_grid, ny); OFF(swpc3d_grid, nz); OFF(swpc3d_grid, nproc_x); OFF(swpc3d_grid, nproc_y);
    OFF(swpc3d_grid, myid); OFF(swpc3d_grid, ibeg); OFF(swpc3d_grid, iend); OFF(swpc3d_grid, jbeg); OFF(swpc3d_grid, jend);
    OFF(swpc3d_grid, ipad); OFF(swpc3d_grid, jpad); OFF(swpc3d_grid, kpad); OFF(swpc3d_grid, ibeg_k); OFF(swpc3d_grid, iend_k);
    OFF(swpc3d_grid, jbeg_k); OFF(swpc3d_grid, jend_k); OFF(swpc3d_grid, kbeg_k); OFF(swpc3d_grid, kend_k); OFF(swpc3d_grid, na);
    OFF(swpc3d_grid, nm); OFF(swpc3d_grid, abc_type); OFF(swpc3d_grid, field_bytes); OFF(swpc3d_grid, device); OFF(swpc3d_grid, reserved);
    OFF(swpc3d_grid, dx); OFF(swpc3d_grid, dy); OFF(swpc3d_grid, dz); OFF(swpc3d_grid, dt); OFF(swpc3d_grid, reserved_f);
    printf("sizeof swpc3d_snap_cfg %u\n", (unsigned)sizeof(swpc3d_snap_cfg));
    OFF(swpc3d_snap_cfg, idec); OFF(swpc3d_snap_cfg, nxs); OFF(swpc3d_snap_cfg, is0); OFF(swpc3d_snap_cfg, k0_xy); OFF(swpc3d_snap_cfg, sw);
    OFF(swpc3d_snap_cfg, M0); OFF(swpc3d_snap_cfg, UC);
    printf("sizeof swpcpsv_grid %u\n", (unsigned)sizeof(swpcpsv_grid));
    OFF(swpcpsv_grid, nx); OFF(swpcpsv_grid, nz); OFF(swpcpsv_grid, device); OFF(swpcpsv_grid, dx); OFF(swpcpsv_grid, dz); OFF(swpcpsv_grid, dt);
    return 0;
}

/* every entry point of the two kernel ABIs, as the object pointers a linker has to resolve */
typedef void (*anyfn)(void);
static anyfn entry_points[] = {
    (anyfn)swpc3d_last_error, (anyfn)swpc3d_version, (anyfn)swpc3d_create, (anyfn)swpc3d_destroy, (anyfn)swpc3d_upload_medium,
    (anyfn)swpc3d_upload_fields, (anyfn)swpc3d_download_fields, (anyfn)swpc3d_zero_state, (anyfn)swpc3d_setup_pml, (anyfn)swpc3d_setup_cerjan,
    (anyfn)swpc3d_set_sources, (anyfn)swpc3d_set_stations, (anyfn)swpc3d_set_wav_products, (anyfn)swpc3d_get_wav_product, (anyfn)swpc3d_set_green,
    (anyfn)swpc3d_green_store, (anyfn)swpc3d_green_source, (anyfn)swpc3d_get_green, (anyfn)swpc3d_update_stress, (anyfn)swpc3d_stressglut,
    (anyfn)swpc3d_comm_stress, (anyfn)swpc3d_update_vel, (anyfn)swpc3d_bodyforce, (anyfn)swpc3d_comm_vel, (anyfn)swpc3d_wav_store, (anyfn)swpc3d_step,
    (anyfn)swpc3d_advance, (anyfn)swpc3d_run, (anyfn)swpc3d_sync, (anyfn)swpc3d_vmax, (anyfn)swpc3d_vmax_global, (anyfn)swpc3d_get_wav,
    (anyfn)swpc3d_snap_setup, (anyfn)swpc3d_snap_step, (anyfn)swpc3d_snap_fetch, (anyfn)swpc3d_snap_fetch_max, (anyfn)swpc3d_snap_fetch_begin, (anyfn)swpc3d_snap_fetch_end,
    (anyfn)swpc3d_reduce_sum,
    (anyfn)swpc3d_nccl_unique_id, (anyfn)swpc3d_comm_init, (anyfn)swpc3d_comm_local, (anyfn)swpc3d_timer_start, (anyfn)swpc3d_timer_stop,
    (anyfn)swpc3d_set_option, (anyfn)swpc3d_get_info,
    (anyfn)swpcpsv_last_error, (anyfn)swpcpsv_version, (anyfn)swpcpsv_create, (anyfn)swpcpsv_destroy, (anyfn)swpcpsv_step, (anyfn)swpcpsv_run,
};

static void *rd(FILE *fp, size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, fp) != bytes) { fprintf(stderr, "c99_host: short read (%lu bytes)\n", (unsigned long)bytes); exit(2); }
    return p;
}
#define CALL(x) do { if (x) { fprintf(stderr, "c99_host: %s -> %s\n", #x, swpc3d_last_error()); return 1; } } while (0)

static int run(const char *in, const char *out) {
    FILE *fp = fopen(in, "rb");
    if (!fp) { perror(in); return 2; }
    swpc3d_grid *g = (swpc3d_grid *)rd(fp, sizeof(swpc3d_grid));
    float *ts = (float *)rd(fp, 3 * sizeof(float));
    int32_t *hdr = (int32_t *)rd(fp, 8 * sizeof(int32_t));   /* nt nsrc nst ntdec_w ntw bf_mode reserved reserved */
    float *scale = (float *)rd(fp, 2 * sizeof(float));        /* M0 UC */
    char *stftype = (char *)rd(fp, 16);
    const int32_t nt = hdr[0], nsrc = hdr[1], nst = hdr[2], ntdec_w = hdr[3], ntw = hdr[4], bf_mode = hdr[5];
    const size_t nxp = (size_t)(g->iend - g->ibeg + 1), nyp = (size_t)(g->jend - g->jbeg + 1);
    const size_t n2 = (nxp + 6 + (size_t)g->ipad) * (nyp + 6 + (size_t)g->jpad), n3 = n2 * (size_t)(g->nz + 6 + g->kpad);
    float *med[5];
    int32_t *map[7];
    float *prof[6];
    const size_t plen[6] = {nxp, nxp, nyp, nyp, (size_t)g->nz, (size_t)g->nz};
    int a;
    for (a = 0; a < 5; a++) med[a] = (float *)rd(fp, n3 * sizeof(float));
    for (a = 0; a < 7; a++) map[a] = (int32_t *)rd(fp, n2 * sizeof(int32_t));
    for (a = 0; a < 6; a++) prof[a] = (float *)rd(fp, plen[a] * 4 * sizeof(float));
    int32_t *sijk = (int32_t *)rd(fp, 3 * (size_t)nsrc * sizeof(int32_t));   /* isrc[], jsrc[], ksrc[] */
    double *mo = (double *)rd(fp, (size_t)nsrc * sizeof(double));
    double *mij = (double *)rd(fp, 6 * (size_t)nsrc * sizeof(double));       /* mxx[], myy[], ... */
    float *prm = (float *)rd(fp, 2 * (size_t)nsrc * sizeof(float));
    int32_t *tijk = (int32_t *)rd(fp, 3 * (size_t)nst * sizeof(int32_t));
    fclose(fp);

    swpc3d_handle *h = NULL;
    CALL(swpc3d_create(g, ts, &h));                                                               /* memory_allocate, kernel__setup */
    CALL(swpc3d_upload_medium(h, med[0], med[1], med[2], med[3], med[4], map[0], map[1], map[2], map[3], map[4], map[5], map[6]));
    if (g->abc_type == SWPC3D_ABC_PML) CALL(swpc3d_setup_pml(h, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5]));
    else { fprintf(stderr, "c99_host: pml cases only\n"); return 2; }
    CALL(swpc3d_set_sources(h, nsrc, sijk, sijk + nsrc, sijk + 2 * nsrc, mo, mij, mij + nsrc, mij + 2 * nsrc, mij + 3 * nsrc, mij + 4 * nsrc,
                            mij + 5 * nsrc, prm, stftype, bf_mode, 0.0f));
    CALL(swpc3d_set_stations(h, nst, tijk, tijk + nst, tijk + 2 * nst, ntdec_w, ntw, scale[0], scale[1]));
    {
        int32_t it;
        for (it = 1; it <= nt; it++) {   /* main.f90:119-139, one call per reference subroutine */
            CALL(swpc3d_wav_store(h, it));
            CALL(swpc3d_update_stress(h));
            CALL(swpc3d_stressglut(h, it));
            CALL(swpc3d_comm_stress(h));
            CALL(swpc3d_update_vel(h));
            CALL(swpc3d_bodyforce(h, it));
            CALL(swpc3d_comm_vel(h));
        }
    }
    CALL(swpc3d_sync(h));
    {
        float vmax[3];
        const size_t fb = (size_t)g->field_bytes;
        char *F = (char *)malloc(9 * n3 * fb);
        float *wav = (float *)calloc((size_t)ntw * 3 * (size_t)(nst > 0 ? nst : 1), sizeof(float));
        FILE *fo;
        if (!F || !wav) return 2;
        CALL(swpc3d_vmax(h, vmax));
        CALL(swpc3d_download_fields(h, F, F + n3 * fb, F + 2 * n3 * fb, F + 3 * n3 * fb, F + 4 * n3 * fb, F + 5 * n3 * fb, F + 6 * n3 * fb,
                                    F + 7 * n3 * fb, F + 8 * n3 * fb));
        if (nst > 0) CALL(swpc3d_get_wav(h, wav));
        fo = fopen(out, "wb");
        if (!fo) { perror(out); return 2; }
        fwrite(vmax, sizeof(float), 3, fo);
        fwrite(F, fb, 9 * n3, fo);
        fwrite(wav, sizeof(float), (size_t)ntw * 3 * (size_t)nst, fo);
        fclose(fo);
        free(F);
        free(wav);
    }
    CALL(swpc3d_destroy(h));
    printf("c99_host: %d steps, %s\n", (int)nt, swpc3d_version());
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "layout")) {
        printf("entry points %u\n", (unsigned)(sizeof(entry_points) / sizeof(entry_points[0])));
        return layout();
    }
    if (argc >= 4 && !strcmp(argv[1], "run")) return run(argv[2], argv[3]);
    fprintf(stderr, "usage: c99_host layout | run <in.bin> <out.bin>\n");
    return 2;
}
