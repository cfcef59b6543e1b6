/*
 * msda_oracle.c -- CPU restatement of the reference's multiscale-deformable-attention path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under msda-triton_b200/ may import, link or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / the reported CPU baseline -- never as the product path.
 *
 * What it restates (all citations are into /root/reference/src/msda_triton/):
 *   level offsets ............ kernels.py:44-64   load_shapes_and_level_offsets (exclusive prefix sum of h*w)
 *   un-normalise ............. kernels.py:141-146 (align_corners ? x*(w-1) : x*w - 0.5, mul THEN sub, no FMA)
 *   floor / neighbours ....... kernels.py:150-153
 *   zeros-mode validity ...... kernels.py:158-162
 *   clamp-in-float then int .. kernels.py:166-169
 *   row addressing ........... kernels.py:180-203 ((level_off + y*w + x) * H*C + hid*C + c)
 *   masked values ............ kernels.py:213-231
 *   bilinear blend ........... kernels.py:235-244 (v00*(1-dy)*(1-dx) + v01*(1-dy)*dx + v10*dy*(1-dx) + v11*dy*dx)
 *   forward reduction ........ kernels.py:339     out = sum_{l,p} aw * sample
 *   grad attention weights ... kernels.py:494
 *   grad sampling points ..... kernels.py:510-524
 *   grad img (scatter) ....... kernels.py:543-553
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the .npz files under tests/golden, which were
 * produced by oracle/make_golden.py from (i) the reference's own Triton kernels executed by Triton's
 * CPU interpreter and (ii) the reference's native grid_sample fallback (frontend.py:15-68).
 *
 * Arithmetic is done in the REAL type of each instantiation (float or double); the summation order over
 * (l, p) is sequential here and a tree in Triton, so fp32 results agree to rounding, not bit-exactly.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MSDA_ORACLE_PAD_ZEROS 0
#define MSDA_ORACLE_PAD_BORDER 1

typedef struct {
    int64_t h, w, off;
} level_t;

/* kernels.py:60-62 -- sizes = h*w ; level_offsets = cumsum(sizes) - sizes */
static level_t *build_levels(const int64_t *shapes, int64_t L) {
    level_t *lv = (level_t *)malloc(sizeof(level_t) * (size_t)(L > 0 ? L : 1));
    int64_t run = 0;
    for (int64_t l = 0; l < L; ++l) {
        lv[l].h = shapes[2 * l + 0];
        lv[l].w = shapes[2 * l + 1];
        lv[l].off = run;
        run += lv[l].h * lv[l].w;
    }
    return lv;
}

#define DEFINE_ORACLE(SFX, REAL, FLOOR, FMIN, FMAX)                                                        \
                                                                                                           \
    typedef struct {                                                                                       \
        int64_t r00, r01, r10, r11; /* pixel-row indices incl. level offset (kernels.py:184-203) */        \
        int m00, m01, m10, m11;     /* zeros-mode validity (kernels.py:227-231); all 1 in border mode */   \
        REAL dx, dy;                /* kernels.py:235-237 */                                               \
    } tap_##SFX;                                                                                           \
                                                                                                           \
    static inline REAL clampr_##SFX(REAL v, REAL lo, REAL hi) { return FMIN(FMAX(v, lo), hi); }            \
                                                                                                           \
    static inline void locate_##SFX(REAL px, REAL py, const level_t *lv, int padding, int align,           \
                                    tap_##SFX *t) {                                                        \
        const REAL wf = (REAL)lv->w, hf = (REAL)lv->h;                                                     \
        REAL x, y;                                                                                         \
        if (align) {                                                                                       \
            x = px * (wf - (REAL)1);                                                                       \
            y = py * (hf - (REAL)1);                                                                       \
        } else {                                                                                           \
            x = px * wf;                                                                                   \
            x = x - (REAL)0.5;                                                                             \
            y = py * hf;                                                                                   \
            y = y - (REAL)0.5;                                                                             \
        }                                                                                                  \
        const REAL x0 = FLOOR(x), y0 = FLOOR(y);                                                           \
        const REAL x1 = x0 + (REAL)1, y1 = y0 + (REAL)1;                                                   \
        int x0m = 1, x1m = 1, y0m = 1, y1m = 1;                                                            \
        if (padding == MSDA_ORACLE_PAD_ZEROS) {                                                            \
            x0m = ((REAL)0 <= x0) && (x0 <= wf - (REAL)1);                                                 \
            x1m = ((REAL)0 <= x1) && (x1 <= wf - (REAL)1);                                                 \
            y0m = ((REAL)0 <= y0) && (y0 <= hf - (REAL)1);                                                 \
            y1m = ((REAL)0 <= y1) && (y1 <= hf - (REAL)1);                                                 \
        }                                                                                                  \
        const int64_t x0c = (int64_t)clampr_##SFX(x0, (REAL)0, wf - (REAL)1);                              \
        const int64_t x1c = (int64_t)clampr_##SFX(x1, (REAL)0, wf - (REAL)1);                              \
        const int64_t y0c = (int64_t)clampr_##SFX(y0, (REAL)0, hf - (REAL)1);                              \
        const int64_t y1c = (int64_t)clampr_##SFX(y1, (REAL)0, hf - (REAL)1);                              \
        t->r00 = lv->off + y0c * lv->w + x0c;                                                              \
        t->r01 = lv->off + y0c * lv->w + x1c;                                                              \
        t->r10 = lv->off + y1c * lv->w + x0c;                                                              \
        t->r11 = lv->off + y1c * lv->w + x1c;                                                              \
        t->m00 = y0m && x0m;                                                                               \
        t->m01 = y0m && x1m;                                                                               \
        t->m10 = y1m && x0m;                                                                               \
        t->m11 = y1m && x1m;                                                                               \
        t->dx = x - x0;                                                                                    \
        t->dy = y - y0;                                                                                    \
    }                                                                                                      \
                                                                                                           \
    /* Forward.  Layouts: img [B,Npix,H,D]; pts [B,Q,H,L,K,2] (x,y); aw [B,Q,H,L,K]; out [B,Q,H,D]. */     \
    int msda_oracle_fwd_##SFX(REAL *out, const REAL *img, const int64_t *shapes, const REAL *pts,          \
                              const REAL *aw, int64_t B, int64_t Npix, int64_t H, int64_t D, int64_t Q,    \
                              int64_t L, int64_t K, int padding, int align) {                              \
        level_t *lv = build_levels(shapes, L);                                                             \
        const int64_t units = B * Q * H;                                                                   \
        _Pragma("omp parallel for schedule(static)") for (int64_t u = 0; u < units; ++u) {                 \
            const int64_t h = u % H, b = u / (H * Q);                                                      \
            const REAL *ib = img + (size_t)b * Npix * H * D + (size_t)h * D;                               \
            REAL *o = out + (size_t)u * D;                                                                 \
            for (int64_t c = 0; c < D; ++c) o[c] = (REAL)0;                                                \
            for (int64_t l = 0; l < L; ++l)                                                                \
                for (int64_t k = 0; k < K; ++k) {                                                          \
                    const size_t pi = ((size_t)u * L + l) * K + k;                                         \
                    tap_##SFX t;                                                                           \
                    locate_##SFX(pts[2 * pi], pts[2 * pi + 1], &lv[l], padding, align, &t);                \
                    const REAL a = aw[pi];                                                                 \
                    const REAL *p00 = ib + (size_t)t.r00 * H * D, *p01 = ib + (size_t)t.r01 * H * D;       \
                    const REAL *p10 = ib + (size_t)t.r10 * H * D, *p11 = ib + (size_t)t.r11 * H * D;       \
                    for (int64_t c = 0; c < D; ++c) {                                                      \
                        const REAL v00 = t.m00 ? p00[c] : (REAL)0, v01 = t.m01 ? p01[c] : (REAL)0;         \
                        const REAL v10 = t.m10 ? p10[c] : (REAL)0, v11 = t.m11 ? p11[c] : (REAL)0;         \
                        const REAL s = v00 * ((REAL)1 - t.dy) * ((REAL)1 - t.dx) +                         \
                                       v01 * ((REAL)1 - t.dy) * (t.dx) + v10 * (t.dy) * ((REAL)1 - t.dx) + \
                                       v11 * (t.dy) * (t.dx);                                              \
                        o[c] += a * s;                                                                     \
                    }                                                                                      \
                }                                                                                          \
        }                                                                                                  \
        free(lv);                                                                                          \
        return 0;                                                                                          \
    }                                                                                                      \
                                                                                                           \
    /* Backward.  gimg is zero-filled here (kernels.py:570).  Parallel over (b,h): rows of different      \
     * (b,h) never alias, so no atomics are needed and the result is run-to-run deterministic. */          \
    int msda_oracle_bwd_##SFX(REAL *gimg, REAL *gpts, REAL *gaw, const REAL *gout, const REAL *img,        \
                              const int64_t *shapes, const REAL *pts, const REAL *aw, int64_t B,           \
                              int64_t Npix, int64_t H, int64_t D, int64_t Q, int64_t L, int64_t K,         \
                              int padding, int align) {                                                    \
        level_t *lv = build_levels(shapes, L);                                                             \
        memset(gimg, 0, sizeof(REAL) * (size_t)B * Npix * H * D);                                          \
        const int64_t BH = B * H;                                                                          \
        _Pragma("omp parallel for schedule(dynamic, 1)") for (int64_t bh = 0; bh < BH; ++bh) {             \
            const int64_t b = bh / H, h = bh % H;                                                          \
            const REAL *ib = img + (size_t)b * Npix * H * D + (size_t)h * D;                               \
            REAL *gb = gimg + (size_t)b * Npix * H * D + (size_t)h * D;                                    \
            for (int64_t q = 0; q < Q; ++q) {                                                              \
                const size_t u = ((size_t)b * Q + q) * H + h;                                              \
                const REAL *go = gout + u * D;                                                             \
                for (int64_t l = 0; l < L; ++l) {                                                          \
                    const REAL xs = align ? (REAL)(lv[l].w - 1) : (REAL)lv[l].w;                           \
                    const REAL ys = align ? (REAL)(lv[l].h - 1) : (REAL)lv[l].h;                           \
                    for (int64_t k = 0; k < K; ++k) {                                                      \
                        const size_t pi = (u * L + l) * K + k;                                             \
                        tap_##SFX t;                                                                       \
                        locate_##SFX(pts[2 * pi], pts[2 * pi + 1], &lv[l], padding, align, &t);            \
                        const REAL a = aw[pi];                                                             \
                        const size_t o00 = (size_t)t.r00 * H * D, o01 = (size_t)t.r01 * H * D;             \
                        const size_t o10 = (size_t)t.r10 * H * D, o11 = (size_t)t.r11 * H * D;             \
                        REAL ga = (REAL)0, gx = (REAL)0, gy = (REAL)0;                                     \
                        for (int64_t c = 0; c < D; ++c) {                                                  \
                            const REAL v00 = t.m00 ? ib[o00 + c] : (REAL)0;                                \
                            const REAL v01 = t.m01 ? ib[o01 + c] : (REAL)0;                                \
                            const REAL v10 = t.m10 ? ib[o10 + c] : (REAL)0;                                \
                            const REAL v11 = t.m11 ? ib[o11 + c] : (REAL)0;                                \
                            const REAL s = v00 * ((REAL)1 - t.dy) * ((REAL)1 - t.dx) +                     \
                                           v01 * ((REAL)1 - t.dy) * (t.dx) +                               \
                                           v10 * (t.dy) * ((REAL)1 - t.dx) + v11 * (t.dy) * (t.dx);        \
                            ga += go[c] * s;                                                               \
                            gx += go[c] * a * xs *                                                         \
                                  (((REAL)1 - t.dy) * (v01 - v00) + t.dy * (v11 - v10));                   \
                            gy += go[c] * a * ys *                                                         \
                                  (((REAL)1 - t.dx) * (v10 - v00) + t.dx * (v11 - v01));                   \
                            if (t.m00) gb[o00 + c] += go[c] * a * ((REAL)1 - t.dy) * ((REAL)1 - t.dx);     \
                            if (t.m01) gb[o01 + c] += go[c] * a * ((REAL)1 - t.dy) * (t.dx);               \
                            if (t.m10) gb[o10 + c] += go[c] * a * (t.dy) * ((REAL)1 - t.dx);               \
                            if (t.m11) gb[o11 + c] += go[c] * a * (t.dy) * (t.dx);                         \
                        }                                                                                  \
                        gaw[pi] = ga;                                                                      \
                        gpts[2 * pi] = gx;                                                                 \
                        gpts[2 * pi + 1] = gy;                                                             \
                    }                                                                                      \
                }                                                                                          \
            }                                                                                              \
        }                                                                                                  \
        free(lv);                                                                                          \
        return 0;                                                                                          \
    }

DEFINE_ORACLE(f32, float, floorf, fminf, fmaxf)
DEFINE_ORACLE(f64, double, floor, fmin, fmax)

/* Level table as the device code must derive it (kernels.py:60-62): rows of {h, w, offset}. */
int msda_oracle_level_table(int64_t *table, const int64_t *shapes, int64_t L) {
    level_t *lv = build_levels(shapes, L);
    for (int64_t l = 0; l < L; ++l) {
        table[3 * l + 0] = lv[l].h;
        table[3 * l + 1] = lv[l].w;
        table[3 * l + 2] = lv[l].off;
    }
    free(lv);
    return 0;
}

int msda_oracle_abi_version(void) { return 1; }

/* Host threads used by the parallel loops above (reported as cpu_baseline.cores by bench.py). */
void msda_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int msda_oracle_max_threads(void) { return omp_get_max_threads(); }
