/*
 * TEST INFRASTRUCTURE — plain-C restatement ("port") of the reference's
 * per-frame image-preparation path and of the responseCalib inner loops.
 *
 * It is the checker for the CUDA path (tests/, __graft_entry__.smoke(), and
 * bench.py's cpu_baseline / --impl reference legs are the only allowed users)
 * and is itself pinned against the reference's own code compiled unmodified
 * (oracle/_ref, see oracle/Makefile and tests/test_oracle_port_vs_ref.py).
 * Parity status: PINNED for rows a1-a5 of SURVEY.md §8 (bit-compared with
 * oracle/_ref on every fixture); UNPINNED for the pyramid (row P: not in the
 * reference at all) and pinned-by-restatement-only for the E/G-step, rmse and
 * leak padding (inline in the reference's main(), no compilable unit).
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/).  Floating-point evaluation order and operand widths are
 * kept exactly; compile with -ffp-contract=off and without -ffast-math.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

/* ------------------------------------------------------------------ FOV model */

typedef struct {
    float in_calib[5];  /* fx fy cx cy omega, relative to the input size */
    float out_calib[5]; /* after renormalisation (FOVUndistorter.cpp:214-218) */
    int in_w, in_h, out_w, out_h;
    int float_math;     /* 0: tan()/sqrt() are the double functions (canonical), 1: float overloads */
} oport_fov;


/* d2t = 2.0f * tan(dist / 2.0f)  (FOVUndistorter.cpp:132, :290).  With the double
 * tan the product is formed in double and narrowed once on assignment. */
static float d2t_of(float dist, int float_math) {
    if (float_math) return 2.0f * tanf(dist / 2.0f);
    return (float)(2.0 * tan((double)(dist / 2.0f)));
}

/* tan(radius*dist)/d2t  (FOVUndistorter.cpp:159-162, :186-189) */
static float trans_radius(float radius, float dist, float d2t, int float_math) {
    if (float_math) return tanf(radius * dist) / d2t;
    return (float)(tan((double)(radius * dist)) / (double)d2t);
}

static float sqrt_sel(float x, int float_math) { return float_math ? sqrtf(x) : (float)sqrt((double)x); }
static float maxf(float a, float b) { return a < b ? b : a; } /* std::max(a,b): returns b iff a<b */

/*
 * Output-camera selection, FOVUndistorter.cpp:131-218.
 * mode: -1 crop, -2 full, 0 explicit (out_calib_in = relative fx fy cx cy, 5th ignored).
 */
void oport_fov_init(oport_fov* f, const float in_calib[5], int in_w, int in_h, int mode,
                    const float out_calib_in[5], int out_w, int out_h, int float_math) {
    memcpy(f->in_calib, in_calib, sizeof(float) * 5);
    f->in_w = in_w; f->in_h = in_h; f->out_w = out_w; f->out_h = out_h; f->float_math = float_math;

    float dist = in_calib[4];
    float d2t = d2t_of(dist, float_math);
    float fx = in_calib[0] * in_w;
    float fy = in_calib[1] * in_h;
    float cx = (float)((double)(in_calib[2] * in_w) - 0.5);   /* :137, double literal */
    float cy = (float)((double)(in_calib[3] * in_h) - 0.5);
    float ofx, ofy, ocx, ocy;

    if (in_calib[4] == 0) {                                   /* :144-150 */
        ofx = in_calib[0] * out_w;
        ofy = in_calib[1] * out_h;
        ocx = (float)((double)(in_calib[2] * out_w) - 0.5);
        ocy = (float)((double)(in_calib[3] * out_h) - 0.5);
    } else if (mode == -1) {                                  /* crop, :151-172 */
        float left_radius = cx / fx;
        float right_radius = (in_w - 1 - cx) / fx;
        float top_radius = cy / fy;
        float bottom_radius = (in_h - 1 - cy) / fy;
        float tl = trans_radius(left_radius, dist, d2t, float_math);
        float tr = trans_radius(right_radius, dist, d2t, float_math);
        float tt = trans_radius(top_radius, dist, d2t, float_math);
        float tb = trans_radius(bottom_radius, dist, d2t, float_math);
        ofy = fy * ((top_radius + bottom_radius) / (tt + tb)) * ((float)out_h / (float)in_h);
        ocy = (tt / top_radius) * ofy * cy / fy;
        ofx = fx * ((left_radius + right_radius) / (tl + tr)) * ((float)out_w / (float)in_w);
        ocx = (tl / left_radius) * ofx * cx / fx;
    } else if (mode == -2) {                                  /* full, :173-205 */
        float left_radius = cx / fx;
        float right_radius = (in_w - 1 - cx) / fx;
        float top_radius = cy / fy;
        float bottom_radius = (in_h - 1 - cy) / fy;
        float tl_r = sqrt_sel(left_radius * left_radius + top_radius * top_radius, float_math);
        float tr_r = sqrt_sel(right_radius * right_radius + top_radius * top_radius, float_math);
        float bl_r = sqrt_sel(left_radius * left_radius + bottom_radius * bottom_radius, float_math);
        float br_r = sqrt_sel(right_radius * right_radius + bottom_radius * bottom_radius, float_math);
        float ttl = trans_radius(tl_r, dist, d2t, float_math);
        float ttr = trans_radius(tr_r, dist, d2t, float_math);
        float tbl = trans_radius(bl_r, dist, d2t, float_math);
        float tbr = trans_radius(br_r, dist, d2t, float_math);
        float hor = maxf(br_r, tr_r) + maxf(bl_r, tl_r);
        float vert = maxf(tr_r, tl_r) + maxf(bl_r, br_r);
        float trans_hor = maxf(tbr, ttr) + maxf(tbl, ttl);
        float trans_vert = maxf(ttr, ttl) + maxf(tbl, tbr);
        ofy = fy * (vert / trans_vert) * ((float)out_h / (float)in_h);
        ocy = maxf(ttl / tl_r, ttr / tr_r) * ofy * cy / fy;
        ofx = fx * (hor / trans_hor) * ((float)out_w / (float)in_w);
        ocx = maxf(tbl / bl_r, ttl / tl_r) * ofx * cx / fx;
    } else {                                                  /* explicit, :206-212 */
        ofx = out_calib_in[0] * out_w;
        ofy = out_calib_in[1] * out_h;
        ocx = (float)((double)(out_calib_in[2] * out_w) - 0.5);
        ocy = (float)((double)(out_calib_in[3] * out_h) - 0.5);
    }
    f->out_calib[0] = ofx / out_w;                            /* :214-218 */
    f->out_calib[1] = ofy / out_h;
    f->out_calib[2] = (float)(((double)ocx + 0.5) / (double)out_w);
    f->out_calib[3] = (float)(((double)ocy + 0.5) / (double)out_h);
    f->out_calib[4] = 0;
}

/* UndistorterFOV::distortCoordinates, FOVUndistorter.cpp:280-319 (valid object assumed). */
void oport_fov_distort(const oport_fov* f, float* xs, float* ys, int n) {
    float dist = f->in_calib[4];
    float d2t = d2t_of(dist, f->float_math);
    float fx = f->in_calib[0] * f->in_w;
    float fy = f->in_calib[1] * f->in_h;
    float cx = (float)((double)(f->in_calib[2] * f->in_w) - 0.5);
    float cy = (float)((double)(f->in_calib[3] * f->in_h) - 0.5);
    float ofx = f->out_calib[0] * f->out_w;
    float ofy = f->out_calib[1] * f->out_h;
    float ocx = f->out_calib[2] * f->out_w - 0.5f;
    float ocy = f->out_calib[3] * f->out_h - 0.5f;
    for (int i = 0; i < n; i++) {
        float ix = (xs[i] - ocx) / ofx;
        float iy = (ys[i] - ocy) / ofy;
        float r = sqrtf(ix * ix + iy * iy);
        float fac = (r == 0 || dist == 0) ? 1 : atanf(r * d2t) / (dist * r);
        xs[i] = fx * fac * ix + cx;
        ys[i] = fy * fac * iy + cy;
    }
}

/* Remap-table build, FOVUndistorter.cpp:224-251.  Returns 1 if any entry was blacked out. */
int oport_fov_build_tables(const oport_fov* f, float* remap_x, float* remap_y) {
    int n = f->out_w * f->out_h, black = 0;
    for (int y = 0; y < f->out_h; y++)
        for (int x = 0; x < f->out_w; x++) {
            remap_x[x + y * f->out_w] = (float)x;
            remap_y[x + y * f->out_w] = (float)y;
        }
    oport_fov_distort(f, remap_x, remap_y, n);
    for (int i = 0; i < n; i++) {
        if (remap_x[i] == 0) remap_x[i] = (float)0.01;
        if (remap_y[i] == 0) remap_y[i] = (float)0.01;
        if (remap_x[i] == (float)(f->in_w - 1)) remap_x[i] = (float)(f->in_w - 1.01);
        if (remap_y[i] == (float)(f->in_h - 1)) remap_y[i] = (float)(f->in_h - 1.01);
        if (!(remap_x[i] > 0 && remap_y[i] > 0 && remap_x[i] < (float)(f->in_w - 1) && remap_y[i] < (float)(f->in_h - 1))) {
            black = 1;
            remap_x[i] = -1;
            remap_y[i] = -1;
        }
    }
    return black;
}

/* Krect / Korg diagonal+offset entries, FOVUndistorter.cpp:257-268: {fx, fy, cx, cy}. */
void oport_fov_K(const oport_fov* f, float krect4[4], float korg4[4]) {
    krect4[0] = f->out_calib[0] * f->out_w;
    krect4[1] = f->out_calib[1] * f->out_h;
    krect4[2] = (float)((double)(f->out_calib[2] * f->out_w) - 0.5);
    krect4[3] = (float)((double)(f->out_calib[3] * f->out_h) - 0.5);
    korg4[0] = f->in_calib[0] * f->in_w;
    korg4[1] = f->in_calib[1] * f->in_h;
    korg4[2] = (float)((double)(f->in_calib[2] * f->in_w) - 0.5);
    korg4[3] = (float)((double)(f->in_calib[3] * f->in_h) - 0.5);
}

/* UndistorterFOV::undistort<float> / <unsigned char>, FOVUndistorter.cpp:341-367. */
#define OPORT_UNDISTORT_BODY(T)                                                              \
    for (int idx = 0; idx < n_out; idx++) {                                                  \
        float xx = remap_x[idx], yy = remap_y[idx];                                          \
        if (xx < 0) { out[idx] = 0; continue; }                                              \
        int xi = (int)xx, yi = (int)yy;                                                      \
        xx -= xi; yy -= yi;                                                                  \
        float xxyy = xx * yy;                                                                \
        const T* s = in + xi + yi * in_w;                                                    \
        out[idx] = xxyy * s[1 + in_w] + (yy - xxyy) * s[in_w] + (xx - xxyy) * s[1]           \
                   + (1 - xx - yy + xxyy) * s[0];                                            \
    }
void oport_undistort_f32(const float* remap_x, const float* remap_y, int in_w, int n_out, const float* in, float* out) {
    OPORT_UNDISTORT_BODY(float)
}
void oport_undistort_u8(const float* remap_x, const float* remap_y, int in_w, int n_out, const unsigned char* in, float* out) {
    OPORT_UNDISTORT_BODY(unsigned char)
}

/* --------------------------------------------------------- photometric model */

/* GInv normalisation + forward G, PhotometricUndistorter.cpp:79-108.
 * Returns 0 if the raw table is not strictly increasing (object stays invalid). */
int oport_photo_tables(const float raw[256], float ginv[256], float g[256]) {
    for (int i = 0; i < 256; i++) ginv[i] = raw[i];
    for (int i = 0; i < 255; i++)
        if (ginv[i + 1] <= ginv[i]) return 0;
    float lo = ginv[0], hi = ginv[255];
    for (int i = 0; i < 256; i++) ginv[i] = (float)(255.0 * (double)(ginv[i] - lo) / (double)(hi - lo));
    for (int i = 1; i < 255; i++)
        for (int s = 1; s < 255; s++)
            if (ginv[s] <= i && ginv[s + 1] >= i) {
                g[i] = s + (i - ginv[s]) / (ginv[s + 1] - ginv[s]);
                break;
            }
    g[0] = 0; g[255] = 255;
    return 1;
}

/* vignette / max -> map, 1.0f/map -> inverse, PhotometricUndistorter.cpp:130-152.  depth = 8 or 16. */
void oport_vignette_maps(const void* pixels, int depth, int n, float* map, float* map_inv) {
    float maxv = 0;
    if (depth == 8) {
        const unsigned char* p = (const unsigned char*)pixels;
        for (int i = 0; i < n; i++) if (p[i] > maxv) maxv = p[i];
        for (int i = 0; i < n; i++) map[i] = p[i] / maxv;
    } else {
        const unsigned short* p = (const unsigned short*)pixels;
        for (int i = 0; i < n; i++) if (p[i] > maxv) maxv = p[i];
        for (int i = 0; i < n; i++) map[i] = p[i] / maxv;
    }
    for (int i = 0; i < n; i++) map_inv[i] = 1.0f / map[i];
}

/* PhotometricUndistorter::unMapImage, PhotometricUndistorter.cpp:173-211.
 * ginv / vinv may be NULL = "not loaded" (validGamma / validVignette false). */
void oport_unmap(const float* ginv, const float* vinv, const unsigned char* in, float* out, int n,
                 int undo_gamma, int undo_vignette, int kill_overexposed) {
    if (!ginv) undo_gamma = 0;
    if (!vinv) undo_vignette = 0;
    if (!undo_gamma && undo_vignette) { undo_vignette = 0; undo_gamma = 0; }
    if (!undo_gamma && !undo_vignette) for (int i = 0; i < n; i++) out[i] = in[i];
    if (undo_gamma && !undo_vignette) for (int i = 0; i < n; i++) out[i] = ginv[in[i]];
    if (undo_gamma && undo_vignette) for (int i = 0; i < n; i++) out[i] = ginv[in[i]] * vinv[i];
    if (kill_overexposed) for (int i = 0; i < n; i++) if (in[i] == 255) out[i] = NAN;
}

/* DatasetReader::getImage mode switch, BenchmarkDatasetReader.h:210-241.
 * out must hold n_out floats when rectify, else in_w*in_h.  tmp: in_w*in_h floats. */
void oport_get_image(const float* remap_x, const float* remap_y, int in_w, int in_h, int n_out,
                     const float* ginv, const float* vinv, const unsigned char* raw, float* tmp, float* out,
                     int rectify, int remove_gamma, int remove_vignette, int nan_overexposed) {
    int n_in = in_w * in_h;
    if (remove_gamma || remove_vignette || nan_overexposed) {
        if (!rectify) oport_unmap(ginv, vinv, raw, out, n_in, remove_gamma, remove_vignette, nan_overexposed);
        else {
            oport_unmap(ginv, vinv, raw, tmp, n_in, remove_gamma, remove_vignette, nan_overexposed);
            oport_undistort_f32(remap_x, remap_y, in_w, n_out, tmp, out);
        }
    } else {
        if (rectify) oport_undistort_u8(remap_x, remap_y, in_w, n_out, raw, out);
        else for (int i = 0; i < n_in; i++) out[i] = raw[i];
    }
}

/* ------------------------------------------------------------------- pyramid */
/* NOT IN THE REFERENCE (parity unpinned).  DSO-convention
 * 2x2 box filter: dst[x,y] = 0.25f*(((a+b)+c)+d), odd trailing row/col dropped. */
void oport_pyr_down(const float* src, int sw, int sh, float* dst) {
    int dw = sw >> 1, dh = sh >> 1;
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            const float* p = src + 2 * x + (size_t)2 * y * sw;
            dst[x + (size_t)y * dw] = 0.25f * (((p[0] + p[1]) + p[sw]) + p[sw + 1]);
        }
}

/* ---------------------------------------------------------------- responseCalib */

/* E-step, main_responseCalib.cpp:320-338.  data: n image-major planes of npix bytes. */
void oport_estep(const unsigned char* data, int n, int npix, const double* t, const double* G, double* E) {
    double* esum = (double*)calloc(npix, sizeof(double));
    double* enum_ = (double*)calloc(npix, sizeof(double));
    for (int i = 0; i < n; i++) {
        const unsigned char* d = data + (size_t)i * npix;
        for (int k = 0; k < npix; k++) {
            int b = d[k];
            if (b == 255) continue;
            enum_[k] += t[i] * t[i];
            esum[k] += G[b] * t[i];
        }
    }
    for (int k = 0; k < npix; k++) {
        E[k] = esum[k] / enum_[k];
        if (E[k] < 0) E[k] = 0;
    }
    free(esum); free(enum_);
}

/* G-step, main_responseCalib.cpp:286-304. */
void oport_gstep(const unsigned char* data, int n, int npix, const double* t, const double* E, double* G) {
    double gsum[256], gnum[256];
    memset(gsum, 0, sizeof gsum); memset(gnum, 0, sizeof gnum);
    for (int i = 0; i < n; i++) {
        const unsigned char* d = data + (size_t)i * npix;
        for (int k = 0; k < npix; k++) {
            int b = d[k];
            if (b == 255) continue;
            gnum[b]++;
            gsum[b] += E[k] * t[i];
        }
    }
    for (int i = 0; i < 256; i++) {
        G[i] = gsum[i] / gnum[i];
        if (!isfinite(G[i]) && i > 1) G[i] = G[i - 1] + (G[i - 1] - G[i - 2]);
    }
}

/* rmse(), main_responseCalib.cpp:50-69: out[0] = 1e5*sqrtl(e/num), out[1] = num. */
void oport_rmse(const unsigned char* data, int n, int npix, const double* t, const double* G, const double* E, double out[2]) {
    long double e = 0, num = 0;
    for (int i = 0; i < n; i++) {
        const unsigned char* d = data + (size_t)i * npix;
        for (int k = 0; k < npix; k++) {
            if (d[k] == 255) continue;
            double r = G[d[k]] - t[i] * E[k];
            if (!isfinite(r)) continue;
            e += r * r * 1e-10;
            num++;
        }
    }
    out[0] = (double)(1e5 * sqrtl(e / num));
    out[1] = (double)num;
}

/* Initial irradiance = per-pixel mean over all images, main_responseCalib.cpp:249-259. */
void oport_einit(const unsigned char* data, int n, int npix, double* E) {
    for (int k = 0; k < npix; k++) E[k] = 0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < npix; k++) E[k] += data[(size_t)i * npix + k];
    for (int k = 0; k < npix; k++) E[k] = E[k] / (double)n;
}

/* Rescale so that G[255] = 255, main_responseCalib.cpp:350-355.  Returns the factor. */
double oport_rescale(int npix, double* E, double* G) {
    double f = 255.0 / G[255];
    for (int i = 0; i < npix; i++) E[i] *= f;
    for (int i = 0; i < 256; i++) G[i] *= f;
    return f;
}

/* Saturation leak padding (3x3 dilation of 255, interior pixels only), one image,
 * `iters` rounds, main_responseCalib.cpp:212-236.  tmp: w*h bytes scratch. */
void oport_leak_padding(unsigned char* img, unsigned char* tmp, int w, int h, int iters) {
    for (int it = 0; it < iters; it++) {
        memcpy(tmp, img, (size_t)w * h);
        for (int y = 1; y < h - 1; y++)
            for (int x = 1; x < w - 1; x++)
                if (img[x + y * w] == 255)
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) tmp[x + dx + w * (y + dy)] = 255;
        memcpy(img, tmp, (size_t)w * h);
    }
}

/* ------------------------------------------------- vignetteCalib optimiser
 *
 * Restatement of the alternating optimisation of main_vignetteCalib.cpp:395-585 (inline in main(), no compilable unit; the
 * front end — aruco detection, homography, image loading — is not part of it).  Inputs as the reference holds them at :395:
 *   images [n][wI*hI] float (NaN = invalidated pixel), p2x/p2y [n][gw*gh] float plane-to-image maps (NaN = not visible; the
 *   reference guarantees 1 < x+0.5 < wI-2 for finite entries, :352-356, so all four bilinear taps are in bounds),
 *   plane_color [gw*gh], vignette [wI*hI].
 * One deviation: the reference reads plane_color UNINITIALISED in its first plane step (new float[] at :381, first use :425);
 * here the caller supplies the starting values.
 * Toolchain-dependent detail: the outlier test is `abs(residual) > oth2` on a double (:423, :481).  Where only `int abs(int)` is
 * visible to unqualified lookup (libstdc++ before GCC 6; also this repo's stand-in headers) the residual is TRUNCATED TO INT
 * first; with <cmath>'s overloads in the global namespace it is fabs.  int_abs selects the former (the author-era behaviour). */
static int oport_vc_outlier(double residual, int oth2, int int_abs) {
    if (!int_abs) return fabs(residual) > oth2;
    /* abs((int)residual) > oth2: residual is a square (>= 0) or NaN; out-of-range / NaN conversions give INT_MIN on x86 */
    if (!(residual < 2147483648.0)) return 0;
    return (int)residual > oth2;
}

/* getInterpolatedElement, main_vignetteCalib.cpp:52-70 */
static float oport_vc_interp(const float* mat, float x, float y, int width) {
    int ix = (int)x, iy = (int)y;
    float dx = x - ix, dy = y - iy, dxdy = dx * dy;
    const float* bp = mat + ix + iy * width;
    return dxdy * bp[1 + width] + (dy - dxdy) * bp[width] + (dx - dxdy) * bp[1] + (1 - dx - dy + dxdy) * bp[0];
}

/* "optimize planeColor", :400-446.  stats[0] = E, stats[1] = R.  ff/fc: gw*gh floats scratch (planeColorFF/FC). */
void oport_vc_plane_step(const float* images, const float* p2x, const float* p2y, int n, int gwgh, int wI, int hI,
                         const float* vignette, float* plane_color, int oth2, int int_abs, float* ff, float* fc, double stats[2]) {
    double E = 0, R = 0;
    size_t npx = (size_t)wI * hI;
    memset(ff, 0, (size_t)gwgh * sizeof(float));
    memset(fc, 0, (size_t)gwgh * sizeof(float));
    for (int img = 0; img < n; img++) {
        const float* mx = p2x + (size_t)img * gwgh;
        const float* my = p2y + (size_t)img * gwgh;
        const float* image = images + (size_t)img * npx;
        for (int pi = 0; pi < gwgh; pi++) {
            if (isnan(mx[pi])) continue;
            float color = oport_vc_interp(image, mx[pi], my[pi], wI);
            float fac = oport_vc_interp(vignette, mx[pi], my[pi], wI);
            if (isnan(fac)) continue;
            if (isnan(color)) continue;
            double residual = (double)((color - plane_color[pi] * fac) * (color - plane_color[pi] * fac));
            if (oport_vc_outlier(residual, oth2, int_abs)) { E += oth2; R++; continue; }
            ff[pi] += fac * fac;
            fc[pi] += color * fac;
            if (isnan(plane_color[pi])) continue;
            E += residual;
            R++;
        }
    }
    for (int pi = 0; pi < gwgh; pi++) plane_color[pi] = (ff[pi] < 1) ? NAN : fc[pi] / ff[pi];
    stats[0] = E; stats[1] = R;
}

/* "optimize vignette", :458-533 incl. the normalisation to max factor 1.  tt/ct: wI*hI floats scratch (vignetteFactorTT/CT). */
void oport_vc_vignette_step(const float* images, const float* p2x, const float* p2y, int n, int gwgh, int wI, int hI,
                            const float* plane_color, float* vignette, int oth2, int int_abs, float* tt, float* ct, double stats[2]) {
    double E = 0, R = 0;
    size_t npx = (size_t)wI * hI;
    memset(tt, 0, npx * sizeof(float));
    memset(ct, 0, npx * sizeof(float));
    for (int img = 0; img < n; img++) {
        const float* mx = p2x + (size_t)img * gwgh;
        const float* my = p2y + (size_t)img * gwgh;
        const float* image = images + (size_t)img * npx;
        for (int pi = 0; pi < gwgh; pi++) {
            if (isnan(mx[pi])) continue;
            float x = mx[pi], y = my[pi];
            float colorImage = oport_vc_interp(image, x, y, wI);
            float fac = oport_vc_interp(vignette, x, y, wI);
            float colorPlane = plane_color[pi];
            if (isnan(colorPlane)) continue;
            if (isnan(colorImage)) continue;
            double residual = (double)((colorImage - colorPlane * fac) * (colorImage - colorPlane * fac));
            if (oport_vc_outlier(residual, oth2, int_abs)) { E += oth2; R++; continue; }
            int ix = (int)x, iy = (int)y;
            float dx = x - ix, dy = y - iy, dxdy = dx * dy;
            size_t b = (size_t)ix + (size_t)iy * wI;
            tt[b] += (1 - dx - dy + dxdy) * colorPlane * colorPlane;
            tt[b + 1] += (dx - dxdy) * colorPlane * colorPlane;
            tt[b + wI] += (dy - dxdy) * colorPlane * colorPlane;
            tt[b + 1 + wI] += dxdy * colorPlane * colorPlane;
            ct[b] += (1 - dx - dy + dxdy) * colorImage * colorPlane;
            ct[b + 1] += (dx - dxdy) * colorImage * colorPlane;
            ct[b + wI] += (dy - dxdy) * colorImage * colorPlane;
            ct[b + 1 + wI] += dxdy * colorImage * colorPlane;
            if (isnan(fac)) continue;
            E += residual;
            R++;
        }
    }
    float maxFac = 0;
    for (size_t i = 0; i < npx; i++) {
        if (tt[i] < 1) vignette[i] = NAN;
        else {
            vignette[i] = ct[i] / tt[i];
            if (vignette[i] > maxFac) maxFac = vignette[i];
        }
    }
    for (size_t i = 0; i < npx; i++) vignette[i] /= maxFac;
    stats[0] = E; stats[1] = R;
}

/* "dilate & smoothe vignette", :542-566: `iters` rounds of a NaN-aware 3x3 mean.  out and tmp: wI*hI floats. */
void oport_vc_smooth(const float* vignette, int wI, int hI, int iters, float* out, float* tmp) {
    size_t npx = (size_t)wI * hI;
    memcpy(out, vignette, npx * sizeof(float));
    for (int it = 0; it < iters; it++) {
        memcpy(tmp, out, npx * sizeof(float));
        for (int y = 0; y < hI; y++)
            for (int x = 0; x < wI; x++) {
                int idx = x + y * wI;
                float sum = 0, num = 0;
                if (x < wI - 1 && y < hI - 1 && !isnan(tmp[idx + 1 + wI])) { sum += tmp[idx + 1 + wI]; num++; }
                if (x < wI - 1 && !isnan(tmp[idx + 1])) { sum += tmp[idx + 1]; num++; }
                if (x < wI - 1 && y > 0 && !isnan(tmp[idx + 1 - wI])) { sum += tmp[idx + 1 - wI]; num++; }
                if (y < hI - 1 && !isnan(tmp[idx + wI])) { sum += tmp[idx + wI]; num++; }
                if (!isnan(tmp[idx])) { sum += tmp[idx]; num++; }
                if (y > 0 && !isnan(tmp[idx - wI])) { sum += tmp[idx - wI]; num++; }
                if (y < hI - 1 && x > 0 && !isnan(tmp[idx - 1 + wI])) { sum += tmp[idx - 1 + wI]; num++; }
                if (x > 0 && !isnan(tmp[idx - 1])) { sum += tmp[idx - 1]; num++; }
                if (y > 0 && x > 0 && !isnan(tmp[idx - 1 - wI])) { sum += tmp[idx - 1 - wI]; num++; }
                if (num > 0) out[idx] = sum / num;
            }
    }
}

/* ---- vignetteCalib front end between the image reader and the optimiser (main_vignetteCalib.cpp:262-358), restated so that
 * the reference PROGRAM (oracle/_ref/vignetteCalib_ref, built against stand-in aruco/OpenCV/Eigen) can be replayed: the 3x3
 * float arithmetic is the stand-in's (oracle/shim/Eigen/Core), not Eigen's. */

/* plane grid -> undistorted image coordinates, :266-282.  H: the homography the program got from cv::findHomography (doubles,
 * narrowed to float at :253-261).  out: gw*gh floats each; distortCoordinates (:284) is applied by the caller afterwards. */
void oport_vc_plane_maps(const double H[9], int gw, int gh, float facw, float fach, float* X, float* Y) {
    float K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Ki[9], Hf[9], HK[9];
    K[0] = gw / facw; K[4] = gh / fach; K[2] = gw / 2; K[5] = gh / 2; K[8] = 1;      /* :196-200 (gw/2 is an int division) */
    {   /* inverse: adjugate / determinant, as the stand-in does it */
        const float c00 = K[4] * K[8] - K[5] * K[7], c01 = K[5] * K[6] - K[3] * K[8], c02 = K[3] * K[7] - K[4] * K[6];
        const float det = K[0] * c00 + K[1] * c01 + K[2] * c02;
        Ki[0] = c00 / det; Ki[3] = c01 / det; Ki[6] = c02 / det;
        Ki[1] = (K[2] * K[7] - K[1] * K[8]) / det;
        Ki[4] = (K[0] * K[8] - K[2] * K[6]) / det;
        Ki[7] = (K[1] * K[6] - K[0] * K[7]) / det;
        Ki[2] = (K[1] * K[5] - K[2] * K[4]) / det;
        Ki[5] = (K[2] * K[3] - K[0] * K[5]) / det;
        Ki[8] = (K[0] * K[4] - K[1] * K[3]) / det;
    }
    for (int i = 0; i < 9; i++) Hf[i] = (float)H[i];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) HK[r * 3 + c] = Hf[r * 3] * Ki[c] + Hf[r * 3 + 1] * Ki[3 + c] + Hf[r * 3 + 2] * Ki[6 + c];
    int idx = 0;
    for (int y = 0; y < gh; y++)
        for (int x = 0; x < gw; x++) {
            float v0 = (float)x, v1 = (float)y, v2 = 1.0f, pp[3];
            for (int r = 0; r < 3; r++) pp[r] = HK[r * 3] * v0 + HK[r * 3 + 1] * v1 + HK[r * 3 + 2] * v2;
            X[idx] = pp[0] / pp[2];
            Y[idx] = pp[1] / pp[2];
            idx++;
        }
}

/* points whose rounded position is not strictly inside the image interior become NaN, :345-357 */
void oport_vc_mask_maps(float* X, float* Y, int n, int wI, int hI) {
    for (int i = 0; i < n; i++) {
        int u_d = X[i] + 0.5;
        int v_d = Y[i] + 0.5;
        if (!(u_d > 1 && v_d > 1 && u_d < wI - 2 && v_d < hI - 2)) { X[i] = NAN; Y[i] = NAN; }
    }
}

/* exposure normalisation and gradient-based invalidation, :288-310 (in place and order-dependent, like the reference) */
void oport_vc_prepare_image(const float* raw, int wI, int hI, float meanExposure, float exposure, int maxAbsGrad, float* image) {
    for (int y = 0; y < hI; y++)
        for (int x = 0; x < wI; x++) image[x + y * wI] = meanExposure * raw[x + y * wI] / exposure;
    for (int y = 2; y < hI - 2; y++)
        for (int x = 2; x < wI - 2; x++)
            for (int deltax = -2; deltax < 3; deltax++)
                for (int deltay = -2; deltay < 3; deltay++)
                    if (fabsf(image[x + y * wI] - image[x + deltax + (y + deltay) * wI]) > maxAbsGrad) {
                        image[x + y * wI] = NAN;
                        image[x + deltax + (y + deltay) * wI] = NAN;
                    }
}

/* ------------------------------------------------- CPU timing loops (bench only) */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct {
    int tid, threads, n_frames, n_distinct, in_w, in_h, n_out, levels, flags;
    const float *rx, *ry, *ginv, *vinv;
    const unsigned char* frames;
} oport_job;

/* photo+rect (+pyramid) loop on `threads` workers; returns wall seconds.
 * out_w/out_h are needed for the pyramid; levels = 1 means level 0 only. */
typedef struct { oport_job j; int out_w, out_h; } oport_job2;

static void* oport_worker2(void* arg) {
    oport_job2* jj = (oport_job2*)arg;
    oport_job* j = &jj->j;
    int n_in = j->in_w * j->in_h;
    float* tmp = (float*)malloc(sizeof(float) * n_in);
    float* out = (float*)malloc(sizeof(float) * (size_t)j->n_out * 2);
    for (int f = j->tid; f < j->n_frames; f += j->threads) {
        const unsigned char* src = j->frames + (size_t)(f % j->n_distinct) * n_in;
        oport_unmap(j->ginv, j->vinv, src, tmp, n_in, j->flags & 1, (j->flags >> 1) & 1, (j->flags >> 2) & 1);
        oport_undistort_f32(j->rx, j->ry, j->in_w, j->n_out, tmp, out);
        float* lv = out; int w = jj->out_w, h = jj->out_h;
        for (int l = 1; l < j->levels; l++) {
            float* nx = lv + (size_t)w * h;
            oport_pyr_down(lv, w, h, nx);
            lv = nx; w >>= 1; h >>= 1;
        }
    }
    free(tmp); free(out);
    return 0;
}

double oport_time_frames(const float* rx, const float* ry, int in_w, int in_h, int out_w, int out_h,
                         const float* ginv, const float* vinv, const unsigned char* frames, int n_distinct,
                         int n_frames, int threads, int flags, int levels) {
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    oport_job2* jobs = (oport_job2*)malloc(sizeof(oport_job2) * threads);
    double t0 = now_s();
    for (int t = 0; t < threads; t++) {
        oport_job2* q = &jobs[t];
        q->j.tid = t; q->j.threads = threads; q->j.n_frames = n_frames; q->j.n_distinct = n_distinct;
        q->j.in_w = in_w; q->j.in_h = in_h; q->j.n_out = out_w * out_h; q->j.levels = levels; q->j.flags = flags;
        q->j.rx = rx; q->j.ry = ry; q->j.ginv = ginv; q->j.vinv = vinv; q->j.frames = frames;
        q->out_w = out_w; q->out_h = out_h;
        if (threads == 1) oport_worker2(q);
        else pthread_create(&th[t], 0, oport_worker2, q);
    }
    if (threads > 1) for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
    double t1 = now_s();
    free(th); free(jobs);
    return t1 - t0;
}
