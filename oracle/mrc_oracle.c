/*
 * mrc_oracle.c -- CPU restatement of the MRC page-decomposition hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA engine: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (archive-pdf-tools_b200/) never links, imports or calls it.
 *
 * Every function is an independent restatement (closed forms / integral images, not the
 * reference's running sums) of a step of internetarchive/archive-pdf-tools v1.5.7; the
 * reference location each one follows is cited.  Pinning status (see DESIGN.md "Oracle"):
 *   - orc_sauvola / orc_denoise / orc_optimise : pinned bit-exact against the reference's own
 *     Cython compiled from /root/reference (oracle/_ref, built by oracle/build_ref.py).
 *   - orc_rgb2gray / orc_resample_bicubic / orc_reduce : pinned bit-exact against Pillow 12.2.
 *   - orc_gauss_blur : pinned bit-exact against scipy 1.18 ndimage.gaussian_filter.
 *   - orc_estimate_sigma, orc_special_gray : third-party arithmetic (scikit-image, PyWavelets)
 *     that is NOT installed and NOT under /root/reference => "parity unpinned": restated from
 *     the published algorithms.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared mrc_oracle.c -o libmrc_oracle.so -lm
 *        (-ffp-contract=off: the reference's x86-64 build has no FMA; keep every double
 *         operation individually rounded.)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * A1  PIL image.convert('L') at internetarchivepdf/mrc.py:358-363 (Pillow Convert.c rgb2l,
 *     L24 macro): L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_rgb2gray(const uint8_t *rgb, int64_t npix, uint8_t *gray)
{
    for (int64_t i = 0; i < npix; i++) {
        uint32_t r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
        gray[i] = (uint8_t)((19595u * r + 38470u * g + 7471u * b + 0x8000u) >> 16);
    }
}

/* ------------------------------------------------------------------------------------------
 * A6  binarise_sauvola, cython/sauvola.pyx:29-222, as called by threshold_image
 *     (mrc.py:58-87, which inverts the result).  Closed form over the clamped window:
 *     rows [max(0,y-o+1), min(H,y+u+1)), cols [max(0,x-l+1), min(W,x+r+1)),
 *     l=(ww+1)/2, r=ww/2, o=(wh+1)/2, u=wh/2   (sauvola.pyx:70-73, 79-126, 128-217)
 *     m = (double)(S / n)  [C int division, cdivision(True)]           (sauvola.pyx:144)
 *     v = (double)(Q / n) - m*m                                         (sauvola.pyx:145)
 *     t = p + m*(k-1);  fg = t<=0 || t*t <= ((m*m)*k2)*v  (k>=0)         (sauvola.pyx:146-147)
 *                       fg = t<=0 && t*t >= ((m*m)*k2)*v  (k<0)          (sauvola.pyx:149-152)
 *     k2 = k*k/R/R                                                       (sauvola.pyx:60)
 *     out_fg[y*W+x] = fg   (== threshold_image()'s returned mask; the Cython itself
 *     stores !fg, sauvola.pyx:153).
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_sauvola(const uint8_t *in, uint8_t *out_fg, int W, int H,
                        int win_w, int win_h, double k, double R)
{
    if (W <= 0 || H <= 0) return 0;
    const int l = (win_w + 1) / 2, r = win_w / 2, o = (win_h + 1) / 2, u = win_h / 2;
    const double k2 = k * k / R / R;
    const double km1 = k - 1.0;
    size_t IW = (size_t)W + 1;
    uint64_t *IS = (uint64_t *)calloc(IW * ((size_t)H + 1), sizeof(uint64_t));
    uint64_t *IQ = (uint64_t *)calloc(IW * ((size_t)H + 1), sizeof(uint64_t));
    if (!IS || !IQ) { free(IS); free(IQ); return -1; }
    for (int y = 0; y < H; y++) {
        uint64_t rs = 0, rq = 0;
        for (int x = 0; x < W; x++) {
            uint64_t p = in[(size_t)y * W + x];
            rs += p; rq += p * p;
            IS[(size_t)(y + 1) * IW + x + 1] = IS[(size_t)y * IW + x + 1] + rs;
            IQ[(size_t)(y + 1) * IW + x + 1] = IQ[(size_t)y * IW + x + 1] + rq;
        }
    }
    for (int y = 0; y < H; y++) {
        int y0 = imax(0, y - o + 1), y1 = imin(H, y + u + 1);
        for (int x = 0; x < W; x++) {
            int x0 = imax(0, x - l + 1), x1 = imin(W, x + r + 1);
            int64_t n = (int64_t)(y1 - y0) * (x1 - x0);
            uint8_t fg;
            if (n <= 0) {
                /* window degenerates only for win sizes < 1; the reference would divide by 0 */
                fg = 0;
            } else {
                int64_t S = (int64_t)(IS[(size_t)y1 * IW + x1] - IS[(size_t)y0 * IW + x1]
                                      - IS[(size_t)y1 * IW + x0] + IS[(size_t)y0 * IW + x0]);
                int64_t Q = (int64_t)(IQ[(size_t)y1 * IW + x1] - IQ[(size_t)y0 * IW + x1]
                                      - IQ[(size_t)y1 * IW + x0] + IQ[(size_t)y0 * IW + x0]);
                double mean = (double)(S / n);
                double mm = mean * mean;
                double variance = (double)(Q / n) - mm;
                double pixel = (double)in[(size_t)y * W + x];
                double tmp = pixel + mean * km1;
                double rhs = mm * k2 * variance;
                if (k >= 0) fg = (tmp <= 0) || (tmp * tmp <= rhs);
                else        fg = (tmp <= 0) && (tmp * tmp >= rhs);
            }
            out_fg[(size_t)y * W + x] = fg;
        }
    }
    free(IS); free(IQ);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A8  fast_mask_denoise, cython/optimiser.pyx:436-472 (called mrc.py:388 with mincnt=4,
 *     n_size=2).  In-place raster order: a set pixel survives iff the (2n+1)^2 count of the
 *     array *being written*, minus itself, is >= mincnt.  Border of width n untouched.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_denoise(uint8_t *mask, int W, int H, int mincnt, int n)
{
    for (int y = n; y < H - n; y++)
        for (int x = n; x < W - n; x++) {
            uint8_t *c = mask + (size_t)y * W + x;
            if (!*c) continue;
            int cnt = 0;
            for (int dy = -n; dy <= n; dy++)
                for (int dx = -n; dx <= n; dx++)
                    cnt += c[(ptrdiff_t)dy * W + dx];
            *c = (uint8_t)((cnt - 1) >= mincnt);
        }
}

/* ------------------------------------------------------------------------------------------
 * A9  optimise_gray/rgb and optimise_gray2/rgb2, cython/optimiser.pyx:22-76, 83-146,
 *     153-273, 280-429 (all four are the same function; the "2" variants are incremental).
 *     out = copy(img); raster order; for every pixel NOT in mask:
 *        box  = [max(0,y-n), min(H,y+n)) x [max(0,x-n), min(W,x+n))
 *        num  = sum_{box, mask} img  +  sum_{[ys,y) x [xs,x)} out        (optimiser.pyx:57-70)
 *        den  = #mask in box + (y-ys)(x-xs)
 *        out  = den > 0 ? num / den : 0          (C truncating division, optimiser.pyx:72-75)
 *     Restated with integral images: static ones for mask*img and mask, and a causal one over
 *     `out` that is extended one finished row at a time.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_optimise(const uint8_t *mask, const uint8_t *img, int W, int H, int C, int n,
                         uint8_t *out)
{
    if (W <= 0 || H <= 0) return 0;
    size_t IW = (size_t)W + 1, IH = (size_t)H + 1;
    int64_t *IM = (int64_t *)calloc(IW * IH, sizeof(int64_t));        /* mask count */
    int64_t *IF = (int64_t *)calloc(IW * IH * C, sizeof(int64_t));    /* mask*img, per chan */
    int64_t *IO = (int64_t *)calloc(IW * IH * C, sizeof(int64_t));    /* out (causal) */
    if (!IM || !IF || !IO) { free(IM); free(IF); free(IO); return -1; }
    for (int y = 0; y < H; y++) {
        int64_t rm = 0, rf[4] = {0, 0, 0, 0};
        for (int x = 0; x < W; x++) {
            int m = mask[(size_t)y * W + x] != 0;
            rm += m;
            IM[(size_t)(y + 1) * IW + x + 1] = IM[(size_t)y * IW + x + 1] + rm;
            for (int c = 0; c < C; c++) {
                if (m) rf[c] += img[((size_t)y * W + x) * C + c];
                IF[((size_t)(y + 1) * IW + x + 1) * C + c] = IF[((size_t)y * IW + x + 1) * C + c] + rf[c];
            }
        }
    }
    memcpy(out, img, (size_t)W * H * C);
    for (int y = 0; y < H; y++) {
        int ys = imax(0, y - n), ye = imin(H, y + n);
        for (int x = 0; x < W; x++) {
            if (mask[(size_t)y * W + x]) continue;
            int xs = imax(0, x - n), xe = imin(W, x + n);
            int64_t den = IM[(size_t)ye * IW + xe] - IM[(size_t)ys * IW + xe]
                        - IM[(size_t)ye * IW + xs] + IM[(size_t)ys * IW + xs]
                        + (int64_t)(y - ys) * (x - xs);
            for (int c = 0; c < C; c++) {
                int64_t fir = IF[((size_t)ye * IW + xe) * C + c] - IF[((size_t)ys * IW + xe) * C + c]
                            - IF[((size_t)ye * IW + xs) * C + c] + IF[((size_t)ys * IW + xs) * C + c];
                int64_t iir = IO[((size_t)y * IW + x) * C + c] - IO[((size_t)ys * IW + x) * C + c]
                            - IO[((size_t)y * IW + xs) * C + c] + IO[((size_t)ys * IW + xs) * C + c];
                out[((size_t)y * W + x) * C + c] = den > 0 ? (uint8_t)((fir + iir) / den) : 0;
            }
        }
        /* row y of `out` is final: extend the causal integral to row y+1 */
        int64_t ro[4] = {0, 0, 0, 0};
        for (int x = 0; x < W; x++)
            for (int c = 0; c < C; c++) {
                ro[c] += out[((size_t)y * W + x) * C + c];
                IO[((size_t)(y + 1) * IW + x + 1) * C + c] = IO[((size_t)y * IW + x + 1) * C + c] + ro[c];
            }
    }
    free(IM); free(IF); free(IO);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A4  scipy.ndimage.gaussian_filter(imgf, sigma) at mrc.py:311 on a float32 image (default
 *     order=0, mode='reflect', truncate=4.0), followed by .astype(uint8) (mrc.py:325).
 *     scipy: radius = int(4*sigma + 0.5); weights exp(-0.5 j^2/sigma^2) normalised to sum 1 in
 *     double; correlate1d along axis 0 then axis 1, symmetric-kernel fast path
 *     (ni_filters.c NI_Correlate1D): acc = x[i]*w0; for j = radius..1: acc += (x[i-j]+x[i+j])*w_j,
 *     double accumulate, float32 store after each axis.  'reflect' = d c b a | a b c d | d c b a.
 *     in/out: float32 H x W.  radius 0 => copy.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_gauss_radius(double sigma) { return (int)(4.0 * sigma + 0.5); }

ORC_API void orc_gauss_weights(double sigma, int radius, double *w /* radius+1, w[j] for |offset|=j */)
{
    double sum = 0.0;
    /* scipy _gaussian_kernel1d: x = arange(-radius, radius+1); phi = exp(-0.5/sigma2 * x**2); phi /= phi.sum() */
    double sigma2 = sigma * sigma;
    for (int j = -radius; j <= radius; j++) sum += exp(-0.5 / sigma2 * (double)(j * j));
    for (int j = 0; j <= radius; j++) w[j] = exp(-0.5 / sigma2 * (double)(j * j)) / sum;
}

static inline int reflect_idx(int i, int n)
{
    /* scipy NI_EXTEND_REFLECT (half-sample symmetric), valid for any offset */
    if (n == 1) return 0;
    int p = 2 * n;
    i %= p; if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

ORC_API int orc_gauss_blur(const float *in, float *out, int W, int H, double sigma)
{
    int radius = orc_gauss_radius(sigma);
    if (radius < 1) { memcpy(out, in, sizeof(float) * (size_t)W * H); return 0; }
    double *w = (double *)malloc(sizeof(double) * (radius + 1));
    float *tmp = (float *)malloc(sizeof(float) * (size_t)W * H);
    if (!w || !tmp) { free(w); free(tmp); return -1; }
    orc_gauss_weights(sigma, radius, w);
    for (int y = 0; y < H; y++)        /* axis 0 */
        for (int x = 0; x < W; x++) {
            double acc = (double)in[(size_t)y * W + x] * w[0];
            for (int j = radius; j >= 1; j--)
                acc += ((double)in[(size_t)reflect_idx(y - j, H) * W + x]
                      + (double)in[(size_t)reflect_idx(y + j, H) * W + x]) * w[j];
            tmp[(size_t)y * W + x] = (float)acc;
        }
    for (int y = 0; y < H; y++)        /* axis 1 */
        for (int x = 0; x < W; x++) {
            double acc = (double)tmp[(size_t)y * W + x] * w[0];
            for (int j = radius; j >= 1; j--)
                acc += ((double)tmp[(size_t)y * W + reflect_idx(x - j, W)]
                      + (double)tmp[(size_t)y * W + reflect_idx(x + j, W)]) * w[j];
            out[(size_t)y * W + x] = (float)acc;
        }
    free(w); free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A3  estimate_noise (mrc.py:273-296) -> mean_estimate_sigma (mrc.py:52-55) -> scikit-image
 *     restoration.estimate_sigma -> PyWavelets dwtn(..., 'db2')['dd'].   PARITY UNPINNED:
 *     scikit-image / PyWavelets are not installed and not under /root/reference; restated from
 *     their published algorithm:
 *       dd = single-level 2-D DWT diagonal detail, mode 'symmetric', float32 arithmetic for
 *            float32 input: per axis (0 then 1) out[o] = sum_{j=0..3} dec_hi[j] * ext(x)[2o+1-j],
 *            o in [0,(N+3)/2), accumulated in float32 in j order, ext = half-sample symmetric.
 *       sigma = median(|dd[dd != 0]|) / 0.6744897501960817   (np.median: float32 mean of the two
 *            middle values for an even count; division in double)
 *     crop: rows [int(h/2-h/4), int(h/2+h/4)), cols likewise; whole image if he==0 or we==0.
 * ---------------------------------------------------------------------------------------- */
static const double ORC_DB2_DEC_HI[4] = {
    -0.48296291314469025, 0.836516303737469, -0.22414386804185735, -0.12940952255092145 };

static inline int sym_idx(int i, int n)
{
    /* PyWavelets MODE_SYMMETRIC: ... x1 x0 | x0 x1 ... x(n-1) | x(n-1) x(n-2) ..., repeated */
    if (n == 1) return 0;
    int p = 2 * n;
    i %= p; if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

ORC_API void orc_noise_crop(int W, int H, int *hs, int *he, int *ws, int *we)
{
    /* mrc.py:278-292 ; python float division then int() truncation */
    *hs = (int)((double)H / 2 - (double)H / 4);
    *he = (int)((double)H / 2 + (double)H / 4);
    *ws = (int)((double)W / 2 - (double)W / 4);
    *we = (int)((double)W / 2 + (double)W / 4);
    if (*he == 0 || *we == 0) { *hs = 0; *he = H; *ws = 0; *we = W; }
}

/* gray: uint8 H x W (the float32 image of mrc.py:372 holds exactly these integers) */
ORC_API double orc_estimate_sigma_crop(const uint8_t *gray, int W, int hs, int he, int ws, int we)
{
    int h = he - hs, w = we - ws;
    if (h <= 0 || w <= 0) return NAN;
    float f[4];
    for (int j = 0; j < 4; j++) f[j] = (float)ORC_DB2_DEC_HI[j];
    int oh = (h + 3) / 2, ow = (w + 3) / 2;
    float *d0 = (float *)malloc(sizeof(float) * (size_t)oh * w);
    float *dd = (float *)malloc(sizeof(float) * (size_t)oh * ow);
    if (!d0 || !dd) { free(d0); free(dd); return NAN; }
    for (int o = 0; o < oh; o++)            /* axis 0: detail along rows */
        for (int x = 0; x < w; x++) {
            float sum = 0.0f;
            for (int j = 0; j < 4; j++) {
                int yy = sym_idx(2 * o + 1 - j, h);
                float p = (float)gray[(size_t)(hs + yy) * W + ws + x];
                float prod = f[j] * p;      /* no FMA: -ffp-contract=off */
                sum = sum + prod;
            }
            d0[(size_t)o * w + x] = sum;
        }
    for (int y = 0; y < oh; y++)            /* axis 1: detail along cols */
        for (int o = 0; o < ow; o++) {
            float sum = 0.0f;
            for (int j = 0; j < 4; j++) {
                float prod = f[j] * d0[(size_t)y * w + sym_idx(2 * o + 1 - j, w)];
                sum = sum + prod;
            }
            dd[(size_t)y * ow + o] = sum;
        }
    size_t cnt = 0, tot = (size_t)oh * ow;
    for (size_t i = 0; i < tot; i++)
        if (dd[i] != 0.0f) dd[cnt++] = fabsf(dd[i]);
    double sigma;
    if (cnt == 0) {
        sigma = NAN;                         /* np.median([]) -> nan */
    } else {
        qsort(dd, cnt, sizeof(float), cmp_float);
        float med;
        if (cnt & 1) med = dd[cnt / 2];
        else { float s = dd[cnt / 2 - 1] + dd[cnt / 2]; med = s / 2.0f; }   /* float32 mean of two */
        sigma = (double)med / 0.6744897501960817;
    }
    free(d0); free(dd);
    return sigma;
}


/* mean_estimate_sigma on a BOOLEAN crop (mrc.py:253-254: estimate_sigma(thres)).  PyWavelets promotes a
 * bool array to float64, so this is the same db2 'dd' band computed in double (True = 1.0) and a float64
 * median.  img01: H x W bytes holding 0/1, row pitch `pitch`.  Third-party algorithm restated from its
 * published form ("parity unpinned", like orc_estimate_sigma_crop). */
static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

ORC_API double orc_estimate_sigma_bool(const uint8_t *img01, int W, int H, int pitch)
{
    if (H <= 0 || W <= 0) return NAN;
    int oh = (H + 3) / 2, ow = (W + 3) / 2;
    double *d0 = (double *)malloc(sizeof(double) * (size_t)oh * W);
    double *dd = (double *)malloc(sizeof(double) * (size_t)oh * ow);
    if (!d0 || !dd) { free(d0); free(dd); return NAN; }
    for (int o = 0; o < oh; o++)
        for (int x = 0; x < W; x++) {
            double sum = 0.0;
            for (int j = 0; j < 4; j++) {
                int yy = sym_idx(2 * o + 1 - j, H);
                double p = img01[(size_t)yy * pitch + x] ? 1.0 : 0.0;
                double prod = ORC_DB2_DEC_HI[j] * p;
                sum = sum + prod;
            }
            d0[(size_t)o * W + x] = sum;
        }
    for (int y = 0; y < oh; y++)
        for (int o = 0; o < ow; o++) {
            double sum = 0.0;
            for (int j = 0; j < 4; j++) {
                double prod = ORC_DB2_DEC_HI[j] * d0[(size_t)y * W + sym_idx(2 * o + 1 - j, W)];
                sum = sum + prod;
            }
            dd[(size_t)y * ow + o] = sum;
        }
    size_t cnt = 0, tot = (size_t)oh * ow;
    for (size_t i = 0; i < tot; i++)
        if (dd[i] != 0.0) dd[cnt++] = fabs(dd[i]);
    double sigma;
    if (cnt == 0) sigma = NAN;
    else {
        qsort(dd, cnt, sizeof(double), cmp_double);
        double med = (cnt & 1) ? dd[cnt / 2] : (dd[cnt / 2 - 1] + dd[cnt / 2]) / 2.0;
        sigma = med / 0.6744897501960817;
    }
    free(d0); free(dd);
    return sigma;
}

ORC_API double orc_estimate_noise(const uint8_t *gray, int W, int H)
{
    int hs, he, ws, we;
    orc_noise_crop(W, H, &hs, &he, &ws, &we);
    return orc_estimate_sigma_crop(gray, W, hs, he, ws, we);
}

/* ------------------------------------------------------------------------------------------
 * A10  PIL Image.thumbnail at mrc.py:420-434 / 454-468 (Pillow: Image.py thumbnail/resize,
 *      Reduce.c, Resample.c).  Host size logic lives in oracle/oracle.py; these are the two
 *      pixel passes.  8-bit fixed point, PRECISION_BITS = 22, horizontal pass then vertical
 *      pass with a uint8 intermediate, BICUBIC a = -0.5, support 2.
 * ---------------------------------------------------------------------------------------- */
static double bicubic_filter(double x)
{
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

static double lanczos_filter(double x)
{
    /* Pillow Resample.c: sinc(x) * sinc(x/3) on [-3,3) */
    if (x < 0.0) x = -x;
    if (x >= 3.0) return 0.0;
    if (x == 0.0) return 1.0;
    double a = x * M_PI, b = a / 3.0;
    return (sin(a) / a) * (sin(b) / b);
}

/* Pillow precompute_coeffs + normalize_coeffs_8bpc.  bounds: 2*outSize ints (xmin, count);
 * kk: outSize*ksize ints.  Returns ksize.  filter: 0 = BICUBIC, 1 = LANCZOS. */
ORC_API int orc_resample_ksize(int inSize, float in0, float in1, int outSize, int filter)
{
    (void)inSize;
    double support0 = filter == 1 ? 3.0 : 2.0;
    double filterscale = (double)(in1 - in0) / outSize;
    if (filterscale < 1.0) filterscale = 1.0;
    return (int)ceil(support0 * filterscale) * 2 + 1;
}

ORC_API int orc_resample_coeffs(int inSize, float in0, float in1, int outSize, int filter,
                                int *bounds, int *kk)
{
    double support0 = filter == 1 ? 3.0 : 2.0;
    double scale, filterscale;
    filterscale = scale = (double)(in1 - in0) / outSize;
    if (filterscale < 1.0) filterscale = 1.0;
    double support = support0 * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    double *k = (double *)malloc(sizeof(double) * ksize);
    if (!k) return -1;
    for (int xx = 0; xx < outSize; xx++) {
        double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0, ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > inSize) xmax = inSize;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; x++) {
            double arg = (x + xmin - center + 0.5) * ss;
            double w = filter == 1 ? lanczos_filter(arg) : bicubic_filter(arg);
            k[x] = w; ww += w;
        }
        for (x = 0; x < xmax; x++) if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; x++) k[x] = 0;
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
        for (x = 0; x < ksize; x++) {
            double v = k[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
        }
    }
    free(k);
    return ksize;
}

static inline uint8_t clip8(int v)
{
    v >>= 22;                      /* arithmetic shift, like Pillow's clip8 lookup */
    return v < 0 ? 0 : v > 255 ? 255 : (uint8_t)v;
}

/* Two-pass resample of an interleaved uint8 image (C channels), box = (x0f,y0f,x1f,y1f). */
ORC_API int orc_resample(const uint8_t *in, int W, int H, int C, uint8_t *out, int OW, int OH,
                         float bx0, float by0, float bx1, float by1, int filter)
{
    int need_h = OW != W || bx0 != 0 || bx1 != W;
    int need_v = OH != H || by0 != 0 || by1 != H;
    int kh = orc_resample_ksize(W, bx0, bx1, OW, filter), kv = orc_resample_ksize(H, by0, by1, OH, filter);
    int *bh = (int *)malloc(sizeof(int) * 2 * OW), *bv = (int *)malloc(sizeof(int) * 2 * OH);
    int *ch = (int *)malloc(sizeof(int) * (size_t)OW * kh), *cv = (int *)malloc(sizeof(int) * (size_t)OH * kv);
    uint8_t *tmp = (uint8_t *)malloc((size_t)OW * H * C);
    if (!bh || !bv || !ch || !cv || !tmp) { free(bh); free(bv); free(ch); free(cv); free(tmp); return -1; }
    orc_resample_coeffs(W, bx0, bx1, OW, filter, bh, ch);
    orc_resample_coeffs(H, by0, by1, OH, filter, bv, cv);
    if (need_h) {
        for (int y = 0; y < H; y++)
            for (int xx = 0; xx < OW; xx++)
                for (int c = 0; c < C; c++) {
                    int ss = 1 << 21;
                    for (int x = 0; x < bh[2 * xx + 1]; x++)
                        ss += in[((size_t)y * W + x + bh[2 * xx]) * C + c] * ch[(size_t)xx * kh + x];
                    tmp[((size_t)y * OW + xx) * C + c] = clip8(ss);
                }
    } else {
        memcpy(tmp, in, (size_t)W * H * C);
    }
    if (need_v) {
        for (int yy = 0; yy < OH; yy++)
            for (int x = 0; x < OW; x++)
                for (int c = 0; c < C; c++) {
                    int ss = 1 << 21;
                    for (int y = 0; y < bv[2 * yy + 1]; y++)
                        ss += tmp[((size_t)(y + bv[2 * yy]) * OW + x) * C + c] * cv[(size_t)yy * kv + y];
                    out[((size_t)yy * OW + x) * C + c] = clip8(ss);
                }
    } else {
        memcpy(out, tmp, (size_t)OW * OH * C);
    }
    free(bh); free(bv); free(ch); free(cv); free(tmp);
    return 0;
}

/* Pillow ImagingReduce (Reduce.c) for a box anchored at (bx,by) of size (bw,bh), factors
 * (fx,fy): out = ((sum + amend) * multiplier) >> 24 with amend = cells/2 and
 * multiplier = division_UINT32(cells, 8) = (uint32)((1<<24) * ... ) -- see oracle tests: the
 * generic formula below is pinned against Pillow for all factors 1..6 on full and partial cells.
 * Edge cells that are only partly inside the box average just the covered pixels. */
static inline uint32_t division_u32(int divider, int result_bits)
{
    uint32_t max_dividend = (1u << result_bits) * (uint32_t)divider;
    float max_int = (1 << 30) * 4.0f;
    return (uint32_t)(max_int / max_dividend);
}

ORC_API void orc_reduce(const uint8_t *in, int W, int H, int C, int bx, int by, int bw, int bh,
                        int fx, int fy, uint8_t *out /* ((bw+fx-1)/fx) x ((bh+fy-1)/fy) */)
{
    (void)H;
    int OW = (bw + fx - 1) / fx, OH = (bh + fy - 1) / fy;
    for (int oy = 0; oy < OH; oy++) {
        int y0 = oy * fy, y1 = imin(bh, y0 + fy);
        for (int ox = 0; ox < OW; ox++) {
            int x0 = ox * fx, x1 = imin(bw, x0 + fx);
            int cells = (y1 - y0) * (x1 - x0);
            uint32_t mult = division_u32(cells, 8), amend = (uint32_t)cells / 2;
            for (int c = 0; c < C; c++) {
                uint32_t ss = amend;
                for (int y = y0; y < y1; y++)
                    for (int x = x0; x < x1; x++)
                        ss += in[((size_t)(by + y) * W + bx + x) * C + c];
                out[((size_t)oy * OW + ox) * C + c] = (uint8_t)((ss * mult) >> 24);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * A12  special_gray_convert, internetarchivepdf/grayconvert.py:38-66 (level_arr :24-31) and
 *      scikit-image color.rgb2hsv (NOT installed: "parity unpinned", restated).
 *      Stage 1 (stats -> thresholds) is done by the caller in Python exactly like the
 *      reference; this is the per-pixel stage:
 *        level:  v = (uint8)((x - minv)/interval), 0 if x<minv, 255 if x>maxv      (:24-31)
 *        a = v * (1/255.)  [skimage img_as_float];  V = max a, delta = max a - min a,
 *        S = delta==0 ? 0 : delta / V;   l = V * (1 - S/2);  out = (uint8)(l*255)   (:63-66)
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_special_gray_pixels(const uint8_t *rgb, int64_t npix, const double *minv,
                                     const double *maxv, uint8_t *out)
{
    const double inv255 = 1.0 / 255.0;
    for (int64_t i = 0; i < npix; i++) {
        double a[3];
        for (int c = 0; c < 3; c++) {
            double x = (double)rgb[3 * i + c];
            double interval = (maxv[c] / 255.) - (minv[c] / 255.);
            uint8_t v;
            if (x < minv[c]) v = 0;
            else if (x > maxv[c]) v = 255;
            else { double q = (x - minv[c]) / interval; v = (uint8_t)(int)q; }
            a[c] = (double)v * inv255;
        }
        double mx = a[0] > a[1] ? a[0] : a[1]; if (a[2] > mx) mx = a[2];
        double mn = a[0] < a[1] ? a[0] : a[1]; if (a[2] < mn) mn = a[2];
        double delta = mx - mn;
        double s = delta == 0.0 ? 0.0 : delta / mx;
        double l = mx * (1 - (s / 2));
        out[i] = (uint8_t)(int)(l * 255);
    }
}
