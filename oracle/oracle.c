/*
 * oracle.c -- CPU restatement of the reference's per-channel snowfall path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED against the reference itself.  tools/make_golden.py runs the unmodified reference
 * (/root/reference, imported with the three shims of oracle/ref_harness.py) and this restatement on the same seeded
 * inputs; tests/test_oracle_golden.py re-checks the committed fixtures (tests/golden/) bit-for-bit.
 *
 * What is restated (reference file:line, all under tools/snowfall/):
 *   process_single_channel   simulation.py:50-194
 *   get_occlusions           simulation.py:298-424
 *   compute_occlusion_dict   simulation.py:231-295   (+ binary_angle_search :197-228)
 *   received_power / xsi     simulation.py:547-569
 *   geometry.angles_to_lines :83-110, distances_of_points_to_lines :113-135, tangents_from_origin :138-190,
 *   tangent_lines_to_tangent_angles :32-80, do_angles_intersect_particles :193-223,
 *   tangent_angles_to_interval_angles :14-29
 *
 * The algorithm is the reference's own: for every beam, a pass over ALL particles of the channel's plane
 * (O(beams x particles)), the three-way disk/wedge test, tangent angles from the tangent-line coefficients,
 * nearest-first claiming of elementary angular sub-intervals, the 1230-sample waveform and its argmax.
 * dtype behaviour follows NumPy 2 (NEP 50) exactly as the reference executes in this image:
 *   - range d and azimuth theta are float32; beam limits are float64 (theta32 -> f64, -/+ radians(div/2));
 *   - the hard target's r_j is float32, so r_j*10, r_j+c*tau and r_j**2 are float32 operations;
 *   - snowflake r_j are float64.
 * Compile with -ffp-contract=off: NumPy never fuses a*b+c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.141592653589793
#define ORC_OK 0
#define ORC_ERR_INDEX 1      /* IndexError: waveform index >= 1230 (range beyond ~120 m), simulation.py:149 */
#define ORC_ERR_NEGATIVE 2   /* AssertionError: new intensity negative, simulation.py:184 */
#define ORC_ERR_TANGENT 3    /* ValueError in geometry.py:72 (not exactly one correct ray) */
#define ORC_ERR_ALLOC 4

#define M_EXT 1230

typedef struct {
    double a1, a2, dist;
} interval_t;

/* numpy add.reduce on a contiguous float64 vector (pairwise sum, 8-way unrolled for n >= 8) */
static double np_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res += a[i];
        return 0.0 + res;
    } else if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return 0.0 + res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum(a, n2) + np_sum(a + n2, n - n2);
    }
}

static double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* geometry.do_angles_intersect_particles (geometry.py:193-223) for one angle / one particle azimuth */
static int ray_on_particle_side(double angle, double phi)
{
    double diff = angle - phi;
    return (fabs(diff) < ORC_PI / 2) || (fabs(diff - 2 * ORC_PI) < ORC_PI / 2) || (fabs(diff + 2 * ORC_PI) < ORC_PI / 2);
}

/* geometry.angles_to_lines (geometry.py:83-110) */
static void angle_to_line(double angle, double *a, double *b)
{
    if (angle == ORC_PI / 2 || angle == 3 * ORC_PI / 2) { *a = 1.0; *b = 0.0; }
    else { *a = -tan(angle); *b = 1.0; }
}

/* geometry.tangents_from_origin (:138-190) + tangent_lines_to_tangent_angles (:32-80) for one particle */
static int tangent_angles(double x, double y, double r, double phi, double out[2])
{
    double a[2], b[2];
    double disc = r * sqrt(x * x + y * y - r * r);
    if (fabs(x) - r == 0) {                      /* one tangent is vertical */
        a[0] = 1.0; b[0] = 0.0;
        a[1] = (y * y - x * x) / (2 * x * y); b[1] = -1.0;
    } else {
        a[0] = (-x * y + disc) / (r * r - x * x);
        a[1] = (-x * y - disc) / (r * r - x * x);
        b[0] = -1.0; b[1] = -1.0;
    }
    for (int i = 0; i < 2; i++) {
        double ray1 = atan(-a[i] / b[i]);
        double ray2 = ray1 + ORC_PI;
        if (ray1 < 0) ray1 = ray1 + 2 * ORC_PI;
        ray1 = fabs(ray1);
        if (b[i] == 0) { ray1 = ORC_PI / 2; ray2 = 3 * ORC_PI / 2; }
        double d1 = ray1 - phi, d2 = ray2 - phi;
        int c1 = (fabs(d1) < ORC_PI / 2) || (fabs(d1 - 2 * ORC_PI) < ORC_PI / 2) || (fabs(d1 + 2 * ORC_PI) < ORC_PI / 2);
        int c2 = (fabs(d2) < ORC_PI / 2) || (fabs(d2 - 2 * ORC_PI) < ORC_PI / 2) || (fabs(d2 + 2 * ORC_PI) < ORC_PI / 2);
        if (c1 == c2) return ORC_ERR_TANGENT;
        out[i] = c1 ? ray1 : ray2;
    }
    if (out[0] > out[1]) { double t = out[0]; out[0] = out[1]; out[1] = t; }     /* angles.sort(axis=1) */
    if (out[1] - out[0] > ORC_PI) { double t = out[0]; out[0] = out[1]; out[1] = t; }   /* swap across the seam */
    return ORC_OK;
}

/* simulation.py:553-569, argument already float64 */
static double xsi64(double R)
{
    if (R <= 0.9) return 0.0;
    if (R >= 1.0) return 1.0;
    double m = (1 - 0) / (1.0 - 0.9);
    double b = 0 - (m * 0.9);
    return m * R + b;
}

/* same with a numpy float32 argument: python floats are weak, so the comparisons and m*R+b run in float32 */
static double xsi32(float R)
{
    if (R <= (float)0.9) return 0.0;
    if (R >= (float)1.0) return 1.0;
    double m = (1 - 0) / (1.0 - 0.9);
    double b = 0 - (m * 0.9);
    float y = (float)m * R;
    y = y + (float)b;
    return (double)y;
}

typedef struct {
    int n;                 /* number of dict entries incl. key -1 (last) */
    double *r;             /* r_j (particles: float64; hard target: float32 value widened) */
    double *ratio;
} occl_t;

/*
 * compute_occlusion_dict (simulation.py:231-295).  `iv` holds L intervals sorted by distance (modified in place for
 * the seam case exactly like the reference mutates `intervals`).  Returns number of claiming particles; their
 * (dist, ratio) are written to out_r/out_ratio in dict order, followed by the hard-target entry.
 */
static int occlusion_dict(double right, double left, interval_t *iv, int L, double current_range,
                          double beam_divergence_deg, double *out_r, double *out_ratio, int *n_out,
                          double *work /* >= 3*(2L+2) doubles */, int *iwork /* >= 2L+2 ints */)
{
    if (right > left) {
        right = right - 2 * ORC_PI;
        for (int j = 0; j < L; j++)
            if (iv[j].a1 > iv[j].a2) iv[j].a1 = iv[j].a1 - 2 * ORC_PI;
    }
    double *ep = work;
    int ne = 0;
    ep[ne++] = right;
    for (int j = 0; j < L; j++) { ep[ne++] = iv[j].a1; ep[ne++] = iv[j].a2; }
    ep[ne++] = left;
    qsort(ep, ne, sizeof(double), cmp_double);
    int nu = 0;                                      /* sorted(set(...)) */
    for (int k = 0; k < ne; k++)
        if (nu == 0 || ep[k] != ep[nu - 1]) ep[nu++] = ep[k];
    ne = nu;
    int nd = ne - 1;
    double *diffs = work + (2 * L + 2);
    double *sel = diffs + (2 * L + 2);
    int *assign = iwork;
    for (int k = 0; k < nd; k++) { diffs[k] = ep[k + 1] - ep[k]; assign[k] = -1; }
    double div_rad = beam_divergence_deg * (ORC_PI / 180.0);        /* np.radians */
    int cnt = 0;
    for (int j = 0; j < L; j++) {
        int i1 = -1, i2 = -1;
        for (int k = 0; k < ne; k++) { if (ep[k] == iv[j].a1) i1 = k; if (ep[k] == iv[j].a2) i2 = k; }
        int made = 0;
        for (int k = i1; k < i2; k++)
            if (assign[k] == -1) { assign[k] = j; made = 1; }
        if (made) {
            int ns = 0;
            for (int k = 0; k < nd; k++) if (assign[k] == j) sel[ns++] = diffs[k];
            double ratio = np_sum(sel, ns) / div_rad;
            out_r[cnt] = iv[j].dist;
            out_ratio[cnt] = clipd(ratio, 0, 1);
            cnt++;
        }
    }
    int ns = 0;
    for (int k = 0; k < nd; k++) if (assign[k] == -1) sel[ns++] = diffs[k];
    double ratio = np_sum(sel, ns) / div_rad;
    out_r[cnt] = current_range;
    out_ratio[cnt] = clipd(ratio, 0, 1);
    *n_out = cnt + 1;
    return cnt;
}

/*
 * One LiDAR channel (process_single_channel, simulation.py:50-194).
 *   x,y,z,intensity : float32[M]   points of this channel (reference order)
 *   theta_in        : float32[M] or NULL.  NULL -> atan2f(y,x) of this host's libm.  (The reference's float32
 *                     np.arctan2 is host/SIMD dependent; golden fixtures carry its bits.)
 *   particles       : float64[Np*3] (x, y, r) -- the plane file of this channel
 *   R               : float64[1230] range grid np.round(np.linspace(0, 120+c*tau, 1230), 2)  (simulation.py:116)
 *   out             : float32[M*5] (x, y, z, intensity, label)
 *   stat_i          : int32[M] optional, number of claiming occluders per beam
 */
int orc_snow_channel(int M, const float *x, const float *y, const float *z, const float *intensity,
                     const float *theta_in, int Np, const double *particles, double beam_divergence_deg,
                     double focal_distance_m, double focal_slope, double min_intensity, double max_intensity,
                     const double *R, float *out, double *intensity_diff_sum, int32_t *n_occluders,
                     float *theta_out)
{
    const double c_light = 299792458.0;
    const double tau_h = 1e-8;
    const double ctau = c_light * tau_h;
    const int ipm = 10;
    const double beta_0 = 1 * 1e-06 / ORC_PI;
    double focal_distance = focal_distance_m * 100;
    double focal_offset = (1 - focal_distance / 13100);
    focal_offset = focal_offset * focal_offset;
    double half = (beam_divergence_deg / 2) * (ORC_PI / 180.0);      /* np.radians(beam_divergence / 2) */

    double *prange = (double *)malloc(sizeof(double) * (Np > 0 ? Np : 1));
    if (!prange) return ORC_ERR_ALLOC;
    for (int p = 0; p < Np; p++) {
        double px = particles[3 * p], py = particles[3 * p + 1];
        prange[p] = sqrt(px * px + py * py);                          /* np.linalg.norm([x, y], axis=0) */
    }
    int cap = 64;
    interval_t *iv = (interval_t *)malloc(sizeof(interval_t) * cap);
    double *work = (double *)malloc(sizeof(double) * 3 * (2 * cap + 2));
    int *iwork = (int *)malloc(sizeof(int) * (2 * cap + 2));
    double *dr = (double *)malloc(sizeof(double) * (cap + 1));
    double *dratio = (double *)malloc(sizeof(double) * (cap + 1));
    double wave[M_EXT];
    double sum_diff = 0.0;
    int rc = ORC_OK;

    for (int i = 0; i < M && rc == ORC_OK; i++) {
        float xi = x[i], yi = y[i], zi = z[i];
        /* np.linalg.norm([x, y, z], axis=0) on float32: sqrt((x*x + y*y) + z*z) in float32, no FMA */
        float s = xi * xi;
        float t = yi * yi;
        s = s + t;
        t = zi * zi;
        s = s + t;
        float d32 = sqrtf(s);
        float th32 = theta_in ? theta_in[i] : atan2f(yi, xi);
        if (theta_out) theta_out[i] = th32;
        if (th32 < 0) th32 = th32 + (float)(2 * ORC_PI);             /* float32 add (python float is weak) */
        double right = (double)th32 - half;
        double left = (double)th32 + half;
        if (right < 0) right = right + 2 * ORC_PI;
        if (left < 0) left = left + 2 * ORC_PI;
        if (right > 2 * ORC_PI) right = right - 2 * ORC_PI;
        if (left > 2 * ORC_PI) left = left - 2 * ORC_PI;
        double d = (double)d32;

        double la_r, lb_r, la_l, lb_l;
        angle_to_line(right, &la_r, &lb_r);
        angle_to_line(left, &la_l, &lb_l);
        double den_r = sqrt(la_r * la_r + lb_r * lb_r), den_l = sqrt(la_l * la_l + lb_l * lb_l);

        int L = 0;
        for (int p = 0; p < Np; p++) {
            if (!(prange[p] < d)) continue;
            double px = particles[3 * p], py = particles[3 * p + 1], pr = particles[3 * p + 2];
            double phi = atan2(py, px);
            if (phi < 0) phi = phi + 2 * ORC_PI;
            int straddle = right > left;
            int standard = (right <= phi) && (phi <= left);
            int seldom = (right - 2 * ORC_PI <= phi) && (phi <= left) && straddle;
            int seldom2 = (right <= phi) && (phi <= left + 2 * ORC_PI) && straddle;
            int center_in = standard || seldom || seldom2;
            double dist_r = fabs((px * la_r + py * lb_r + 0.0) / den_r);
            double dist_l = fabs((px * la_l + py * lb_l + 0.0) / den_l);
            int right_hit = (dist_r < pr) && ray_on_particle_side(right, phi);
            int left_hit = (dist_l < pr) && ray_on_particle_side(left, phi);
            if (!(center_in || right_hit || left_hit)) continue;
            double ta[2];
            rc = tangent_angles(px, py, pr, phi, ta);
            if (rc != ORC_OK) break;
            if (right_hit) ta[0] = right;
            if (left_hit) ta[1] = left;
            if (L == cap) {
                cap *= 2;
                iv = (interval_t *)realloc(iv, sizeof(interval_t) * cap);
                work = (double *)realloc(work, sizeof(double) * 3 * (2 * cap + 2));
                iwork = (int *)realloc(iwork, sizeof(int) * (2 * cap + 2));
                dr = (double *)realloc(dr, sizeof(double) * (cap + 1));
                dratio = (double *)realloc(dratio, sizeof(double) * (cap + 1));
                if (!iv || !work || !iwork || !dr || !dratio) { rc = ORC_ERR_ALLOC; break; }
            }
            iv[L].a1 = ta[0]; iv[L].a2 = ta[1]; iv[L].dist = prange[p];
            L++;
        }
        if (rc != ORC_OK) break;
        /* stable insertion sort by distance (np.argsort, :416) */
        for (int a = 1; a < L; a++) {
            interval_t key = iv[a];
            int b = a - 1;
            while (b >= 0 && iv[b].dist > key.dist) { iv[b + 1] = iv[b]; b--; }
            iv[b + 1] = key;
        }
        int n_entries = 1, n_claim = 0;
        if (L > 0) n_claim = occlusion_dict(right, left, iv, L, d, beam_divergence_deg, dr, dratio, &n_entries, work, iwork);
        if (n_occluders) n_occluders[i] = n_claim;

        float *o = out + 5 * (size_t)i;
        o[0] = xi; o[1] = yi; o[2] = zi; o[3] = intensity[i];
        if (n_entries <= 1) { o[4] = 0.0f; continue; }

        memset(wave, 0, sizeof(wave));
        double i_orig = 0.9 * max_intensity;
        double CA_P0 = i_orig / beta_0;
        for (int j = 0; j < n_entries; j++) {
            int is_hard = (j == n_entries - 1);
            int start_index, end_index;
            double amp, rj = dr[j];
            if (!is_hard) {
                start_index = (int)ceil(rj * ipm);
                end_index = (int)(floor((rj + ctau) * ipm) + 1);
                amp = (CA_P0 * beta_0 * dratio[j] * xsi64(rj)) / (rj * rj);
            } else {
                float rf = d32;
                float v = rf * (float)ipm;
                start_index = (int)ceilf(v);
                float w = rf + (float)ctau;
                w = w * (float)ipm;
                end_index = (int)(floorf(w) + 1);
                float r2 = rf * rf;                                   /* float32 ** 2 */
                amp = (CA_P0 * beta_0 * dratio[j] * xsi32(rf)) / (double)r2;
            }
            for (int k = start_index; k < end_index; k++) {
                if (k >= M_EXT || k < 0) { rc = ORC_ERR_INDEX; break; }
                double sn = sin((ORC_PI * (R[k] - rj)) / ctau);
                wave[k] += amp * (sn * sn);
            }
            if (rc != ORC_OK) break;
        }
        if (rc != ORC_OK) break;
        int kmax = 0;
        for (int k = 1; k < M_EXT; k++) if (wave[k] > wave[kmax]) kmax = k;
        double i_max = wave[kmax];
        double d_max = ((double)kmax / ipm) - (ctau / 2);
        double q = 1 - d_max / 120;
        i_max += max_intensity * focal_slope * fabs(focal_offset - q * q);
        i_max = clipd(i_max, min_intensity, max_intensity);
        long new_i = (long)i_max;                                     /* int() truncation */
        if (fabs(d_max - d) < 2 * (1.0 / ipm)) {
            o[4] = 1.0f;
            sum_diff += i_orig - (double)new_i;
        } else {
            o[4] = 2.0f;
            double scale = d_max / d;
            o[0] = (float)((double)xi * scale);
            o[1] = (float)((double)yi * scale);
            o[2] = (float)((double)zi * scale);
        }
        if (new_i < 0) { rc = ORC_ERR_NEGATIVE; break; }
        o[3] = (float)clipd((double)new_i, min_intensity, max_intensity);
    }
    *intensity_diff_sum = sum_diff;
    free(prange); free(iv); free(work); free(iwork); free(dr); free(dratio);
    return rc;
}

/*
 * Whole cloud, channel fan-out of augment() (simulation.py:488-514) with OpenMP standing in for the reference's
 * process pool.  Points must already be grouped by channel: channel c owns rows ch_off[c] .. ch_off[c+1].
 *   pts      : float32[N*5] (x,y,z,intensity,channel) channel-sorted
 *   planes   : float64 particle tables concatenated; plane k (0-based file index k+1) = rows pl_off[k]..pl_off[k+1]
 *   order    : int32[64] channel -> plane index (simulation.py:70,78)
 *   sensor   : float64[64*4] focal_distance[m], focal_slope, min_intensity, max_intensity per channel
 */
int orc_snow_cloud(const float *pts, const int64_t *ch_off, int n_channels, const float *theta_in,
                   const double *planes, const int64_t *pl_off, const int32_t *order, const double *sensor,
                   double beam_divergence_deg, const double *R, float *out, double *intensity_diff_sum,
                   int32_t *n_occluders, float *theta_out, int n_threads)
{
    int rc_all = ORC_OK;
    double total = 0.0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : total)
#endif
    for (int c = 0; c < n_channels; c++) {
        int64_t b = ch_off[c], e = ch_off[c + 1];
        int M = (int)(e - b);
        if (M <= 0) continue;
        float *buf = (float *)malloc(sizeof(float) * 4 * (size_t)M);
        for (int i = 0; i < M; i++) {
            buf[i] = pts[5 * (b + i)];
            buf[M + i] = pts[5 * (b + i) + 1];
            buf[2 * M + i] = pts[5 * (b + i) + 2];
            buf[3 * M + i] = pts[5 * (b + i) + 3];
        }
        int k = order[c];
        double s = 0.0;
        int rc = orc_snow_channel(M, buf, buf + M, buf + 2 * M, buf + 3 * M, theta_in ? theta_in + b : NULL,
                                  (int)(pl_off[k + 1] - pl_off[k]), planes + 3 * pl_off[k], beam_divergence_deg,
                                  sensor[4 * c], sensor[4 * c + 1], sensor[4 * c + 2], sensor[4 * c + 3], R,
                                  out + 5 * b, &s, n_occluders ? n_occluders + b : NULL,
                                  theta_out ? theta_out + b : NULL);
        total += s;
        free(buf);
        if (rc != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = rc;
        }
    }
    *intensity_diff_sum = total;
    (void)n_threads;
    return rc_all;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* test hook: compute_occlusion_dict (simulation.py:231-295) on caller-supplied intervals (already sorted by distance).
 * intervals: L rows (a1, a2, dist).  Writes up to L+1 (r, ratio) pairs, the hard target last; returns the count. */
int orc_occlusion_dict(double right, double left, const double *intervals, int L, double current_range,
                       double beam_divergence_deg, double *out_r, double *out_ratio)
{
    interval_t *iv = (interval_t *)malloc(sizeof(interval_t) * (L > 0 ? L : 1));
    double *work = (double *)malloc(sizeof(double) * 3 * (2 * L + 2));
    int *iwork = (int *)malloc(sizeof(int) * (2 * L + 2));
    for (int j = 0; j < L; j++) { iv[j].a1 = intervals[3 * j]; iv[j].a2 = intervals[3 * j + 1]; iv[j].dist = intervals[3 * j + 2]; }
    int n = 0;
    occlusion_dict(right, left, iv, L, current_range, beam_divergence_deg, out_r, out_ratio, &n, work, iwork);
    free(iv); free(work); free(iwork);
    return n;
}
