// snowfall.cu -- batched snowfall augmentation on device-resident clouds.
//
// Pipeline per call (all on one stream):
//   k_channel_sort   stable counting sort of every cloud by channel        (tools/snowfall/simulation.py:447)
//   k_snowfall       one thread per beam: range/azimuth, candidate scan of ONE azimuth bucket of the channel's
//                    snowflake plane, exact float64 disk/wedge test, nearest-first claiming of the beam's angular
//                    sub-intervals, summed sin^2 waveform + argmax, relabel / move the point, threshold + FOV keep
//                    flag, per-cloud statistics                            (simulation.py:50-194, 231-424, 516-540)
//   k_compact        stable stream compaction of the kept rows per cloud   (simulation.py:523,540)
//   k_finalize       stats (num_attenuated, num_removed, avg_intensity_diff) (simulation.py:525-542)
//
// Numerics: everything the reference computes in float32 under NumPy 2 (range d, azimuth theta, the hard target's
// waveform window and r^2) is computed in float32 with round-to-nearest, non-fused intrinsics so it is bit-identical;
// the geometric narrow phase, the occlusion ratios and the waveform run in float64.
#include "common.cuh"

namespace {

constexpr int SNOW_TPB = 128;
constexpr int SORT_TPB = 1024;
constexpr int CHANNEL_BINS = LSS_N_CHANNELS + 1;   // + "not a valid channel"

struct DevArgs {
    // tables
    const ParticleRec *rec;
    const BroadEntry *entries;
    const int32_t *bucket_start;
    int n_buckets;
    int n_planes;
    double inv_w, w;
    // per call
    const float *pts;            // channel-sorted rows
    const float *theta;          // sorted theta or null
    const int64_t *cloud_off;    // [B+1] device
    const int32_t *order;        // [B*64] device
    const double *thresh;        // [B*3] device or null
    const SensorConst *sensor;
    const CameraConst *camera;
    const double *R;
    double half_div;             // radians(beam_divergence / 2)
    double div_rad;              // radians(beam_divergence)
    uint32_t flags;
    float *aug;                  // [N*5] un-compacted augmented rows
    uint8_t *keep;               // [N]
    int32_t *nocc;               // optional
    double *stats;               // [B*4]: num_attenuated, num_removed, avg_diff, diff_sum
    int *counters;               // [B*2]: num_attenuated(kept), num_removed
    int *status;
};

__device__ __forceinline__ void raise_status(int *status, int code) { atomicMax(status, code); }

// ---------------------------------------------------------------------------------------------------------------------
// channel sort: one CTA per cloud, tiles of SORT_TPB rows processed in order => stable
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int channel_bin(float ch)
{
    int c = (int)ch;
    return (ch >= 0.0f && ch < 64.0f && (float)c == ch) ? c : LSS_N_CHANNELS;
}

__global__ void __launch_bounds__(SORT_TPB) k_channel_sort(const float *__restrict__ pts, const float *__restrict__ theta,
                                                            const int64_t *__restrict__ cloud_off,
                                                            float *__restrict__ sorted, float *__restrict__ theta_sorted,
                                                            int32_t *__restrict__ perm)
{
    __shared__ int base[CHANNEL_BINS];
    __shared__ int warp_cnt[SORT_TPB / 32][CHANNEL_BINS];
    const int b = blockIdx.x;
    const int64_t beg = cloud_off[b];
    const int n = (int)(cloud_off[b + 1] - beg);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *src = pts + beg * 5;

    if (tid < CHANNEL_BINS) base[tid] = 0;
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += SORT_TPB) {      // warp-aggregated histogram (sorted inputs hit one bin)
        const int i = t0 + tid;
        const int bin = i < n ? channel_bin(src[(int64_t)i * 5 + 4]) : CHANNEL_BINS;
        const unsigned m = __match_any_sync(0xffffffffu, bin);
        if (bin < CHANNEL_BINS && lane == __ffs(m) - 1) atomicAdd(&base[bin], __popc(m));
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int c = 0; c < CHANNEL_BINS; c++) { int t = base[c]; base[c] = run; run += t; }
    }
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += SORT_TPB) {
        for (int k = tid; k < (SORT_TPB / 32) * CHANNEL_BINS; k += SORT_TPB) (&warp_cnt[0][0])[k] = 0;
        __syncthreads();
        const int i = t0 + tid;
        float row[5];
        int bin = CHANNEL_BINS;      // inactive
        if (i < n) {
#pragma unroll
            for (int k = 0; k < 5; k++) row[k] = src[(int64_t)i * 5 + k];
            bin = channel_bin(row[4]);
        }
        unsigned m = __match_any_sync(0xffffffffu, bin);
        int rank = __popc(m & ((1u << lane) - 1u));
        if (bin < CHANNEL_BINS && rank == 0) warp_cnt[warp][bin] = __popc(m);
        __syncthreads();
        if (tid < CHANNEL_BINS) {
            int run = base[tid];
            for (int wv = 0; wv < SORT_TPB / 32; wv++) { int t = warp_cnt[wv][tid]; warp_cnt[wv][tid] = run; run += t; }
            base[tid] = run;
        }
        __syncthreads();
        if (i < n) {
            int dst = warp_cnt[warp][bin] + rank;
            float *o = sorted + (beg + dst) * 5;
#pragma unroll
            for (int k = 0; k < 5; k++) o[k] = row[k];
            if (theta_sorted) theta_sorted[beg + dst] = theta[beg + i];
            if (perm) perm[beg + dst] = i;
        }
        __syncthreads();
    }
}

__global__ void k_identity_perm(const int64_t *__restrict__ cloud_off, int32_t *__restrict__ perm)
{
    const int b = blockIdx.y;
    const int64_t beg = cloud_off[b];
    const int n = (int)(cloud_off[b + 1] - beg);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[beg + i] = i;
}

// ---------------------------------------------------------------------------------------------------------------------
// per-beam solve
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool within(double diff, double tol)
{
    return (fabs(diff) < tol) || (fabs(diff - LSS_TWO_PI) < tol) || (fabs(diff + LSS_TWO_PI) < tol);
}

__device__ __forceinline__ double xsi64(double r)
{
    // simulation.py:553-569
    if (r <= 0.9) return 0.0;
    if (r >= 1.0) return 1.0;
    const double m = (1 - 0) / (1.0 - 0.9);
    const double b = 0 - (m * 0.9);
    return __dadd_rn(__dmul_rn(m, r), b);
}

__device__ __forceinline__ double xsi32(float r)
{
    // same with a float32 argument: NumPy 2 keeps the comparison and m*R+b in float32
    if (r <= 0.9f) return 0.0;
    if (r >= 1.0f) return 1.0;
    const double m = (1 - 0) / (1.0 - 0.9);
    const double b = 0 - (m * 0.9);
    return (double)__fadd_rn(__fmul_rn((float)m, r), (float)b);
}

__global__ void __launch_bounds__(SNOW_TPB) k_snowfall(DevArgs a)
{
    const int b = blockIdx.y;
    const int64_t beg = a.cloud_off[b];
    const int n = (int)(a.cloud_off[b + 1] - beg);
    const int i = blockIdx.x * SNOW_TPB + threadIdx.x;
    const bool active = i < n;
    const int lane = threadIdx.x & 31;

    float px = 0, py = 0, pz = 0, pint = 0, pch = 0;
    if (active) {
        const float *row = a.pts + (beg + i) * 5;
        px = row[0]; py = row[1]; pz = row[2]; pint = row[3]; pch = row[4];
    }
    // np.linalg.norm([x, y, z], axis=0) in float32: sqrt((x*x + y*y) + z*z), no FMA   (simulation.py:89)
    const float d32 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    const double d = (double)d32;

    float out_x = px, out_y = py, out_z = pz, out_i = pint, out_l = pch;
    double diff = 0.0;
    int n_claim = 0;
    const int ch = channel_bin(pch);
    const double ctau = 299792458.0 * 1e-8;

    // per-beam lists (local memory): hits (a1, a2, range) -> after claiming: pulses (amplitude, range, window)
    double ha1[LSS_MAX_OCC + 1], ha2[LSS_MAX_OCC], hr[LSS_MAX_OCC + 1];
    int ks[LSS_MAX_OCC + 1], ke[LSS_MAX_OCC + 1];
    int n_pulses = 0;            // > 0: this beam has a waveform to solve (claiming particles + hard target)

    if (active && ch < LSS_N_CHANNELS) {
        out_l = 0.0f;
        // ---- beam limits (simulation.py:91-101) -------------------------------------------------------------------
        float th32 = a.theta ? a.theta[beg + i] : (float)atan2((double)py, (double)px);
        if (th32 < 0.0f) th32 = __fadd_rn(th32, 6.2831855f);
        const double thd = (double)th32;
        double right = thd - a.half_div, left = thd + a.half_div;
        if (right < 0) right += LSS_TWO_PI;
        if (left < 0) left += LSS_TWO_PI;
        if (right > LSS_TWO_PI) right -= LSS_TWO_PI;
        if (left > LSS_TWO_PI) left -= LSS_TWO_PI;
        const bool straddle = right > left;

        // ---- candidate scan: one azimuth bucket of this channel's plane -------------------------------------------
        int L = 0;
        bool overflow = false;
        const int plane = a.order[b * LSS_N_CHANNELS + ch];
        if (plane >= 0 && plane < a.n_planes && thd == thd) {
            double thm = thd >= LSS_TWO_PI ? thd - LSS_TWO_PI : thd;
            int bk = (int)(thm * a.inv_w);
            bk = bk < 0 ? 0 : (bk >= a.n_buckets ? a.n_buckets - 1 : bk);
            const float th_rel = (float)(thm - (bk + 0.5) * a.w);
            const int32_t *bs = a.bucket_start + (int64_t)plane * (a.n_buckets + 1) + bk;
            const int e0 = bs[0], e1 = bs[1];
            for (int e = e0; e < e1; e++) {
                const BroadEntry en = __ldg(&a.entries[e]);
                if (!(en.x < d32)) break;                       // sorted by range: nothing nearer follows
                if (!(fabsf(en.y - th_rel) <= en.z)) continue;   // float32 broad phase (conservative)
                const ParticleRec *rp = a.rec + __float_as_int(en.w);
                const double rho = rp->rho;
                if (!(rho < d)) continue;                        // simulation.py:345 (strict, float64)
                const double phi = rp->phi, alpha = rp->alpha;
                // simulation.py:359-365 centre inside the beam
                bool inside = (right <= phi) && (phi <= left);
                if (straddle) inside = inside || ((right - LSS_TWO_PI <= phi) && (phi <= left)) ||
                                       ((right <= phi) && (phi <= left + LSS_TWO_PI));
                // simulation.py:371-385: disk crosses a limit ray  <=>  |phi - limit| < asin(r/rho)  (mod 2 pi)
                const bool right_hit = within(right - phi, alpha);
                const bool left_hit = within(left - phi, alpha);
                if (!(inside || right_hit || left_hit)) continue;
                if (L == LSS_MAX_OCC) { overflow = true; break; }
                const double a1 = right_hit ? right : rp->t_right;   // geometry.py:26-27
                const double a2 = left_hit ? left : rp->t_left;
                int j = L - 1;                                   // insertion by range (np.argsort, :416)
                while (j >= 0 && hr[j] > rho) { ha1[j + 1] = ha1[j]; ha2[j + 1] = ha2[j]; hr[j + 1] = hr[j]; j--; }
                ha1[j + 1] = a1; ha2[j + 1] = a2; hr[j + 1] = rho;
                L++;
            }
        }
        if (overflow) raise_status(a.status, LSS_ERR_OCCLUDER_OVERFLOW);

        if (L > 0 && !overflow) {
            // ---- compute_occlusion_dict (simulation.py:231-295) ---------------------------------------------------
            // The reference splits the beam into elementary sub-intervals between all sorted end points and lets the
            // particles claim, nearest first, every still-unclaimed piece inside their own interval.  Equivalent
            // formulation used here: keep the union of the intervals claimed so far as a list of disjoint, non-touching
            // intervals; a particle claims |[a1,a2]| - |[a1,a2] n union| and is dropped iff [a1,a2] is contained in
            // one union interval (or a1 >= a2, the reference's empty range(i1, i2)).  The hard target gets what is left
            // between the smallest and the largest end point -- including the ~2 pi gap of the seam quirk.
            double rb = right;
            if (straddle) {
                rb = right - LSS_TWO_PI;
                for (int j = 0; j < L; j++) if (ha1[j] > ha2[j]) ha1[j] -= LSS_TWO_PI;
            }
            double ulo[LSS_MAX_OCC], uhi[LSS_MAX_OCC];
            int nu = 0;
            double ep_min = fmin(rb, left), ep_max = fmax(rb, left), claimed_total = 0.0;
            int P = 0;          // pulses: claiming particles in range order, then the hard target
            for (int j = 0; j < L; j++) {
                const double lo = ha1[j], hi = ha2[j];
                ep_min = fmin(ep_min, fmin(lo, hi));
                ep_max = fmax(ep_max, fmax(lo, hi));
                if (!(lo < hi)) continue;
                bool contained = false;
                double cov = 0.0;
                for (int u = 0; u < nu; u++) {
                    contained |= (ulo[u] <= lo) && (hi <= uhi[u]);
                    const double ov = fmin(hi, uhi[u]) - fmax(lo, ulo[u]);
                    if (ov > 0.0) cov += ov;
                }
                if (contained) continue;
                const double claimed = (hi - lo) - cov;
                claimed_total += claimed;
                double nlo = lo, nhi = hi;      // merge [lo, hi] into the union (absorb overlapping / touching pieces)
                int w = 0;
                for (int u = 0; u < nu; u++) {
                    if (ulo[u] <= nhi && uhi[u] >= nlo && ulo[u] <= hi && uhi[u] >= lo) {
                        nlo = fmin(nlo, ulo[u]);
                        nhi = fmax(nhi, uhi[u]);
                    } else {
                        ulo[w] = ulo[u]; uhi[w] = uhi[u]; w++;
                    }
                }
                ulo[w] = nlo; uhi[w] = nhi;
                nu = w + 1;
                double ratio = claimed / a.div_rad;
                ratio = ratio < 0 ? 0 : (ratio > 1 ? 1 : ratio);
                hr[P] = hr[j];          // P <= j: safe in place
                ha1[P] = ratio;
                P++;
            }
            n_claim = P;
            double ratio_hard = ((ep_max - ep_min) - claimed_total) / a.div_rad;
            ratio_hard = ratio_hard < 0 ? 0 : (ratio_hard > 1 ? 1 : ratio_hard);

            if (P > 0) {
                // ---- pulses of the waveform (simulation.py:137-149) -------------------------------------------------
                const double beta_0 = 1 * 1e-06 / LSS_PI;
                const double i_orig = 0.9 * a.sensor->max_intensity[ch];
                const double A = (i_orig / beta_0) * beta_0;        // CA_P0 * beta_0 (quirk: every pulse uses it)
                double *amp = ha1, *rj = hr;                         // reuse the hit arrays
                bool bad = false;
                for (int j = 0; j < P; j++) {
                    const double r = rj[j];
                    ks[j] = (int)ceil(r * 10);
                    ke[j] = (int)(floor((r + ctau) * 10) + 1);
                    amp[j] = (A * amp[j] * xsi64(r)) / (r * r);
                    bad |= (ke[j] > LSS_M_EXT) || (ks[j] < 0);
                }
                {   // hard target: r_j is float32 => float32 index arithmetic and r^2
                    ks[P] = (int)ceilf(__fmul_rn(d32, 10.0f));
                    ke[P] = (int)(floorf(__fmul_rn(__fadd_rn(d32, (float)ctau), 10.0f)) + 1.0f);
                    rj[P] = d;
                    amp[P] = (A * ratio_hard * xsi32(d32)) / (double)__fmul_rn(d32, d32);
                    bad |= (ke[P] > LSS_M_EXT) || (ks[P] < 0);
                }
                if (bad) raise_status(a.status, LSS_ERR_RANGE_INDEX);
                else n_pulses = P + 1;
            }
        }
    }

    // ---- argmax of the summed waveform (simulation.py:148-153), warp-cooperative ------------------------------------
    // Only samples inside some pulse window are non-zero.  Pulses whose windows overlap form a group whose samples are
    // summed in full (in dict order, like the reference's i[k] +=); an isolated pulse A sin^2(pi (R - r)/(c tau)) is
    // unimodal and symmetric about r + c tau / 2, so its maximum over the grid is at one of the three samples around
    // the sample nearest to the peak.  The owner lane publishes its pulses and candidate segments in shared memory,
    // the 32 lanes evaluate the candidate samples in parallel and reduce to the first maximum (np.argmax).
    __shared__ double s_amp[SNOW_TPB / 32][LSS_MAX_OCC + 1];
    __shared__ double s_r[SNOW_TPB / 32][LSS_MAX_OCC + 1];
    __shared__ int s_win[SNOW_TPB / 32][LSS_MAX_OCC + 1];
    __shared__ int4 s_seg[SNOW_TPB / 32][LSS_MAX_OCC + 2];       // first sample, first pulse, last pulse, prefix
    double best = 0.0;
    int kbest = 0;
    {
        const int wid = threadIdx.x >> 5;
        unsigned todo = __ballot_sync(0xffffffffu, n_pulses > 0);
        while (todo) {
            const int owner = __ffs(todo) - 1;
            todo &= todo - 1;
            int nseg = 0, T = 0;
            if (lane == owner) {
                const double inv_step = (double)(LSS_M_EXT - 1) / (120 + ctau);
                for (int j = 0; j < n_pulses; j++) {
                    s_amp[wid][j] = ha1[j];
                    s_r[wid][j] = hr[j];
                    s_win[wid][j] = ks[j] | (ke[j] << 16);
                }
                int j = 0;
                while (j < n_pulses) {
                    int g1 = j, k_lo = ks[j], k_hi = ke[j];
                    while (g1 + 1 < n_pulses && ks[g1 + 1] < k_hi) {
                        g1++;
                        k_lo = min(k_lo, ks[g1]);
                        k_hi = max(k_hi, ke[g1]);
                    }
                    if (g1 == j) {
                        const int k0 = (int)rint((hr[j] + ctau / 2) * inv_step);
                        k_lo = max(k_lo, k0 - 1);
                        k_hi = min(k_hi, k0 + 2);
                    }
                    if (k_hi > k_lo) {
                        s_seg[wid][nseg] = make_int4(k_lo, j, g1, T);
                        T += k_hi - k_lo;
                        nseg++;
                    }
                    j = g1 + 1;
                }
            }
            __syncwarp();
            nseg = __shfl_sync(0xffffffffu, nseg, owner);
            T = __shfl_sync(0xffffffffu, T, owner);
            double bv = 0.0;
            int bk = 0;
            for (int c = lane; c < T; c += 32) {
                int sgi = 0;
                while (sgi + 1 < nseg && s_seg[wid][sgi + 1].w <= c) sgi++;
                const int4 sg = s_seg[wid][sgi];
                const int k = sg.x + (c - sg.w);
                const double Rk = __ldg(&a.R[k]);
                double v = 0.0;
                for (int q = sg.y; q <= sg.z; q++) {
                    const int wn = s_win[wid][q];
                    if (k >= (wn & 0xffff) && k < (wn >> 16)) {
                        const double sn = sin((LSS_PI * (Rk - s_r[wid][q])) / ctau);
                        v += s_amp[wid][q] * (sn * sn);
                    }
                }
                if (v > bv) { bv = v; bk = k; }               // ascending k per lane: first maximum
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, sh);
                const int ok = __shfl_xor_sync(0xffffffffu, bk, sh);
                if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
            }
            if (lane == owner) { best = bv; kbest = bv > 0.0 ? bk : 0; }
            __syncwarp();
        }
    }

    if (n_pulses > 0) {
        // ---- new range / intensity / label (simulation.py:151-188) -------------------------------------------------
        const double max_i = a.sensor->max_intensity[ch];
        const double min_i = a.sensor->min_intensity[ch];
        const double i_orig = 0.9 * max_i;
        const double d_max = ((double)kbest / 10) - (ctau / 2);
        const double q1 = 1 - d_max / 120;
        double i_max = best + max_i * a.sensor->focal_slope[ch] * fabs(a.sensor->focal_offset[ch] - q1 * q1);
        i_max = i_max < min_i ? min_i : (i_max > max_i ? max_i : i_max);
        const long long new_i = (long long)i_max;       // int(): truncation
        if (fabs(d_max - d) < 2 * (1.0 / 10)) {
            out_l = 1.0f;
            diff = i_orig - (double)new_i;
        } else {
            out_l = 2.0f;
            const double sc = d_max / d;
            out_x = (float)((double)px * sc);
            out_y = (float)((double)py * sc);
            out_z = (float)((double)pz * sc);
        }
        if (new_i < 0) raise_status(a.status, LSS_ERR_NEGATIVE_INTENSITY);
        double ci = (double)new_i;
        ci = ci < min_i ? min_i : (ci > max_i ? max_i : ci);
        out_i = (float)ci;
    }

    // ---- cloud-level post: round, threshold, FOV (simulation.py:516-540) ------------------------------------------
    bool keep = active;
    bool keep_thr = active;      // kept by the threshold filter: num_attenuated is counted BEFORE the FOV filter (:525)
    bool removed = false;
    if (active) {
        out_i = rintf(out_i);
        if (a.flags & LSS_FLAG_THRESHOLD_FILTER) {
            const double *p = a.thresh + 3 * b;
            const double d2 = (double)__fmul_rn(d32, d32);
            const double thr = __dadd_rn(__dadd_rn(__dmul_rn(p[0], d2), __dmul_rn(p[1], d)), p[2]);
            keep = (out_l == 2.0f) || ((double)out_i > thr);
        }
        keep_thr = keep;
        if (keep && (a.flags & LSS_FLAG_CAMERA_FOV)) {
            const float *M = a.camera->M, *P2 = a.camera->P2;
            float rx = fmaf(out_z, M[6], fmaf(out_y, M[3], out_x * M[0])) + M[9];
            float ry = fmaf(out_z, M[7], fmaf(out_y, M[4], out_x * M[1])) + M[10];
            float rz = fmaf(out_z, M[8], fmaf(out_y, M[5], out_x * M[2])) + M[11];
            float u = fmaf(rz, P2[2], fmaf(ry, P2[1], rx * P2[0])) + P2[3];
            float v = fmaf(rz, P2[6], fmaf(ry, P2[5], rx * P2[4])) + P2[7];
            float wd = fmaf(rz, P2[10], fmaf(ry, P2[9], rx * P2[8])) + P2[11];
            u = u / rz;
            v = v / rz;
            const float depth = wd - P2[11];
            keep = (u >= 0.0f) && (u < (float)a.camera->img_w) && (v >= 0.0f) && (v < (float)a.camera->img_h) &&
                   (depth >= 0.0f);
        }
        removed = !keep;
        float *o = a.aug + (beg + i) * 5;
        o[0] = out_x; o[1] = out_y; o[2] = out_z; o[3] = out_i; o[4] = out_l;
        a.keep[beg + i] = keep ? 1 : 0;
        if (a.nocc) a.nocc[beg + i] = n_claim;
    }
    // per-cloud statistics: warp-aggregated
    const unsigned m_att = __ballot_sync(0xffffffffu, keep_thr && out_l == 1.0f && ch < LSS_N_CHANNELS);
    const unsigned m_rem = __ballot_sync(0xffffffffu, removed);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, s);
    if (lane == 0) {
        if (m_att) atomicAdd(&a.counters[2 * b], __popc(m_att));
        if (m_rem) atomicAdd(&a.counters[2 * b + 1], __popc(m_rem));
        if (diff != 0.0) atomicAdd(&a.stats[4 * b + 3], diff);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stable compaction, one CTA per cloud
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_TPB) k_compact(const float *__restrict__ aug, const uint8_t *__restrict__ keep,
                                                       const int64_t *__restrict__ cloud_off, float *__restrict__ out,
                                                       int32_t *__restrict__ counts)
{
    __shared__ int warp_tot[SORT_TPB / 32];
    __shared__ int run_s;
    const int b = blockIdx.x;
    const int64_t beg = cloud_off[b];
    const int n = (int)(cloud_off[b + 1] - beg);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) run_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += SORT_TPB) {
        const int i = t0 + tid;
        const bool k = (i < n) && keep[beg + i];
        const unsigned m = __ballot_sync(0xffffffffu, k);
        const int rank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = run_s;
        for (int wv = 0; wv < warp; wv++) off += warp_tot[wv];
        if (k) {
            const float *s = aug + (beg + i) * 5;
            float *o = out + (beg + off + rank) * 5;
#pragma unroll
            for (int q = 0; q < 5; q++) o[q] = s[q];
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int wv = 0; wv < SORT_TPB / 32; wv++) t += warp_tot[wv];
            run_s += t;
        }
        __syncthreads();
    }
    if (tid == 0) counts[b] = run_s;
}

__global__ void k_finalize(double *stats, const int *counters, int n_clouds)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_clouds) return;
    const int n_att = counters[2 * b], n_rem = counters[2 * b + 1];
    const double sum = stats[4 * b + 3];
    stats[4 * b + 0] = (double)n_att;
    stats[4 * b + 1] = (double)n_rem;
    stats[4 * b + 2] = n_att > 0 ? (double)(long long)(sum / (double)n_att) : 0.0;   // int(sum / n), :527-530
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
    int64_t sorted, theta, keep, perm, aug, cloud_off, order, thresh, counters, prepass, prepass_bytes, total;
};

WsLayout ws_layout(int64_t n_total, int n_clouds)
{
    WsLayout w;
    int64_t o = 0;
    w.sorted = o;    o = align_up(o + n_total * 5 * 4, 256);
    w.theta = o;     o = align_up(o + n_total * 4, 256);
    w.keep = o;      o = align_up(o + n_total, 256);
    w.perm = o;      o = align_up(o + n_total * 4, 256);
    w.aug = o;       o = align_up(o + n_total * 5 * 4, 256);
    w.cloud_off = o; o = align_up(o + (int64_t)(n_clouds + 1) * 8, 256);
    w.order = o;     o = align_up(o + (int64_t)n_clouds * LSS_N_CHANNELS * 4, 256);
    w.thresh = o;    o = align_up(o + (int64_t)n_clouds * 3 * 8, 256);
    w.counters = o;  o = align_up(o + (int64_t)n_clouds * 2 * 4, 256);
    w.prepass_bytes = lss_prepass_ws_bytes(n_total, n_clouds);
    w.prepass = o;   o = align_up(o + w.prepass_bytes, 256);
    w.total = o;
    return w;
}

}  // namespace

int64_t lss_snowfall_ws_bytes(int64_t n_total, int n_clouds)
{
    if (n_total < 0 || n_clouds < 0) return -1;
    return ws_layout(n_total, n_clouds).total;
}

lss_status lss_snowfall_run(lss_engine *e, const SnowfallArgs &s, cudaStream_t stream)
{
    const int B = s.n_clouds;
    const int64_t N = s.h_cloud_offsets[B] - s.h_cloud_offsets[0];
    if (s.h_cloud_offsets[0] != 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud_offsets[0] must be 0");
    int64_t max_n = 0;
    for (int b = 0; b < B; b++) {
        int64_t n = s.h_cloud_offsets[b + 1] - s.h_cloud_offsets[b];
        if (n < 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud_offsets must be non-decreasing");
        if (n >= (1LL << 31)) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud too large");
        max_n = n > max_n ? n : max_n;
    }
    for (int k = 0; k < B * LSS_N_CHANNELS; k++)
        if (s.h_order[k] < 0 || s.h_order[k] >= s.ts->n_planes)
            return lss_fail(e, LSS_ERR_NO_TABLE, "order[] names a plane that is not in the table set");
    const WsLayout w = ws_layout(N, B);
    if (s.workspace_bytes < w.total || !s.d_workspace) return lss_fail(e, LSS_ERR_WORKSPACE, "workspace too small");
    if ((s.flags & LSS_FLAG_THRESHOLD_FILTER) && !s.h_thresh_poly && !(s.flags & LSS_FLAG_DEVICE_PREPASS))
        return lss_fail(e, LSS_ERR_INVALID_ARG, "threshold filter needs h_thresh_poly or LSS_FLAG_DEVICE_PREPASS");
    if ((s.flags & LSS_FLAG_CAMERA_FOV) && !e->has_camera)
        return lss_fail(e, LSS_ERR_NO_SENSOR, "camera calibration not set");
    const double div_rad = s.beam_divergence_deg * (LSS_PI / 180.0);
    if (!(div_rad > 0) || div_rad > s.ts->max_div_rad * (1 + 1e-12))
        return lss_fail(e, LSS_ERR_INVALID_ARG, "beam_divergence exceeds the value the table set was built for");

    char *ws = (char *)s.d_workspace;
    float *d_sorted = (float *)(ws + w.sorted);
    float *d_theta_sorted = (float *)(ws + w.theta);
    uint8_t *d_keep = (uint8_t *)(ws + w.keep);
    int32_t *d_perm = s.d_out_perm ? s.d_out_perm : (int32_t *)(ws + w.perm);
    float *d_aug = s.d_out_full ? s.d_out_full : (float *)(ws + w.aug);
    int64_t *d_off = (int64_t *)(ws + w.cloud_off);
    int32_t *d_order = (int32_t *)(ws + w.order);
    double *d_thresh = (double *)(ws + w.thresh);
    int *d_counters = (int *)(ws + w.counters);

    LSS_CUDA_CHECK(e, cudaMemcpyAsync(d_off, s.h_cloud_offsets, sizeof(int64_t) * (B + 1), cudaMemcpyHostToDevice, stream));
    LSS_CUDA_CHECK(e, cudaMemcpyAsync(d_order, s.h_order, sizeof(int32_t) * B * LSS_N_CHANNELS, cudaMemcpyHostToDevice, stream));
    if (s.h_thresh_poly)
        LSS_CUDA_CHECK(e, cudaMemcpyAsync(d_thresh, s.h_thresh_poly, sizeof(double) * 3 * B, cudaMemcpyHostToDevice, stream));
    LSS_CUDA_CHECK(e, cudaMemsetAsync(d_counters, 0, sizeof(int) * 2 * B, stream));
    LSS_CUDA_CHECK(e, cudaMemsetAsync(s.d_out_stats, 0, sizeof(double) * 4 * B, stream));
    if (N == 0 || B == 0) {
        if (B) LSS_CUDA_CHECK(e, cudaMemsetAsync(s.d_out_counts, 0, sizeof(int32_t) * B, stream));
        return LSS_OK;
    }

    const float *d_pts_sorted = s.d_points;
    const float *d_theta = s.d_theta;
    if (!(s.flags & LSS_FLAG_ASSUME_SORTED)) {
        {
            KernelTimer kt(e, LSS_K_SORT, stream);
            k_channel_sort<<<B, SORT_TPB, 0, stream>>>(s.d_points, s.d_theta, d_off, d_sorted,
                                                       s.d_theta ? d_theta_sorted : nullptr, d_perm);
        }
        d_pts_sorted = d_sorted;
        d_theta = s.d_theta ? d_theta_sorted : nullptr;
    } else if (s.d_out_perm) {
        dim3 g((unsigned)((max_n + 255) / 256), B);
        KernelTimer kt(e, LSS_K_SORT, stream);
        k_identity_perm<<<g, 256, 0, stream>>>(d_off, d_perm);
    }

    if ((s.flags & LSS_FLAG_THRESHOLD_FILTER) && (s.flags & LSS_FLAG_DEVICE_PREPASS) && !s.h_thresh_poly) {
        // plane + laser parameters + threshold polynomial from the channel-sorted cloud (simulation.py:449-467)
        lss_status ps = lss_prepass_run(e, d_pts_sorted, d_off, nullptr, s.h_cloud_offsets, B, 0.5, s.noise_floor, 0, 0, 1, nullptr,
                                        d_thresh, nullptr, ws + w.prepass, w.prepass_bytes, nullptr, stream);
        if (ps != LSS_OK) return ps;
    }

    DevArgs a;
    a.rec = s.ts->d_rec;
    a.entries = s.ts->d_entries;
    a.bucket_start = s.ts->d_bucket_start;
    a.n_buckets = s.ts->n_buckets;
    a.n_planes = s.ts->n_planes;
    a.w = LSS_TWO_PI / s.ts->n_buckets;
    a.inv_w = s.ts->n_buckets / LSS_TWO_PI;
    a.pts = d_pts_sorted;
    a.theta = d_theta;
    a.cloud_off = d_off;
    a.order = d_order;
    a.thresh = d_thresh;
    a.sensor = e->d_sensor;
    a.camera = e->d_camera;
    a.R = e->d_R;
    a.half_div = (s.beam_divergence_deg / 2) * (LSS_PI / 180.0);
    a.div_rad = div_rad;
    a.flags = s.flags;
    a.aug = d_aug;
    a.keep = d_keep;
    a.nocc = s.d_out_nocc;
    a.stats = s.d_out_stats;
    a.counters = d_counters;
    a.status = e->d_status;
    dim3 grid((unsigned)((max_n + SNOW_TPB - 1) / SNOW_TPB), B);
    {
        KernelTimer kt(e, LSS_K_SNOWFALL, stream);
        k_snowfall<<<grid, SNOW_TPB, 0, stream>>>(a);
    }
    {
        KernelTimer kt(e, LSS_K_COMPACT, stream);
        k_compact<<<B, SORT_TPB, 0, stream>>>(d_aug, d_keep, d_off, s.d_out_points, s.d_out_counts);
    }
    {
        KernelTimer kt(e, LSS_K_FINALIZE, stream);
        k_finalize<<<(B + 127) / 128, 128, 0, stream>>>(s.d_out_stats, d_counters, B);
    }
    LSS_CUDA_CHECK(e, cudaGetLastError());
    return LSS_OK;
}
