// snowfall.cu -- batched snowfall augmentation on device-resident clouds.
//
// Pipeline per call (on the caller's stream; the pre-pass is forked onto an engine side stream and joined before k_keep):
//   [pre-pass]       ground plane + noise-threshold polynomial (prepass.cu), concurrent with the beam kernels
//                                                                                 (tools/snowfall/simulation.py:449-467)
//   k_scan           (solve.cu) every beam, rows in INPUT order: range / azimuth, walk of ONE azimuth bucket of the channel's
//                    snowflake plane; un-occluded beams are finished, the others pushed to the solve list with their hits
//   k_list_sort      counting sort of the solve list by work class (a solve warp runs as long as its slowest lane)
//   k_solve          (solve.cu) the listed beams: nearest-first claiming of the beam's angular sub-intervals, summed
//                    sin^2 waveform + argmax, relabel / move the point, label-1 statistics  (simulation.py:50-194, 231-424)
//   k_overflow       beams with more than 63 occluders (rare): one beam per thread, per-thread lists of up to 128 hits
//   k_keep           threshold (original range) + FOV keep flag, per-tile channel histogram of the kept rows,
//                    num_attenuated / num_removed                                 (simulation.py:516-540)
//   k_tile_scan      per cloud: exclusive scan of the tile histograms -> destination of every (tile, channel) run;
//                    kept-row count and stats (num_attenuated, num_removed, avg_intensity_diff)  (simulation.py:525-542)
//   k_scatter        stable scatter of the kept rows to "sorted by channel, compacted" order -- the reference's
//                    pc[pc[:, 4].argsort()] (simulation.py:447) and its boolean-mask filters (:523,540) in ONE pass
//
// The reference sorts first and filters last; both are pure row permutations/selections that commute with the per-beam
// solve, so they are applied once, at the end (82 B/point of traffic instead of 122).
//
// Numerics: everything the reference computes in float32 under NumPy 2 (range d, azimuth theta, the hard target's
// waveform window and r^2) is computed in float32 with round-to-nearest, non-fused intrinsics so it is bit-identical;
// the geometric narrow phase, the occlusion ratios and the waveform run in float64.
#include "beam.cuh"
#include <cstdlib>

#define LSS_CHECK_STATUS(call) do { const lss_status _st = (call); if (_st != LSS_OK) return _st; } while (0)

namespace {

// Cold or bulky pieces are kept out of line and data-dependent loops are not unrolled: the beam kernel is
// instruction-fetch sensitive (profiles/: `no_inst` stalls grow with its SASS size), every instruction that is not
// on the common path costs fetch bandwidth for all resident warps.
// one waveform sample: sum over the pulses q0..q1 whose window contains k, in dict order (simulation.py:148-149)
__device__ __noinline__ double waveform_sample(int k, double Rk, int q0, int q1, const double *amp, const double *r,
                                               const int *ks, const int *ke)
{
    const double inv_ctau = 1.0 / (299792458.0 * 1e-8);
    double v = 0.0;
#pragma unroll 1
    for (int q = q0; q <= q1; q++)
        if (k >= ks[q] && k < ke[q]) {
            const double sn = sinpi((Rk - r[q]) * inv_ctau);
            v += amp[q] * (sn * sn);
        }
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// overflow kernel: the beams the solve kernel's shared-memory arena cannot take (more than 63 occluders).  This is round
// 1's per-beam kernel reduced to its list mode: it walks the bucket again, keeps per-thread lists (local memory) and
// evaluates whole windows with a sinpi per sample -- slow, general, and only ever run on a handful of beams.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SNOW_TPB, 1) k_overflow(DevArgs a)
{
    constexpr int CAP = SLOW_CAP;
    __shared__ double s_amp[SNOW_WARPS][POOL];
    __shared__ double s_r[SNOW_WARPS][POOL];
    __shared__ int s_win[SNOW_WARPS][POOL];
    __shared__ unsigned s_cand[SNOW_WARPS][CCAP];                   // sample | first pulse << 11 | last pulse << 18 | lane << 25
    __shared__ double s_best[SNOW_WARPS][32];
    __shared__ int s_kbest[SNOW_WARPS][32];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // one listed beam per thread: the lanes of a warp may belong to different clouds
    const int slot = blockIdx.x * SNOW_TPB + threadIdx.x;
    const int cnt = min(*a.count_in, a.cap_in);
    if (blockIdx.x * SNOW_TPB >= cnt) return;
    const bool active = slot < cnt;
    const unsigned long long it = active ? a.list_in[slot] : 0ull;
    const int b = (int)((it >> 32) & 0xffffu);
    const int i = (int)(it & 0xffffffffu);
    const int64_t beg = a.cloud_off[b];
    float px = 0, py = 0, pz = 0, pint = 0, pch = 0;
    if (active) {
        const float *row = a.pts + (beg + i) * 5;
        px = row[0]; py = row[1]; pz = row[2]; pint = row[3]; pch = row[4];
    }
    // np.linalg.norm([x, y, z], axis=0) in float32: sqrt((x*x + y*y) + z*z), no FMA   (simulation.py:89)
    const float d32 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    const double d = (double)d32;

    float out_x = px, out_y = py, out_z = pz, out_i = pint, out_l = pch;
    long long att_new_i = -1;    // >= 0: this beam was attenuated (label 1) to this integer intensity
    int n_claim = 0;
    const int ch = channel_bin(pch);
    const double ctau = 299792458.0 * 1e-8;

    // per-beam lists (local memory): hits (a1, a2, range) -> after claiming: pulses (amplitude, range, window)
    double ha1[CAP + 1], ha2[CAP], hr[CAP + 1];
    int ks[CAP + 1], ke[CAP + 1];
    bool deferred = false;       // only with a further overflow list (a.list_out): not used by the engine's launcher
    int n_pulses = 0;            // > 0: this beam has a waveform to solve (claiming particles + hard target)

    if (active && ch < LSS_N_CHANNELS) {
        out_l = 0.0f;
        // ---- beam limits (simulation.py:91-101) -------------------------------------------------------------------
        float th32 = a.theta ? a.theta[beg + i] : azimuth32(py, px);
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
#pragma unroll 1
            for (int e = e0; e < e1; e++) {
                const EntryView en = lss_decode(__ldg(&a.entries[e]), a.zbase);
                if (!(en.x < d32)) break;                       // sorted by range: nothing nearer follows
                if (!(fabsf(en.y - th_rel) <= en.z)) continue;   // float32 broad phase (conservative)
                const long long pi = a.plane_off[plane] + en.idx;
                const ParticleRec *rp = a.rec + pi;
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
                if (L == CAP) { overflow = true; break; }
                const double a1 = right_hit ? right : a.tan[pi].t_right;   // geometry.py:26-27
                const double a2 = left_hit ? left : a.tan[pi].t_left;
                int j = L - 1;                                   // insertion by range (np.argsort, :416)
#pragma unroll 1
                while (j >= 0 && hr[j] > rho) { ha1[j + 1] = ha1[j]; ha2[j + 1] = ha2[j]; hr[j + 1] = hr[j]; j--; }
                ha1[j + 1] = a1; ha2[j + 1] = a2; hr[j + 1] = rho;
                L++;
            }
        }
        if (overflow) {
            if (a.list_out == nullptr) {
                raise_status(a.status, LSS_ERR_OCCLUDER_OVERFLOW);
            } else {
                const int slot = atomicAdd(a.count_out, 1);
                if (slot < a.cap_out) a.list_out[slot] = ((unsigned long long)b << 32) | (unsigned)i;
                else raise_status(a.status, LSS_ERR_OCCLUDER_OVERFLOW);
                deferred = slot < a.cap_out;
            }
        }
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
#pragma unroll 1
                for (int j = 0; j < L; j++) if (ha1[j] > ha2[j]) ha1[j] -= LSS_TWO_PI;
            }
            double ulo[CAP], uhi[CAP];
            int nu = 0;
            double ep_min = fmin(rb, left), ep_max = fmax(rb, left), claimed_total = 0.0;
            int P = 0;          // pulses: claiming particles in range order, then the hard target
#pragma unroll 1
            for (int j = 0; j < L; j++) {
                const double lo = ha1[j], hi = ha2[j];
                ep_min = fmin(ep_min, fmin(lo, hi));
                ep_max = fmax(ep_max, fmax(lo, hi));
                if (!(lo < hi)) continue;
                bool contained = false;
                double cov = 0.0;
#pragma unroll 1
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
#pragma unroll 1
                for (int u = 0; u < nu; u++) {
                    if (ulo[u] <= hi && uhi[u] >= lo) {
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
#pragma unroll 1
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
    // the sample nearest to the peak.  Every lane publishes its pulses and one descriptor per candidate sample in the
    // warp's pool; the candidates of all beams of the warp are then evaluated 32 at a time and reduced per beam with a
    // segmented scan (first maximum wins, np.argmax).
    const double inv_step = (double)(LSS_M_EXT - 1) / (120 + ctau);
    const double inv_ctau = 1.0 / ctau;
    // candidate segments of this beam: calls f(k_lo, k_hi, first pulse, last pulse) in ascending sample order
    auto for_each_segment = [&](auto &&f) {
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
            if (k_hi > k_lo) f(k_lo, k_hi, j, g1);
            j = g1 + 1;
        }
    };
    int T_all = 0;
    if (n_pulses > 0) for_each_segment([&](int k_lo, int k_hi, int, int) { T_all += k_hi - k_lo; });
    double best = 0.0;
    int kbest = 0;
    bool coop = n_pulses > 0;
    if (T_all > CCAP || n_pulses > POOL) {   // pathological beam (dozens of overlapping pulses): solve it in this thread
        coop = false;
        for_each_segment([&](int k_lo, int k_hi, int j0, int j1) {
#pragma unroll 1
            for (int k = k_lo; k < k_hi; k++) {
                const double v = waveform_sample(k, __ldg(&a.R[k]), j0, j1, ha1, hr, ks, ke);
                if (v > best) { best = v; kbest = k; }
            }
        });
    }
    s_best[wid][lane] = 0.0;
    s_kbest[wid][lane] = 0;
    unsigned remaining = __ballot_sync(0xffffffffu, coop);
    while (remaining) {
        const bool rem = (remaining >> lane) & 1u;
        const int mine = rem ? n_pulses : 0, myT = rem ? T_all : 0;
        int incl = mine, cincl = myT;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, s);
            const int u = __shfl_up_sync(0xffffffffu, cincl, s);
            if (lane >= s) { incl += t; cincl += u; }
        }
        // a prefix of the remaining lanes (each has <= 49 pulses and <= CCAP candidates)
        const bool in_batch = rem && incl <= POOL && cincl <= CCAP;
        const unsigned batch = __ballot_sync(0xffffffffu, in_batch);
        remaining &= ~batch;
        const int total = __shfl_sync(0xffffffffu, cincl, 31 - __clz(batch));
        if (in_batch) {
            const int poff = incl - mine;
            int coff = cincl - myT;
#pragma unroll 1
            for (int j = 0; j < mine; j++) {
                s_amp[wid][poff + j] = ha1[j];
                s_r[wid][poff + j] = hr[j];
                s_win[wid][poff + j] = ks[j] | (ke[j] << 16);
            }
            for_each_segment([&](int k_lo, int k_hi, int j0, int j1) {
                const unsigned hi_bits = ((unsigned)(poff + j0) << 11) | ((unsigned)(poff + j1) << 18) | ((unsigned)lane << 25);
#pragma unroll 4
                for (int k = k_lo; k < k_hi; k++) s_cand[wid][coff++] = (unsigned)k | hi_bits;
            });
        }
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
            const int c = base + lane;
            int owner = -1, k = 0;
            double v = 0.0;
            if (c < total) {
                const unsigned desc = s_cand[wid][c];
                k = desc & 2047u;
                owner = desc >> 25;
                const int q0 = (desc >> 11) & 127u, q1 = (desc >> 18) & 127u;
                const double Rk = __ldg(&a.R[k]);
#pragma unroll 1
                for (int q = q0; q <= q1; q++) {
                    const int wn = s_win[wid][q];
                    if (k >= (wn & 0xffff) && k < (wn >> 16)) {
                        // sin(pi (R - r) / (c tau)) of simulation.py:549, as sinpi of the normalised offset
                        const double sn = sinpi((Rk - s_r[wid][q]) * inv_ctau);
                        v += s_amp[wid][q] * (sn * sn);      // pulses in dict order, like the reference's i[k] +=
                    }
                }
            }
            // per-beam argmax over the lanes holding candidates of the same beam: peer groups by beam, then three
            // integer reductions (non-negative doubles order like their bit patterns): max high word, max low word
            // among those, min sample index among the exact ties -> the first maximum, like np.argmax
            const unsigned peers = __match_any_sync(0xffffffffu, owner);
            const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
            const unsigned vhi = (unsigned)(vb >> 32), vlo = (unsigned)vb;
            const unsigned mhi = __reduce_max_sync(peers, vhi);
            const unsigned mlo = __reduce_max_sync(peers, vhi == mhi ? vlo : 0u);
            const bool is_max = (vhi == mhi) && (vlo == mlo);
            const unsigned kmin = __reduce_min_sync(peers, is_max ? (unsigned)k : 0xffffffffu);
            if (owner >= 0 && lane == __ffs(peers) - 1) {
                const double vmax = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
                if (vmax > s_best[wid][owner]) { s_best[wid][owner] = vmax; s_kbest[wid][owner] = (int)kmin; }
            }
            __syncwarp();
        }
    }
    if (coop) {
        best = s_best[wid][lane];
        kbest = best > 0.0 ? s_kbest[wid][lane] : 0;
    }

    if (n_pulses > 0) {
        // ---- new range / intensity / label (simulation.py:151-188) -------------------------------------------------
        const double max_i = a.sensor->max_intensity[ch];
        const double min_i = a.sensor->min_intensity[ch];
        const double d_max = ((double)kbest / 10) - (ctau / 2);
        const double q1 = 1 - d_max / 120;
        double i_max = best + max_i * a.sensor->focal_slope[ch] * fabs(a.sensor->focal_offset[ch] - q1 * q1);
        i_max = i_max < min_i ? min_i : (i_max > max_i ? max_i : i_max);
        const long long new_i = (long long)i_max;       // int(): truncation
        if (fabs(d_max - d) < 2 * (1.0 / 10)) {
            out_l = 1.0f;
            att_new_i = new_i;                          // intensity_diff_sum += i_orig - new_i   (simulation.py:170)
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

    // ---- np.round of the intensity column (simulation.py:516); the threshold / FOV filters, which need the pre-pass
    // polynomial, are applied by k_keep so that the pre-pass can run next to the beam kernels -------------------------
    if (active) {
        out_i = rintf(out_i);
        if (!deferred && a.nocc) a.nocc[beg + i] = n_claim;
    }
    const bool counted = active && !deferred;      // a deferred beam is accounted for by the overflow kernel
    // one unrelated beam per thread: plain stores; integer atomics aggregated over lanes hitting the same counter
    if (counted) {
        float *row = a.aug + (beg + i) * 5;
        row[0] = out_x; row[1] = out_y; row[2] = out_z; row[3] = out_i; row[4] = out_l;
    }
    // label-1 beams per channel and the sum of their new intensities (simulation.py:170): integer atomics, order
    // independent => bit-reproducible
    const bool on = counted && att_new_i >= 0;
    const int key = b * LSS_N_CHANNELS + (ch < LSS_N_CHANNELS ? ch : 0);
    const unsigned mk = __match_any_sync(0xffffffffu, on ? key : -1);
    if (on && lane == __ffs(mk) - 1) atomicAdd(a.att_cnt + key, (unsigned)__popc(mk));
    const unsigned m = __match_any_sync(0xffffffffu, on ? b : -1);
    const unsigned sum = __reduce_add_sync(m, on ? (unsigned)att_new_i : 0u);
    if (on && lane == __ffs(m) - 1) atomicAdd(&a.att_sum[b], (unsigned long long)sum);
}

// ---------------------------------------------------------------------------------------------------------------------
// counting sort of the solve list by work class, so that the 32 beams of a solve-kernel warp cost about the same
// (the warp runs as long as its slowest lane).  hdr: [0] entries, [LIST_CLASSES + c] class counts (from the scan kernel),
// [2 * LIST_CLASSES + c] class cursors.  Order inside a class is arbitrary: the results do not depend on list order.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HIT_POS_PER_BEAM = 6;                 // capacity of the hit-position array per beam of the batch
constexpr int SORT_PER_THREAD = 8;
__global__ void __launch_bounds__(256) k_list_sort(const SolveItem *__restrict__ in, SolveItem *out, int *hdr, int cap)
{
    __shared__ int base[LIST_CLASSES], hist[LIST_CLASSES], blk[LIST_CLASSES];
    const int cnt = min(hdr[0], cap);
    const int first = blockIdx.x * (256 * SORT_PER_THREAD);
    if (first >= cnt) return;
    for (int c = threadIdx.x; c < LIST_CLASSES; c += 256) hist[c] = 0;
    if (threadIdx.x < 32) {                                  // exclusive scan of the class counts, one warp
        int run = 0;
        for (int c0 = 0; c0 < LIST_CLASSES; c0 += 32) {
            const int v = hdr[LIST_CLASSES + c0 + threadIdx.x];
            int incl = v;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, s);
                if ((int)threadIdx.x >= s) incl += t;
            }
            base[c0 + threadIdx.x] = run + incl - v;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    // rank inside the block (shared-memory counters, warp-aggregated), then ONE global cursor update per class and
    // block: the neighbouring beams of a cloud fall into few classes, per-entry global atomics would serialise
    const int lane = threadIdx.x & 31;
    int cls_of[SORT_PER_THREAD], rank[SORT_PER_THREAD];
#pragma unroll
    for (int k = 0; k < SORT_PER_THREAD; k++) {
        const int slot = first + k * 256 + threadIdx.x;
        const bool on = slot < cnt;
        const int cls = on ? (int)(in[slot].key >> 48) : -1;
        cls_of[k] = cls;
        const unsigned m = __match_any_sync(0xffffffffu, cls);
        const int leader = __ffs(m) - 1;
        int r = 0;
        if (on && lane == leader) r = atomicAdd(&hist[cls], __popc(m));
        r = __shfl_sync(0xffffffffu, r, leader);
        rank[k] = r + __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < LIST_CLASSES; c += 256)
        blk[c] = hist[c] ? atomicAdd(&hdr[2 * LIST_CLASSES + c], hist[c]) : 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_PER_THREAD; k++)
        if (cls_of[k] >= 0) {
            const int slot = first + k * 256 + threadIdx.x;
            const uint4 *src = reinterpret_cast<const uint4 *>(in + slot);
            uint4 *dst = reinterpret_cast<uint4 *>(out + base[cls_of[k]] + blk[cls_of[k]] + rank[k]);
            dst[0] = src[0];
            dst[1] = src[1];
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// keep pass: threshold filter on the ORIGINAL range (simulation.py:518-523), camera FOV filter (:532-537), per-tile
// channel histograms for the scatter pass, num_attenuated / num_removed.  One CTA per tile of 1024 rows: the tile's
// histogram rows are written, not accumulated, so they need no zero fill.  Runs after the beam kernels AND the
// pre-pass (which may have run concurrently with them on another stream).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) k_keep(DevArgs a)
{
    __shared__ unsigned h_keep[NBINS], h_all[NBINS];
    __shared__ int s_cnt[2];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int64_t beg = a.cloud_off[b];
    const int n = (int)(a.cloud_off[b + 1] - beg);
    if (tile * TILE >= n) return;
    if (threadIdx.x < NBINS) { h_keep[threadIdx.x] = 0; h_all[threadIdx.x] = 0; }
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = tile * TILE + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool active = i < n;
    bool keep = false, att = false;
    int ch = -1;
    if (active) {
        const float *in = a.pts + (beg + i) * 5;
        const float *row = a.aug + (beg + i) * 5;
        const float px = __ldcs(in), py = __ldcs(in + 1), pz = __ldcs(in + 2);
        ch = channel_bin(__ldcs(in + 4));
        const float out_x = row[0], out_y = row[1], out_z = row[2], out_i = row[3], out_l = row[4];
        keep = true;
        if (a.flags & LSS_FLAG_THRESHOLD_FILTER) {
            const float d32 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
            const double d = (double)d32;
            const double *p = a.thresh + 3 * b;
            const double d2 = (double)__fmul_rn(d32, d32);
            const double thr = __dadd_rn(__dadd_rn(__dmul_rn(p[0], d2), __dmul_rn(p[1], d)), p[2]);
            keep = (out_l == 2.0f) || ((double)out_i > thr);
        }
        att = keep && out_l == 1.0f && ch < LSS_N_CHANNELS;   // num_attenuated is counted BEFORE the FOV filter (:525)
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
        a.code_keep[beg + i] = keep ? (uint8_t)ch : (uint8_t)255;
        if (a.code_all) a.code_all[beg + i] = (uint8_t)ch;
    }
    // warp-aggregated shared-memory histograms
    {
        const int ck = (active && keep) ? ch : -1;
        const unsigned mk = __match_any_sync(0xffffffffu, ck);
        if (ck >= 0 && lane == __ffs(mk) - 1) atomicAdd(&h_keep[ck], (unsigned)__popc(mk));
        if (a.hist_all) {
            const unsigned ma = __match_any_sync(0xffffffffu, active ? ch : -1);
            if (active && lane == __ffs(ma) - 1) atomicAdd(&h_all[ch], (unsigned)__popc(ma));
        }
        const unsigned m_att = __ballot_sync(0xffffffffu, att);
        const unsigned m_rem = __ballot_sync(0xffffffffu, active && !keep);
        if (lane == 0) {
            if (m_att) atomicAdd(&s_cnt[0], __popc(m_att));
            if (m_rem) atomicAdd(&s_cnt[1], __popc(m_rem));
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS) {
        const size_t hrow = ((size_t)a.tile_base[b] + tile) * NBINS + threadIdx.x;
        a.hist_keep[hrow] = h_keep[threadIdx.x];
        if (a.hist_all) a.hist_all[hrow] = h_all[threadIdx.x];
    }
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&a.counters[2 * b + threadIdx.x], s_cnt[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------------
// tile scan: hist[tile][bin] -> exclusive destination offsets in (bin-major, tile-minor) order.  One CTA per cloud,
// one warp per bin at a time, tiles scanned 32 at a time.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_tile_scan(unsigned *hist, const int32_t *__restrict__ tile_base,
                                                     int32_t *counts /* or null */, double *stats, const int *counters,
                                                     const unsigned *att_cnt, const unsigned long long *att_sum,
                                                     const SensorConst *sensor)
{
    __shared__ unsigned bin_total[NBINS];
    const int b = blockIdx.x;
    const int n_tiles = tile_base[b + 1] - tile_base[b];
    unsigned *h = hist + (size_t)tile_base[b] * NBINS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < NBINS; c += 32) {
        unsigned run = 0;
        for (int t0 = 0; t0 < n_tiles; t0 += 32) {
            const int t = t0 + lane;
            const unsigned v = t < n_tiles ? h[(size_t)t * NBINS + c] : 0u;
            unsigned incl = v;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= s) incl += o; }
            if (t < n_tiles) h[(size_t)t * NBINS + c] = run + incl - v;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) bin_total[c] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int c = 0; c < NBINS; c++) { const unsigned t = bin_total[c]; bin_total[c] = run; run += t; }
        if (counts) counts[b] = (int32_t)run;
        if (stats) {
            const int n_att = counters[2 * b], n_rem = counters[2 * b + 1];
            // intensity_diff_sum = sum over attenuated beams of (0.9 * max_intensity - new_i)   (simulation.py:140,170)
            double sum = 0.0;
            for (int c = 0; c < LSS_N_CHANNELS; c++)
                sum += (double)att_cnt[b * LSS_N_CHANNELS + c] * (0.9 * sensor->max_intensity[c]);
            sum -= (double)att_sum[b];
            stats[4 * b + 3] = sum;
            stats[4 * b + 0] = (double)n_att;
            stats[4 * b + 1] = (double)n_rem;
            stats[4 * b + 2] = n_att > 0 ? (double)(long long)(sum / (double)n_att) : 0.0;   // int(sum / n), :527-530
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_tiles * NBINS; k += blockDim.x) h[k] += bin_total[k % NBINS];
}

// ---------------------------------------------------------------------------------------------------------------------
// stable scatter of one tile: row i with code c goes to tile_off[tile][c] + (number of earlier rows of the tile with
// the same code).  Grid (tiles, clouds), 1024 threads.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE) k_scatter(const float *__restrict__ aug, const uint8_t *__restrict__ code,
                                                   const unsigned *__restrict__ tile_off,
                                                   const int64_t *__restrict__ cloud_off,
                                                   const int32_t *__restrict__ tile_base,
                                                   float *__restrict__ out, const int32_t *__restrict__ nocc_in,
                                                   int32_t *__restrict__ nocc_out, int32_t *__restrict__ perm_out)
{
    __shared__ unsigned warp_cnt[TILE / 32][NBINS];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int64_t beg = cloud_off[b];
    const int n = (int)(cloud_off[b + 1] - beg);
    const int i = tile * TILE + threadIdx.x;
    if (tile * TILE >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < (TILE / 32) * NBINS; k += TILE) (&warp_cnt[0][0])[k] = 0;
    __syncthreads();
    int c = -1;
    if (i < n) {
        const int cc = code[beg + i];
        c = cc < NBINS ? cc : -1;
    }
    const unsigned m = __match_any_sync(0xffffffffu, c);
    const int rank = __popc(m & ((1u << lane) - 1u));
    if (c >= 0 && rank == 0) warp_cnt[warp][c] = __popc(m);
    __syncthreads();
    if (threadIdx.x < NBINS) {
        unsigned run = tile_off[((size_t)tile_base[b] + tile) * NBINS + threadIdx.x];
        for (int wv = 0; wv < TILE / 32; wv++) { const unsigned t = warp_cnt[wv][threadIdx.x]; warp_cnt[wv][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
    if (c >= 0) {
        const int64_t dst = beg + warp_cnt[warp][c] + rank;
        const float *s = aug + (beg + i) * 5;
        float *o = out + dst * 5;
#pragma unroll
        for (int q = 0; q < 5; q++) o[q] = s[q];
        if (nocc_out) nocc_out[dst] = nocc_in[beg + i];
        if (perm_out) perm_out[dst] = i;
    }
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
    int64_t aug, code_keep, code_all, nocc, hist_keep, hist_all, hist_rows, cloud_off, tile_base, order, thresh, counters,
        counters_bytes, ovf, prepass, prepass_bytes, total;
};

WsLayout ws_layout(int64_t n_total, int n_clouds)
{
    WsLayout w;
    w.hist_rows = n_total / TILE + (int64_t)n_clouds + 1;          // >= sum over clouds of ceil(n_b / TILE)
    int64_t o = 0;
    w.aug = o;        o = align_up(o + n_total * 5 * 4, 256);
    w.code_keep = o;  o = align_up(o + n_total, 256);
    w.code_all = o;   o = align_up(o + n_total, 256);
    w.nocc = o;       o = align_up(o + n_total * 4, 256);
    w.hist_keep = o;  o = align_up(o + w.hist_rows * NBINS * 4, 256);
    w.hist_all = o;   o = align_up(o + w.hist_rows * NBINS * 4, 256);
    w.cloud_off = o;  o = align_up(o + (int64_t)(n_clouds + 1) * 8, 256);
    w.tile_base = o;  o = align_up(o + (int64_t)(n_clouds + 1) * 4, 256);
    w.order = o;      o = align_up(o + (int64_t)n_clouds * LSS_N_CHANNELS * 4, 256);
    w.thresh = o;     o = align_up(o + (int64_t)n_clouds * 3 * 8, 256);
    // counters: int[B*2] | unsigned att_cnt[B*64] | unsigned long long att_sum[B]
    w.counters_bytes = align_up((int64_t)n_clouds * 2 * 4, 8) + (int64_t)n_clouds * LSS_N_CHANNELS * 4 + (int64_t)n_clouds * 8;
    w.counters = o;   o = align_up(o + w.counters_bytes, 256);
    // list header | overflow list | solve list, unsorted + sorted (every beam may have occluders)
    //   | hit particle indices (int32; HIT_POS_PER_BEAM per beam of the batch on average: the surveyed densities give 1-5 occluders
    //   on a third of the beams)
    w.ovf = o;        o = align_up(o + LIST_HDR_BYTES + (int64_t)OVF_LIST_CAP * 8 + 2 * n_total * (int64_t)sizeof(SolveItem) +
                                   (n_total * HIT_POS_PER_BEAM + 4096) * 4, 256);
    w.prepass_bytes = lss_prepass_ws_bytes(n_total, n_clouds);
    w.prepass = o;    o = align_up(o + w.prepass_bytes, 256);
    w.total = o;
    return w;
}

}  // namespace

int64_t lss_snowfall_ws_bytes(int64_t n_total, int n_clouds)
{
    if (n_total < 0 || n_clouds < 0) return -1;
    return ws_layout(n_total, n_clouds).total;
}

int64_t lss_snowfall_ws_cloud_off(int64_t n_total, int n_clouds) { return ws_layout(n_total, n_clouds).cloud_off; }

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
    if ((s.flags & LSS_FLAG_THRESHOLD_FILTER) && !s.h_thresh_poly && !s.d_thresh_poly && !(s.flags & LSS_FLAG_DEVICE_PREPASS))
        return lss_fail(e, LSS_ERR_INVALID_ARG, "threshold filter needs h_thresh_poly or LSS_FLAG_DEVICE_PREPASS");
    if ((s.flags & LSS_FLAG_CAMERA_FOV) && !e->has_camera)
        return lss_fail(e, LSS_ERR_NO_SENSOR, "camera calibration not set");
    const double div_rad = s.beam_divergence_deg * (LSS_PI / 180.0);
    if (!(div_rad > 0) || div_rad > s.ts->max_div_rad * (1 + 1e-12))
        return lss_fail(e, LSS_ERR_INVALID_ARG, "beam_divergence exceeds the value the table set was built for");
    const int max_tiles = (int)((max_n + TILE - 1) / TILE);
    std::vector<int32_t> h_tile_base(B + 1, 0);
    for (int b = 0; b < B; b++)
        h_tile_base[b + 1] = h_tile_base[b] + (int32_t)((s.h_cloud_offsets[b + 1] - s.h_cloud_offsets[b] + TILE - 1) / TILE);
    if ((s.d_out_perm || s.d_out_nocc) && !s.d_out_full)
        return lss_fail(e, LSS_ERR_INVALID_ARG, "d_out_perm / d_out_nocc need d_out_full");

    char *ws = (char *)s.d_workspace;
    float *d_aug = (float *)(ws + w.aug);
    uint8_t *d_code_keep = (uint8_t *)(ws + w.code_keep);
    const bool want_all = s.d_out_full != nullptr;
    uint8_t *d_code_all = want_all ? (uint8_t *)(ws + w.code_all) : nullptr;
    int32_t *d_nocc_tmp = s.d_out_nocc ? (int32_t *)(ws + w.nocc) : nullptr;
    unsigned *d_hist_keep = (unsigned *)(ws + w.hist_keep);
    unsigned *d_hist_all = want_all ? (unsigned *)(ws + w.hist_all) : nullptr;
    int64_t *d_off = (int64_t *)(ws + w.cloud_off);
    int32_t *d_tile_base = (int32_t *)(ws + w.tile_base);
    int32_t *d_order = (int32_t *)(ws + w.order);
    double *d_thresh = (double *)(ws + w.thresh);
    int *d_counters = (int *)(ws + w.counters);
    unsigned *d_att_cnt = (unsigned *)(ws + w.counters + align_up((int64_t)B * 2 * 4, 8));
    unsigned long long *d_att_sum = (unsigned long long *)((char *)d_att_cnt + (int64_t)B * LSS_N_CHANNELS * 4);

    LSS_CUDA_CHECK(e, lss_stage_upload(e, d_off, s.h_cloud_offsets, sizeof(int64_t) * (B + 1), stream));
    LSS_CUDA_CHECK(e, lss_stage_upload(e, d_tile_base, h_tile_base.data(), sizeof(int32_t) * (B + 1), stream));
    LSS_CUDA_CHECK(e, lss_stage_upload(e, d_order, s.h_order, sizeof(int32_t) * B * LSS_N_CHANNELS, stream));
    if (s.h_thresh_poly && !s.d_thresh_poly)
        LSS_CUDA_CHECK(e, lss_stage_upload(e, d_thresh, s.h_thresh_poly, sizeof(double) * 3 * B, stream));
    int *d_counts2 = (int *)(ws + w.ovf);                                   // [0] solve list, [1] overflow list
    {
        ZeroRegions z;
        z.add(d_counters, w.counters_bytes);
        z.add(s.d_out_stats, sizeof(double) * 4 * B);
        if (N == 0 || B == 0) {
            z.add(s.d_out_counts, sizeof(int32_t) * B);
            LSS_CUDA_CHECK(e, lss_zero_async(e, z, stream));
            return LSS_OK;
        }
        z.add(d_counts2, LIST_HDR_BYTES);                 // (the tile histograms are written whole by k_keep)
        LSS_CUDA_CHECK(e, lss_zero_async(e, z, stream));
    }

    DevArgs a;
    a.rec = s.ts->d_rec;
    a.tan = s.ts->d_tan;
    a.plane_off = s.ts->d_plane_off;
    a.zbase = s.ts->zbase;
    a.entries = s.ts->d_entries;
    a.bucket_start = s.ts->d_bucket_start;
    a.n_buckets = s.ts->n_buckets;
    a.n_planes = s.ts->n_planes;
    a.w = LSS_TWO_PI / s.ts->n_buckets;
    a.inv_w = s.ts->n_buckets / LSS_TWO_PI;
    a.pts = s.d_points;
    a.theta = s.d_theta;
    a.cloud_off = d_off;
    a.order = d_order;
    a.thresh = s.d_thresh_poly ? s.d_thresh_poly : d_thresh;
    a.sensor = e->d_sensor;
    a.camera = e->d_camera;
    a.R = e->d_R;
    a.wtab = e->d_wtab;
    a.half_div = (s.beam_divergence_deg / 2) * (LSS_PI / 180.0);
    a.div_rad = div_rad;
    a.flags = s.flags;
    a.aug = d_aug;
    a.code_keep = d_code_keep;
    a.code_all = d_code_all;
    a.nocc = d_nocc_tmp;
    a.hist_keep = d_hist_keep;
    a.hist_all = d_hist_all;
    a.tile_base = d_tile_base;
    a.stats = s.d_out_stats;
    a.counters = d_counters;
    a.att_cnt = d_att_cnt;
    a.att_sum = d_att_sum;
    a.status = e->d_status;
    unsigned long long *d_ovf_list = (unsigned long long *)(ws + w.ovf + LIST_HDR_BYTES);
    SolveItem *d_solve_list = (SolveItem *)(d_ovf_list + OVF_LIST_CAP);
    SolveItem *d_sorted_list = d_solve_list + N;
    a.hit_idx = (int *)(d_sorted_list + N);
    // Device pre-pass: plane + laser parameters + threshold polynomial (simulation.py:449-467), on the cloud as given.
    // Only k_keep needs its result, so it runs on one of the engine's high-priority side streams next to the beam kernels (a
    // chain of small latency-bound kernels).
    const bool device_prepass = (s.flags & LSS_FLAG_THRESHOLD_FILTER) && (s.flags & LSS_FLAG_DEVICE_PREPASS) &&
                                !s.h_thresh_poly && !s.d_thresh_poly;
    // Where to fork it: the persistent solve kernel holds every SM's registers until its last tile, so a chain forked
    // AFTER the scan (which would let the scan kernel compact the mounting-window points as a by-product,
    // PrepassIO::window_staged) finds no room for its 1024-thread CTAs and becomes the critical path (measured: step
    // 1.07 ms instead of 1.01).  Default: fork before the scan, whose CTAs retire continuously; LSS_FUSE_WINDOW=1 selects
    // the other order for experiments.
    static const bool fuse_window_env = getenv("LSS_FUSE_WINDOW") && getenv("LSS_FUSE_WINDOW")[0] == '1';
    const bool fuse_window = fuse_window_env && !s.h_plane_in;
    a.win_stage = nullptr;
    a.win_tile_cnt = nullptr;
    if (device_prepass && fuse_window) lss_prepass_window_staging(ws + w.prepass, N, B, &a.win_stage, &a.win_tile_cnt);
    cudaEvent_t ev_join = nullptr;
    auto fork_prepass = [&]() -> lss_status {
        cudaStream_t side = nullptr;
        cudaEvent_t ev_fork = nullptr;
        LSS_CUDA_CHECK(e, lss_side_stream(e, &side, &ev_fork, &ev_join));
        LSS_CUDA_CHECK(e, cudaEventRecord(ev_fork, stream));        // (with fuse_window: after the scan has staged the window points)
        LSS_CUDA_CHECK(e, cudaStreamWaitEvent(side, ev_fork, 0));
        PrepassIO io;
        io.h_plane_in = s.h_plane_in;
        io.h_ymins_in = s.h_ymins_in;
        io.d_poly_out = d_thresh;
        io.window_staged = a.win_stage != nullptr;
        lss_status ps = lss_prepass_run(e, s.d_points, d_off, nullptr, s.h_cloud_offsets, B, 0.5, s.noise_floor, 0, 0, 1,
                                io, ws + w.prepass, w.prepass_bytes, nullptr, side);
        const cudaError_t je = cudaEventRecord(ev_join, side);
        if (ps != LSS_OK || je != cudaSuccess) {
            cudaStreamWaitEvent(stream, ev_join, 0);                // never leave the side stream dangling
            return ps != LSS_OK ? ps : lss_fail(e, LSS_ERR_CUDA, "event record failed");
        }
        return LSS_OK;
    };
    if (device_prepass && !fuse_window) LSS_CHECK_STATUS(fork_prepass());
    a.hit_cap = (int)std::min<int64_t>(N * HIT_POS_PER_BEAM + 4096, 0x7fffffff);
    {
        KernelTimer kt(e, LSS_K_SNOWFALL, stream);
        const int items_cap = (int)std::min<int64_t>(N, 0x7fffffff);
        // 1. scan: all beams; the ones without occluders are finished, the others go to the solve list with their hit masks
        a.list_in = nullptr; a.count_in = nullptr; a.cap_in = 0;
        a.list_out = nullptr; a.count_out = nullptr; a.cap_out = 0;
        a.hdr = d_counts2;
        a.items_out = d_solve_list; a.items_in = nullptr; a.items_cap = items_cap;
        {
            KernelTimer ks(e, LSS_K_SCAN, stream);
            lss_launch_scan(a, max_n, B, stream);
        }
        if (device_prepass && fuse_window) LSS_CHECK_STATUS(fork_prepass());
        // 2. solve: the listed beams, sorted by work class, one warp per tile of 32 (persistent grid)
        k_list_sort<<<(unsigned)((N + 256 * SORT_PER_THREAD - 1) / (256 * SORT_PER_THREAD)), 256, 0, stream>>>(d_solve_list, d_sorted_list, d_counts2, items_cap);
        a.items_in = d_sorted_list; a.items_out = nullptr;
        a.count_in = d_counts2; a.cap_in = items_cap;
        a.list_out = d_ovf_list; a.count_out = d_counts2 + 1; a.cap_out = OVF_LIST_CAP;
        {
            KernelTimer ks(e, LSS_K_SOLVE, stream);
            lss_launch_solve(a, d_counts2 + 2, e->n_sm, stream);
        }
        // 3. overflow: beams with more occluders than the solve kernel's arena takes per beam (rare), round-1 list kernel
        a.list_in = d_ovf_list; a.count_in = d_counts2 + 1; a.cap_in = OVF_LIST_CAP;
        a.list_out = nullptr; a.count_out = nullptr; a.cap_out = 0;
        k_overflow<<<OVF_LIST_CAP / SNOW_TPB, SNOW_TPB, 0, stream>>>(a);
        e->launches += 1;           // (+ one each from the three KernelTimer brackets = 4 launches)
    }
    if (ev_join) LSS_CUDA_CHECK(e, cudaStreamWaitEvent(stream, ev_join, 0));
    {
        KernelTimer kt(e, LSS_K_FINALIZE, stream);
        k_keep<<<dim3(max_tiles, B), TILE, 0, stream>>>(a);
    }
    {
        KernelTimer kt(e, LSS_K_SORT, stream);
        k_tile_scan<<<B, 1024, 0, stream>>>(d_hist_keep, d_tile_base, s.d_out_counts, s.d_out_stats, d_counters,
                                            d_att_cnt, d_att_sum, e->d_sensor);
    }
    {
        KernelTimer kt(e, LSS_K_COMPACT, stream);
        k_scatter<<<dim3(max_tiles, B), TILE, 0, stream>>>(d_aug, d_code_keep, d_hist_keep, d_off, d_tile_base,
                                                           s.d_out_points, nullptr, nullptr, nullptr);
    }
    if (want_all) {     // un-filtered, channel-sorted debug views (tests): full rows, original index, occluder counts
        KernelTimer kt(e, LSS_K_COMPACT, stream);
        k_tile_scan<<<B, 1024, 0, stream>>>(d_hist_all, d_tile_base, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        k_scatter<<<dim3(max_tiles, B), TILE, 0, stream>>>(d_aug, d_code_all, d_hist_all, d_off, d_tile_base, s.d_out_full,
                                                           d_nocc_tmp, s.d_out_nocc, s.d_out_perm);
        e->launches++;
    }
    LSS_CUDA_CHECK(e, cudaGetLastError());
    return LSS_OK;
}
