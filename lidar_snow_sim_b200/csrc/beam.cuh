// beam.cuh -- argument block, constants and small device helpers shared by the per-beam kernels
// (snowfall.cu: scan / overflow / keep / scatter; solve.cu: the dense solve kernel).
#pragma once
#include "common.cuh"

// One beam of the solve list, written by the scan kernel (32 bytes).  The scan has already walked the beam's whole
// bucket prefix and tested every candidate exactly, so it hands over WHICH particles hit (their indices, in prefix
// order, in the hit array) and the azimuth it used.
struct __align__(16) SolveItem {
    unsigned long long key;       // work class << 48 | cloud << 32 | row
    int hit_off;                  // first of the beam's L entries of hit_idx[]
    int L;                        // occluders
    float th32;                   // beam azimuth in [0, 2 pi) as the scan used it
    int pad0;
    long long pad1;
};

// argument block of the per-beam kernels (global type: it crosses translation units)
struct DevArgs {
    // tables
    const ParticleRec *rec;
    const ParticleTan *tan;
    const int64_t *plane_off;    // [n_planes + 1] first particle of each plane (entries hold plane-local indices)
    float zbase;                 // decode base of the entries' half width (lss_decode)
    const BroadEntry *entries;
    const int32_t *bucket_start;
    int n_buckets;
    int n_planes;
    double inv_w, w;
    // per call
    const float *pts;            // rows in input order
    const float *theta;          // optional, input order
    const int64_t *cloud_off;    // [B+1] device
    const int32_t *order;        // [B*64] device
    const double *thresh;        // [B*3] device or null
    const SensorConst *sensor;
    const CameraConst *camera;
    const double *R;
    const double2 *wtab;         // [1230] (sin, cos) of pi * R[k] / (c tau): waveform phase table of the solve kernel
    double half_div;             // radians(beam_divergence / 2)
    double div_rad;              // radians(beam_divergence)
    uint32_t flags;
    float *aug;                  // [N*5] augmented rows, input order
    uint8_t *code_keep;          // [N] channel bin of a kept row, 255 = dropped
    uint8_t *code_all;           // optional [N] channel bin of every row (un-filtered debug output)
    int32_t *nocc;               // optional [N], input order
    unsigned *hist_keep;         // [sum of tiles * NBINS]; cloud b owns rows tile_base[b] .. tile_base[b+1]
    unsigned *hist_all;          // optional
    const int32_t *tile_base;    // [B+1] device
    double *stats;               // [B*4]: num_attenuated, num_removed, avg_diff, diff_sum
    int *counters;               // [B*2]: num_attenuated (threshold-kept), num_removed
    unsigned *att_cnt;           // [B*64] label-1 beams per channel (all of them, simulation.py:170)
    unsigned long long *att_sum; // [B] sum of their new integer intensities
    int *status;
    // work lists (cloud << 32 | row): the scan kernel defers every beam that has occluders to the dense solve kernel,
    // which in turn defers beams with more than SOLVE_LCAP occluders to the overflow kernel
    const unsigned long long *list_in;
    const int *count_in;
    int cap_in;
    unsigned long long *list_out;
    int *count_out;
    int cap_out;
    // solve list (scan kernel -> sort -> solve kernel); hdr = the list header ints (LIST_HDR_BYTES)
    SolveItem *items_out;
    const SolveItem *items_in;
    int *hdr;                    // [0] listed beams, [1] overflow beams, [2] tile cursor, [3] hit positions used, class counts ...
    int items_cap;
    int *hit_idx;                // particle indices of the hits of the listed beams
    int hit_cap;
    // optional by-product of the scan kernel for the concurrent pre-pass: the mounting-window points of calculate_plane
    // (tools/wet_ground/planes.py:21-27) compacted per 32-row tile (prepass.cu, PrepassIO::window_staged)
    float *win_stage;
    int *win_tile_cnt;
};

namespace {

constexpr int SNOW_TPB = 128;
constexpr int SNOW_WARPS = SNOW_TPB / 32;
constexpr int TILE = 1024;                          // rows per scatter tile
constexpr int NBINS = LSS_N_CHANNELS + 1;           // + "not a valid channel" (sorted last)
constexpr int POOL = 128;                           // pulses a warp publishes per cooperative batch
constexpr int CCAP = 512;                           // candidate samples a warp evaluates per cooperative batch
constexpr int SLOW_CAP = 128;                       // occluders per beam held by the overflow kernel (per-thread lists)
constexpr int OVF_LIST_CAP = 1 << 16;               // beams the overflow kernel can take per call
constexpr int LIST_HDR_BYTES = 2048;  // ints: [0] solve count, [1] overflow count, [C..2C) class counts, [2C..3C) cursors
constexpr int LIST_CLASSES = 128;  // solve list is counting-sorted by work class (occluder count) before the solve kernel

__device__ __forceinline__ void raise_status(int *status, int code) { atomicMax(status, code); }

__device__ __forceinline__ int channel_bin(float ch)
{
    int c = (int)ch;
    return (ch >= 0.0f && ch < 64.0f && (float)c == ch) ? c : LSS_N_CHANNELS;
}

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


// Beam azimuth: (float)atan2((double)y, (double)x) -- the correctly rounded float32 of the float64 atan2
// (simulation.py:91 uses a host-dependent float32 np.arctan2).
//
// CUDA's float64 atan2 was 26-36 % of the scan kernel's instructions.  Fast path: a = min/max of |x|, |y|, nearest table
// point c = i / 32, atan(a) = atan(c) + atan(t) with t = (a - c) / (1 + a c) = (mn - c mx) / (mx + c mn) (ONE division),
// |t| <= 1/64, so four terms of the series give atan(t) to 6e-18; quadrant by symmetry.  Error <= 5e-16 absolute
// (checked against extended precision on 2e6 arguments: 4.4e-16).  If the float64 value lies within 1e-13 relative of a
// float32 rounding boundary -- where that error could change the float32 result -- the library atan2 decides
// (~3e-6 of the beams), so the result is the library's everywhere.
__device__ __noinline__ float azimuth32_slow(float y, float x)
{
    return (float)atan2((double)y, (double)x);
}

// atan(i / 32), i = 0 .. 32 (global memory -> L1: lanes index it divergently, the constant cache would serialise them)
__device__ const double ATAN_I32[33] = {
        0.0, 0.031239833430268277, 0.06241880999595735, 0.09347678115858947,
        0.12435499454676144, 0.15499674192394097, 0.18534794999569476, 0.21535769969773805,
        0.24497866312686414, 0.2741674511196588, 0.3028848683749714, 0.3310960767041321,
        0.35877067027057225, 0.38588266939807375, 0.4124104415973873, 0.43833655985795783,
        0.4636476090008061, 0.48833395105640554, 0.5123894603107377, 0.5358112379604637,
        0.5585993153435624, 0.5807563535676704, 0.6022873461349642, 0.6231993299340659,
        0.6435011087932844, 0.6632029927060933, 0.6823165548747481, 0.7008544078844502,
        0.7188299996216245, 0.7362574289814281, 0.7531512809621944, 0.7695264804056583,
        0.7853981633974483};

__device__ __forceinline__ float azimuth32(float yf, float xf)
{
    const float axf = fabsf(xf), ayf = fabsf(yf);
    const float mxf = fmaxf(axf, ayf), mnf = fminf(axf, ayf);
    // zeros, infinities, NaNs, denormal-range ratios: the library handles the special cases
    if (!(mnf > 0.0f) || !(mxf < 3.0e38f) || !(mnf > mxf * 1e-30f) || xf != xf || yf != yf) return azimuth32_slow(yf, xf);
    const int i = (int)rintf(__fdividef(mnf, mxf) * 32.0f);
    const double c = (double)i * (1.0 / 32.0), mx = (double)mxf, mn = (double)mnf;
    const double t = fma(-c, mx, mn) / fma(c, mn, mx);
    const double t2 = t * t;
    double p = fma(t2, 1.0 / 9.0, -1.0 / 7.0);
    p = fma(t2, p, 1.0 / 5.0);
    p = fma(t2, p, -1.0 / 3.0);
    double r = __ldg(&ATAN_I32[i]) + fma(t * t2, p, t);
    if (ayf > axf) r = 1.5707963267948966 - r;
    if (xf < 0.0f) r = LSS_PI - r;
    if (yf < 0.0f) r = -r;
    const float f = (float)r;
    // how far is r from the nearest float32 rounding boundary?  spacing of f's binade: 2^(e - 23)
    const float ulp = __int_as_float(max((__float_as_int(fabsf(f)) & 0x7f800000) - (23 << 23), 1 << 23));
    const double slack = 0.5 * (double)ulp - fabs(r - (double)f);
    if (slack < 1e-13 * fabs(r)) return azimuth32_slow(yf, xf);
    return f;
}

}  // namespace

// solve.cu: the dense solve kernel over the (sorted) solve list; beams it cannot take (more than SOLVE_LCAP occluders or
// a bucket prefix longer than it tracks) go to list_out for the overflow kernel.  hdr[2] is its tile cursor (zeroed).
void lss_launch_scan(const DevArgs &a, int64_t max_rows, int n_clouds, cudaStream_t stream);
void lss_launch_solve(const DevArgs &a, int *tile_cursor, int n_sm, cudaStream_t stream);
