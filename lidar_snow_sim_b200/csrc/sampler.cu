// sampler.cu -- snowflake table sampler: sequential dart throwing of non-overlapping disks
// (tools/snowfall/sampling.py:90-194), host-native C++ with a uniform grid instead of the reference's O(N^2) scan.
//
// The random stream is NumPy's: PCG64 (XSL-RR 128/64) + Generator.uniform + Generator.exponential (256-layer
// ziggurat), consumed in exactly the reference's order, so that for the same np.random.Generator state the accepted
// disks, their order and their count are the reference's.  Accept / reject decisions use the reference's own
// expressions evaluated in float64 without fused multiply-add (this file is compiled with -ffp-contract=off).
// The only values that are not guaranteed bit-identical are cos/sin of the centre angle: NumPy dispatches float64
// cos/sin to SIMD kernels on some hosts (<= 1 ulp from libm); everything else is libm-free arithmetic.
//
// The grid is a pure accelerator: a new disk can only overlap accepted disks whose centres lie within
// (r_new + r_max) of its own centre, and every such disk lives in the 3x3 cell neighbourhood because the cell size
// (0.25 m) exceeds 2 * r_max = 2 cm by far.
#include "common.cuh"
#include "zig_tables.h"
#include <cmath>
#include <cstring>
#include <thread>

namespace {

typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state, inc;
    inline uint64_t next64()
    {
        const u128 mult = ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL;
        state = state * mult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((-rot) & 63));
    }
    inline double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
    inline double uniform(double low, double high) { return low + (high - low) * next_double(); }
    double standard_exponential()
    {
        for (;;) {
            uint64_t ri = next64();
            ri >>= 3;
            const unsigned idx = (unsigned)(ri & 0xFF);
            ri >>= 8;
            double we, fe0, fe1;
            memcpy(&we, &ZIG_WE_BITS[idx], 8);
            const double x = (double)ri * we;
            if (ri < ZIG_KE[idx]) return x;
            if (idx == 0) return ZIG_EXP_R - log1p(-next_double());
            memcpy(&fe0, &ZIG_FE_BITS[idx - 1], 8);
            memcpy(&fe1, &ZIG_FE_BITS[idx], 8);
            if ((fe0 - fe1) * next_double() + fe1 < exp(-x)) return x;
        }
    }
    inline double exponential(double scale) { return scale * standard_exponential(); }
};

struct Grid {
    double R0, cell;
    int n;
    std::vector<int32_t> head;     // per cell: index of the most recent disk, -1 = empty
    std::vector<int32_t> next;     // per disk
    Grid(double R0_, double cell_) : R0(R0_), cell(cell_)
    {
        n = (int)std::ceil(2 * R0 / cell) + 2;
        head.assign((size_t)n * n, -1);
    }
    inline int coord(double v) const
    {
        int c = (int)std::floor((v + R0) / cell) + 1;
        return c < 0 ? 0 : (c >= n ? n - 1 : c);
    }
};

}  // namespace

extern "C" {

// dart_throwing(occupancy_ratio, precipitation_rate, R_0, rng, distribution)     (sampling.py:90-194)
//   distribution: 0 = 'gunn' (Gunn & Marshall 1958), 1 = 'sekhon' (Sekhon & Srivastava 1970)
//   pcg_state:    in/out, 4 x uint64 = PCG64 {state_hi, state_lo, inc_hi, inc_lo} of a numpy Generator
//   h_xyr:        out, capacity rows of (x, y, r) float64
// Returns LSS_ERR_WORKSPACE (and the required row count in *n_out is NOT known) if capacity is exceeded.
// The reference squares SCALARS with `** 2` (python float.__pow__ / numpy's scalar power), i.e. through libm's pow(),
// which is within 0.52 ulp but not always the correctly rounded x*x; GCC would fold pow(x, 2.0) into x*x, so the call
// goes through a volatile pointer.  (Array expressions like (sx - x) ** 2 use np.square = x*x, sampling.py:170.)
static double (*volatile libm_pow)(double, double) = pow;
static inline double sq(double v) { return libm_pow(v, 2.0); }

lss_status lss_dart_throwing(double occupancy_ratio, double precipitation_rate, double R_0, int distribution,
                             uint64_t *pcg_state, double *h_xyr, int64_t capacity, int64_t *n_out)
{
    if (!pcg_state || !h_xyr || !n_out || !(occupancy_ratio > 0) || !(R_0 > 0) || !(precipitation_rate > 0))
        return LSS_ERR_INVALID_ARG;
    double rate;
    if (distribution == 0) rate = 25.5 * pow(precipitation_rate, -0.48);        // sampling.py:81-87
    else if (distribution == 1) rate = 22.9 * pow(precipitation_rate, -0.45);   // sampling.py:72-78
    else return LSS_ERR_INVALID_ARG;
    const double scale_cm = 1 / rate;                                            // sampling.py:115
    const double PI = 3.141592653589793;
    Pcg64 g;
    g.state = ((u128)pcg_state[0] << 64) | pcg_state[1];
    g.inc = ((u128)pcg_state[2] << 64) | pcg_state[3];

    Grid grid(R_0, 0.25);
    grid.next.reserve(1 << 16);
    int64_t n = 0;
    double area_occupied = 0.0;
    const double area_occupied_global = occupancy_ratio * PI * sq(R_0);          // sampling.py:124
    const double R0sq = sq(R_0);
    lss_status st = LSS_OK;
    while (area_occupied < area_occupied_global) {
        const double length = sqrt(g.uniform(0, R0sq));                          // :145
        const double angle = g.uniform(0, 2) * PI;                               // :146
        const double x = length * cos(angle);
        const double y = length * sin(angle);
        double dia;
        do { dia = g.exponential(scale_cm * 10); } while (dia > 20);             // :151-154 (mm)
        dia = dia / 1000;                                                        // :157
        const double height = g.uniform(-dia / 2, dia / 2);                      // :160
        const double half = dia / 2;
        const double disk_radius = sqrt(sq(half) - sq(height));                 // :163
        if (sq(x) + sq(y) <= sq(disk_radius)) continue;                          // :166
        const int cx = grid.coord(x), cy = grid.coord(y);
        bool overlap = false;
        for (int gy = cy - 1; gy <= cy + 1 && !overlap; gy++) {
            if (gy < 0 || gy >= grid.n) continue;
            for (int gx = cx - 1; gx <= cx + 1 && !overlap; gx++) {
                if (gx < 0 || gx >= grid.n) continue;
                for (int32_t k = grid.head[(size_t)gy * grid.n + gx]; k >= 0; k = grid.next[k]) {
                    const double dx = h_xyr[3 * k] - x, dy = h_xyr[3 * k + 1] - y, rr = h_xyr[3 * k + 2] + disk_radius;
                    if (dx * dx + dy * dy <= rr * rr) { overlap = true; break; }   // :170
                }
            }
        }
        if (overlap) continue;
        if (n >= capacity) { st = LSS_ERR_WORKSPACE; break; }
        h_xyr[3 * n] = x;
        h_xyr[3 * n + 1] = y;
        h_xyr[3 * n + 2] = disk_radius;
        grid.next.push_back(grid.head[(size_t)cy * grid.n + cx]);
        grid.head[(size_t)cy * grid.n + cx] = (int32_t)n;
        n++;
        area_occupied += PI * sq(disk_radius);                                   // :181-182
    }
    pcg_state[0] = (uint64_t)(g.state >> 64);
    pcg_state[1] = (uint64_t)g.state;
    *n_out = n;
    return st;
}

// Many planes at once (the 64 planes of one (rate, velocity) configuration, sampling.py:410-413), one host thread
// per plane up to n_threads.  pcg_states: n_planes x 4 uint64 (in/out).  Plane k is written to
// h_xyr + 3 * k * capacity_per_plane; h_counts[k] rows are valid.
lss_status lss_dart_throwing_planes(int n_planes, double occupancy_ratio, double precipitation_rate, double R_0,
                                    int distribution, uint64_t *pcg_states, double *h_xyr, int64_t capacity_per_plane,
                                    int64_t *h_counts, int n_threads)
{
    if (n_planes <= 0 || !pcg_states || !h_xyr || !h_counts) return LSS_ERR_INVALID_ARG;
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > n_planes) n_threads = n_planes;
    std::vector<int> status(n_planes, LSS_OK);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t]() {
            for (int k = t; k < n_planes; k += n_threads)
                status[k] = lss_dart_throwing(occupancy_ratio, precipitation_rate, R_0, distribution,
                                              pcg_states + 4 * k, h_xyr + 3 * (size_t)k * capacity_per_plane,
                                              capacity_per_plane, h_counts + k);
        });
    }
    for (auto &th : pool) th.join();
    for (int k = 0; k < n_planes; k++)
        if (status[k] != LSS_OK) return (lss_status)status[k];
    return LSS_OK;
}

}  // extern "C"
