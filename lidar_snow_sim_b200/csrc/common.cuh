// common.cuh -- shared types of the snowfall engine (device + host side of the C ABI in include/lidar_snow_sim.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/lidar_snow_sim.h"

#define LSS_PI 3.141592653589793
#define LSS_TWO_PI 6.283185307179586
#define LSS_M_EXT 1230            // samples of the range grid R (tools/snowfall/simulation.py:111-116)
#define LSS_ANG_MARGIN 1e-5       // rad; safety margin of the float32 broad phase (0.3 % of the 3 mrad beam)

// Exact per-particle record used by the float64 narrow phase.  All angles in [0, 2 pi).
struct __align__(16) ParticleRec {
    double phi;       // azimuth of the disk centre            (simulation.py:351-352)
    double rho;       // planar range sqrt(x^2 + y^2)           (simulation.py:332,413)
    double alpha;     // angular half width asin(r / rho)
    double t_right;   // tangent angles, (right, left) ordered as geometry.py:32-80 leaves them
    double t_left;
    double r;         // disk radius
};

// Broad-phase entry: one per (particle, azimuth bucket it can touch).  16 B -> one LDG.128 per candidate.
//   x = rho rounded DOWN to float32 (entries of a bucket are sorted by it)
//   y = phi - bucket centre, wrapped to (-pi, pi]
//   z = alpha + max_beam_divergence/2 + margin, rounded UP
//   w = index of the ParticleRec (bit pattern of an int32)
typedef float4 BroadEntry;

struct TableSet {
    int n_planes = 0;
    int n_buckets = 0;              // azimuth buckets per plane (power of two not required)
    double max_div_rad = 0.0;
    int64_t n_particles = 0;
    int64_t n_entries = 0;
    ParticleRec *d_rec = nullptr;       // [n_particles]
    BroadEntry *d_entries = nullptr;    // [n_entries]
    int32_t *d_bucket_start = nullptr;  // [n_planes * (n_buckets + 1)] global entry index
    int64_t bytes = 0;
};

struct SensorConst {
    double focal_offset[LSS_N_CHANNELS];   // (1 - focal_distance*100/13100)^2   (simulation.py:74-76)
    double focal_slope[LSS_N_CHANNELS];
    double min_intensity[LSS_N_CHANNELS];
    double max_intensity[LSS_N_CHANNELS];
};

struct CameraConst {
    float M[12];      // (V2C^T R0^T) as 4x3 row-major: rect = [x y z 1] . M     (calibration_kitti.py:65-73)
    float P2[12];     // 3x4 row-major                                           (calibration_kitti.py:75-84)
    int img_h, img_w;
};

struct lss_engine {
    int device = 0;
    bool has_sensor = false;
    bool has_camera = false;
    SensorConst sensor;
    CameraConst camera;
    SensorConst *d_sensor = nullptr;
    CameraConst *d_camera = nullptr;
    double *d_R = nullptr;              // range grid, LSS_M_EXT doubles
    int *d_status = nullptr;            // latched asynchronous device status
    std::map<int, TableSet> tables;
    int next_table_id = 1;
    int64_t launches = 0;
    std::string last_error;
    // optional per-kernel timing (lss_set_profiling): CUDA events on the launching stream around every kernel
    bool profiling = false;
    struct TimedLaunch { int kernel; cudaEvent_t beg, end; };
    std::vector<TimedLaunch> timed;
    double kernel_ms[16] = {0};
    int64_t kernel_calls[16] = {0};
};

enum { LSS_K_SORT = 0, LSS_K_PREPASS = 1, LSS_K_SNOWFALL = 2, LSS_K_COMPACT = 3, LSS_K_FINALIZE = 4, LSS_K_WET = 5,
       LSS_K_COUNT = 6 };

struct KernelTimer {        // RAII: records begin/end events when profiling is on
    lss_engine *e; cudaStream_t s; int idx = -1;
    KernelTimer(lss_engine *e_, int kernel, cudaStream_t s_) : e(e_), s(s_) {
        e->launches++;
        if (!e->profiling) return;
        lss_engine::TimedLaunch t; t.kernel = kernel;
        cudaEventCreate(&t.beg); cudaEventCreate(&t.end);
        cudaEventRecord(t.beg, s);
        e->timed.push_back(t); idx = (int)e->timed.size() - 1;
    }
    ~KernelTimer() { if (idx >= 0) cudaEventRecord(e->timed[idx].end, s); }
};

#define LSS_CUDA_CHECK(e, call)                                                                          \
    do {                                                                                                 \
        cudaError_t _err = (call);                                                                       \
        if (_err != cudaSuccess) {                                                                       \
            char _buf[512];                                                                              \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_err),      \
                     __FILE__, __LINE__);                                                                \
            (e)->last_error = _buf;                                                                      \
            return LSS_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static inline lss_status lss_fail(lss_engine *e, lss_status s, const char *msg)
{
    if (e) e->last_error = msg;
    return s;
}

// implemented in tables.cu
lss_status lss_build_tables(lss_engine *e, TableSet &ts, const double *d_xyr, const int64_t *h_plane_offsets,
                            cudaStream_t stream);
// implemented in snowfall.cu
struct SnowfallArgs {
    const TableSet *ts;
    const float *d_points;
    const int64_t *h_cloud_offsets;
    int n_clouds;
    const int32_t *h_order;
    double beam_divergence_deg;
    const float *d_theta;
    const double *h_thresh_poly;
    double noise_floor;
    uint32_t flags;
    float *d_out_points;
    int32_t *d_out_counts;
    double *d_out_stats;
    float *d_out_full;
    int32_t *d_out_perm;
    int32_t *d_out_nocc;
    void *d_workspace;
    int64_t workspace_bytes;
};
lss_status lss_snowfall_run(lss_engine *e, const SnowfallArgs &a, cudaStream_t stream);
// implemented in prepass.cu
struct CloudPre {            // per-cloud scratch / results, float64
    double w[3], h;          // plane
    double nw;               // |w|
    int n_window;            // points in the mounting window
    int n_ground;
    double ymax;             // |max(I / cos)|           (histogram range, augmentation.py:233)
    double lin[2];           // first regression  I/cos ~ lin0 * d + lin1       (augmentation.py:216-219)
    double pmin[2];          // second regression over the per-range-bin minima (augmentation.py:249)
    double poly[3];          // np.polyfit(d, noise*cos, 2): highest power first  (simulation.py:467)
    float z_med, mad;
    int best_trial;
    int flat;                // flat-earth fallback taken
};
// p . w exactly as written, without FMA contraction, so that every kernel classifies a point identically
__device__ __forceinline__ double lss_plane_dot(double x, double y, double z, const double *w)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(x, w[0]), __dmul_rn(y, w[1])), __dmul_rn(z, w[2]));
}
int64_t lss_prepass_ws_bytes(int64_t n_total, int n_clouds);
lss_status lss_prepass_run(lss_engine *e, const float *d_pts, const int64_t *d_cloud_off, const int32_t *d_cloud_cnt,
                           const int64_t *h_cloud_off, int n_clouds, double delta, double noise_floor, int flat_earth,
                           int range64, int raise_few_ground, const double *h_plane_in,
                           double *d_poly_out, double *d_plane_out, void *d_ws, int64_t ws_bytes, void **cloudpre_out,
                           cudaStream_t stream);
int64_t lss_snowfall_ws_bytes(int64_t n_total, int n_clouds);
