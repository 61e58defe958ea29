// voxelize.cu -- point-range mask + voxelisation of (augmented) clouds on the device, so that the augmented batch goes
// from the augmentation kernels to the detector input without a host round trip.
//
// Replaces, for every cloud of a batch,
//   DataProcessor.mask_points_and_boxes_outside_range    lib/OpenPCDet/pcdet/datasets/processor/data_processor.py:78-91
//       (points part: common_utils.mask_points_by_range, lib/OpenPCDet/pcdet/utils/common_utils.py:60-63 -- x / y only,
//        both ends inclusive)
//   DataProcessor.transform_points_to_voxels              data_processor.py:115-143 -> VoxelGeneratorWrapper (:15-58)
//       -> spconv's point-to-voxel generator (third party, not vendored; rule restated in oracle/voxel.py):
//          float32 c = floor((p - range_min) / voxel_size) per axis, points outside the grid skipped, voxels numbered by
//          FIRST APPEARANCE in point order, at most max_voxels voxels (points of later voxels are skipped), the first
//          max_points points of a voxel kept in point order; coordinates stored (z, y, x)
//   the batch index column of DatasetTemplate.collate_batch  lib/OpenPCDet/pcdet/datasets/dataset.py:199-204
//
// The rule is sequential in the reference; its result only depends on, per voxel, the smallest point index (= order of
// first appearance) and the max_points smallest point indices (= the points kept).  Both are order-independent
// reductions:
//   k_vox_insert    hash table per cloud (open addressing, 64-bit voxel key): atomicMin of the point index, count
//   k_vox_flags / k_vox_scan / k_vox_assign   a point is "first of its voxel" iff the table's minimum is its own index;
//                   exclusive scan of those flags in point order = the voxel number; coordinates, counts
//   k_vox_cascade   the max_points smallest indices of every kept voxel: a cascade of atomicMin over max_points levels
//                   (the value displaced from / rejected by level t moves on to level t + 1: level t ends up with the
//                   (t+1)-th smallest index whatever the interleaving)
//   k_vox_write     every point finds its rank in its voxel's level list and copies its row there
// Everything is integer / float32 arithmetic without reassociation: bit-identical to the sequential rule.
#include "common.cuh"
#include <climits>

namespace {

constexpr int VTILE = 1024;
constexpr unsigned long long VOX_EMPTY = ~0ull;

struct VoxArgs {
    const float *pts;
    int F;
    const int64_t *cloud_off;
    const int32_t *cloud_cnt;      // optional: valid rows per cloud slot (slot-compacted input)
    float lo[3], hi[3], vs[3];
    int gs[3];
    int max_points, max_voxels, mask_xy, n_clouds;
    unsigned long long *h_key;     // [2N + B] hash slots; cloud b owns [2 off[b] + b, 2 off[b+1] + b + 1)
    int *h_first, *h_count, *h_vid;
    int *slot_of;                  // [N] slot of each point's voxel, -1 = point not in the grid
    int *tile_cnt, *tile_off;
    const int32_t *tile_base;      // [B+1]
    int *top;                      // [B * max_voxels * max_points]
    float *out_vox;                // [B * max_voxels * max_points * F]
    int32_t *out_coords;           // [B * max_voxels * 4]  (batch index, z, y, x)
    int32_t *out_num;              // [B * max_voxels]
    int32_t *out_nvox;             // [B]
};

__device__ __forceinline__ int cloud_rows(const VoxArgs &a, int b)
{
    return a.cloud_cnt ? a.cloud_cnt[b] : (int)(a.cloud_off[b + 1] - a.cloud_off[b]);
}

// voxel coordinate of a point, or false if it is masked / outside the grid
__device__ __forceinline__ bool voxel_of(const VoxArgs &a, const float *row, int &cx, int &cy, int &cz)
{
    const float x = row[0], y = row[1], z = row[2];
    if (a.mask_xy && !(x >= a.lo[0] && x <= a.hi[0] && y >= a.lo[1] && y <= a.hi[1])) return false;   // common_utils.py:60-63
    const float fx = floorf(__fdiv_rn(__fsub_rn(x, a.lo[0]), a.vs[0]));
    const float fy = floorf(__fdiv_rn(__fsub_rn(y, a.lo[1]), a.vs[1]));
    const float fz = floorf(__fdiv_rn(__fsub_rn(z, a.lo[2]), a.vs[2]));
    if (!(fx >= 0.0f && fx < (float)a.gs[0] && fy >= 0.0f && fy < (float)a.gs[1] && fz >= 0.0f && fz < (float)a.gs[2]))
        return false;
    cx = (int)fx; cy = (int)fy; cz = (int)fz;
    return true;
}

__global__ void k_fill32(uint32_t *p, unsigned long long n, uint32_t v)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

__global__ void __launch_bounds__(256) k_vox_insert(VoxArgs a)
{
    const int b = blockIdx.y;
    const int64_t beg = a.cloud_off[b];
    const int n = cloud_rows(a, b);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int cx, cy, cz, slot = -1;
    if (voxel_of(a, a.pts + (beg + i) * a.F, cx, cy, cz)) {
        const unsigned long long key = ((unsigned long long)cz * a.gs[1] + cy) * a.gs[0] + cx;
        const long long base = 2 * beg + b;
        const unsigned cap = (unsigned)(2 * (a.cloud_off[b + 1] - beg) + 1);
        unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ULL) >> 32) % cap;
        for (;;) {
            const unsigned long long prev = atomicCAS(&a.h_key[base + h], VOX_EMPTY, key);
            if (prev == VOX_EMPTY || prev == key) break;
            h = h + 1 == cap ? 0 : h + 1;
        }
        slot = (int)(base + h);
        atomicMin(&a.h_first[slot], i);
        atomicAdd(&a.h_count[slot], 1);
    }
    a.slot_of[beg + i] = slot;
}

__global__ void __launch_bounds__(VTILE) k_vox_flags(VoxArgs a)
{
    __shared__ int cnt;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int n = cloud_rows(a, b);
    const int n_tiles = a.tile_base[b + 1] - a.tile_base[b];
    if (tile >= n_tiles) return;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int i = tile * VTILE + threadIdx.x;
    bool first = false;
    if (i < n) {
        const int slot = a.slot_of[a.cloud_off[b] + i];
        first = slot >= 0 && a.h_first[slot] == i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) a.tile_cnt[a.tile_base[b] + tile] = cnt;
}

__global__ void __launch_bounds__(1024) k_vox_scan(VoxArgs a)
{
    __shared__ int wsum[32];
    __shared__ int run;
    const int b = blockIdx.x;
    const int t0 = a.tile_base[b], nt = a.tile_base[b + 1] - t0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) run = 0;
    __syncthreads();
    for (int base = 0; base < nt; base += 1024) {
        const int t = base + tid;
        const int v = t < nt ? a.tile_cnt[t0 + t] : 0;
        int incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, s); if (lane >= s) incl += u; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int o = run;
        for (int wv = 0; wv < warp; wv++) o += wsum[wv];
        if (t < nt) a.tile_off[t0 + t] = o + incl - v;
        __syncthreads();
        if (tid == 1023) run = o + incl;
        __syncthreads();
    }
    if (tid == 0) a.out_nvox[b] = run < a.max_voxels ? run : a.max_voxels;
}

__global__ void __launch_bounds__(VTILE) k_vox_assign(VoxArgs a)
{
    __shared__ int wcnt[VTILE / 32];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int n = cloud_rows(a, b);
    const int n_tiles = a.tile_base[b + 1] - a.tile_base[b];
    if (tile >= n_tiles) return;
    const int64_t beg = a.cloud_off[b];
    const int i = tile * VTILE + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int slot = -1;
    bool first = false;
    if (i < n) {
        slot = a.slot_of[beg + i];
        first = slot >= 0 && a.h_first[slot] == i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    if (!first) return;
    int vid = a.tile_off[a.tile_base[b] + tile] + __popc(m & ((1u << lane) - 1u));
    for (int wv = 0; wv < warp; wv++) vid += wcnt[wv];
    if (vid >= a.max_voxels) { a.h_vid[slot] = -1; return; }          // later voxels are skipped (and their points)
    a.h_vid[slot] = vid;
    int cx, cy, cz;
    voxel_of(a, a.pts + (beg + i) * a.F, cx, cy, cz);
    int32_t *c = a.out_coords + ((size_t)b * a.max_voxels + vid) * 4;
    c[0] = b; c[1] = cz; c[2] = cy; c[3] = cx;                         // collate_batch's batch index + spconv's (z, y, x)
    const int cnt = a.h_count[slot];
    a.out_num[(size_t)b * a.max_voxels + vid] = cnt < a.max_points ? cnt : a.max_points;
}

__global__ void __launch_bounds__(256) k_vox_cascade(VoxArgs a)
{
    const int b = blockIdx.y;
    const int64_t beg = a.cloud_off[b];
    const int n = cloud_rows(a, b);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int slot = a.slot_of[beg + i];
    if (slot < 0) return;
    const int vid = a.h_vid[slot];
    if (vid < 0) return;
    int *top = a.top + ((size_t)b * a.max_voxels + vid) * a.max_points;
    int v = i;
    for (int t = 0; t < a.max_points; t++) {
        const int old = atomicMin(&top[t], v);
        if (old == INT_MAX) break;          // the level was empty: v stays, nothing moves on
        if (old > v) v = old;               // v took the level: the displaced index moves on (else v itself does)
    }
}

__global__ void __launch_bounds__(256) k_vox_write(VoxArgs a)
{
    const int b = blockIdx.y;
    const int64_t beg = a.cloud_off[b];
    const int n = cloud_rows(a, b);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int slot = a.slot_of[beg + i];
    if (slot < 0) return;
    const int vid = a.h_vid[slot];
    if (vid < 0) return;
    const int *top = a.top + ((size_t)b * a.max_voxels + vid) * a.max_points;
    for (int t = 0; t < a.max_points; t++) {
        const int v = top[t];
        if (v == i) {
            const float *row = a.pts + (beg + i) * a.F;
            float *o = a.out_vox + (((size_t)b * a.max_voxels + vid) * a.max_points + t) * a.F;
            for (int f = 0; f < a.F; f++) o[f] = row[f];
            return;
        }
        if (v > i) return;                  // levels ascend: this point is not among the first max_points
    }
}

inline int64_t align_up(int64_t v, int64_t al) { return (v + al - 1) / al * al; }

struct VoxLayout { int64_t off, key, first, count, vid, slot_of, tile_base, tile_cnt, tile_off, top, total, n_slots, n_tiles; };

VoxLayout vox_layout(int64_t n_total, int n_clouds, int max_points, int max_voxels)
{
    VoxLayout L;
    L.n_slots = 2 * n_total + n_clouds + 1;
    L.n_tiles = n_total / VTILE + n_clouds + 1;
    int64_t o = 0;
    L.off = o;       o = align_up(o + (int64_t)(n_clouds + 1) * 8, 256);
    L.key = o;       o = align_up(o + L.n_slots * 8, 256);
    L.first = o;     o = align_up(o + L.n_slots * 4, 256);
    L.count = o;     o = align_up(o + L.n_slots * 4, 256);
    L.vid = o;       o = align_up(o + L.n_slots * 4, 256);
    L.slot_of = o;   o = align_up(o + n_total * 4, 256);
    L.tile_base = o; o = align_up(o + (int64_t)(n_clouds + 1) * 4, 256);
    L.tile_cnt = o;  o = align_up(o + L.n_tiles * 4, 256);
    L.tile_off = o;  o = align_up(o + L.n_tiles * 4, 256);
    L.top = o;       o = align_up(o + (int64_t)n_clouds * max_voxels * max_points * 4, 256);
    L.total = o;
    return L;
}

cudaError_t fill32(lss_engine *e, void *p, unsigned long long words, uint32_t v, cudaStream_t st)
{
    if (!words) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<unsigned long long>((words + 1023) / 1024, 148 * 16);
    k_fill32<<<blocks, 256, 0, st>>>((uint32_t *)p, words, v);
    e->launches++;
    return cudaGetLastError();
}

}  // namespace

extern "C" {

int64_t lss_voxelize_workspace_bytes(int64_t n_total, int n_clouds, int max_points_per_voxel, int max_voxels)
{
    if (n_total < 0 || n_clouds < 0 || max_points_per_voxel <= 0 || max_voxels <= 0) return -1;
    return vox_layout(n_total, n_clouds, max_points_per_voxel, max_voxels).total;
}

lss_status lss_voxelize_batch(lss_engine *e, const float *d_points, int n_features, const int64_t *h_cloud_offsets,
                              const int32_t *d_cloud_counts, int n_clouds, const float *h_point_cloud_range,
                              const float *h_voxel_size, int max_points_per_voxel, int max_voxels, int mask_xy_range,
                              float *d_out_voxels, int32_t *d_out_coords, int32_t *d_out_num_points,
                              int32_t *d_out_n_voxels, void *d_workspace, int64_t workspace_bytes, void *stream)
{
    if (!e) return LSS_ERR_INVALID_ARG;
    if (!h_cloud_offsets || n_clouds < 0 || !h_point_cloud_range || !h_voxel_size || !d_out_voxels || !d_out_coords ||
        !d_out_num_points || !d_out_n_voxels || !d_workspace)
        return lss_fail(e, LSS_ERR_INVALID_ARG, "null argument");
    if (n_features < 3 || max_points_per_voxel <= 0 || max_voxels <= 0)
        return lss_fail(e, LSS_ERR_INVALID_ARG, "n_features >= 3, max_points_per_voxel > 0, max_voxels > 0 required");
    if (h_cloud_offsets[0] != 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud_offsets[0] must be 0");
    const int B = n_clouds;
    const int64_t N = h_cloud_offsets[B];
    if (N >= (1LL << 30)) return lss_fail(e, LSS_ERR_INVALID_ARG, "batch too large");
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    const VoxLayout L = vox_layout(N, B, max_points_per_voxel, max_voxels);
    if (workspace_bytes < L.total) return lss_fail(e, LSS_ERR_WORKSPACE, "workspace too small");
    VoxArgs a;
    a.pts = d_points;
    a.F = n_features;
    a.cloud_cnt = d_cloud_counts;
    a.max_points = max_points_per_voxel;
    a.max_voxels = max_voxels;
    a.mask_xy = mask_xy_range;
    a.n_clouds = B;
    for (int j = 0; j < 3; j++) {
        a.lo[j] = h_point_cloud_range[j];
        a.hi[j] = h_point_cloud_range[3 + j];
        a.vs[j] = h_voxel_size[j];
        if (!(a.vs[j] > 0.0f) || !(a.hi[j] > a.lo[j])) return lss_fail(e, LSS_ERR_INVALID_ARG, "bad range / voxel size");
        // data_processor.py:117-118: np.round((range[3:6] - range[0:3]) / voxel_size), float32
        const float q = (a.hi[j] - a.lo[j]) / a.vs[j];
        a.gs[j] = (int)nearbyintf(q);
        if (a.gs[j] <= 0 || q > 2.0e9f) return lss_fail(e, LSS_ERR_INVALID_ARG, "bad grid size");
    }
    char *ws = (char *)d_workspace;
    int64_t *d_off = (int64_t *)(ws + L.off);
    a.cloud_off = d_off;
    a.h_key = (unsigned long long *)(ws + L.key);
    a.h_first = (int *)(ws + L.first);
    a.h_count = (int *)(ws + L.count);
    a.h_vid = (int *)(ws + L.vid);
    a.slot_of = (int *)(ws + L.slot_of);
    a.tile_base = (const int32_t *)(ws + L.tile_base);
    a.tile_cnt = (int *)(ws + L.tile_cnt);
    a.tile_off = (int *)(ws + L.tile_off);
    a.top = (int *)(ws + L.top);
    a.out_vox = d_out_voxels;
    a.out_coords = d_out_coords;
    a.out_num = d_out_num_points;
    a.out_nvox = d_out_n_voxels;

    const size_t n_vox_all = (size_t)B * max_voxels;
    if (B == 0) return LSS_OK;
    std::vector<int32_t> h_tb(B + 1, 0);
    int64_t max_n = 0;
    for (int b = 0; b < B; b++) {
        const int64_t nb = h_cloud_offsets[b + 1] - h_cloud_offsets[b];
        if (nb < 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud_offsets must be non-decreasing");
        max_n = std::max(max_n, nb);
        h_tb[b + 1] = h_tb[b] + (int32_t)((nb + VTILE - 1) / VTILE);
    }
    LSS_CUDA_CHECK(e, lss_stage_upload(e, d_off, h_cloud_offsets, sizeof(int64_t) * (B + 1), st));
    LSS_CUDA_CHECK(e, lss_stage_upload(e, ws + L.tile_base, h_tb.data(), sizeof(int32_t) * (B + 1), st));
    {
        KernelTimer kt(e, LSS_K_VOXEL, st);
        LSS_CUDA_CHECK(e, fill32(e, a.h_key, (unsigned long long)L.n_slots * 2, 0xffffffffu, st));
        LSS_CUDA_CHECK(e, fill32(e, a.h_first, (unsigned long long)L.n_slots, (uint32_t)INT_MAX, st));
        LSS_CUDA_CHECK(e, fill32(e, a.h_count, (unsigned long long)L.n_slots, 0u, st));
        LSS_CUDA_CHECK(e, fill32(e, a.top, (unsigned long long)n_vox_all * max_points_per_voxel, (uint32_t)INT_MAX, st));
        LSS_CUDA_CHECK(e, fill32(e, d_out_voxels, (unsigned long long)n_vox_all * max_points_per_voxel * n_features, 0u, st));
        LSS_CUDA_CHECK(e, fill32(e, d_out_coords, (unsigned long long)n_vox_all * 4, 0u, st));
        LSS_CUDA_CHECK(e, fill32(e, d_out_num_points, (unsigned long long)n_vox_all, 0u, st));
        if (max_n > 0) {
            const dim3 g256((unsigned)((max_n + 255) / 256), B), gt((unsigned)((max_n + VTILE - 1) / VTILE), B);
            k_vox_insert<<<g256, 256, 0, st>>>(a);
            k_vox_flags<<<gt, VTILE, 0, st>>>(a);
            k_vox_scan<<<B, 1024, 0, st>>>(a);
            k_vox_assign<<<gt, VTILE, 0, st>>>(a);
            k_vox_cascade<<<g256, 256, 0, st>>>(a);
            k_vox_write<<<g256, 256, 0, st>>>(a);
            e->launches += 5;
        } else {
            LSS_CUDA_CHECK(e, fill32(e, d_out_n_voxels, (unsigned long long)B, 0u, st));
        }
    }
    LSS_CUDA_CHECK(e, cudaGetLastError());
    return LSS_OK;
}

}  // extern "C"
