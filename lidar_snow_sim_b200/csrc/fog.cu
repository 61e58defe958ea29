// fog.cu -- batched fog simulation on device-resident clouds.
//
// Reference: lib/LiDAR_fog_sim/fog_simulation.py
//   P_R_fog_hard :183-189   Beer-Lambert attenuation of the hard target: I <- round(exp(-2 alpha r_0) I), float32
//   P_R_fog_soft :192-296   soft target from the integral look-up table keyed by round(r_0, 1): if the fog response
//                           beats the attenuated return, the point moves to the fog distance (with a noise factor drawn
//                           from the caller's generator) and takes the response as intensity; optional gain
//   simulate_fog :299-316   hard, then soft
//
// Kernels (one thread per point, tiles of 256 points staged through shared memory for coalesced I/O, HBM bound: reads F x 4 B, writes F x 8 B + 1 B per point):
//   k_fog_count   fog mask per point -> fog points per tile
//   k_fog_scan    per cloud: exclusive scan of the tile counts (rank of a fog point in POINT ORDER = position of its draw
//                 in the generator's stream), num_fog_responses
//   k_fog_apply   everything: hard, soft, noise (k-th PCG64 output by jump-ahead), min / max response, max intensity
//   k_fog_gain    intensity *= 255 / ceil(max intensity)                                            (:282-285)
//
// Numerics: float32 where NumPy 2 computes in float32 (r_0, exp, the hard-target product, r_0 ** 2, r_0 -/+ noise),
// float64 elsewhere, no FMA contraction.  Two places are host-defined in the reference and therefore parity by
// tolerance, not by bits: the float32 np.exp and the scalar float32 power r_0 ** 2 (neither is correctly
// rounded on every host; the device uses the correctly rounded values), and pow() of the v2 / v3 noise factors.
#include "common.cuh"

namespace {

constexpr int FOG_TILE = 256;                     // points per CTA; rows are staged through shared memory (coalesced I/O)
constexpr int FOG_STAGE_F = 8;                    // ... for up to this many features; wider rows are accessed in place
constexpr int LUT_N = 2001;

struct FogArgs {
    const float *pts;            // [N * F]
    int F;
    const int64_t *cloud_off;    // [B + 1] device
    const int32_t *tile_base;    // [B + 1] device
    const double *lut;           // [LUT_N * 2] (fog_distance, fog_response)
    double alpha, beta, beta_0;
    int hard, soft, gain;
    int noise, variant;          // variant 1..4; 4 = externally drawn values (ext_noise, by rank)
    const unsigned long long *rng;   // [B * 4] PCG64 state_hi, state_lo, inc_hi, inc_lo per cloud, or null
    const double *ext_noise;     // [N] by (cloud offset + rank) or null
    double *out;                 // [N * F]
    uint8_t *mask;               // [N]
    int32_t *rank;               // [N] or null
    int *tile_cnt;               // [tiles]
    int *tile_off;               // [tiles] exclusive, per cloud
    unsigned long long *info;    // [B * 4] bit patterns: min response, max response, count, max intensity (ordered)
};

// correctly rounded float32 exp via float64 (np.exp on float32 is a host SIMD kernel, < 1 ulp but host defined)
__device__ __forceinline__ float exp32(float x) { return (float)exp((double)x); }

// order-preserving map of a double onto unsigned integers (atomicMax over values of either sign)
__device__ __forceinline__ unsigned long long ord_of(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to(unsigned long long o)
{
    const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)b);
}

struct Soft { bool fog; double resp, fog_distance; float r0, hard_i; };

__device__ __forceinline__ Soft soft_target(const FogArgs &a, const float *row)
{
    Soft s;
    const float x = row[0], y = row[1], z = row[2], I = row[3];
    // np.linalg.norm(float32 rows): sqrt((x*x + y*y) + z*z) in float32
    s.r0 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    s.hard_i = I;
    if (a.hard) {
        // np.round(np.exp(-2 * alpha * r_0) * I): -2*alpha is a Python float (weak), the array op runs in float32
        const float coef = (float)(-2.0 * a.alpha);
        s.hard_i = rintf(__fmul_rn(exp32(__fmul_rn(coef, s.r0)), I));
    }
    s.fog = false; s.resp = 0.0; s.fog_distance = 0.0;
    if (a.soft) {
        // key = float(str(round(r_0, 1))), min(key, 200): index rint(float32(r_0 * 10)) capped at 2000   (:212-214)
        const float k10 = rintf(__fmul_rn(s.r0, 10.0f));
        int k = (k10 >= (float)(LUT_N - 1)) ? LUT_N - 1 : (int)k10;
        k = k < 0 ? 0 : k;
        s.fog_distance = a.lut[2 * k];
        double r = __dmul_rn(a.lut[2 * k + 1], (double)I);                  // * original intensity       (:216)
        r = __dmul_rn(r, (double)__fmul_rn(s.r0, s.r0));                    // * r_0 ** 2 (float32)
        r = __ddiv_rn(__dmul_rn(r, a.beta), a.beta_0);
        s.resp = fmin(r, 255.0);                                            // :219
        s.fog = s.resp > (double)s.hard_i;                                  // :221
    }
    return s;
}

// coalesced copy of the tile's rows into shared memory; returns the row of thread `threadIdx.x`
__device__ __forceinline__ const float *stage_rows(const FogArgs &a, int64_t first_row, int rows, float *s_in)
{
    if (a.F > FOG_STAGE_F) return a.pts + (first_row + threadIdx.x) * a.F;
    const float *src = a.pts + first_row * a.F;
    const int nf = rows * a.F;
    for (int f = threadIdx.x; f < nf; f += FOG_TILE) s_in[f] = __ldcs(src + f);
    __syncthreads();
    return s_in + threadIdx.x * a.F;
}

__global__ void __launch_bounds__(FOG_TILE) k_fog_count(FogArgs a)
{
    __shared__ float s_in[FOG_TILE * FOG_STAGE_F];
    __shared__ int cnt;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int64_t beg = a.cloud_off[b];
    const int n = (int)(a.cloud_off[b + 1] - beg);
    if (tile * FOG_TILE >= n) return;
    if (threadIdx.x == 0) cnt = 0;
    const int i = tile * FOG_TILE + threadIdx.x;
    const float *row = stage_rows(a, beg + (int64_t)tile * FOG_TILE, min(FOG_TILE, n - tile * FOG_TILE), s_in);
    __syncthreads();
    bool fog = false;
    if (i < n) fog = soft_target(a, row).fog;
    const unsigned m = __ballot_sync(0xffffffffu, fog);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) a.tile_cnt[a.tile_base[b] + tile] = cnt;
}

__global__ void __launch_bounds__(1024) k_fog_scan(FogArgs a)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int b = blockIdx.x;
    const int t0 = a.tile_base[b], nt = a.tile_base[b + 1] - t0;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nt; base += 1024) {
        const int t = base + threadIdx.x;
        const int v = t < nt ? a.tile_cnt[t0 + t] : 0;
        int incl = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, s);
            if ((threadIdx.x & 31) >= s) incl += u;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += warp_tot[w];
        if (t < nt) a.tile_off[t0 + t] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.info[4 * b + 2] = (unsigned long long)carry;       // num_fog_responses
        a.info[4 * b] = ~0ull;                                // running minimum of the responses
    }
}

// PCG64 (XSL-RR 128/64, numpy's default bit generator): the (k+1)-th state after `st` by jump-ahead, then its output
__device__ __forceinline__ void mul128(unsigned long long ah, unsigned long long al, unsigned long long bh,
                                       unsigned long long bl, unsigned long long &rh, unsigned long long &rl)
{
    rl = al * bl;
    rh = __umul64hi(al, bl) + ah * bl + al * bh;
}
__device__ __forceinline__ void add128(unsigned long long &ah, unsigned long long &al, unsigned long long bh,
                                       unsigned long long bl)
{
    const unsigned long long lo = al + bl;
    ah = ah + bh + (lo < al ? 1ull : 0ull);
    al = lo;
}
__device__ double pcg64_kth_double(const unsigned long long *st, unsigned long long k)
{
    unsigned long long cur_mh = 0x2360ED051FC65DA4ull, cur_ml = 0x4385DF649FCCF645ull;   // multiplier
    unsigned long long cur_ph = st[2], cur_pl = st[3];                                      // increment
    unsigned long long acc_mh = 0, acc_ml = 1, acc_ph = 0, acc_pl = 0;
    unsigned long long delta = k + 1;                                                       // draw k uses state k+1
    while (delta) {
        if (delta & 1ull) {
            mul128(acc_mh, acc_ml, cur_mh, cur_ml, acc_mh, acc_ml);
            unsigned long long th, tl;
            mul128(acc_ph, acc_pl, cur_mh, cur_ml, th, tl);
            add128(th, tl, cur_ph, cur_pl);
            acc_ph = th; acc_pl = tl;
        }
        unsigned long long th, tl, oh = cur_mh, ol = cur_ml;
        add128(oh, ol, 0, 1);                                      // cur_mult + 1
        mul128(oh, ol, cur_ph, cur_pl, th, tl);
        cur_ph = th; cur_pl = tl;
        mul128(cur_mh, cur_ml, cur_mh, cur_ml, cur_mh, cur_ml);
        delta >>= 1;
    }
    unsigned long long sh, sl;
    mul128(acc_mh, acc_ml, st[0], st[1], sh, sl);
    add128(sh, sl, acc_ph, acc_pl);
    const unsigned long long x = sh ^ sl;
    const unsigned rot = (unsigned)(sh >> 58);
    const unsigned long long out = (x >> rot) | (x << ((64u - rot) & 63u));
    return (double)(out >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(FOG_TILE) k_fog_apply(FogArgs a)
{
    __shared__ float s_in[FOG_TILE * FOG_STAGE_F];
    __shared__ double s_out[FOG_TILE * FOG_STAGE_F];
    __shared__ int warp_cnt[FOG_TILE / 32];
    __shared__ unsigned long long s_min, s_max, s_imax;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int64_t beg = a.cloud_off[b];
    const int n = (int)(a.cloud_off[b + 1] - beg);
    if (tile * FOG_TILE >= n) return;
    if (threadIdx.x == 0) { s_min = ~0ull; s_max = 0ull; s_imax = 0ull; }
    const int i = tile * FOG_TILE + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool active = i < n;
    const int rows = min(FOG_TILE, n - tile * FOG_TILE);
    const bool staged = a.F <= FOG_STAGE_F;
    Soft s;
    s.fog = false;
    const float *row = stage_rows(a, beg + (int64_t)tile * FOG_TILE, rows, s_in);
    if (active) s = soft_target(a, row);
    const unsigned m = __ballot_sync(0xffffffffu, active && s.fog);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    double out_i = 0.0;
    if (active) {
        double *o = staged ? s_out + threadIdx.x * a.F : a.out + (beg + i) * a.F;
        if (!a.soft) {                                             // hard only: the rows stay float32 valued
            for (int f = 0; f < a.F; f++) o[f] = (double)row[f];
            o[3] = (double)s.hard_i;
            out_i = o[3];
            a.mask[beg + i] = 0;
        } else if (!s.fog) {
            for (int f = 0; f < a.F; f++) o[f] = (double)row[f];   // augmented_pc[i] = pc[i]               (:276)
            o[3] = (double)s.hard_i;
            out_i = o[3];
            a.mask[beg + i] = 0;
            if (a.rank) a.rank[beg + i] = -1;
        } else {
            int rank = a.tile_off[a.tile_base[b] + tile] + __popc(m & ((1u << lane) - 1u));
            for (int w = 0; w < wid; w++) rank += warp_cnt[w];
            const double scaling = __ddiv_rn(s.fog_distance, (double)s.r0);                  // :227
            double px = __dmul_rn((double)row[0], scaling), py = __dmul_rn((double)row[1], scaling),
                   pz = __dmul_rn((double)row[2], scaling);
            if (a.noise > 0) {
                double factor = 1.0;
                bool have = true;
                if (a.variant == 4) {
                    // additive = r_noise * beta(2, 20) with r_noise = 10 (:207-208, :258-260); the beta draws are the
                    // caller's (rejection sampling does not jump ahead)
                    have = a.ext_noise != nullptr;
                    if (have) {
                        const double additive = __dmul_rn(10.0, a.ext_noise[beg + rank]);
                        factor = __ddiv_rn(__dadd_rn(s.fog_distance, additive), s.fog_distance);
                    }
                } else {
                    have = a.rng != nullptr || a.ext_noise != nullptr;
                    const double u = a.ext_noise ? a.ext_noise[beg + rank]
                                                 : (a.rng ? pcg64_kth_double(a.rng + 4 * b, (unsigned long long)rank) : 0.0);
                    if (a.variant == 1) {
                        // RNG.uniform(low=r_0 - noise, high=r_0 + noise): float32 limits, low + (high - low) * u
                        const double low = (double)__fsub_rn(s.r0, (float)a.noise);
                        const double high = (double)__fadd_rn(s.r0, (float)a.noise);
                        const double dn = __dadd_rn(low, __dmul_rn(__dsub_rn(high, low), u));
                        factor = __ddiv_rn((double)s.r0, dn);                                   // :241-242
                    } else if (a.variant == 2) {
                        const double power = __dadd_rn(-1.0, __dmul_rn(2.0, u));               // uniform(-1, 1)
                        factor = pow(fmax(1.0, (double)a.noise / 5), power);                    // :247-248
                    } else {
                        const double power = __dadd_rn(-0.5, __dmul_rn(1.5, u));               // uniform(-0.5, 1)
                        factor = pow(fmax(1.0, (double)a.noise * 4 / 10), power);               // :253-254
                    }
                }
                if (have) { px = __dmul_rn(px, factor); py = __dmul_rn(py, factor); pz = __dmul_rn(pz, factor); }
            }
            o[0] = px; o[1] = py; o[2] = pz; o[3] = s.resp;
            if (a.F > 4) o[4] = (double)row[4];                    // only the 5th feature is carried over (:231-233)
            for (int f = 5; f < a.F; f++) o[f] = 0.0;
            out_i = s.resp;
            a.mask[beg + i] = 1;
            if (a.rank) a.rank[beg + i] = rank;
            atomicMin(&s_min, ord_of(s.resp));
            atomicMax(&s_max, ord_of(s.resp));
        }
        if (a.gain) atomicMax(&s_imax, ord_of(out_i));
    }
    __syncthreads();
    if (staged) {                                                  // coalesced store of the tile's float64 rows
        double *dst = a.out + (beg + (int64_t)tile * FOG_TILE) * a.F;
        const int nf = rows * a.F;
        for (int f = threadIdx.x; f < nf; f += FOG_TILE) __stcs(dst + f, s_out[f]);
    }
    if (threadIdx.x == 0) {
        if (s_min != ~0ull) { atomicMin(&a.info[4 * b], s_min); atomicMax(&a.info[4 * b + 1], s_max); }
        if (a.gain) atomicMax(&a.info[4 * b + 3], s_imax);
    }
}

__global__ void __launch_bounds__(FOG_TILE) k_fog_gain(FogArgs a)
{
    const int b = blockIdx.y;
    const int64_t beg = a.cloud_off[b];
    const int n = (int)(a.cloud_off[b + 1] - beg);
    const int i = blockIdx.x * FOG_TILE + threadIdx.x;
    if (i >= n) return;
    // max_intensity = np.ceil(max(augmented_pc[:, 3])); gain_factor = 255 / max_intensity; column *= gain_factor
    const double gain_factor = __ddiv_rn(255.0, ceil(ord_to(a.info[4 * b + 3])));
    double *o = a.out + (beg + i) * a.F + 3;
    *o = __dmul_rn(*o, gain_factor);
}

// info: ordered bit patterns -> (min response, max response, count) as doubles; a cloud without fog points reports
// (inf, 0, 0) like the reference's initial values (:201-203)
__global__ void k_fog_info(FogArgs a, int B, double *info_out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const unsigned long long cnt = a.info[4 * b + 2];
    info_out[3 * b] = cnt ? ord_to(a.info[4 * b]) : __longlong_as_double(0x7ff0000000000000LL);
    info_out[3 * b + 1] = cnt ? ord_to(a.info[4 * b + 1]) : 0.0;
    info_out[3 * b + 2] = (double)cnt;
}

struct FogLayout { int64_t off, tile_base, tile_cnt, tile_off, info, rng, total; };

FogLayout fog_layout(int64_t n_total, int n_clouds)
{
    FogLayout L;
    const int64_t tiles = n_total / FOG_TILE + n_clouds + 1;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { const int64_t at = o; o = (o + bytes + 255) / 256 * 256; return at; };
    L.off = take((int64_t)(n_clouds + 1) * 8);
    L.tile_base = take((int64_t)(n_clouds + 1) * 4);
    L.tile_cnt = take(tiles * 4);
    L.tile_off = take(tiles * 4);
    L.info = take((int64_t)n_clouds * 4 * 8);
    L.rng = take((int64_t)n_clouds * 4 * 8);
    L.total = o;
    return L;
}

}  // namespace

int64_t lss_fog_workspace_bytes(int64_t n_total, int n_clouds)
{
    if (n_total < 0 || n_clouds < 0) return -1;
    return fog_layout(n_total, n_clouds).total;
}

lss_status lss_fog_batch(lss_engine *e, const float *d_points, int n_features, const int64_t *h_cloud_offsets,
                                    int n_clouds, double alpha, double beta, double beta_0, const double *d_lut,
                                    uint32_t flags, int noise, int noise_variant, const uint64_t *h_rng_state,
                                    const double *d_ext_noise, double *d_out, uint8_t *d_out_fog_mask,
                                    int32_t *d_out_rank, double *d_out_info, void *d_workspace, int64_t workspace_bytes,
                                    void *stream)
{
    if (!e) return LSS_ERR_INVALID_ARG;
    if (!h_cloud_offsets || n_clouds < 0 || !d_out || !d_out_fog_mask || !d_out_info || !d_workspace)
        return lss_fail(e, LSS_ERR_INVALID_ARG, "null argument");
    if (n_features < 4 || n_features > 16) return lss_fail(e, LSS_ERR_INVALID_ARG, "n_features must be in 4..16");
    if (n_clouds > 65535) return lss_fail(e, LSS_ERR_INVALID_ARG, "at most 65535 clouds per call");
    if (h_cloud_offsets[0] != 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud_offsets[0] must be 0");
    const bool soft = flags & LSS_FOG_SOFT;
    if (soft && !d_lut) return lss_fail(e, LSS_ERR_INVALID_ARG, "the soft target needs the integral look-up table");
    if (noise > 0 && soft && (noise_variant < 1 || noise_variant > 4))
        return lss_fail(e, LSS_ERR_INVALID_ARG, "noise variant must be 1..4 (NotImplementedError in the reference)");
    const int B = n_clouds;
    const int64_t N = h_cloud_offsets[B];
    int64_t max_n = 0;
    std::vector<int32_t> h_tb(B + 1, 0);
    for (int b = 0; b < B; b++) {
        const int64_t n = h_cloud_offsets[b + 1] - h_cloud_offsets[b];
        if (n < 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud_offsets must be non-decreasing");
        if (n >= (1LL << 31)) return lss_fail(e, LSS_ERR_INVALID_ARG, "cloud too large");
        max_n = std::max(max_n, n);
        h_tb[b + 1] = h_tb[b] + (int32_t)((n + FOG_TILE - 1) / FOG_TILE);
    }
    if (!d_points && N > 0) return lss_fail(e, LSS_ERR_INVALID_ARG, "null points");
    const FogLayout L = fog_layout(N, B);
    if (workspace_bytes < L.total) return lss_fail(e, LSS_ERR_WORKSPACE, "workspace too small");
    if (B == 0) return LSS_OK;
    DeviceGuard g(e->device);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_workspace;

    FogArgs a;
    a.pts = d_points;
    a.F = n_features;
    a.cloud_off = (const int64_t *)(ws + L.off);
    a.tile_base = (const int32_t *)(ws + L.tile_base);
    a.lut = d_lut;
    a.alpha = alpha; a.beta = beta; a.beta_0 = beta_0;
    a.hard = (flags & LSS_FOG_HARD) ? 1 : 0;
    a.soft = soft ? 1 : 0;
    a.gain = (soft && (flags & LSS_FOG_GAIN)) ? 1 : 0;
    a.noise = noise;
    a.variant = noise_variant;
    a.rng = nullptr;
    a.ext_noise = d_ext_noise;
    a.out = d_out;
    a.mask = d_out_fog_mask;
    a.rank = d_out_rank;
    a.tile_cnt = (int *)(ws + L.tile_cnt);
    a.tile_off = (int *)(ws + L.tile_off);
    a.info = (unsigned long long *)(ws + L.info);

    LSS_CUDA_CHECK(e, lss_stage_upload(e, ws + L.off, h_cloud_offsets, sizeof(int64_t) * (B + 1), st));
    LSS_CUDA_CHECK(e, lss_stage_upload(e, ws + L.tile_base, h_tb.data(), sizeof(int32_t) * (B + 1), st));
    if (h_rng_state && soft && noise > 0 && noise_variant != 4 && !d_ext_noise) {
        LSS_CUDA_CHECK(e, lss_stage_upload(e, ws + L.rng, h_rng_state, sizeof(uint64_t) * 4 * B, st));
        a.rng = (const unsigned long long *)(ws + L.rng);
    }
    {
        ZeroRegions z;
        z.add(ws + L.info, (size_t)B * 4 * 8);
        LSS_CUDA_CHECK(e, lss_zero_async(e, z, st));
    }
    const int max_tiles = (int)((max_n + FOG_TILE - 1) / FOG_TILE);
    if (N > 0) {
        KernelTimer kt(e, LSS_K_FOG, st);
        if (soft) {
            k_fog_count<<<dim3(max_tiles, B), FOG_TILE, 0, st>>>(a);
            k_fog_scan<<<B, 1024, 0, st>>>(a);
            e->launches += 2;
        }
        k_fog_apply<<<dim3(max_tiles, B), FOG_TILE, 0, st>>>(a);
        if (a.gain) { k_fog_gain<<<dim3(max_tiles, B), FOG_TILE, 0, st>>>(a); e->launches++; }
    }
    k_fog_info<<<(B + 127) / 128, 128, 0, st>>>(a, B, d_out_info);
    e->launches++;
    LSS_CUDA_CHECK(e, cudaGetLastError());
    return LSS_OK;
}
