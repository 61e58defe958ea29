// solve.cu -- the two per-beam kernels of the snowfall path.
//
//   k_scan   every beam of the batch, one thread per beam, rows in INPUT order: float32 range / azimuth, walk of the
//            beam's azimuth bucket of its channel's snowflake plane (float32 broad phase, exact float64 disk / wedge test)
//            over the WHOLE prefix of entries nearer than the target.  Beams without occluder (~2/3) are finished here;
//            the others are pushed to the solve list together with what the walk found: the particle indices of the
//            hits (in prefix order) and the azimuth.                         (tools/snowfall/simulation.py:80-101, 329-390)
//   k_solve  the listed beams, sorted by work class: tangent angles of the hits, nearest-first claiming of the beam's
//            angular sub-intervals, summed sin^2 waveform + argmax, relabel / move the point, label-1 statistics.
//                                                                              (simulation.py:118-188, 231-295, 391-424)
//
// Design of k_solve (round 2; the round-1 kernel kept per-thread lists in local memory -- 180 MB of it across the
// resident threads, thrashing L2 --, walked the bucket a second time, published one descriptor per waveform sample
// lane-serially and evaluated a float64 sinpi per sample and pulse):
//   * persistent grid, one warp = one tile of 32 listed beams, tiles handed out by an atomic cursor (the list is sorted
//     costliest class first, so the tail is cheap tiles);
//   * NO local memory: the beams of a warp share a shared-memory arena of ARENA slots, allocated exactly
//     (occluders + 1 per beam) with a warp scan of the counts the scan kernel delivered;
//   * the hits are loaded COOPERATIVELY: arena slot s is filled by lane s mod 32, whatever beam it belongs to (owner by
//     a shuffle binary search over the offsets, the slot's particle = the r-th hit index the scan stored for the owner),
//     so the index -> record loads of all beams are in flight together; the owner then orders its few slots by range;
//   * nearest-first claiming runs in place in the arena: the union list lives in the slots of the already processed
//     hits, pulses (range, ratio) are compacted to the front;
//   * waveform: sin(pi (R_k - r) / (c tau)) = sin(pi a_k) cos(pi b) - cos(pi a_k) sin(pi b) with a_k = R_k / (c tau) from a
//     1230-entry table (host, extended precision) and b = r / (c tau) evaluated once per pulse (one sincospi), so a
//     sample costs a 16-byte load and six float64 operations instead of a sinpi;
//   * the sample axis is cut into pieces with a fixed set of active pulses; on a piece the summed waveform is a single
//     sinusoid, so only the three samples around its analytic peak (clipped to the piece) are evaluated -- exactly, in
//     dict order like the reference's i[k] +=; first maximum wins, np.argmax -- instead of whole windows; the pieces of
//     all 32 beams are handled one per lane, with uniform code.
//
// Beams the arena cannot take (more than SOLVE_LCAP occluders) go to the overflow list and are redone by the
// overflow kernel (snowfall.cu, k_overflow: round 1's list kernel), which has no such limit below 128.
#include "beam.cuh"

namespace {

constexpr int SOLVE_TPB = 128;
constexpr int SOLVE_WARPS = SOLVE_TPB / 32;
#ifndef LSS_SOLVE_CTAS
#define LSS_SOLVE_CTAS 6
#endif
#ifndef LSS_SOLVE_ARENA
#define LSS_SOLVE_ARENA 160
#endif
constexpr int SOLVE_CTAS_PER_SM = LSS_SOLVE_CTAS;
constexpr int ARENA = LSS_SOLVE_ARENA;                 // slots per warp: sum over the 32 beams of (occluders + 1); more -> extra round
constexpr int SOLVE_LCAP = 63;             // occluders per beam handled here (needs LCAP + 1 <= ARENA)
constexpr unsigned FULL = 0xffffffffu;

struct Beam {                              // what the narrow phase needs of a beam
    double d, right, left;
    bool straddle;
};

// beam limits of simulation.py:91-101 from the float32 azimuth (already shifted into [0, 2 pi))
__device__ __forceinline__ void beam_limits(float th32, double half_div, Beam &bm)
{
    const double thd = (double)th32;
    double right = thd - half_div, left = thd + half_div;
    if (right < 0) right += LSS_TWO_PI;
    if (left < 0) left += LSS_TWO_PI;
    if (right > LSS_TWO_PI) right -= LSS_TWO_PI;
    if (left > LSS_TWO_PI) left -= LSS_TWO_PI;
    bm.right = right; bm.left = left; bm.straddle = right > left;
}

// simulation.py:345-385 for one particle: planar range strictly below the target range, centre inside the beam or
// disk crossing one of the two limit rays
__device__ __forceinline__ bool exact_hit(const ParticleRec *rp, const Beam &bm, double &rho, bool &right_hit, bool &left_hit)
{
    rho = rp->rho;
    if (!(rho < bm.d)) return false;
    const double phi = rp->phi, alpha = rp->alpha;
    bool inside = (bm.right <= phi) && (phi <= bm.left);
    if (bm.straddle) inside = inside || ((bm.right - LSS_TWO_PI <= phi) && (phi <= bm.left)) ||
                              ((bm.right <= phi) && (phi <= bm.left + LSS_TWO_PI));
    right_hit = within(bm.right - phi, alpha);
    left_hit = within(bm.left - phi, alpha);
    return inside || right_hit || left_hit;
}

// ---------------------------------------------------------------------------------------------------------------------
// scan: all beams
// ---------------------------------------------------------------------------------------------------------------------
// Two phases per warp (32 consecutive rows), because a thread-per-beam loop that tests a candidate exactly as soon as it
// finds one pays the latency of that dependent record load in EVERY iteration in which any lane of the warp has a
// candidate (measured: the exact tests ran at 5 of 32 lanes and dominated the kernel):
//   A  each lane walks its beam's bucket prefix with the float32 broad phase only -- a streaming read of 8-byte
//      entries, four loads in flight -- and notes the particle indices of the survivors (shared memory, SURV_CAP per lane);
//   B  the survivors of all 32 beams are tested exactly by ALL lanes, one survivor per lane and round (owner by a
//      shuffle binary search), so the record loads of the whole warp are in flight together; hits are flagged in a
//      per-beam bit mask.
// Lanes with more than SURV_CAP survivors (extreme densities) fall back to the serial walk.
constexpr int SURV_CAP = 20;

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    const long long b = __double_as_longlong(v);
    const int lo = __shfl_sync(FULL, (int)(unsigned)b, src), hi = __shfl_sync(FULL, (int)(b >> 32), src);
    return __longlong_as_double(((long long)hi << 32) | (unsigned)lo);
}

#ifndef LSS_CLASS_BY_RANGE
#define LSS_CLASS_BY_RANGE 0
#endif
#ifndef LSS_SCAN_CTAS
#define LSS_SCAN_CTAS 10
#endif
__global__ void __launch_bounds__(SNOW_TPB, LSS_SCAN_CTAS) k_scan(DevArgs a)
{
    __shared__ float s_rows[SNOW_WARPS][32 * 5];                   // per-warp coalesced staging of 32 rows (in and out)
    __shared__ int s_idx[SNOW_WARPS][32][SURV_CAP];                // plane-local particle index of each lane's survivors
    __shared__ unsigned s_hit[SNOW_WARPS][32];                     // bit r: survivor r of this lane is a hit
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y, blk0 = blockIdx.x * SNOW_TPB, i = blk0 + threadIdx.x;
    const int64_t beg = a.cloud_off[b];
    const int n = (int)(a.cloud_off[b + 1] - beg);
    if (blk0 >= n) return;
    const bool active = i < n;
    const int w0 = blk0 + 32 * wid;                                 // first row of this warp
    const int nf_w = max(0, min(32, n - w0)) * 5;                   // floats of this warp's rows
    if (nf_w <= 0) return;                                          // (whole warps only; no block-wide barrier below)
    float px = 0, py = 0, pz = 0, pint = 0, pch = 0;
    {   // coalesced load of the warp's 32 rows (160 floats)
        const float *src = a.pts + (beg + w0) * 5;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int f = q * 32 + lane;
            if (f < nf_w) s_rows[wid][f] = __ldcs(src + f);         // streamed once: do not displace the table index in L2
        }
        __syncwarp();
        if (active) {
            const float *row = &s_rows[wid][5 * lane];
            px = row[0]; py = row[1]; pz = row[2]; pint = row[3]; pch = row[4];
        }
    }
    if (a.win_stage) {      // by-product for the pre-pass: this warp's mounting-window points, compacted (planes.py:21-27)
        const bool in = active && lss_in_window(px, py, pz);
        const unsigned m = __ballot_sync(FULL, in);
        if (in) {
            float *o = a.win_stage + (beg + w0 + __popc(m & ((1u << lane) - 1u))) * 3;
            o[0] = px; o[1] = py; o[2] = pz;
        }
        if (lane == 0) a.win_tile_cnt[lss_window_tile0(beg, b) + w0 / 32] = __popc(m);
    }
    // np.linalg.norm([x, y, z], axis=0) in float32: sqrt((x*x + y*y) + z*z), no FMA   (simulation.py:89)
    const float d32 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    const int ch = channel_bin(pch);
    float out_l = pch;
    int e0 = 0, e1 = 0, ns = 0;
    long long pbase = 0;                                            // first particle of the beam's plane
    bool slow = false;
    float th32 = 0.0f, th_rel = 0.0f;
    Beam bm;
    bm.d = (double)d32; bm.right = bm.left = 0.0; bm.straddle = false;
    if (active && ch < LSS_N_CHANNELS) {
        out_l = 0.0f;
        th32 = a.theta ? a.theta[beg + i] : azimuth32(py, px);
        if (th32 < 0.0f) th32 = __fadd_rn(th32, 6.2831855f);
        beam_limits(th32, a.half_div, bm);
        const double thd = (double)th32;
        const int plane = a.order[b * LSS_N_CHANNELS + ch];
        if (plane >= 0 && plane < a.n_planes && thd == thd) {
            const double thm = thd >= LSS_TWO_PI ? thd - LSS_TWO_PI : thd;
            int bk = (int)(thm * a.inv_w);
            bk = bk < 0 ? 0 : (bk >= a.n_buckets ? a.n_buckets - 1 : bk);
            th_rel = (float)(thm - (bk + 0.5) * a.w);
            const int32_t *bs = a.bucket_start + (int64_t)plane * (a.n_buckets + 1) + bk;
            e0 = bs[0];
            e1 = bs[1];
            pbase = a.plane_off[plane];
            // ---- phase A: broad phase over the prefix of entries nearer than the target, four loads in flight --------------
            int *pos = s_idx[wid][lane];
            bool stop = false;
#pragma unroll 1
            for (int e = e0; e < e1 && !stop; e += 4) {
                BroadEntry raw[4];
#pragma unroll
                for (int q = 0; q < 4; q++) raw[q] = __ldg(&a.entries[min(e + q, e1 - 1)]);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (stop || e + q >= e1) continue;
                    const EntryView en = lss_decode(raw[q], a.zbase);
                    if (!(en.x < d32)) { stop = true; continue; }               // sorted by range: nothing nearer follows
                    if (!(fabsf(en.y - th_rel) <= en.z)) continue;               // float32 broad phase (conservative)
                    if (ns < SURV_CAP) pos[ns] = en.idx;
                    else slow = true;
                    ns++;
                }
            }
        }
    }
    // ---- phase B: exact tests of the survivors of the whole warp, one per lane and round --------------------------------------
    s_hit[wid][lane] = 0u;
    const int mine = slow ? 0 : ns;
    int incl = mine;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, sft);
        if (lane >= sft) incl += t;
    }
    const int off = incl - mine;
    const int total = __shfl_sync(FULL, incl, 31);
    __syncwarp();
#pragma unroll 1
    for (int s0 = 0; s0 < total; s0 += 32) {
        const int s = s0 + lane;
        int j = 0;                                              // owner: the last lane whose offset is <= s
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const int c = j + step;
            const int oc = __shfl_sync(FULL, off, c & 31);
            if (c < 32 && oc <= s) j = c;
        }
        const int r = s - __shfl_sync(FULL, off, j);
        const long long pbj = __shfl_sync(FULL, pbase, j);
        Beam bj;
        bj.d = shfl_f64(bm.d, j);
        bj.right = shfl_f64(bm.right, j);
        bj.left = shfl_f64(bm.left, j);
        bj.straddle = bj.right > bj.left;
        if (s < total) {
            double rho;
            bool rh, lh;
            if (exact_hit(a.rec + pbj + s_idx[wid][j][r], bj, rho, rh, lh)) atomicOr(&s_hit[wid][j], 1u << r);
        }
    }
    __syncwarp();
    unsigned hits = s_hit[wid][lane];
    int L = __popc(hits);
    if (slow) {                                                 // serial fallback: count the hits of the whole prefix
        L = 0;
#pragma unroll 1
        for (int e = e0; e < e1; e++) {
            const EntryView en = lss_decode(__ldg(&a.entries[e]), a.zbase);
            if (!(en.x < d32)) break;
            if (!(fabsf(en.y - th_rel) <= en.z)) continue;
            double rho;
            bool rh, lh;
            if (exact_hit(a.rec + pbase + en.idx, bm, rho, rh, lh)) L++;
        }
    }
    // ---- beams with occluders: warp-aggregated push to the solve list, hit positions to the position array -------------------
    {
        const bool push = L > 0;
        const unsigned pm = __ballot_sync(FULL, push);
        int lincl = push ? L : 0;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const int t = __shfl_up_sync(FULL, lincl, sft);
            if (lane >= sft) lincl += t;
        }
        const int ltotal = __shfl_sync(FULL, lincl, 31);
        if (pm) {
            int base = 0, hbase = 0;
            const int leader = __ffs(pm) - 1;
            if (lane == leader) {
                base = atomicAdd(a.hdr, __popc(pm));
                hbase = atomicAdd(a.hdr + 3, ltotal);
            }
            base = __shfl_sync(FULL, base, leader);
            hbase = __shfl_sync(FULL, hbase, leader);
            if (push) {
                // work class = number of occluders (then far / near target): what a beam costs the solve kernel -- claiming,
                // pulses, the sweep over the window ends -- is per-beam serial work proportional to it, and a warp runs as long as
                // its slowest lane, so the 32 beams of a tile should have the same count.  The costliest class comes first so
                // that the kernel's tail is cheap tiles.  (Round 1 sorted by target range: the cost was the window samples then.)
#if LSS_CLASS_BY_RANGE
                const int cls = LIST_CLASSES - 1 - min(LIST_CLASSES - 1, (int)(d32 * (LIST_CLASSES / 100.0f)));
#else
                const int cls = LIST_CLASSES - 1 - min(LIST_CLASSES - 1, 2 * min(L, 63) + (d32 > 40.0f ? 1 : 0));
#endif
                const int slot = base + __popc(pm & ((1u << lane) - 1u));
                const int hoff = hbase + lincl - L;
                const bool fits = hoff + L <= a.hit_cap;        // position array full: the beam goes to the overflow kernel
                if (slot < a.items_cap) {
                    SolveItem it;
                    it.key = ((unsigned long long)cls << 48) | ((unsigned long long)b << 32) | (unsigned)i;
                    it.hit_off = hoff;
                    it.L = fits ? L : 0x7fff;
                    it.th32 = th32;
                    it.pad0 = 0;
                    it.pad1 = 0;
                    a.items_out[slot] = it;
                }
                if (fits) {
                    int *hp = a.hit_idx + hoff;                 // particle indices of the hits, in prefix (~ range) order
                    if (!slow) {
                        const int *pos = s_idx[wid][lane];
#pragma unroll 1
                        for (int k = 0; hits; hits &= hits - 1) hp[k++] = (int)pbase + pos[__ffs(hits) - 1];
                    } else {
                        int k = 0;
#pragma unroll 1
                        for (int e = e0; e < e1 && k < L; e++) {
                            const EntryView en = lss_decode(__ldg(&a.entries[e]), a.zbase);
                            if (!(en.x < d32)) break;
                            if (!(fabsf(en.y - th_rel) <= en.z)) continue;
                            double rho;
                            bool rh, lh;
                            if (exact_hit(a.rec + pbase + en.idx, bm, rho, rh, lh)) hp[k++] = (int)pbase + en.idx;
                        }
                    }
                }
                const unsigned cm = __match_any_sync(pm, cls);
                if (lane == __ffs(cm) - 1) atomicAdd(a.hdr + LIST_CLASSES + cls, __popc(cm));
            }
        }
    }
    // ---- rows back through shared memory (coalesced store); the listed beams' rows are rewritten by the solve kernel ------
    __syncwarp();
    if (active) {
        float *row = &s_rows[wid][5 * lane];
        row[3] = rintf(pint);                                       // np.round of the intensity column (simulation.py:516)
        row[4] = out_l;
        if (a.nocc) a.nocc[beg + i] = 0;
    }
    __syncwarp();
    {
        float *dst = a.aug + (beg + w0) * 5;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int f = q * 32 + lane;
            if (f < nf_w) dst[f] = s_rows[wid][f];                  // read back by k_keep / k_scatter from L2
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// solve: the listed beams
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_win(int ks, int ke, int k0)
{
    return (unsigned long long)(unsigned)ks | ((unsigned long long)(unsigned)ke << 11) | ((unsigned long long)(unsigned)k0 << 22);
}

// sin / cos of pi * r / (c tau) to ~1 ulp of the ANGLE: quotient in two pieces (the float64 quotient alone is off by up
// to 4e-15 in an argument of ~40), sincospi of the leading piece, first-order correction for the rest
__device__ __forceinline__ void pulse_phase(double r, double &sb, double &cb)
{
    const double ctau = 299792458.0 * 1e-8, inv_ctau = 1.0 / (299792458.0 * 1e-8);
    const double bh = r * inv_ctau;
    const double bl = fma(-bh, ctau, r) * inv_ctau;
    double s, c;
    sincospi(bh, &s, &c);
    const double corr = LSS_PI * bl;
    sb = fma(corr, c, s);
    cb = fma(-corr, s, c);
}

__global__ void __launch_bounds__(SOLVE_TPB, SOLVE_CTAS_PER_SM) k_solve(DevArgs a, int *tile_cursor)
{
    __shared__ double s_arena[SOLVE_WARPS][4][ARENA];
    __shared__ unsigned long long s_piece[SOLVE_WARPS][2 * ARENA];  // piece descriptors, then the pieces' maxima
    __shared__ int s_piece_k[SOLVE_WARPS][2 * ARENA];               // ... and where they are
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long *PD = s_piece[wid];
    int *PK = s_piece_k[wid];
    // slot arrays of this warp.  Phase 1 (hits): A0 = a1, A1 = a2, A2 = planar range.  Claiming: A0 / A1 prefix = union
    // list, A2 / A3 prefix = (range, ratio) of the claiming particles.  Phase 2 (pulses): A0 = amplitude, A1 = sin phase,
    // A3 = cos phase, A2 = packed (first sample, end sample, sample nearest to the peak).
    double *A0 = s_arena[wid][0], *A1 = s_arena[wid][1], *A2 = s_arena[wid][2], *A3 = s_arena[wid][3];
    unsigned long long *W = reinterpret_cast<unsigned long long *>(A2);

    const int cnt = min(*a.count_in, a.cap_in);
    const int n_tiles = (cnt + 31) >> 5;
    const double ctau = 299792458.0 * 1e-8;
    const double inv_step = (double)(LSS_M_EXT - 1) / (120 + ctau);

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(tile_cursor, 1);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= n_tiles) break;
        const int slot = tile * 32 + lane;
        const bool active = slot < cnt;
        SolveItem it;
        it.key = 0ull; it.hit_off = 0; it.L = 0; it.th32 = 0.0f; it.pad0 = 0; it.pad1 = 0;
        if (active) it = a.items_in[slot];
        const int b = (int)((it.key >> 32) & 0xffffu);
        const int i = (int)(it.key & 0xffffffffu);
        const int64_t beg = a.cloud_off[b];
        float px = 0, py = 0, pz = 0, pint = 0, pch = 0;
        if (active) {
            const float *row = a.pts + (beg + i) * 5;
            px = row[0]; py = row[1]; pz = row[2]; pint = row[3]; pch = row[4];
        }
        // np.linalg.norm([x, y, z], axis=0) in float32: sqrt((x*x + y*y) + z*z), no FMA   (simulation.py:89)
        const float d32 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
        const int ch = channel_bin(pch);

        float out_x = px, out_y = py, out_z = pz, out_i = pint, out_l = active ? 0.0f : pch;
        long long att_new_i = -1;
        int n_claim = 0;

        // what the scan kernel found on this beam's bucket prefix: L hits, their particle indices in hit_idx[hit_off ..]
        const int L = active ? it.L : 0;
        Beam bm;
        bm.d = (double)d32;
        beam_limits(it.th32, a.half_div, bm);

        const bool deferred = L > SOLVE_LCAP;
        if (deferred) {
            const int s2 = atomicAdd(a.count_out, 1);
            if (s2 < a.cap_out) a.list_out[s2] = ((unsigned long long)b << 32) | (unsigned)i;
            else raise_status(a.status, LSS_ERR_OCCLUDER_OVERFLOW);
        }
        const int need = (L > 0 && !deferred) ? L + 1 : 0;

        // ---- rounds: as many beams of the tile as fit into the arena (normally all of them) ---------------------------
        unsigned remaining = __ballot_sync(FULL, need > 0);
        while (remaining) {
            const bool rem = (remaining >> lane) & 1u;
            const int mine = rem ? need : 0;
            int incl = mine;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, s);
                if (lane >= s) incl += t;
            }
            const bool in_round = rem && incl <= ARENA;
            remaining &= ~__ballot_sync(FULL, in_round);
            const int off = incl - mine;                            // exclusive offsets: non-decreasing over the lanes
            const int total = (int)__reduce_max_sync(FULL, in_round ? (unsigned)incl : 0u);
            int n_pulses = 0;
            double amax = 0.0;                                      // largest pulse amplitude of this beam (prunes the pieces)
            double best = 0.0;
            int kbest = 0;

            // ---- cooperative fill: arena slot s <- the r-th hit of its owner beam, (a1, a2, range) --------------------
#pragma unroll 1
            for (int s0 = 0; s0 < total; s0 += 32) {
                const int s = s0 + lane;
                int j = 0;                                          // owner: the last lane whose offset is <= s
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int c = j + step;
                    const int oc = __shfl_sync(FULL, off, c & 31);
                    if (c < 32 && oc <= s) j = c;
                }
                const int r = s - __shfl_sync(FULL, off, j);
                const int Lj = __shfl_sync(FULL, L, j);
                const int hoj = __shfl_sync(FULL, it.hit_off, j);
                const int inr = __shfl_sync(FULL, (int)in_round, j);
                Beam bj;
                bj.d = shfl_f64(bm.d, j);
                bj.right = shfl_f64(bm.right, j);
                bj.left = shfl_f64(bm.left, j);
                bj.straddle = bj.right > bj.left;
                if (s < total && inr && r < Lj) {                   // (slot L of a beam is its hard target: filled later)
                    const int pi = a.hit_idx[hoj + r];
                    const ParticleTan tn = a.tan[pi];               // (both records requested before either is used)
                    double rho;
                    bool rh, lh;
                    exact_hit(a.rec + pi, bj, rho, rh, lh);
                    A0[s] = rh ? bj.right : tn.t_right;             // geometry.py:26-27: a limit ray the disk crosses clips
                    A1[s] = lh ? bj.left : tn.t_left;
                    A2[s] = rho;
                }
            }
            __syncwarp();

            if (in_round) {
                // order by range (np.argsort, simulation.py:416); the prefix is sorted by the float32 range already, so this
                // insertion sort almost never moves anything; equal ranges keep their prefix order
#pragma unroll 1
                for (int p = 1; p < L; p++) {
                    const double rho = A2[off + p];
                    if (!(A2[off + p - 1] > rho)) continue;
                    const double a1 = A0[off + p], a2 = A1[off + p];
                    int q = p - 1;
#pragma unroll 1
                    while (q >= 0 && A2[off + q] > rho) { A0[off + q + 1] = A0[off + q]; A1[off + q + 1] = A1[off + q]; A2[off + q + 1] = A2[off + q]; q--; }
                    A0[off + q + 1] = a1; A1[off + q + 1] = a2; A2[off + q + 1] = rho;
                }

                // ---- compute_occlusion_dict (simulation.py:231-295) ------------------------------------------------------
                // Union-list formulation (see snowfall.cu): a particle claims |[a1,a2]| - |[a1,a2] n union of the claims so
                // far| and is dropped iff its interval is contained in one union interval (or is empty); the hard target
                // gets what is left between the smallest and the largest end point -- seam quirk included.
                double rb = bm.right;
                if (bm.straddle) rb = bm.right - LSS_TWO_PI;
                int nu = 0, P = 0;
                double ep_min = fmin(rb, bm.left), ep_max = fmax(rb, bm.left), claimed_total = 0.0;
#pragma unroll 1
                for (int j = 0; j < L; j++) {
                    double lo = A0[off + j];
                    const double hi = A1[off + j], rho = A2[off + j];
                    if (bm.straddle && lo > hi) lo -= LSS_TWO_PI;              // simulation.py:260-263
                    ep_min = fmin(ep_min, fmin(lo, hi));
                    ep_max = fmax(ep_max, fmax(lo, hi));
                    if (!(lo < hi)) continue;
                    // one pass over the union list: is [lo, hi] inside one of its intervals (then nothing is claimed and the list
                    // stays as it is -- no earlier interval can have overlapped, so nothing has been moved yet), how much of it
                    // is covered, and the merged list (overlapping / touching intervals absorbed into [nlo, nhi], the others
                    // compacted in place)
                    bool contained = false;
                    double cov = 0.0, nlo = lo, nhi = hi;
                    int w = 0;
#pragma unroll 1
                    for (int u = 0; u < nu; u++) {
                        const double ul = A0[off + u], uh = A1[off + u];
                        if ((ul <= lo) && (hi <= uh)) { contained = true; break; }
                        if (ul <= hi && uh >= lo) {
                            const double ov = fmin(hi, uh) - fmax(lo, ul);
                            if (ov > 0.0) cov += ov;
                            nlo = fmin(nlo, ul);
                            nhi = fmax(nhi, uh);
                        } else {
                            A0[off + w] = ul; A1[off + w] = uh; w++;
                        }
                    }
                    if (contained) continue;
                    const double claimed = (hi - lo) - cov;
                    claimed_total += claimed;
                    A0[off + w] = nlo; A1[off + w] = nhi;       // w <= nu <= P <= j: only slots of processed hits
                    nu = w + 1;
                    double ratio = claimed / a.div_rad;
                    ratio = ratio < 0 ? 0 : (ratio > 1 ? 1 : ratio);
                    A2[off + P] = rho;
                    A3[off + P] = ratio;
                    P++;
                }
                n_claim = P;
                double ratio_hard = ((ep_max - ep_min) - claimed_total) / a.div_rad;
                ratio_hard = ratio_hard < 0 ? 0 : (ratio_hard > 1 ? 1 : ratio_hard);

                if (P > 0) {
                    // ---- pulses of the waveform (simulation.py:137-149) --------------------------------------------------
                    const double beta_0 = 1 * 1e-06 / LSS_PI;
                    const double i_orig = 0.9 * a.sensor->max_intensity[ch];
                    const double A = (i_orig / beta_0) * beta_0;        // CA_P0 * beta_0 (quirk: every pulse uses it)
                    bool bad = false;
#pragma unroll 1
                    for (int j = 0; j < P; j++) {
                        const double r = A2[off + j], ratio = A3[off + j];
                        const int ks = (int)ceil(r * 10);
                        const int ke = (int)(floor((r + ctau) * 10) + 1);
                        bad |= (ke > LSS_M_EXT) || (ks < 0);
                        double sb, cb;
                        pulse_phase(r, sb, cb);
                        const double amp = (A * ratio * xsi64(r)) / (r * r);
                        amax = fmax(amax, amp);
                        A0[off + j] = amp;
                        A1[off + j] = sb;
                        A3[off + j] = cb;
                        W[off + j] = pack_win(ks, ke, (int)rint((r + ctau / 2) * inv_step));
                    }
                    {   // hard target: r_j is float32 => float32 index arithmetic and r^2
                        const int ks = (int)ceilf(__fmul_rn(d32, 10.0f));
                        const int ke = (int)(floorf(__fmul_rn(__fadd_rn(d32, (float)ctau), 10.0f)) + 1.0f);
                        bad |= (ke > LSS_M_EXT) || (ks < 0);
                        double sb, cb;
                        pulse_phase(bm.d, sb, cb);
                        const double amp = (A * ratio_hard * xsi32(d32)) / (double)__fmul_rn(d32, d32);
                        amax = fmax(amax, amp);
                        A0[off + P] = amp;
                        A1[off + P] = sb;
                        A3[off + P] = cb;
                        W[off + P] = pack_win(ks, ke, (int)rint((bm.d + ctau / 2) * inv_step));
                    }
                    if (bad) raise_status(a.status, LSS_ERR_RANGE_INDEX);
                    else n_pulses = P + 1;
                }
            }

            // ---- argmax of the summed waveform (simulation.py:148-153) -----------------------------------------------------
            // Only samples inside some pulse window are non-zero.  Pulse ranges ascend, so first and end samples of the
            // windows ascend too and a sweep over them cuts the sample axis into PIECES on which the set of active pulses is
            // a fixed index range [qa, qb].  On a piece the waveform is ONE sinusoid of period c tau (29.96 samples; a piece
            // is never longer than a window, 31 samples):
            //     sum_q A_q sin^2(pi (R - r_q) / (c tau)) = C - |Z| / 2 * cos(2 pi R / (c tau) - arg Z),   Z = sum_q A_q e^(2 pi i r_q / (c tau))
            // so its maximum over the piece's samples is at the sample nearest to the analytic peak R* closest to the piece's
            // middle -- or, if that lies outside, at the piece's end nearest to it.  The three samples around R*, clipped to
            // the piece, are evaluated exactly (float64, pulses summed in dict order like the reference's i[k] +=; the first
            // maximum wins, np.argmax); R* only SELECTS them (float32 is ample: the grid is 0.1 m, its rounding to 0.01 m moves
            // the nearest sample by at most one).  If the pulses cancel (|Z| tiny: every sample of the piece has the same value
            // up to rounding) the whole piece is evaluated.  A single pulse peaks at its stored sample k0.
            //   1. every lane sweeps its own pulses and writes piece descriptors (cheap, divergent);
            //   2. the pieces of ALL beams of the round are handled one per lane (owner by shuffle binary search): uniform
            //      code, full warps -- a lane-local version ran at 6 of 32 lanes and the round-2a version summed whole
            //      windows (47 % of the kernel's instructions);
            //   3. every lane takes the first maximum over its own pieces.
            int np_mine = 0;
            if (n_pulses > 0) {
                // Pruning: the largest pulse alone contributes A_max sin^2 >= 0.9966 A_max at its stored peak sample (at most
                // half a grid step + the grid's 0.005 m rounding away from r + c tau / 2), every term of the sum is >= 0 and
                // float64 addition of non-negative terms is monotone, so the waveform's maximum is >= 0.99 A_max.  A piece whose
                // active amplitudes sum to less than that (sin^2 <= 1) cannot hold the argmax and is not even written down.
                const double lb = 0.99 * amax;
                double asum = 0.0;                                  // running sum of the active amplitudes (drift << the margin)
                int qa = 0, qb = -1, nxt = 0;
                int ks_n = (int)(W[off] & 2047u);                   // first sample of the next pulse to start (cached)
                int ke_a = 0;                                       // end sample of the oldest active pulse (cached)
                int k = ks_n;
                unsigned long long *pd = PD + 2 * off;
#pragma unroll 1
                for (;;) {
#pragma unroll 1
                    while (ks_n <= k) {
                        qb = nxt++;
                        asum += A0[off + qb];
                        ks_n = nxt < n_pulses ? (int)(W[off + nxt] & 2047u) : 4096;
                    }
#pragma unroll 1
                    while (qa <= qb) {
                        ke_a = (int)((W[off + qa] >> 11) & 2047u);
                        if (ke_a > k) break;
                        asum -= A0[off + qa];
                        qa++;
                    }
                    if (qa > qb) {
                        asum = 0.0;
                        if (nxt >= n_pulses) break;
                        k = ks_n;
                        continue;
                    }
                    const int pend = min(ke_a, ks_n);               // the oldest active pulse ends first, or the next one starts
                    // piece [k, pend), active pulses off + qa .. off + qb (at most 2 n_pulses - 1 pieces: they fit 2 (L + 1) slots)
                    if (asum * 1.0001 >= lb)
                        pd[np_mine++] = (unsigned long long)(unsigned)k | ((unsigned long long)(unsigned)pend << 11) |
                                        ((unsigned long long)(unsigned)(off + qa) << 22) | ((unsigned long long)(unsigned)(off + qb) << 32);
                    k = pend;
                }
            }
            int pincl = np_mine;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                const int t = __shfl_up_sync(FULL, pincl, sft);
                if (lane >= sft) pincl += t;
            }
            const int pex = pincl - np_mine;
            const int ptotal = __shfl_sync(FULL, pincl, 31);
            __syncwarp();
            {
                const float step = (float)((120 + ctau) / (double)(LSS_M_EXT - 1));
                const float ctau_f = (float)ctau;
#pragma unroll 1
                for (int f0 = 0; f0 < ptotal; f0 += 32) {
                    const int f = f0 + lane;
                    int j = 0;                                      // owner: the last lane whose piece offset is <= f
#pragma unroll
                    for (int stp = 16; stp; stp >>= 1) {
                        const int c = j + stp;
                        const int oc = __shfl_sync(FULL, pex, c & 31);
                        if (c < 32 && oc <= f) j = c;
                    }
                    const int slot = 2 * __shfl_sync(FULL, off, j) + (f - __shfl_sync(FULL, pex, j));
                    if (f < ptotal) {
                        const unsigned long long d = PD[slot];
                        const int lo = (int)(d & 2047u), hi = (int)((d >> 11) & 2047u);
                        const int qa = (int)((d >> 22) & 1023u), qb = (int)((d >> 32) & 1023u);
                        double pbest = 0.0;
                        int pk = lo;
                        auto eval = [&](int c) {
                            const double2 t = __ldg(&a.wtab[c]);
                            double v = 0.0;
#pragma unroll 1
                            for (int q = qa; q <= qb; q++) {
                                const double sn = t.x * A3[q] - t.y * A1[q];
                                v += A0[q] * (sn * sn);
                            }
                            if (v > pbest) { pbest = v; pk = c; }   // ascending c: the first maximum stays
                        };
                        int kc;
                        bool all = false;
                        if (qa == qb) {
                            kc = (int)((W[qa] >> 22) & 2047u);
                        } else {
                            float zx = 0.0f, zy = 0.0f, asum = 0.0f;
#pragma unroll 1
                            for (int q = qa; q <= qb; q++) {
                                const float A = (float)A0[q], sn = (float)A1[q], cs = (float)A3[q];
                                zx += A * (cs * cs - sn * sn);
                                zy += A * (2.0f * sn * cs);
                                asum += A;
                            }
                            all = zx * zx + zy * zy < 1e-6f * asum * asum;
                            const float r0 = (atan2f(zy, zx) + 3.14159265f) * (ctau_f * 0.15915494f);
                            const float rc = 0.5f * (float)(lo + hi - 1) * step;
                            kc = (int)rintf((r0 + rintf((rc - r0) / ctau_f) * ctau_f) / step);
                        }
                        const int c_lo = all ? lo : max(lo, min(hi - 1, kc - 1));
                        const int c_hi = all ? hi - 1 : min(hi - 1, max(lo, kc + 1));
#pragma unroll 1
                        for (int c = c_lo; c <= c_hi; c++) eval(c);
                        PD[slot] = (unsigned long long)__double_as_longlong(pbest);
                        PK[slot] = pk;
                    }
                }
            }
            __syncwarp();
#pragma unroll 1
            for (int r = 0; r < np_mine; r++) {                     // pieces ascend in sample index: strict > keeps the first maximum
                const double v = __longlong_as_double((long long)PD[2 * off + r]);
                if (v > best) { best = v; kbest = PK[2 * off + r]; }
            }
            __syncwarp();

            if (n_pulses > 0) {
                // ---- new range / intensity / label (simulation.py:151-188) -------------------------------------------------
                const double max_i = a.sensor->max_intensity[ch];
                const double min_i = a.sensor->min_intensity[ch];
                const double d_max = ((double)kbest / 10) - (ctau / 2);
                const double q1 = 1 - d_max / 120;
                double i_max = best + max_i * a.sensor->focal_slope[ch] * fabs(a.sensor->focal_offset[ch] - q1 * q1);
                i_max = i_max < min_i ? min_i : (i_max > max_i ? max_i : i_max);
                const long long new_i = (long long)i_max;       // int(): truncation
                if (fabs(d_max - bm.d) < 2 * (1.0 / 10)) {
                    out_l = 1.0f;
                    att_new_i = new_i;                          // intensity_diff_sum += i_orig - new_i   (simulation.py:170)
                } else {
                    out_l = 2.0f;
                    const double sc = d_max / bm.d;
                    out_x = (float)((double)px * sc);
                    out_y = (float)((double)py * sc);
                    out_z = (float)((double)pz * sc);
                }
                if (new_i < 0) raise_status(a.status, LSS_ERR_NEGATIVE_INTENSITY);
                double ci = (double)new_i;
                ci = ci < min_i ? min_i : (ci > max_i ? max_i : ci);
                out_i = (float)ci;
            }
            __syncwarp();
        }

        // ---- np.round of the intensity column (simulation.py:516), store, label-1 statistics (simulation.py:170) ----------
        const bool counted = active && !deferred;          // a deferred beam is written by the overflow kernel
        if (counted) {
            float *row = a.aug + (beg + i) * 5;
            row[0] = out_x; row[1] = out_y; row[2] = out_z; row[3] = rintf(out_i); row[4] = out_l;
            if (a.nocc) a.nocc[beg + i] = n_claim;
        }
        {
            const bool on = counted && att_new_i >= 0;
            const int key = b * LSS_N_CHANNELS + (ch < LSS_N_CHANNELS ? ch : 0);
            const unsigned mk = __match_any_sync(FULL, on ? key : -1);
            if (on && lane == __ffs(mk) - 1) atomicAdd(a.att_cnt + key, (unsigned)__popc(mk));
            const unsigned mb = __match_any_sync(FULL, on ? b : -1);
            const unsigned sum = __reduce_add_sync(mb, on ? (unsigned)att_new_i : 0u);
            if (on && lane == __ffs(mb) - 1) atomicAdd(&a.att_sum[b], (unsigned long long)sum);
        }
    }
}

}  // namespace

namespace {
__global__ void k_debug_azimuth(const float *y, const float *x, float *out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = azimuth32(y[i], x[i]);
}
}  // namespace

extern "C" lss_status lss_debug_azimuth(lss_engine *e, const float *d_y, const float *d_x, int64_t n, float *d_out, void *stream)
{
    if (!e || !d_y || !d_x || !d_out || n < 0) return LSS_ERR_INVALID_ARG;
    if (n == 0) return LSS_OK;
    DeviceGuard g(e->device);
    k_debug_azimuth<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_y, d_x, d_out, n);
    e->launches++;
    LSS_CUDA_CHECK(e, cudaGetLastError());
    return LSS_OK;
}

void lss_launch_scan(const DevArgs &a, int64_t max_rows, int n_clouds, cudaStream_t stream)
{
    const dim3 grid((unsigned)((max_rows + SNOW_TPB - 1) / SNOW_TPB), (unsigned)n_clouds);
    k_scan<<<grid, SNOW_TPB, 0, stream>>>(a);
}

void lss_launch_solve(const DevArgs &a, int *tile_cursor, int n_sm, cudaStream_t stream)
{
    k_solve<<<(unsigned)(n_sm * SOLVE_CTAS_PER_SM), SOLVE_TPB, 0, stream>>>(a, tile_cursor);
}
