// kernels_msm.cu -- Pippenger bucket MSM for variable bases (bls.LinCombG1, bls/bls_kilic.go:132-150):
// the kernels around the per-thread bodies of msm.cuh.  Integer-pipe bound like every G1 kernel; the design
// goal is parallel slack: n = 4096 gives 8192 GLV terms x 13 windows, spread over (bucket, slice) threads so
// that no thread adds more than a handful of points, with warp-shuffle trees for every partial-sum reduction.
#include "kernels.h"
#include "msm.cuh"
#include "quad.cuh"
#include <cooperative_groups.h>

namespace b200 {

static inline unsigned grid_for(size_t total, unsigned block) { return (unsigned)((total + block - 1) / block); }

// ---- warp-shuffle helpers: a Jacobian point is 36 words ---------------------------------------------------
__device__ __forceinline__ G1J g1_shfl_xor(const G1J& p, unsigned lane_mask) {
    G1J r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        r.x.l[i] = __shfl_xor_sync(0xffffffffu, p.x.l[i], lane_mask);
        r.y.l[i] = __shfl_xor_sync(0xffffffffu, p.y.l[i], lane_mask);
        r.z.l[i] = __shfl_xor_sync(0xffffffffu, p.z.l[i], lane_mask);
    }
    return r;
}
// butterfly sum over groups of `width` adjacent lanes (power of two): every lane ends with the group's sum
__device__ __forceinline__ void g1_warp_sum(G1J& acc, unsigned width) {
    for (unsigned off = width >> 1; off >= 1; off >>= 1) {
        G1J other = g1_shfl_xor(acc, off);
        g1_add_ni(&acc, &acc, &other);
    }
}

// ---- step 1 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_msm_recode(MsmPlan p, const G1J* __restrict__ pts, const Fr* __restrict__ k, int k_is_mont,
                                                    int16_t* digits, uint32_t* counts, Fp* bx, uint32_t* not_affine) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    Fr s = ld_vec(k + i);
    if (k_is_mont) s = fe_from_mont(s);
    G1J pt = ld_vec(pts + i);
    msm_recode_point(p, i, pt, s, digits, counts, bx, not_affine);
}

// ---- step 2: exclusive prefix of counts[w][0..B] (entry 0, the zero digit, is never counted) --------------
__global__ void __launch_bounds__(1024) k_msm_scan(MsmPlan p, const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t part[1024];
    const unsigned w = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const unsigned len = p.B + 1, per = (len + T - 1) / T;
    const uint32_t* c = counts + (size_t)w * len;
    uint32_t* o = offsets + (size_t)w * len;
    const unsigned lo = tid * per, hi = lo + per < len ? lo + per : len;
    uint32_t s = 0;
    for (unsigned j = lo; j < hi; j++) s += c[j];
    part[tid] = s;
    __syncthreads();
    for (unsigned d = 1; d < T; d <<= 1) {            // Hillis-Steele inclusive scan of the per-thread totals
        uint32_t v = tid >= d ? part[tid - d] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = tid ? part[tid - 1] : 0u;
    for (unsigned j = lo; j < hi; j++) { o[j] = run; run += c[j]; }
}

// ---- step 3 ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_msm_scatter(MsmPlan p, const int16_t* __restrict__ digits, const uint32_t* __restrict__ offsets,
                                                     uint32_t* cursors, uint32_t* sorted) {
    size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)p.W * p.T) return;
    msm_scatter_term(p, (unsigned)(g / p.T), g % p.T, digits, offsets, cursors, sorted);
}

// ---- step 4: (window, bucket, slice) units; the S slices of a bucket are adjacent -------------------------------
// QUAD = false: one thread per unit (large problems: every lane has a list of its own, throughput bound).
// QUAD = true:  one quad per unit (quad.cuh: an addition in 5 dependent products instead of 11 / 16) for problems that
//               cannot fill the chip anyway, where the depth of the chain of additions is what is waited for.
template <bool QUAD>
__global__ void __launch_bounds__(128, 4) k_msm_accumulate(MsmPlan p, const G1J* __restrict__ pts, const Fp* __restrict__ bx,
                                                           const uint32_t* __restrict__ sorted, const uint32_t* __restrict__ counts,
                                                           const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ not_affine,
                                                           G1J* __restrict__ buckets) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t u = QUAD ? tid >> 2 : tid;        // unit index; the units of a window start at a multiple of 32
    unsigned w = 0;
    while (w + 1 < p.W && u >= p.unit_off[w + 1]) w++;
    const unsigned S = p.Sw[w];
    const size_t local = u - p.unit_off[w];
    const unsigned b = (unsigned)(local / S) + 1, slice = (unsigned)(local % S);
    const bool valid = u < p.unit_off[p.W] && b <= p.Bw[w];
    G1J acc = G1J::infinity();
    uint32_t qbegin = 0, qcount = 0;
    if (valid) {
        const size_t slot = (size_t)w * (p.B + 1) + b;
        const uint32_t len = counts[slot], off = offsets[slot];
        const uint32_t per = (len + S - 1) / S;
        uint32_t begin = slice * per, end = begin + per;
        if (begin > len) begin = len;
        if (end > len) end = len;
        if (!QUAD) msm_accumulate_slice(p, pts, bx, sorted + (size_t)w * p.T, off + begin, off + end, *not_affine == 0, &acc);
        qbegin = off + begin; qcount = end - begin;
    }
    if (QUAD) {
        // warp-collective quad operations: every quad walks max(count over the warp) steps, idle ones pass active = false
        const bool affine = *not_affine == 0;
        const uint32_t* sw = sorted + (size_t)w * p.T;
        const uint32_t steps = __reduce_max_sync(0xffffffffu, qcount);
        for (uint32_t k = 0; k < steps; k++) {
            const bool act = k < qcount;
            G1J q = G1J::infinity();
            if (act) {
                const uint32_t v = sw[qbegin + k];
                const size_t t = v & 0x7fffffffu;
                const bool second = t >= p.n;
                const size_t i = second ? t - p.n : t;
                const bool neg = ((v >> 31) != 0) != second;
                q.x = second ? ld_vec(bx + i) : ld_vec(&pts[i].x);
                q.y = ld_vec(&pts[i].y);
                if (neg) q.y = fe_neg(q.y);
                if (!affine) q.z = ld_vec(&pts[i].z);
            }
            if (affine) {
                G1A qa; qa.x = q.x; qa.y = q.y;
                quad_add_mixed(&acc, &qa, act);
            } else {
                quad_add(&acc, &q, act);
            }
        }
    }
    // the slices of a bucket: adjacent lanes (thread units) or adjacent quads (quad units); whole warps get here, and a
    // warp lies inside one window, so S is uniform over it
    if (!QUAD) {
        if (S > 1) g1_warp_sum(acc, S);
    } else if (S > 1) {
        quad_group_sum(acc, S);
    }
    if (valid && slice == 0 && (!QUAD || (threadIdx.x & 3u) == 0)) st_vec(buckets + (size_t)w * p.B + (b - 1), acc);
}

// ---- step 5: thread = (window, segment of L buckets) ---------------------------------------------------------
// (a quad per segment was measured slower here: the collective small multiplication runs an addition on every bit)
__global__ void __launch_bounds__(128, 4) k_msm_segments(MsmPlan p, const G1J* __restrict__ buckets, G1J* __restrict__ segs) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned nseg = p.B / p.L;
    if (g >= (size_t)p.W * nseg) return;
    const unsigned w = (unsigned)(g / nseg), s = (unsigned)(g % nseg);
    G1J out;
    msm_reduce_segment(buckets + (size_t)w * p.B, s * p.L + 1, p.L, &out);
    st_vec(segs + g, out);
}

// ---- step 6: one CTA (64 quads) per window: sum of the window's segments ----------------------------------------
#define MSM_WINDOWS_THREADS 256
__global__ void __launch_bounds__(MSM_WINDOWS_THREADS) k_msm_windows(MsmPlan p, const G1J* __restrict__ segs, G1J* __restrict__ wsums) {
    __shared__ G1J part[MSM_WINDOWS_THREADS / 32];
    const unsigned w = blockIdx.x, tid = threadIdx.x, quad = tid >> 2, nseg = p.B / p.L;
    G1J acc = G1J::infinity();
    for (unsigned base = 0; base < nseg; base += MSM_WINDOWS_THREADS / 4) {
        const unsigned s = base + quad;
        const bool act = s < nseg;
        G1J v = act ? ld_vec(segs + (size_t)w * nseg + s) : G1J::infinity();
        quad_add(&acc, &v, act);
    }
    quad_warp_sum(acc);
    if ((tid & 31) == 0) part[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {                                  // warp 0: quad j takes the sum of warp j (8 warps, 8 quads), then a shuffle tree
        G1J v = part[tid >> 2];
        quad_warp_sum(v);
        if (tid == 0) st_vec(wsums + w, v);
    }
}
static_assert(MSM_WINDOWS_THREADS == 256, "warp 0 holds one quad per warp of the CTA");

// ---- steps 5 + 6 for windows of up to 512 buckets: one CTA per window, Q = min(B, 64) quads, L = B / Q buckets each ----------
// sum_b b Bucket[b] with b = q L + j + 1:   sum_q A_q + L sum_(k >= 1) Suf_k,   A_q = sum_j (j + 1) Bucket[q L + j + 1] (running
// sums, 2 L additions), R_q = sum_j Bucket[..], Suf_k = sum_(q >= k) R_q by a Hillis-Steele suffix scan over the quads (log2 Q
// levels through shared memory), L Suf by log2 L doublings, then one tree over the quads.  Depth: 2 L + 2 log2 Q + 1 quad
// additions + log2 L doublings (n = 4096: 15 additions, ~0.1 ms) where the segment + window kernels above take ~0.4 ms because of
// their one-lane additions and the small multiplication by the segment's base index.
#define MSM_SCAN_QUADS 64      // two warps per SM sub-partition: with 128 the four warps of a sub-partition share its multiplier and every level takes twice as long
__global__ void __launch_bounds__(4 * MSM_SCAN_QUADS) k_msm_window_scan(MsmPlan p, const G1J* __restrict__ buckets, G1J* __restrict__ wsums) {
    __shared__ G1J sh[MSM_SCAN_QUADS];
    __shared__ G1J part[MSM_SCAN_QUADS / 8];
    static_assert(MSM_SCAN_QUADS <= 64, "warp 0 holds one quad per warp of the CTA");
    const unsigned w = blockIdx.x, tid = threadIdx.x, q = tid >> 2, Q = blockDim.x >> 2, L = p.B / Q;
    const G1J* bkt = buckets + (size_t)w * p.B + (size_t)q * L;
    G1J run = ld_vec(bkt + (L - 1)), acc = run;
    for (int j = (int)L - 2; j >= 0; j--) {
        G1J v = ld_vec(bkt + j);
        quad_add(&run, &v, true);
        quad_add(&acc, &run, true);
    }
    G1J suf = run;
    for (unsigned d = 1; d < Q; d <<= 1) {
        if ((tid & 3u) == 0) sh[q] = suf;
        __syncthreads();
        const bool act = q + d < Q;
        G1J other = act ? sh[q + d] : G1J::infinity();
        __syncthreads();
        quad_add(&suf, &other, act);
    }
    for (unsigned l = L; l > 1; l >>= 1) quad_dbl(&suf, true);           // L Suf_q
    quad_add(&acc, &suf, q >= 1);                                         // quad 0 has base index 0
    quad_warp_sum(acc);
    const unsigned nwarp = blockDim.x >> 5;
    if ((tid & 31u) == 0) part[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {                                  // warp 0: quad j takes the sum of warp j (at most 8 warps), then a shuffle tree
        const unsigned j = tid >> 2;
        G1J v = j < nwarp ? part[j] : G1J::infinity();
        quad_warp_sum(v);
        if (tid == 0) st_vec(wsums + w, v);
    }
}

// The same reduction for windows of 128 .. 1024 buckets over a CLUSTER of four CTAs (4 x 32 quads = 128 quads, one quad warp
// per SM sub-partition on four SMs, so no two warps share a multiplier): the suffix scan reads the neighbouring CTAs' values
// through distributed shared memory.  n = 4096 (512 buckets, L = 4): 23 quad additions + 2 doublings deep.
#define MSM_CL_CTAS 4
#define MSM_CL_QUADS 32
__global__ void __cluster_dims__(MSM_CL_CTAS, 1, 1) __launch_bounds__(4 * MSM_CL_QUADS)
k_msm_window_scan_cluster(MsmPlan p, const G1J* __restrict__ buckets, G1J* __restrict__ wsums) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ G1J sh[MSM_CL_QUADS];
    __shared__ G1J part[MSM_CL_QUADS / 8];
    __shared__ G1J cta_total;
    const unsigned rank = cluster.block_rank(), w = blockIdx.x / MSM_CL_CTAS, tid = threadIdx.x, q = tid >> 2;
    const unsigned Q = MSM_CL_CTAS * MSM_CL_QUADS, gq = rank * MSM_CL_QUADS + q, L = p.B / Q;
    const G1J* bkt = buckets + (size_t)w * p.B + (size_t)gq * L;
    G1J run = ld_vec(bkt + (L - 1)), acc = run;
    for (int j = (int)L - 2; j >= 0; j--) {
        G1J v = ld_vec(bkt + j);
        quad_add(&run, &v, true);
        quad_add(&acc, &run, true);
    }
    G1J suf = run;
    for (unsigned d = 1; d < Q; d <<= 1) {
        if ((tid & 3u) == 0) sh[q] = suf;
        cluster.sync();
        const unsigned src = gq + d;
        const bool act = src < Q;
        G1J other = G1J::infinity();
        if (act) {
            const G1J* remote = cluster.map_shared_rank(sh, src / MSM_CL_QUADS);
            other = remote[src % MSM_CL_QUADS];
        }
        cluster.sync();
        quad_add(&suf, &other, act);
    }
    for (unsigned l = L; l > 1; l >>= 1) quad_dbl(&suf, true);
    quad_add(&acc, &suf, gq >= 1);
    quad_warp_sum(acc);
    if ((tid & 31u) == 0) part[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {
        const unsigned j = tid >> 2;
        G1J v = j < MSM_CL_QUADS / 8 ? part[j] : G1J::infinity();
        quad_warp_sum(v);
        if (tid == 0) cta_total = v;
    }
    cluster.sync();
    if (rank == 0 && tid < 32) {
        const unsigned j = tid >> 2;
        G1J v = G1J::infinity();
        if (j < MSM_CL_CTAS) v = *cluster.map_shared_rank(&cta_total, j);
        quad_warp_sum(v);
        if (tid == 0) st_vec(wsums + w, v);
    }
    cluster.sync();                                  // keep every CTA's shared memory alive until rank 0 has read it
}

// ---- step 7: one CTA, quad w doubles the sum of window w  c w  times; then the sum over the windows -------------
__global__ void __launch_bounds__(160) k_msm_horner(MsmPlan p, const G1J* __restrict__ wsums, G1J* __restrict__ out) {
    __shared__ G1J part[5];
    const unsigned tid = threadIdx.x, w = tid >> 2;
    G1J acc = G1J::infinity();
    unsigned dbls = 0;
    if (w < p.W) { acc = ld_vec(wsums + w); dbls = p.c * w; }
    const unsigned steps = __reduce_max_sync(0xffffffffu, dbls);
    for (unsigned d = 0; d < steps; d++) quad_dbl(&acc, d < dbls);
    quad_warp_sum(acc);
    const unsigned nwarp = blockDim.x >> 5;
    if ((tid & 31) == 0) part[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {                                  // quad j of warp 0 takes the sum of warp j, then a shuffle tree
        G1J v = (tid >> 2) < nwarp ? part[tid >> 2] : G1J::infinity();
        quad_warp_sum(v);
        if (tid == 0) st_vec(out, v);
    }
}

// a quad per (bucket, slice) unit only while the expanded launch stays at about one warp per SM sub-partition
static const size_t kMsmQuadUnits = (size_t)148 * 4 * 32;   // one warp per SM sub-partition

// ---- workspace layout -------------------------------------------------------------------------------------
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
struct MsmWorkspace {
    int16_t* digits; uint32_t* counts; uint32_t* offsets; uint32_t* cursors; uint32_t* flag; uint32_t* sorted;
    Fp* bx; G1J* buckets; G1J* segs; G1J* wsums;
    size_t zero_bytes;      // counts + cursors + flag are contiguous at `counts` and cleared per call
};
static size_t msm_layout(const MsmPlan& p, char* base, MsmWorkspace* ws) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align256(bytes); return base ? base + at : (char*)nullptr; };
    const size_t slots = (size_t)p.W * (p.B + 1);
    char* counts = take(slots * 4);
    char* cursors = take(slots * 4);
    char* flag = take(4);
    const size_t zero_end = o;
    char* offsets = take(slots * 4);
    char* digits = take((size_t)p.W * p.T * 2);
    char* sorted = take((size_t)p.W * p.T * 4);
    char* bx = take(p.n * sizeof(Fp));
    char* buckets = take((size_t)p.W * p.B * sizeof(G1J));
    char* segs = take((size_t)p.W * (p.B / p.L) * sizeof(G1J));
    char* wsums = take((size_t)p.W * sizeof(G1J));
    if (ws) {
        ws->counts = (uint32_t*)counts; ws->cursors = (uint32_t*)cursors; ws->flag = (uint32_t*)flag; ws->offsets = (uint32_t*)offsets;
        ws->digits = (int16_t*)digits; ws->sorted = (uint32_t*)sorted; ws->bx = (Fp*)bx; ws->buckets = (G1J*)buckets;
        ws->segs = (G1J*)segs; ws->wsums = (G1J*)wsums; ws->zero_bytes = zero_end;
    }
    return o;
}
size_t msm_workspace_bytes(size_t n) {
    if (n == 0) return 256;
    MsmPlan p = msm_plan(n);
    return msm_layout(p, nullptr, nullptr);
}

// out (one internal Jacobian point) = sum_i k[i] pts[i]
void launch_g1_msm(const G1J* pts, const Fr* k, int k_is_mont, size_t n, void* workspace, G1J* out, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_MSM, st);
    if (n == 0) { launch_g1_fill_infinity(out, 1, st); return; }
    MsmPlan p = msm_plan(n);
    MsmWorkspace ws;
    msm_layout(p, (char*)workspace, &ws);
    cudaMemsetAsync(ws.counts, 0, ws.zero_bytes, st);
    k_msm_recode<<<grid_for(n, 128), 128, 0, st>>>(p, pts, k, k_is_mont, ws.digits, ws.counts, ws.bx, ws.flag);
    k_msm_scan<<<p.W, 1024, 0, st>>>(p, ws.counts, ws.offsets);
    k_msm_scatter<<<grid_for((size_t)p.W * p.T, 256), 256, 0, st>>>(p, ws.digits, ws.offsets, ws.cursors, ws.sorted);
    const size_t units = p.unit_off[p.W];
    cudaMemsetAsync(ws.buckets, 0, (size_t)p.W * p.B * sizeof(G1J), st);   // buckets no unit owns (beyond Bw[w]) are infinity
    // a quad per unit only pays when the expanded launch is still a fraction of one wave (measured at n = 4096: 213 k
    // quad lanes take 1.16 ms where 53 k thread units take 0.24 ms -- a warp instruction costs the same with 8 points as with 32)
    bool quad_ok = units * 4 <= kMsmQuadUnits;
    for (unsigned w = 0; w < p.W; w++) quad_ok = quad_ok && p.Sw[w] <= 8;
    // 64-thread CTAs: a launch below one wave then spreads evenly over the SMs (CTAs of 128 leave some SMs with 3 and others with 2)
    if (quad_ok) k_msm_accumulate<true><<<grid_for(units * 4, 64), 64, 0, st>>>(p, pts, ws.bx, ws.sorted, ws.counts, ws.offsets, ws.flag, ws.buckets);
    else k_msm_accumulate<false><<<grid_for(units, 64), 64, 0, st>>>(p, pts, ws.bx, ws.sorted, ws.counts, ws.offsets, ws.flag, ws.buckets);
    if (p.B >= MSM_CL_CTAS * MSM_CL_QUADS && p.B <= 8 * MSM_CL_CTAS * MSM_CL_QUADS) {   // 128 .. 1024 buckets: cluster of four CTAs per window
        k_msm_window_scan_cluster<<<p.W * MSM_CL_CTAS, 4 * MSM_CL_QUADS, 0, st>>>(p, ws.buckets, ws.wsums);
        g_launch_count += 6;
    } else if (p.B <= 8 * MSM_SCAN_QUADS) {          // fewer buckets (n below ~1000): one CTA per window
        const unsigned quads = p.B < MSM_SCAN_QUADS ? p.B : MSM_SCAN_QUADS;
        k_msm_window_scan<<<p.W, 4 * quads, 0, st>>>(p, ws.buckets, ws.wsums);
        g_launch_count += 6;
    } else {
        k_msm_segments<<<grid_for((size_t)p.W * (p.B / p.L), 128), 128, 0, st>>>(p, ws.buckets, ws.segs);
        k_msm_windows<<<p.W, MSM_WINDOWS_THREADS, 0, st>>>(p, ws.segs, ws.wsums);
        g_launch_count += 7;
    }
    k_msm_horner<<<1, (unsigned)((p.W * 4 + 31) / 32 * 32), 0, st>>>(p, ws.wsums, out);
}

}  // namespace b200
