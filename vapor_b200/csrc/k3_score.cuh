// Kernel 3: dot cleaning + diagonal-distance statistics + gates + per-read score.
//
// One CTA per task (= one read of one SV/allele).  It restates, on the sparse hit lists of
// kernel 2, the reference functions
//   dis_cluster_2 / clean_dotdata_diagnal_and_anti_diagnal   vapor_vali/Simple_function.pyx:566-580, 432-448
//   dis_cluster / clean_dotdata_(anti_)diagnal_m1b            :551-564, 404-430
//   eu_dis_abs_calcu, eu_dis_dots_within_10perc               :705-708, 730-733
//   dis_to_diagnal_most_abundant_defined (+number_cluster,
//     find_longest_list, unify_list), eu_dis_dir_calcu         :582-591, 1104-1118, 788-792, 710-722
//   the three calcu_vapor_single_read_score_* gate ladders     :182-203, 241-257, 277-294
//   and the per-read combine of the L2 drivers                 :1718-1726, 1913-1915
//
// The reference clusters by sorting; here a *bitmap* of the occupied values of y-x (resp. y+x)
// in shared memory plays the sorted list: chain groups ("successive difference < 10") are runs
// of set bits separated by >= 9 clear bits, found with shifts + popc; a dot's group is the rank
// of its run (popc prefix), group sizes are counted per dot with warp-aggregated shared-memory
// atomics.  The cost of a clustering round is O(dots + range/32), the scratch 1.2 bytes per
// value of the range.  All statistics are exact integer sums; the only floating-point
// operations are the final IEEE divisions the reference also performs, so results are
// bit-identical, not merely within tolerance.
#pragma once
#include "common.cuh"

namespace vb {

#ifndef K3_THREADS_N
#define K3_THREADS_N 256
#endif
constexpr int K3_THREADS = K3_THREADS_N;

// scratch layout in 32-bit words for a value range of NB: two group sets (y-x and y+x are clustered side by
// side, so no flag has to travel through global memory between them) + one small histogram
__host__ __device__ inline int k3_words_bitmap(int nb) { return (nb + 31) / 32 + 1; }
__host__ __device__ inline int k3_words_groups(int nb) { return nb / 10 + 4; }
__host__ __device__ inline size_t k3_scratch_words(int nb) {
    return 2 * (3 * (size_t)k3_words_bitmap(nb) + (size_t)k3_words_groups(nb)) + (size_t)k3_words_groups(nb) + 8;
}

struct K3Groups {       // chain groups of one value axis
    uint32_t* present;  // bitmap of occupied values
    uint32_t* start;    // bitmap of group starts
    uint32_t* wpref;    // [W+1] exclusive prefix of popc(start[w])
    uint32_t* gsize;    // [ng]   dots in every group
};

struct K3Scratch {
    K3Groups D, A;      // groups of y-x (diagonals) and of y+x (anti-diagonals)
    uint32_t* aux;      // [nb/10+4] small histogram (median of the modal sub-bin)
};

struct K3Shared {       // block-wide accumulators (static shared memory)
    unsigned long long u64a, u64b;
    long long s64;
    unsigned int u32a, u32b, u32c, gmaxD, gmaxA;
    int imin, imax;
    int ng;
    unsigned int cnt11[11];
    int sel;            // selected bin or -1
    int m2min, m2max;
    long long icpt2;    // twice the intercept
    unsigned int warp_tot[K3_THREADS / 32];
    unsigned int warp_tot2[K3_THREADS / 32];
};

__device__ __forceinline__ void k3_setup_scratch(K3Scratch& s, uint32_t* base, int nb_cap) {
    const int wb = k3_words_bitmap(nb_cap), wg = k3_words_groups(nb_cap);
    K3Groups* sets[2] = {&s.D, &s.A};
    for (int i = 0; i < 2; ++i) {
        sets[i]->present = base; base += wb;
        sets[i]->start = base;   base += wb;
        sets[i]->wpref = base;   base += wb;
        sets[i]->gsize = base;   base += wg;
    }
    s.aux = base;
}

// inclusive prefix sum of arr[0..N) in place; every thread of the block must call it
__device__ void k3_block_scan(uint32_t* arr, int N, K3Shared& sh) {
    const int tid = threadIdx.x, T = K3_THREADS;
    int C = (N + T - 1) / T;
    C |= 1;                                            // odd chunk: conflict-free strided access
    const int b0 = min(tid * C, N), b1 = min(b0 + C, N);
    uint32_t sum = 0;
    for (int i = b0; i < b1; ++i) sum += arr[i];
    uint32_t inc = sum;
    const int lane = tid & 31, warp = tid >> 5;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) sh.warp_tot[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += sh.warp_tot[w];
    uint32_t run = base + inc - sum;
    for (int i = b0; i < b1; ++i) { run += arr[i]; arr[i] = run; }
    __syncthreads();
}

// The same for two arrays of the same length at once (the y-x and y+x group sets): one pair of barriers instead of two.
__device__ void k3_block_scan2(uint32_t* a0, uint32_t* a1, int N, K3Shared& sh) {
    const int tid = threadIdx.x, T = K3_THREADS;
    int C = (N + T - 1) / T;
    C |= 1;
    const int b0 = min(tid * C, N), b1 = min(b0 + C, N);
    uint32_t s0 = 0, s1 = 0;
    for (int i = b0; i < b1; ++i) { s0 += a0[i]; s1 += a1[i]; }
    uint32_t i0 = s0, i1 = s1;
    const int lane = tid & 31, warp = tid >> 5;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v0 = __shfl_up_sync(0xFFFFFFFFu, i0, o), v1 = __shfl_up_sync(0xFFFFFFFFu, i1, o);
        if (lane >= o) { i0 += v0; i1 += v1; }
    }
    if (lane == 31) { sh.warp_tot[warp] = i0; sh.warp_tot2[warp] = i1; }
    __syncthreads();
    uint32_t r0 = i0 - s0, r1 = i1 - s1;
    for (int w = 0; w < warp; ++w) { r0 += sh.warp_tot[w]; r1 += sh.warp_tot2[w]; }
    for (int i = b0; i < b1; ++i) { r0 += a0[i]; a0[i] = r0; r1 += a1[i]; a1[i] = r1; }
    __syncthreads();
}

// Counter increment for a whole warp at once.  The dots of a read pile up in one or two groups, so a plain
// shared-memory atomic per dot serialises 32 ways; lanes with the same index elect one lane to add their
// count.  Call from converged code; `idx < 0` = this lane has nothing to add.
__device__ __forceinline__ void k3_warp_count_add(uint32_t* C, int idx) {
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, idx < 0 ? -1 - lane : idx);
    if (idx >= 0 && lane == __ffs(peers) - 1) atomicAdd(&C[idx], (uint32_t)__popc(peers));
}

__device__ __forceinline__ int k3_group_of(const K3Groups& s, int b) {
    const int w = b >> 5;
    return (int)s.wpref[w] + __popc(s.start[w] & (0xFFFFFFFFu >> (31 - (b & 31)))) - 1;
}
__device__ __forceinline__ uint32_t k3_group_size(const K3Groups& s, int b) {
    return s.gsize[k3_group_of(s, b)];
}

// steps of a clustering round on one group set (every thread of the block calls them; barriers are the caller's)
__device__ __forceinline__ void k3_groups_clear(K3Groups& s, int W) {
    for (int w = threadIdx.x; w < W; w += K3_THREADS) s.present[w] = 0u;
    if (threadIdx.x == 0) s.wpref[0] = 0;
}
__device__ __forceinline__ void k3_groups_mark(K3Groups& s, int b) {
    const uint32_t bit = 1u << (b & 31);
    if (!(s.present[b >> 5] & bit)) atomicOr(&s.present[b >> 5], bit);      // most dots re-set a bit already set
}
__device__ __forceinline__ void k3_groups_starts(K3Groups& s, int W) {
    for (int w = threadIdx.x; w < W; w += K3_THREADS) {
        const uint32_t cur = s.present[w], prev = w ? s.present[w - 1] : 0u;
        const unsigned long long comb = ((unsigned long long)cur << 32) | prev;
        unsigned long long near = 0;
        #pragma unroll
        for (int sft = 1; sft <= 9; ++sft) near |= comb << sft;
        const uint32_t st = cur & ~(uint32_t)(near >> 32);
        s.start[w] = st;
        s.wpref[w + 1] = __popc(st);
    }
}
__device__ __forceinline__ uint32_t k3_groups_max(const K3Groups& s, int ng) {      // block-wide maximum via the caller's atomicMax
    uint32_t lmax = 0;
    for (int g = threadIdx.x; g < ng; g += K3_THREADS) lmax = max(lmax, s.gsize[g]);
    #pragma unroll
    for (int o = 16; o; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xFFFFFFFFu, lmax, o));
    return lmax;
}

// Pass 0 (span of x over all hits + checksum of the hit list) can ride on the first pass of a cleaning: the caller decides what
// to clean from the hit counts alone and applies the span gates afterwards, so a plot's hits are read one time less.
struct K3P0 { int minx, maxx; unsigned long long cs; };
__device__ __forceinline__ void k3_p0_init(K3Shared& sh) { sh.imin = 0x7FFFFFFF; sh.imax = -1; sh.u64a = 0; }     // one thread, before a barrier
__device__ __forceinline__ void k3_p0_reduce(K3Shared& sh, int lmin, int lmax, unsigned long long lsum) {          // every thread, before a barrier
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        lmin = min(lmin, __shfl_xor_sync(0xFFFFFFFFu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xFFFFFFFFu, lmax, o));
        lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&sh.imin, lmin); atomicMax(&sh.imax, lmax); atomicAdd(&sh.u64a, lsum); }
}

// Chain groups of the values binf(hit) in [0, nb) over the hits of a plot (binf < 0 = hit not taken): runs of
// occupied values whose gaps are < 10, and their sizes.  On return k3_group_size(set, value) answers per value and
// `gmax` (a field of sh) holds the largest group.  Every thread of the block must call it.
template <typename BinF>
__device__ void k3_build_groups(const uint2* hits, uint32_t H, K3Groups& s, int nb, K3Shared& sh, unsigned int K3Shared::*gmax, BinF binf,
                                K3P0* p0 = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int W = (nb + 31) >> 5;
    k3_groups_clear(s, W);
    if (tid == 0) { sh.*gmax = 0; if (p0) k3_p0_init(sh); }
    __syncthreads();
    {
        int lmin = 0x7FFFFFFF, lmax = -1;
        unsigned long long lsum = 0;
        for (uint32_t i = tid; i < H; i += K3_THREADS) {
            const uint2 h = hits[i];
            if (p0) { lmin = min(lmin, (int)h.x); lmax = max(lmax, (int)h.x); lsum += hit_mix(h.x, h.y & HIT_Y_MASK); }
            const int b = binf(h);
            if (b >= 0) k3_groups_mark(s, b);
        }
        if (p0) k3_p0_reduce(sh, lmin, lmax, lsum);
    }
    __syncthreads();
    if (p0) { p0->minx = sh.imin; p0->maxx = sh.imax; p0->cs = sh.u64a; }
    k3_groups_starts(s, W);
    __syncthreads();
    k3_block_scan(s.wpref + 1, W, sh);
    const int ng = (int)s.wpref[W];
    for (int g = tid; g < ng; g += K3_THREADS) s.gsize[g] = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < H; base += K3_THREADS) {            // uniform trip count: warp-aggregated adds
        const uint32_t i = base + tid;
        int g = -1;
        if (i < H) { const int b = binf(hits[i]); if (b >= 0) g = k3_group_of(s, b); }
        k3_warp_count_add(s.gsize, g);
    }
    __syncthreads();
    const uint32_t lmax = k3_groups_max(s, ng);
    if (lane == 0 && lmax) atomicMax(&(sh.*gmax), lmax);
    __syncthreads();
}

__device__ __forceinline__ void k3_zero(uint32_t* a, int n) {
    for (int i = threadIdx.x; i < n; i += K3_THREADS) a[i] = 0;
}

struct PlotView {
    uint2* hits;
    uint32_t H;
    int n, m;            // bins: d = y - x + (m-1) in [0, n+m-2], a = x + y in [0, n+m-2]
};

struct PlotStat {
    uint32_t nclean;
    unsigned long long sumabs;     // sum |x-y| over clean dots                (ABS)
    uint32_t cnt10;                // eu_dis_dots_within_10perc over clean dots (W10)
    double dir;                    // |eu_dis_dir_calcu| after re-centring      (REDEF)
};

// pass 0: span of x over all hits + checksum of the hit list
__device__ void k3_pass0(const PlotView& v, K3Shared& sh, int& minx, int& maxx, unsigned long long& csum) {
    if (threadIdx.x == 0) { sh.imin = 0x7FFFFFFF; sh.imax = -1; sh.u64a = 0; }
    __syncthreads();
    int lmin = 0x7FFFFFFF, lmax = -1;
    unsigned long long lsum = 0;
    for (uint32_t i = threadIdx.x; i < v.H; i += K3_THREADS) {
        const uint2 h = v.hits[i];
        const int x = (int)h.x;
        lmin = min(lmin, x); lmax = max(lmax, x);
        lsum += hit_mix(h.x, h.y & HIT_Y_MASK);
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        lmin = min(lmin, __shfl_xor_sync(0xFFFFFFFFu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xFFFFFFFFu, lmax, o));
        lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&sh.imin, lmin); atomicMax(&sh.imax, lmax); atomicAdd(&sh.u64a, lsum);
    }
    __syncthreads();
    minx = sh.imin; maxx = sh.imax; csum = sh.u64a;
    __syncthreads();
}

// Chain groups of y-x into set D and of y+x into set A in one go: every hit is read twice instead of four times,
// and both sets stay available, so no flag has to be written back between the two clusterings.
__device__ void k3_build_groups_both(const PlotView& v, K3Scratch& s, K3Shared& sh, K3P0* p0) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int nb = v.n + v.m - 1, moff = v.m - 1;
    const int W = (nb + 31) >> 5;
    k3_groups_clear(s.D, W);
    k3_groups_clear(s.A, W);
    if (tid == 0) { sh.gmaxD = 0; sh.gmaxA = 0; if (p0) k3_p0_init(sh); }
    __syncthreads();
    {
        int lmin = 0x7FFFFFFF, lmax = -1;
        unsigned long long lsum = 0;
        for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
            const uint2 h = v.hits[i];
            const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
            if (p0) { lmin = min(lmin, x); lmax = max(lmax, x); lsum += hit_mix(h.x, (uint32_t)y); }
            k3_groups_mark(s.D, y - x + moff);
            k3_groups_mark(s.A, y + x);
        }
        if (p0) k3_p0_reduce(sh, lmin, lmax, lsum);
    }
    __syncthreads();
    if (p0) { p0->minx = sh.imin; p0->maxx = sh.imax; p0->cs = sh.u64a; }
    k3_groups_starts(s.D, W);
    k3_groups_starts(s.A, W);
    __syncthreads();
    k3_block_scan2(s.D.wpref + 1, s.A.wpref + 1, W, sh);
    const int ngD = (int)s.D.wpref[W], ngA = (int)s.A.wpref[W];
    for (int g = tid; g < ngD; g += K3_THREADS) s.D.gsize[g] = 0u;
    for (int g = tid; g < ngA; g += K3_THREADS) s.A.gsize[g] = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < v.H; base += K3_THREADS) {          // uniform trip count: warp-aggregated adds
        const uint32_t i = base + tid;
        int gd = -1, ga = -1;
        if (i < v.H) {
            const uint2 h = v.hits[i];
            const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
            gd = k3_group_of(s.D, y - x + moff);
            ga = k3_group_of(s.A, y + x);
        }
        k3_warp_count_add(s.D.gsize, gd);
        k3_warp_count_add(s.A.gsize, ga);
    }
    __syncthreads();
    const uint32_t mD = k3_groups_max(s.D, ngD), mA = k3_groups_max(s.A, ngA);
    if (lane == 0) { if (mD) atomicMax(&sh.gmaxD, mD); if (mA) atomicMax(&sh.gmaxA, mA); }
    __syncthreads();
}

// clean_dotdata_diagnal_and_anti_diagnal (Simple_function.pyx:432-448): keep a dot unless its y-x
// chain group and its y+x chain group both have <= 10 members.  Returns count and sum |x-y| of the kept dots;
// sets HIT_F_CLEAN on them when `write_flags` (the REDEF statistics read it back).
__device__ void k3_clean_a6(const PlotView& v, K3Scratch& s, K3Shared& sh, PlotStat& st, bool write_flags, K3P0* p0) {
    const int moff = v.m - 1;
    k3_build_groups_both(v, s, sh, p0);
    if (threadIdx.x == 0) { sh.u32a = 0; sh.u64a = 0; }
    __syncthreads();
    uint32_t lcnt = 0; unsigned long long lsum = 0;
    for (uint32_t i = threadIdx.x; i < v.H; i += K3_THREADS) {
        const uint2 h = v.hits[i];
        const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
        const bool keep = k3_group_size(s.D, y - x + moff) > 10u || k3_group_size(s.A, y + x) > 10u;
        if (write_flags) v.hits[i].y = (uint32_t)y | (keep ? HIT_F_CLEAN : 0u);
        if (keep) { ++lcnt; lsum += (unsigned long long)abs(x - y); }
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        lcnt += __shfl_xor_sync(0xFFFFFFFFu, lcnt, o);
        lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh.u32a, lcnt); atomicAdd(&sh.u64a, lsum); }
    __syncthreads();
    st.nclean = sh.u32a; st.sumabs = sh.u64a;
    __syncthreads();
}

// dis_cluster keep rule (Simple_function.pyx:560-563): groups with > 50 members, or, when there is
// none, every group tied for the maximum size.
__device__ __forceinline__ bool k3_a7_keep(uint32_t sz, uint32_t gmax) {
    return gmax > 50u ? sz > 50u : sz == gmax;
}

// W10 cleaning (Simple_function.pyx:281-288): dis_cluster on y-x, then dis_cluster on y+x over the
// dots the first step did not keep; union.  Returns count and the within-16% count.
__device__ void k3_clean_w10(const PlotView& v, K3Scratch& s, K3Shared& sh, PlotStat& st, K3P0* p0) {
    st.nclean = 0; st.cnt10 = 0;
    if (v.H == 0) {                                     // uniform across the block
        if (p0) { p0->minx = 0x7FFFFFFF; p0->maxx = -1; p0->cs = 0; }
        return;
    }
    const int nb = v.n + v.m - 1;
    const int moff = v.m - 1;
    if (threadIdx.x == 0) { sh.u32a = 0; sh.u32b = 0; sh.u32c = 0; }
    k3_build_groups(v.hits, v.H, s.D, nb, sh, &K3Shared::gmaxD,
                    [moff](const uint2& h) { return (int)(h.y & HIT_Y_MASK) - (int)h.x + moff; }, p0);
    const uint32_t gmax1 = sh.gmaxD;
    const K3Groups D = s.D;
    auto kept1 = [D, moff, gmax1](const uint2& h) {
        return k3_a7_keep(k3_group_size(D, (int)(h.y & HIT_Y_MASK) - (int)h.x + moff), gmax1);
    };
    uint32_t lkept = 0;
    for (uint32_t i = threadIdx.x; i < v.H; i += K3_THREADS) lkept += kept1(v.hits[i]);
    #pragma unroll
    for (int o = 16; o; o >>= 1) lkept += __shfl_xor_sync(0xFFFFFFFFu, lkept, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh.u32a, lkept);
    __syncthreads();
    const uint32_t kept1_n = sh.u32a;
    const bool have_left = kept1_n < v.H;
    uint32_t gmax2 = 0;
    if (have_left) {                                     // second clustering, on y+x, over the dots the first did not keep
        k3_build_groups(v.hits, v.H, s.A, nb, sh, &K3Shared::gmaxA,
                        [kept1](const uint2& h) { return kept1(h) ? -1 : (int)(h.y & HIT_Y_MASK) + (int)h.x; });
        gmax2 = sh.gmaxA;
    }
    uint32_t lcnt = 0, l10 = 0;
    for (uint32_t i = threadIdx.x; i < v.H; i += K3_THREADS) {
        const uint2 h = v.hits[i];
        const uint32_t y = h.y & HIT_Y_MASK;
        bool keep = kept1(h);
        if (!keep && have_left) keep = k3_a7_keep(k3_group_size(s.A, (int)y + (int)h.x), gmax2);
        if (keep) {
            ++lcnt;
            const int x = (int)h.x;
            // abs(float(x-y)/float(x)) < 0.16 for x > 0  (Simple_function.pyx:732-733) as the exact integer test
            // 25 |x-y| < 4 x: a quotient of two integers below 2^28 that is not 4/25 differs from it by far more than an
            // ulp, and an exact 4/25 rounds to the literal 0.16 itself (not below it)
            if (x > 0 && 25ll * (long long)abs(x - (int)y) < 4ll * (long long)x) ++l10;
        }
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        lcnt += __shfl_xor_sync(0xFFFFFFFFu, lcnt, o);
        l10 += __shfl_xor_sync(0xFFFFFFFFu, l10, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh.u32b, lcnt); atomicAdd(&sh.u32c, l10); }
    __syncthreads();
    st.nclean = sh.u32b; st.cnt10 = sh.u32c;
    __syncthreads();
}

// bin of value d among the 11 edges mn + i*(mx-mn)/10 as number_cluster assigns it
// (Simple_function.pyx:1104-1118).  "d < mn + i*range/10.0" in doubles is exactly
// "10*(d-mn) < i*range" in integers (kernel 3), so the bin is floor(10*(d-mn)/range),
// and everything lands in the last bin when range == 0.
__device__ __forceinline__ int k3_bin11(int d, int mn, int range) {
    if (range <= 0) return 10;
    // d - mn <= range < 2^28 (sequence lengths are below 2^27), so 10 (d - mn) fits 32 unsigned bits.  The quotient is at
    // most 10: a single-precision estimate is off by less than one, two integer multiplies put it right (a 32-bit
    // division is ~20 instructions and this runs six times per dot of a REDEF task).
    const uint32_t num = 10u * (uint32_t)(d - mn), den = (uint32_t)range;
    if (num > 10u * den) return 11;                             // only for a d outside [mn, mn + range]: no bin
    int q = (int)((float)num * __frcp_rn((float)den));
    q = min(max(q, 0), 10);
    if ((uint32_t)q * den > num) --q;
    else if ((uint32_t)(q + 1) * den <= num) ++q;
    return q;
}

// dis_to_diagnal_most_abundant_defined + eu_dis_dir_calcu on the clean dots (Simple_function.pyx:582-591,
// 248-251, 710-722).  Needs HIT_F_CLEAN flags from k3_clean_a6 and st.nclean > 0.
__device__ void k3_redef_stat(const PlotView& v, K3Scratch& s, K3Shared& sh, PlotStat& st) {
    const int tid = threadIdx.x, lane = tid & 31;
    // level 1: range of d over the clean dots
    if (tid == 0) { sh.imin = 0x7FFFFFFF; sh.imax = -0x7FFFFFFF; sh.icpt2 = 0; sh.sel = -1; }
    if (tid < 11) sh.cnt11[tid] = 0;
    __syncthreads();
    int lmin = 0x7FFFFFFF, lmax = -0x7FFFFFFF;
    for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
        const uint2 h = v.hits[i];
        if (h.y & HIT_F_CLEAN) {
            const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
            lmin = min(lmin, d); lmax = max(lmax, d);
        }
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        lmin = min(lmin, __shfl_xor_sync(0xFFFFFFFFu, lmin, o));
        lmax = max(lmax, __shfl_xor_sync(0xFFFFFFFFu, lmax, o));
    }
    if (lane == 0) { atomicMin(&sh.imin, lmin); atomicMax(&sh.imax, lmax); }
    __syncthreads();
    const int mn1 = sh.imin, rg1 = sh.imax - sh.imin;
    for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
        const uint2 h = v.hits[i];
        if (h.y & HIT_F_CLEAN) atomicAdd(&sh.cnt11[k3_bin11((int)(h.y & HIT_Y_MASK) - (int)h.x, mn1, rg1)], 1u);
    }
    __syncthreads();
    if (tid == 0) {                                    // find_longest_list: exactly one modal bin?
        unsigned best = 0; int nb = 0, bi = -1;
        for (int b = 0; b < 11; ++b) best = max(best, sh.cnt11[b]);
        for (int b = 0; b < 11; ++b) if (sh.cnt11[b] == best) { ++nb; bi = b; }
        sh.sel = (nb == 1) ? bi : -1;
        sh.m2min = 0x7FFFFFFF; sh.m2max = -0x7FFFFFFF;
    }
    __syncthreads();
    const int b1 = sh.sel;
    if (b1 >= 0) {                                     // uniform
        // level 2 inside the modal bin
        __syncthreads();
        if (tid < 11) sh.cnt11[tid] = 0;
        lmin = 0x7FFFFFFF; lmax = -0x7FFFFFFF;
        for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
            const uint2 h = v.hits[i];
            if (h.y & HIT_F_CLEAN) {
                const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
                if (k3_bin11(d, mn1, rg1) == b1) { lmin = min(lmin, d); lmax = max(lmax, d); }
            }
        }
        #pragma unroll
        for (int o = 16; o; o >>= 1) {
            lmin = min(lmin, __shfl_xor_sync(0xFFFFFFFFu, lmin, o));
            lmax = max(lmax, __shfl_xor_sync(0xFFFFFFFFu, lmax, o));
        }
        if (lane == 0) { atomicMin(&sh.m2min, lmin); atomicMax(&sh.m2max, lmax); }
        __syncthreads();
        const int mn2 = sh.m2min, rg2 = sh.m2max - sh.m2min;
        for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
            const uint2 h = v.hits[i];
            if (h.y & HIT_F_CLEAN) {
                const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
                if (k3_bin11(d, mn1, rg1) == b1) atomicAdd(&sh.cnt11[k3_bin11(d, mn2, rg2)], 1u);
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned best = 0; int nb = 0, bj = -1;
            for (int b = 0; b < 11; ++b) best = max(best, sh.cnt11[b]);
            for (int b = 0; b < 11; ++b) if (sh.cnt11[b] == best) { ++nb; bj = b; }
            sh.sel = (nb == 1) ? bj : -1;
        }
        __syncthreads();
        const int b2 = sh.sel;
        if (b2 >= 0) {                                 // uniform: np.median of that sub-bin
            const uint32_t c = sh.cnt11[b2];
            k3_zero(s.aux, rg2 + 1);
            __syncthreads();
            for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
                const uint2 h = v.hits[i];
                if (h.y & HIT_F_CLEAN) {
                    const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
                    if (k3_bin11(d, mn1, rg1) == b1 && k3_bin11(d, mn2, rg2) == b2) atomicAdd(&s.aux[d - mn2], 1u);
                }
            }
            __syncthreads();
            // np.median of the sub-bin: the values of rank (c-1)/2 and c/2, found on the inclusive prefix of the value counts
            k3_block_scan(s.aux, rg2 + 1, sh);
            {
                const uint32_t r0 = (c - 1) / 2, r1 = c / 2;
                for (int t = tid; t <= rg2; t += K3_THREADS) {
                    const uint32_t lo = t ? s.aux[t - 1] : 0u, hi = s.aux[t];
                    if (lo <= r0 && hi > r0) sh.m2min = t;                 // m2min / m2max are free again: reused for the two ranks
                    if (lo <= r1 && hi > r1) sh.m2max = t;
                }
            }
            __syncthreads();
            if (tid == 0) sh.icpt2 = (long long)(sh.m2min + mn2) + (long long)(sh.m2max + mn2);
            __syncthreads();
        }
    }
    __syncthreads();
    // eu_dis_dir_calcu on (x + intercept, y): in doubled integers A = 2(x'-y), B = 2x'
    const long long icpt2 = sh.icpt2;
    if (tid == 0) { sh.s64 = 0; sh.u32a = 0; }
    __syncthreads();
    long long lsum = 0; uint32_t lcnt = 0;
    for (uint32_t i = tid; i < v.H; i += K3_THREADS) {
        const uint2 h = v.hits[i];
        if (h.y & HIT_F_CLEAN) {
            const long long y = (long long)(h.y & HIT_Y_MASK);
            const long long B = 2ll * (long long)h.x + icpt2;
            const long long A = B - 2ll * y;
            // |(A/2) / (B/2)| > 0.1 (x' == 0: |A/2| / 1 > 0.1, i.e. A != 0) as the exact integer test 10 |A| > |B|: the
            // quotient of two integers below 2^30 that is not 1/10 differs from it by far more than an ulp, and an exact
            // 1/10 rounds to the literal 0.1 itself (not above it)
            const bool far = (B == 0) ? (A != 0) : (10ll * llabs(A) > llabs(B));
            if (far) { lsum += A; ++lcnt; }
        }
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        lsum += __shfl_xor_sync(0xFFFFFFFFu, lsum, o);
        lcnt += __shfl_xor_sync(0xFFFFFFFFu, lcnt, o);
    }
    if (lane == 0) { atomicAdd((unsigned long long*)&sh.s64, (unsigned long long)lsum); atomicAdd(&sh.u32a, lcnt); }
    __syncthreads();
    if (sh.u32a == 0) st.dir = 0.0001;
    else st.dir = fabs(((double)sh.s64 * 0.5) / (double)sh.u32a);
    __syncthreads();
}

struct EvalResult {
    double a, b;         // the [a, b] pair the reference function returns
    bool valid;          // not (0 in pair)
};

__device__ __forceinline__ double k3_pair_score(const EvalResult& e) { return 1.0 - e.b / e.a; }

struct K3Params {
    const Task* tasks;
    const int32_t* task_ids;     // tasks of this launch
    const int32_t* out_ids;      // where task_ids[i]'s results go (null: at the task's own index)
    int n_ids;
    const Plot* plots;
    const uint32_t* cnt;         // hits per plot
    const int32_t* op_status;
    uint2* hits;
    uint32_t* gscratch;          // global scratch, [gridDim.x][scratch_words] (only when smem is too small)
    int nb_cap;                  // bins the scratch can hold
    int use_global;
    unsigned int* queue;         // warp kernel (k3_warp.cuh): next task of this launch
    double* task_score; uint8_t* task_status; double* task_stat; uint32_t* task_hits;
    unsigned long long* task_hitsum;
};

// One evaluation of one reference mode on the (ref, alt) plots.  Block-uniform control flow.
// Every phase below runs in a two-trip loop over (ref plot, alt plot) that is deliberately NOT unrolled, and the
// kernel calls this function from one call site: fully inlined and unrolled the kernel was 318 KB of code and spent
// more time waiting for instruction fetch than for anything else.
__device__ void k3_eval(int mode, const PlotView* pv /*[2]: ref, alt*/, int len_ref, int len_alt,
                        K3Scratch& s, K3Shared& sh, EvalResult& out, unsigned long long* cs /*[2]*/,
                        int* minx /*[2]*/, int* maxx /*[2]*/, bool have_pass0)
{
    out.a = 0; out.b = 0; out.valid = false;
    const double Hr = (double)pv[0].H, Ha = (double)pv[1].H;
    const double Lr = (double)len_ref, La = (double)len_alt;
    // What will be cleaned unless a span gate stops it follows from the hit counts alone (first steps of the ladders,
    // Simple_function.pyx:187-190, 280-281, 244-246): the cleaning runs first, with pass 0 (spans + checksums; the caller hands
    // them over when it has them) riding on its first pass over the hits, and the span gates are applied afterwards.  A task
    // whose spans fail them was cleaned for nothing -- rare --, every other task reads its hits one time less.
    bool want_a6 = false, want_w10 = false;
    if (mode == 0) want_a6 = (pv[0].H > 2 && pv[1].H > 2) && (Hr / fmin(Lr, La) > 0.1);
    else if (mode == 1) want_w10 = fmax(Hr / Lr, Ha / La) > 0.1;
    else want_a6 = Hr / Lr > 0.1 && Ha / La > 0.1;
    PlotStat st[2] = {};
    K3P0 p0[2];
    const bool fuse = !have_pass0 && (want_a6 || want_w10);
    if (!have_pass0 && !fuse) {
        #pragma unroll 1
        for (int w = 0; w < 2; ++w) k3_pass0(pv[w], sh, minx[w], maxx[w], cs[w]);
    }
    if (want_a6) {
        #pragma unroll 1
        for (int w = 0; w < 2; ++w) k3_clean_a6(pv[w], s, sh, st[w], mode == 2, fuse ? &p0[w] : nullptr);
    } else if (want_w10) {
        #pragma unroll 1
        for (int w = 0; w < 2; ++w) k3_clean_w10(pv[w], s, sh, st[w], fuse ? &p0[w] : nullptr);
    }
    if (fuse) {
        #pragma unroll 1
        for (int w = 0; w < 2; ++w) { minx[w] = p0[w].minx; maxx[w] = p0[w].maxx; cs[w] = p0[w].cs; }
    }
    const double span_r = (double)(maxx[0] - minx[0]) / Lr, span_a = (double)(maxx[1] - minx[1]) / La;
    bool clean_a6 = false, clean_w10 = false;
    if (mode == 0) {                                   // ABS, Simple_function.pyx:187-203
        if (!want_a6) return;
        const bool rs = span_r > 0.6, as = span_a > 0.6;
        if (rs && as) clean_a6 = true;
        else if (rs) { out.a = 1.1; out.b = 2.1; }
        else if (as) { out.a = 2.1; out.b = 1.1; }
    } else if (mode == 1) {                            // W10, Simple_function.pyx:280-294
        if (!want_w10) return;
        clean_w10 = true;
    } else {                                           // REDEF, Simple_function.pyx:244-257
        if (!want_a6) return;
        if (!(span_r > 0.7 && span_a > 0.7)) return;
        clean_a6 = true;
    }
    if ((clean_a6 || clean_w10) && st[0].nclean > 0 && st[1].nclean > 0) {
        if (mode == 0) {
            out.a = (double)st[0].sumabs / (double)st[0].nclean;      // np.mean of exact integers
            out.b = (double)st[1].sumabs / (double)st[1].nclean;
        } else if (mode == 1) {
            out.a = (double)st[1].cnt10; out.b = (double)st[0].cnt10;  // swapped on purpose (:290)
        } else {
            #pragma unroll 1
            for (int w = 0; w < 2; ++w) k3_redef_stat(pv[w], s, sh, st[w]);
            out.a = st[0].dir; out.b = st[1].dir;
        }
    }
    out.valid = (out.a != 0.0) && (out.b != 0.0);       // `if not 0 in pair`
}

#ifndef K3_MINB
#define K3_MINB 6                              // measured: 6 resident CTAs per SM (42 registers) beats 4 (64)
#endif

__global__ void __launch_bounds__(K3_THREADS, K3_MINB)
k3_score_reads(const K3Params p)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    __shared__ K3Shared sh;
    K3Scratch s;
    k3_setup_scratch(s, p.use_global ? p.gscratch + (size_t)blockIdx.x * k3_scratch_words(p.nb_cap) : s_dyn, p.nb_cap);

    for (int it = blockIdx.x; it < p.n_ids; it += gridDim.x) {
        const int tin = p.task_ids[it];
        const int tix = p.out_ids ? p.out_ids[it] : tin;
        const Task t = p.tasks[tin];
        PlotView pv[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (t.plot[i] >= 0) {
                const Plot pl = p.plots[t.plot[i]];
                // a plot that overflowed its first-pass capacity is incomplete (slots near the end of its list may never have
                // been written): it counts as empty here, and its task is scored again by redo_wave with the exact capacity
                const uint32_t found = p.cnt[t.plot[i]];
                pv[i].hits = p.hits + pl.hit_off; pv[i].H = found > pl.cap ? 0u : found; pv[i].n = pl.n; pv[i].m = pl.m;
            } else { pv[i].hits = nullptr; pv[i].H = 0; pv[i].n = 0; pv[i].m = 0; }
        }
        const bool bad = p.op_status[t.read_op] != 0;
        EvalResult ev[2] = {{0, 0, false}, {0, 0, false}};
        unsigned long long cs[4] = {0, 0, 0, 0};
        if (!bad) {
            const int n_eval = (t.mode == 3) ? 2 : 1;                   // the simple-DEL rule asks ABS, then W10
            int minx[4], maxx[4];
            #pragma unroll 1
            for (int e = 0; e < n_eval; ++e) {
                // the W10 opinion of the simple-DEL rule looks at the same two plots unless the structures hold lower case
                const bool same = e == 1 && t.plot[2] == t.plot[0] && t.plot[3] == t.plot[1];
                if (same) { minx[2] = minx[0]; maxx[2] = maxx[0]; minx[3] = minx[1]; maxx[3] = maxx[1]; cs[2] = cs[0]; cs[3] = cs[1]; }
                k3_eval((t.mode == 3) ? e : t.mode, pv + 2 * e, t.len_ref, t.len_alt, s, sh, ev[e], cs + 2 * e,
                        minx + 2 * e, maxx + 2 * e, same);
            }
        }
        const EvalResult ea = ev[0], eb = ev[1];
        if (threadIdx.x == 0) {
            double score = 0.0; uint8_t status = 0;
            if (bad) status = 2;
            else if (t.mode == 3) {                    // simple-DEL rule, Simple_function.pyx:1718-1726
                if (ea.valid && eb.valid) { const double s1 = k3_pair_score(ea), s2 = k3_pair_score(eb); score = (s2 < s1) ? s2 : s1; status = 1; }
                else if (ea.valid) { score = k3_pair_score(ea); status = 1; }
                else if (eb.valid) { score = k3_pair_score(eb); status = 1; }
            } else if (ea.valid) { score = k3_pair_score(ea); status = 1; }
            p.task_score[tix] = score;
            p.task_status[tix] = status;
            p.task_stat[4 * tix + 0] = ea.a; p.task_stat[4 * tix + 1] = ea.b;
            p.task_stat[4 * tix + 2] = eb.a; p.task_stat[4 * tix + 3] = eb.b;
            for (int i = 0; i < 4; ++i) { p.task_hits[4 * tix + i] = pv[i].H; p.task_hitsum[4 * tix + i] = cs[i]; }
        }
        __syncthreads();
    }
}

}  // namespace vb
