// Kernel 3, warp variant: one WARP per task, several tasks per CTA, no block-wide barrier anywhere.
//
// Same reference functions as k3_score.cuh (vapor_vali/Simple_function.pyx:182-203, 241-257, 277-294, 404-448,
// 551-591, 705-733, 788-792, 1104-1118, 1718-1726, 1913-1915) and the same bitmap chain-grouping, re-cut for tasks
// with small value ranges.  It can take ranges up to 26 624 bins; by default it gets the classes up to 8 192 bins
// (63 % of the tasks of a 50 bp - 5 kb SV list), where it needs 23 ns per task against 63 ns for the CTA kernel -- above
// that its scratch leaves too few warps per SM and the CTA-per-task kernel wins (api.cu, k3w_nclass):
//
//   * a task is a few hundred to a few thousand dots.  With 256 threads on it, two thirds of the instructions of the
//     CTA kernel were fixed cost per phase (about 70 phases per task: barriers, block reductions through shared
//     atomics, scans of short arrays) and every phase exposed one global-memory latency.  A warp pays a shuffle
//     reduction per phase, keeps four 256-byte loads in flight per pass, and 32 tasks are resident per SM
//     instead of 6;
//   * scratch per task is 0.78 bytes per bin instead of 1.95: the group-start bitmap overwrites the occupancy
//     bitmap in place (the neighbour word travels by shuffle), word prefixes and group sizes are 16-bit (a plot of
//     this class holds at most n + m + 32 < 65 536 dots in its first-pass slab; the tasks of plots that overflowed it
//     are re-scored by the CTA kernel, api.cu launch_k3_class), the median histogram of REDEF reuses the group-size arrays;
//   * the passes over a plot's dots are fused across the two opinions of the simple-DEL rule (ABS, then W10 on the
//     same two plots, :1718-1726): both need the chain groups of y-x over all dots, built once -- 7 passes over the
//     dots instead of 10;
//   * the per-dot floating-point tests of the reference are evaluated as the equivalent exact integer tests
//     (|x-y|/x < 0.16  <=>  25|x-y| < 4x;  |A/B| > 0.1  <=>  10|A| > |B|: the quotient of two integers below 2^20
//     that differs from 4/25 or 1/10 differs by far more than an ulp, and an exact 4/25 or 1/10 rounds to the
//     literal itself), the 11-edge binning by a float estimate corrected with two integer multiplies;
//   * every pass body exists once (the four dots in flight rotate through one register pair) and the three stages
//     have one call site each: with the bodies unrolled four times the kernel was 180 KB of code and instruction
//     fetch was its first stall reason;
//   * pass 0 (span of x, checksum) rides on the first cleaning pass: what to clean follows from the hit counts, the span
//     gates are applied afterwards (a task whose spans fail them was cleaned for nothing -- rare);
//   * tasks are pulled from a per-launch queue (one atomic per task), so warps never wait for the slowest
//     task of a CTA.
// Results are bit-identical to the CTA kernel's (all GPU tests and the soak run both through the oracle).
#pragma once
#include "k3_score.cuh"

namespace vb {

constexpr int K3W_TEAMS = 4;                    // warps (= tasks in flight) per CTA
constexpr int K3W_MAX_NB = 26624;               // largest value range handled here
constexpr unsigned K3W_FULL = 0xFFFFFFFFu;

// scratch of one team in 32-bit words: [bitsD][bitsA][prefD][prefA][gsD][gsA]
__host__ __device__ inline int k3w_words_bits(int nb) { return nb / 32 + 2; }
__host__ __device__ inline int k3w_words_pref(int nb) { return (k3w_words_bits(nb) + 2) / 2; }
__host__ __device__ inline int k3w_words_gs(int nb)   { return (nb / 10 + 6) / 2; }
__host__ __device__ inline int k3w_scratch_words(int nb) {
    return 2 * (k3w_words_bits(nb) + k3w_words_pref(nb) + k3w_words_gs(nb));
}

struct K3WSet {            // chain groups of one value axis
    uint32_t* bits;        // occupancy bitmap, then (in place) bitmap of group starts
    uint16_t* pref;        // [W] group starts before word w
    uint32_t* gs;          // group sizes, two 16-bit counters per word
};

__device__ __forceinline__ void k3w_clear(uint32_t* a, int n, int lane) {
    #pragma unroll 2
    for (int i = lane; i < n; i += 32) a[i] = 0u;
}
__device__ __forceinline__ void k3w_mark(uint32_t* bits, int b) {
    const uint32_t bit = 1u << (b & 31);
    if (!(bits[b >> 5] & bit)) atomicOr(&bits[b >> 5], bit);       // most dots re-set a bit already set
}
// occupancy -> group starts (a set bit with 9 clear bits below it) in place, + exclusive word prefix; returns #groups
__device__ __forceinline__ int k3w_starts(const K3WSet& s, int W, int lane) {
    uint32_t carry = 0u, run = 0u;
    #pragma unroll 1
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        const uint32_t cur = w < W ? s.bits[w] : 0u;
        uint32_t prev = __shfl_up_sync(K3W_FULL, cur, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(K3W_FULL, cur, 31);
        const unsigned long long comb = ((unsigned long long)cur << 32) | prev;
        const unsigned long long t1 = comb | (comb << 1);           // shifts 0..1
        const unsigned long long t2 = t1 | (t1 << 2);               // 0..3
        const unsigned long long t4 = t2 | (t2 << 4);               // 0..7
        const unsigned long long near = (t4 << 1) | (comb << 9);    // 1..9
        const uint32_t st = cur & ~(uint32_t)(near >> 32);
        const uint32_t c = (uint32_t)__popc(st);
        uint32_t inc = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(K3W_FULL, inc, o);
            if (lane >= o) inc += v;
        }
        if (w < W) { s.bits[w] = st; s.pref[w] = (uint16_t)(run + inc - c); }
        run += __shfl_sync(K3W_FULL, inc, 31);
    }
    __syncwarp();
    return (int)run;
}
__device__ __forceinline__ int k3w_group(const K3WSet& s, int b) {
    const int w = b >> 5;
    return (int)s.pref[w] + __popc(s.bits[w] & (0xFFFFFFFFu >> (31 - (b & 31)))) - 1;
}
__device__ __forceinline__ uint32_t k3w_gsize(const K3WSet& s, int b) {
    const int g = k3w_group(s, b);
    return (s.gs[g >> 1] >> ((g & 1) << 4)) & 0xFFFFu;
}
// Counter increment for the whole warp (converged code only; idx < 0 = nothing to add).  The dots of a read pile up in
// one or two groups: the lanes that share the first valid lane's index add their count with ONE atomic, the others add
// 1 each (match.any did this exactly, but its latency was 9 % of the kernel's stall samples).
__device__ __forceinline__ void k3w_count_add16(uint32_t* gs, int g, int lane) {
    const unsigned valid = __ballot_sync(K3W_FULL, g >= 0);
    if (valid == 0u) return;                                        // warp-uniform
    const int leader = __ffs(valid) - 1;
    const int g0 = __shfl_sync(K3W_FULL, g, leader);
    const unsigned same = __ballot_sync(K3W_FULL, g == g0);
    if (g == g0) { if (lane == leader) atomicAdd(&gs[g0 >> 1], (uint32_t)__popc(same) << ((g0 & 1) << 4)); }
    else if (g >= 0) atomicAdd(&gs[g >> 1], 1u << ((g & 1) << 4));
}
__device__ __forceinline__ void k3w_count_add32(uint32_t* c, int idx, int lane) {
    const unsigned valid = __ballot_sync(K3W_FULL, idx >= 0);
    if (valid == 0u) return;
    const int leader = __ffs(valid) - 1;
    const int i0 = __shfl_sync(K3W_FULL, idx, leader);
    const unsigned same = __ballot_sync(K3W_FULL, idx == i0);
    if (idx == i0) { if (lane == leader) atomicAdd(&c[i0], (uint32_t)__popc(same)); }
    else if (idx >= 0) atomicAdd(&c[idx], 1u);
}
__device__ __forceinline__ uint32_t k3w_gmax(const uint32_t* gs, int ng, int lane) {
    uint32_t m = 0;
    #pragma unroll 2
    for (int i = lane; i < (ng + 1) / 2; i += 32) { const uint32_t v = gs[i]; m = max(m, max(v & 0xFFFFu, v >> 16)); }
    return __reduce_max_sync(K3W_FULL, m);
}
__device__ __forceinline__ unsigned long long k3w_sum64(unsigned long long v) {
    #pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(K3W_FULL, v, o);
    return v;
}

// Every dot of a plot, four coalesced loads in flight per lane; f(hit, index, valid) is called from converged code.
// The body exists ONCE per pass (the four dots rotate through one register pair): with the body unrolled four times
// the kernel was 180 KB of code and its top stall reason was instruction fetch.
template <typename F>
__device__ __forceinline__ void k3w_each(const uint2* hits, uint32_t H, int lane, F&& f) {
    #pragma unroll 1
    for (uint32_t base = 0; base < H; base += 128) {
        const uint32_t i0 = base + lane;
        uint2 h0 = i0 < H ? hits[i0] : make_uint2(0u, 0u);
        uint2 h1 = i0 + 32 < H ? hits[i0 + 32] : make_uint2(0u, 0u);
        uint2 h2 = i0 + 64 < H ? hits[i0 + 64] : make_uint2(0u, 0u);
        uint2 h3 = i0 + 96 < H ? hits[i0 + 96] : make_uint2(0u, 0u);
        const uint32_t nu = min(4u, (H - base + 31u) >> 5);         // warp-uniform
        #pragma unroll 1
        for (uint32_t u = 0; u < nu; ++u) {
            const uint32_t i = i0 + 32 * u;
            f(h0, i, i < H);
            h0 = h1; h1 = h2; h2 = h3;
        }
    }
}

// bin of value d among the 11 edges of number_cluster (see k3_bin11): floor(10 (d - mn) / range), in [0, 10]
__device__ __forceinline__ int k3w_bin11(int d, int mn, int range, float inv_range) {
    if (range <= 0) return 10;
    const uint32_t num = 10u * (uint32_t)(d - mn);
    int q = (int)((float)num * inv_range);
    q = min(max(q, 0), 10);
    if ((uint32_t)q * (uint32_t)range > num) --q;
    else if ((uint32_t)(q + 1) * (uint32_t)range <= num) ++q;
    return q;
}

struct K3WStat {
    uint32_t n6; unsigned long long sumabs; int dmin, dmax;    // a6: kept dots, sum |x-y|, range of y-x over them
    uint32_t nw, c10;                                          // w10: kept dots, dots within 16 %
};

// pass 0: span of x over all dots + checksum of the dot list
__device__ __forceinline__ void k3w_pass0(const PlotView& v, int lane, int& minx, int& maxx, unsigned long long& csum) {
    int lmin = 0x7FFFFFFF, lmax = -1;
    unsigned long long lsum = 0;
    k3w_each(v.hits, v.H, lane, [&](const uint2& h, uint32_t, bool ok) {
        if (ok) {
            const int x = (int)h.x;
            lmin = min(lmin, x); lmax = max(lmax, x);
            lsum += hit_mix(h.x, h.y & HIT_Y_MASK);
        }
    });
    minx = __reduce_min_sync(K3W_FULL, lmin);
    maxx = __reduce_max_sync(K3W_FULL, lmax);
    csum = k3w_sum64(lsum);
}

// The cleaning of one plot for up to two opinions at once:
//   want6   clean_dotdata_diagnal_and_anti_diagnal (:432-448): keep a dot unless its y-x chain group and its y+x chain
//           group both have <= 10 members (-> n6, sumabs, range of y-x over the kept dots; HIT_F_CLEAN when `flags`)
//   want10  the W10 cleaning (:281-288): dis_cluster on y-x, then dis_cluster on y+x over the dots the first step did not
//           keep, union (-> nw, c10)
// Both start from the chain groups of y-x over all dots, built once.
// The first pass over the dots also does pass 0 (span of x, checksum): the caller decides what to clean from the hit
// counts alone and applies the span gates afterwards, so a task's plots are read one time less.
__device__ __forceinline__ void k3w_clean(const PlotView& v, const K3WSet& D, const K3WSet& A, int lane,
                                          bool want6, bool want10, bool flags, K3WStat& st,
                                          int& minx, int& maxx, unsigned long long& csum) {
    st.n6 = 0; st.sumabs = 0; st.dmin = 0x7FFFFFFF; st.dmax = -0x7FFFFFFF; st.nw = 0; st.c10 = 0;
    if (v.H == 0 || !(want6 || want10)) { k3w_pass0(v, lane, minx, maxx, csum); return; }   // warp-uniform
    const int nb = v.n + v.m - 1, moff = v.m - 1;
    const int W = (nb + 31) >> 5;
    const uint32_t H = v.H;
    uint2* hits = v.hits;
    k3w_clear(D.bits, W, lane);
    if (want6) k3w_clear(A.bits, W, lane);
    __syncwarp();
    {
        int lmin = 0x7FFFFFFF, lmax = -1;
        unsigned long long lsum = 0;
        k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
            if (ok) {
                const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
                lmin = min(lmin, x); lmax = max(lmax, x);
                lsum += hit_mix(h.x, (uint32_t)y);
                k3w_mark(D.bits, y - x + moff);
                if (want6) k3w_mark(A.bits, y + x);
            }
        });
        minx = __reduce_min_sync(K3W_FULL, lmin);
        maxx = __reduce_max_sync(K3W_FULL, lmax);
        csum = k3w_sum64(lsum);
    }
    __syncwarp();
    const int ngD = k3w_starts(D, W, lane);
    const int ngA = want6 ? k3w_starts(A, W, lane) : 0;
    k3w_clear(D.gs, (ngD + 1) / 2, lane);
    k3w_clear(A.gs, (ngA + 1) / 2, lane);
    __syncwarp();
    k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
        int gd = -1, ga = -1;
        if (ok) {
            const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
            gd = k3w_group(D, y - x + moff);
            if (want6) ga = k3w_group(A, y + x);
        }
        k3w_count_add16(D.gs, gd, lane);
        if (want6) k3w_count_add16(A.gs, ga, lane);
    });
    __syncwarp();
    const uint32_t gmax1 = want10 ? k3w_gmax(D.gs, ngD, lane) : 0u;
    // one pass: the a6 verdict of every dot, and the first W10 verdict (both read only group sizes)
    uint32_t n6 = 0, sum6 = 0, k1 = 0, c10 = 0;
    int dmin = 0x7FFFFFFF, dmax = -0x7FFFFFFF;
    k3w_each(hits, H, lane, [&](const uint2& h, uint32_t i, bool ok) {
        if (ok) {
            const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
            const uint32_t szD = k3w_gsize(D, y - x + moff);
            if (want6) {
                const bool keep = szD > 10u || k3w_gsize(A, y + x) > 10u;
                if (flags) hits[i].y = (uint32_t)y | (keep ? HIT_F_CLEAN : 0u);
                if (keep) { ++n6; sum6 += (uint32_t)abs(x - y); dmin = min(dmin, y - x); dmax = max(dmax, y - x); }
            }
            if (want10 && k3_a7_keep(szD, gmax1)) {
                ++k1;
                if (x > 0 && 25 * abs(x - y) < 4 * x) ++c10;        // abs(float(x-y)/float(x)) < 0.16, :732-733
            }
        }
    });
    if (want6) {
        st.n6 = __reduce_add_sync(K3W_FULL, n6);
        st.sumabs = k3w_sum64((unsigned long long)sum6);
        st.dmin = __reduce_min_sync(K3W_FULL, dmin);
        st.dmax = __reduce_max_sync(K3W_FULL, dmax);
    }
    if (!want10) return;
    const uint32_t kept1 = __reduce_add_sync(K3W_FULL, k1);
    uint32_t extra = 0;
    if (kept1 < H) {                                                // second clustering, on y+x, over the dots the first did not keep
        __syncwarp();
        k3w_clear(A.bits, W, lane);
        __syncwarp();
        k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
            if (ok) {
                const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
                if (!k3_a7_keep(k3w_gsize(D, y - x + moff), gmax1)) k3w_mark(A.bits, y + x);
            }
        });
        __syncwarp();
        const int ng2 = k3w_starts(A, W, lane);
        k3w_clear(A.gs, (ng2 + 1) / 2, lane);
        __syncwarp();
        k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
            int ga = -1;
            if (ok) {
                const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
                if (!k3_a7_keep(k3w_gsize(D, y - x + moff), gmax1)) ga = k3w_group(A, y + x);
            }
            k3w_count_add16(A.gs, ga, lane);
        });
        __syncwarp();
        const uint32_t gmax2 = k3w_gmax(A.gs, ng2, lane);
        k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
            if (ok) {
                const int x = (int)h.x, y = (int)(h.y & HIT_Y_MASK);
                if (!k3_a7_keep(k3w_gsize(D, y - x + moff), gmax1) && k3_a7_keep(k3w_gsize(A, y + x), gmax2)) {
                    ++extra;
                    if (x > 0 && 25 * abs(x - y) < 4 * x) ++c10;
                }
            }
        });
        extra = __reduce_add_sync(K3W_FULL, extra);
    }
    st.nw = kept1 + extra;
    st.c10 = __reduce_add_sync(K3W_FULL, c10);
}

// dis_to_diagnal_most_abundant_defined + eu_dis_dir_calcu on the clean dots (:582-591, 248-251, 710-722); see
// k3_redef_stat.  Needs the HIT_F_CLEAN flags and the range [dmin, dmax] of y-x over the clean dots from k3w_clean.
// `aux` = the team's group-size area (free again), at least nb/10 + 5 words.
__device__ __forceinline__ double k3w_redef_stat(const PlotView& v, uint32_t* aux, int lane, int mn1, int mx1) {
    const uint2* hits = v.hits;
    const uint32_t H = v.H;
    const int rg1 = mx1 - mn1;
    const float inv1 = rg1 > 0 ? 1.0f / (float)rg1 : 0.0f;
    auto modal = [&](int& sel, uint32_t& cnt) {                    // find_longest_list: exactly one modal bin?
        const uint32_t c = lane < 11 ? aux[lane] : 0u;
        const uint32_t best = __reduce_max_sync(K3W_FULL, c);
        const unsigned who = __ballot_sync(K3W_FULL, lane < 11 && c == best);
        sel = (__popc(who) == 1) ? (__ffs(who) - 1) : -1;
        cnt = best;
    };
    if (lane < 11) aux[lane] = 0u;
    __syncwarp();
    k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
        int b = -1;
        if (ok && (h.y & HIT_F_CLEAN)) b = k3w_bin11((int)(h.y & HIT_Y_MASK) - (int)h.x, mn1, rg1, inv1);
        k3w_count_add32(aux, b, lane);
    });
    __syncwarp();
    int b1; uint32_t c1;
    modal(b1, c1);
    long long icpt2 = 0;
    if (b1 >= 0) {                                                  // level 2 inside the modal bin
        int lmin = 0x7FFFFFFF, lmax = -0x7FFFFFFF;
        k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
            if (ok && (h.y & HIT_F_CLEAN)) {
                const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
                if (k3w_bin11(d, mn1, rg1, inv1) == b1) { lmin = min(lmin, d); lmax = max(lmax, d); }
            }
        });
        const int mn2 = __reduce_min_sync(K3W_FULL, lmin), rg2 = __reduce_max_sync(K3W_FULL, lmax) - mn2;
        const float inv2 = rg2 > 0 ? 1.0f / (float)rg2 : 0.0f;
        __syncwarp();
        if (lane < 11) aux[lane] = 0u;
        __syncwarp();
        k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
            int b = -1;
            if (ok && (h.y & HIT_F_CLEAN)) {
                const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
                if (k3w_bin11(d, mn1, rg1, inv1) == b1) b = k3w_bin11(d, mn2, rg2, inv2);
            }
            k3w_count_add32(aux, b, lane);
        });
        __syncwarp();
        int b2; uint32_t c;
        modal(b2, c);
        if (b2 >= 0) {                                              // np.median of that sub-bin
            __syncwarp();
            k3w_clear(aux, rg2 + 1, lane);
            __syncwarp();
            k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
                int t = -1;
                if (ok && (h.y & HIT_F_CLEAN)) {
                    const int d = (int)(h.y & HIT_Y_MASK) - (int)h.x;
                    if (k3w_bin11(d, mn1, rg1, inv1) == b1 && k3w_bin11(d, mn2, rg2, inv2) == b2) t = d - mn2;
                }
                k3w_count_add32(aux, t, lane);
            });
            __syncwarp();
            // the values of rank (c-1)/2 and c/2 on the running prefix of the value counts
            const uint32_t r0 = (c - 1) / 2, r1 = c / 2;
            uint32_t run = 0;
            int t0 = -1, t1 = -1;
            #pragma unroll 1
            for (int base = 0; base <= rg2 && t1 < 0; base += 32) {
                const int t = base + lane;
                const uint32_t cv = t <= rg2 ? aux[t] : 0u;
                uint32_t inc = cv;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(K3W_FULL, inc, o);
                    if (lane >= o) inc += u;
                }
                const uint32_t hi = run + inc, lo = hi - cv;
                const unsigned f0 = __ballot_sync(K3W_FULL, lo <= r0 && hi > r0);
                const unsigned f1 = __ballot_sync(K3W_FULL, lo <= r1 && hi > r1);
                if (f0) t0 = base + __ffs(f0) - 1;
                if (f1) t1 = base + __ffs(f1) - 1;
                run += __shfl_sync(K3W_FULL, inc, 31);
            }
            icpt2 = (long long)(t0 + mn2) + (long long)(t1 + mn2);
        }
    }
    __syncwarp();
    // eu_dis_dir_calcu on (x + intercept, y) in doubled integers: A = 2(x'-y), B = 2x'; a dot counts when |A/B| > 0.1
    // (x' == 0: |A/2| / 1 > 0.1, i.e. A != 0)
    const int ic = (int)icpt2;
    long long lsum = 0; uint32_t lcnt = 0;
    k3w_each(hits, H, lane, [&](const uint2& h, uint32_t, bool ok) {
        if (ok && (h.y & HIT_F_CLEAN)) {
            const int B = 2 * (int)h.x + ic;
            const int A = B - 2 * (int)(h.y & HIT_Y_MASK);
            const bool far = (B == 0) ? (A != 0) : (10ll * (long long)abs(A) > (long long)abs(B));
            if (far) { lsum += A; ++lcnt; }
        }
    });
    const long long s64 = (long long)k3w_sum64((unsigned long long)lsum);
    const uint32_t cnt = __reduce_add_sync(K3W_FULL, lcnt);
    return cnt == 0 ? 0.0001 : fabs(((double)s64 * 0.5) / (double)cnt);
}

// The part of the gate ladders that needs only the hit counts: what an opinion WILL clean unless a span gate (which needs
// pass 0) stops it.  Same expressions as in k3w_gates, so that plan.a6 / plan.w10 there imply the same here.
__device__ __forceinline__ void k3w_hgates(int mode, uint32_t Hr_, uint32_t Ha_, int len_ref, int len_alt, bool& a6, bool& w10) {
    const double Hr = (double)Hr_, Ha = (double)Ha_, Lr = (double)len_ref, La = (double)len_alt;
    a6 = false; w10 = false;
    if (mode == 0) a6 = (Hr_ > 2 && Ha_ > 2) && (Hr / fmin(Lr, La) > 0.1);
    else if (mode == 1) w10 = fmax(Hr / Lr, Ha / La) > 0.1;
    else a6 = Hr / Lr > 0.1 && Ha / La > 0.1;
}

struct K3WPlan {           // what the gates of one opinion ask for
    bool a6, w10;
    double a, b;           // the pair when the gates decide alone
};

// gate ladders of the three modes (:187-203, 280-294, 244-257) up to the point where the dots must be cleaned
__device__ __forceinline__ K3WPlan k3w_gates(int mode, uint32_t Hr_, uint32_t Ha_, int len_ref, int len_alt,
                                             int minr, int maxr, int mina, int maxa) {
    K3WPlan g{false, false, 0.0, 0.0};
    const double Hr = (double)Hr_, Ha = (double)Ha_, Lr = (double)len_ref, La = (double)len_alt;
    const double span_r = (double)((long long)maxr - minr) / Lr, span_a = (double)((long long)maxa - mina) / La;
    if (mode == 0) {
        if (!(Hr_ > 2 && Ha_ > 2)) return g;
        if (!(Hr / fmin(Lr, La) > 0.1)) return g;
        const bool rs = span_r > 0.6, as = span_a > 0.6;
        if (rs && as) g.a6 = true;
        else if (rs) { g.a = 1.1; g.b = 2.1; }
        else if (as) { g.a = 2.1; g.b = 1.1; }
    } else if (mode == 1) {
        if (!(fmax(Hr / Lr, Ha / La) > 0.1)) return g;
        g.w10 = true;
    } else {
        if (!(Hr / Lr > 0.1 && Ha / La > 0.1)) return g;
        if (!(span_r > 0.7 && span_a > 0.7)) return g;
        g.a6 = true;
    }
    return g;
}

#ifndef K3W_MINB
#define K3W_MINB 8
#endif

__global__ void __launch_bounds__(32 * K3W_TEAMS, K3W_MINB)
k3w_score_reads(const K3Params p)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    const int lane = threadIdx.x & 31, team = threadIdx.x >> 5;
    K3WSet D, A;
    {
        const int wb = k3w_words_bits(p.nb_cap), wp = k3w_words_pref(p.nb_cap), wg = k3w_words_gs(p.nb_cap);
        uint32_t* base = s_dyn + (size_t)team * k3w_scratch_words(p.nb_cap);
        D.bits = base; A.bits = base + wb;
        D.pref = reinterpret_cast<uint16_t*>(base + 2 * wb); A.pref = reinterpret_cast<uint16_t*>(base + 2 * wb + wp);
        D.gs = base + 2 * wb + 2 * wp; A.gs = D.gs + wg;
    }
    uint32_t* aux = D.gs;                                           // gsD | gsA, contiguous: nb/10 + 5 words at least

    while (true) {
        int it = 0;
        if (lane == 0) it = (int)atomicAdd(p.queue, 1u);
        it = __shfl_sync(K3W_FULL, it, 0);
        if (it >= p.n_ids) break;
        const int tin = p.task_ids[it];
        const int tix = p.out_ids ? p.out_ids[it] : tin;
        const Task t = p.tasks[tin];
        PlotView pv[4];
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (t.plot[i] >= 0) {
                const Plot pl = p.plots[t.plot[i]];
                const uint32_t found = p.cnt[t.plot[i]];            // an overflowed plot counts as empty (see k3_score_reads)
                pv[i].hits = p.hits + pl.hit_off; pv[i].H = found > pl.cap ? 0u : found; pv[i].n = pl.n; pv[i].m = pl.m;
            } else { pv[i].hits = nullptr; pv[i].H = 0; pv[i].n = 0; pv[i].m = 0; }
        }
        const bool bad = p.op_status[t.read_op] != 0;
        double ea = 0, eb = 0, fa = 0, fb = 0;                      // the pairs of the two opinions
        unsigned long long cs[4] = {0, 0, 0, 0};
        if (!bad) {
            const bool two = t.mode == 3;                           // the simple-DEL rule asks ABS, then W10
            const bool same = two && t.plot[2] == t.plot[0] && t.plot[3] == t.plot[1];
            int minx[4] = {0, 0, 0, 0}, maxx[4] = {0, 0, 0, 0};
            const int n_own = (two && !same) ? 4 : 2;               // plots this task looks at on their own
            const int mode0 = two ? 0 : t.mode;
            // what to clean follows from the hit counts; the span gates (pass 0, done inside the first cleaning pass) come after
            bool h6, h10, hx, h10b = false;
            k3w_hgates(mode0, pv[0].H, pv[1].H, t.len_ref, t.len_alt, h6, h10);
            if (two) k3w_hgates(1, pv[2].H, pv[3].H, t.len_ref, t.len_alt, hx, h10b);
            K3WStat sx[4];
            const bool fused = same && h10b;                        // both opinions look at the same two plots: one cleaning call
            #pragma unroll 1
            for (int e = 0; e < n_own; ++e) {
                const bool first = e < 2;
                k3w_clean(pv[e], D, A, lane, first && h6, first ? (h10 || fused) : h10b, first && mode0 == 2, sx[e], minx[e], maxx[e], cs[e]);
                __syncwarp();
            }
            if (same) { minx[2] = minx[0]; maxx[2] = maxx[0]; minx[3] = minx[1]; maxx[3] = maxx[1]; cs[2] = cs[0]; cs[3] = cs[1]; sx[2] = sx[0]; sx[3] = sx[1]; }
            const K3WPlan g0 = k3w_gates(mode0, pv[0].H, pv[1].H, t.len_ref, t.len_alt, minx[0], maxx[0], minx[1], maxx[1]);
            K3WPlan g1{false, false, 0.0, 0.0};
            if (two) g1 = k3w_gates(1, pv[2].H, pv[3].H, t.len_ref, t.len_alt, minx[2], maxx[2], minx[3], maxx[3]);
            ea = g0.a; eb = g0.b;
            const K3WStat* s0 = sx; const K3WStat* s1 = sx + 2;
            if (g0.a6 && s0[0].n6 > 0 && s0[1].n6 > 0) {
                if (mode0 == 0) {
                    ea = (double)s0[0].sumabs / (double)s0[0].n6;    // np.mean of exact integers
                    eb = (double)s0[1].sumabs / (double)s0[1].n6;
                } else {
                    double dir[2];
                    #pragma unroll 1
                    for (int w = 0; w < 2; ++w) {
                        dir[w] = k3w_redef_stat(pv[w], aux, lane, s0[w].dmin, s0[w].dmax);
                        __syncwarp();
                    }
                    ea = dir[0]; eb = dir[1];
                }
            } else if (g0.w10 && s0[0].nw > 0 && s0[1].nw > 0) {
                ea = (double)s0[1].c10; eb = (double)s0[0].c10;      // swapped on purpose (:290)
            }
            if (two && g1.w10 && s1[0].nw > 0 && s1[1].nw > 0) { fa = (double)s1[1].c10; fb = (double)s1[0].c10; }
        }
        if (lane == 0) {
            const bool va = (ea != 0.0) && (eb != 0.0), vb_ = (fa != 0.0) && (fb != 0.0);   // `if not 0 in pair`
            double score = 0.0; uint8_t status = 0;
            if (bad) status = 2;
            else if (t.mode == 3) {                                 // simple-DEL rule, :1718-1726
                const double s1 = 1.0 - eb / ea, s2 = 1.0 - fb / fa;
                if (va && vb_) { score = (s2 < s1) ? s2 : s1; status = 1; }
                else if (va) { score = s1; status = 1; }
                else if (vb_) { score = s2; status = 1; }
            } else if (va) { score = 1.0 - eb / ea; status = 1; }
            p.task_score[tix] = score;
            p.task_status[tix] = status;
            p.task_stat[4 * tix + 0] = ea; p.task_stat[4 * tix + 1] = eb;
            p.task_stat[4 * tix + 2] = fa; p.task_stat[4 * tix + 3] = fb;
            #pragma unroll
            for (int i = 0; i < 4; ++i) { p.task_hits[4 * tix + i] = pv[i].H; p.task_hitsum[4 * tix + i] = cs[i]; }
        }
        __syncwarp();
    }
}

}  // namespace vb
