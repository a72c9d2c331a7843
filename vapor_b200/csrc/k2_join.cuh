// Kernel 2, join variant: the recurrence plot as a radix-partitioned join instead of an all-pairs tile sweep.
//
// Reference semantics are those of k2_tile.cuh: the dot (i, j) exists when structure k-mer i equals read k-mer j
// forward or reverse-complemented (vapor_vali/Simple_function.pyx:964-979), twice when the read k-mer is its own
// reverse complement (:959-960, :1419-1421).  The reference finds the dots with a Python dict keyed by the k-mer
// string (kmerhits, :951-983): it never looks at a cell whose k-mers differ.  The tile kernel looks at all n x m
// cells; this kernel looks only at cells whose canonical words share their top `bits` bits:
//
//   kernel 1b (k1b_build_tables): every structure-side operand chunk becomes a table -- its valid words
//       counting-sorted by bucket key, the position of each sorted word, the bucket offsets, and a membership bitmap
//       over 4 more key bits than the buckets use (common.cuh, TabChunk);
//   k2_join_match: one CTA stages one table in shared memory with ONE TMA bulk copy (cp.async.bulk + mbarrier)
//       and streams the read words of every plot that uses the table past it, straight from HBM/L2 in natural order
//       (coalesced, four loads in flight per lane); every warp owns a contiguous share of the item's read words.
//       Step 1, branch-free and converged: a lane tests its word against the membership bitmap (one shared-memory
//       load).  77 % of the read words of a 15 %-error read match nothing, and 9 in 10 of those leave here.  The
//       survivors are compacted (ballot + popc) into a per-warp list in shared memory.
//       Step 2, whenever the list holds 32 survivors: one survivor per lane scans its bucket (two 16-bit offsets, 1-2
//       entries on average) remembering up to two equal entries in registers -- no side effects inside the
//       divergent loop -- and the warp parks the matched cells in its hit queue by ballot + popc (no atomics).
//       Rounds are full warps by construction, the first version of this kernel ran its bucket scans with 17 of 32
//       lanes active and spent 4.2 warp instructions per read word; this one spends about 1.
//       The queue is emitted with k2_flush of k2_tile.cuh (confirmation of hashed words, multiplicity of
//       palindromes, QC counters, one global atomic per batch) -- the hit set is identical to the tile kernel's and
//       to the reference's, only its order differs.
//
// Cells evaluated = sum over the surviving read words of the size of their bucket; the kernel counts them
// (K2JParams::evaluated) so that throughput can be quoted on evaluated cells.
#pragma once
#include "common.cuh"
#include "k2_tile.cuh"

namespace vb {

constexpr int K1B_THREADS = 256;
constexpr int K2J_WARPS     = 8;                // warps per CTA (tables up to 44 KB: 4 CTAs per SM)
#ifndef K2J_WARPS_BIG_N
#define K2J_WARPS_BIG_N 12
#endif
constexpr int K2J_WARPS_BIG = K2J_WARPS_BIG_N;    // ... for the largest tables (2 CTAs per SM by shared memory): 12 warps at 85 registers beat 16 at 64 (measured, -2 %)
constexpr int K2J_UNROLL  = 4;                  // read words per lane and step (loads in flight)
#ifndef K2J_PLOTS_N
#define K2J_PLOTS_N 24
#endif
#ifndef K2J_WORDS_N
#define K2J_WORDS_N 65536
#endif
constexpr int K2J_PLOTS_PER_ITEM = K2J_PLOTS_N;      // most plots one CTA streams past its table ...
constexpr int K2J_WORDS_PER_ITEM = K2J_WORDS_N;      // ... and about how many read words: items of similar length

__host__ __device__ __forceinline__ uint32_t k2j_key(uint32_t word, int bits) {
    return (word & 0x3FFFFFFFu) >> (30 - bits);
}

// ---- kernel 1b: counting sort of one table chunk ---------------------------------------------------------
__global__ void __launch_bounds__(K1B_THREADS)
k1b_build_tables(const TabChunk* __restrict__ chunks, const Operand* __restrict__ ops,
                 const uint32_t* __restrict__ hash, uint8_t* __restrict__ table)
{
    __shared__ uint32_t s_cnt[(1 << K2J_MAX_BITS) + 1];
    __shared__ uint32_t s_bm[(1 << 16) / 32];                    // membership bitmap over the top fbits bits
    __shared__ uint32_t s_wtot[K1B_THREADS / 32];
    const TabChunk c = chunks[blockIdx.x];
    const int NB = 1 << c.bits;
    const int fbits = k2j_fbits(c.len), FW = (1 << fbits) / 32;
    const int tid = threadIdx.x;
    for (int q = tid; q <= NB; q += K1B_THREADS) s_cnt[q] = 0u;
    for (int q = tid; q < FW; q += K1B_THREADS) s_bm[q] = 0u;
    __syncthreads();
    const uint32_t* w = hash + ops[c.op].hash_off + c.pos0;
    for (int i = tid; i < c.len; i += K1B_THREADS) {
        const uint32_t word = w[i];
        if (word <= H_MAX_VALID) {
            atomicAdd(&s_cnt[k2j_key(word, c.bits)], 1u);
            const uint32_t fine = k2j_key(word, fbits), bit = 1u << (fine & 31u);
            if (!(s_bm[fine >> 5] & bit)) atomicOr(&s_bm[fine >> 5], bit);
        }
    }
    __syncthreads();
    // exclusive scan of s_cnt[0 .. NB): every thread owns NB / 256 consecutive counters (NB >= 256)
    {
        const int per = NB / K1B_THREADS;
        const int b0 = tid * per;
        uint32_t sum = 0;
        for (int q = 0; q < per; ++q) sum += s_cnt[b0 + q];
        uint32_t inc = sum;
        const int lane = tid & 31, warp = tid >> 5;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t run = inc - sum;
        for (int x = 0; x < warp; ++x) run += s_wtot[x];
        for (int q = 0; q < per; ++q) { const uint32_t v = s_cnt[b0 + q]; s_cnt[b0 + q] = run; run += v; }
        if (tid == K1B_THREADS - 1) s_cnt[NB] = run;              // valid words in the chunk
    }
    __syncthreads();
    const int lp = k2j_lp(c.len);
    uint32_t* tw = reinterpret_cast<uint32_t*>(table + c.blob_off);
    uint16_t* tp = reinterpret_cast<uint16_t*>(table + c.blob_off + 4 * (size_t)lp);
    uint16_t* toff = reinterpret_cast<uint16_t*>(table + c.blob_off + 6 * (size_t)lp);
    const uint32_t total = s_cnt[NB];
    uint32_t* tbm = reinterpret_cast<uint32_t*>(table + c.blob_off + 6 * (size_t)lp + 2 * (size_t)(NB + 8));
    for (int q = tid; q < FW; q += K1B_THREADS) tbm[q] = s_bm[q];
    for (int q = tid; q < NB + 8; q += K1B_THREADS) toff[q] = (uint16_t)(q <= NB ? s_cnt[q] : total);
    for (int i = (int)total + tid; i < lp; i += K1B_THREADS) { tw[i] = H_STRUCT_INVALID; tp[i] = 0; }   // defined padding
    __syncthreads();
    for (int i = tid; i < c.len; i += K1B_THREADS) {
        const uint32_t word = w[i];
        if (word <= H_MAX_VALID) {
            const uint32_t slot = atomicAdd(&s_cnt[k2j_key(word, c.bits)], 1u);
            tw[slot] = word;
            tp[slot] = (uint16_t)i;
        }
    }
}

struct K2JParams {
    const JoinItem* items;
    const int32_t* jplots;       // plot indices (into `plots`) grouped by structure operand
    const TabChunk* chunks;
    const Plot* plots;           // plots of this launch's wave
    const Operand* ops;
    const uint32_t* hash;
    const uint8_t* code;
    const uint8_t* table;
    uint2* hits;
    uint32_t* cnt;               // hits found per plot (may exceed cap)
    uint32_t* overflow;          // set when any plot exceeded its capacity
    uint32_t* qc;                // QC counters of PLOT_QC plots (or null)
    unsigned long long* evaluated;   // += word compares done
};

#ifndef K2J_MINB
#define K2J_MINB 3                               // 3 CTAs of 8 warps per SM at 85 registers: no spills in the probe loop; 4 at 64 registers
#endif                                          // spilled the words in flight and was 3 % slower (measured)
constexpr int K2J_QCAP = 96;                    // per-warp queue of matched cells awaiting emission
constexpr int K2J_LCAP = 160;                   // per-warp list of read words that passed the filter (< 32 carried + 128 new)

__device__ __forceinline__ uint32_t k2j_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t k2j_lds16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (uint32_t)v; }

struct K2JPlot {            // what the warps need to know about one plot of the item (shared memory, filled once per CTA)
    K2Strip st;
    const uint32_t* rw;     // the read's k-mer words
    int n, m, xoff, nblk;   // read k-mers, structure k-mers after the cut, structure coordinate of table position 0, 128-word blocks
};

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, (K2J_MINB * K2J_WARPS) / WARPS)
k2_join_match(const K2JParams p)
{
    extern __shared__ __align__(128) uint8_t s_blob[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint2 s_queue[WARPS][K2J_QCAP];
    __shared__ __align__(8) uint2 s_list[WARPS][K2J_LCAP];
    __shared__ __align__(8) K2JPlot s_plot[K2J_PLOTS_PER_ITEM];
    __shared__ unsigned long long s_eval;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const JoinItem item = p.items[blockIdx.x];
    const TabChunk ch = p.chunks[item.chunk];
    if (threadIdx.x == 0) {
        s_eval = 0ull;
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&s_bar, (uint32_t)ch.blob_bytes);
        tma_bulk_g2s(s_blob, p.table + ch.blob_off, (uint32_t)ch.blob_bytes, &s_bar);
    }
    // the item's plots, one thread each: three dependent global loads per plot happen here, once and in parallel, instead of
    // in front of every warp's visit of the plot
    const int n_plots = item.jp_end - item.jp_begin;
    if ((int)threadIdx.x < n_plots) {
        const int pid = p.jplots[item.jp_begin + threadIdx.x];
        const Plot pl = p.plots[pid];
        const Operand opr = p.ops[pl.read_op];
        const Operand ops_ = p.ops[pl.struct_op];
        K2JPlot& q = s_plot[threadIdx.x];
        q.st.code_read = p.code + opr.code_off; q.st.code_struct = p.code + ops_.code_off + pl.miss;
        q.st.cnt = p.cnt + pid; q.st.hits = p.hits + pl.hit_off; q.st.overflow = p.overflow; q.st.cap = pl.cap; q.st.k = opr.k; q.st.swap = false;
        q.st.qc = (pl.kind & PLOT_QC) ? p.qc + pl.hit_off * QC_WORDS : nullptr;
        q.rw = p.hash + opr.hash_off;
        q.n = pl.n; q.m = pl.m; q.xoff = ch.pos0 - pl.miss;
        q.nblk = pl.m > 0 ? (pl.n + 32 * K2J_UNROLL - 1) / (32 * K2J_UNROLL) : 0;
    }
    __syncthreads();                                     // the barrier is initialised before anyone waits on it; s_plot is filled
    const int lp = k2j_lp(ch.len);
    const uint32_t a_tw = smem_u32(s_blob);              // sorted words
    const uint32_t a_tp = a_tw + 4u * (uint32_t)lp;      // their positions
    const uint32_t a_off = a_tw + 6u * (uint32_t)lp;     // bucket offsets
    const uint32_t a_bm = a_off + 2u * (uint32_t)((1 << ch.bits) + 8);   // membership bitmap
    const int shift = 30 - ch.bits, fshift = 30 - k2j_fbits(ch.len);
    uint2* queue = s_queue[warp];
    uint2* list = s_list[warp];
    uint32_t evaluated = 0;
    int ln = 0, qn = 0;                                  // survivors listed, cells queued (warp-uniform)
    bool waited = false;

    // every warp owns a contiguous share of the item's 128-word read blocks
    int total = 0;
    for (int j = 0; j < n_plots; ++j) total += s_plot[j].nblk;
    const int w_lo = (int)(((long long)total * warp) / WARPS), w_hi = (int)(((long long)total * (warp + 1)) / WARPS);
    int blk0 = 0;                                        // read blocks of the item's earlier plots
    for (int j = 0; j < n_plots; ++j) {
        const K2JPlot& pq = s_plot[j];
        const int nblk = pq.nblk;
        const int b_lo = max(w_lo - blk0, 0), b_hi = min(w_hi - blk0, nblk);
        blk0 += nblk;
        if (b_lo >= b_hi) continue;                      // warp-uniform
        const K2Strip& st = pq.st;
        const uint32_t* rw = pq.rw;
        const int n = pq.n, xoff = pq.xoff;
        uint32_t r[K2J_UNROLL], rn[K2J_UNROLL];          // this block's words, the next block's (in flight while this one is probed)
        #pragma unroll
        for (int u = 0; u < K2J_UNROLL; ++u) { const int i = b_lo * 32 * K2J_UNROLL + 32 * u + lane; rn[u] = i < n ? rw[i] : H_READ_PAD; }
        for (int b = b_lo; b < b_hi; ++b) {
            const int base = b * 32 * K2J_UNROLL + lane;
            #pragma unroll
            for (int u = 0; u < K2J_UNROLL; ++u) r[u] = rn[u];
            if (b + 1 < b_hi) {
                #pragma unroll
                for (int u = 0; u < K2J_UNROLL; ++u) { const int i = base + 32 * K2J_UNROLL + 32 * u; rn[u] = i < n ? rw[i] : H_READ_PAD; }
            }
            if (!waited) { mbar_wait(&s_bar, 0); waited = true; }
            // step 1: membership test, survivors compacted into the list
            #pragma unroll
            for (int u = 0; u < K2J_UNROLL; ++u) {
                const uint32_t word = r[u];
                const uint32_t fine = (word & 0x3FFFFFFFu) >> fshift;
                const bool pass = word <= H_MAX_VALID && ((k2j_lds32(a_bm + 4u * (fine >> 5)) >> (fine & 31u)) & 1u);
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, pass);
                if (pass) list[ln + __popc(pm & lt)] = make_uint2(word, (uint32_t)(base + 32 * u));
                ln += __popc(pm);
            }
            // step 2: full rounds of 32 survivors (the plot's last block also takes the remainder)
            const bool last = b + 1 == b_hi;
            while (ln >= 32 || (last && ln > 0)) {
                const int n_act = min(ln, 32);
                __syncwarp();
                const bool have = lane < n_act;
                ln -= n_act;
                const uint2 e = have ? list[ln + lane] : make_uint2(H_READ_PAD, 0u);
                __syncwarp();                            // everybody holds its survivor before the list is written again
                const uint32_t word = e.x;
                uint32_t idx = 0, end = 0;
                if (have) {
                    const uint32_t ka = a_off + 2u * ((word & 0x3FFFFFFFu) >> shift);
                    idx = k2j_lds16(ka); end = k2j_lds16(ka + 2u);
                }
                evaluated += end - idx;
                int m0 = -1, m1 = -1, m2 = -1;           // first two equal entries; m2: where a third one sits (repeats)
                for (; idx < end; ++idx) {
                    if (k2j_lds32(a_tw + 4u * idx) == word) {
                        if (m0 < 0) m0 = (int)idx; else if (m1 < 0) m1 = (int)idx; else { m2 = (int)idx; break; }
                    }
                }
                int x0 = -1, x1 = -1;
                if (m0 >= 0) x0 = xoff + (int)k2j_lds16(a_tp + 2u * (uint32_t)m0);      // negative: before the cut, not a cell of the plot
                if (m1 >= 0) x1 = xoff + (int)k2j_lds16(a_tp + 2u * (uint32_t)m1);
                const unsigned b0 = __ballot_sync(0xFFFFFFFFu, x0 >= 0), b1 = __ballot_sync(0xFFFFFFFFu, x1 >= 0);
                if (x0 >= 0) queue[qn + __popc(b0 & lt)] = make_uint2((uint32_t)x0 | (word & 0xC0000000u), e.y);
                qn += __popc(b0);
                if (x1 >= 0) queue[qn + __popc(b1 & lt)] = make_uint2((uint32_t)x1 | (word & 0xC0000000u), e.y);
                qn += __popc(b1);
                if (__any_sync(0xFFFFFFFFu, m2 >= 0)) {  // rare: a word that sits three or more times in its bucket
                    uint32_t i2 = m2 >= 0 ? (uint32_t)m2 : end;
                    while (true) {
                        int x = -1;
                        while (i2 < end && x < 0) {
                            const uint32_t ii = i2++;
                            if (k2j_lds32(a_tw + 4u * ii) == word) x = xoff + (int)k2j_lds16(a_tp + 2u * ii);
                        }
                        const unsigned bb = __ballot_sync(0xFFFFFFFFu, x >= 0);
                        if (bb == 0u) break;
                        if (qn + 32 > K2J_QCAP) { __syncwarp(); k2_flush(st, queue, qn, lane); qn = 0; __syncwarp(); }
                        if (x >= 0) queue[qn + __popc(bb & lt)] = make_uint2((uint32_t)x | (word & 0xC0000000u), e.y);
                        qn += __popc(bb);
                    }
                }
                if (qn > K2J_QCAP - 64 || (last && ln == 0 && qn > 0)) { __syncwarp(); k2_flush(st, queue, qn, lane); qn = 0; __syncwarp(); }
            }
        }
        if (qn > 0) { __syncwarp(); k2_flush(st, queue, qn, lane); qn = 0; __syncwarp(); }   // (a last block without survivors)
    }
    if (!waited) mbar_wait(&s_bar, 0);                   // nobody leaves while the bulk copy may still be landing
    unsigned long long ev = evaluated;
    #pragma unroll
    for (int o = 16; o; o >>= 1) ev += __shfl_xor_sync(0xFFFFFFFFu, ev, o);
    if (lane == 0 && ev) atomicAdd(&s_eval, ev);
    __syncthreads();
    if (threadIdx.x == 0 && s_eval && p.evaluated) atomicAdd(p.evaluated, s_eval);
}

}  // namespace vb
