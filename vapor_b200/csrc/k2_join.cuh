// Kernel 2, join variant: the recurrence plot as a radix-partitioned join instead of an all-pairs tile sweep.
//
// Reference semantics are those of k2_tile.cuh: the dot (i, j) exists when structure k-mer i equals read k-mer j
// forward or reverse-complemented (vapor_vali/Simple_function.pyx:964-979), twice when the read k-mer is its own
// reverse complement (:959-960, :1419-1421).  The reference finds the dots with a Python dict keyed by the k-mer
// string (kmerhits, :951-983): it never looks at a cell whose k-mers differ.  The tile kernel looks at all n x m
// cells; this kernel looks only at cells whose canonical words share their top `bits` bits:
//
//   kernel 1b (k1b_build_tables): every structure-side operand chunk becomes a table -- its valid words
//       counting-sorted by bucket key, the position of each sorted word, the bucket offsets (common.cuh, TabChunk);
//   k2_join_match: one CTA stages one table in shared memory with ONE TMA bulk copy (cp.async.bulk + mbarrier)
//       and streams the read words of every plot that uses the table past it, 128 read words per warp and step,
//       straight from HBM/L2 in natural order (coalesced, four loads in flight per lane).  A lane looks its
//       word's bucket up (two 16-bit offsets) and compares the word with the bucket's entries (1.3 on average), each
//       lane on its own; a matched cell takes a slot of the per-warp queue with a shared-memory atomic, and the warp
//       emits the queue with k2_flush of k2_tile.cuh (confirmation of hashed words, multiplicity of palindromes, QC
//       counters, one global atomic per batch) -- the hit set is identical to the tile kernel's and to the
//       reference's, only its order differs.
//
// Cells evaluated = sum over read words of the size of their bucket (about n * (1 + m / 2^bits) per plot instead of
// n * m); the kernel counts them (K2JParams::evaluated) so that throughput can be quoted on evaluated cells.
#pragma once
#include "common.cuh"
#include "k2_tile.cuh"

namespace vb {

constexpr int K1B_THREADS = 256;
constexpr int K2J_WARPS     = 8;                // warps per CTA (tables up to 44 KB: 4 CTAs per SM)
constexpr int K2J_WARPS_BIG = 16;               // ... for the largest tables (2 CTAs per SM by shared memory: keep 32 warps per SM)
constexpr int K2J_UNROLL  = 4;                  // read words per lane and step (loads in flight)
constexpr int K2J_PLOTS_PER_ITEM = 8;           // most plots one CTA streams past its table ...
constexpr int K2J_WORDS_PER_ITEM = 24576;       // ... and about how many read words: items of similar length

__host__ __device__ __forceinline__ uint32_t k2j_key(uint32_t word, int bits) {
    return (word & 0x3FFFFFFFu) >> (30 - bits);
}

// ---- kernel 1b: counting sort of one table chunk ---------------------------------------------------------
__global__ void __launch_bounds__(K1B_THREADS)
k1b_build_tables(const TabChunk* __restrict__ chunks, const Operand* __restrict__ ops,
                 const uint32_t* __restrict__ hash, uint8_t* __restrict__ table)
{
    __shared__ uint32_t s_cnt[(1 << K2J_MAX_BITS) + 1];
    __shared__ uint32_t s_wtot[K1B_THREADS / 32];
    const TabChunk c = chunks[blockIdx.x];
    const int NB = 1 << c.bits;
    const int tid = threadIdx.x;
    for (int q = tid; q <= NB; q += K1B_THREADS) s_cnt[q] = 0u;
    __syncthreads();
    const uint32_t* w = hash + ops[c.op].hash_off + c.pos0;
    for (int i = tid; i < c.len; i += K1B_THREADS) {
        const uint32_t word = w[i];
        if (word <= H_MAX_VALID) atomicAdd(&s_cnt[k2j_key(word, c.bits)], 1u);
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
#define K2J_MINB 4
#endif
constexpr int K2J_QCAP = 96;                    // per-warp queue of matched cells awaiting emission

__device__ __forceinline__ uint32_t k2j_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t k2j_lds16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (uint32_t)v; }

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, (K2J_MINB * K2J_WARPS) / WARPS)
k2_join_match(const K2JParams p)
{
    extern __shared__ __align__(128) uint8_t s_blob[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint2 s_queue[WARPS][K2J_QCAP];
    __shared__ uint32_t s_qn[WARPS];
    __shared__ K2Strip s_strip[WARPS];
    __shared__ unsigned long long s_eval;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const JoinItem item = p.items[blockIdx.x];
    const TabChunk ch = p.chunks[item.chunk];
    if (threadIdx.x == 0) {
        s_eval = 0ull;
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&s_bar, (uint32_t)ch.blob_bytes);
        tma_bulk_g2s(s_blob, p.table + ch.blob_off, (uint32_t)ch.blob_bytes, &s_bar);
    }
    if (lane == 0) s_qn[warp] = 0u;
    __syncthreads();                                     // the barrier is initialised before anyone waits on it
    const int lp = k2j_lp(ch.len);
    const uint32_t a_tw = smem_u32(s_blob);              // sorted words
    const uint32_t a_tp = a_tw + 4u * (uint32_t)lp;      // their positions
    const uint32_t a_off = a_tw + 6u * (uint32_t)lp;     // bucket offsets
    const int shift = 30 - ch.bits;
    uint2* queue = s_queue[warp];
    uint32_t* qn = &s_qn[warp];
    K2Strip& st = s_strip[warp];
    uint32_t evaluated = 0;
    bool waited = false;

    int blk0 = 0;                                        // 128-word read blocks of the item's earlier plots
    for (int jp = item.jp_begin; jp < item.jp_end; ++jp) {
        const int pid = p.jplots[jp];
        const Plot pl = p.plots[pid];
        const int nblk = (pl.n + 32 * K2J_UNROLL - 1) / (32 * K2J_UNROLL);
        // blocks are dealt round-robin to the warps across the item's plots: warp w takes block b when (blk0 + b) % W == w
        int b = (warp - blk0 % WARPS + WARPS) % WARPS;
        blk0 += nblk;
        if (b >= nblk || pl.m <= 0) continue;            // warp-uniform
        const Operand opr = p.ops[pl.read_op];
        __syncwarp();
        if (lane == 0) {
            const Operand ops_ = p.ops[pl.struct_op];
            st.code_read = p.code + opr.code_off; st.code_struct = p.code + ops_.code_off + pl.miss;
            st.cnt = p.cnt + pid; st.hits = p.hits + pl.hit_off; st.overflow = p.overflow; st.cap = pl.cap; st.k = opr.k; st.swap = false;
            st.qc = (pl.kind & PLOT_QC) ? p.qc + pl.hit_off * QC_WORDS : nullptr;
        }
        __syncwarp();
        const uint32_t* rw = p.hash + opr.hash_off;
        const int xoff = ch.pos0 - pl.miss;              // structure coordinate of table position 0 after the cut
        for (; b < nblk; b += WARPS) {
            const int base = b * 32 * K2J_UNROLL + lane;
            uint32_t r[K2J_UNROLL];
            #pragma unroll
            for (int u = 0; u < K2J_UNROLL; ++u) r[u] = (base + 32 * u < pl.n) ? rw[base + 32 * u] : H_READ_PAD;
            if (!waited) { mbar_wait(&s_bar, 0); waited = true; }
            #pragma unroll
            for (int u = 0; u < K2J_UNROLL; ++u) {
                const uint32_t word = r[u];
                uint32_t idx = 0, end = 0;
                if (word <= H_MAX_VALID) {
                    const uint32_t ka = a_off + 2u * ((word & 0x3FFFFFFFu) >> shift);
                    idx = k2j_lds16(ka); end = k2j_lds16(ka + 2u);
                }
                evaluated += end - idx;
                // every lane scans its own bucket; a matched cell takes a queue slot with a shared-memory atomic.  A lane that
                // finds the queue full stops at that entry: the warp then empties the queue and the lane carries on.
                while (true) {
                    for (; idx < end; ++idx) {
                        if (k2j_lds32(a_tw + 4u * idx) != word) continue;
                        const int x = xoff + (int)k2j_lds16(a_tp + 2u * idx);
                        if (x < 0) continue;
                        const uint32_t slot = atomicAdd(qn, 1u);
                        if (slot >= (uint32_t)K2J_QCAP) break;
                        queue[slot] = make_uint2((uint32_t)x | (word & 0xC0000000u), (uint32_t)(base + 32 * u));
                    }
                    __syncwarp();                                    // every lane's slots and queue entries are visible
                    const unsigned more = __ballot_sync(0xFFFFFFFFu, idx < end);
                    const uint32_t have = min(*qn, (uint32_t)K2J_QCAP);
                    if (more == 0u && have <= (uint32_t)(K2J_QCAP - 48)) break;
                    __syncwarp();
                    k2_flush(st, queue, (int)have, lane);
                    __syncwarp();
                    if (lane == 0) *qn = 0u;
                    __syncwarp();
                    if (more == 0u) break;
                }
            }
        }
        __syncwarp();
        const uint32_t have = min(*qn, (uint32_t)K2J_QCAP);
        if (have) { k2_flush(st, queue, (int)have, lane); __syncwarp(); if (lane == 0) *qn = 0u; }
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
