// Kernel 2: recurrence-plot tile kernel.
//
// Reference semantics: for every structure k-mer i and read k-mer j the reference emits the
// dot (i, j) when the strings are equal forward or reverse-complemented
// (vapor_vali/Simple_function.pyx:964-979), a self-reverse-complement read k-mer emitting it
// twice (:959-960, :1419-1421).  The reference gets there with a Python dict; this kernel
// evaluates every cell of the n x m plot on 32-bit canonical k-mer words and never materialises
// the matrix: only the sparse hit list reaches HBM.
//
// Decomposition: a *strip* is 32*R consecutive k-mer words of one axis (R per lane, in registers:
// the "rows") against up to K2_TS consecutive words of the other axis staged in shared memory by a
// 1-D TMA bulk copy (cp.async.bulk + mbarrier, one buffer per warp, no block-wide barrier
// anywhere: the "stream").  Each warp of a persistent grid pulls strips from a global queue.
// Equality is symmetric, so either axis can be the rows: the full read chunks of a plot take the
// read k-mers as rows and stream the structure; the last, partial read chunk is tiled transposed
// (structure k-mers as rows, the few left-over read k-mers streamed) when that wastes fewer padded
// cells -- the host decides per plot (PLOT_TAIL_T).
//
// Inner loop, dual pipe.  B200 issues integer compares (ISETP) on the alu pipe and integer
// multiply-adds (IMAD) on the fma pipe, each at 64 lanes/clk/SM (tools/microbench2.cu).  A lane's
// R = NI + 8*NP rows are therefore split: NI rows are compared one ISETP each, and every further
// group of 8 rows is folded into the polynomial P(v) = prod_q (v - r_q) mod 2^32, evaluated for
// each streamed word v by Horner's rule (8 fma-pipe ops) and tested against zero once.
// P(v) == 0 whenever v equals one of the 8 rows (no false negatives); a zero without a match needs
// the factors' trailing zero bits to sum to >= 32 (about 3.5e-5 per evaluation on mixed words) and
// only costs a visit to the exact path.  The streamed word is the FIRST source operand of every
// instruction of the loop: ptxas keeps the order, and the operand-reuse cache then feeds it to
// consecutive ISETP / IMAD without a register-file read -- the register file delivers two operands
// per clock and SM sub-partition, and a 2-read ISETP next to a 3-read IMAD would otherwise cap the
// issue rate at 0.8 instructions/clk (measured: 95 -> 106 cells/clk/SM on the isolated loop).
// A warp vote every 32 streamed k-mers asks whether any lane saw a candidate; a lane that did
// leaves a 4-byte note (block, lane, row group) in a per-warp shared-memory list and the inner
// loop moves on.  The notes are resolved later, one note per lane: the 8 rows of the noted lane
// (fetched by shuffle) against the 32 words of the noted block, first a collective-free scan that
// remembers matching positions, then one parked dot per lane and step.  Parked dots are confirmed
// (hash words with bit 31 set are checked on the code strings) and appended in batches of 32 with
// one atomic.
#pragma once
#include "common.cuh"

namespace vb {

constexpr int K2_D       = 8;                  // rows folded into one polynomial
constexpr int K2_TS      = 2048;               // streamed k-mers per strip
constexpr int K2_WARPS   = 4;                  // warps per CTA
constexpr int K2_THREADS = 32 * K2_WARPS;
constexpr int K2_SBUF    = K2_TS + 40;         // words per warp buffer (alignment shift + vote-block padding)
constexpr int K2_QCAP    = 64;                 // per-warp queue of matched cells awaiting emission
constexpr int K2_NCAP    = 192;                // per-warp list of candidate notes awaiting resolution

// tile variants: (rows by ISETP, row polynomials): 0 = (16,0), 1 = (14,2), 2 = (16,2), 3 = (12,2), 4 = (13,2).
// Variant 4 is the product default: with the exact path and loop overhead also on the alu pipe, 13 + 2 compares
// against 16 multiply-adds per streamed word balances the two pipes best (measured).
constexpr int K2_NVARIANT = 5;
__host__ __device__ constexpr int k2_variant_ni(int v) { return v == 0 ? 16 : (v == 1 ? 14 : (v == 2 ? 16 : (v == 3 ? 12 : 13))); }
__host__ __device__ constexpr int k2_variant_np(int v) { return v == 0 ? 0 : 2; }
__host__ __device__ constexpr int k2_variant_rows(int v) { return 32 * (k2_variant_ni(v) + K2_D * k2_variant_np(v)); }

// Strips of one plot (host and device agree through these two functions).
// Full read chunks first: strip = (read chunk rc, structure chunk cc), rc fastest.  Then the tail chunk
// of t = n % rows read k-mers: either as one more read chunk (ceil(m / K2_TS) strips) or transposed
// (ceil(m / rows) strips of structure rows, the t read k-mers streamed; t < rows <= K2_TS).
__host__ __device__ inline int64_t k2_cells_padded_tail(int rows, int t, int m, bool transposed) {
    return transposed ? (int64_t)((m + rows - 1) / rows) * rows * ((t + 31) & ~31)
                      : (int64_t)rows * ((m + 31) & ~31);
}
__host__ __device__ inline int k2_tail_strips(int rows, int t, int m, bool transposed) {
    if (t == 0) return 0;
    return transposed ? (m + rows - 1) / rows : (m + K2_TS - 1) / K2_TS;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Four consecutive code bytes from any alignment: one or two aligned 32-bit loads + a funnel shift.  Reads up to 3 bytes
// before p and 3 bytes past p + 3 (inside the code buffer: it starts 256-byte aligned and is allocated with 64 bytes of slack).
__device__ __forceinline__ uint32_t k2_load4(const uint8_t* p) {
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p - a);
    const uint32_t lo = q[0];
    return a ? __funnelshift_r(lo, q[1], 8u * a) : lo;
}

// Confirm a candidate on the code strings: structure k-mer == read k-mer, forward or reverse-complemented -- four bases per
// step (byte-wise compare of packed codes; the reverse complement is a byte reversal + invert_base on the four codes at once:
// codes with bit 2 set (N, n, invalid) keep their value, the others flip their low two bits), the last k mod 4 bases one by one.
// Every match of a hashed word (k > 15, or a k-mer holding N / lower case) comes through here.
__device__ __noinline__ bool verify_kmer(const uint8_t* __restrict__ cr, const uint8_t* __restrict__ cs, int k) {
    bool fwd = true, rev = true;
    int t = 0;
    for (; t + 4 <= k; t += 4) {
        const uint32_t s = k2_load4(cs + t) & 0x0F0F0F0Fu;
        const uint32_t r = k2_load4(cr + t) & 0x0F0F0F0Fu;
        fwd &= (s == r);
        uint32_t q = __byte_perm(k2_load4(cr + k - 4 - t) & 0x0F0F0F0Fu, 0u, 0x0123);    // bases k-1-t, k-2-t, k-3-t, k-4-t
        q ^= 0x03030303u & ~(((q >> 2) & 0x01010101u) * 3u);
        rev &= (s == q);
        if (!(fwd | rev)) return false;
    }
    for (; t < k; ++t) {
        const int s = cs[t] & 15;
        fwd &= (s == (cr[t] & 15));
        rev &= (s == comp_code(cr[k - 1 - t] & 15));
    }
    return fwd | rev;
}

struct K2Params {
    const Plot* plots;          // plots of this launch
    const Operand* ops;
    const int64_t* strip_prefix;   // [n_plots+1] cumulative strips
    int n_plots;
    long long n_strips;
    long long strip_base;          // strip_prefix values are offset by this
    const uint32_t* hash;
    const uint8_t* code;
    uint2* hits;
    uint32_t* cnt;              // [n_plots] hits found (may exceed cap)
    unsigned long long* queue;  // strip queue head
    uint32_t* overflow;         // set when any plot exceeded its capacity
    uint32_t* qc;               // [n_qc_plots * QC_WORDS] counters of PLOT_QC plots (may be null when there is none)
};

// Row groups of a lane: a candidate flag covers one group.  Groups 0 and 1 are the two halves of the ISETP rows,
// groups 2 and 3 the rows of the two polynomials.
template <int NI, int NP> struct K2Groups {
    static constexpr int NG = 2 + NP;
    __host__ __device__ static constexpr int base(int g) { return g == 0 ? 0 : (g == 1 ? NI / 2 : NI + K2_D * (g - 2)); }
    __host__ __device__ static constexpr int count(int g) { return g == 0 ? NI / 2 : (g == 1 ? NI - NI / 2 : K2_D); }
};

// Matched cells are parked in a per-warp shared-memory queue by the exact path and emitted here, one cell
// per lane: hashed words are confirmed on the code strings, the hit count of the plot is bumped once per
// batch (a single atomic for up to 32 cells instead of one round trip per cell) and the hits are appended.
struct K2Strip {                 // what emission needs to know about the current strip
    const uint8_t* code_read;    // code string of the read operand
    const uint8_t* code_struct;  // code string of the structure operand, at the cut
    uint32_t* cnt;               // hit counter of the plot
    uint2* hits;                 // hit list of the plot
    uint32_t* qc;                // QC counters of the plot, or null
    uint32_t* overflow;
    uint32_t cap;
    int k;
    bool swap;                   // rows = structure k-mers, stream = read k-mers
};

__device__ __forceinline__ void k2_flush(const K2Strip& st, const uint2* queue, int qn, int lane)
{
    for (int base = 0; base < qn; base += 32) {
        const int i = base + lane;
        const bool have = i < qn;
        const uint2 e = have ? queue[i] : make_uint2(0u, 0u);
        const int cs_ = (int)(e.x & 0x0FFFFFFFu), cr_ = (int)e.y;
        const int x = st.swap ? cr_ : cs_;                           // structure k-mer
        const int y = st.swap ? cs_ : cr_;                           // read k-mer
        bool ok = have;
        if (have && (e.x & H_NEEDS_VERIFY)) ok = verify_kmer(st.code_read + y, st.code_struct + x, st.k);
        const uint32_t mult = ok ? 1u + ((e.x >> 30) & 1u) : 0u;     // H_PALINDROME: the reference appends the dot twice
        if (st.qc) {                                                 // self-plot QC: count, store nothing
            if (mult) {
                atomicAdd(&st.qc[0], mult);
                if (x == y) atomicAdd(&st.qc[1], mult);
                else if (x > y) {
                    atomicAdd(&st.qc[2], mult);
                    atomicMin(&st.qc[3], (uint32_t)x); atomicMax(&st.qc[4], (uint32_t)x);
                    atomicMin(&st.qc[5], (uint32_t)y); atomicMax(&st.qc[6], (uint32_t)y);
                }
            }
            continue;
        }
        uint32_t incl = mult;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t slot0 = 0;
        if (lane == 31 && incl) slot0 = atomicAdd(st.cnt, incl);
        slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 31);
        if (mult) {
            const uint32_t slot = slot0 + incl - mult;
            if (slot + mult <= st.cap) {
                st.hits[slot] = make_uint2((uint32_t)x, (uint32_t)y);
                if (mult == 2) st.hits[slot + 1] = make_uint2((uint32_t)x, (uint32_t)y);
            } else {
                *st.overflow = 1u;
            }
        }
    }
}

// Park matched cells in the per-warp hit queue (about 1e-4 of all cells get here): `hm` = rows j of this lane's
// note equal to the streamed word `ve` at coordinate `cs`; row j sits at coordinate `cr0 + 32 j`.  Slots by ballot +
// popc; the queue is emitted when it cannot take another round.  Warp-collective.
__device__ __forceinline__ void k2_park(uint32_t hm, uint32_t ve, int cs, int cr0, const K2Strip& st, uint2* queue, int& qn, int lane)
{
    while (true) {
        const unsigned act = __ballot_sync(0xFFFFFFFFu, hm != 0u);
        if (act == 0u) break;
        if (qn + __popc(act) > K2_QCAP) { __syncwarp(); k2_flush(st, queue, qn, lane); qn = 0; __syncwarp(); }
        if (hm) {
            const int j = __ffs(hm) - 1;
            hm &= hm - 1;
            queue[qn + __popc(act & ((1u << lane) - 1u))] = make_uint2((uint32_t)cs | (ve & 0xC0000000u), (uint32_t)(cr0 + 32 * j));
        }
        qn += __popc(act);
    }
}

// Resolve candidate notes, one note per lane and round: the 8 rows of the noted lane's group (fetched by shuffle:
// every lane names its own source lane) against the 32 streamed words of the noted block.  Lanes walk the block in
// rotated order so the shared-memory reads of a round never share a bank.  Matches are parked in the hit queue.
template <int NI, int NP, int R>
__device__ __forceinline__ void k2_resolve(const uint32_t (&r)[R], const uint32_t* notes, int nn, const uint32_t* sb,
                                           int stream_origin, int row0, const K2Strip& st, uint2* queue, int& qn, int lane)
{
    using G = K2Groups<NI, NP>;
    for (int base = 0; base < nn; base += 32) {
        const bool have = base + lane < nn;
        const uint32_t note = have ? notes[base + lane] : (uint32_t)(lane << 3);
        const int b = (int)(note >> 8), L = (int)(note >> 3) & 31, g = (int)(note & 7u);
        uint32_t rows[K2_D];
        #pragma unroll
        for (int j = 0; j < K2_D; ++j) rows[j] = H_ROW_PAD;
        #pragma unroll
        for (int gg = 0; gg < G::NG; ++gg) {
            #pragma unroll
            for (int j = 0; j < G::count(gg); ++j) {
                const uint32_t w = __shfl_sync(0xFFFFFFFFu, r[G::base(gg) + j], L);
                if (have && g == gg) rows[j] = w;
            }
        }
        const int gbase = g == 0 ? G::base(0) : (g == 1 ? G::base(1) : (g == 2 ? G::base(2) : G::base(3)));
        const uint32_t* sblk = sb + b * 32;
        // pass 1, no warp-collective work: every lane walks its block and remembers (up to 4) positions whose word
        // equals one of its 8 rows; predicated bookkeeping only, so the loop runs at full rate
        uint32_t rec = 0, cnt = 0;
        #pragma unroll 4
        for (int t = 0; t < 32; ++t) {
            const int tt = (t + lane) & 31;
            const uint32_t v = sblk[tt];
            bool any = false;
            #pragma unroll
            for (int j = 0; j < K2_D; ++j) any |= (v == rows[j]);
            if (any) { rec = (rec << 8) | (uint32_t)tt; ++cnt; }
        }
        // pass 2: the remembered positions, one per lane and step (a note has one dot in the typical case)
        const unsigned many = __ballot_sync(0xFFFFFFFFu, cnt > 4u);     // repeats: more matches than rec holds
        for (uint32_t step = 0; step < 4; ++step) {
            const bool mine = step < min(cnt, 4u);
            if (__ballot_sync(0xFFFFFFFFu, mine) == 0u) break;
            const int tt = (int)((rec >> (8 * step)) & 31u);
            const uint32_t ve = sblk[tt];
            uint32_t hm = 0;
            if (mine && !((many >> lane) & 1u)) {
                #pragma unroll
                for (int j = 0; j < K2_D; ++j) if (rows[j] == ve) hm |= 1u << j;
            }
            k2_park(hm, ve, stream_origin + b * 32 + tt, row0 + gbase * 32 + L, st, queue, qn, lane);
        }
        if (many) {                                                 // rare: walk the block again, parking as we go
            for (int t = 0; t < 32; ++t) {
                const int tt = (t + lane) & 31;
                const uint32_t ve = sblk[tt];
                uint32_t hm = 0;
                if ((many >> lane) & 1u) {
                    #pragma unroll
                    for (int j = 0; j < K2_D; ++j) if (rows[j] == ve) hm |= 1u << j;
                }
                k2_park(hm, ve, stream_origin + b * 32 + tt, row0 + gbase * 32 + L, st, queue, qn, lane);
            }
        }
    }
}

#ifndef K2_MINB
#define K2_MINB 4                              // resident CTAs per SM the register allocation aims for
#endif

template <int NI, int NP>
__global__ void __launch_bounds__(K2_THREADS, K2_MINB)
k2_tile_match(const K2Params p)
{
    constexpr int R = NI + K2_D * NP;              // rows per lane; row q of lane l is strip row q*32 + l
    constexpr int ROWS = 32 * R;
    static_assert(NP <= 2 && R <= 32, "row split");
    __shared__ __align__(16) uint32_t s_buf[K2_WARPS][K2_SBUF];
    __shared__ __align__(8) uint64_t s_bar[K2_WARPS];
    __shared__ __align__(8) uint2 s_queue[K2_WARPS][K2_QCAP];
    __shared__ uint32_t s_notes[K2_WARPS][K2_NCAP];
    __shared__ K2Strip s_strip[K2_WARPS];           // read only by the (rare) emission code: keeps it out of registers

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t* sb = s_buf[warp];
    uint64_t* bar = &s_bar[warp];
    uint2* queue = s_queue[warp];
    uint32_t* notes = s_notes[warp];
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;

    while (true) {
        unsigned long long strip = 0;
        if (lane == 0) strip = atomicAdd(p.queue, 1ull);
        strip = __shfl_sync(0xFFFFFFFFu, strip, 0);
        if (strip >= (unsigned long long)p.n_strips) break;

        // which plot: last index with prefix <= strip.  A 32-ary search, one probe per lane and a vote per level:
        // 4 dependent global loads for a million plots instead of 20
        int lo = 0, hi = p.n_plots;
        while (hi - lo > 1) {
            const int span = hi - lo;
            const int step = (span + 31) >> 5;                       // lane l probes lo + (l+1)*step
            const int probe = lo + (lane + 1) * step;
            const bool le = probe < hi && (unsigned long long)(p.strip_prefix[probe] - p.strip_base) <= strip;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, le);       // prefix is non-decreasing: a run of low lanes
            const int nle = __popc(m);
            const int new_lo = lo + nle * step;
            hi = min(hi, new_lo + step);
            lo = new_lo;
        }
        const Plot pl = p.plots[lo];
        const int local = (int)(strip - (unsigned long long)(p.strip_prefix[lo] - p.strip_base));
        const Operand opr = p.ops[pl.read_op];
        const Operand ops_ = p.ops[pl.struct_op];
        const long long read_words = opr.hash_off;                   // element index of read k-mer 0
        const long long struct_words = ops_.hash_off + pl.miss;      // element index of structure k-mer 0 (after the cut)

        // ---- decode the strip: which words are rows (registers), which are streamed (shared memory) ----
        const int n_full = pl.n / ROWS;                              // full read chunks
        bool swap = false;                                           // true: rows = structure k-mers, stream = read k-mers
        long long rows_elem, stream_elem;
        int rows_valid, stream_valid, row0, stream0;                 // row0/stream0: coordinate of the first row / streamed word
        if (local < pl.n_main_strips) {
            const int rc = local % n_full, cc = local / n_full;      // read chunk fastest: neighbours share the structure chunk in L2
            row0 = rc * ROWS; rows_valid = ROWS; rows_elem = read_words + row0;
            stream0 = cc * K2_TS; stream_valid = min(K2_TS, pl.m - stream0); stream_elem = struct_words + stream0;
        } else {
            const int lt = local - pl.n_main_strips;
            const int t0 = n_full * ROWS, t = pl.n - t0;             // the tail read chunk
            if (pl.kind & PLOT_TAIL_T) {
                swap = true;
                row0 = lt * ROWS; rows_valid = min(ROWS, pl.m - row0); rows_elem = struct_words + row0;
                stream0 = t0; stream_valid = t; stream_elem = read_words + t0;
            } else {
                row0 = t0; rows_valid = t; rows_elem = read_words + t0;
                stream0 = lt * K2_TS; stream_valid = min(K2_TS, pl.m - stream0); stream_elem = struct_words + stream0;
            }
        }

        K2Strip& st = s_strip[warp];
        if (lane == 0) {
            st.code_read = p.code + opr.code_off; st.code_struct = p.code + ops_.code_off + pl.miss;
            st.cnt = p.cnt + lo; st.hits = p.hits + pl.hit_off; st.overflow = p.overflow; st.cap = pl.cap; st.k = opr.k; st.swap = swap;
            st.qc = (pl.kind & PLOT_QC) ? p.qc + pl.hit_off * QC_WORDS : nullptr;
        }
        int qn = 0;                                                  // cells parked in the queue
        int nn = 0;                                                  // candidate notes awaiting resolution

        // ---- stage the streamed words with one TMA bulk copy ---------------------------------------
        const int shift = (int)(stream_elem & 3);                    // TMA wants a 16-byte aligned source
        const uint32_t bytes = (uint32_t)(((shift + stream_valid + 3) & ~3) * 4);
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic writes to sb
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s(sb, p.hash + (stream_elem - shift), bytes, bar);
        }

        // ---- row words into registers (coalesced: row q of lane l = strip row q*32+l) ---------------
        uint32_t r[R];
        {
            const uint32_t* rh = p.hash + rows_elem + lane;
            const int left = rows_valid - lane;                      // row q of this lane exists while q*32 < left
            #pragma unroll
            for (int q = 0; q < R; ++q) r[q] = (q * 32 < left) ? rh[q * 32] : H_ROW_PAD;
        }
        // ---- row polynomials: P_g(v) = prod over the valid rows j of group g of (v - r[NI+8g+j]), as
        //      c[g][0] v^8 + c[g][1] v^7 + ... + c[g][8].  Padding rows are left out of the product (their
        //      common value would stack trailing zero bits and fake candidates). -------------------------
        uint32_t c[NP > 0 ? NP : 1][K2_D + 1];
        #pragma unroll
        for (int g = 0; g < NP; ++g) {
            uint32_t a[K2_D + 1];                                    // a[i] = coefficient of v^i
            a[0] = 1u;
            #pragma unroll
            for (int i = 1; i <= K2_D; ++i) a[i] = 0u;
            #pragma unroll
            for (int j = 0; j < K2_D; ++j) {
                const uint32_t w = r[NI + K2_D * g + j];
                if (w != H_ROW_PAD) {
                    #pragma unroll
                    for (int i = j + 1; i >= 1; --i) a[i] = a[i - 1] - w * a[i];
                    a[0] = 0u - w * a[0];
                }
            }
            #pragma unroll
            for (int i = 0; i <= K2_D; ++i) c[g][i] = a[K2_D - i];
        }

        mbar_wait(bar, parity);
        parity ^= 1;
        // vote blocks start at the 16-byte aligned head of the buffer (LDS.128): the `shift` words in front of
        // the chunk and the tail padding are overwritten with a word no row equals
        const int nblk = (shift + stream_valid + 31) >> 5;
        if (lane < shift) sb[lane] = H_STREAM_PAD;
        for (int i = shift + stream_valid + lane; i < nblk * 32; i += 32) sb[i] = H_STREAM_PAD;
        __syncwarp();

        for (int b = 0; b < nblk; ++b) {
            const uint32_t* sblk = sb + b * 32;
            bool pi0 = false, pi1 = false, pg0 = false, pg1 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v4 = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t v = vw[e];
                    uint32_t acc[NP > 0 ? NP : 1];
                    #pragma unroll
                    // the streamed word is written as the FIRST source of every instruction: ptxas keeps the order, and the
                    // operand-reuse cache then serves v to consecutive ISETP/IMAD without a register-file read (+12 %)
                    for (int g = 0; g < NP; ++g) acc[g] = v * c[g][0] + c[g][1];
                    #pragma unroll
                    for (int i = 2; i <= K2_D; ++i)
                        #pragma unroll
                        for (int g = 0; g < NP; ++g) acc[g] = v * acc[g] + c[g][i];
                    #pragma unroll
                    for (int q = 0; q < NI - NI / 2; ++q) {
                        if (q < NI / 2) pi0 |= (v == r[q]);
                        pi1 |= (v == r[NI / 2 + q]);
                    }
                    if (NP > 0) pg0 |= (acc[0] == 0u);
                    if (NP > 1) pg1 |= (acc[NP > 1 ? 1 : 0] == 0u);
                }
            }
            // ---- candidates: a lane with a flagged row group leaves a note (block, lane, group) and moves on ----
            unsigned act = __ballot_sync(0xFFFFFFFFu, pi0 | pi1 | pg0 | pg1);
            if (act == 0u) continue;
            unsigned flags = (pi0 ? 1u : 0u) | (pi1 ? 2u : 0u) | (pg0 ? 4u : 0u) | (pg1 ? 8u : 0u);
            while (act) {                                            // one pass per flagged group of the busiest lane (almost always 1)
                if (nn + __popc(act) > K2_NCAP) {
                    __syncwarp();
                    k2_resolve<NI, NP>(r, notes, nn, sb, stream0 - shift, row0, st, queue, qn, lane);
                    nn = 0;
                    __syncwarp();
                }
                if (flags) {
                    const int g = __ffs(flags) - 1;
                    flags &= flags - 1;
                    notes[nn + __popc(act & ((1u << lane) - 1u))] = (uint32_t)((b << 8) | (lane << 3) | g);
                }
                nn += __popc(act);
                act = __ballot_sync(0xFFFFFFFFu, flags != 0u);
            }
        }
        __syncwarp();
        if (nn) { k2_resolve<NI, NP>(r, notes, nn, sb, stream0 - shift, row0, st, queue, qn, lane); nn = 0; }
        __syncwarp();
        if (qn) k2_flush(st, queue, qn, lane);
        __syncwarp();       // every lane is done with sb before the next strip's copy lands
    }
}

}  // namespace vb
