// Kernel 2: recurrence-plot tile kernel.
//
// Reference semantics: for every structure k-mer i and read k-mer j the reference emits the
// dot (i, j) when the strings are equal forward or reverse-complemented
// (vapor_vali/Simple_function.pyx:964-979), a self-reverse-complement read k-mer emitting it
// twice (:959-960, :1419-1421).  The reference gets there with a Python dict; this kernel
// evaluates every cell of the n x m plot with one 32-bit compare on canonical k-mer words and
// never materialises the matrix: only the sparse hit list reaches HBM.
//
// Decomposition: a *strip* is 32*R consecutive read k-mers (R per lane, in registers) against
// K2_TS consecutive structure k-mers staged in shared memory by a 1-D TMA bulk copy
// (cp.async.bulk + mbarrier, one buffer per warp, no block-wide barrier anywhere).  Each warp
// of a persistent grid pulls strips from a global queue.  The inner loop is R compares per
// broadcast shared-memory word, OR-accumulated into predicates; a warp vote every 32
// structure k-mers sends the rare blocks that hold a match to a warp-cooperative slow path
// that locates, confirms (hash words with bit 31 set) and appends the hits.
#pragma once
#include "common.cuh"

namespace vb {

constexpr int K2_R       = 16;                 // read k-mers per lane
constexpr int K2_ROWS    = 32 * K2_R;          // read k-mers per strip
constexpr int K2_TS      = 2048;               // structure k-mers per strip
constexpr int K2_WARPS   = 4;                  // warps per CTA
constexpr int K2_THREADS = 32 * K2_WARPS;
constexpr int K2_SBUF    = K2_TS + 40;         // words per warp buffer (alignment shift + vote-block padding)

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

// Confirm a candidate on the code strings: structure k-mer == read k-mer, forward or reverse-complemented.
__device__ __forceinline__ bool verify_kmer(const uint8_t* __restrict__ cr, const uint8_t* __restrict__ cs, int k) {
    bool fwd = true, rev = true;
    for (int t = 0; t < k; ++t) {
        int s = cs[t] & 15;
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
};

__global__ void __launch_bounds__(K2_THREADS)
k2_tile_match(const K2Params p)
{
    __shared__ __align__(16) uint32_t s_buf[K2_WARPS][K2_SBUF];
    __shared__ __align__(8) uint64_t s_bar[K2_WARPS];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t* sb = s_buf[warp];
    uint64_t* bar = &s_bar[warp];
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

        // which plot: last index with prefix <= strip
        int lo = 0, hi = p.n_plots;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if ((unsigned long long)(p.strip_prefix[mid] - p.strip_base) <= strip) lo = mid; else hi = mid;
        }
        const Plot pl = p.plots[lo];
        const int local = (int)(strip - (unsigned long long)(p.strip_prefix[lo] - p.strip_base));
        const int n_rc = (pl.n + K2_ROWS - 1) / K2_ROWS;
        const int rc = local % n_rc;             // read chunk fastest: neighbours share the structure chunk in L2
        const int cc = local / n_rc;
        const Operand opr = p.ops[pl.read_op];
        const Operand ops_ = p.ops[pl.struct_op];

        // ---- stage the structure words with one TMA bulk copy ------------------------------------
        const int x0 = cc * K2_TS;
        const int valid = min(K2_TS, pl.m - x0);
        const long long src_elem = ops_.hash_off + pl.miss + x0;
        const int shift = (int)(src_elem & 3);                       // TMA wants 16-byte aligned source
        const uint32_t bytes = (uint32_t)(((shift + valid + 3) & ~3) * 4);
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic writes to sb
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s(sb, p.hash + (src_elem - shift), bytes, bar);
        }

        // ---- read words into registers while the copy is in flight -------------------------------
        uint32_t r[K2_R];
        const int ybase = rc * K2_ROWS + lane * K2_R;
        {
            const uint32_t* rh = p.hash + opr.hash_off + ybase;
            #pragma unroll
            for (int q = 0; q < K2_R; q += 4) {
                if (ybase + q + 3 < pl.n) {
                    uint4 v = *reinterpret_cast<const uint4*>(rh + q);
                    r[q] = v.x; r[q + 1] = v.y; r[q + 2] = v.z; r[q + 3] = v.w;
                } else {
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) r[q + e] = (ybase + q + e < pl.n) ? rh[q + e] : H_READ_PAD;
                }
            }
        }

        mbar_wait(bar, parity);
        parity ^= 1;
        const int nblk = (valid + 31) >> 5;
        for (int i = shift + valid + lane; i < shift + nblk * 32; i += 32) sb[i] = H_STRUCT_INVALID;
        __syncwarp();

        const uint32_t* s = sb + shift;
        for (int b = 0; b < nblk; ++b) {
            const uint32_t* sblk = s + b * 32;
            bool p0 = false, p1 = false, p2 = false, p3 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint32_t v0 = sblk[jj], v1 = sblk[jj + 1], v2 = sblk[jj + 2], v3 = sblk[jj + 3];
                #pragma unroll
                for (int q = 0; q < K2_R; ++q) {
                    p0 |= (r[q] == v0);
                    p1 |= (r[q] == v1);
                    p2 |= (r[q] == v2);
                    p3 |= (r[q] == v3);
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, p0 | p1 | p2 | p3);
            if (mask == 0) continue;

            // ---- slow path: lane t takes structure k-mer t of the block against each flagged lane's rows
            const uint32_t v = sblk[lane];
            const int x = x0 + b * 32 + lane;
            while (mask) {
                const int L = __ffs(mask) - 1;
                mask &= mask - 1;
                #pragma unroll
                for (int q = 0; q < K2_R; ++q) {
                    const uint32_t rq = __shfl_sync(0xFFFFFFFFu, r[q], L);
                    if (rq == v) {
                        const int y = rc * K2_ROWS + L * K2_R + q;
                        const uint8_t* cr = p.code + opr.code_off + y;
                        bool ok = true;
                        if (v & H_NEEDS_VERIFY)
                            ok = verify_kmer(cr, p.code + ops_.code_off + pl.miss + x, opr.k);
                        if (ok) {
                            const uint32_t mult = 1u + (uint32_t)(cr[0] >> 7);
                            const uint32_t slot = atomicAdd(&p.cnt[lo], mult);
                            if (slot + mult <= pl.cap) {
                                uint2* out = p.hits + pl.hit_off + slot;
                                out[0] = make_uint2((uint32_t)x, (uint32_t)y);
                                if (mult == 2) out[1] = make_uint2((uint32_t)x, (uint32_t)y);
                            } else {
                                *p.overflow = 1u;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();       // every lane is done with sb before the next strip's copy lands
    }
}

}  // namespace vb
