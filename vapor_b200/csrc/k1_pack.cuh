// Kernel 1: sequence packing + k-mer words.
//
// Reference semantics: kmerhits hashes every read k-mer forward and reverse-complemented
// (vapor_vali/Simple_function.pyx:957-960 with subkeys, :1403-1422) after key_modify
// (:908-949) and probes with every structure k-mer (:964-975).  "s == r or s == rc(r)"
// is "canon(s) == canon(r)", so one canonical word per k-mer position is enough.
//
// Per operand this kernel writes
//   code[i]  : 4-bit alphabet code of base i (| 0x80 when the k-mer starting at i is its own
//              reverse complement), and
//   hash[i]  : 32-bit canonical k-mer word.  For k <= 15 and a pure ACGT k-mer the word is a
//              bijective mix of the exact canonical 2-bit code (< 2^30, injective: no confirmation
//              needed); otherwise
//              bit 31 is set and the word is a symmetric hash of (forward, revcomp) that the
//              tile kernel confirms on the code strings.  Bit 30 marks a k-mer that is its own
//              reverse complement: the reference appends such a read position twice.
// 1 B/base read, 1 B/base + 4 B/position written.  Bases are read with aligned 32-bit loads and turned into codes
// through a shared-memory copy of the alphabet table, then (k <= 15) packed 2 bits per base; every thread owns 8
// consecutive positions and cuts each window's word out of two packed words with one funnel shift (k1_run_exact) --
// O(1) per position with no state carried between positions; k > 15 rolls two polynomial hashes instead (k1_run).
// Results leave as 16-byte (words) and 8-byte (codes) stores.
#pragma once
#include "common.cuh"

namespace vb {

#ifndef K1_THREADS_N
#define K1_THREADS_N 256
#endif
constexpr int K1_THREADS = K1_THREADS_N;
constexpr int K1_CHUNK   = 8 * K1_THREADS; // positions per CTA (8 per thread)
constexpr int K1_MAXK    = 40;

__constant__ uint8_t c_code_lut[256];

__device__ __forceinline__ uint64_t fmix64(uint64_t z) {
    z ^= z >> 33; z *= 0xFF51AFD7ED558CCDull;
    z ^= z >> 33; z *= 0xC4CEB9FE1A85EC53ull;
    z ^= z >> 33;
    return z;
}

// Bijection on [0, 2^30): exact words stay exact (equal iff the canonical k-mers are equal) while their
// low bits stop being "the last bases of the k-mer" -- the tile kernel's row polynomials live in
// Z/2^32 and give a false candidate when differences share many trailing zero bits.
__host__ __device__ __forceinline__ uint32_t mix30(uint32_t x) {
    x ^= x >> 15; x = (x * 0x2C1B3C6Du) & 0x3FFFFFFFu;
    x ^= x >> 12; x = (x * 0x297A2D39u) & 0x3FFFFFFFu;
    x ^= x >> 15;
    return x;
}

// Hashed words (k > 15, or a k-mer holding N / lower case).  The k-mer and its reverse complement are hashed as
// polynomials over Z/2^64 -- f = sum (c_t + 1) B^(k-1-t), r = sum (comp(c_t) + 1) B^t -- and the word is a symmetric
// mix of the two, so that a k-mer and its reverse complement get the same word.  Both sums can be *rolled* from one
// position to the next (multiply by B, resp. by B^-1): O(1) per position.
constexpr uint64_t K1_B = 0x9E3779B97F4A7C15ull | 1ull;

__device__ __forceinline__ uint64_t k1_inv64(uint64_t b) {       // inverse of an odd number mod 2^64 (Newton)
    uint64_t x = b;                                              // b*b = 1 mod 8
    #pragma unroll
    for (int i = 0; i < 5; ++i) x *= 2ull - b * x;
    return x;
}

__device__ __forceinline__ uint32_t k1_hash_finish(uint64_t f, uint64_t r, const uint8_t* w, int k, bool& pal) {
    pal = false;
    if (f == r) {                          // candidate palindrome: confirm exactly
        pal = true;
        for (int t = 0; t < k; ++t) pal &= (w[t] == comp_code(w[k - 1 - t]));
    }
    const uint64_t cmin = f < r ? f : r, cmax = f < r ? r : f;
    const uint32_t h30 = (uint32_t)(fmix64(cmin ^ (cmax * 0xD6E8FEB86659FD93ull)) >> 34);
    uint32_t h = H_NEEDS_VERIFY | (pal ? H_PALINDROME : 0u) | h30;
    if (h > H_MAX_VALID) h -= 4;
    return h;
}

// one k-mer from scratch (k <= 15 windows that hold N / lower case: rare)
__device__ __forceinline__ uint32_t k1_hashed_word(const uint8_t* w, int k, bool& pal) {
    uint64_t f = 0, r = 0;
    #pragma unroll 1
    for (int t = 0; t < k; ++t) {
        f = f * K1_B + (uint64_t)(w[t] + 1);
        r = r * K1_B + (uint64_t)(comp_code(w[k - 1 - t]) + 1);
    }
    return k1_hash_finish(f, r, w, k, pal);
}

constexpr int K1_RUN = K1_CHUNK / K1_THREADS;      // consecutive positions per thread (8)

// Words and codes of one thread's run of K1_RUN positions.  EXACT: k <= 15 (rolling 2-bit words), else rolling
// polynomial hashes; two instantiations so that neither carries the other's state in registers.
template <bool EXACT>
__device__ __forceinline__ bool k1_run(const uint8_t* s_code, const Operand& op, int base0, int k,
                                       uint32_t* __restrict__ hash, uint8_t* __restrict__ code)
{
    const int npos_code = min(K1_CHUNK, op.len - base0);          // bases owned by this chunk
    const int npos_hash = min(K1_CHUNK, op.n - base0);            // k-mers owned by this chunk (may be <= 0)
    const bool is_read = (op.flags & OPF_READ) != 0;
    bool bad_read = false;
    const int p0 = threadIdx.x * K1_RUN;
    if (p0 < npos_code) {
        uint32_t words[K1_RUN];
        uint8_t codes[K1_RUN];
        // rolling state of the window [p, p+k): 2-bit forward / reverse-complement words (k <= 15), how many of its
        // codes are not plain ACGT, how many are invalid
        const uint32_t mask = EXACT ? ((k == 16 ? 0u : (1u << (2 * k))) - 1u) : 0u;
        uint32_t f = 0, r = 0;
        uint64_t hf = 0, hr = 0, bk1 = 1, binv = 0;              // k > 15: rolling polynomial hashes, B^(k-1), B^-1
        int n_np = 0, n_inv = 0;
        if (p0 < npos_hash) {
            if (!EXACT) {
                for (int t = 1; t < k; ++t) bk1 *= K1_B;
                binv = k1_inv64(K1_B);
            }
            uint64_t bt = 1;                                     // B^t
            for (int t = 0; t < k; ++t) {
                const int c = s_code[p0 + t];
                n_np += (c >= 4); n_inv += (c == CODE_INVALID);
                if (EXACT) {
                    f = ((f << 2) | (uint32_t)(c & 3)) & mask;
                    r = (r >> 2) | ((uint32_t)(3 - (c & 3)) << (2 * (k - 1)));
                } else {
                    hf = hf * K1_B + (uint64_t)(c + 1);
                    hr += (uint64_t)(comp_code(c) + 1) * bt;
                    bt *= K1_B;
                }
            }
        }
        #pragma unroll
        for (int i = 0; i < K1_RUN; ++i) {
            const int p = p0 + i;
            const int c0 = (p < npos_code) ? s_code[p] : 0;
            uint32_t h = H_STRUCT_INVALID;
            bool pal = false;
            if (p < npos_hash) {
                if (n_inv > 0) {
                    if (is_read) { bad_read = true; h = H_READ_PAD; }
                } else if (EXACT && n_np == 0) {
                    pal = (f == r);
                    h = mix30(min(f, r)) | (pal ? H_PALINDROME : 0u);
                } else if (EXACT) {
                    h = k1_hashed_word(s_code + p, k, pal);
                } else {
                    h = k1_hash_finish(hf, hr, s_code + p, k, pal);
                }
                if (p + 1 < npos_hash) {                           // roll the window one base on
                    const int cn = s_code[p + k];
                    n_np += (cn >= 4) - (c0 >= 4);
                    n_inv += (cn == CODE_INVALID) - (c0 == CODE_INVALID);
                    if (EXACT) {
                        f = ((f << 2) | (uint32_t)(cn & 3)) & mask;
                        r = (r >> 2) | ((uint32_t)(3 - (cn & 3)) << (2 * (k - 1)));
                    } else {
                        hf = (hf - (uint64_t)(c0 + 1) * bk1) * K1_B + (uint64_t)(cn + 1);
                        hr = (hr - (uint64_t)(comp_code(c0) + 1)) * binv + (uint64_t)(comp_code(cn) + 1) * bk1;
                    }
                }
            }
            words[i] = h;
            codes[i] = (uint8_t)(c0 | (pal ? 0x80 : 0));
        }
        // ---- stores: 8 words = two 16-byte stores, 8 codes = one 8-byte store when the run is complete --------
        uint32_t* hp = hash + op.hash_off + base0 + p0;
        uint8_t* cp = code + op.code_off + base0 + p0;
        if (p0 + K1_RUN <= npos_hash) {
            reinterpret_cast<uint4*>(hp)[0] = make_uint4(words[0], words[1], words[2], words[3]);
            reinterpret_cast<uint4*>(hp)[1] = make_uint4(words[4], words[5], words[6], words[7]);
        } else {
            #pragma unroll
            for (int i = 0; i < K1_RUN; ++i) if (p0 + i < npos_hash) hp[i] = words[i];
        }
        if (p0 + K1_RUN <= npos_code) {
            uint2 v;
            v.x = codes[0] | (codes[1] << 8) | (codes[2] << 16) | ((uint32_t)codes[3] << 24);
            v.y = codes[4] | (codes[5] << 8) | (codes[6] << 16) | ((uint32_t)codes[7] << 24);
            *reinterpret_cast<uint2*>(cp) = v;
        } else {
            #pragma unroll
            for (int i = 0; i < K1_RUN; ++i) if (p0 + i < npos_code) cp[i] = codes[i];
        }
    }
    return bad_read;
}


// Words and codes of one thread's run of K1_RUN positions for k <= 15, from the chunk's 2-bit packed bases (16 per word,
// base i at bits 2i) and its "not plain ACGT" / "invalid" bit strings (32 per word).  With g = the window's bases packed
// first-base-lowest, the reverse-complement word of the rolling form (r = sum (3 - c_t) << 2t) is ~g, and the forward word
// (f = sum c_t << 2(k-1-t)) is g with its k base pairs in reverse order: brev + a swap inside the pairs.  No per-thread
// window set-up, no state carried from one position to the next: 30 instructions per position instead of 128.
__device__ __forceinline__ bool k1_run_exact(const uint8_t* s_code, const uint32_t* s_p2, const uint32_t* s_np, const uint32_t* s_inv,
                                             const Operand& op, int base0, int k,
                                             uint32_t* __restrict__ hash, uint8_t* __restrict__ code)
{
    const int npos_code = min(K1_CHUNK, op.len - base0);          // bases owned by this chunk
    const int npos_hash = min(K1_CHUNK, op.n - base0);            // k-mers owned by this chunk (may be <= 0)
    const bool is_read = (op.flags & OPF_READ) != 0;
    bool bad_read = false;
    const int p0 = threadIdx.x * K1_RUN;
    if (p0 < npos_code) {
        const uint32_t mask = (1u << (2 * k)) - 1u, kmask = (1u << k) - 1u;
        const int wi = p0 >> 4, off = (2 * p0) & 31, bi = p0 >> 5, boff = p0 & 31;
        const uint32_t w0 = s_p2[wi], w1 = s_p2[wi + 1];
        const uint32_t n0 = s_np[bi], n1 = s_np[bi + 1], v0 = s_inv[bi], v1 = s_inv[bi + 1];
        const int rsh = 32 - 2 * k;
        uint32_t* hp = hash + op.hash_off + base0 + p0;
        uint8_t* cp = code + op.code_off + base0 + p0;
        #pragma unroll
        for (int half = 0; half < K1_RUN / 4; ++half) {             // four positions at a time: one 16-byte and one 4-byte store
            uint32_t words[4];
            uint32_t codes = 0;
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 4 * half + j, p = p0 + i;
                const uint32_t c0 = (p < npos_code) ? s_code[p] : 0u;
                uint32_t h = H_STRUCT_INVALID;
                bool pal = false;
                if (p < npos_hash) {
                    if ((__funnelshift_r(n0, n1, boff + i) & kmask) == 0u) {
                        const uint32_t g = __funnelshift_r(w0, w1, off + 2 * i) & mask;
                        const uint32_t r = ~g & mask;
                        const uint32_t b = __brev(g);
                        const uint32_t f = (((b & 0x55555555u) << 1) | ((b >> 1) & 0x55555555u)) >> rsh;
                        pal = (f == r);
                        h = mix30(min(f, r)) | (pal ? H_PALINDROME : 0u);
                    } else if ((__funnelshift_r(v0, v1, boff + i) & kmask) != 0u) {
                        if (is_read) { bad_read = true; h = H_READ_PAD; }
                    } else {
                        h = k1_hashed_word(s_code + p, k, pal);
                    }
                }
                words[j] = h;
                codes |= (c0 | (pal ? 0x80u : 0u)) << (8 * j);
            }
            const int q0 = p0 + 4 * half;
            if (q0 + 4 <= npos_hash) {
                reinterpret_cast<uint4*>(hp)[half] = make_uint4(words[0], words[1], words[2], words[3]);
            } else {
                #pragma unroll
                for (int j = 0; j < 4; ++j) if (q0 + j < npos_hash) hp[4 * half + j] = words[j];
            }
            if (q0 + 4 <= npos_code) {
                reinterpret_cast<uint32_t*>(cp)[half] = codes;
            } else {
                #pragma unroll
                for (int j = 0; j < 4; ++j) if (q0 + j < npos_code) cp[4 * half + j] = (uint8_t)(codes >> (8 * j));
            }
        }
    }
    return bad_read;
}

#ifndef K1_MINB
#define K1_MINB 8                              // measured: 8 resident CTAs per SM (32 registers)
#endif

// operand of every CTA of a k1_pack_kmers launch (one thread per operand): the kernel used to find it by a 17-step binary
// search over the chunk prefix -- dependent global loads that were 39 % of its stall samples
__global__ void k1_map_chunks(const int32_t* __restrict__ chunk_prefix, int chunk_base, int n_ops, int32_t* __restrict__ cta_op) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_ops) return;
    for (int c = chunk_prefix[o] - chunk_base, e = chunk_prefix[o + 1] - chunk_base; c < e; ++c) cta_op[c] = o;
}

constexpr int K1_PACK_GROUPS = (K1_CHUNK + 64) / 8;              // 8-base groups packed per chunk (the chunk + the longest window + slack)

__global__ void __launch_bounds__(K1_THREADS, K1_MINB)
k1_pack_kmers(const uint8_t* __restrict__ seq, const Operand* __restrict__ ops,
              const int32_t* __restrict__ chunk_prefix,   // [n_ops+1] cumulative chunk counts, offset by chunk_base
              const int32_t* __restrict__ cta_op,         // operand of every CTA (k1_map_chunks)
              int chunk_base, int n_ops, uint32_t* __restrict__ hash, uint8_t* __restrict__ code,
              int32_t* __restrict__ op_status)
{
    __shared__ uint8_t s_lut[256];                               // the alphabet table, out of the constant cache:
                                                                 // a per-lane index there would replay 32 ways
    __shared__ __align__(16) uint8_t s_raw[K1_CHUNK + K1_MAXK + 24];
    __shared__ __align__(4) uint16_t s_p2[K1_PACK_GROUPS + 2];   // 2-bit codes, 8 bases per entry (k <= 15)
    __shared__ __align__(4) uint8_t s_np[K1_PACK_GROUPS + 8];    // "not plain ACGT" bits, 8 bases per entry
    __shared__ __align__(4) uint8_t s_inv[K1_PACK_GROUPS + 8];   // "invalid character" bits
    for (int i = threadIdx.x; i < 256; i += K1_THREADS) s_lut[i] = c_code_lut[i];
    const int bid = blockIdx.x;
    const int lo = cta_op[bid];
    const Operand op = ops[lo];
    const int chunk = bid - (chunk_prefix[lo] - chunk_base);
    const int base0 = chunk * K1_CHUNK;
    const int k = op.k;
    const int nload = min(K1_CHUNK + k - 1, op.len - base0);     // bases this CTA needs
    const bool upper = (op.flags & OPF_UPPER) != 0;
    const uint8_t* src = seq + op.seq_begin + base0;

    // ---- bases -> codes: aligned 32-bit loads, four table look-ups, one 32-bit shared store ----------------
    const int head = (int)(reinterpret_cast<uintptr_t>(src) & 3);    // s_code[i] = code of base i lives at s_raw[head + i]
    __syncthreads();                                                 // s_lut is complete
    {
        const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src - head);
        const int nwords = (head + nload + 3) >> 2;
        for (int j = threadIdx.x; j < nwords; j += K1_THREADS) {
            uint32_t w = src4[j], c4 = 0;
            #pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t b = (w >> (8 * e)) & 0xFFu;
                if (upper && b >= 'a' && b <= 'z') b -= 32;          // str.upper() on ASCII
                c4 |= (uint32_t)s_lut[b] << (8 * e);
            }
            reinterpret_cast<uint32_t*>(s_raw)[j] = c4;
        }
    }
    __syncthreads();
    const uint8_t* s_code = s_raw + head;

    bool bad_read;
    if (k <= 15) {
        // pack the chunk: 2 bits per base + the two flag bits per base, 8 bases per thread and step
        for (int g8 = threadIdx.x; g8 < K1_PACK_GROUPS; g8 += K1_THREADS) {
            uint32_t p2 = 0, np = 0, iv = 0;
            #pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = 8 * g8 + e;
                const uint32_t c = i < nload ? s_code[i] : 0u;
                p2 |= (c & 3u) << (2 * e);
                np |= (uint32_t)(c >= 4u) << e;
                iv |= (uint32_t)(c == (uint32_t)CODE_INVALID) << e;
            }
            s_p2[g8] = (uint16_t)p2; s_np[g8] = (uint8_t)np; s_inv[g8] = (uint8_t)iv;
        }
        __syncthreads();
        bad_read = k1_run_exact(s_code, reinterpret_cast<const uint32_t*>(s_p2), reinterpret_cast<const uint32_t*>(s_np),
                                reinterpret_cast<const uint32_t*>(s_inv), op, base0, k, hash, code);
    } else {
        bad_read = k1_run<false>(s_code, op, base0, k, hash, code);
    }
    if (bad_read) atomicOr(&op_status[lo], 1);
}

}  // namespace vb
