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
// HBM-bound: 1 B/base read, 1 B/base + 4 B/position written.  Bases are read with aligned 32-bit loads and turned
// into codes through a shared-memory copy of the alphabet table; every thread then owns 8 consecutive positions and
// *rolls* the 2-bit forward / reverse-complement words of its window from one position to the next (k <= 15), so the
// work per position is O(1), not O(k); results leave as 16-byte (words) and 8-byte (codes) stores.
#pragma once
#include "common.cuh"

namespace vb {

constexpr int K1_THREADS = 256;
constexpr int K1_CHUNK   = 2048;          // positions per CTA
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


#ifndef K1_MINB
#define K1_MINB 8                              // measured: 8 resident CTAs per SM (32 registers)
#endif

__global__ void __launch_bounds__(K1_THREADS, K1_MINB)
k1_pack_kmers(const uint8_t* __restrict__ seq, const Operand* __restrict__ ops,
              const int32_t* __restrict__ chunk_prefix,   // [n_ops+1] cumulative chunk counts, offset by chunk_base
              int chunk_base, int n_ops, uint32_t* __restrict__ hash, uint8_t* __restrict__ code,
              int32_t* __restrict__ op_status)
{
    __shared__ uint8_t s_lut[256];                               // the alphabet table, out of the constant cache:
                                                                 // a per-lane index there would replay 32 ways
    __shared__ __align__(16) uint8_t s_raw[K1_CHUNK + K1_MAXK + 24];
    __shared__ int s_op;
    s_lut[threadIdx.x] = c_code_lut[threadIdx.x];
    // locate (operand, chunk) of this CTA: one thread searches, everybody reads the answer (the search is 17 dependent
    // loads for a 100 000-operand wave -- a fifth of the kernel's instructions when all 256 threads repeat it)
    const int bid = blockIdx.x;
    if (threadIdx.x == 0) {
        int lo_ = 0, hi_ = n_ops;                 // last op with prefix <= blockIdx.x
        while (hi_ - lo_ > 1) {
            const int mid = (lo_ + hi_) >> 1;
            if (chunk_prefix[mid] - chunk_base <= bid) lo_ = mid; else hi_ = mid;
        }
        s_op = lo_;
    }
    __syncthreads();
    const int lo = s_op;
    const Operand op = ops[lo];
    const int chunk = bid - (chunk_prefix[lo] - chunk_base);
    const int base0 = chunk * K1_CHUNK;
    const int k = op.k;
    const int nload = min(K1_CHUNK + k - 1, op.len - base0);     // bases this CTA needs
    const bool upper = (op.flags & OPF_UPPER) != 0;
    const uint8_t* src = seq + op.seq_begin + base0;

    // ---- bases -> codes: aligned 32-bit loads, four table look-ups, one 32-bit shared store ----------------
    const int head = (int)(reinterpret_cast<uintptr_t>(src) & 3);    // s_code[i] = code of base i lives at s_raw[head + i]
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

    const bool bad_read = (k <= 15) ? k1_run<true>(s_code, op, base0, k, hash, code) : k1_run<false>(s_code, op, base0, k, hash, code);
    if (bad_read) atomicOr(&op_status[lo], 1);
}

}  // namespace vb
