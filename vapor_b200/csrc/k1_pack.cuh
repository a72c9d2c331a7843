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
// HBM-bound: 1 B/base read, 1 B/base + 4 B/position written.
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

__global__ void __launch_bounds__(K1_THREADS)
k1_pack_kmers(const uint8_t* __restrict__ seq, const Operand* __restrict__ ops,
              const int32_t* __restrict__ chunk_prefix,   // [n_ops+1] cumulative chunk counts
              int n_ops, uint32_t* __restrict__ hash, uint8_t* __restrict__ code,
              int32_t* __restrict__ op_status)
{
    __shared__ uint8_t s_code[K1_CHUNK + K1_MAXK + 8];
    // locate (operand, chunk) of this CTA
    int lo = 0, hi = n_ops;                 // last op with prefix <= blockIdx.x
    const int bid = blockIdx.x;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk_prefix[mid] <= bid) lo = mid; else hi = mid;
    }
    const Operand op = ops[lo];
    const int chunk = bid - chunk_prefix[lo];
    const int base0 = chunk * K1_CHUNK;
    const int k = op.k;
    const int nload = min(K1_CHUNK + k - 1, op.len - base0);     // bases this CTA needs
    const bool upper = (op.flags & OPF_UPPER) != 0;
    const uint8_t* src = seq + op.seq_begin + base0;

    for (int i = threadIdx.x; i < nload; i += K1_THREADS) {
        uint8_t b = src[i];
        if (upper && b >= 'a' && b <= 'z') b -= 32;              // str.upper() on ASCII
        s_code[i] = c_code_lut[b];
    }
    __syncthreads();

    const int npos_code = min(K1_CHUNK, op.len - base0);          // bases owned by this chunk
    const int npos_hash = min(K1_CHUNK, op.n - base0);            // k-mers owned by this chunk (may be <= 0)
    const bool is_read = (op.flags & OPF_READ) != 0;
    bool bad_read = false;

    for (int i = threadIdx.x; i < npos_code; i += K1_THREADS) {
        int c0 = s_code[i];
        uint32_t h = H_STRUCT_INVALID;
        bool pal = false;
        if (i < npos_hash) {
            bool invalid = false, pure = true;
            #pragma unroll 1
            for (int t = 0; t < k; ++t) {
                int c = s_code[i + t];
                invalid |= (c == CODE_INVALID);
                pure &= (c < 4);
            }
            if (invalid) {
                if (is_read) { bad_read = true; h = H_READ_PAD; }
            } else if (pure && k <= 15) {
                uint32_t f = 0, r = 0;
                #pragma unroll 1
                for (int t = 0; t < k; ++t) {
                    f = (f << 2) | (uint32_t)s_code[i + t];
                    r = (r << 2) | (uint32_t)(3 - s_code[i + k - 1 - t]);
                }
                pal = (f == r);
                h = mix30(min(f, r)) | (pal ? H_PALINDROME : 0u);
            } else {
                const uint64_t B = 0x9E3779B97F4A7C15ull | 1ull;
                uint64_t f = 0, r = 0;
                #pragma unroll 1
                for (int t = 0; t < k; ++t) {
                    f = f * B + (uint64_t)(s_code[i + t] + 1);
                    r = r * B + (uint64_t)(comp_code(s_code[i + k - 1 - t]) + 1);
                }
                if (f == r) {                          // candidate palindrome: confirm exactly
                    pal = true;
                    for (int t = 0; t < k; ++t)
                        pal &= (s_code[i + t] == comp_code(s_code[i + k - 1 - t]));
                }
                uint64_t cmin = f < r ? f : r, cmax = f < r ? r : f;
                uint32_t h30 = (uint32_t)(fmix64(cmin ^ (cmax * 0xD6E8FEB86659FD93ull)) >> 34);
                h = H_NEEDS_VERIFY | (pal ? H_PALINDROME : 0u) | h30;
                if (h > H_MAX_VALID) h -= 4;
            }
            hash[op.hash_off + base0 + i] = h;
        }
        code[op.code_off + base0 + i] = (uint8_t)(c0 | (pal ? 0x80 : 0));
    }
    if (bad_read) atomicOr(&op_status[lo], 1);
}

}  // namespace vb
