// vapor_b200 C-ABI: host-side planning + launches of kernels 1-4.  See include/vapor_b200.h.
//
// There is deliberately no CPU implementation of the scoring path in this file: every result
// an entry point returns was computed by the kernels in k1..k4; without a CUDA device
// vapor_gpu_open fails and nothing else can be called.
#include "../../include/vapor_b200.h"
#include "common.cuh"
#include "k1_pack.cuh"
#include "k2_tile.cuh"
#include "k2_join.cuh"
#include "k3_score.cuh"
#include "k3_warp.cuh"
#include "k4_genotype.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

using namespace vb;

namespace {

std::string g_open_error;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return VAPOR_E_CUDA;                                                          \
        }                                                                                 \
    } while (0)

template <typename T>
struct DevBuf {                      // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// A wave = a run of whole SVs whose transient device data (k-mer words, codes, tables, hit lists) fits the budget.
// Offsets inside Operand / TabChunk / Plot are relative to the wave's buffers, which the next wave reuses.
struct Wave {
    int64_t task_begin, task_end;
    int64_t plot_begin, plot_end;
    int64_t op_begin, op_end;
    int64_t chunk_begin, chunk_end;              // table chunks (join mode)
    int64_t hit_elems, hash_elems, code_bytes, table_bytes;
    int64_t bytes() const { return 8 * hit_elems + 4 * hash_elems + code_bytes + table_bytes; }
};

// kernel-3 scratch classes: bins that fit in shared memory, then a global-memory fallback
// Classes 0 .. K3W_NCLASS-1 run on the warp-per-task kernel (k3_warp.cuh), the next one on the CTA-per-task kernel with
// shared-memory scratch, the last one on the CTA kernel with global scratch.
constexpr int K3_NCLASS = 7;
constexpr int K3W_NCLASS = 5;
const int k3_class_cap[K3_NCLASS - 1] = {2048, 4096, 8192, 16384, K3W_MAX_NB, 110000};

// join-kernel launch classes by table blob size (dynamic shared memory of the launch)
constexpr int K2J_NCLASS = 3;
const int k2j_class_cap[K2J_NCLASS] = {k2j_blob_bytes(2048, k2j_bits(2048)), k2j_blob_bytes(4096, k2j_bits(4096)), k2j_blob_bytes(K2J_CH, k2j_bits(K2J_CH))};

// the k2 plan of a set of plots: strips for the tile kernel, or table chunks + items for the join kernel
struct JoinPlan {
    std::vector<JoinItem> items;                 // grouped by (wave, class)
    std::vector<int64_t> item_off;               // [n_waves * K2J_NCLASS + 1]
    std::vector<int32_t> jplots;                 // plot ids, grouped by structure operand
};

struct PlanShard;
struct Handle {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // kernel 3 of wave i runs here while kernels 1-2 of wave i+1 run on `stream`
    int overlap = 0;                 // 1 = kernel 3 of wave i on stream2 under kernels 1-2 of wave i+1 (measured: -2.5 % of the step, but the
                                     // phase times then overlap; off by default so that every kernel is timed alone)
    std::vector<cudaEvent_t> ev_k2, ev_k3;       // per wave: kernel 2 done (stream), kernel 3 done (stream2)
    cudaEvent_t ev_run0 = nullptr, ev_run1 = nullptr;
    std::string err;
    int64_t hit_budget = 0;          // bytes; 0 = default
    int64_t default_hit_budget = (int64_t)6 << 30;
    int k2_ctas_per_sm = 0;          // persistent-grid size of the tile kernel = this x SM count; 0 = occupancy
    int k2_occupancy[K2_NVARIANT] = {0, 0, 0, 0, 0};
    int tile_variant = 4;            // inner-loop variant of the tile kernel (k2_variant_*), fixed at upload
    int plan_variant = 4;            // variant the resident plan's strips were cut for
    int k2_mode = 1;                 // 0 = all-pairs tile kernel (k2_tile.cuh), 1 = join kernel (k2_join.cuh)
    int plan_mode = 1;               // mode the resident plan was made for
    int plan_threads = 0;            // host threads for planning (0 = auto)
    int debug_sync = getenv("VAPOR_DEBUG_SYNC") ? atoi(getenv("VAPOR_DEBUG_SYNC")) : 0;   // bit mask: 1 k1, 2 k1b, 4 k2, 8 k3

    // plan (host)
    std::vector<Operand> ops;
    std::vector<Plot> plots;
    std::vector<Task> tasks;
    std::vector<int32_t> chunk_prefix;
    std::vector<int64_t> strip_prefix;
    std::vector<Wave> waves;
    std::vector<int64_t> group_off;              // [n_groups+1] task ranges of the planning groups (SVs)
    std::vector<PlanShard>* shards = nullptr;    // planner scratch, kept between calls (capacity reuse)
    ~Handle();
    std::vector<int32_t> class_ids;              // task ids grouped by (wave, class)
    std::vector<int64_t> class_off;              // [n_waves*K3_NCLASS+1]
    std::vector<TabChunk> chunks;                // join mode: table chunks of all structure-side operands
    JoinPlan jp;
    int64_t table_bytes = 0;
    int max_nb = 0;
    int64_t n_task = 0, n_sv = 0, n_seq = 0, seq_total = 0;
    int64_t hash_elems = 0, code_bytes = 0, max_wave_hits = 0;
    bool resident = false, ran = false;
    vapor_timings_t tm{};
    int64_t tm_padded_cells = 0;     // cells the tile kernel evaluates including strip padding
    std::map<int32_t, int64_t> ovf_loc;   // plots whose hit list sits in d_ovf_hits after the last run: plot -> element offset

    // device
    DevBuf<uint8_t> d_seq, d_code, d_task_status, d_sv_gt, d_table;
    DevBuf<Operand> d_ops;
    DevBuf<Plot> d_plots;
    DevBuf<Task> d_tasks;
    DevBuf<TabChunk> d_chunks;
    DevBuf<JoinItem> d_items;
    DevBuf<int32_t> d_chunk_prefix, d_op_status, d_class_ids, d_sv_nscore, d_jplots, d_k1map;
    DevBuf<int64_t> d_strip_prefix, d_sv_off;
    DevBuf<uint32_t> d_hash, d_cnt, d_task_hits, d_gscratch, d_ovf_flags, d_qc, d_k3q;
    int k3_mode = 1;                 // 1 = warp-per-task kernel for the classes it covers, 0 = CTA-per-task kernel everywhere
    int k3w_nclass = 3;              // classes 0 .. k3w_nclass-1 go to the warp kernel (measured: it wins up to 8 192 bins)
    DevBuf<uint2> d_hits, d_hits2, d_ovf_hits;   // hit slabs of even / odd waves (d_hits2 only with overlap and > 1 wave)
    DevBuf<double> d_task_score, d_task_stat, d_pos, d_sv_qs, d_sv_gs, d_sv_gq;
    DevBuf<unsigned long long> d_task_hitsum, d_queue, d_stats;
    unsigned long long* h_stats = nullptr;       // pinned: [0] hits, [1] evaluated cells
    uint32_t* h_flags = nullptr;                 // pinned: overflow flag per wave
    size_t h_flags_cap = 0;

    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    size_t ev_used = 0;
    struct Span { size_t ev; int cat; };
    std::vector<Span> spans;
};

enum { CAT_H2D = 0, CAT_PACK, CAT_TABLE, CAT_TILE, CAT_SCORE, CAT_SCORE_W, CAT_GENO, CAT_D2H, CAT_N };   // CAT_SCORE_W: the warp kernel 3

int span_begin(Handle* h, int cat, cudaStream_t st = nullptr) {
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        h->ev_pool.push_back({a, b});
    }
    size_t i = h->ev_used++;
    cudaEventRecord(h->ev_pool[i].first, st ? st : h->stream);
    h->spans.push_back({i, cat});
    return (int)i;
}
void span_end(Handle* h, int i, cudaStream_t st = nullptr) { cudaEventRecord(h->ev_pool[i].second, st ? st : h->stream); }

void collect_spans(Handle* h, float* acc /*CAT_N*/) {
    for (auto& s : h->spans) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev_pool[s.ev].first, h->ev_pool[s.ev].second);
        acc[s.cat] += ms;
    }
    h->spans.clear();
    h->ev_used = 0;
}

uint8_t host_code_lut[256];
void build_lut() {
    for (int i = 0; i < 256; ++i) host_code_lut[i] = CODE_INVALID;
    const char* acgt = "ACGT";
    for (int i = 0; i < 4; ++i) {
        host_code_lut[(int)acgt[i]] = (uint8_t)i;
        host_code_lut[(int)acgt[i] + 32] = (uint8_t)(8 + i);
    }
    host_code_lut[(int)'N'] = 4; host_code_lut[(int)'n'] = 12;
    const char* iupac = "RYSWKMBDHV";                 // key_modify, Simple_function.pyx:909-948
    for (int i = 0; i < 10; ++i) { host_code_lut[(int)iupac[i]] = 4; host_code_lut[(int)iupac[i] + 32] = 12; }
}

inline int64_t align4(int64_t v) { return (v + 3) & ~(int64_t)3; }

// Cut one plot into strips for the tile kernel (see k2_tile.cuh): sets kind/n_main_strips, returns the strip count.
int64_t cut_strips(Plot& p, int rows, int64_t* padded_cells) {
    p.kind &= ~PLOT_TAIL_T; p.n_main_strips = 0;
    if (p.n <= 0 || p.m <= 0) return 0;
    const int n_full = p.n / rows, t = p.n % rows;
    const int ncol = (p.m + K2_TS - 1) / K2_TS;
    p.n_main_strips = n_full * ncol;
    int64_t strips = p.n_main_strips;
    if (padded_cells) *padded_cells += (int64_t)n_full * rows * ((p.m + 31) & ~31);
    if (t > 0) {
        const int64_t c_n = k2_cells_padded_tail(rows, t, p.m, false), c_t = k2_cells_padded_tail(rows, t, p.m, true);
        const bool tr = c_t < c_n;
        if (tr) p.kind |= PLOT_TAIL_T;
        strips += k2_tail_strips(rows, t, p.m, tr);
        if (padded_cells) *padded_cells += tr ? c_t : c_n;
    }
    return strips;
}

int k3_class_of(int nb) {
    for (int c = 0; c < K3_NCLASS - 1; ++c) if (nb <= k3_class_cap[c]) return c;
    return K3_NCLASS - 1;
}

// ------------------------------------------------------------------------------------------
// planning: operands (k-mer word arrays), plots, tasks, waves
// ------------------------------------------------------------------------------------------

// Table chunks of every structure-side operand (join kernel), wave by wave: blob offsets restart in every wave.
// op_chunk0[i] = first chunk of operand i, or -1.  Returns the largest table size of a wave.
int64_t build_chunks(const std::vector<Operand>& ops, std::vector<Wave>& waves, std::vector<TabChunk>& chunks, std::vector<int32_t>& op_chunk0) {
    chunks.clear();
    op_chunk0.assign(ops.size(), -1);
    int64_t worst = 0;
    for (Wave& w : waves) {
        int64_t off = 0;
        w.chunk_begin = (int64_t)chunks.size();
        for (int64_t i = w.op_begin; i < w.op_end; ++i) {
            const Operand& o = ops[i];
            if (!(o.flags & OPF_TABLE) || o.n <= 0) continue;
            op_chunk0[i] = (int32_t)chunks.size();
            for (int p0 = 0; p0 < o.n; p0 += K2J_CH) {
                TabChunk c{};
                c.op = (int32_t)i; c.pos0 = p0; c.len = std::min(K2J_CH, o.n - p0); c.bits = k2j_bits(c.len);
                c.blob_bytes = k2j_blob_bytes(c.len, c.bits); c.blob_off = off;
                off += c.blob_bytes;
                chunks.push_back(c);
            }
        }
        w.chunk_end = (int64_t)chunks.size();
        w.table_bytes = off;
        worst = std::max(worst, off);
    }
    return worst;
}

int k2j_class_of(int blob_bytes) {
    for (int c = 0; c < K2J_NCLASS - 1; ++c) if (blob_bytes <= k2j_class_cap[c]) return c;
    return K2J_NCLASS - 1;
}

// Items of the join kernel for every wave (= contiguous plot range): the wave's plots grouped by structure operand,
// every group cut into runs of K2J_PLOTS_PER_ITEM plots, one item per (table chunk of the operand, run).
void build_join_items(const std::vector<Plot>& plots, const std::vector<Wave>& waves, const std::vector<TabChunk>& chunks,
                      const std::vector<int32_t>& op_chunk0, JoinPlan& jp)
{
    jp.items.clear(); jp.jplots.clear();
    jp.item_off.assign(waves.size() * K2J_NCLASS + 1, 0);
    jp.jplots.reserve(plots.size());
    std::vector<int32_t> count, sorted;
    std::vector<JoinItem> cls[K2J_NCLASS];
    for (size_t wi = 0; wi < waves.size(); ++wi) {
        const int64_t pb = waves[wi].plot_begin, pe = waves[wi].plot_end;
        int32_t lo = INT32_MAX, hi = -1;
        for (int64_t i = pb; i < pe; ++i) {
            const Plot& p = plots[i];
            if (p.n <= 0 || p.m <= 0) continue;
            lo = std::min(lo, p.struct_op); hi = std::max(hi, p.struct_op);
        }
        for (auto& v : cls) v.clear();
        if (hi >= lo) {
            // counting sort of the wave's non-empty plots by structure operand (stable: task order inside a group)
            count.assign((size_t)(hi - lo) + 2, 0);
            for (int64_t i = pb; i < pe; ++i) { const Plot& p = plots[i]; if (p.n > 0 && p.m > 0) ++count[p.struct_op - lo + 1]; }
            for (size_t q = 1; q < count.size(); ++q) count[q] += count[q - 1];
            const int32_t total = count.back();
            sorted.resize((size_t)total);
            {
                std::vector<int32_t> cur(count.begin(), count.end() - 1);
                for (int64_t i = pb; i < pe; ++i) { const Plot& p = plots[i]; if (p.n > 0 && p.m > 0) sorted[cur[p.struct_op - lo]++] = (int32_t)i; }
            }
            for (int32_t op = lo; op <= hi; ++op) {
                const int32_t g0 = count[op - lo], g1 = count[op - lo + 1];
                if (g1 == g0 || op_chunk0[op] < 0) continue;
                const int32_t jbase = (int32_t)jp.jplots.size();
                jp.jplots.insert(jp.jplots.end(), sorted.begin() + g0, sorted.begin() + g1);
                for (int32_t c = op_chunk0[op]; c < (int32_t)chunks.size() && chunks[c].op == op; ++c) {
                    const int k = k2j_class_of(chunks[c].blob_bytes);
                    // runs of plots of about K2J_WORDS_PER_ITEM read words (at most K2J_PLOTS_PER_ITEM plots): items of similar length
                    int32_t a = 0;
                    while (a < g1 - g0) {
                        int32_t e = a; int64_t words = 0;
                        while (e < g1 - g0 && e - a < K2J_PLOTS_PER_ITEM && (e == a || words + plots[sorted[g0 + e]].n <= K2J_WORDS_PER_ITEM)) {
                            words += plots[sorted[g0 + e]].n; ++e;
                        }
                        cls[k].push_back(JoinItem{c, jbase + a, jbase + e, 0});
                        a = e;
                    }
                }
            }
        }
        for (int k = 0; k < K2J_NCLASS; ++k) {
            jp.items.insert(jp.items.end(), cls[k].begin(), cls[k].end());
            jp.item_off[wi * K2J_NCLASS + k + 1] = (int64_t)jp.items.size();
        }
    }
}

// ---- the batch planner ---------------------------------------------------------------------------------------
// Groups = SVs (or runs of 64 tasks when the batch has no SV table).  Everything a group needs -- its operands (one
// per distinct (sequence, k, casing, role)), plots, tasks, table chunks, join items -- is local to the group, so the
// plan is made in three phases:
//   A (parallel over contiguous group ranges): each shard builds its groups' records with group-relative ids/offsets;
//   B (serial, O(groups)): waves are cut at group boundaries by the memory budget, every group gets its bases;
//   C (parallel): the records are copied to their final places with the bases added.
// The result does not depend on the number of threads (tests/test_host_logic.py pins the digest).
struct GroupSize {
    int32_t n_ops, n_plots, n_tasks, n_chunks, n_jplots, n_k1chunks;
    int32_t n_items[K2J_NCLASS];
    int32_t n_cls[K3_NCLASS];
    int64_t n_strips;
    int64_t hash_elems, code_bytes, table_bytes, hit_elems;
};
struct GroupBase {
    int64_t op, plot, chunk, jplot, k1chunk, strip;
    int64_t hash, code, table, hit;
    int64_t item[K2J_NCLASS];
    int64_t cls[K3_NCLASS];
};
struct PlanShard {
    int64_t g0 = 0, g1 = 0;
    std::vector<Operand> ops; std::vector<Plot> plots; std::vector<Task> tasks; std::vector<TabChunk> chunks;
    std::vector<JoinItem> items; std::vector<uint8_t> item_class; std::vector<int32_t> jplots;
    std::vector<uint8_t> task_class; std::vector<int32_t> k1chunks; std::vector<int32_t> strips;
    int64_t cells = 0, padded_cells = 0, bases = 0, probe_words = 0;
    int max_nb = 1;
    int rc = VAPOR_OK; std::string err;
};

Handle::~Handle() { delete shards; }

void plan_shard(const Handle* h, const vapor_batch_t* in, const std::vector<int64_t>& goff, PlanShard& sh, std::vector<GroupSize>& gsize) {
    const int64_t n_seq = in->n_seq;
    const int mode_k2 = h->k2_mode;
    const int rows = k2_variant_rows(h->tile_variant);
    struct Key { int32_t seq, k, flags, op; };
    std::vector<Key> local;                      // operand cache of the current group
    std::vector<std::pair<int32_t, int8_t>> lower;   // has-lower-case memo of the current group's structure sequences
    std::vector<int32_t> op_chunk0, members;
    auto fail = [&](const char* msg) { sh.rc = VAPOR_E_ARG; sh.err = msg; };
    for (int64_t g = sh.g0; g < sh.g1 && sh.rc == VAPOR_OK; ++g) {
        GroupSize gs{};
        local.clear(); lower.clear();
        const size_t op0 = sh.ops.size(), plot0 = sh.plots.size();
        auto seq_has_lower = [&](int32_t s) -> bool {
            for (auto& e : lower) if (e.first == s) return e.second != 0;
            const uint8_t* p = in->seq_bytes + in->seq_off[s];
            const int64_t L = in->seq_off[s + 1] - in->seq_off[s];
            int8_t f = 0;
            int64_t i = 0;
            // ASCII lower-case letters have bit 5 set, upper-case ones do not: skip 8 bytes at a time while no byte has it
            for (; i + 8 <= L; i += 8) {
                uint64_t w8; memcpy(&w8, p + i, 8);
                if (w8 & 0x2020202020202020ull) {
                    for (int j = 0; j < 8; ++j) if (p[i + j] >= 'a' && p[i + j] <= 'z') { f = 1; break; }
                    if (f) break;
                }
            }
            for (; !f && i < L; ++i) if (p[i] >= 'a' && p[i] <= 'z') f = 1;
            lower.push_back({s, f});
            return f != 0;
        };
        auto get_op = [&](int32_t seq, int k, int flags) -> int32_t {
            for (auto& e : local) if (e.seq == seq && e.k == k && e.flags == flags) return e.op;
            Operand o{};
            o.seq_begin = in->seq_off[seq];
            o.len = (int32_t)(in->seq_off[seq + 1] - in->seq_off[seq]);
            o.k = k; o.flags = flags;
            o.n = std::max(0, o.len - k + 1);
            o.hash_off = gs.hash_elems; o.code_off = gs.code_bytes;
            gs.hash_elems += align4(o.n) + 8;
            gs.code_bytes += (o.len + 8 + 15) & ~15;     // 16-byte aligned code strings (8-byte stores in kernel 1)
            const int k1c = std::max(1, (o.len + K1_CHUNK - 1) / K1_CHUNK);
            sh.k1chunks.push_back(k1c); gs.n_k1chunks += k1c;
            sh.bases += o.len;
            sh.ops.push_back(o);
            const int32_t id = (int32_t)(sh.ops.size() - 1 - op0);
            local.push_back({seq, k, flags, id});
            return id;
        };
        auto add_plot = [&](int32_t rop, int32_t sop, int32_t miss_bp) -> int32_t {
            const Operand& r = sh.ops[op0 + rop];
            Operand& s = sh.ops[op0 + sop];
            // the reference slices ref_seq[miss_bp:] (Simple_function.pyx:185-186): a negative miss_bp, which
            // cigar2alignstart_by_pos can return, counts from the end of the string as Python slicing does
            const int32_t miss = miss_bp >= 0 ? miss_bp : std::max(0, s.len + miss_bp);
            Plot p{};
            p.read_op = rop; p.struct_op = sop; p.miss = miss;
            p.n = r.n;
            p.m = (miss <= s.len) ? std::max(0, s.len - miss - s.k + 1) : 0;
            p.cap = (p.n > 0 && p.m > 0) ? (uint32_t)align4((int64_t)p.n + p.m + 32) : 0u;
            p.hit_off = gs.hit_elems;
            gs.hit_elems += p.cap;
            s.flags |= OPF_TABLE;
            sh.max_nb = std::max(sh.max_nb, p.n + p.m - 1);
            sh.cells += (int64_t)p.n * p.m;
            if (p.n > 0 && p.m > 0) sh.probe_words += p.n;
            int64_t ns = 0;
            if (mode_k2 == 0) ns = cut_strips(p, rows, &sh.padded_cells);
            sh.strips.push_back((int32_t)ns); gs.n_strips += ns;
            sh.plots.push_back(p);
            return (int32_t)(sh.plots.size() - 1 - plot0);
        };
        for (int64_t t = goff[g]; t < goff[g + 1]; ++t) {
            const int32_t rs = in->task_read[t], fs = in->task_ref[t], as = in->task_alt[t];
            const int32_t miss = in->task_miss[t];
            const int k = in->task_k[t], mode = in->task_mode[t];
            if (rs < 0 || rs >= n_seq || fs < 0 || fs >= n_seq || as < 0 || as >= n_seq) { fail("task sequence index out of range"); break; }
            if (k < 1 || k > K1_MAXK) { fail("window_size k must be 1..40"); break; }
            if (mode < 0 || mode > 3) { fail("unknown mode"); break; }
            Task tk{};
            tk.mode = mode;
            tk.len_ref = (int32_t)(in->seq_off[fs + 1] - in->seq_off[fs]);
            tk.len_alt = (int32_t)(in->seq_off[as + 1] - in->seq_off[as]);
            tk.read_op = get_op(rs, k, OPF_READ);
            const bool abs_first = (mode == VAPOR_MODE_ABS || mode == VAPOR_MODE_ABS_AND_W10);
            const int fl_ref = (abs_first && seq_has_lower(fs)) ? OPF_UPPER : 0;
            const int fl_alt = (abs_first && seq_has_lower(as)) ? OPF_UPPER : 0;
            tk.plot[0] = add_plot(tk.read_op, get_op(fs, k, fl_ref), miss);
            tk.plot[1] = add_plot(tk.read_op, get_op(as, k, fl_alt), miss);
            tk.plot[2] = tk.plot[3] = -1;
            if (mode == VAPOR_MODE_ABS_AND_W10) {       // W10 does not upper-case (Simple_function.pyx:277-279)
                tk.plot[2] = fl_ref ? add_plot(tk.read_op, get_op(fs, k, 0), miss) : tk.plot[0];
                tk.plot[3] = fl_alt ? add_plot(tk.read_op, get_op(as, k, 0), miss) : tk.plot[1];
            }
            int nb = 1;
            for (int i = 0; i < 4; ++i) if (tk.plot[i] >= 0) { const Plot& p = sh.plots[plot0 + tk.plot[i]]; nb = std::max(nb, p.n + p.m - 1); }
            const int cls = k3_class_of(nb);
            sh.task_class.push_back((uint8_t)cls); ++gs.n_cls[cls];
            sh.tasks.push_back(tk);
        }
        gs.n_ops = (int32_t)(sh.ops.size() - op0); gs.n_plots = (int32_t)(sh.plots.size() - plot0);
        gs.n_tasks = (int32_t)(goff[g + 1] - goff[g]);
        // join kernel: table chunks of the group's structure-side operands, then its items
        if (mode_k2 == 1) {
            op_chunk0.assign((size_t)gs.n_ops, -1);
            for (int32_t o = 0; o < gs.n_ops; ++o) {
                const Operand& op = sh.ops[op0 + o];
                if (!(op.flags & OPF_TABLE) || op.n <= 0) continue;
                op_chunk0[o] = gs.n_chunks;
                for (int p0 = 0; p0 < op.n; p0 += K2J_CH) {
                    TabChunk c{};
                    c.op = o; c.pos0 = p0; c.len = std::min(K2J_CH, op.n - p0); c.bits = k2j_bits(c.len);
                    c.blob_bytes = k2j_blob_bytes(c.len, c.bits); c.blob_off = gs.table_bytes;
                    gs.table_bytes += c.blob_bytes;
                    sh.chunks.push_back(c); ++gs.n_chunks;
                }
            }
            const size_t chunk0 = sh.chunks.size() - (size_t)gs.n_chunks;
            for (int32_t o = 0; o < gs.n_ops; ++o) {
                if (op_chunk0[o] < 0) continue;
                members.clear();                                   // plots on this table, in task order
                for (int32_t q = 0; q < gs.n_plots; ++q) { const Plot& p = sh.plots[plot0 + q]; if (p.struct_op == o && p.n > 0 && p.m > 0) members.push_back(q); }
                if (members.empty()) continue;
                const int32_t jbase = gs.n_jplots;
                sh.jplots.insert(sh.jplots.end(), members.begin(), members.end());
                gs.n_jplots += (int32_t)members.size();
                for (int32_t c = op_chunk0[o]; c < gs.n_chunks && sh.chunks[chunk0 + c].op == o; ++c) {
                    const int kc = k2j_class_of(sh.chunks[chunk0 + c].blob_bytes);
                    // runs of plots of about K2J_WORDS_PER_ITEM read words (at most K2J_PLOTS_PER_ITEM plots): items of similar length
                    int32_t a = 0;
                    const int32_t nm = (int32_t)members.size();
                    while (a < nm) {
                        int32_t e = a; int64_t words = 0;
                        while (e < nm && e - a < K2J_PLOTS_PER_ITEM && (e == a || words + sh.plots[plot0 + members[e]].n <= K2J_WORDS_PER_ITEM)) {
                            words += sh.plots[plot0 + members[e]].n; ++e;
                        }
                        sh.items.push_back(JoinItem{c, jbase + a, jbase + e, 0});
                        sh.item_class.push_back((uint8_t)kc); ++gs.n_items[kc];
                        a = e;
                    }
                }
            }
        }
        gsize[(size_t)g] = gs;
    }
}

void place_shard(Handle* h, const PlanShard& sh, const std::vector<GroupSize>& gsize, const std::vector<GroupBase>& gbase) {
    size_t io = 0, ip = 0, it = 0, ic = 0, ii = 0, ij = 0;             // cursors into the shard's records
    for (int64_t g = sh.g0; g < sh.g1; ++g) {
        const GroupSize& gs = gsize[(size_t)g];
        const GroupBase& gb = gbase[(size_t)g];
        int32_t k1run = (int32_t)gb.k1chunk;
        for (int32_t o = 0; o < gs.n_ops; ++o, ++io) {
            Operand op = sh.ops[io];
            op.hash_off += gb.hash; op.code_off += gb.code;
            h->ops[(size_t)gb.op + o] = op;
            h->chunk_prefix[(size_t)gb.op + o] = k1run;
            k1run += sh.k1chunks[io];
        }
        int64_t srun = gb.strip;
        for (int32_t q = 0; q < gs.n_plots; ++q, ++ip) {
            Plot p = sh.plots[ip];
            p.read_op += (int32_t)gb.op; p.struct_op += (int32_t)gb.op; p.hit_off += gb.hit;
            h->plots[(size_t)gb.plot + q] = p;
            srun += sh.strips[ip];
            h->strip_prefix[(size_t)gb.plot + q + 1] = srun;
        }
        int64_t cls_run[K3_NCLASS];
        for (int c = 0; c < K3_NCLASS; ++c) cls_run[c] = gb.cls[c];
        const int64_t task0 = h->group_off[(size_t)g];
        for (int32_t t = 0; t < gs.n_tasks; ++t, ++it) {
            Task tk = sh.tasks[it];
            for (int i = 0; i < 4; ++i) if (tk.plot[i] >= 0) tk.plot[i] += (int32_t)gb.plot;
            tk.read_op += (int32_t)gb.op;
            h->tasks[(size_t)task0 + t] = tk;
            h->class_ids[(size_t)cls_run[sh.task_class[it]]++] = (int32_t)(task0 + t);
        }
        for (int32_t c = 0; c < gs.n_chunks; ++c, ++ic) {
            TabChunk ch = sh.chunks[ic];
            ch.op += (int32_t)gb.op; ch.blob_off += gb.table;
            h->chunks[(size_t)gb.chunk + c] = ch;
        }
        for (int32_t j = 0; j < gs.n_jplots; ++j, ++ij) h->jp.jplots[(size_t)gb.jplot + j] = sh.jplots[ij] + (int32_t)gb.plot;
        int64_t item_run[K2J_NCLASS];
        for (int c = 0; c < K2J_NCLASS; ++c) item_run[c] = gb.item[c];
        int32_t n_it = 0;
        for (int c = 0; c < K2J_NCLASS; ++c) n_it += gs.n_items[c];
        for (int32_t x = 0; x < n_it; ++x, ++ii) {
            JoinItem item = sh.items[ii];
            item.chunk += (int32_t)gb.chunk; item.jp_begin += (int32_t)gb.jplot; item.jp_end += (int32_t)gb.jplot;
            h->jp.items[(size_t)item_run[sh.item_class[ii]]++] = item;
        }
    }
}

template <typename F>
void run_threads(int n, F&& f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)n - 1);
    for (int i = 1; i < n; ++i) th.emplace_back([&f, i]() { f(i); });
    f(0);
    for (auto& t : th) t.join();
}

int plan_batch(Handle* h, const vapor_batch_t* in) {
    if (!in || !in->seq_off || (!in->seq_bytes && in->n_seq > 0 && in->seq_off[in->n_seq] > 0)) { h->err = "NULL batch arrays"; return VAPOR_E_ARG; }
    if (in->n_task > 0 && (!in->task_read || !in->task_ref || !in->task_alt || !in->task_miss || !in->task_k || !in->task_mode)) {
        h->err = "NULL task arrays"; return VAPOR_E_ARG;
    }
    if (in->n_sv > 0 && !in->sv_task_off) { h->err = "NULL sv_task_off"; return VAPOR_E_ARG; }
    const int64_t n_seq = in->n_seq, n_task = in->n_task, n_sv = in->n_sv;
    for (int64_t i = 0; i < n_seq; ++i) {
        int64_t L = in->seq_off[i + 1] - in->seq_off[i];
        if (L < 0 || L >= (1ll << 27)) { h->err = "sequence length out of range (0 .. 2^27)"; return VAPOR_E_ARG; }
    }
    if (n_sv > 0 && (in->sv_task_off[0] != 0 || in->sv_task_off[n_sv] != n_task)) { h->err = "sv_task_off must cover [0, n_task]"; return VAPOR_E_ARG; }
    for (int64_t s = 0; s < n_sv; ++s)
        if (in->sv_task_off[s + 1] < in->sv_task_off[s]) { h->err = "sv_task_off must be non-decreasing"; return VAPOR_E_ARG; }

    const bool trace = getenv("VAPOR_PLAN_TRACE") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[plan] %-10s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tp0).count());
        tp0 = now;
    };
    // groups
    std::vector<int64_t>& goff = h->group_off;
    goff.clear();
    if (n_sv > 0) goff.assign(in->sv_task_off, in->sv_task_off + n_sv + 1);
    else { for (int64_t t = 0; t < n_task; t += 64) goff.push_back(t); goff.push_back(n_task); }
    const int64_t n_groups = (int64_t)goff.size() - 1;

    // phase A
    int threads = h->plan_threads;
    if (threads <= 0) {
        const char* e = getenv("VAPOR_PLAN_THREADS");
        threads = e ? atoi(e) : (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / 4u));
    }
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, n_groups / 64 + 1));
    if (!h->shards) h->shards = new std::vector<PlanShard>();
    std::vector<PlanShard>& shards = *h->shards;
    shards.resize((size_t)threads);
    for (PlanShard& sh : shards) {
        sh.ops.clear(); sh.plots.clear(); sh.tasks.clear(); sh.chunks.clear(); sh.items.clear(); sh.item_class.clear(); sh.jplots.clear();
        sh.task_class.clear(); sh.k1chunks.clear(); sh.strips.clear();
        sh.cells = sh.padded_cells = sh.bases = sh.probe_words = 0; sh.max_nb = 1; sh.rc = VAPOR_OK; sh.err.clear();
    }
    {   // contiguous group ranges of about equal task counts
        int64_t g = 0;
        for (int i = 0; i < threads; ++i) {
            shards[(size_t)i].g0 = g;
            const int64_t want = n_task * (i + 1) / threads;
            while (g < n_groups && (i == threads - 1 || goff[(size_t)g + 1] <= want)) ++g;
            shards[(size_t)i].g1 = g;
        }
        shards.back().g1 = n_groups;
    }
    std::vector<GroupSize> gsize((size_t)n_groups);
    run_threads(threads, [&](int i) {
        PlanShard& sh = shards[(size_t)i];
        const int64_t nt = goff[(size_t)sh.g1] - goff[(size_t)sh.g0];
        sh.tasks.reserve((size_t)nt); sh.plots.reserve((size_t)nt * 2); sh.ops.reserve((size_t)nt + (size_t)(sh.g1 - sh.g0) * 4);
        plan_shard(h, in, goff, sh, gsize);
    });
    for (const PlanShard& sh : shards) if (sh.rc != VAPOR_OK) { h->err = sh.err; return sh.rc; }
    lap("phase A");

    // phase B: waves (cut at group boundaries by the memory budget: an eighth of the device memory free at open(),
    // at most 16 GB, unless the caller set it) and the bases of every group
    const int64_t budget_bytes = std::max<int64_t>(h->hit_budget > 0 ? h->hit_budget : h->default_hit_budget, (int64_t)1 << 19);
    h->waves.clear();
    std::vector<GroupBase> gbase((size_t)n_groups);
    std::vector<int32_t> gwave((size_t)n_groups);
    int64_t n_ops = 0, n_plots = 0, n_chunks = 0, n_jplots = 0, n_k1 = 0, n_strips = 0;
    {
        Wave cur{};
        for (int64_t g = 0; g < n_groups; ++g) {
            const GroupSize& gs = gsize[(size_t)g];
            if (cur.bytes() >= budget_bytes && cur.task_end > cur.task_begin) {
                h->waves.push_back(cur);
                Wave nw{};
                nw.task_begin = nw.task_end = cur.task_end; nw.plot_begin = nw.plot_end = cur.plot_end; nw.op_begin = nw.op_end = cur.op_end;
                nw.chunk_begin = nw.chunk_end = cur.chunk_end;
                cur = nw;
            }
            GroupBase& gb = gbase[(size_t)g];
            gb.op = n_ops; gb.plot = n_plots; gb.chunk = n_chunks; gb.jplot = n_jplots; gb.k1chunk = n_k1; gb.strip = n_strips;
            gb.hash = cur.hash_elems; gb.code = cur.code_bytes; gb.table = cur.table_bytes; gb.hit = cur.hit_elems;
            gwave[(size_t)g] = (int32_t)h->waves.size();
            n_ops += gs.n_ops; n_plots += gs.n_plots; n_chunks += gs.n_chunks; n_jplots += gs.n_jplots; n_k1 += gs.n_k1chunks; n_strips += gs.n_strips;
            cur.hash_elems += gs.hash_elems; cur.code_bytes += gs.code_bytes; cur.table_bytes += gs.table_bytes; cur.hit_elems += gs.hit_elems;
            cur.task_end = goff[(size_t)g + 1]; cur.plot_end = n_plots; cur.op_end = n_ops; cur.chunk_end = n_chunks;
        }
        if (cur.task_end > cur.task_begin) h->waves.push_back(cur);
    }
    const size_t n_waves = h->waves.size();
    // (wave, class) offsets of the join items and of the kernel-3 task lists; groups take their slots in order
    h->jp.item_off.assign(n_waves * K2J_NCLASS + 1, 0);
    h->class_off.assign(n_waves * K3_NCLASS + 1, 0);
    for (int64_t g = 0; g < n_groups; ++g) {
        const size_t wi = std::min<size_t>((size_t)gwave[(size_t)g], n_waves ? n_waves - 1 : 0);
        if (!n_waves) break;
        for (int c = 0; c < K2J_NCLASS; ++c) h->jp.item_off[wi * K2J_NCLASS + c + 1] += gsize[(size_t)g].n_items[c];
        for (int c = 0; c < K3_NCLASS; ++c) h->class_off[wi * K3_NCLASS + c + 1] += gsize[(size_t)g].n_cls[c];
    }
    for (size_t i = 1; i < h->jp.item_off.size(); ++i) h->jp.item_off[i] += h->jp.item_off[i - 1];
    for (size_t i = 1; i < h->class_off.size(); ++i) h->class_off[i] += h->class_off[i - 1];
    {
        std::vector<int64_t> icur(h->jp.item_off.begin(), h->jp.item_off.end()), ccur(h->class_off.begin(), h->class_off.end());
        for (int64_t g = 0; g < n_groups && n_waves; ++g) {
            const size_t wi = std::min<size_t>((size_t)gwave[(size_t)g], n_waves - 1);
            GroupBase& gb = gbase[(size_t)g];
            for (int c = 0; c < K2J_NCLASS; ++c) { gb.item[c] = icur[wi * K2J_NCLASS + c]; icur[wi * K2J_NCLASS + c] += gsize[(size_t)g].n_items[c]; }
            for (int c = 0; c < K3_NCLASS; ++c) { gb.cls[c] = ccur[wi * K3_NCLASS + c]; ccur[wi * K3_NCLASS + c] += gsize[(size_t)g].n_cls[c]; }
        }
    }

    lap("phase B");
    // phase C
    h->ops.resize((size_t)n_ops); h->plots.resize((size_t)n_plots); h->tasks.resize((size_t)n_task);
    h->chunk_prefix.assign((size_t)n_ops + 1, 0); h->chunk_prefix[(size_t)n_ops] = (int32_t)n_k1;
    h->strip_prefix.assign((size_t)n_plots + 1, 0);
    h->class_ids.resize((size_t)n_task);
    h->chunks.resize((size_t)n_chunks); h->jp.jplots.resize((size_t)n_jplots); h->jp.items.resize((size_t)h->jp.item_off.back());
    lap("C alloc");
    run_threads(threads, [&](int i) { place_shard(h, shards[(size_t)i], gsize, gbase); });
    lap("phase C");

    h->plan_variant = h->tile_variant;
    h->plan_mode = h->k2_mode;
    h->max_nb = 1; h->max_wave_hits = 0; h->hash_elems = 0; h->code_bytes = 0; h->table_bytes = 0;
    int64_t cells = 0, bases = 0, padded = 0, probe = 0, table_total = 0;
    for (const PlanShard& sh : shards) { h->max_nb = std::max(h->max_nb, sh.max_nb); cells += sh.cells; bases += sh.bases; padded += sh.padded_cells; probe += sh.probe_words; }
    for (const Wave& w : h->waves) {
        h->max_wave_hits = std::max(h->max_wave_hits, w.hit_elems);
        h->hash_elems = std::max(h->hash_elems, w.hash_elems);
        h->code_bytes = std::max(h->code_bytes, w.code_bytes);
        h->table_bytes = std::max(h->table_bytes, w.table_bytes);
        table_total += w.table_bytes;
    }
    h->hash_elems += 16; h->code_bytes += 64;
    h->tm_padded_cells = padded;
    h->n_task = n_task; h->n_sv = n_sv; h->n_seq = n_seq;
    h->seq_total = n_seq > 0 ? in->seq_off[n_seq] : 0;
    h->tm = vapor_timings_t{};
    h->tm.cells = cells; h->tm.n_plots = n_plots; h->tm.n_operands = n_ops;
    h->tm.padded_cells = padded;
    h->tm.n_strips = h->plan_mode == 0 ? n_strips : (int64_t)h->jp.items.size();
    h->tm.n_waves = (int64_t)n_waves; h->tm.bases = bases;
    h->tm.k2_mode = h->plan_mode;
    h->tm.table_bytes = table_total;
    h->tm.probe_words = probe;
    return VAPOR_OK;
}

int ensure_flags(Handle* h, size_t n) {
    if (n <= h->h_flags_cap) return VAPOR_OK;
    if (h->h_flags) cudaFreeHost(h->h_flags);
    h->h_flags = nullptr; h->h_flags_cap = 0;
    CK(cudaMallocHost(&h->h_flags, (n + 64) * sizeof(uint32_t)));
    h->h_flags_cap = n + 64;
    return VAPOR_OK;
}

int upload_impl(Handle* h, const vapor_batch_t* in) {
    h->resident = false; h->ran = false;
    CK(cudaSetDevice(h->device));
    // The sequence bytes are the bulk of the upload and need no planning: start that copy first, so it runs
    // (from pinned host memory) underneath the host-side planning below.
    int sp = -1;
    if (in && in->seq_off && in->n_seq > 0 && in->seq_bytes && in->seq_off[0] == 0 && in->seq_off[in->n_seq] > 0 &&
        in->seq_off[in->n_seq] < ((int64_t)1 << 40)) {
        const size_t total = (size_t)in->seq_off[in->n_seq];
        CK(h->d_seq.ensure(total + 64));
        sp = span_begin(h, CAT_H2D);
        CK(cudaMemcpyAsync(h->d_seq.p, in->seq_bytes, total, cudaMemcpyHostToDevice, h->stream));
    }
    auto t0 = std::chrono::steady_clock::now();
    int rc = plan_batch(h, in);
    if (rc) { if (sp >= 0) { span_end(h, sp); cudaStreamSynchronize(h->stream); h->spans.clear(); h->ev_used = 0; } return rc; }
    h->tm.host_prep_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const size_t nt = (size_t)h->n_task, nsv = (size_t)h->n_sv;
    CK(h->d_seq.ensure((size_t)h->seq_total + 64));
    CK(h->d_ops.ensure(h->ops.size() + 1));
    CK(h->d_plots.ensure(h->plots.size() + 1));
    CK(h->d_tasks.ensure(nt + 1));
    CK(h->d_chunk_prefix.ensure(h->chunk_prefix.size()));
    {
        int64_t most = 0;                            // kernel-1 CTAs of the largest wave
        for (const Wave& w : h->waves) most = std::max<int64_t>(most, h->chunk_prefix[w.op_end] - h->chunk_prefix[w.op_begin]);
        CK(h->d_k1map.ensure((size_t)most + 1));
    }
    CK(h->d_class_ids.ensure(nt + 1));
    CK(h->d_sv_off.ensure(nsv + 1));
    CK(h->d_hash.ensure((size_t)h->hash_elems));
    CK(h->d_code.ensure((size_t)h->code_bytes));
    CK(h->d_op_status.ensure(h->ops.size() + 1));
    CK(h->d_cnt.ensure(h->plots.size() + 1));
    CK(h->d_hits.ensure((size_t)h->max_wave_hits + 64));
    if (h->overlap && h->waves.size() > 1) CK(h->d_hits2.ensure((size_t)h->max_wave_hits + 64));
    CK(h->d_task_score.ensure(nt + 1)); CK(h->d_task_status.ensure(nt + 1));
    CK(h->d_task_stat.ensure(4 * nt + 4)); CK(h->d_task_hits.ensure(4 * nt + 4)); CK(h->d_task_hitsum.ensure(4 * nt + 4));
    CK(h->d_pos.ensure(nt + 1));
    CK(h->d_sv_qs.ensure(nsv + 1)); CK(h->d_sv_gs.ensure(nsv + 1)); CK(h->d_sv_gq.ensure(nsv + 1));
    CK(h->d_sv_gt.ensure(nsv + 1)); CK(h->d_sv_nscore.ensure(nsv + 1));
    CK(h->d_queue.ensure(4)); CK(h->d_stats.ensure(4)); CK(h->d_ovf_flags.ensure(h->waves.size() + 1));
    CK(h->d_k3q.ensure((h->waves.size() + 1) * K3_NCLASS));
    { int rcf = ensure_flags(h, h->waves.size() + 1); if (rcf) return rcf; }
    if (k3_class_of(h->max_nb) == K3_NCLASS - 1)
        CK(h->d_gscratch.ensure((size_t)2 * h->sm_count * k3_scratch_words(h->max_nb)));
    if (h->plan_mode == 0) {
        CK(h->d_strip_prefix.ensure(h->strip_prefix.size()));
    } else {
        CK(h->d_chunks.ensure(h->chunks.size() + 1)); CK(h->d_items.ensure(h->jp.items.size() + 1));
        CK(h->d_jplots.ensure(h->jp.jplots.size() + 1)); CK(h->d_table.ensure((size_t)h->table_bytes + 64));
    }

    if (sp < 0) {
        sp = span_begin(h, CAT_H2D);
        if (h->seq_total > 0) CK(cudaMemcpyAsync(h->d_seq.p, in->seq_bytes, (size_t)h->seq_total, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(h->d_ops.p, h->ops.data(), h->ops.size() * sizeof(Operand), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_plots.p, h->plots.data(), h->plots.size() * sizeof(Plot), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_tasks.p, h->tasks.data(), nt * sizeof(Task), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_chunk_prefix.p, h->chunk_prefix.data(), h->chunk_prefix.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_class_ids.p, h->class_ids.data(), nt * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (h->plan_mode == 0) {
        CK(cudaMemcpyAsync(h->d_strip_prefix.p, h->strip_prefix.data(), h->strip_prefix.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    } else {
        CK(cudaMemcpyAsync(h->d_chunks.p, h->chunks.data(), h->chunks.size() * sizeof(TabChunk), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_items.p, h->jp.items.data(), h->jp.items.size() * sizeof(JoinItem), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_jplots.p, h->jp.jplots.data(), h->jp.jplots.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    }
    if (nsv) CK(cudaMemcpyAsync(h->d_sv_off.p, in->sv_task_off, (nsv + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    span_end(h, sp);
    CK(cudaStreamSynchronize(h->stream));
    float acc[CAT_N] = {0};
    collect_spans(h, acc);
    h->tm.h2d_ms = acc[CAT_H2D];
    h->resident = true;
    return VAPOR_OK;
}

}  // namespace

namespace {

// persistent grid: one CTA per resident CTA slot (occupancy x SM count), strips pulled from the queue
template <int V>
void launch_k2_variant(Handle* h, const K2Params& kp) {
    auto kern = k2_tile_match<k2_variant_ni(V), k2_variant_np(V)>;
    if (h->k2_occupancy[V] == 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K2_THREADS, 0) != cudaSuccess || occ < 1) occ = 4;
        h->k2_occupancy[V] = occ;
    }
    const int per_sm = h->k2_ctas_per_sm > 0 ? h->k2_ctas_per_sm : h->k2_occupancy[V];
    const int grid = (int)std::min<int64_t>((kp.n_strips + K2_WARPS - 1) / K2_WARPS, (int64_t)h->sm_count * per_sm);
    kern<<<grid, K2_THREADS, 0, h->stream>>>(kp);
}
void launch_k2(Handle* h, const K2Params& kp, int variant) {
    switch (variant) {
        case 0:  launch_k2_variant<0>(h, kp); break;
        case 2:  launch_k2_variant<2>(h, kp); break;
        case 3:  launch_k2_variant<3>(h, kp); break;
        case 1:  launch_k2_variant<1>(h, kp); break;
        default: launch_k2_variant<4>(h, kp); break;
    }
}

// the join kernel over the items [o0, o1) of every class of one wave; returns the launches made
int launch_k2_join(Handle* h, const JoinPlan& jp, size_t wi, K2JParams kp) {
    int launches = 0;
    const JoinItem* items0 = kp.items;
    for (int c = 0; c < K2J_NCLASS; ++c) {
        const int64_t o0 = jp.item_off[wi * K2J_NCLASS + c], o1 = jp.item_off[wi * K2J_NCLASS + c + 1];
        if (o1 == o0) continue;
        kp.items = items0 + o0;
        if (c == K2J_NCLASS - 1) k2_join_match<K2J_WARPS_BIG><<<(unsigned)(o1 - o0), 32 * K2J_WARPS_BIG, (size_t)k2j_class_cap[c], h->stream>>>(kp);
        else                     k2_join_match<K2J_WARPS><<<(unsigned)(o1 - o0), 32 * K2J_WARPS, (size_t)k2j_class_cap[c], h->stream>>>(kp);
        ++launches;
    }
    return launches;
}

// kernel 3 over task ids [o0, o1) of scratch class c
// `slot` = this launch's task-queue counter in d_k3q (zeroed by the caller)
// `allow_warp` = false for the re-scored tasks of an overflowed wave: their plots hold more dots than n + m + 32 -- possibly more
// than the 65 535 the warp kernel's 16-bit group sizes can count -- so they always go to the CTA-per-task kernel.
void launch_k3_class(Handle* h, K3Params kp, int c, int max_nb, size_t slot, cudaStream_t st, bool allow_warp = true) {
    if (c < h->k3w_nclass && h->k3_mode == 1 && allow_warp) {
        kp.nb_cap = k3_class_cap[c]; kp.use_global = 0; kp.gscratch = nullptr; kp.queue = h->d_k3q.p + slot;
        const size_t smem = (size_t)K3W_TEAMS * k3w_scratch_words(kp.nb_cap) * sizeof(uint32_t);
        const int per_sm = std::max(1, std::min(K3W_MINB, (int)((size_t)(226 * 1024) / (smem + 1024))));
        const int grid = (int)std::min<int64_t>(((int64_t)kp.n_ids + K3W_TEAMS - 1) / K3W_TEAMS, (int64_t)h->sm_count * per_sm);
        k3w_score_reads<<<grid, 32 * K3W_TEAMS, smem, st>>>(kp);
    } else if (c < K3_NCLASS - 1) {
        kp.nb_cap = k3_class_cap[c]; kp.use_global = 0; kp.gscratch = nullptr;
        const size_t smem = k3_scratch_words(kp.nb_cap) * sizeof(uint32_t);
        k3_score_reads<<<kp.n_ids, K3_THREADS, smem, st>>>(kp);
    } else {
        kp.nb_cap = max_nb; kp.use_global = 1; kp.gscratch = h->d_gscratch.p;
        const int grid = std::min(kp.n_ids, 2 * h->sm_count);
        k3_score_reads<<<grid, K3_THREADS, 0, st>>>(kp);
    }
}

// kernel 1 over the operands of one wave; returns the launches made (0 or 1)
int launch_k1_wave(Handle* h, const Wave& w) {
    const int n_ops = (int)(w.op_end - w.op_begin);
    if (n_ops <= 0) return 0;
    const int base = h->chunk_prefix[w.op_begin], n_chunks = h->chunk_prefix[w.op_end] - base;
    if (n_chunks <= 0) return 0;
    k1_map_chunks<<<(n_ops + 255) / 256, 256, 0, h->stream>>>(h->d_chunk_prefix.p + w.op_begin, base, n_ops, h->d_k1map.p);
    k1_pack_kmers<<<n_chunks, K1_THREADS, 0, h->stream>>>(
        h->d_seq.p, h->d_ops.p + w.op_begin, h->d_chunk_prefix.p + w.op_begin, h->d_k1map.p, base, n_ops, h->d_hash.p, h->d_code.p,
        h->d_op_status.p + w.op_begin);
    return 2;
}

__global__ void k_sum_counts(const uint32_t* __restrict__ cnt, long long n, unsigned long long* out) {
    unsigned long long s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += cnt[i];
    #pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// A wave whose first pass overflowed some hit lists (repeats: more dots than n + m + 32): the tasks that own such a
// plot are scored again from copies of their plots with exact capacities in d_ovf_hits.  The resident plan
// (h->plots / d_plots) is never modified, so a later run() on the same batch starts from the same state.
int redo_wave(Handle* h, size_t wi, int64_t* launches) {
    const Wave& w = h->waves[wi];
    const int n_plots = (int)(w.plot_end - w.plot_begin);
    std::vector<uint32_t> cnt_host((size_t)n_plots);
    CK(cudaMemcpy(cnt_host.data(), h->d_cnt.p + w.plot_begin, n_plots * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::vector<int32_t> remap((size_t)n_plots, -1);
    std::vector<Plot> re_plots; std::vector<Task> re_tasks; std::vector<int32_t> re_out; std::vector<int64_t> re_prefix{0};
    const int rows = k2_variant_rows(h->tile_variant);
    int64_t extra = 0;
    h->ovf_loc.clear();
    for (int64_t t = w.task_begin; t < w.task_end; ++t) {
        const Task& tk = h->tasks[t];
        bool over = false;
        for (int i = 0; i < 4; ++i) if (tk.plot[i] >= 0) {
            const Plot& p = h->plots[tk.plot[i]];
            if (cnt_host[tk.plot[i] - w.plot_begin] > p.cap) over = true;
        }
        if (!over) continue;
        Task nt = tk;
        for (int i = 0; i < 4; ++i) if (tk.plot[i] >= 0) {
            const int32_t li = (int32_t)(tk.plot[i] - w.plot_begin);
            if (remap[li] < 0) {
                Plot p = h->plots[tk.plot[i]];
                if (cnt_host[li] > (1u << 30)) { h->err = "a plot produced more than 2^30 hits"; return VAPOR_E_CAPACITY; }
                p.cap = (uint32_t)align4(std::max<uint32_t>(cnt_host[li], 4u));
                p.hit_off = extra;
                h->ovf_loc[tk.plot[i]] = extra;
                extra += p.cap;
                re_prefix.push_back(re_prefix.back() + cut_strips(p, rows, nullptr));
                remap[li] = (int32_t)re_plots.size();
                re_plots.push_back(p);
            }
            nt.plot[i] = remap[li];
        }
        re_tasks.push_back(nt);
        re_out.push_back((int32_t)t);
    }
    h->tm.n_overflow_plots += (int64_t)re_plots.size();
    if (re_tasks.empty()) return VAPOR_OK;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if ((size_t)extra * sizeof(uint2) + ((size_t)1 << 30) > free_b + h->d_ovf_hits.cap * sizeof(uint2)) {
        h->err = "hit lists of repetitive plots exceed device memory"; return VAPOR_E_CAPACITY;
    }
    CK(h->d_ovf_hits.ensure((size_t)extra + 64));
    // group the re-scored tasks by kernel-3 scratch class
    std::vector<int32_t> ids[K3_NCLASS];
    int re_max_nb = 1;
    for (size_t i = 0; i < re_tasks.size(); ++i) {
        int nb = 1;
        for (int q = 0; q < 4; ++q) if (re_tasks[i].plot[q] >= 0) { const Plot& p = re_plots[re_tasks[i].plot[q]]; nb = std::max(nb, p.n + p.m - 1); }
        ids[k3_class_of(nb)].push_back((int32_t)i);
        re_max_nb = std::max(re_max_nb, nb);
    }
    std::vector<int32_t> id_flat, out_flat; int64_t id_off[K3_NCLASS + 1] = {0};
    for (int c = 0; c < K3_NCLASS; ++c) {
        for (int32_t i : ids[c]) { id_flat.push_back(i); out_flat.push_back(re_out[i]); }
        id_off[c + 1] = (int64_t)id_flat.size();
    }
    DevBuf<Plot> d_re; DevBuf<Task> d_rt; DevBuf<uint32_t> d_recnt, d_reflag; DevBuf<int64_t> d_repre; DevBuf<int32_t> d_ids, d_out;
    int rc = VAPOR_OK;
    auto cleanup = [&]() { d_re.release(); d_rt.release(); d_recnt.release(); d_reflag.release(); d_repre.release(); d_ids.release(); d_out.release(); };
    #define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return VAPOR_E_CUDA; } } while (0)
    CKR(d_re.ensure(re_plots.size())); CKR(d_rt.ensure(re_tasks.size())); CKR(d_recnt.ensure(re_plots.size())); CKR(d_reflag.ensure(4));
    CKR(d_repre.ensure(re_prefix.size())); CKR(d_ids.ensure(id_flat.size())); CKR(d_out.ensure(out_flat.size()));
    if (k3_class_of(re_max_nb) == K3_NCLASS - 1) CKR(h->d_gscratch.ensure((size_t)2 * h->sm_count * k3_scratch_words(std::max(re_max_nb, h->max_nb))));
    CKR(cudaMemcpyAsync(d_re.p, re_plots.data(), re_plots.size() * sizeof(Plot), cudaMemcpyHostToDevice, h->stream));
    CKR(cudaMemcpyAsync(d_rt.p, re_tasks.data(), re_tasks.size() * sizeof(Task), cudaMemcpyHostToDevice, h->stream));
    CKR(cudaMemcpyAsync(d_repre.p, re_prefix.data(), re_prefix.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    CKR(cudaMemcpyAsync(d_ids.p, id_flat.data(), id_flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CKR(cudaMemcpyAsync(d_out.p, out_flat.data(), out_flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CKR(cudaMemsetAsync(d_recnt.p, 0, re_plots.size() * sizeof(uint32_t), h->stream));
    CKR(cudaMemsetAsync(d_reflag.p, 0, 4 * sizeof(uint32_t), h->stream));
    CKR(cudaMemsetAsync(h->d_k3q.p + h->waves.size() * K3_NCLASS, 0, K3_NCLASS * sizeof(uint32_t), h->stream));
    if (h->waves.size() > 1) *launches += launch_k1_wave(h, w);      // later waves reused the word / code buffers
    CKR(cudaMemsetAsync(h->d_queue.p, 0, 4 * sizeof(unsigned long long), h->stream));
    int sp = span_begin(h, CAT_TILE);
    if (re_prefix.back() > 0) {
        K2Params kp{};
        kp.plots = d_re.p; kp.ops = h->d_ops.p; kp.strip_prefix = d_repre.p;
        kp.n_plots = (int)re_plots.size(); kp.n_strips = re_prefix.back(); kp.strip_base = 0;
        kp.hash = h->d_hash.p; kp.code = h->d_code.p; kp.hits = h->d_ovf_hits.p;
        kp.cnt = d_recnt.p; kp.queue = h->d_queue.p; kp.overflow = d_reflag.p;
        launch_k2(h, kp, h->tile_variant);
        ++*launches;
    }
    span_end(h, sp);
    sp = span_begin(h, CAT_SCORE);
    for (int c = 0; c < K3_NCLASS; ++c) {
        if (id_off[c + 1] == id_off[c]) continue;
        K3Params kp{};
        kp.tasks = d_rt.p; kp.task_ids = d_ids.p + id_off[c]; kp.out_ids = d_out.p + id_off[c]; kp.n_ids = (int)(id_off[c + 1] - id_off[c]);
        kp.plots = d_re.p; kp.cnt = d_recnt.p; kp.op_status = h->d_op_status.p; kp.hits = h->d_ovf_hits.p;
        kp.task_score = h->d_task_score.p; kp.task_status = h->d_task_status.p; kp.task_stat = h->d_task_stat.p;
        kp.task_hits = h->d_task_hits.p; kp.task_hitsum = h->d_task_hitsum.p;
        launch_k3_class(h, kp, c, std::max(re_max_nb, h->max_nb), h->waves.size() * K3_NCLASS + c, h->stream, /*allow_warp=*/false);
        ++*launches;
    }
    span_end(h, sp);
    CKR(cudaGetLastError());
    CKR(cudaStreamSynchronize(h->stream));
    #undef CKR
    cleanup();
    return rc;
}

// VAPOR_DEBUG_SYNC=1: synchronise after every launch so that a device fault names the kernel that raised it
#define CKL(name, bit)                                                                                \
    do {                                                                                              \
        if (h->debug_sync & (bit)) {                                                                        \
            cudaError_t e_ = cudaStreamSynchronize(h->stream);                                        \
            if (e_ != cudaSuccess) { h->err = std::string(name) + ": " + cudaGetErrorString(e_); return VAPOR_E_CUDA; } \
        }                                                                                             \
    } while (0)

int run_impl(Handle* h) {
    if (!h->resident) { h->err = "no resident batch: call vapor_gpu_upload first"; return VAPOR_E_STATE; }
    CK(cudaSetDevice(h->device));
    const size_t nsv = (size_t)h->n_sv;
    int64_t launches = 0;
    h->tm.n_overflow_plots = 0;
    h->ovf_loc.clear();
    h->spans.clear(); h->ev_used = 0;
    CK(cudaEventRecord(h->ev_run0, h->stream));

    CK(cudaMemsetAsync(h->d_op_status.p, 0, (h->ops.size() + 1) * sizeof(int32_t), h->stream));
    CK(cudaMemsetAsync(h->d_cnt.p, 0, (h->plots.size() + 1) * sizeof(uint32_t), h->stream));
    CK(cudaMemsetAsync(h->d_ovf_flags.p, 0, (h->waves.size() + 1) * sizeof(uint32_t), h->stream));
    CK(cudaMemsetAsync(h->d_stats.p, 0, 4 * sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->d_k3q.p, 0, (h->waves.size() + 1) * K3_NCLASS * sizeof(uint32_t), h->stream));

    // Kernel 3 of wave i runs on stream2 while kernels 1, 1b and 2 of wave i+1 run on stream: both sides are bound by
    // latency and instruction issue at half occupancy, so together they fill the SMs better than one after the other.
    // The hit slabs alternate between two buffers; kernel 2 of wave i waits for kernel 3 of wave i-2.
    const bool ovl = h->overlap && !h->debug_sync && h->waves.size() > 1 && h->d_hits2.p != nullptr;
    cudaStream_t s3 = ovl ? h->stream2 : h->stream;
    if (ovl) {
        while (h->ev_k2.size() < h->waves.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            h->ev_k2.push_back(a); h->ev_k3.push_back(b);
        }
    }
    int sp;
    for (size_t wi = 0; wi < h->waves.size(); ++wi) {
        const Wave& w = h->waves[wi];
        uint2* wave_hits = (ovl && (wi & 1)) ? h->d_hits2.p : h->d_hits.p;
        if (ovl && wi >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_k3[wi - 2], 0));
        // ---- kernel 1 (+ 1b: tables of the structure-side operands) of this wave's operands ---------------
        sp = span_begin(h, CAT_PACK);
        launches += launch_k1_wave(h, w);
        span_end(h, sp);
        CKL("k1_pack_kmers", 1);
        if (h->plan_mode == 1 && w.chunk_end > w.chunk_begin) {
            sp = span_begin(h, CAT_TABLE);
            k1b_build_tables<<<(unsigned)(w.chunk_end - w.chunk_begin), K1B_THREADS, 0, h->stream>>>(
                h->d_chunks.p + w.chunk_begin, h->d_ops.p, h->d_hash.p, h->d_table.p);
            ++launches;
            span_end(h, sp);
            CKL("k1b_build_tables", 2);
        }
        CK(cudaGetLastError());
        // ---- kernel 2 ---------------------------------------------------------------------
        sp = span_begin(h, CAT_TILE);
        if (h->plan_mode == 0) {
            const int n_plots = (int)(w.plot_end - w.plot_begin);
            const int64_t sbase = h->strip_prefix[w.plot_begin];
            const int64_t n_strips = h->strip_prefix[w.plot_end] - sbase;
            CK(cudaMemsetAsync(h->d_queue.p, 0, 4 * sizeof(unsigned long long), h->stream));
            if (n_strips > 0) {
                K2Params kp{};
                kp.plots = h->d_plots.p + w.plot_begin;
                kp.ops = h->d_ops.p;
                kp.strip_prefix = h->d_strip_prefix.p + w.plot_begin;
                kp.n_plots = n_plots;
                kp.n_strips = n_strips;
                kp.strip_base = sbase;
                kp.hash = h->d_hash.p; kp.code = h->d_code.p;
                kp.hits = wave_hits;
                kp.cnt = h->d_cnt.p + w.plot_begin;
                kp.queue = h->d_queue.p;
                kp.overflow = h->d_ovf_flags.p + wi;
                launch_k2(h, kp, h->plan_variant);
                ++launches;
            }
        } else {
            K2JParams kp{};
            kp.items = h->d_items.p; kp.jplots = h->d_jplots.p; kp.chunks = h->d_chunks.p; kp.plots = h->d_plots.p;
            kp.ops = h->d_ops.p; kp.hash = h->d_hash.p; kp.code = h->d_code.p; kp.table = h->d_table.p;
            kp.hits = wave_hits; kp.cnt = h->d_cnt.p; kp.overflow = h->d_ovf_flags.p + wi; kp.qc = nullptr;
            kp.evaluated = h->d_stats.p + 1;
            launches += launch_k2_join(h, h->jp, wi, kp);
        }
        span_end(h, sp);
        CK(cudaGetLastError());
        CKL("kernel 2", 4);
        // ---- kernel 3, one launch per scratch class ----------------------------------------------
        if (ovl) { CK(cudaEventRecord(h->ev_k2[wi], h->stream)); CK(cudaStreamWaitEvent(s3, h->ev_k2[wi], 0)); }
        const int n_warp_classes = h->k3_mode == 1 ? h->k3w_nclass : 0;       // the warp kernel's launches are timed apart
        sp = span_begin(h, n_warp_classes > 0 ? CAT_SCORE_W : CAT_SCORE, s3);
        for (int c = 0; c < K3_NCLASS; ++c) {
            if (c == n_warp_classes && c > 0) { span_end(h, sp, s3); sp = span_begin(h, CAT_SCORE, s3); }
            const int64_t o0 = h->class_off[wi * K3_NCLASS + c], o1 = h->class_off[wi * K3_NCLASS + c + 1];
            if (o1 == o0) continue;
            K3Params kp{};
            kp.tasks = h->d_tasks.p; kp.task_ids = h->d_class_ids.p + o0; kp.out_ids = nullptr; kp.n_ids = (int)(o1 - o0);
            kp.plots = h->d_plots.p; kp.cnt = h->d_cnt.p; kp.op_status = h->d_op_status.p;
            kp.hits = wave_hits;
            kp.task_score = h->d_task_score.p; kp.task_status = h->d_task_status.p; kp.task_stat = h->d_task_stat.p;
            kp.task_hits = h->d_task_hits.p; kp.task_hitsum = h->d_task_hitsum.p;
            launch_k3_class(h, kp, c, h->max_nb, wi * K3_NCLASS + c, s3);
            ++launches;
            CKL("k3_score_reads", 8);
        }
        span_end(h, sp, s3);
        if (ovl) CK(cudaEventRecord(h->ev_k3[wi], s3));
        CK(cudaGetLastError());
    }
    if (ovl) {                                       // the summaries wait for the last two waves' scores
        const size_t nw = h->waves.size();
        CK(cudaStreamWaitEvent(h->stream, h->ev_k3[nw - 1], 0));
        if (nw >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_k3[nw - 2], 0));
    }
    // ---- kernel 4 -------------------------------------------------------------------------
    auto genotype = [&]() -> int {
        int s4 = span_begin(h, CAT_GENO);
        if (nsv) {
            k4_genotype<<<(unsigned)((nsv + 127) / 128), 128, 0, h->stream>>>(
                h->d_sv_off.p, (int)nsv, h->d_task_score.p, h->d_task_status.p, h->d_pos.p,
                h->d_sv_qs.p, h->d_sv_gs.p, h->d_sv_gq.p, h->d_sv_gt.p, h->d_sv_nscore.p);
            ++launches;
        }
        span_end(h, s4);
        CK(cudaGetLastError());
        return VAPOR_OK;
    };
    { int rc = genotype(); if (rc) return rc; }
    if (!h->plots.empty()) {
        k_sum_counts<<<std::min<int64_t>(((int64_t)h->plots.size() + 255) / 256, 4 * h->sm_count), 256, 0, h->stream>>>(
            h->d_cnt.p, (long long)h->plots.size(), h->d_stats.p);
        ++launches;
    }
    CK(cudaEventRecord(h->ev_run1, h->stream));
    // one host round trip per run: overflow flags of all waves + the run's counters
    CK(cudaMemcpyAsync(h->h_flags, h->d_ovf_flags.p, (h->waves.size() + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_stats, h->d_stats.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float acc[CAT_N] = {0};
    collect_spans(h, acc);
    float tot = 0; cudaEventElapsedTime(&tot, h->ev_run0, h->ev_run1);
    bool any_ovf = false;
    for (size_t wi = 0; wi < h->waves.size(); ++wi) any_ovf |= h->h_flags[wi] != 0;
    if (any_ovf) {                                   // rare: repeats.  Re-score the affected tasks, then the summaries again
        CK(cudaEventRecord(h->ev_run0, h->stream));
        for (size_t wi = 0; wi < h->waves.size(); ++wi) {
            if (!h->h_flags[wi]) continue;
            int rc = redo_wave(h, wi, &launches);
            if (rc) { h->spans.clear(); h->ev_used = 0; return rc; }
        }
        { int rc = genotype(); if (rc) return rc; }
        CK(cudaEventRecord(h->ev_run1, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        collect_spans(h, acc);
        float t2 = 0; cudaEventElapsedTime(&t2, h->ev_run0, h->ev_run1);
        tot += t2;
    }
    h->tm.pack_ms = acc[CAT_PACK]; h->tm.table_ms = acc[CAT_TABLE]; h->tm.tile_ms = acc[CAT_TILE];
    h->tm.score_ms = acc[CAT_SCORE] + acc[CAT_SCORE_W]; h->tm.score_warp_ms = acc[CAT_SCORE_W]; h->tm.genotype_ms = acc[CAT_GENO];
    h->tm.total_ms = tot;
    h->tm.launches = launches;
    h->tm.hits = (int64_t)h->h_stats[0];
    h->tm.evaluated_cells = h->plan_mode == 0 ? h->tm.padded_cells : (int64_t)h->h_stats[1];
    h->ran = true;
    return VAPOR_OK;
}

int fetch_impl(Handle* h, vapor_out_t* out) {
    if (!h->ran) { h->err = "nothing to fetch: call vapor_gpu_run first"; return VAPOR_E_STATE; }
    if (!out) { h->err = "NULL out"; return VAPOR_E_ARG; }
    CK(cudaSetDevice(h->device));
    const size_t nt = (size_t)h->n_task, nsv = (size_t)h->n_sv;
    int sp = span_begin(h, CAT_D2H);
    if (nt) {
        if (out->task_score) CK(cudaMemcpyAsync(out->task_score, h->d_task_score.p, nt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (out->task_status) CK(cudaMemcpyAsync(out->task_status, h->d_task_status.p, nt, cudaMemcpyDeviceToHost, h->stream));
        if (out->task_stat) CK(cudaMemcpyAsync(out->task_stat, h->d_task_stat.p, 4 * nt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (out->task_hits) CK(cudaMemcpyAsync(out->task_hits, h->d_task_hits.p, 4 * nt * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        if (out->task_hitsum) CK(cudaMemcpyAsync(out->task_hitsum, h->d_task_hitsum.p, 4 * nt * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    }
    if (nsv) {
        if (out->sv_qs) CK(cudaMemcpyAsync(out->sv_qs, h->d_sv_qs.p, nsv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (out->sv_gs) CK(cudaMemcpyAsync(out->sv_gs, h->d_sv_gs.p, nsv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (out->sv_gq) CK(cudaMemcpyAsync(out->sv_gq, h->d_sv_gq.p, nsv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (out->sv_gt) CK(cudaMemcpyAsync(out->sv_gt, h->d_sv_gt.p, nsv, cudaMemcpyDeviceToHost, h->stream));
        if (out->sv_nscore) CK(cudaMemcpyAsync(out->sv_nscore, h->d_sv_nscore.p, nsv * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    span_end(h, sp);
    CK(cudaStreamSynchronize(h->stream));
    float acc[CAT_N] = {0};
    collect_spans(h, acc);
    h->tm.d2h_ms = acc[CAT_D2H];
    return VAPOR_OK;
}

// ---- integer issue-rate microbenchmarks ------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) k_int_peak(const uint32_t* __restrict__ in, uint32_t* out, int iters) {
    uint32_t r[32];
    #pragma unroll
    for (int q = 0; q < 32; ++q) r[q] = in[(threadIdx.x + q * 7) & 1023];
    uint32_t v = in[threadIdx.x & 1023];
    if (WHICH == 0) {
        bool p0 = false, p1 = false, p2 = false, p3 = false;
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int q = 0; q < 32; q += 4) {
                p0 |= (r[q] == v); p1 |= (r[q + 1] == v); p2 |= (r[q + 2] == v); p3 |= (r[q + 3] == v);
            }
            v += 0x9E3779B9u;
        }
        if (p0 | p1 | p2 | p3) out[0] = v;
    } else if (WHICH == 1) {
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = (r[q] ^ v) & (r[(q + 1) & 31] | v);
            v += 0x9E3779B9u;
        }
        uint32_t a = 0;
        #pragma unroll
        for (int q = 0; q < 32; ++q) a ^= r[q];
        if (a == 0x12345678u) out[0] = a;
    } else if (WHICH == 3) {
        // both integer pipes at once: 16 independent LOP3 chains (alu pipe) + 16 independent IMAD chains (fma pipe)
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int q = 0; q < 16; ++q) {
                r[q] = (r[q] ^ v) & (r[(q + 1) & 15] | v);
                r[16 + q] = r[16 + q] * v + r[16 + ((q + 3) & 15)];
            }
            v += 0x9E3779B9u;
        }
        uint32_t a = 0;
        #pragma unroll
        for (int q = 0; q < 32; ++q) a ^= r[q];
        if (a == 0x12345678u) out[0] = a;
    } else if (WHICH == 5) {
        // fma pipe alone: 32 independent IMAD chains, the streamed word first
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = v * r[q] + r[(q + 3) & 31];
            v += 0x9E3779B9u;
        }
        uint32_t a = 0;
        #pragma unroll
        for (int q = 0; q < 32; ++q) a ^= r[q];
        if (a == 0x12345678u) out[0] = a;
    } else if (WHICH == 6) {
        // both pipes, the instruction mix of the tile kernel but none of its structure: 16 independent compare-accumulates
        // (ISETP, alu pipe) + 16 independent multiply-adds (IMAD, fma pipe) per step, the streamed word first in each
        bool p0 = false, p1 = false, p2 = false, p3 = false;
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int q = 0; q < 16; q += 4) {
                p0 |= (v == r[q]); p1 |= (v == r[q + 1]); p2 |= (v == r[q + 2]); p3 |= (v == r[q + 3]);
            }
            #pragma unroll
            for (int q = 16; q < 32; ++q) r[q] = v * r[q] + r[16 + ((q + 3) & 15)];
            v += 0x9E3779B9u;
        }
        uint32_t a = 0;
        #pragma unroll
        for (int q = 16; q < 32; ++q) a ^= r[q];
        if ((p0 | p1 | p2 | p3) && a == 0x12345678u) out[0] = a;
    } else {
        for (int it = 0; it < iters; ++it) {
            #pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = r[q] + v + r[(q + 5) & 31];
            v += 0x9E3779B9u;
        }
        uint32_t a = 0;
        #pragma unroll
        for (int q = 0; q < 32; ++q) a ^= r[q];
        if (a == 0x12345678u) out[0] = a;
    }
}

// The tile kernel's inner loop in isolation (no queue, no TMA, no exact path): 14 compare-accumulates (alu pipe) +
// 2 x 8 Horner steps (fma pipe) + 2 zero tests per shared word, words fetched with LDS.128, one vote per 32 words,
// the shared word first in every instruction.  32 integer instructions per lane and word.
__global__ void __launch_bounds__(128) k_tile_loop_peak(const uint32_t* __restrict__ in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint32_t s[4][2048];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < 2048; i += 32) s[warp][i] = in[(i * 7 + warp) & 1023] | 1u;
    uint32_t r[14], c[2][8];
    #pragma unroll
    for (int q = 0; q < 14; ++q) r[q] = in[(threadIdx.x * 14 + q + 3 * blockIdx.x) & 1023] & ~1u;
    #pragma unroll
    for (int p = 0; p < 2; ++p)
        #pragma unroll
        for (int q = 0; q < 8; ++q) c[p][q] = in[(threadIdx.x * 16 + p * 8 + q + 5 * blockIdx.x) & 1023] | 1u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < 2048 / 32; ++b) {
            const uint32_t* sblk = s[warp] + b * 32;
            bool p0 = false, p1 = false, p2 = false, p3 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v4 = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t v = vw[e];
                    uint32_t a0 = v + c[0][0], a1 = v + c[1][0];
                    #pragma unroll
                    for (int q = 1; q < 8; ++q) { a0 = v * a0 + c[0][q]; a1 = v * a1 + c[1][q]; }
                    #pragma unroll
                    for (int q = 0; q < 14; q += 2) { p0 |= (v == r[q]); p1 |= (v == r[q + 1]); }
                    p2 |= (a0 == 0u); p3 |= (a1 == 0u);
                }
            }
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, p0 | p1 | p2 | p3);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

}  // namespace

// ==============================================================================================
// extern "C"
// ==============================================================================================
extern "C" {

int vapor_b200_abi_version(void) { return VAPOR_B200_ABI_VERSION; }

int vapor_gpu_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

uint64_t vapor_hit_mix(uint32_t x, uint32_t y) { return hit_mix(x, y); }

int vapor_gpu_pci_bus_id(int device, char* buf, int len) {
    if (!buf || len < 13) return VAPOR_E_ARG;
    return cudaDeviceGetPCIBusId(buf, len, device) == cudaSuccess ? VAPOR_OK : VAPOR_E_CUDA;
}

const char* vapor_gpu_last_error(void* handle) {
    if (!handle) return g_open_error.c_str();
    return static_cast<Handle*>(handle)->err.c_str();
}

int vapor_gpu_open(int device, void** handle) {
    if (!handle) { g_open_error = "NULL handle pointer"; return VAPOR_E_ARG; }
    *handle = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_open_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); vapor_b200 has no CPU fallback";
        return VAPOR_E_CUDA;
    }
    if (device < 0 || device >= ndev) { g_open_error = "device index out of range"; return VAPOR_E_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_open_error = cudaGetErrorString(e); return VAPOR_E_CUDA; }
    Handle* h = new Handle();
    h->device = device;
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, device);
    h->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > 0)
            h->default_hit_budget = std::max<int64_t>((int64_t)1 << 30, std::min<int64_t>((int64_t)16 << 30, (int64_t)(free_b / 8)));
    }
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_open_error = cudaGetErrorString(e); delete h; return VAPOR_E_CUDA; }
    build_lut();
    e = cudaMemcpyToSymbol(c_code_lut, host_code_lut, 256);
    if (e != cudaSuccess) { g_open_error = std::string("kernel image not loadable on this device: ") + cudaGetErrorString(e); cudaStreamDestroy(h->stream); delete h; return VAPOR_E_CUDA; }
    e = cudaFuncSetAttribute(k3_score_reads, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(k3_scratch_words(k3_class_cap[K3_NCLASS - 2]) * sizeof(uint32_t)));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3w_score_reads, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)(K3W_TEAMS * k3w_scratch_words(K3W_MAX_NB) * sizeof(uint32_t)));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_join_match<K2J_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2j_class_cap[K2J_NCLASS - 1]);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2_join_match<K2J_WARPS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2j_class_cap[K2J_NCLASS - 1]);
    if (e != cudaSuccess) { g_open_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); cudaStreamDestroy(h->stream); delete h; return VAPOR_E_CUDA; }
    e = cudaEventCreate(&h->ev_run0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev_run1);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_stats, 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) { g_open_error = std::string("event / pinned allocation: ") + cudaGetErrorString(e); cudaStreamDestroy(h->stream); delete h; return VAPOR_E_CUDA; }
    *handle = h;
    return VAPOR_OK;
}

int vapor_gpu_close(void* handle) {
    if (!handle) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    h->d_seq.release(); h->d_code.release(); h->d_task_status.release(); h->d_sv_gt.release(); h->d_table.release();
    h->d_ops.release(); h->d_plots.release(); h->d_tasks.release(); h->d_chunks.release(); h->d_items.release();
    h->d_chunk_prefix.release(); h->d_op_status.release(); h->d_class_ids.release(); h->d_sv_nscore.release(); h->d_jplots.release(); h->d_k1map.release();
    h->d_strip_prefix.release(); h->d_sv_off.release();
    h->d_hash.release(); h->d_cnt.release(); h->d_task_hits.release(); h->d_gscratch.release(); h->d_ovf_flags.release(); h->d_qc.release(); h->d_k3q.release();
    h->d_hits.release(); h->d_hits2.release(); h->d_ovf_hits.release();
    for (cudaEvent_t e_ : h->ev_k2) cudaEventDestroy(e_);
    for (cudaEvent_t e_ : h->ev_k3) cudaEventDestroy(e_);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    h->d_task_score.release(); h->d_task_stat.release(); h->d_pos.release(); h->d_sv_qs.release(); h->d_sv_gs.release(); h->d_sv_gq.release();
    h->d_task_hitsum.release(); h->d_queue.release(); h->d_stats.release();
    if (h->h_stats) cudaFreeHost(h->h_stats);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->ev_run0) cudaEventDestroy(h->ev_run0);
    if (h->ev_run1) cudaEventDestroy(h->ev_run1);
    for (auto& ev : h->ev_pool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    cudaStreamDestroy(h->stream);
    delete h;
    return VAPOR_OK;
}

int vapor_gpu_set_hit_budget(void* handle, int64_t bytes) {
    if (!handle) return VAPOR_E_ARG;
    static_cast<Handle*>(handle)->hit_budget = bytes;
    return VAPOR_OK;
}

int vapor_gpu_set_option(void* handle, const char* name, int64_t value) {
    if (!handle || !name) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    const std::string n(name);
    if (n == "hit_budget_bytes") { h->hit_budget = value; return VAPOR_OK; }
    if (n == "tile_variant") {
        if (value < 0 || value >= K2_NVARIANT) { h->err = "tile_variant out of range"; return VAPOR_E_ARG; }
        h->tile_variant = (int)value; h->resident = false; h->ran = false;       // strips must be re-cut
        return VAPOR_OK;
    }
    if (n == "k2_mode") {
        if (value < 0 || value > 1) { h->err = "k2_mode must be 0 (tile) or 1 (join)"; return VAPOR_E_ARG; }
        h->k2_mode = (int)value; h->resident = false; h->ran = false;             // the kernel-2 plan must be rebuilt
        return VAPOR_OK;
    }
    if (n == "k3_warp_classes") {
        if (value < 0 || value > K3W_NCLASS) { h->err = "k3_warp_classes out of range"; return VAPOR_E_ARG; }
        h->k3w_nclass = (int)value; return VAPOR_OK;
    }
    if (n == "overlap") {
        if (value < 0 || value > 1) { h->err = "overlap must be 0 or 1"; return VAPOR_E_ARG; }
        h->overlap = (int)value; h->resident = false; h->ran = false;             // the second hit slab is allocated at upload
        return VAPOR_OK;
    }
    if (n == "k3_mode") {
        if (value < 0 || value > 1) { h->err = "k3_mode must be 0 (CTA per task) or 1 (warp per task)"; return VAPOR_E_ARG; }
        h->k3_mode = (int)value; return VAPOR_OK;
    }
    if (n == "plan_threads") {
        if (value < 0 || value > 256) { h->err = "plan_threads out of range"; return VAPOR_E_ARG; }
        h->plan_threads = (int)value; return VAPOR_OK;
    }
    if (n == "k2_ctas_per_sm") {
        if (value < 0 || value > 32) { h->err = "k2_ctas_per_sm out of range"; return VAPOR_E_ARG; }
        h->k2_ctas_per_sm = (int)value; return VAPOR_OK;
    }
    h->err = "unknown option: " + n;
    return VAPOR_E_ARG;
}

int vapor_gpu_upload(void* handle, const vapor_batch_t* in) {
    if (!handle) return VAPOR_E_ARG;
    return upload_impl(static_cast<Handle*>(handle), in);
}
int vapor_gpu_run(void* handle) {
    if (!handle) return VAPOR_E_ARG;
    return run_impl(static_cast<Handle*>(handle));
}
int vapor_gpu_fetch(void* handle, vapor_out_t* out) {
    if (!handle) return VAPOR_E_ARG;
    return fetch_impl(static_cast<Handle*>(handle), out);
}
int vapor_gpu_score(void* handle, const vapor_batch_t* in, vapor_out_t* out) {
    if (!handle) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    int rc = upload_impl(h, in);
    if (rc) return rc;
    rc = run_impl(h);
    if (rc) return rc;
    return fetch_impl(h, out);
}

int vapor_gpu_last_timings(void* handle, vapor_timings_t* t) {
    if (!handle || !t) return VAPOR_E_ARG;
    *t = static_cast<Handle*>(handle)->tm;
    return VAPOR_OK;
}

int vapor_gpu_dotdata(void* handle, int k, const uint8_t* read, int64_t read_len,
                      const uint8_t* structure, int64_t struct_len,
                      int32_t* xy, int64_t cap, int64_t* n_hits)
{
    if (!handle) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    if (!n_hits || read_len < 0 || struct_len < 0 || (cap > 0 && !xy)) { h->err = "bad dotdata arguments"; return VAPOR_E_ARG; }
    // one task in W10 mode (no upper-casing) over (read, structure, structure): plot 0 is the wanted plot
    std::vector<uint8_t> seq((size_t)(read_len + struct_len));
    if (read_len) memcpy(seq.data(), read, (size_t)read_len);
    if (struct_len) memcpy(seq.data() + read_len, structure, (size_t)struct_len);
    int64_t off[3] = {0, read_len, read_len + struct_len};
    int32_t tr = 0, tf = 1, ta = 1, tmiss = 0;
    uint8_t tk = (uint8_t)k, tm_ = VAPOR_MODE_W10;
    int64_t sv_off[2] = {0, 1};
    vapor_batch_t b{};
    b.seq_bytes = seq.data(); b.seq_off = off; b.n_seq = 2; b.n_task = 1;
    b.task_read = &tr; b.task_ref = &tf; b.task_alt = &ta; b.task_miss = &tmiss; b.task_k = &tk; b.task_mode = &tm_;
    b.n_sv = 1; b.sv_task_off = sv_off;
    if (k < 1 || k > K1_MAXK) { h->err = "window_size k must be 1..40"; return VAPOR_E_ARG; }
    int rc = upload_impl(h, &b);
    if (rc) return rc;
    rc = run_impl(h);
    if (rc) return rc;
    int32_t st = 0;
    CK(cudaMemcpy(&st, h->d_op_status.p + h->tasks[0].read_op, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (st) { h->err = "BADREAD: read holds a character outside ACGTN/acgtn after key_modify (reference raises KeyError)"; return VAPOR_E_ARG; }
    const Plot& p = h->plots[h->tasks[0].plot[0]];
    uint32_t cnt = 0;
    CK(cudaMemcpy(&cnt, h->d_cnt.p + h->tasks[0].plot[0], sizeof(uint32_t), cudaMemcpyDeviceToHost));
    *n_hits = cnt;
    std::vector<uint2> hits(cnt);
    const auto ovf = h->ovf_loc.find(h->tasks[0].plot[0]);      // a repetitive plot was re-run into the overflow buffer
    const uint2* src = ovf != h->ovf_loc.end() ? h->d_ovf_hits.p + ovf->second : h->d_hits.p + p.hit_off;
    if (cnt) CK(cudaMemcpy(hits.data(), src, cnt * sizeof(uint2), cudaMemcpyDeviceToHost));
    for (auto& v : hits) v.y &= HIT_Y_MASK;
    std::sort(hits.begin(), hits.end(), [](const uint2& a, const uint2& b2) { return a.x != b2.x ? a.x < b2.x : a.y < b2.y; });
    const int64_t nw = std::min<int64_t>(cap, cnt);
    for (int64_t i = 0; i < nw; ++i) { xy[2 * i] = (int32_t)hits[i].x; xy[2 * i + 1] = (int32_t)hits[i].y; }
    return VAPOR_OK;
}

int vapor_gpu_selfplot_qc(void* handle, const uint8_t* seq_bytes, const int64_t* seq_off, int64_t n_seq,
                          const uint8_t* k, int64_t* out)
{
    if (!handle) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    if (n_seq < 0 || (n_seq > 0 && (!seq_off || !k || !out))) { h->err = "bad selfplot_qc arguments"; return VAPOR_E_ARG; }
    if (n_seq == 0) return VAPOR_OK;
    const int64_t total = seq_off[n_seq];
    if (seq_off[0] != 0 || total < 0 || (total > 0 && !seq_bytes)) { h->err = "bad selfplot_qc arguments"; return VAPOR_E_ARG; }
    CK(cudaSetDevice(h->device));
    h->resident = false; h->ran = false;                       // plan and device buffers are reused
    // one read-role operand per sequence serves both axes of its self-plot
    std::vector<Operand> ops((size_t)n_seq);
    std::vector<Plot> plots((size_t)n_seq);
    std::vector<int32_t> chunk_prefix((size_t)n_seq + 1, 0);
    std::vector<int64_t> strip_prefix((size_t)n_seq + 1, 0);
    const int mode = h->k2_mode;
    const int k2_rows = k2_variant_rows(h->tile_variant);
    int64_t hash_off = 0, code_off = 0;
    for (int64_t i = 0; i < n_seq; ++i) {
        const int64_t L = seq_off[i + 1] - seq_off[i];
        if (L < 0 || L >= (1ll << 27)) { h->err = "sequence length out of range (0 .. 2^27)"; return VAPOR_E_ARG; }
        if (k[i] < 1 || k[i] > K1_MAXK) { h->err = "window_size k must be 1..40"; return VAPOR_E_ARG; }
        Operand o{};
        o.seq_begin = seq_off[i]; o.len = (int32_t)L; o.k = k[i]; o.flags = OPF_READ | OPF_TABLE;
        o.n = std::max(0, o.len - o.k + 1);
        o.hash_off = hash_off; o.code_off = code_off;
        hash_off += align4(o.n) + 8; code_off += (o.len + 8 + 15) & ~15;
        ops[i] = o;
        Plot p{};
        p.read_op = p.struct_op = (int32_t)i; p.miss = 0; p.n = p.m = o.n; p.cap = 0; p.kind = PLOT_QC; p.hit_off = i;
        if (mode == 0) strip_prefix[i + 1] = strip_prefix[i] + cut_strips(p, k2_rows, nullptr);
        plots[i] = p;
        chunk_prefix[i + 1] = chunk_prefix[i] + std::max(1, (o.len + K1_CHUNK - 1) / K1_CHUNK);
    }
    const size_t ns = (size_t)n_seq;
    std::vector<TabChunk> chunks; JoinPlan jp; int64_t table_bytes = 0;
    if (mode == 1) {
        std::vector<int32_t> op_chunk0;
        Wave w1{};
        w1.plot_end = (int64_t)ns; w1.op_end = (int64_t)ns;
        std::vector<Wave> one{w1};
        table_bytes = build_chunks(ops, one, chunks, op_chunk0);
        build_join_items(plots, one, chunks, op_chunk0, jp);
    }
    CK(h->d_seq.ensure((size_t)total + 64)); CK(h->d_ops.ensure(ns + 1)); CK(h->d_plots.ensure(ns + 1));
    CK(h->d_chunk_prefix.ensure(ns + 1));
    CK(h->d_hash.ensure((size_t)hash_off + 16)); CK(h->d_code.ensure((size_t)code_off + 64));
    CK(h->d_op_status.ensure(ns + 1)); CK(h->d_cnt.ensure(ns + 1)); CK(h->d_qc.ensure(ns * QC_WORDS));
    CK(h->d_queue.ensure(4)); CK(h->d_ovf_flags.ensure(4)); CK(h->d_stats.ensure(4));
    if (total > 0) CK(cudaMemcpyAsync(h->d_seq.p, seq_bytes, (size_t)total, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_ops.p, ops.data(), ns * sizeof(Operand), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_plots.p, plots.data(), ns * sizeof(Plot), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_chunk_prefix.p, chunk_prefix.data(), (ns + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    if (mode == 0) {
        CK(h->d_strip_prefix.ensure(ns + 1));
        CK(cudaMemcpyAsync(h->d_strip_prefix.p, strip_prefix.data(), (ns + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    } else if (!chunks.empty()) {
        CK(h->d_chunks.ensure(chunks.size() + 1)); CK(h->d_items.ensure(jp.items.size() + 1));
        CK(h->d_jplots.ensure(jp.jplots.size() + 1)); CK(h->d_table.ensure((size_t)table_bytes + 64));
        CK(cudaMemcpyAsync(h->d_chunks.p, chunks.data(), chunks.size() * sizeof(TabChunk), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_items.p, jp.items.data(), jp.items.size() * sizeof(JoinItem), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_jplots.p, jp.jplots.data(), jp.jplots.size() * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    }
    std::vector<uint32_t> qc_init(ns * QC_WORDS, 0u);
    for (size_t i = 0; i < ns; ++i) { qc_init[i * QC_WORDS + 3] = 0xFFFFFFFFu; qc_init[i * QC_WORDS + 5] = 0xFFFFFFFFu; }
    CK(cudaMemcpyAsync(h->d_qc.p, qc_init.data(), qc_init.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_op_status.p, 0, (ns + 1) * sizeof(int32_t), h->stream));
    CK(cudaMemsetAsync(h->d_cnt.p, 0, (ns + 1) * sizeof(uint32_t), h->stream));
    CK(cudaMemsetAsync(h->d_queue.p, 0, 4 * sizeof(unsigned long long), h->stream));
    CK(cudaMemsetAsync(h->d_ovf_flags.p, 0, 4 * sizeof(uint32_t), h->stream));
    CK(cudaMemsetAsync(h->d_stats.p, 0, 4 * sizeof(unsigned long long), h->stream));
    CK(h->d_k1map.ensure((size_t)chunk_prefix.back() + 1));
    if (chunk_prefix.back() > 0) {
        k1_map_chunks<<<(unsigned)((ns + 255) / 256), 256, 0, h->stream>>>(h->d_chunk_prefix.p, 0, (int)ns, h->d_k1map.p);
        k1_pack_kmers<<<chunk_prefix.back(), K1_THREADS, 0, h->stream>>>(
            h->d_seq.p, h->d_ops.p, h->d_chunk_prefix.p, h->d_k1map.p, 0, (int)ns, h->d_hash.p, h->d_code.p, h->d_op_status.p);
    }
    CK(cudaGetLastError());
    if (mode == 0) {
        if (strip_prefix.back() > 0) {
            K2Params kp{};
            kp.plots = h->d_plots.p; kp.ops = h->d_ops.p; kp.strip_prefix = h->d_strip_prefix.p;
            kp.n_plots = (int)ns; kp.n_strips = strip_prefix.back(); kp.strip_base = 0;
            kp.hash = h->d_hash.p; kp.code = h->d_code.p; kp.hits = nullptr; kp.cnt = h->d_cnt.p;
            kp.queue = h->d_queue.p; kp.overflow = h->d_ovf_flags.p; kp.qc = h->d_qc.p;
            launch_k2(h, kp, h->tile_variant);
        }
    } else if (!chunks.empty()) {
        k1b_build_tables<<<(unsigned)chunks.size(), K1B_THREADS, 0, h->stream>>>(h->d_chunks.p, h->d_ops.p, h->d_hash.p, h->d_table.p);
        K2JParams kp{};
        kp.items = h->d_items.p; kp.jplots = h->d_jplots.p; kp.chunks = h->d_chunks.p; kp.plots = h->d_plots.p;
        kp.ops = h->d_ops.p; kp.hash = h->d_hash.p; kp.code = h->d_code.p; kp.table = h->d_table.p;
        kp.hits = nullptr; kp.cnt = h->d_cnt.p; kp.overflow = h->d_ovf_flags.p; kp.qc = h->d_qc.p;
        kp.evaluated = h->d_stats.p + 1;
        launch_k2_join(h, jp, 0, kp);
    }
    CK(cudaGetLastError());
    std::vector<int32_t> st(ns);
    CK(cudaMemcpyAsync(qc_init.data(), h->d_qc.p, qc_init.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(st.data(), h->d_op_status.p, ns * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < ns; ++i) {
        for (int j = 0; j < 7; ++j) out[i * 8 + j] = (int64_t)qc_init[i * QC_WORDS + j];
        out[i * 8 + 7] = st[i] ? VAPOR_ST_BADREAD : VAPOR_ST_SCORED;
    }
    return VAPOR_OK;
}

int vapor_gpu_summarize(void* handle, const double* scores, const int64_t* sv_off, int64_t n_sv,
                        double* sv_qs, double* sv_gs, double* sv_gq, uint8_t* sv_gt, int32_t* sv_nscore)
{
    if (!handle) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    if (n_sv < 0 || (n_sv > 0 && !sv_off)) { h->err = "bad summarize arguments"; return VAPOR_E_ARG; }
    if (n_sv == 0) return VAPOR_OK;
    const int64_t n = sv_off[n_sv];
    if (sv_off[0] != 0 || n < 0 || (n > 0 && !scores)) { h->err = "bad summarize arguments"; return VAPOR_E_ARG; }
    CK(cudaSetDevice(h->device));
    h->resident = false; h->ran = false;            // the task buffers are reused
    const size_t nt = (size_t)n, nsv = (size_t)n_sv;
    CK(h->d_task_score.ensure(nt + 1)); CK(h->d_task_status.ensure(nt + 1)); CK(h->d_pos.ensure(nt + 1));
    CK(h->d_sv_off.ensure(nsv + 1));
    CK(h->d_sv_qs.ensure(nsv + 1)); CK(h->d_sv_gs.ensure(nsv + 1)); CK(h->d_sv_gq.ensure(nsv + 1));
    CK(h->d_sv_gt.ensure(nsv + 1)); CK(h->d_sv_nscore.ensure(nsv + 1));
    if (nt) CK(cudaMemcpyAsync(h->d_task_score.p, scores, nt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_task_status.p, 1, nt + 1, h->stream));
    CK(cudaMemcpyAsync(h->d_sv_off.p, sv_off, (nsv + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    k4_genotype<<<(unsigned)((nsv + 127) / 128), 128, 0, h->stream>>>(
        h->d_sv_off.p, (int)nsv, h->d_task_score.p, h->d_task_status.p, h->d_pos.p,
        h->d_sv_qs.p, h->d_sv_gs.p, h->d_sv_gq.p, h->d_sv_gt.p, h->d_sv_nscore.p);
    CK(cudaGetLastError());
    if (sv_qs) CK(cudaMemcpyAsync(sv_qs, h->d_sv_qs.p, nsv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (sv_gs) CK(cudaMemcpyAsync(sv_gs, h->d_sv_gs.p, nsv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (sv_gq) CK(cudaMemcpyAsync(sv_gq, h->d_sv_gq.p, nsv * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (sv_gt) CK(cudaMemcpyAsync(sv_gt, h->d_sv_gt.p, nsv, cudaMemcpyDeviceToHost, h->stream));
    if (sv_nscore) CK(cudaMemcpyAsync(sv_nscore, h->d_sv_nscore.p, nsv * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VAPOR_OK;
}

int vapor_gpu_host_alloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) return VAPOR_E_ARG;
    cudaError_t e = cudaMallocHost(p, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { g_open_error = cudaGetErrorString(e); return VAPOR_E_CUDA; }
    return VAPOR_OK;
}
int vapor_gpu_host_free(void* p) {
    if (!p) return VAPOR_OK;
    return cudaFreeHost(p) == cudaSuccess ? VAPOR_OK : VAPOR_E_CUDA;
}

int vapor_host_plan(const vapor_batch_t* in, int k2_mode, int threads, int64_t wave_budget_bytes, double* ms, uint64_t* digest, int64_t* counts) {
    Handle h;                                    // no device is touched: planning is host work
    h.k2_mode = k2_mode ? 1 : 0; h.plan_threads = threads; h.hit_budget = wave_budget_bytes;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = plan_batch(&h, in);
    if (ms) *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc) { g_open_error = h.err; return rc; }
    if (digest) {                                // FNV-1a over everything the kernels will read
        uint64_t d = 1469598103934665603ull;
        auto mixin = [&](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); for (size_t i = 0; i < n; ++i) { d ^= b[i]; d *= 1099511628211ull; } };
        mixin(h.ops.data(), h.ops.size() * sizeof(Operand)); mixin(h.plots.data(), h.plots.size() * sizeof(Plot));
        mixin(h.tasks.data(), h.tasks.size() * sizeof(Task)); mixin(h.chunk_prefix.data(), h.chunk_prefix.size() * sizeof(int32_t));
        mixin(h.class_ids.data(), h.class_ids.size() * sizeof(int32_t)); mixin(h.class_off.data(), h.class_off.size() * sizeof(int64_t));
        mixin(h.chunks.data(), h.chunks.size() * sizeof(TabChunk)); mixin(h.jp.items.data(), h.jp.items.size() * sizeof(JoinItem));
        mixin(h.jp.jplots.data(), h.jp.jplots.size() * sizeof(int32_t)); mixin(h.jp.item_off.data(), h.jp.item_off.size() * sizeof(int64_t));
        mixin(h.strip_prefix.data(), h.strip_prefix.size() * sizeof(int64_t));
        for (const Wave& w : h.waves) mixin(&w, sizeof(Wave));
        *digest = d;
    }
    if (counts) {
        counts[0] = (int64_t)h.ops.size(); counts[1] = (int64_t)h.plots.size(); counts[2] = (int64_t)h.tasks.size();
        counts[3] = (int64_t)h.waves.size(); counts[4] = (int64_t)h.chunks.size(); counts[5] = (int64_t)h.jp.items.size();
        counts[6] = h.tm.cells; counts[7] = h.max_wave_hits;
    }
    return VAPOR_OK;
}

int vapor_gpu_int_peak(void* handle, int which, double* lane_ops_per_s) {
    if (!handle || !lane_ops_per_s || which < 0 || which > 6) return VAPOR_E_ARG;
    Handle* h = static_cast<Handle*>(handle);
    CK(cudaSetDevice(h->device));
    DevBuf<uint32_t> in, out;
    CK(in.ensure(1024)); CK(out.ensure(4));
    std::vector<uint32_t> hin(1024);
    for (int i = 0; i < 1024; ++i) hin[i] = 0x01000193u * (uint32_t)(i + 1) ^ 0xA5A5A5A5u;
    CK(cudaMemcpy(in.p, hin.data(), 1024 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    const int grid = h->sm_count * 8, iters = 1 << 15;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(a, h->stream));
        double ops = (double)grid * 256.0 * (double)iters * 32.0;
        if (which == 0) k_int_peak<0><<<grid, 256, 0, h->stream>>>(in.p, out.p, iters);
        else if (which == 1) k_int_peak<1><<<grid, 256, 0, h->stream>>>(in.p, out.p, iters);
        else if (which == 3) k_int_peak<3><<<grid, 256, 0, h->stream>>>(in.p, out.p, iters);
        else if (which == 5) k_int_peak<5><<<grid, 256, 0, h->stream>>>(in.p, out.p, iters);
        else if (which == 6) k_int_peak<6><<<grid, 256, 0, h->stream>>>(in.p, out.p, iters);
        else if (which == 4) {
            const int g4 = h->sm_count * 4, reps = 48;
            k_tile_loop_peak<<<g4, 128, 0, h->stream>>>(in.p, out.p, reps);
            ops = (double)g4 * 128.0 * (double)reps * 2048.0 * 32.0;          // 32 integer instructions per lane and word
        }
        else k_int_peak<2><<<grid, 256, 0, h->stream>>>(in.p, out.p, iters);
        CK(cudaEventRecord(b, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    in.release(); out.release();
    *lane_ops_per_s = best;
    return VAPOR_OK;
}

}  // extern "C"
