// Native region extraction for the scoring path: see include/vapor_hostio.h.
//
// Host code only (C++17 + zlib).  It restates, for many windows per call and on several threads,
//   ref_seq_readin            vapor_vali/Simple_function.pyx:1203-1217   (samtools faidx)
//   chop_pacbio_read_by_pos   :339-354                                   (samtools view + cut to the window)
//   cigar2alignstart_by_pos   :309-337                                   (CIGAR walk with the reference's quirks)
//   minimize_pacbio_read_list :1091-1102                                 (at most 20 reads, smallest miss_bp first)
// and is pinned against vapor_b200/Simple_function.py's versions of the same functions (themselves pinned to the
// reference by tests/test_cli_host.py) and against spec-derived BAM fixtures (tests/test_hostio.py).
#include "../../include/vapor_b200.h"
#include "../../include/vapor_hostio.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>
#include <zlib.h>

namespace {

thread_local std::string t_err;
int fail(int code, const std::string& msg) { t_err = msg; return code; }

template <typename F>
void run_threads(int n, F&& f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)n - 1);
    for (int i = 1; i < n; ++i) th.emplace_back([&f, i]() { f(i); });
    f(0);
    for (auto& t : th) t.join();
}

// ================================================================================================
// FASTA
// ================================================================================================
struct FaiEntry { int64_t length, offset, lb, lw; };

struct Fasta {
    std::string path;
    int fd = -1;
    std::unordered_map<std::string, FaiEntry> index;
    std::vector<uint8_t> bytes;      // result of the last fetch_many
    std::vector<int64_t> off;
};

bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}

// <path>.fai with the five columns of `samtools faidx`
bool build_fai(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    struct Row { std::string name; int64_t length, offset, lb, lw; };
    std::vector<Row> rows;
    std::string line;
    bool have = false;
    Row cur{};
    int64_t pos = 0;
    std::vector<char> buf(1 << 20);
    std::string carry;
    auto handle_line = [&](const char* p, size_t n) {          // n includes the newline when there is one
        if (n > 0 && p[0] == '>') {
            if (have) rows.push_back(cur);
            have = true;
            size_t a = 1, b = 1;
            while (b < n && !isspace((unsigned char)p[b])) ++b;
            cur = Row{std::string(p + a, b - a), 0, pos + (int64_t)n, 0, 0};
        } else if (have) {
            size_t bases = n;
            while (bases > 0 && (p[bases - 1] == '\n' || p[bases - 1] == '\r')) --bases;
            if (cur.lb == 0 && bases) { cur.lb = (int64_t)bases; cur.lw = (int64_t)n; }
            cur.length += (int64_t)bases;
        }
        pos += (int64_t)n;
    };
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        size_t s = 0;
        for (size_t i = 0; i < got; ++i) {
            if (buf[i] == '\n') {
                if (!carry.empty()) { carry.append(buf.data() + s, i + 1 - s); handle_line(carry.data(), carry.size()); carry.clear(); }
                else handle_line(buf.data() + s, i + 1 - s);
                s = i + 1;
            }
        }
        if (s < got) carry.append(buf.data() + s, got - s);
    }
    if (!carry.empty()) handle_line(carry.data(), carry.size());
    if (have) rows.push_back(cur);
    fclose(f);
    FILE* o = fopen((path + ".fai").c_str(), "w");
    if (!o) return false;
    for (auto& r : rows) fprintf(o, "%s\t%lld\t%lld\t%lld\t%lld\n", r.name.c_str(), (long long)r.length, (long long)r.offset, (long long)r.lb, (long long)r.lw);
    fclose(o);
    return true;
}

// bases start..end (1-based inclusive), clipped like samtools; appended to `out`
void fasta_fetch(const Fasta& fa, const std::string& chrom, int64_t start, int64_t end, std::string& out, std::vector<char>& tmp) {
    auto it = fa.index.find(chrom);
    if (it == fa.index.end()) return;
    const FaiEntry& e = it->second;
    start = std::max<int64_t>(start, 1);
    end = std::min<int64_t>(end, e.length);
    if (end < start || e.lb <= 0) return;
    const int64_t s0 = start - 1, e0 = end;
    const int64_t b0 = e.offset + (s0 / e.lb) * e.lw + s0 % e.lb;
    const int64_t b1 = e.offset + ((e0 - 1) / e.lb) * e.lw + (e0 - 1) % e.lb + 1;
    tmp.resize((size_t)(b1 - b0));
    int64_t done = 0;
    while (done < b1 - b0) {
        ssize_t r = pread(fa.fd, tmp.data() + done, (size_t)(b1 - b0 - done), (off_t)(b0 + done));
        if (r <= 0) break;
        done += r;
    }
    out.reserve(out.size() + (size_t)done);
    for (int64_t i = 0; i < done; ++i) if (tmp[(size_t)i] != '\n' && tmp[(size_t)i] != '\r') out.push_back(tmp[(size_t)i]);
}

// ================================================================================================
// alignment records
// ================================================================================================
// CIGAR operations as BAM stores them: len << 4 | op, op = index into "MIDNSHP=X"
constexpr char CIG_OPS[] = "MIDNSHP=XB";

// the tokens the reference's regex (\d+)([MIDNSHP=X]) finds in a SAM CIGAR string
void parse_cigar_text(const char* s, size_t n, std::vector<uint32_t>& ops) {
    uint64_t num = 0; bool have = false;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        if (c >= '0' && c <= '9') { num = have ? std::min<uint64_t>(num * 10 + (uint64_t)(c - '0'), (1ull << 28) - 1) : (uint64_t)(c - '0'); have = true; continue; }
        if (have) {
            const char* p = (const char*)memchr("MIDNSHP=X", c, 9);
            if (p) ops.push_back((uint32_t)(num << 4) | (uint32_t)(p - "MIDNSHP=X"));
        }
        have = false; num = 0;
    }
}

inline int64_t ref_span(const uint32_t* ops, size_t n) {     // M, D, N, =, X; at least 1
    int64_t s = 0;
    for (size_t i = 0; i < n; ++i) { const uint32_t op = ops[i] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) s += ops[i] >> 4; }
    return s > 0 ? s : 1;
}

// cigar2alignstart_by_pos (Simple_function.pyx:309-337): S, M, =, I advance the read; M, =, D the reference; X, N, H, P
// nothing; the loop stops after the first operation that puts the reference position past start - 1.
inline void cigar_walk(const uint32_t* ops, size_t n, int64_t align_start, int64_t start, int64_t& read_off, int64_t& miss) {
    int64_t read_rec = 0, align_rec = align_start;
    int last = -1;
    for (size_t i = 0; i < n; ++i) {
        const int64_t len = ops[i] >> 4; const int op = (int)(ops[i] & 15);
        if (op > 8) continue;                            // 'B' never matches the reference's pattern
        if (op == 4) read_rec += len;                    // S
        else if (op == 0 || op == 7) { read_rec += len; align_rec += len; }   // M, =
        else if (op == 2) align_rec += len;              // D
        else if (op == 1) read_rec += len;               // I
        last = op;
        if (align_rec > start - 1) break;
    }
    const int64_t start_dis = align_rec - start;
    if (last == 0 || last == 7) { read_off = read_rec - start_dis; miss = 0; }
    else { read_off = read_rec; miss = start_dis; }
}

struct RecView {             // what chop needs of one record
    int64_t pos;             // 1-based POS
    const uint32_t* ops; size_t n_ops;
    const char* seq; size_t seq_len;
    const char* qname; size_t qname_len;
};

struct Kept { std::string seq, qname; int64_t miss; };

// chop_pacbio_read_by_pos for one record (Simple_function.pyx:345-352), Python slicing rules included
inline void chop_record(const RecView& r, int64_t start, int64_t end, int64_t flank, std::vector<Kept>& out) {
    if (!(r.pos < start + 1)) return;
    int64_t a, miss;
    cigar_walk(r.ops, r.n_ops, r.pos, start, a, miss);
    if ((double)miss > (double)flank / 2.0) return;
    const int64_t len = (int64_t)r.seq_len;
    int64_t from = a < 0 ? std::max<int64_t>(0, len + a) : std::min<int64_t>(a, len);     // seq[a:]
    const int64_t tlen = len - from;
    const int64_t L = end - start - miss;
    if (!(tlen > L)) return;
    const int64_t take = L < 0 ? std::max<int64_t>(0, tlen + L) : std::min<int64_t>(L, tlen);   // target[:L]
    out.push_back(Kept{std::string(r.seq + from, (size_t)take), std::string(r.qname, r.qname_len), miss});
}

// ---- SAM text: whole file in memory, per contig in file order ----------------------------------------------
struct SamRec { int64_t pos, end; uint64_t ops_off; uint32_t n_ops; uint64_t seq_off; uint32_t seq_len; uint64_t qname_off; uint32_t qname_len; };
struct SamContig { std::vector<SamRec> recs; std::vector<int64_t> starts; bool sorted = true; int64_t max_span = 1; };

struct Aln {
    std::string path;
    bool is_bam = false;
    // SAM text
    std::unordered_map<std::string, SamContig> contigs;
    std::vector<uint32_t> ops; std::string seqs, qnames;
    // BAM
    int fd = -1;
    std::vector<std::string> refs;
    std::unordered_map<std::string, int> tid;
    int64_t first_rec = 0;                       // virtual offset of the first record
    bool have_bai = false;
    struct BaiRef { std::unordered_map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins; std::vector<uint64_t> ioff; };
    std::vector<BaiRef> bai;
};

bool load_sam(Aln& a) {
    gzFile f = gzopen(a.path.c_str(), "rb");
    if (!f) return false;
    gzbuffer(f, 1 << 20);
    std::string line;
    std::vector<char> buf(1 << 20);
    std::vector<const char*> fld; std::vector<size_t> flen;
    auto handle = [&](const char* p, size_t n) {
        while (n > 0 && (p[n - 1] == '\n')) --n;
        if (n == 0 || p[0] == '@') return;
        fld.clear(); flen.clear();
        size_t s = 0;
        for (size_t i = 0; i <= n; ++i) if (i == n || p[i] == '\t') { fld.push_back(p + s); flen.push_back(i - s); s = i + 1; }
        if (fld.size() < 10) {                               // not tab separated: any whitespace
            fld.clear(); flen.clear();
            size_t i = 0;
            while (i < n) {
                while (i < n && isspace((unsigned char)p[i])) ++i;
                size_t b = i;
                while (i < n && !isspace((unsigned char)p[i])) ++i;
                if (i > b) { fld.push_back(p + b); flen.push_back(i - b); }
            }
            if (fld.size() < 10) return;
        }
        if (flen[2] == 1 && fld[2][0] == '*') return;
        SamRec r{};
        r.pos = strtoll(std::string(fld[3], flen[3]).c_str(), nullptr, 10);
        r.ops_off = a.ops.size();
        parse_cigar_text(fld[5], flen[5], a.ops);
        r.n_ops = (uint32_t)(a.ops.size() - r.ops_off);
        r.end = r.pos + ref_span(a.ops.data() + r.ops_off, r.n_ops) - 1;
        r.seq_off = a.seqs.size(); r.seq_len = (uint32_t)flen[9]; a.seqs.append(fld[9], flen[9]);
        r.qname_off = a.qnames.size(); r.qname_len = (uint32_t)flen[0]; a.qnames.append(fld[0], flen[0]);
        a.contigs[std::string(fld[2], flen[2])].recs.push_back(r);
    };
    int got;
    while ((got = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
        size_t s = 0;
        for (size_t i = 0; i < (size_t)got; ++i) {
            if (buf[i] == '\n') {
                if (!line.empty()) { line.append(buf.data() + s, i + 1 - s); handle(line.data(), line.size()); line.clear(); }
                else handle(buf.data() + s, i + 1 - s);
                s = i + 1;
            }
        }
        if (s < (size_t)got) line.append(buf.data() + s, (size_t)got - s);
    }
    if (!line.empty()) handle(line.data(), line.size());
    gzclose(f);
    for (auto& kv : a.contigs) {
        SamContig& c = kv.second;
        c.starts.reserve(c.recs.size());
        for (auto& r : c.recs) { c.starts.push_back(r.pos); c.max_span = std::max(c.max_span, r.end - r.pos + 1); }
        c.sorted = std::is_sorted(c.starts.begin(), c.starts.end());
    }
    return true;
}

// ---- BGZF / BAM ---------------------------------------------------------------------------------------------
struct Bgzf {                    // one cursor (per thread): a decompressed block and a position in it
    int fd;
    int64_t block_start = -1, block_len = 0;
    std::vector<uint8_t> data, cbuf;
    size_t upos = 0;
    explicit Bgzf(int fd_) : fd(fd_) {}

    bool load(int64_t coffset) {
        uint8_t hdr[18];
        data.clear(); block_start = coffset; block_len = 0; upos = 0;
        if (pread(fd, hdr, 18, (off_t)coffset) < 18) return false;
        if (!(hdr[0] == 0x1f && hdr[1] == 0x8b && hdr[2] == 8 && (hdr[3] & 4))) return false;
        const int xlen = hdr[10] | (hdr[11] << 8);
        std::vector<uint8_t> extra((size_t)xlen);
        if (pread(fd, extra.data(), (size_t)xlen, (off_t)(coffset + 12)) < xlen) return false;
        int bsize = -1;
        for (int i = 0; i + 4 <= xlen;) {
            const int slen = extra[(size_t)i + 2] | (extra[(size_t)i + 3] << 8);
            if (extra[(size_t)i] == 66 && extra[(size_t)i + 1] == 67 && i + 6 <= xlen) bsize = extra[(size_t)i + 4] | (extra[(size_t)i + 5] << 8);
            i += 4 + slen;
        }
        if (bsize < 0) return false;
        const int clen = bsize - xlen - 19;
        if (clen < 0) return false;
        cbuf.resize((size_t)clen + 8);
        if (pread(fd, cbuf.data(), (size_t)clen + 8, (off_t)(coffset + 12 + xlen)) < clen + 8) return false;
        const uint32_t isize = cbuf[(size_t)clen + 4] | (cbuf[(size_t)clen + 5] << 8) | (cbuf[(size_t)clen + 6] << 16) | ((uint32_t)cbuf[(size_t)clen + 7] << 24);
        data.resize(isize);
        if (isize) {
            z_stream zs{};
            if (inflateInit2(&zs, -15) != Z_OK) return false;
            zs.next_in = cbuf.data(); zs.avail_in = (uInt)clen;
            zs.next_out = data.data(); zs.avail_out = (uInt)isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) { data.clear(); return false; }
        }
        block_len = bsize + 1;
        return true;
    }
    void seek(int64_t voffset) {
        const int64_t co = voffset >> 16;
        if (co != block_start || block_len == 0) load(co);
        upos = (size_t)(voffset & 0xFFFF);
    }
    int64_t tell() const { return (block_start << 16) | (int64_t)upos; }
    size_t read(uint8_t* out, size_t n) {
        size_t done = 0;
        while (done < n) {
            if (upos >= data.size()) {
                if (block_len == 0 && block_start >= 0 && data.empty()) { if (!load(block_start)) break; if (data.empty() && block_len == 0) break; continue; }
                if (!load(block_start + block_len)) break;
                continue;                                   // an empty block (the EOF marker) falls through to the next load
            }
            const size_t take = std::min(n - done, data.size() - upos);
            memcpy(out + done, data.data() + upos, take);
            upos += take; done += take;
        }
        return done;
    }
};

inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd_u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

bool load_bam_header(Aln& a) {
    a.fd = open(a.path.c_str(), O_RDONLY);
    if (a.fd < 0) return false;
    Bgzf bg(a.fd);
    bg.seek(0);
    uint8_t b4[4];
    if (bg.read(b4, 4) < 4 || memcmp(b4, "BAM\1", 4) != 0) return false;
    if (bg.read(b4, 4) < 4) return false;
    int32_t l_text = rd_i32(b4);
    std::vector<uint8_t> skip((size_t)std::max(l_text, 0));
    if (l_text > 0 && bg.read(skip.data(), (size_t)l_text) < (size_t)l_text) return false;
    if (bg.read(b4, 4) < 4) return false;
    const int32_t n_ref = rd_i32(b4);
    for (int32_t i = 0; i < n_ref; ++i) {
        if (bg.read(b4, 4) < 4) return false;
        const int32_t l_name = rd_i32(b4);
        std::vector<uint8_t> nm((size_t)l_name);
        if (bg.read(nm.data(), (size_t)l_name) < (size_t)l_name) return false;
        a.refs.emplace_back((const char*)nm.data(), (size_t)std::max(0, l_name - 1));
        a.tid[a.refs.back()] = i;
        if (bg.read(b4, 4) < 4) return false;
    }
    a.first_rec = bg.tell();
    // index: <path>.bai or <stem>.bai
    std::string raw;
    std::string stem = a.path.substr(0, a.path.rfind('.') == std::string::npos ? a.path.size() : a.path.rfind('.'));
    if (read_file(a.path + ".bai", raw) || read_file(stem + ".bai", raw)) {
        if (raw.size() >= 8 && memcmp(raw.data(), "BAI\1", 4) == 0) {
            const uint8_t* p = (const uint8_t*)raw.data(); size_t o = 4;
            const int32_t nr = rd_i32(p + o); o += 4;
            a.bai.resize((size_t)std::max(nr, 0));
            bool ok = true;
            for (int32_t r = 0; r < nr && ok; ++r) {
                if (o + 4 > raw.size()) { ok = false; break; }
                const int32_t n_bin = rd_i32(p + o); o += 4;
                for (int32_t b = 0; b < n_bin; ++b) {
                    if (o + 8 > raw.size()) { ok = false; break; }
                    const uint32_t bin = rd_u32(p + o); const int32_t n_chunk = rd_i32(p + o + 4); o += 8;
                    if (o + 16 * (size_t)n_chunk > raw.size()) { ok = false; break; }
                    auto& v = a.bai[(size_t)r].bins[bin];
                    for (int32_t c = 0; c < n_chunk; ++c) v.push_back({rd_u64(p + o + 16 * (size_t)c), rd_u64(p + o + 16 * (size_t)c + 8)});
                    o += 16 * (size_t)n_chunk;
                }
                if (!ok || o + 4 > raw.size()) { ok = false; break; }
                const int32_t n_intv = rd_i32(p + o); o += 4;
                if (o + 8 * (size_t)n_intv > raw.size()) { ok = false; break; }
                for (int32_t i = 0; i < n_intv; ++i) a.bai[(size_t)r].ioff.push_back(rd_u64(p + o + 8 * (size_t)i));
                o += 8 * (size_t)n_intv;
            }
            a.have_bai = ok;
        }
    }
    return true;
}

// records of `samtools view file chrom:start-end` from a BAM, in file order
template <typename F>
void bam_fetch(const Aln& a, Bgzf& bg, const std::string& chrom, int64_t start, int64_t end, F&& emit) {
    auto it = a.tid.find(chrom);
    if (it == a.tid.end()) return;
    const int tid = it->second;
    const int64_t beg0 = std::max<int64_t>(start - 1, 0), end0 = std::max<int64_t>(end, 1);
    int64_t voff = a.first_rec;
    if (a.have_bai && (size_t)tid < a.bai.size()) {
        const Aln::BaiRef& br = a.bai[(size_t)tid];
        const uint64_t lin = br.ioff.empty() ? 0 : br.ioff[(size_t)std::min<int64_t>(beg0 >> 14, (int64_t)br.ioff.size() - 1)];
        uint64_t best = ~0ull; bool any = false;
        auto visit = [&](uint32_t bin) {
            auto f = br.bins.find(bin);
            if (f == br.bins.end()) return;
            for (auto& c : f->second) if (c.second > lin) { best = std::min(best, c.first); any = true; }
        };
        const int64_t e1 = end0 - 1;
        visit(0);
        const int shifts[5] = {26, 23, 20, 17, 14}; const uint32_t bases[5] = {1, 9, 73, 585, 4681};
        for (int l = 0; l < 5; ++l) for (int64_t b = bases[l] + (beg0 >> shifts[l]); b <= bases[l] + (e1 >> shifts[l]); ++b) visit((uint32_t)b);
        if (!any) return;
        voff = (int64_t)(lin ? std::max(best, lin) : best);
    }
    bg.seek(voff);
    std::vector<uint8_t> rec;
    std::vector<uint32_t> ops;
    std::string seq;
    static const char dec[] = "=ACMGRSVTWYHKDBN";
    while (true) {
        uint8_t b4[4];
        if (bg.read(b4, 4) < 4) return;
        const int32_t bs = rd_i32(b4);
        if (bs < 32) return;
        rec.resize((size_t)bs);
        if (bg.read(rec.data(), (size_t)bs) < (size_t)bs) return;
        const int32_t rtid = rd_i32(rec.data()), pos0 = rd_i32(rec.data() + 4);
        const int l_rn = rec[8];
        const int n_cig = rec[12] | (rec[13] << 8);
        const int32_t l_seq = rd_i32(rec.data() + 16);
        if (rtid != tid) {
            if (rtid > tid || rtid < 0) return;
            continue;
        }
        const int64_t pos = (int64_t)pos0 + 1;
        if (pos > end) return;
        size_t p = 32;
        const char* qname = (const char*)rec.data() + p; const size_t qlen = l_rn > 0 ? (size_t)l_rn - 1 : 0;
        p += (size_t)l_rn;
        ops.resize((size_t)n_cig);
        for (int i = 0; i < n_cig; ++i) ops[(size_t)i] = rd_u32(rec.data() + p + 4 * (size_t)i);
        p += 4 * (size_t)n_cig;
        const size_t nb = ((size_t)std::max(l_seq, 0) + 1) / 2;
        const uint8_t* sq = rec.data() + p;
        p += nb + (size_t)std::max(l_seq, 0);             // SEQ + QUAL
        // more than 65535 operations: the CIGAR field holds <l_seq>S<ref_len>N and the real one sits in CG:B,I
        if (n_cig == 2 && (ops[0] & 15) == 4 && (int32_t)(ops[0] >> 4) == l_seq && (ops[1] & 15) == 3) {
            size_t q = p;
            while (q + 3 <= rec.size()) {
                const char t0 = (char)rec[q], t1 = (char)rec[q + 1], ty = (char)rec[q + 2];
                q += 3;
                size_t sz = 0;
                if (ty == 'A' || ty == 'c' || ty == 'C') sz = 1;
                else if (ty == 's' || ty == 'S') sz = 2;
                else if (ty == 'i' || ty == 'I' || ty == 'f') sz = 4;
                else if (ty == 'Z' || ty == 'H') { while (q + sz < rec.size() && rec[q + sz]) ++sz; ++sz; }
                else if (ty == 'B') {
                    if (q + 5 > rec.size()) break;
                    const char sub = (char)rec[q]; const uint32_t cnt = rd_u32(rec.data() + q + 1);
                    const size_t es = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
                    if (t0 == 'C' && t1 == 'G' && sub == 'I' && q + 5 + 4 * (size_t)cnt <= rec.size()) {
                        ops.resize(cnt);
                        for (uint32_t i = 0; i < cnt; ++i) ops[i] = rd_u32(rec.data() + q + 5 + 4 * (size_t)i);
                        break;
                    }
                    sz = 5 + es * (size_t)cnt;
                } else break;
                q += sz;
            }
        }
        const int64_t rend = pos0 + ref_span(ops.data(), ops.size());
        if (rend < start) continue;
        seq.clear();
        if (l_seq <= 0) seq = "*";
        else {
            seq.resize((size_t)l_seq);
            for (int32_t i = 0; i < l_seq; ++i) { const uint8_t b = sq[(size_t)i >> 1]; seq[(size_t)i] = dec[(i & 1) ? (b & 15) : (b >> 4)]; }
        }
        emit(RecView{pos, ops.data(), ops.size(), seq.data(), seq.size(), qname, qlen});
    }
}

template <typename F>
void sam_fetch(const Aln& a, const std::string& chrom, int64_t start, int64_t end, F&& emit) {
    auto it = a.contigs.find(chrom);
    if (it == a.contigs.end()) return;
    const SamContig& c = it->second;
    size_t lo = 0, hi = c.recs.size();
    if (c.sorted) {
        lo = (size_t)(std::lower_bound(c.starts.begin(), c.starts.end(), start - c.max_span + 1) - c.starts.begin());
        hi = (size_t)(std::upper_bound(c.starts.begin(), c.starts.end(), end) - c.starts.begin());
    }
    for (size_t i = lo; i < hi; ++i) {
        const SamRec& r = c.recs[i];
        if (r.pos <= end && r.end >= start)
            emit(RecView{r.pos, a.ops.data() + r.ops_off, r.n_ops, a.seqs.data() + r.seq_off, r.seq_len, a.qnames.data() + r.qname_off, r.qname_len});
    }
}

struct ReadsOwner {
    vapor_io_reads_t pub;        // first member: the public view
    std::vector<int64_t> win_off, seq_off, qname_off;
    std::vector<uint8_t> seq_bytes, qname_bytes;
    std::vector<int32_t> miss;
};

}  // namespace

extern "C" {

const char* vapor_io_last_error(void) { return t_err.c_str(); }

int vapor_io_fasta_open(const char* path, void** fasta) {
    if (!path || !fasta) return fail(VAPOR_E_ARG, "NULL argument");
    *fasta = nullptr;
    Fasta* fa = new Fasta();
    fa->path = path;
    struct stat st;
    if (stat((fa->path + ".fai").c_str(), &st) != 0 && !build_fai(fa->path)) { delete fa; return fail(VAPOR_E_ARG, std::string("cannot index ") + path); }
    std::string raw;
    if (!read_file(fa->path + ".fai", raw)) { delete fa; return fail(VAPOR_E_ARG, std::string("cannot read ") + path + ".fai"); }
    size_t s = 0;
    while (s < raw.size()) {
        size_t e = raw.find('\n', s);
        if (e == std::string::npos) e = raw.size();
        std::vector<std::string> f;
        size_t a = s;
        for (size_t i = s; i <= e; ++i) if (i == e || raw[i] == '\t') { f.emplace_back(raw, a, i - a); a = i + 1; }
        if (f.size() < 5) {                                  // whitespace separated
            f.clear();
            size_t i = s;
            while (i < e) { while (i < e && isspace((unsigned char)raw[i])) ++i; size_t b = i; while (i < e && !isspace((unsigned char)raw[i])) ++i; if (i > b) f.emplace_back(raw, b, i - b); }
        }
        if (f.size() >= 5) fa->index[f[0]] = FaiEntry{atoll(f[1].c_str()), atoll(f[2].c_str()), atoll(f[3].c_str()), atoll(f[4].c_str())};
        s = e + 1;
    }
    fa->fd = open(path, O_RDONLY);
    if (fa->fd < 0) { delete fa; return fail(VAPOR_E_ARG, std::string("cannot open ") + path); }
    *fasta = fa;
    return VAPOR_OK;
}

int vapor_io_fasta_close(void* fasta) {
    if (!fasta) return VAPOR_E_ARG;
    Fasta* fa = static_cast<Fasta*>(fasta);
    if (fa->fd >= 0) close(fa->fd);
    delete fa;
    return VAPOR_OK;
}

int vapor_io_fasta_fetch(void* fasta, const char* chrom, int64_t start, int64_t end, char* out, int64_t cap, int64_t* len) {
    if (!fasta || !chrom || !len || (cap > 0 && !out)) return fail(VAPOR_E_ARG, "NULL argument");
    std::string s; std::vector<char> tmp;
    fasta_fetch(*static_cast<Fasta*>(fasta), chrom, start, end, s, tmp);
    *len = (int64_t)s.size();
    if (cap > 0) memcpy(out, s.data(), (size_t)std::min<int64_t>(cap, (int64_t)s.size()));
    return VAPOR_OK;
}

int vapor_io_fasta_fetch_many(void* fasta, int64_t n, const char* chroms, const int64_t* chrom_off, const int64_t* start,
                              const int64_t* end, int threads, const uint8_t** bytes, const int64_t** off)
{
    if (!fasta || n < 0 || (n > 0 && (!chroms || !chrom_off || !start || !end)) || !bytes || !off) return fail(VAPOR_E_ARG, "NULL argument");
    Fasta* fa = static_cast<Fasta*>(fasta);
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 4, n / 64 + 1));
    std::vector<std::vector<std::string>> parts((size_t)threads);
    run_threads(threads, [&](int t) {
        const int64_t a = n * t / threads, b = n * (t + 1) / threads;
        std::vector<char> tmp;
        auto& v = parts[(size_t)t];
        v.resize((size_t)(b - a));
        for (int64_t i = a; i < b; ++i) fasta_fetch(*fa, chroms + chrom_off[i], start[i], end[i], v[(size_t)(i - a)], tmp);
    });
    fa->off.assign((size_t)n + 1, 0);
    int64_t i = 0, total = 0;
    for (auto& v : parts) for (auto& s : v) { total += (int64_t)s.size(); fa->off[(size_t)++i] = total; }
    fa->bytes.resize((size_t)total);
    i = 0;
    for (auto& v : parts) for (auto& s : v) { memcpy(fa->bytes.data() + fa->off[(size_t)i], s.data(), s.size()); ++i; }
    *bytes = fa->bytes.data(); *off = fa->off.data();
    return VAPOR_OK;
}

int vapor_io_aln_open(const char* path, void** aln) {
    if (!path || !aln) return fail(VAPOR_E_ARG, "NULL argument");
    *aln = nullptr;
    Aln* a = new Aln();
    a->path = path;
    std::string low = a->path;
    std::transform(low.begin(), low.end(), low.begin(), [](unsigned char c) { return (char)tolower(c); });
    auto ends = [&](const char* suf) { const size_t n = strlen(suf); return low.size() >= n && low.compare(low.size() - n, n, suf) == 0; };
    bool text = ends(".sam") || ends(".sam.gz");
    if (!text) {
        unsigned char m[2] = {0, 0};
        FILE* f = fopen(path, "rb");
        if (!f) { delete a; return fail(VAPOR_E_ARG, std::string("cannot open ") + path); }
        const size_t g = fread(m, 1, 2, f);
        fclose(f);
        text = !(g == 2 && m[0] == 0x1f && m[1] == 0x8b);
    }
    if (text) {
        if (!load_sam(*a)) { delete a; return fail(VAPOR_E_ARG, std::string("cannot read ") + path); }
    } else {
        a->is_bam = true;
        if (!load_bam_header(*a)) { if (a->fd >= 0) close(a->fd); delete a; return fail(VAPOR_E_ARG, std::string(path) + ": not a BAM file"); }
    }
    *aln = a;
    return VAPOR_OK;
}

int vapor_io_aln_close(void* aln) {
    if (!aln) return VAPOR_E_ARG;
    Aln* a = static_cast<Aln*>(aln);
    if (a->fd >= 0) close(a->fd);
    delete a;
    return VAPOR_OK;
}

int vapor_io_chop_many(void* const* alns, int n_aln, int64_t n_win, const char* chroms, const int64_t* chrom_off,
                       const int64_t* start, const int64_t* end, const int64_t* flank, int max_reads, int threads,
                       vapor_io_reads_t** result)
{
    if (!result || n_aln < 0 || n_win < 0 || (n_aln > 0 && !alns) || (n_win > 0 && (!chroms || !chrom_off || !start || !end || !flank)))
        return fail(VAPOR_E_ARG, "NULL argument");
    *result = nullptr;
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 4, n_win / 16 + 1));
    struct Part { std::vector<Kept> kept; std::vector<int64_t> count; int64_t seen = 0; };
    std::vector<Part> parts((size_t)threads);
    run_threads(threads, [&](int t) {
        const int64_t a = n_win * t / threads, b = n_win * (t + 1) / threads;
        Part& P = parts[(size_t)t];
        P.count.reserve((size_t)(b - a));
        std::vector<std::unique_ptr<Bgzf>> cur((size_t)n_aln);
        std::vector<Kept> win;
        for (int64_t w = a; w < b; ++w) {
            win.clear();
            const std::string chrom(chroms + chrom_off[w]);
            for (int f = 0; f < n_aln; ++f) {
                const Aln& A = *static_cast<const Aln*>(alns[f]);
                auto emit = [&](const RecView& r) { ++P.seen; chop_record(r, start[w], end[w], flank[w], win); };
                if (A.is_bam) {
                    if (!cur[(size_t)f]) cur[(size_t)f].reset(new Bgzf(A.fd));
                    bam_fetch(A, *cur[(size_t)f], chrom, start[w], end[w], emit);
                } else sam_fetch(A, chrom, start[w], end[w], emit);
            }
            if (max_reads > 0 && (int64_t)win.size() > max_reads) {      // minimize_pacbio_read_list: smallest miss_bp first, file order inside
                std::stable_sort(win.begin(), win.end(), [](const Kept& x, const Kept& y) { return x.miss < y.miss; });
                win.resize((size_t)max_reads);
            }
            P.count.push_back((int64_t)win.size());
            for (auto& k : win) P.kept.push_back(std::move(k));
        }
    });
    ReadsOwner* R = new ReadsOwner();
    R->win_off.assign((size_t)n_win + 1, 0);
    int64_t w = 0, n_reads = 0, seq_total = 0, qn_total = 0, seen = 0;
    for (auto& P : parts) {
        for (int64_t c : P.count) { n_reads += c; R->win_off[(size_t)++w] = n_reads; }
        for (auto& k : P.kept) { seq_total += (int64_t)k.seq.size(); qn_total += (int64_t)k.qname.size(); }
        seen += P.seen;
    }
    R->seq_off.assign((size_t)n_reads + 1, 0); R->qname_off.assign((size_t)n_reads + 1, 0);
    R->seq_bytes.resize((size_t)seq_total); R->qname_bytes.resize((size_t)qn_total); R->miss.resize((size_t)n_reads);
    int64_t i = 0, so = 0, qo = 0;
    for (auto& P : parts) for (auto& k : P.kept) {
        memcpy(R->seq_bytes.data() + so, k.seq.data(), k.seq.size()); so += (int64_t)k.seq.size();
        memcpy(R->qname_bytes.data() + qo, k.qname.data(), k.qname.size()); qo += (int64_t)k.qname.size();
        R->miss[(size_t)i] = (int32_t)k.miss;
        ++i;
        R->seq_off[(size_t)i] = so; R->qname_off[(size_t)i] = qo;
    }
    R->pub.n_win = n_win; R->pub.win_off = R->win_off.data(); R->pub.seq_off = R->seq_off.data(); R->pub.seq_bytes = R->seq_bytes.data();
    R->pub.miss = R->miss.data(); R->pub.qname_off = R->qname_off.data(); R->pub.qname_bytes = R->qname_bytes.data();
    R->pub.n_records_seen = seen;
    *result = &R->pub;
    return VAPOR_OK;
}

int vapor_io_reads_free(vapor_io_reads_t* result) {
    if (!result) return VAPOR_OK;
    delete reinterpret_cast<ReadsOwner*>(result);
    return VAPOR_OK;
}

int vapor_host_scatter_runs(void* dst, const void* src, int64_t elem_bytes, const int64_t* run_dst, const int64_t* run_len, int64_t n_runs, int threads) {
    if (n_runs < 0 || elem_bytes <= 0 || (n_runs > 0 && (!dst || !src || !run_dst || !run_len))) return fail(VAPOR_E_ARG, "bad scatter arguments");
    if (n_runs == 0) return VAPOR_OK;
    std::vector<int64_t> src_off((size_t)n_runs + 1, 0);
    for (int64_t r = 0; r < n_runs; ++r) { if (run_len[r] < 0 || run_dst[r] < 0) return fail(VAPOR_E_ARG, "negative run"); src_off[(size_t)r + 1] = src_off[(size_t)r] + run_len[r]; }
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 2, n_runs / 4096 + 1));
    uint8_t* d = static_cast<uint8_t*>(dst); const uint8_t* s_ = static_cast<const uint8_t*>(src);
    run_threads(threads, [&](int t) {
        for (int64_t r = n_runs * t / threads; r < n_runs * (t + 1) / threads; ++r)
            memcpy(d + run_dst[r] * elem_bytes, s_ + src_off[(size_t)r] * elem_bytes, (size_t)(run_len[r] * elem_bytes));
    });
    return VAPOR_OK;
}

int vapor_io_cigar2alignstart(const char* cigar, int64_t align_start, int64_t start, int64_t* out) {
    if (!cigar || !out) return fail(VAPOR_E_ARG, "NULL argument");
    std::vector<uint32_t> ops;
    parse_cigar_text(cigar, strlen(cigar), ops);
    cigar_walk(ops.data(), ops.size(), align_start, start, out[0], out[1]);
    return VAPOR_OK;
}

}  // extern "C"
