// hostio.cu -- native host-side text ingest of the hot path; no device code.
//
//   tab_project_file   the BED projection of a LASTZ-style .tab file: what `awk '!/^#/ {print $1,$3,$4;}'` feeds into
//                      sort | bedtools genomecov (wrappers.py:1120-1128, 827-835, 1201-1220), as dictionary-encoded
//                      scaffold ids + integer columns, parsed from an mmap by several threads;
//   fasta_read_file    every record of a FASTA file with line breaks removed -- the loader behind chromlens / splitFasta
//                      (utils.py:274-309, 502-557; the reference uses Biopython's SeqIO.parse), two parallel passes
//                      (count, compact) over the mmap.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <thread>
#include <unordered_map>

#include "internal.cuh"

namespace mb2 {

namespace {

struct MappedFile {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    explicit MappedFile(const char* path) {
        fd = ::open(path, O_RDONLY);
        MB2_REQUIRE(fd >= 0, -2, std::string("cannot open ") + path);
        struct stat st;
        MB2_REQUIRE(::fstat(fd, &st) == 0, -2, std::string("cannot stat ") + path);
        n = (size_t)st.st_size;
        if (n) {
            void* m = ::mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            MB2_REQUIRE(m != MAP_FAILED, -2, std::string("cannot map ") + path);
            p = (const char*)m;
            ::madvise(m, n, MADV_SEQUENTIAL);
        }
    }
    ~MappedFile() {
        if (p) ::munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
};

int pick_threads(int requested, size_t bytes) {
    int t = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
    t = std::max(1, std::min(t, 64));
    const size_t by_size = std::max<size_t>(1, bytes / (1u << 20));   // at least 1 MiB per thread
    return (int)std::min<size_t>((size_t)t, by_size);
}

template <typename F>
void run_threads(int nt, F&& f) {
    if (nt <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; t++) th.emplace_back([&f, t] { f(t); });
    for (auto& x : th) x.join();
}

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }

struct TabChunk {
    std::vector<int32_t> chrom;                    // chunk-local name ids
    std::vector<int64_t> start, end;
    std::vector<std::string_view> names;           // chunk-local dictionary, first-appearance order (views into the mmap)
    std::string err;
    size_t err_off = 0;
};

// [+-]digits -> int64; false if anything else. '%' characters are ignored: the reference strips them from the projected
// columns with sed 's/%//g' before bedtools reads them (wrappers.py:1125).
inline bool parse_int(const char* b, const char* e, int64_t& v) {
    while (b < e && *b == '%') b++;
    while (e > b && e[-1] == '%') e--;
    if (b == e) return false;
    bool neg = false;
    if (*b == '-' || *b == '+') { neg = *b == '-'; b++; if (b == e) return false; }
    if (e - b > 18) return false;
    int64_t x = 0;
    for (; b < e; b++) {
        const unsigned d = (unsigned)(*b - '0');
        if (d > 9) return false;
        x = x * 10 + (int64_t)d;
    }
    v = neg ? -x : x;
    return true;
}

void parse_tab_range(const char* base, size_t lo, size_t hi, TabChunk& c) {
    std::unordered_map<std::string_view, int32_t> dict;
    const char* p = base + lo;
    const char* const end = base + hi;
    const size_t guess = (hi - lo) / 48 + 16;
    c.chrom.reserve(guess); c.start.reserve(guess); c.end.reserve(guess);
    std::string_view last_name;
    int32_t last_id = -1;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        const char* next = nl ? nl + 1 : end;
        const char* q = p;
        if (le > q && le[-1] == '\r') le--;
        if (q < le && *q != '#') {                 // awk '!/^#/': only a '#' in column one makes a comment line
            const char* fb[4]; const char* fe[4];
            int nf = 0;
            while (nf < 4) {
                while (q < le && is_blank(*q)) q++;
                if (q >= le) break;
                fb[nf] = q;
                while (q < le && !is_blank(*q)) q++;
                fe[nf] = q;
                nf++;
            }
            if (nf > 0) {                          // a line of blanks only is skipped
                int64_t s = 0, e = 0;
                if (nf < 4 || !parse_int(fb[2], fe[2], s) || !parse_int(fb[3], fe[3], e)) {
                    if (c.err.empty()) { c.err = nf < 4 ? "fewer than 4 fields" : "columns 3 and 4 must be integers"; c.err_off = (size_t)(p - base); }
                    return;
                }
                const std::string_view name(fb[0], (size_t)(fe[0] - fb[0]));
                int32_t id;
                if (last_id >= 0 && name == last_name) id = last_id;       // rows of one scaffold come in runs
                else {
                    auto it = dict.find(name);
                    if (it == dict.end()) { id = (int32_t)c.names.size(); dict.emplace(name, id); c.names.push_back(name); }
                    else id = it->second;
                    last_name = name; last_id = id;
                }
                c.chrom.push_back(id); c.start.push_back(s); c.end.push_back(e);
            }
        }
        p = next;
    }
}

}  // namespace

void tab_project_file(const char* path, int nthreads, TabHits& out) {
    out = TabHits();
    MappedFile mf(path);
    if (mf.n == 0) return;
    const int nt = pick_threads(nthreads, mf.n);
    // chunk boundaries at line starts
    std::vector<size_t> cut(nt + 1, 0);
    cut[nt] = mf.n;
    for (int t = 1; t < nt; t++) {
        size_t pos = mf.n / nt * t;
        const char* nl = (const char*)memchr(mf.p + pos, '\n', mf.n - pos);
        cut[t] = nl ? (size_t)(nl - mf.p) + 1 : mf.n;
    }
    for (int t = 1; t <= nt; t++) cut[t] = std::max(cut[t], cut[t - 1]);
    std::vector<TabChunk> chunks(nt);
    run_threads(nt, [&](int t) { parse_tab_range(mf.p, cut[t], cut[t + 1], chunks[t]); });
    for (int t = 0; t < nt; t++) {
        if (!chunks[t].err.empty()) {
            size_t line = 1;
            for (size_t k = 0; k < chunks[t].err_off; k++) line += mf.p[k] == '\n';
            throw Error(-4, std::string(path) + ": line " + std::to_string(line) + ": " + chunks[t].err);
        }
    }
    // merge the dictionaries in file order (global ids = order of first appearance)
    std::unordered_map<std::string_view, int32_t> dict;
    std::vector<std::vector<int32_t>> remap(nt);
    size_t total = 0;
    for (int t = 0; t < nt; t++) {
        remap[t].resize(chunks[t].names.size());
        for (size_t k = 0; k < chunks[t].names.size(); k++) {
            auto it = dict.find(chunks[t].names[k]);
            if (it == dict.end()) {
                const int32_t id = (int32_t)out.names.size();
                dict.emplace(chunks[t].names[k], id);
                out.names.emplace_back(chunks[t].names[k]);
                remap[t][k] = id;
            } else remap[t][k] = it->second;
        }
        total += chunks[t].chrom.size();
    }
    out.chrom.resize(total); out.start.resize(total); out.end.resize(total);
    std::vector<size_t> off(nt + 1, 0);
    for (int t = 0; t < nt; t++) off[t + 1] = off[t] + chunks[t].chrom.size();
    run_threads(nt, [&](int t) {
        const TabChunk& c = chunks[t];
        const size_t o = off[t];
        for (size_t k = 0; k < c.chrom.size(); k++) out.chrom[o + k] = remap[t][c.chrom[k]];
        if (!c.start.empty()) {
            memcpy(out.start.data() + o, c.start.data(), c.start.size() * sizeof(int64_t));
            memcpy(out.end.data() + o, c.end.data(), c.end.size() * sizeof(int64_t));
        }
    });
}

// ---------------------------------------------------------------------------------------------- FASTA
namespace {
inline bool seq_byte(unsigned char c) { return c != '\n' && c != '\r' && c != ' '; }
struct Piece { size_t lo, hi; int rec; uint64_t kept, dst; };
}  // namespace

void fasta_read_file(const char* path, int nthreads, FastaData& out) {
    MappedFile mf(path);
    if (mf.n == 0) return;
    // record starts: '>' at the start of the file or right after a newline
    std::vector<size_t> starts;
    for (const char* p = mf.p; p < mf.p + mf.n;) {
        const char* g = (const char*)memchr(p, '>', (size_t)(mf.p + mf.n - p));
        if (!g) break;
        if (g == mf.p || g[-1] == '\n') starts.push_back((size_t)(g - mf.p));
        p = g + 1;
    }
    const int nrec = (int)starts.size();
    out.ids.resize(nrec); out.headers.resize(nrec); out.off.assign(nrec + 1, 0);
    std::vector<Piece> pieces;
    constexpr size_t PIECE = 4u << 20;
    for (int r = 0; r < nrec; r++) {
        const size_t s = starts[r], e = r + 1 < nrec ? starts[r + 1] : mf.n;
        const char* nl = (const char*)memchr(mf.p + s, '\n', e - s);
        size_t he = nl ? (size_t)(nl - mf.p) : e;
        const size_t body = nl ? he + 1 : e;
        while (he > s + 1 && mf.p[he - 1] == '\r') he--;
        out.headers[r].assign(mf.p + s + 1, he - (s + 1));
        const std::string& h = out.headers[r];
        size_t a = 0;
        while (a < h.size() && isspace((unsigned char)h[a])) a++;
        size_t b = a;
        while (b < h.size() && !isspace((unsigned char)h[b])) b++;
        out.ids[r] = h.substr(a, b - a);
        for (size_t lo = body; lo < e; lo += PIECE) pieces.push_back(Piece{lo, std::min(e, lo + PIECE), r, 0, 0});
    }
    const int nt = pick_threads(nthreads, mf.n);
    const size_t np = pieces.size();
    // pass 1: bytes kept per piece
    run_threads(nt, [&](int t) {
        for (size_t k = (size_t)t; k < np; k += (size_t)nt) {
            const unsigned char* p = (const unsigned char*)mf.p + pieces[k].lo;
            const size_t n = pieces[k].hi - pieces[k].lo;
            uint64_t cnt = 0;
            for (size_t i = 0; i < n; i++) cnt += seq_byte(p[i]) ? 1u : 0u;
            pieces[k].kept = cnt;
        }
    });
    uint64_t total = 0;
    for (size_t k = 0; k < np; k++) { pieces[k].dst = total; total += pieces[k].kept; out.off[pieces[k].rec + 1] += pieces[k].kept; }
    for (int r = 0; r < nrec; r++) out.off[r + 1] += out.off[r];
    out.total = total;
    out.seq = (uint8_t*)malloc(total + 1);           // pages are first touched by the compacting threads
    MB2_REQUIRE(out.seq != nullptr, -5, "fasta_read_file: out of memory");
    // pass 2: compact
    run_threads(nt, [&](int t) {
        for (size_t k = (size_t)t; k < np; k += (size_t)nt) {
            const unsigned char* p = (const unsigned char*)mf.p + pieces[k].lo;
            const unsigned char* const e = (const unsigned char*)mf.p + pieces[k].hi;
            uint8_t* d = out.seq + pieces[k].dst;
            while (p < e) {                                   // whole lines by memcpy when they hold nothing to drop
                const unsigned char* nl = (const unsigned char*)memchr(p, '\n', (size_t)(e - p));
                const unsigned char* le = nl ? nl : e;
                const size_t n = (size_t)(le - p);
                if (n && !memchr(p, '\r', n) && !memchr(p, ' ', n)) { memcpy(d, p, n); d += n; }
                else for (size_t i = 0; i < n; i++) if (seq_byte(p[i])) *d++ = p[i];
                p = nl ? nl + 1 : e;
            }
        }
    });
}


// ---------------------------------------------------------------------------------------------- .tab formatter
// The awk/sed/sort filter that follows every LASTZ call of the reference (wrappers.py:1044-1056): keep length1 >= minLen and
// printed identity ('%.1f') >= minIdt, print the 10 columns, sort each (target, query) block by start1 numerically and then by
// the whole line's bytes. Blocks come out ordered by (t_id, q_id).
namespace {
inline char* put_int(char* p, long long v) {
    char tmp[24];
    int k = 0;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (k) *p++ = tmp[--k];
    return p;
}
// '%.1f' of 100*nm/nc exactly as printf rounds the double; the integer fast path is taken unless the value sits within 1e-6 of
// a rounding tie, where the C library decides
inline char* put_pct(char* p, int nm, int nc, double& printed) {
    const double ratio = nc > 0 ? 100.0 * (double)nm / (double)nc : 0.0;
    const double t = ratio * 10.0 + 0.5;
    const double fl = floor(t);
    if (ratio >= 0.0 && ratio < 1e7 && t - fl > 1e-6 && (fl + 1.0) - t > 1e-6) {
        const long long tenths = (long long)fl;
        p = put_int(p, tenths / 10); *p++ = '.'; *p++ = (char)('0' + tenths % 10);
        printed = (double)tenths / 10.0;
        return p;
    }
    char buf[40];
    const int n = snprintf(buf, sizeof(buf), "%.1f", ratio);
    memcpy(p, buf, (size_t)n);
    printed = strtod(buf, nullptr);
    return p + n;
}
}  // namespace

void format_tab_blocks(const int32_t* t_id, const int32_t* q_id, const int32_t* strand, const int32_t* start1, const int32_t* end1,
                       const int32_t* start2, const int32_t* end2, const int32_t* score, const int32_t* nmatch, const int32_t* ncols,
                       uint64_t n, const char* const* tnames, int nt, const char* const* qnames, int nq, double min_len, double min_idt,
                       TabText& out) {
    out = TabText();
    struct Row { int32_t t, q, s1; uint32_t len; size_t off; };
    std::vector<Row> rows;
    rows.reserve(n);
    std::vector<size_t> tlen(nt), qlen(nq);
    size_t maxname = 0;
    for (int k = 0; k < nt; k++) { tlen[k] = strlen(tnames[k]); maxname = std::max(maxname, tlen[k]); }
    for (int k = 0; k < nq; k++) { qlen[k] = strlen(qnames[k]); maxname = std::max(maxname, qlen[k]); }
    std::vector<char> pool;
    pool.resize((size_t)n * (2 * maxname + 96) + 16);
    char* w = pool.data();
    for (uint64_t k = 0; k < n; k++) {
        const long long len1 = (long long)end1[k] - (long long)start1[k] + 1;
        if ((double)len1 < min_len) continue;                                    // awk '0+$5 >= minLen'
        MB2_REQUIRE(t_id[k] >= 0 && t_id[k] < nt && q_id[k] >= 0 && q_id[k] < nq, -2, "format_tab: scaffold index out of range");
        char* const r0 = w;
        memcpy(w, tnames[t_id[k]], tlen[t_id[k]]); w += tlen[t_id[k]];
        *w++ = '\t'; *w++ = '+'; *w++ = '\t';
        w = put_int(w, start1[k]); *w++ = '\t';
        w = put_int(w, end1[k]); *w++ = '\t';
        memcpy(w, qnames[q_id[k]], qlen[q_id[k]]); w += qlen[q_id[k]];
        *w++ = '\t'; *w++ = strand[k] ? '-' : '+'; *w++ = '\t';
        w = put_int(w, start2[k]); *w++ = '\t';
        w = put_int(w, end2[k]); *w++ = '\t';
        w = put_int(w, score[k]); *w++ = '\t';
        double printed;
        w = put_pct(w, nmatch[k], ncols[k], printed);                            // what LASTZ prints, and what awk then compares
        *w++ = '\n';
        if (printed < min_idt) { w = r0; continue; }                             // awk '0+$13 >= minIdt'
        rows.push_back(Row{t_id[k], q_id[k], start1[k], (uint32_t)(w - r0), (size_t)(r0 - pool.data())});
    }
    const char* base = pool.data();
    std::sort(rows.begin(), rows.end(), [base](const Row& a, const Row& b) {
        if (a.t != b.t) return a.t < b.t;
        if (a.q != b.q) return a.q < b.q;
        if (a.s1 != b.s1) return a.s1 < b.s1;
        const int c = memcmp(base + a.off, base + b.off, std::min(a.len, b.len));   // bytes, shorter prefix first
        return c != 0 ? c < 0 : a.len < b.len;
    });
    out.text.reserve((size_t)(w - pool.data()));
    for (size_t k = 0; k < rows.size(); k++) {
        if (k == 0 || rows[k].t != rows[k - 1].t || rows[k].q != rows[k - 1].q) {
            out.t_id.push_back(rows[k].t); out.q_id.push_back(rows[k].q); out.off.push_back(out.text.size()); out.nrows.push_back(0);
        }
        out.nrows.back()++;
        out.text.append(base + rows[k].off, rows[k].len);
    }
    out.off.push_back(out.text.size());
}


// ---------------------------------------------------------------------------------------------- mimeo map: .tab -> GFF3 rows
// import_Align + writeGFFlines of the reference (wrappers.py:33-117, 443-522) without the DataFrame: rows whose first
// non-blank character is not '#', split on white space; keep int(end1) - int(start1) >= minLen and float(identity) >= minIdt;
// stable sort by the STRING values of (tName, tStart, tEnd, tStrand); UID = prefix_<row number zero-filled to the width of
// the row count>; one GFF3 line per row with every field printed exactly as it stands in the file.
namespace {
struct MapRow { uint64_t off; uint16_t fo[10], fl[10]; };
struct MapChunk { std::vector<MapRow> rows; std::string err; size_t err_off = 0; };
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

void parse_map_range(const char* base, size_t lo, size_t hi, double min_len, double min_idt, MapChunk& c) {
    const char* p = base + lo;
    const char* const end = base + hi;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        const char* q = p;
        while (q < le && is_space(*q)) q++;
        if (!(q < le && *q == '#')) {
            MapRow r; r.off = (uint64_t)(p - base);
            int nf = 0;
            while (nf < 10) {
                while (q < le && is_space(*q)) q++;
                if (q >= le) break;
                const char* b = q;
                while (q < le && !is_space(*q)) q++;
                if (b - p > 65535 || q - b > 65535) { nf = -1; break; }
                r.fo[nf] = (uint16_t)(b - p); r.fl[nf] = (uint16_t)(q - b);
                nf++;
            }
            int64_t s1 = 0, e1 = 0;
            double idt = 0;
            bool ok = nf == 10 && parse_int(p + r.fo[2], p + r.fo[2] + r.fl[2], s1) && parse_int(p + r.fo[3], p + r.fo[3] + r.fl[3], e1);
            if (ok) {
                char tmp[64];
                const size_t L = r.fl[9];
                ok = L > 0 && L < sizeof(tmp);
                if (ok) { memcpy(tmp, p + r.fo[9], L); tmp[L] = 0; char* endp = nullptr; idt = strtod(tmp, &endp); ok = endp == tmp + L; }
            }
            if (!ok) {
                if (c.err.empty()) { c.err = nf != 10 ? "fewer than 10 fields" : "start1/end1 must be integers and identity a number"; c.err_off = (size_t)(p - base); }
                return;
            }
            if ((double)(e1 - s1) >= min_len && idt >= min_idt) c.rows.push_back(r);
        }
        p = nl ? nl + 1 : end;
    }
}
}  // namespace

uint64_t map_gff_rows(const char* path, const char* prefix, double min_len, double min_idt, const char* ftype, int nthreads, std::string& out) {
    out.clear();
    MappedFile mf(path);
    if (mf.n == 0) return 0;
    const int nt = pick_threads(nthreads, mf.n);
    std::vector<size_t> cut(nt + 1, 0);
    cut[nt] = mf.n;
    for (int t = 1; t < nt; t++) {
        const size_t pos = mf.n / nt * t;
        const char* nl = (const char*)memchr(mf.p + pos, '\n', mf.n - pos);
        cut[t] = nl ? (size_t)(nl - mf.p) + 1 : mf.n;
    }
    for (int t = 1; t <= nt; t++) cut[t] = std::max(cut[t], cut[t - 1]);
    std::vector<MapChunk> chunks(nt);
    run_threads(nt, [&](int t) { parse_map_range(mf.p, cut[t], cut[t + 1], min_len, min_idt, chunks[t]); });
    size_t total = 0;
    for (int t = 0; t < nt; t++) {
        if (!chunks[t].err.empty()) {
            size_t line = 1;
            for (size_t k = 0; k < chunks[t].err_off; k++) line += mf.p[k] == '\n';
            throw Error(-4, std::string(path) + ": line " + std::to_string(line) + ": " + chunks[t].err);
        }
        total += chunks[t].rows.size();
    }
    std::vector<MapRow> rows;
    rows.reserve(total);
    for (int t = 0; t < nt; t++) rows.insert(rows.end(), chunks[t].rows.begin(), chunks[t].rows.end());   // file order
    const char* base = mf.p;
    auto field = [base](const MapRow& r, int f) { return std::string_view(base + r.off + r.fo[f], r.fl[f]); };
    std::stable_sort(rows.begin(), rows.end(), [&](const MapRow& a, const MapRow& b) {
        for (int f : {0, 2, 3, 1}) {                                   // tName, tStart, tEnd, tStrand as strings
            const int c = field(a, f).compare(field(b, f));
            if (c) return c < 0;
        }
        return false;
    });
    const std::string stem = (prefix && *prefix) ? prefix : "BHit";
    const int width = (int)std::to_string(total).size();
    const std::string ft = ftype ? ftype : "BHit";
    out.reserve(total * 160);
    char num[32];
    for (size_t k = 0; k < rows.size(); k++) {
        const MapRow& r = rows[k];
        auto add = [&](int f) { out.append(base + r.off + r.fo[f], r.fl[f]); };
        add(0); out += "\tmimeo-map\t"; out += ft; out += '\t'; add(2); out += '\t'; add(3); out += '\t'; add(8); out += '\t'; add(1);
        out += "\t.\tID="; out += stem; out += '_';
        snprintf(num, sizeof(num), "%0*llu", width, (unsigned long long)(k + 1));
        out += num;
        out += ";identity="; add(9); out += ";B_locus="; add(4); out += '_'; add(5); out += '_'; add(6); out += '_'; add(7); out += '\n';
    }
    return total;
}


// ------------------------------------------------------------------------------------------------
// fasta_split_file: one `<outdir>/<id>.fa` per record, header line kept whole, sequence wrapped at `width` columns -- what
// splitFasta (utils.py:274-309: SeqIO.parse + SeqIO.write per record) leaves in --adir / --bdir. Records are written by
// `nthreads` workers (one file each). With `unique`, the records before the first repeated id are written and the call then
// fails with -7 naming the id (the reference exits at that record); without it a later record replaces an earlier file of
// the same id, so only the last one of each id is written. Returns the number of files written.
// ------------------------------------------------------------------------------------------------
namespace {
void write_all(int fd, const char* p, size_t n, const std::string& path) {
    while (n) {
        const ssize_t w = ::write(fd, p, n);
        if (w < 0) throw Error(-5, "write " + path + ": " + strerror(errno));
        p += w; n -= (size_t)w;
    }
}
void write_record(const std::string& path, const std::string& header, const uint8_t* seq, uint64_t n, int width) {
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) throw Error(-5, "open " + path + ": " + strerror(errno));
    try {
        std::string buf;
        const size_t lines_per_block = std::max<size_t>(1, ((size_t)1 << 20) / (size_t)(width + 1));
        buf.reserve(lines_per_block * (size_t)(width + 1) + header.size() + 2);
        buf.push_back('>'); buf += header; buf.push_back('\n');
        uint64_t i = 0;
        while (i < n) {
            for (size_t l = 0; l < lines_per_block && i < n; l++) {
                const uint64_t m = std::min<uint64_t>((uint64_t)width, n - i);
                buf.append((const char*)seq + i, (size_t)m);
                buf.push_back('\n');
                i += m;
            }
            write_all(fd, buf.data(), buf.size(), path);
            buf.clear();
        }
        if (!buf.empty()) write_all(fd, buf.data(), buf.size(), path);
    } catch (...) { ::close(fd); throw; }
    if (::close(fd) != 0) throw Error(-5, "close " + path + ": " + strerror(errno));
}
}  // namespace

uint64_t fasta_split_file(const char* path, const char* outdir, bool unique, int width, int nthreads) {
    MB2_REQUIRE(width > 0, -2, "fasta_split_file: width must be positive");
    FastaData f;
    fasta_read_file(path, nthreads, f);
    const size_t n = f.ids.size();
    std::unordered_map<std::string_view, size_t> last;     // id -> index of its last record
    size_t stop = n;                                        // first record whose id was seen before
    for (size_t r = 0; r < n; r++) {
        auto it = last.find(f.ids[r]);
        if (it != last.end() && stop == n) stop = r;
        last[f.ids[r]] = r;
    }
    const size_t upto = unique ? stop : n;
    std::vector<size_t> todo;
    for (size_t r = 0; r < upto; r++)
        if (unique || last[f.ids[r]] == r) todo.push_back(r);
    for (size_t r : todo)
        MB2_REQUIRE(!f.ids[r].empty() && f.ids[r].find('/') == std::string::npos, -2,
                    std::string(path) + ": record id '" + f.ids[r] + "' cannot name a file");
    const int nt = std::max(1, std::min<int>(nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency(), (int)std::max<size_t>(todo.size(), 1)));
    std::vector<std::string> errs(nt);
    std::atomic<size_t> next{0};
    const std::string dir(outdir);
    run_threads(nt, [&](int t) {
        try {
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= todo.size()) break;
                const size_t r = todo[k];
                write_record(dir + "/" + f.ids[r] + ".fa", f.headers[r], f.seq + f.off[r], f.off[r + 1] - f.off[r], width);
            }
        } catch (const std::exception& e) { errs[t] = e.what(); }
    });
    for (const auto& e : errs) if (!e.empty()) throw Error(-5, e);
    if (unique && stop < n) throw Error(-7, f.ids[stop]);   // message = the repeated id
    return todo.size();
}


// ------------------------------------------------------------------------------------------------
// format_segment_gff: the awk that ends every coverage block of the reference's script (wrappers.py:1166-1173, 885-891,
// 1257-1264): one GFF3 feature row per merged run,
//   chrom \t source \t label \t start \t end \t . \t + \t . \t ID=<prefix>_%05d \n      counter from first_id within the block
// ------------------------------------------------------------------------------------------------
namespace {
inline char* put_uint(char* p, uint64_t v, int min_width) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int k = n; k < min_width; k++) *p++ = '0';
    while (n) *p++ = tmp[--n];
    return p;
}
}  // namespace

uint64_t format_segment_gff(const int32_t* chrom, const int32_t* start, const int32_t* end, uint64_t n, const char* const* names,
                            int nnames, const char* source, const char* label, const char* prefix, uint64_t first_id, int nthreads,
                            std::string& out) {
    out.clear();
    if (n == 0) return 0;
    const std::string mid = std::string("\t") + source + "\t" + label + "\t";
    const std::string pre = std::string("\t.\t+\t.\tID=") + prefix + "_";
    std::vector<std::string_view> nm((size_t)nnames);
    for (int k = 0; k < nnames; k++) nm[k] = names[k];
    for (uint64_t i = 0; i < n; i++)
        MB2_REQUIRE(chrom[i] >= 0 && chrom[i] < nnames, -2, "format_segment_gff: scaffold index out of range");
    const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>(nthreads > 0 ? (uint64_t)nthreads : std::thread::hardware_concurrency(), n / 4096 + 1));
    std::vector<uint64_t> cut(nt + 1), bytes(nt + 1, 0);
    for (int t = 0; t <= nt; t++) cut[t] = n * (uint64_t)t / (uint64_t)nt;
    const size_t fixed = mid.size() + pre.size() + 11 + 11 + 1 + 20 + 1;     // two ints, tab, id digits, newline
    run_threads(nt, [&](int t) {
        uint64_t b = 0;
        for (uint64_t i = cut[t]; i < cut[t + 1]; i++) b += nm[chrom[i]].size() + fixed;
        bytes[t + 1] = b;                                                        // upper bound of the thread's text
    });
    for (int t = 0; t < nt; t++) bytes[t + 1] += bytes[t];
    std::string buf(bytes[nt], '\0');
    std::vector<uint64_t> used(nt, 0);
    run_threads(nt, [&](int t) {
        char* const p0 = &buf[bytes[t]];
        char* p = p0;
        for (uint64_t i = cut[t]; i < cut[t + 1]; i++) {
            const std::string_view& c = nm[chrom[i]];
            memcpy(p, c.data(), c.size()); p += c.size();
            memcpy(p, mid.data(), mid.size()); p += mid.size();
            p = put_int(p, (long long)start[i]); *p++ = '\t';
            p = put_int(p, (long long)end[i]);
            memcpy(p, pre.data(), pre.size()); p += pre.size();
            p = put_uint(p, first_id + i, 5); *p++ = '\n';
        }
        used[t] = (uint64_t)(p - p0);
    });
    uint64_t total = 0;                                                          // close the gaps between the threads' pieces
    for (int t = 0; t < nt; t++) {
        if (bytes[t] != total) memmove(&buf[total], &buf[bytes[t]], used[t]);
        total += used[t];
    }
    buf.resize(total);
    out.swap(buf);
    return n;
}

}  // namespace mb2
