// ratings_io.cpp -- ratings-file ingest for the factorization path: the text formats the BASELINE.json
// shapes are named after -> the (users[], items[], ratings[]) triplets of MatrixFactorizationSGD.java:109, with the
// sparse ids of the files compacted to dense row numbers (ascending original id). Host-only code, no CUDA.
//   MFSGD_FORMAT_TRIPLETS      "user <sep> item <sep> rating [<sep> anything]" per line, <sep> any run of tab, blank, ',',
//                              ';', ':' or '|': MovieLens u.data (tab), ratings.csv (comma, header line), ratings.dat ("::");
//                              fields may be quoted ("1","31","2.5"), a UTF-8 BOM is skipped; lines that do not start with a
//                              digit (headers, comments, blank lines) are skipped -- except a signed number, which is a data
//                              line with an id this interface cannot hold and is reported, not dropped
//   MFSGD_FORMAT_NETFLIX_PRIZE "movie:" lines, each followed by that movie's "customer,rating[,date]" lines
//                              (combined_data_*.txt / mv_*.txt of the Netflix Prize set)
//   MFSGD_FORMAT_AUTO          NETFLIX_PRIZE if the first line that starts with a digit (after a BOM, blanks, quotes) is "<digits>:", else TRIPLETS
// The file is mmapped and parsed by up to 16 threads (one slice of lines each), ids are ranked through a presence table:
// 45 M ratings/s (1.2 GB/s) on 8 cores for a 10 M-line ratings.csv, 17 M ratings/s single-threaded.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <new>
#include <thread>
#include <vector>

#include "../../include/mfsgd.h"

namespace mfsgd {
int set_error(int code, const char* fmt, ...);   // engine.cu: thread-local message behind mfsgd_last_error()
}

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap(const_cast<char*>(p), n);
        if (fd >= 0) close(fd);
    }
};

inline bool is_quote(char c) { return c == '"' || c == '\''; }
// field separators; quotes count as separators so that a quoted CSV ("1","31","2.5") reads like a plain one
inline bool is_sep(char c) { return c == '\t' || c == ' ' || c == ',' || c == ';' || c == ':' || c == '|' || c == '\r' || is_quote(c); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// unsigned decimal integer at s (< end); returns false if none / overflow
inline bool parse_u63(const char*& s, const char* end, int64_t* out) {
    if (s >= end || !is_digit(*s)) return false;
    uint64_t v = 0;
    while (s < end && is_digit(*s)) {
        v = v * 10 + (uint64_t)(*s - '0');
        if (v > (uint64_t)INT64_MAX / 16) return false;
        s++;
    }
    *out = (int64_t)v;
    return true;
}

// decimal number [+-]ddd[.ddd][e[+-]dd] at s; plain digits on the fast path, strtod for exponents
inline bool parse_rating(const char*& s, const char* end, float* out) {
    const char* b = s;
    bool neg = false;
    if (s < end && (*s == '-' || *s == '+')) neg = (*s++ == '-');
    if (s >= end || (!is_digit(*s) && *s != '.')) return false;
    double v = 0.0;
    int digits = 0;
    while (s < end && is_digit(*s)) { v = v * 10.0 + (*s++ - '0'); digits++; }
    if (s < end && *s == '.') {
        s++;
        double scale = 0.1;
        while (s < end && is_digit(*s)) { v += scale * (*s++ - '0'); scale *= 0.1; digits++; }
    }
    if (digits == 0) return false;
    if (s < end && (*s == 'e' || *s == 'E')) {     // rare: let strtod do it on a bounded copy
        char buf[64];
        size_t len = 0;
        const char* t = b;
        while (t < end && len + 1 < sizeof(buf) && !is_sep(*t) && *t != '\n') buf[len++] = *t++;
        buf[len] = 0;
        char* stop = nullptr;
        v = strtod(buf, &stop);
        if (stop == buf) return false;
        s = b + (stop - buf);
        neg = false;
    }
    *out = (float)(neg ? -v : v);
    return std::isfinite(*out);
}

inline const char* skip_seps(const char* s, const char* end) {
    while (s < end && is_sep(*s)) s++;
    return s;
}
inline const char* next_line(const char* s, const char* end) {
    const char* nl = (const char*)memchr(s, '\n', (size_t)(end - s));
    return nl ? nl + 1 : end;
}

struct Slice {
    std::vector<int64_t> u, i;
    std::vector<float> r;
    int64_t max_u = 0, max_i = 0;
    const char* error = nullptr;      // static message; error_at = start of the offending line
    const char* error_at = nullptr;
};

// "<digits>:" alone on the line starting at s (line ends at nl)? -> the movie id
inline bool movie_header(const char* s, const char* nl, int64_t* movie) {
    while (s < nl && (*s == ' ' || *s == '\t')) s++;
    int64_t a = 0;
    if (!parse_u63(s, nl, &a) || s >= nl || *s != ':') return false;
    s++;
    while (s < nl && (*s == ' ' || *s == '\r' || *s == '\n')) s++;
    if (s < nl) return false;
    *movie = a;
    return true;
}

// lines [begin, end) of the file starting at `base` (begin is a line start)
void parse_slice(const char* base, const char* begin, const char* end, int format, Slice& out) {
    const size_t guess = (size_t)(end - begin) / 12 + 16;
    out.u.reserve(guess); out.i.reserve(guess); out.r.reserve(guess);
    int64_t movie = -1;
    if (format == MFSGD_FORMAT_NETFLIX_PRIZE && begin > base) {       // the movie this slice continues: nearest header above
        const char* line_end = begin;                                  // one past the '\n' of the previous line
        while (line_end > base) {
            const char* ls = line_end - 1;                             // at the '\n' (or last byte)
            while (ls > base && ls[-1] != '\n') ls--;
            if (movie_header(ls, line_end, &movie)) break;
            line_end = ls;
        }
    }
    const char* s = begin;
    while (s < end) {
        const char* const line = s;
        const char* nl = next_line(s, end);
        if (line == base && nl - s >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) s += 3;   // UTF-8 BOM
        while (s < nl && (*s == ' ' || *s == '\t' || is_quote(*s))) s++;
        if (s < nl && (*s == '-' || *s == '+') && s + 1 < nl && is_digit(s[1])) {     // a data line, not a header: say so instead of dropping it
            out.error = "signed id (ids are non-negative integers)";
            out.error_at = line;
            return;
        }
        if (s >= nl || !is_digit(*s)) { s = nl; continue; }            // header, comment, blank line
        int64_t a = 0, b = 0;
        float r = 0.f;
        if (format == MFSGD_FORMAT_NETFLIX_PRIZE) {
            if (movie_header(s, nl, &movie)) { s = nl; continue; }
            if (!parse_u63(s, nl, &a)) { out.error = "bad id"; out.error_at = line; return; }
            if (movie < 0) { out.error = "rating before the first \"movie:\" line"; out.error_at = line; return; }
            s = skip_seps(s, nl);
            if (!parse_rating(s, nl, &r)) { out.error = "bad rating"; out.error_at = line; return; }
            out.u.push_back(a); out.i.push_back(movie); out.r.push_back(r);
            out.max_u = std::max(out.max_u, a); out.max_i = std::max(out.max_i, movie);
        } else {
            if (!parse_u63(s, nl, &a)) { out.error = "bad id"; out.error_at = line; return; }
            s = skip_seps(s, nl);
            if (!parse_u63(s, nl, &b)) { out.error = "bad item id"; out.error_at = line; return; }
            s = skip_seps(s, nl);
            if (!parse_rating(s, nl, &r)) { out.error = "bad rating"; out.error_at = line; return; }
            out.u.push_back(a); out.i.push_back(b); out.r.push_back(r);
            out.max_u = std::max(out.max_u, a); out.max_i = std::max(out.max_i, b);
        }
        s = nl;
    }
}

// Runs f(t) for every slice, side by side when asked to. An exception inside a worker thread (std::bad_alloc from a
// growing vector) would end in std::terminate: it is caught there and re-thrown as std::bad_alloc on the calling thread,
// where mfsgd_read_ratings turns it into MFSGD_E_OOM (include/mfsgd.h: no exceptions, no abort).
template <typename F>
void for_each_slice(size_t n_slices, bool parallel, F f) {
    if (!parallel || n_slices == 1) {
        for (size_t t = 0; t < n_slices; t++) f(t);
        return;
    }
    std::atomic<bool> failed(false);
    std::vector<std::thread> pool;
    try {
        for (size_t t = 0; t < n_slices; t++)
            pool.emplace_back([&failed, &f, t]() {
                try {
                    f(t);
                } catch (...) {
                    failed.store(true);
                }
            });
    } catch (...) {          // thread creation failed: finish what runs, then report
        failed.store(true);
    }
    for (auto& th : pool) th.join();
    if (failed.load()) throw std::bad_alloc();
}

// Dense rank (ascending original id) of every id of every slice, written to dense[offset[t] + j]; ids_out = the sorted
// distinct ids. which = 0: users, 1: items. Slices work side by side; only the prefix over the id range is serial.
int compact(const std::vector<Slice>& slices, const std::vector<size_t>& offset, int which, int32_t* dense,
            std::vector<int64_t>& ids_out) {
    auto ids_of = [&](size_t t) -> const std::vector<int64_t>& { return which == 0 ? slices[t].u : slices[t].i; };
    int64_t mx = 0;
    for (size_t t = 0; t < slices.size(); t++) mx = std::max(mx, which == 0 ? slices[t].max_u : slices[t].max_i);
    if (mx < (int64_t)1 << 28) {               // presence table + prefix ranks: O(n / threads + max id)
        std::vector<int32_t> rank((size_t)mx + 2, 0);
        int32_t* const table = rank.data();
        for_each_slice(slices.size(), true, [&](size_t t) {
            for (int64_t v : ids_of(t)) __atomic_store_n(&table[(size_t)v], 1, __ATOMIC_RELAXED);    // same value from every writer
        });
        int32_t next = 0;
        for (size_t v = 0; v <= (size_t)mx; v++) {
            if (table[v]) {
                ids_out.push_back((int64_t)v);
                table[v] = next++;
            }
        }
        for_each_slice(slices.size(), true, [&](size_t t) {
            const std::vector<int64_t>& raw = ids_of(t);
            int32_t* d = dense + offset[t];
            for (size_t j = 0; j < raw.size(); j++) d[j] = table[(size_t)raw[j]];
        });
    } else {                                   // sparse 63-bit ids: sort + binary search
        for (size_t t = 0; t < slices.size(); t++) ids_out.insert(ids_out.end(), ids_of(t).begin(), ids_of(t).end());
        std::sort(ids_out.begin(), ids_out.end());
        ids_out.erase(std::unique(ids_out.begin(), ids_out.end()), ids_out.end());
        if (ids_out.size() > (size_t)INT32_MAX) return -1;
        for_each_slice(slices.size(), true, [&](size_t t) {
            const std::vector<int64_t>& raw = ids_of(t);
            int32_t* d = dense + offset[t];
            for (size_t j = 0; j < raw.size(); j++) d[j] = (int32_t)(std::lower_bound(ids_out.begin(), ids_out.end(), raw[j]) - ids_out.begin());
        });
    }
    return 0;
}

template <typename T>
T* copy_out(const std::vector<T>& v) {
    T* p = (T*)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
    if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

}  // namespace

extern "C" void mfsgd_free_ratings(mfsgd_ratings* r) {
    if (!r) return;
    free(r->users);
    free(r->items);
    free(r->ratings);
    free(r->user_ids);
    free(r->item_ids);
    memset(r, 0, sizeof(*r));
}

static int read_ratings_body(const char* path, int32_t format, mfsgd_ratings* out);

extern "C" int mfsgd_read_ratings(const char* path, int32_t format, mfsgd_ratings* out) {
    try {
        return read_ratings_body(path, format, out);
    } catch (const std::bad_alloc&) {
        if (out) mfsgd_free_ratings(out);
        return mfsgd::set_error(MFSGD_E_OOM, "out of host memory while reading %s", path ? path : "(null)");
    } catch (...) {
        if (out) mfsgd_free_ratings(out);
        return mfsgd::set_error(MFSGD_E_STATE, "internal error while reading %s", path ? path : "(null)");
    }
}

static int read_ratings_body(const char* path, int32_t format, mfsgd_ratings* out) {
    using mfsgd::set_error;
    if (!path || !out) return set_error(MFSGD_E_INVALID_ARG, "path or out is null");
    if (format < MFSGD_FORMAT_AUTO || format > MFSGD_FORMAT_NETFLIX_PRIZE) return set_error(MFSGD_E_INVALID_ARG, "unknown format %d", format);
    memset(out, 0, sizeof(*out));
    Mapped m;
    m.fd = open(path, O_RDONLY);
    if (m.fd < 0) return set_error(MFSGD_E_INVALID_ARG, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(m.fd, &st) != 0 || !S_ISREG(st.st_mode)) return set_error(MFSGD_E_INVALID_ARG, "%s is not a regular file", path);
    m.n = (size_t)st.st_size;
    if (m.n > 0) {
        void* p = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
        if (p == MAP_FAILED) { m.n = 0; return set_error(MFSGD_E_OOM, "mmap of %s failed: %s", path, strerror(errno)); }
        m.p = (const char*)p;
        madvise(p, m.n, MADV_SEQUENTIAL);
    }
    const char* s = m.p;
    const char* const end = m.p + m.n;
    if (format == MFSGD_FORMAT_AUTO) {
        format = MFSGD_FORMAT_TRIPLETS;
        for (const char* t = s; t < end; t = next_line(t, end)) {
            const char* c = t;
            if (t == s && end - c >= 3 && (unsigned char)c[0] == 0xEF && (unsigned char)c[1] == 0xBB && (unsigned char)c[2] == 0xBF) c += 3;   // UTF-8 BOM
            while (c < end && (*c == ' ' || *c == '\t' || is_quote(*c))) c++;
            if (c < end && is_digit(*c)) {
                while (c < end && is_digit(*c)) c++;
                const char* d = c;
                if (d < end && *d == ':') {
                    d++;
                    while (d < end && (*d == ' ' || *d == '\r')) d++;
                    if (d >= end || *d == '\n') format = MFSGD_FORMAT_NETFLIX_PRIZE;
                }
                break;
            }
        }
    }
    // Parse in parallel: the file is cut into one slice per thread at line starts; a Netflix-Prize slice first looks
    // backwards for the "movie:" line it continues. MFSGD_IO_THREADS overrides the thread count (tests force many
    // slices on small files).
    int threads = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    threads = (int)std::min<size_t>((size_t)threads, m.n / ((size_t)4 << 20) + 1);
    if (const char* e = getenv("MFSGD_IO_THREADS")) threads = std::max(1, std::min(256, atoi(e)));
    std::vector<const char*> cut((size_t)threads + 1, end);
    cut[0] = s;
    for (int t = 1; t < threads; t++) {
        const char* c = m.p + m.n * (size_t)t / (size_t)threads;
        if (c < cut[(size_t)t - 1]) c = cut[(size_t)t - 1];
        cut[(size_t)t] = (c == m.p) ? c : next_line(c - 1, end);      // first line start at or after c
    }
    std::vector<Slice> slices((size_t)threads);
    auto work = [&](int t) { parse_slice(m.p, cut[(size_t)t], cut[(size_t)t + 1], format, slices[(size_t)t]); };
    for_each_slice((size_t)threads, threads > 1, [&](size_t t) { work((int)t); });
    std::vector<size_t> offset(slices.size() + 1, 0);
    for (size_t t = 0; t < slices.size(); t++) {
        const Slice& sl = slices[t];
        if (sl.error) {           // first failing slice in file order = first bad line of the file
            int64_t line_no = 1;
            for (const char* c = m.p; c < sl.error_at; c++) line_no += (*c == '\n');
            return set_error(MFSGD_E_INVALID_ARG, "%s:%lld: %s", path, (long long)line_no, sl.error);
        }
        offset[t + 1] = offset[t] + sl.r.size();
    }
    const size_t n = offset.back();
    std::vector<int64_t> uid, iid;
    out->users = (int32_t*)malloc(std::max<size_t>(1, n) * 4);
    out->items = (int32_t*)malloc(std::max<size_t>(1, n) * 4);
    out->ratings = (float*)malloc(std::max<size_t>(1, n) * 4);
    if (!out->users || !out->items || !out->ratings) { mfsgd_free_ratings(out); return set_error(MFSGD_E_OOM, "out of host memory for %zu ratings", n); }
    for_each_slice(slices.size(), true, [&](size_t t) {
        if (!slices[t].r.empty()) memcpy(out->ratings + offset[t], slices[t].r.data(), slices[t].r.size() * sizeof(float));
    });
    if (compact(slices, offset, 0, out->users, uid) != 0 || compact(slices, offset, 1, out->items, iid) != 0) {
        mfsgd_free_ratings(out);
        return set_error(MFSGD_E_INVALID_ARG, "%s: more than 2^31-1 distinct ids", path);
    }
    out->user_ids = copy_out(uid);
    out->item_ids = copy_out(iid);
    if (!out->user_ids || !out->item_ids) { mfsgd_free_ratings(out); return set_error(MFSGD_E_OOM, "out of host memory"); }
    out->n = (int64_t)n;
    out->n_users = (int32_t)uid.size();
    out->n_items = (int32_t)iid.size();
    out->format = format;
    return MFSGD_OK;
}
