// ratings_io.cpp -- ratings-file ingest for the factorization path: the text formats the BASELINE.json
// shapes are named after -> the (users[], items[], ratings[]) triplets of MatrixFactorizationSGD.java:109, with the
// sparse ids of the files compacted to dense row numbers (ascending original id). Host-only code, no CUDA.
//   MFSGD_FORMAT_TRIPLETS      "user <sep> item <sep> rating [<sep> anything]" per line, <sep> any run of tab, blank, ',',
//                              ';', ':' or '|': MovieLens u.data (tab), ratings.csv (comma, header line), ratings.dat ("::");
//                              lines that do not start with a digit (headers, comments, blank lines) are skipped
//   MFSGD_FORMAT_NETFLIX_PRIZE "movie:" lines, each followed by that movie's "customer,rating[,date]" lines
//                              (combined_data_*.txt / mv_*.txt of the Netflix Prize set)
//   MFSGD_FORMAT_AUTO          NETFLIX_PRIZE if the first line that starts with a digit is "<digits>:", else TRIPLETS
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

inline bool is_sep(char c) { return c == '\t' || c == ' ' || c == ',' || c == ';' || c == ':' || c == '|' || c == '\r'; }
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

// dense rank of every id (ascending original id); fills ids_out with the sorted distinct ids
int compact(const std::vector<int64_t>& raw, int32_t* dense, std::vector<int64_t>& ids_out) {
    int64_t mx = 0;
    for (int64_t v : raw) mx = std::max(mx, v);
    if (mx < (int64_t)1 << 28) {            // presence table + prefix ranks: O(n + max id)
        std::vector<int32_t> rank((size_t)mx + 2, 0);
        for (int64_t v : raw) rank[(size_t)v] = 1;
        int32_t next = 0;
        for (size_t v = 0; v <= (size_t)mx; v++) {
            if (rank[v]) {
                ids_out.push_back((int64_t)v);
                rank[v] = next++;
            }
        }
        for (size_t t = 0; t < raw.size(); t++) dense[t] = rank[(size_t)raw[t]];
    } else {                                   // sparse 63-bit ids: sort + binary search
        ids_out = raw;
        std::sort(ids_out.begin(), ids_out.end());
        ids_out.erase(std::unique(ids_out.begin(), ids_out.end()), ids_out.end());
        if (ids_out.size() > (size_t)INT32_MAX) return -1;
        for (size_t t = 0; t < raw.size(); t++)
            dense[t] = (int32_t)(std::lower_bound(ids_out.begin(), ids_out.end(), raw[t]) - ids_out.begin());
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

extern "C" int mfsgd_read_ratings(const char* path, int32_t format, mfsgd_ratings* out) {
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
            while (c < end && (*c == ' ' || *c == '\t')) c++;
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
    std::vector<int64_t> ru, ri;
    std::vector<float> rr;
    const size_t guess = m.n / 12 + 16;
    ru.reserve(guess); ri.reserve(guess); rr.reserve(guess);
    int64_t line_no = 0, movie = -1;
    while (s < end) {
        line_no++;
        const char* line = s;
        const char* nl = next_line(s, end);
        while (s < nl && (*s == ' ' || *s == '\t')) s++;
        if (s >= nl || !is_digit(*s)) { s = nl; continue; }            // header, comment, blank line
        int64_t a = 0, b = 0;
        float r = 0.f;
        if (!parse_u63(s, nl, &a)) return set_error(MFSGD_E_INVALID_ARG, "%s:%lld: bad id", path, (long long)line_no);
        if (format == MFSGD_FORMAT_NETFLIX_PRIZE) {
            const char* t = s;
            if (t < nl && *t == ':') {
                t++;
                while (t < nl && (*t == ' ' || *t == '\r' || *t == '\n')) t++;
                if (t >= nl) { movie = a; s = nl; continue; }            // "movie:" header
            }
            if (movie < 0) return set_error(MFSGD_E_INVALID_ARG, "%s:%lld: rating before the first \"movie:\" line", path, (long long)line_no);
            s = skip_seps(s, nl);
            if (!parse_rating(s, nl, &r)) return set_error(MFSGD_E_INVALID_ARG, "%s:%lld: bad rating", path, (long long)line_no);
            ru.push_back(a); ri.push_back(movie); rr.push_back(r);
        } else {
            s = skip_seps(s, nl);
            if (!parse_u63(s, nl, &b)) return set_error(MFSGD_E_INVALID_ARG, "%s:%lld: bad item id", path, (long long)line_no);
            s = skip_seps(s, nl);
            if (!parse_rating(s, nl, &r)) return set_error(MFSGD_E_INVALID_ARG, "%s:%lld: bad rating", path, (long long)line_no);
            ru.push_back(a); ri.push_back(b); rr.push_back(r);
        }
        (void)line;
        s = nl;
    }
    const size_t n = rr.size();
    std::vector<int64_t> uid, iid;
    out->users = (int32_t*)malloc(std::max<size_t>(1, n) * 4);
    out->items = (int32_t*)malloc(std::max<size_t>(1, n) * 4);
    out->ratings = copy_out(rr);
    if (!out->users || !out->items || !out->ratings) { mfsgd_free_ratings(out); return set_error(MFSGD_E_OOM, "out of host memory for %zu ratings", n); }
    if (compact(ru, out->users, uid) != 0 || compact(ri, out->items, iid) != 0) {
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
