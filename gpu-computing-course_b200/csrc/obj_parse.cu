// obj_parse.cu — host-side OBJ ingest with the reference parser's dialect, multi-threaded.
//
// Reference: loadObj (reference CollisionDetection/load_obj.h:24-103) reads the file line by line
// with getline(buffer, 255) and sscanf: 4.2 s per million triangles on one core, three
// orders of magnitude more than the GPU pipeline it feeds. Same observable behaviour here:
//   - only lines that start with "v " are vertices, parsed like sscanf("v %f %f %f") (load_obj.h:48-52);
//   - only lines that start with "f " are faces, parsed like sscanf("f %d/%d %d/%d %d/%d")
//     (load_obj.h:68); the /vt integers are read and dropped; indices are 1-based (load_obj.h:81-83);
//   - every other line is skipped; triangle ID = order of the face lines (load_obj.h:94);
//   - a last line without '\n' is dropped (load_obj.h:41 tests eof() after getline);
//   - a face may only name vertices defined on earlier lines (load_obj.h:76-79,89).
// Where the reference exit()s (load_obj.h:34,60,73) or misbehaves (lines of 255+ characters put the
// stream into a failed state and the loop never ends), this returns B200CD_E_IO / B200CD_E_PARSE with
// the line number of the FIRST offending line.
// Method: the file is cut into chunks at line boundaries; pass 1 counts vertex / face / all lines per
// chunk, a prefix sum gives every chunk its output offsets and the number of vertices defined before it,
// pass 2 parses the chunks in parallel (std::from_chars: correctly rounded like strtof/%f, no locale).
#include <omp.h>

#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200cd {

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }  // isspace minus '\n'

// "%f": skip white space, optional sign, decimal / inf / nan (hex floats and out-of-range values go through strtof)
inline bool scan_float(const char*& p, const char* end, float& out) {
    while (p < end && is_space(*p)) ++p;
    if (p >= end) return false;
    const char* q = p;
    if (*q == '+') ++q;  // from_chars does not take a leading '+'
    auto r = std::from_chars(q, end, out, std::chars_format::general);
    const bool hexish = r.ec == std::errc() && r.ptr < end && (*r.ptr == 'x' || *r.ptr == 'X') && r.ptr - q <= 2;
    if (r.ec == std::errc() && !hexish) {
        p = r.ptr;
        return true;
    }
    // rare: hex float, out of range (%f still stores +-inf / 0 / a denormal), anything from_chars refuses
    char tmp[64];
    size_t len = std::min<size_t>(sizeof tmp - 1, (size_t)(end - p));
    memcpy(tmp, p, len);
    tmp[len] = '\0';
    char* stop = nullptr;
    out = strtof(tmp, &stop);
    if (stop == tmp) return false;
    p += stop - tmp;
    return true;
}

// "%d": skip white space, optional sign, at least one digit
inline bool scan_int(const char*& p, const char* end, long long& out) {
    while (p < end && is_space(*p)) ++p;
    if (p >= end) return false;
    bool neg = false;
    if (*p == '+' || *p == '-') {
        neg = *p == '-';
        ++p;
    }
    if (p >= end || *p < '0' || *p > '9') return false;
    long long v = 0;
    while (p < end && *p >= '0' && *p <= '9') {
        if (v < (1ll << 40)) v = v * 10 + (*p - '0');  // saturate: anything this large is out of range anyway
        ++p;
    }
    out = neg ? -v : v;
    return true;
}

struct Chunk {
    size_t begin = 0, end = 0;          // byte range, ends right after a '\n'
    size_t nv = 0, nf = 0, nlines = 0;  // counted in pass 1
    size_t v0 = 0, f0 = 0, line0 = 0;   // prefix sums
    size_t err_line = 0;                // first bad line in this chunk (1-based, global), 0 = none
    int err_kind = 0;                   // 1 vertex format, 2 face format, 3 face references an undefined vertex, 4 line too long
};

}  // namespace

int obj_parse(const char* path, std::vector<float>& xyz, std::vector<uint32_t>& idx, std::string& err) {
    FILE* fp = fopen(path, "rb");
    if (!fp) {
        err = std::string("cannot open ") + path;  // load_obj.h:31-35
        return B200CD_E_IO;
    }
    std::vector<char> text;
    {
        fseek(fp, 0, SEEK_END);
        long sz = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        if (sz > 0) {
            text.resize((size_t)sz);
            size_t got = fread(text.data(), 1, text.size(), fp);
            text.resize(got);
        }
        // files whose size cannot be told up front (pipes): read the rest the slow way
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof chunk, fp)) > 0) text.insert(text.end(), chunk, chunk + got);
        const bool bad = ferror(fp) != 0;
        fclose(fp);
        if (bad) {
            err = std::string("read error on ") + path;
            return B200CD_E_IO;
        }
    }
    // a last line without '\n' is dropped: load_obj.h:41 tests eof() after getline
    size_t usable = text.size();
    while (usable > 0 && text[usable - 1] != '\n') --usable;
    const char* base = text.data();

    const int nthreads = std::max(1, std::min(omp_get_max_threads(), 64));
    const int nchunks = (int)std::min<size_t>((size_t)nthreads * 4, std::max<size_t>(1, usable / (1 << 16)));
    std::vector<Chunk> chunks((size_t)nchunks);
    {
        size_t pos = 0;
        for (int c = 0; c < nchunks; ++c) {
            chunks[c].begin = pos;
            size_t target = (c + 1 == nchunks) ? usable : std::max(pos, usable / nchunks * (size_t)(c + 1));
            while (target < usable && text[target - 1] != '\n') ++target;  // advance to just after a newline
            if (target > usable) target = usable;
            chunks[c].end = pos = target;
        }
    }
    // ---- pass 1: count
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int c = 0; c < nchunks; ++c) {
        Chunk& ch = chunks[c];
        const char* p = base + ch.begin;
        const char* e = base + ch.end;
        while (p < e) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(e - p)));
            const size_t len = (size_t)(nl - p);
            ++ch.nlines;
            if (len >= 2 && p[1] == ' ') {
                if (p[0] == 'v') ++ch.nv;
                else if (p[0] == 'f') ++ch.nf;
            }
            p = nl + 1;
        }
    }
    size_t tv = 0, tf = 0, tl = 0;
    for (Chunk& ch : chunks) {
        ch.v0 = tv; ch.f0 = tf; ch.line0 = tl;
        tv += ch.nv; tf += ch.nf; tl += ch.nlines;
    }
    if (tv > (1ull << B200CD_MAX_TRIS_LOG2) || tf > (1ull << B200CD_MAX_TRIS_LOG2)) {
        err = "more than 2^30 vertices or triangles";
        return B200CD_E_TOOBIG;
    }
    xyz.assign(3 * tv, 0.f);
    idx.assign(3 * tf, 0u);
    // ---- pass 2: parse
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int c = 0; c < nchunks; ++c) {
        Chunk& ch = chunks[c];
        const char* p = base + ch.begin;
        const char* e = base + ch.end;
        size_t v = ch.v0, f = ch.f0, line = ch.line0;
        auto fail = [&](int kind) {
            if (!ch.err_line) { ch.err_line = line; ch.err_kind = kind; }
        };
        while (p < e && !ch.err_line) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(e - p)));
            const size_t len = (size_t)(nl - p);
            ++line;
            if (len >= 255) {  // getline(buffer, 255) would set failbit and the reference never terminates
                fail(4);
            } else if (len >= 2 && p[1] == ' ' && p[0] == 'v') {  // load_obj.h:48-61
                const char* q = p + 1;
                float a, b, d;
                if (scan_float(q, nl, a) && scan_float(q, nl, b) && scan_float(q, nl, d)) {
                    xyz[3 * v] = a; xyz[3 * v + 1] = b; xyz[3 * v + 2] = d;
                    ++v;
                } else {
                    fail(1);
                }
            } else if (len >= 2 && p[1] == ' ' && p[0] == 'f') {  // load_obj.h:64-102
                const char* q = p + 1;
                long long vi[3], ti;
                bool ok = true;
                for (int k = 0; k < 3 && ok; ++k) {
                    ok = scan_int(q, nl, vi[k]) && q < nl && *q == '/';  // the '/' literal matches without skipping blanks
                    if (ok) {
                        ++q;
                        ok = scan_int(q, nl, ti);
                    }
                }
                if (!ok) {
                    fail(2);
                } else {
                    const long long v_size = (long long)v + 1;  // vertices defined so far + 1 (load_obj.h:76-79,89)
                    if (vi[0] < 1 || vi[1] < 1 || vi[2] < 1 || vi[0] >= v_size || vi[1] >= v_size || vi[2] >= v_size) {
                        fail(3);
                    } else {
                        idx[3 * f] = (uint32_t)(vi[0] - 1); idx[3 * f + 1] = (uint32_t)(vi[1] - 1); idx[3 * f + 2] = (uint32_t)(vi[2] - 1);
                        ++f;
                    }
                }
            }
            p = nl + 1;
        }
    }
    for (const Chunk& ch : chunks) {  // chunks are in file order: the first one with an error holds the first bad line
        if (!ch.err_line) continue;
        static const char* what[] = {"", "vertex not in 'v x y z' format", "face not in 'f v/vt v/vt v/vt' format",
                                     "face references a vertex not yet defined", "longer than 254 characters"};
        err = "line " + std::to_string(ch.err_line) + ": " + what[ch.err_kind];
        xyz.clear();
        idx.clear();
        return B200CD_E_PARSE;
    }
    return B200CD_OK;
}

}  // namespace b200cd
