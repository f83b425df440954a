// sig_scan.hpp -- purpose-built reader for the one field the train core consumes from a sourmash
// signature file: document[0]["signatures"][0]["mins"] (reference src/cpp/main.cpp:62-84, which
// builds a full nlohmann::json DOM to get at it).  This is a single forward pass over the bytes of
// the file: containers and strings that are not on that path are skipped without being
// materialised, and the hash list is converted digit by digit into uint64.
//
// Behaviour kept from the reference:
//   * only the FIRST record and its FIRST sub-signature are looked at, whatever their ksize;
//   * hashes may use the full uint64 range;
//   * a file that cannot be opened is reported on stderr and yields an empty sketch (:68-71).
// A document that does not have that path (not an array, no "signatures", no "mins", malformed
// JSON before the field) is a hard error here, as it is in the reference (uncaught json exception).
#pragma once
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace sigscan {

enum Status { OK = 0, CANNOT_OPEN = 1, MALFORMED = 2 };

static const uint64_t kPow10[9] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull};

struct Scanner {
    const char* p;
    const char* end;
    std::string why;

    bool fail(const char* msg) { why = msg; return false; }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }

    // p at opening quote; leaves p after the closing quote.  `out` may be null (skip only).
    bool string(std::string* out) {
        if (p >= end || *p != '"') return fail("expected string");
        p++;
        const char* s = p;
        bool esc = false;
        while (p < end) {
            const char c = *p;
            if (c == '\\') { esc = true; p += 2; continue; }
            if (c == '"') break;
            p++;
        }
        if (p >= end) return fail("unterminated string");
        if (out) {
            out->assign(s, p - s);
            (void)esc;  // keys on the path ("signatures", "mins") never contain escapes
        }
        p++;
        return true;
    }

    bool skip_value() {
        ws();
        if (p >= end) return fail("unexpected end of document");
        const char c = *p;
        if (c == '"') return string(nullptr);
        if (c == '{' || c == '[') {
            // skip a whole container: track depth, honour strings
            int depth = 0;
            while (p < end) {
                const char d = *p;
                if (d == '"') { if (!string(nullptr)) return false; continue; }
                if (d == '{' || d == '[') depth++;
                else if (d == '}' || d == ']') { depth--; if (depth == 0) { p++; return true; } }
                p++;
            }
            return fail("unterminated container");
        }
        // number / true / false / null
        while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\t' && *p != '\r') p++;
        return true;
    }

    // eight ASCII digits (first digit in the lowest byte) -> their value; SWAR, no per-digit loop
    // number of leading (lowest-address) bytes of c that are ASCII digits, 0..8
    static inline uint32_t leading_digits(uint64_t c) {
        const uint64_t t = c ^ 0x3030303030303030ull;                                         // digits -> 0..9
        const uint64_t nondigit = ((t + 0x7676767676767676ull) | t) & 0x8080808080808080ull;   // high bit where t >= 10
        return nondigit ? (uint32_t)(__builtin_ctzll(nondigit) >> 3) : 8u;
    }
    // the k (1..7) leading digits of c as an eight-digit field with leading '0's
    static inline uint64_t pad_digits(uint64_t c, uint32_t k) {
        return k == 8 ? c : ((c << (8u * (8u - k))) | (0x3030303030303030ull >> (8u * k)));
    }
    static inline uint64_t parse_eight(uint64_t c) {
        c -= 0x3030303030303030ull;
        c = (c * 10u) + (c >> 8);                                            // pairs of digits
        return (((c & 0x000000FF000000FFull) * (100u + (1000000ull << 32))) +
                (((c >> 16) & 0x000000FF000000FFull) * (1u + (10000ull << 32)))) >> 32;
    }

    // p at '[' of the mins array.  Hashes at scaled = 1000 are 13-17 digit literals: the digits are converted eight
    // at a time (same value modulo 2^64 as digit-by-digit accumulation); anything that is not a plain unsigned literal
    // takes the general route below.
    bool uint_array(std::vector<uint64_t>& out) {
        if (p >= end || *p != '[') return fail("\"mins\" is not an array");
        p++;
        for (;;) {
            if (p < end && *p == ',') p++;                    // the common separator, no whitespace
            else {
                ws();
                if (p >= end) return fail("unterminated \"mins\" array");
                if (*p == ']') { p++; return true; }
                if (*p == ',') { p++; continue; }
            }
            if (p >= end) return fail("unterminated \"mins\" array");
            const char* s = p;
            uint64_t v = 0;
            if (end - p >= 16) {
                uint64_t c;
                memcpy(&c, p, 8);
                uint32_t k = leading_digits(c);
                if (k == 8) {
                    v = parse_eight(c);
                    p += 8;
                    memcpy(&c, p, 8);
                    k = leading_digits(c);
                    if (k) { v = v * kPow10[k] + parse_eight(pad_digits(c, k)); p += k; }
                } else if (k) {
                    v = parse_eight(pad_digits(c, k));
                    p += k;
                }
            }
            while (p < end && (unsigned)(*p - '0') <= 9u) { v = v * 10u + (uint64_t)(*p - '0'); p++; }
            if (p == s || (p < end && (*p == '.' || *p == 'e' || *p == 'E'))) {
                if (p == s && (*p == ']' || *p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) continue;   // "[ ]", ", ]": the top of the loop decides
                // signed or floating literal: the reference casts whatever number it finds
                char* q = nullptr;
                errno = 0;
                if (*s == '-' || *s == '+') {
                    const long long sv = strtoll(s, &q, 10);
                    if (q == s) return fail("bad number in \"mins\"");
                    if (q < end && (*q == '.' || *q == 'e' || *q == 'E')) { const double dv = strtod(s, &q); v = (uint64_t)(long long)dv; }
                    else v = (uint64_t)sv;
                } else {
                    const double dv = strtod(s, &q);
                    if (q == s) return fail("bad number in \"mins\"");
                    v = (uint64_t)dv;
                }
                p = q;
            }
            out.push_back(v);
        }
    }

    // p at '{'; calls on_key for every member; on_key returns 1 = it consumed the value,
    // 0 = skip the value, -1 = error, 2 = consumed and stop scanning this object.
    template <typename F>
    bool object(F on_key, bool* stopped) {
        if (p >= end || *p != '{') return fail("expected object");
        p++;
        std::string key;
        for (;;) {
            ws();
            if (p >= end) return fail("unterminated object");
            if (*p == '}') { p++; return true; }
            if (*p == ',') { p++; continue; }
            if (!string(&key)) return false;
            ws();
            if (p >= end || *p != ':') return fail("expected ':'");
            p++;
            ws();
            const int r = on_key(key);
            if (r < 0) return false;
            if (r == 2) { *stopped = true; return true; }
            if (r == 0 && !skip_value()) return false;
        }
    }
};

// mins of document[0]["signatures"][0]; appended to `out`.
inline Status parse_mins(const char* data, size_t len, std::vector<uint64_t>& out, std::string* why) {
    Scanner sc{data, data + len, {}};
    auto bad = [&](const char* m) { if (why) *why = sc.why.empty() ? m : sc.why; return MALFORMED; };
    sc.ws();
    if (sc.p >= sc.end || *sc.p != '[') return bad("document is not an array");
    sc.p++;
    sc.ws();
    bool found_sigs = false, found_mins = false, stop = false;
    auto on_sub = [&](const std::string& k) -> int {
        if (k != "mins") return 0;
        if (!sc.uint_array(out)) return -1;
        found_mins = true;
        return 2;
    };
    auto on_rec = [&](const std::string& k) -> int {
        if (k != "signatures") return 0;
        found_sigs = true;
        if (sc.p >= sc.end || *sc.p != '[') { sc.fail("\"signatures\" is not an array"); return -1; }
        sc.p++;
        sc.ws();
        bool s2 = false;
        if (!sc.object(on_sub, &s2)) return -1;
        return 2;
    };
    if (!sc.object(on_rec, &stop)) return bad("malformed record");
    if (!found_sigs) return bad("no \"signatures\" in the first record");
    if (!found_mins) return bad("no \"mins\" in the first signature");
    return OK;
}

// Whole-file read into a reusable buffer, then parse_mins.
// The run side of the reference loads every sketch through load_signature_with_ksize (src/yacht/utils.py:31-51): ALL records and
// sub-signatures of the file are candidates, exactly one must carry the requested k-mer size.  Appends that one's mins to `out`;
// *n_match = how many sub-signatures of that ksize the file holds (the caller raises unless it is 1, like the reference).
inline Status parse_mins_ksize(const char* data, size_t len, int ksize, std::vector<uint64_t>& out, std::string* why, int* n_match) {
    Scanner sc{data, data + len, {}};
    auto bad = [&](const char* m) { if (why) *why = sc.why.empty() ? m : sc.why; return MALFORMED; };
    *n_match = 0;
    sc.ws();
    if (sc.p >= sc.end || *sc.p != '[') return bad("document is not an array");
    sc.p++;
    const char* mins_at = nullptr;      // where the "mins" array of the first matching sub-signature starts
    for (;;) {                          // records
        sc.ws();
        if (sc.p >= sc.end) return bad("unterminated document");
        if (*sc.p == ']') break;
        if (*sc.p == ',') { sc.p++; continue; }
        bool stop = false;
        auto on_rec = [&](const std::string& k) -> int {
            if (k != "signatures") return 0;
            if (sc.p >= sc.end || *sc.p != '[') { sc.fail("\"signatures\" is not an array"); return -1; }
            sc.p++;
            for (;;) {                  // sub-signatures
                sc.ws();
                if (sc.p >= sc.end) { sc.fail("unterminated \"signatures\""); return -1; }
                if (*sc.p == ']') { sc.p++; return 1; }
                if (*sc.p == ',') { sc.p++; continue; }
                long long this_k = -1;
                const char* this_mins = nullptr;
                bool s2 = false;
                auto on_sub = [&](const std::string& kk) -> int {
                    if (kk == "ksize") {
                        char* q = nullptr;
                        this_k = strtoll(sc.p, &q, 10);
                        if (q == sc.p) { sc.fail("bad \"ksize\""); return -1; }
                        sc.p = q;
                        return 1;
                    }
                    if (kk == "mins") { this_mins = sc.p; return 0; }      // remembered, skipped for now
                    return 0;
                };
                if (!sc.object(on_sub, &s2)) return -1;
                if (this_k == (long long)ksize) {
                    if (*n_match == 0) mins_at = this_mins;
                    (*n_match)++;
                }
            }
        };
        if (!sc.object(on_rec, &stop)) return bad("malformed record");
    }
    if (*n_match >= 1 && mins_at) {
        Scanner m{mins_at, data + len, {}};
        if (!m.uint_array(out)) { if (why) *why = m.why; return MALFORMED; }
    }
    return OK;
}

inline Status read_file(const std::string& path, std::vector<char>& buf, size_t* got_out) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return CANNOT_OPEN;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return CANNOT_OPEN; }
    const size_t len = (size_t)st.st_size;
    if (buf.size() < len + 1) buf.resize(len + 1);
    size_t got = 0;
    while (got < len) {
        const ssize_t r = read(fd, buf.data() + got, len - got);
        if (r < 0) { if (errno == EINTR) continue; close(fd); return CANNOT_OPEN; }
        if (r == 0) break;
        got += (size_t)r;
    }
    close(fd);
    buf[got] = 0;
    *got_out = got;
    return OK;
}

inline Status read_mins_ksize(const std::string& path, int ksize, std::vector<char>& buf, std::vector<uint64_t>& out, std::string* why, int* n_match) {
    size_t got = 0;
    *n_match = 0;
    const Status st = read_file(path, buf, &got);
    if (st != OK) return st;
    return parse_mins_ksize(buf.data(), got, ksize, out, why, n_match);
}

inline Status read_mins(const std::string& path, std::vector<char>& buf, std::vector<uint64_t>& out, std::string* why) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return CANNOT_OPEN;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return CANNOT_OPEN; }
    const size_t len = (size_t)st.st_size;
    if (buf.size() < len + 1) buf.resize(len + 1);
    size_t got = 0;
    while (got < len) {
        const ssize_t r = read(fd, buf.data() + got, len - got);
        if (r < 0) { if (errno == EINTR) continue; close(fd); return CANNOT_OPEN; }
        if (r == 0) break;
        got += (size_t)r;
    }
    close(fd);
    buf[got] = 0;
    return parse_mins(buf.data(), got, out, why);
}

}  // namespace sigscan
