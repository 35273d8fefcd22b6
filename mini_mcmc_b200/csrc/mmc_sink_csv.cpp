// CSV sink, host half (src/io/csv.rs:47-77,110-147): header `chain,observation,dim_0..`, then one record per
// (chain, observation) in chain-major order, every value printed the way Rust's `Display` prints it: integers as
// decimal digits, floats as the SHORTEST decimal string that round-trips in the value's own type, never in
// exponent notation (`42`, `1.1`, `0.0000001`), `NaN`, `inf`, `-inf`.  std::to_chars(fixed) has exactly these
// semantics.  Chains are formatted by the host threads in parallel and written in order.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/minimcmc.h"

namespace mmc {
void set_error(const char *fmt, ...);
int widen_threads();

namespace {

inline void put_u64(std::string &s, uint64_t v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    s.append(buf, r.ptr);
}

// Rust's float Display: the shortest digit string that round-trips (flt2dec "shortest"), laid out positionally and
// padded with zeros (f32::MAX prints as 340282350000000000000000000000000000000).  std::to_chars(scientific) yields
// the same shortest digits; std::to_chars(fixed) would print the exact integer digits instead of the zero padding.
template <typename F>
inline void put_float(std::string &s, F v) {
    if (std::isnan(v)) { s += "NaN"; return; }
    if (std::isinf(v)) { s += v < 0 ? "-inf" : "inf"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    const char *p = buf;
    if (*p == '-') { s += '-'; ++p; }
    char digits[32];
    int nd = 0;
    for (; p < r.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    int exp10 = 0;
    if (p < r.ptr) std::from_chars(p + 1 + (p[1] == '+' ? 1 : 0), r.ptr, exp10);
    while (nd > 1 && digits[nd - 1] == '0') --nd;   // "1.0e0" style never occurs, but keep the invariant explicit
    if (exp10 >= 0) {
        for (int i = 0; i <= exp10; ++i) s += i < nd ? digits[i] : '0';
        if (nd > exp10 + 1) {
            s += '.';
            s.append(digits + exp10 + 1, nd - exp10 - 1);
        }
    } else {
        s += "0.";
        s.append((size_t)(-exp10 - 1), '0');
        s.append(digits, nd);
    }
}

template <typename T>
void format_rows(const T *data, int64_t n, int32_t dim, int64_t c_begin, int64_t c_end, std::string &out) {
    for (int64_t c = c_begin; c < c_end; ++c) {
        for (int64_t t = 0; t < n; ++t) {
            put_u64(out, (uint64_t)c);
            out += ',';
            put_u64(out, (uint64_t)t);
            const T *row = data + (c * n + t) * dim;
            for (int32_t d = 0; d < dim; ++d) {
                out += ',';
                if constexpr (std::is_floating_point<T>::value) put_float(out, row[d]);
                else put_u64(out, (uint64_t)row[d]);
            }
            out += '\n';
        }
    }
}

template <typename T>
int write_csv(const T *data, int64_t chains, int64_t n, int32_t dim, FILE *f) {
    const int nt = std::max(1, widen_threads());
    // blocks of chains sized to ~32 MB of text per round so memory stays bounded
    const int64_t rows_per_round = std::max<int64_t>(1, (int64_t)(32 << 20) / std::max<int64_t>(1, 12 * (dim + 2)));
    const int64_t chains_per_round = std::max<int64_t>(nt, rows_per_round / std::max<int64_t>(1, n));
    std::vector<std::string> parts((size_t)nt);
    for (int64_t c0 = 0; c0 < chains; c0 += chains_per_round) {
        const int64_t c1 = std::min(chains, c0 + chains_per_round);
        const int64_t per = (c1 - c0 + nt - 1) / nt;
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const int64_t b = c0 + t * per, e = std::min(c1, b + per);
            parts[t].clear();
            if (b >= e) continue;
            if (nt == 1) format_rows(data, n, dim, b, e, parts[t]);
            else th.emplace_back([=, &parts]() { format_rows(data, n, dim, b, e, parts[t]); });
        }
        for (auto &x : th) x.join();
        for (int t = 0; t < nt; ++t)
            if (!parts[t].empty() && fwrite(parts[t].data(), 1, parts[t].size(), f) != parts[t].size()) return -1;
    }
    return 0;
}

}  // namespace
}  // namespace mmc

extern "C" int mmc_save_csv(const void *sample_host, int32_t dtype, int64_t chains, int64_t n, int32_t dim, const char *filename) {
    using namespace mmc;
    if (!filename || chains < 0 || n < 0 || dim < 0 || (!sample_host && chains * n * dim > 0)) {
        set_error("mmc_save_csv: bad arguments");
        return MMC_ERR_INVALID;
    }
    FILE *f = fopen(filename, "wb");
    if (!f) {
        set_error("mmc_save_csv: cannot create %s: %s", filename, strerror(errno));
        return MMC_ERR_INVALID;
    }
    std::string header = "chain,observation";
    for (int32_t d = 0; d < dim; ++d) header += ",dim_" + std::to_string(d);
    header += '\n';
    int rc = fwrite(header.data(), 1, header.size(), f) == header.size() ? 0 : -1;
    if (rc == 0 && chains * n > 0) {
        if (dtype == MMC_F32) rc = write_csv(static_cast<const float *>(sample_host), chains, n, dim, f);
        else if (dtype == MMC_F64) rc = write_csv(static_cast<const double *>(sample_host), chains, n, dim, f);
        else if (dtype == MMC_U64) rc = write_csv(static_cast<const uint64_t *>(sample_host), chains, n, dim, f);
        else { fclose(f); set_error("mmc_save_csv: unknown dtype %d", dtype); return MMC_ERR_INVALID; }
    }
    if (fclose(f) != 0) rc = -1;
    if (rc) {
        set_error("mmc_save_csv: write to %s failed: %s", filename, strerror(errno));
        return MMC_ERR_INVALID;
    }
    return MMC_OK;
}
