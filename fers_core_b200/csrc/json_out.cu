// json_out.cu -- host-side writer for the bulk of a ResultsBundle document (the
// reference boundary returns JSON text, fers_core/fers/fers.py:300-311; at a few thousand members the text is tens
// of MB and Python's json.dumps spends seconds on float formatting).  A map of equally-shaped records
//   {"<key 0>": <template with numbers>, "<key 1>": ..., ...}
// is written from a template string (byte 0x01 marks a number slot, 0x02 an integer slot) and a dense
// [n_records][n_slots] array:
// shortest round-trip decimal form (std::to_chars), NaN / Infinity spelled as Python's json module spells them.
// No CUDA in here; it lives in the library so that the drop-in module needs nothing but libfers_b200.so.
#include <charconv>
#include <cmath>
#include <cstring>
#include <string>

#include "common.cuh"

namespace {

inline char* put_number(char* p, double v) {
    if (std::isnan(v)) { memcpy(p, "NaN", 3); return p + 3; }
    if (std::isinf(v)) {
        if (v < 0) { memcpy(p, "-Infinity", 9); return p + 9; }
        memcpy(p, "Infinity", 8); return p + 8;
    }
    auto r = std::to_chars(p, p + 32, v);
    // keep the value a JSON *float* where Python would print one ("5" -> "5.0"): the text then parses to the same types
    bool plain_int = true;
    for (char* q = p; q < r.ptr; ++q)
        if (*q == '.' || *q == 'e' || *q == 'E' || *q == 'n' || *q == 'i') { plain_int = false; break; }
    if (plain_int) { r.ptr[0] = '.'; r.ptr[1] = '0'; return r.ptr + 2; }
    return r.ptr;
}

}  // namespace

extern "C" {

/* Upper bound of the bytes fers_b200_json_records writes. */
size_t fers_b200_json_records_bound(size_t template_len, int n_slots, long long n_records, size_t keys_bytes) {
    return 2 + (size_t)n_records * (template_len + (size_t)n_slots * 32 + 4) + keys_bytes + 16;
}

/* keys: n_records NUL-terminated strings back to back (already JSON-safe: decimal ids); values [n_records][n_slots];
 * out receives '{' "key":record , ... '}' (no terminating NUL); *written_out = bytes written. */
int fers_b200_json_records(const char* tmpl, size_t template_len, int n_slots, long long n_records, const char* keys,
                           const double* values, char* out, size_t out_cap, size_t* written_out) {
    FERS_REQUIRE(tmpl && out && written_out && (n_records == 0 || (keys && (values || n_slots == 0))), "json_records: bad arguments");
    int slots = 0;
    for (size_t i = 0; i < template_len; ++i) slots += (tmpl[i] == '\x01' || tmpl[i] == '\x02');
    FERS_REQUIRE(slots == n_slots, "json_records: template slot count does not match n_slots");
    size_t keys_bytes = 0;
    {
        const char* k = keys;
        for (long long r = 0; r < n_records; ++r) { size_t l = strlen(k); keys_bytes += l + 1; k += l + 1; }
    }
    FERS_REQUIRE(out_cap >= fers_b200_json_records_bound(template_len, n_slots, n_records, keys_bytes), "json_records: output too small");
    char* p = out;
    *p++ = '{';
    const char* k = keys;
    for (long long r = 0; r < n_records; ++r) {
        if (r) *p++ = ',';
        *p++ = '"';
        const size_t kl = strlen(k);
        memcpy(p, k, kl); p += kl; k += kl + 1;
        *p++ = '"'; *p++ = ':';
        const double* v = values + (size_t)r * n_slots;
        const char* t = tmpl;
        const char* te = tmpl + template_len;
        while (t < te) {
            const char* m = t;
            while (m < te && *m != '\x01' && *m != '\x02') ++m;
            memcpy(p, t, (size_t)(m - t)); p += m - t;
            if (m == te) break;
            if (*m == '\x01') p = put_number(p, *v++);
            else p = std::to_chars(p, p + 24, (long long)std::llround(*v++)).ptr;      // integer slot
            t = m + 1;
        }
    }
    *p++ = '}';
    *written_out = (size_t)(p - out);
    return 0;
}

}  // extern "C"
