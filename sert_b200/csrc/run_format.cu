// Host-side formatter of TREC run files (the step behind scoring): emits the lines
//   "<subject> Q0 <object> <rank> <relevance> <model>\n"
// of cvangysel trec_utils.write_run (trec_utils.py:573-580) for rankings that are already sorted arrays.  The one part
// that is not a memcpy is the relevance value: write_run prints '{0}'.format(value), i.e. Python's repr(float) --
// the shortest digit string that round-trips, fixed notation for decimal exponents in [-4, 16), exponent notation
// with at least two exponent digits otherwise (CPython: float_repr_style 'short', format code 'r').
// std::to_chars(double) yields the same shortest digits; only the layout rules are restated here.
// No CUDA in this file: it is part of libsert_b200.so so that one ctypes library serves the whole path.
#include <charconv>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace sert {

// writes repr(v) at dst (at least 32 bytes), returns the length
static int format_repr(double v, char *dst) {
  if (std::isnan(v)) { memcpy(dst, "nan", 3); return 3; }
  if (std::isinf(v)) {
    if (v < 0) { memcpy(dst, "-inf", 4); return 4; }
    memcpy(dst, "inf", 3);
    return 3;
  }
  char sci[40];
  const auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);
  const int n = (int)(res.ptr - sci);
  int pos = 0, out = 0;
  if (sci[0] == '-') { dst[out++] = '-'; pos = 1; }
  // mantissa digits (without the point) and the decimal exponent
  char digits[24];
  int nd = 0;
  while (pos < n && sci[pos] != 'e') {
    if (sci[pos] != '.') digits[nd++] = sci[pos];
    ++pos;
  }
  int exp10 = 0;
  if (pos < n) {
    ++pos;                                   // 'e'
    const bool neg = sci[pos] == '-';
    ++pos;                                   // sign (to_chars always writes one)
    while (pos < n) exp10 = exp10 * 10 + (sci[pos++] - '0');
    if (neg) exp10 = -exp10;
  }
  const int decpt = exp10 + 1;               // position of the decimal point relative to the first digit
  if (decpt > -4 && decpt <= 16) {
    if (decpt <= 0) {
      dst[out++] = '0'; dst[out++] = '.';
      for (int i = 0; i < -decpt; ++i) dst[out++] = '0';
      memcpy(dst + out, digits, nd); out += nd;
    } else if (decpt >= nd) {
      memcpy(dst + out, digits, nd); out += nd;
      for (int i = nd; i < decpt; ++i) dst[out++] = '0';
      dst[out++] = '.'; dst[out++] = '0';
    } else {
      memcpy(dst + out, digits, decpt); out += decpt;
      dst[out++] = '.';
      memcpy(dst + out, digits + decpt, nd - decpt); out += nd - decpt;
    }
    return out;
  }
  dst[out++] = digits[0];
  if (nd > 1) {
    dst[out++] = '.';
    memcpy(dst + out, digits + 1, nd - 1); out += nd - 1;
  }
  dst[out++] = 'e';
  int e = decpt - 1;
  dst[out++] = e < 0 ? '-' : '+';
  if (e < 0) e = -e;
  char tmp[8];
  int ne = 0;
  do { tmp[ne++] = (char)('0' + e % 10); e /= 10; } while (e > 0);
  if (ne < 2) tmp[ne++] = '0';
  while (ne > 0) dst[out++] = tmp[--ne];
  return out;
}

}  // namespace sert

extern "C" {

int64_t sert_format_run(const char *subject_blob, const int64_t *subject_off, const char *object_blob,
                        const int64_t *object_off, const int32_t *line_subject, const int32_t *line_object,
                        const int32_t *line_rank, const double *line_relevance, int64_t n_lines,
                        const char *model_name, char *out, int64_t capacity) {
  if (!subject_blob || !subject_off || !object_blob || !object_off || !line_subject || !line_object || !line_rank ||
      !line_relevance || !model_name || (!out && capacity > 0) || n_lines < 0) {
    sert::set_error("sert_format_run: null argument");
    return -1;
  }
  const size_t model_len = strlen(model_name);
  int64_t at = 0;
  for (int64_t i = 0; i < n_lines; ++i) {
    const int64_t s0 = subject_off[line_subject[i]], s1 = subject_off[line_subject[i] + 1];
    const int64_t o0 = object_off[line_object[i]], o1 = object_off[line_object[i] + 1];
    const int64_t need = (s1 - s0) + 4 + (o1 - o0) + 1 + 11 + 1 + 32 + 1 + (int64_t)model_len + 1;
    if (at + need > capacity) {
      sert::set_error("sert_format_run: output buffer too small");
      return -2;
    }
    memcpy(out + at, subject_blob + s0, (size_t)(s1 - s0)); at += s1 - s0;
    memcpy(out + at, " Q0 ", 4); at += 4;
    memcpy(out + at, object_blob + o0, (size_t)(o1 - o0)); at += o1 - o0;
    out[at++] = ' ';
    const auto r = std::to_chars(out + at, out + at + 11, line_rank[i]);
    at = r.ptr - out;
    out[at++] = ' ';
    at += sert::format_repr(line_relevance[i], out + at);
    out[at++] = ' ';
    memcpy(out + at, model_name, model_len); at += (int64_t)model_len;
    out[at++] = '\n';
  }
  return at;
}

}  // extern "C"
