// hostprep.cu -- the host data path either side of the GPU kernels, as a compiled, R-free routine:
// what train_magma / train_magmaclust / pred_* do to the `data` tibble before any arithmetic
// (/root/reference/R/training.R:690-720, R/prediction.R:534-567,937-976, R/em-magma.R:27-30):
//   * inputs rounded with signif(x, 6)                       -- R nmath/fprec.c
//   * Reference = paste of the input columns with ":" through as.character(<double>): 15 significant digits, fixed
//     notation unless scientific is strictly shorter          -- R format.c (formatReal / scientific)
//   * rows ordered by ID string, then by Reference string, both byte-wise (dplyr::arrange, C locale)
//   * union grid = sorted unique Reference keys; grid_index = match(Reference, all_input) - 1
// The result is exactly the CSR layout mb_upload_dataset takes.  No CUDA call in this file.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/magma_b200.h"

void mb_set_error(const char* file, int line, const char* msg);
#define HP_FAIL(code, msg) do { mb_set_error(__FILE__, __LINE__, msg); return code; } while (0)

// signif(): R's fprec() with the max10e split of the power
extern "C" double mb_r_signif(double x, int digits) {
  if (!isfinite(x) || x == 0.0) return x;
  int dig = digits;
  if (dig > 22) return x;
  if (dig < 1) dig = 1;
  double sgn = 1.0;
  if (x < 0.0) { sgn = -1.0; x = -x; }
  const double l10 = log10(x);
  int e10 = (int)(dig - 1 - floor(l10));
  double p10 = 1.0;
  if (e10 > 308) { p10 = pow(10.0, e10 - 308); e10 = 308; }
  if (e10 > 0) {
    const double pow10 = pow(10.0, e10);
    return sgn * (nearbyint((x * pow10) * p10) / pow10) / p10;
  }
  const double pow10 = pow(10.0, -e10);
  return sgn * (nearbyint(x / pow10) * pow10);
}

static std::string r_as_character(double v) {
  if (isnan(v)) return "NA";
  if (isinf(v)) return v > 0 ? "Inf" : "-Inf";
  if (v == 0.0) return "0";
  const bool neg = v < 0;
  char buf[64];
  snprintf(buf, sizeof buf, "%.14e", v);
  const double ref = strtod(buf, nullptr);
  int nsig = 15;
  for (int n = 1; n <= 15; ++n) {
    snprintf(buf, sizeof buf, "%.*e", n - 1, v);
    if (strtod(buf, nullptr) == ref) { nsig = n; break; }
  }
  snprintf(buf, sizeof buf, "%.*e", nsig - 1, fabs(v));
  char* ep = strchr(buf, 'e');
  const int e = atoi(ep + 1);
  std::string mant(buf, ep - buf), digits;
  for (char ch : mant) if (ch != '.') digits.push_back(ch);
  const int wexp = abs(e) < 100 ? 2 : 3;
  const int w_sci = (neg ? 1 : 0) + (nsig > 1 ? nsig + 1 : 1) + 2 + wexp;
  const int left = e >= 0 ? e + 1 : 1;
  const int rgt = std::max(0, nsig - e - 1);
  const int w_fix = (neg ? 1 : 0) + left + (rgt > 0 ? rgt + 1 : 0);
  std::string s;
  if (w_fix <= w_sci) {
    if (e >= 0) {
      std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)e + 1));
      while ((int)ip.size() < e + 1) ip.push_back('0');
      s = ip;
      if (rgt > 0) s += "." + digits.substr(e + 1);
    } else {
      s = "0." + std::string(-e - 1, '0') + digits;
    }
  } else {
    char eb[16];
    snprintf(eb, sizeof eb, "e%c%0*d", e < 0 ? '-' : '+', wexp, abs(e));
    s = mant + eb;
  }
  return (neg ? "-" : "") + s;
}

extern "C" int mb_r_as_character(double x, char* out, int cap) {
  const std::string s = r_as_character(x);
  if (!out || cap <= (int)s.size()) HP_FAIL(-1, "buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

static std::string reference_key(const double* row, int d) {
  std::string k;
  for (int c = 0; c < d; ++c) { if (c) k.push_back(':'); k += r_as_character(row[c]); }
  return k;
}

extern "C" int mb_reference_key(const double* row, int d, int digits, char* out, int cap) {
  if (!row || d < 1 || !out) HP_FAIL(-1, "bad argument");
  std::vector<double> r(d);
  for (int c = 0; c < d; ++c) r[c] = mb_r_signif(row[c], digits);
  const std::string s = reference_key(r.data(), d);
  if (cap <= (int)s.size()) HP_FAIL(-1, "buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

extern "C" int mb_prepare_rows(int64_t n_rows, int d, const double* X, const double* y, const char* const* ids,
                               int digits, const double* grid_extra, int64_t n_extra, int64_t* perm, double* X_out,
                               double* y_out, int64_t* n_tasks_out, int64_t* offs_out, int32_t* grid_index_out,
                               int32_t* n_grid_out, double* grid_X_out) {
  if (n_rows < 1 || d < 1 || !X || !ids || !perm || !X_out || !n_tasks_out || !offs_out || !grid_index_out ||
      !n_grid_out || !grid_X_out || (n_extra > 0 && !grid_extra))
    HP_FAIL(-1, "bad argument");
  // signif + keys (one string per distinct value would be cheaper; rows are few next to the arithmetic they feed)
  std::vector<double> Xr((size_t)n_rows * d);
  std::vector<std::string> keys(n_rows);
  for (int64_t r = 0; r < n_rows; ++r) {
    for (int c = 0; c < d; ++c) Xr[r * d + c] = mb_r_signif(X[r * d + c], digits);
    keys[r] = reference_key(&Xr[r * d], d);
  }
  std::vector<int64_t> ord(n_rows);
  for (int64_t r = 0; r < n_rows; ++r) ord[r] = r;
  std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) {
    const int ci = strcmp(ids[a], ids[b]);            // strcmp compares unsigned bytes: C-locale order
    if (ci != 0) return ci < 0;
    return strcmp(keys[a].c_str(), keys[b].c_str()) < 0;
  });
  int64_t nt = 0;
  for (int64_t q = 0; q < n_rows; ++q) {
    const int64_t r = ord[q];
    perm[q] = r;
    for (int c = 0; c < d; ++c) X_out[q * d + c] = Xr[r * d + c];
    if (y && y_out) y_out[q] = y[r];
    if (q == 0 || strcmp(ids[r], ids[ord[q - 1]]) != 0) offs_out[nt++] = q;
    else if (keys[r] == keys[ord[q - 1]])
      HP_FAIL(3, "At least one individual have several Outputs on the same grid point.");   // training.R:703-714
  }
  offs_out[nt] = n_rows;
  *n_tasks_out = nt;
  // union grid: unique keys of the rows (first occurrence in the ORIGINAL row order gives the coordinates, they are
  // equal for equal keys up to signif) plus the extra grid rows, sorted byte-wise
  struct G { const std::string* k; const double* x; };
  std::vector<std::string> ekeys(n_extra);
  std::vector<double> Er((size_t)n_extra * d);
  std::vector<G> all;
  all.reserve(n_rows + n_extra);
  for (int64_t r = 0; r < n_rows; ++r) all.push_back({&keys[r], &Xr[r * d]});
  for (int64_t r = 0; r < n_extra; ++r) {
    for (int c = 0; c < d; ++c) Er[r * d + c] = mb_r_signif(grid_extra[r * d + c], digits);
    ekeys[r] = reference_key(&Er[r * d], d);
    all.push_back({&ekeys[r], &Er[r * d]});
  }
  std::stable_sort(all.begin(), all.end(), [](const G& a, const G& b) { return strcmp(a.k->c_str(), b.k->c_str()) < 0; });
  std::vector<const std::string*> gk;
  int64_t U = 0;
  for (size_t q = 0; q < all.size(); ++q) {
    if (q > 0 && *all[q].k == *all[q - 1].k) continue;
    for (int c = 0; c < d; ++c) grid_X_out[U * d + c] = all[q].x[c];
    gk.push_back(all[q].k);
    ++U;
  }
  if (U > INT32_MAX) HP_FAIL(-1, "grid too large");
  *n_grid_out = (int32_t)U;
  for (int64_t q = 0; q < n_rows; ++q) {
    const std::string& k = keys[ord[q]];
    auto it = std::lower_bound(gk.begin(), gk.end(), &k, [](const std::string* a, const std::string* b) {
      return strcmp(a->c_str(), b->c_str()) < 0;
    });
    grid_index_out[q] = (int32_t)(it - gk.begin());
  }
  return 0;
}
