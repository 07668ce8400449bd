// (C3) Host tail of the scorer: from the packed device results of a batch (AUSE slice sums, prologue sums, AUCE
// interval histogram) to the reference's curves and scalars.  HOST code: O(100) arithmetic per image, but as ~130
// small numpy calls it cost 0.2 ms + 17 us per image and bounded the batched scorer; here it is ~2 us per image.
//
// Reference arithmetic (file:line under /root/reference/nerfuncertainty), evaluated operation by operation in the
// reference's own dtypes so that the results equal the numpy expressions bit for bit
// (tests/test_host_logic.py::test_native_score_tail_equals_numpy_tail):
//   metrics/ause.py:15-20, 29-34   err_sorted[:c].mean() per ratio (float32; torch.sqrt for rmse)
//   metrics/ause.py:36-44          max(oracle curve, by-uncertainty curve) -- Python's max(): the first maximal element,
//                                  a NaN never wins but a leading NaN sticks; the oracle curve is a float32 array that
//                                  is divided in float32 unless the float64 by-uncertainty maximum is larger --,
//                                  normalisation, np.trapz of the gap over the 100 ratios
//   metrics/auce.py:24-54          coverage = count / n, interval length = 2 z mean(sigma), the three error curves and
//                                  their np.trapz areas over the 99 alphas
//   scripts/eval_uncertainty.py:323-333, 404-412   nll / avg_var / mse means as float32
// np.trapz(y, x) = (diff(x) * (y[1:] + y[:-1]) / 2.0).sum(): numpy adds a contiguous float64 row pairwise -- eight
// interleaved partial sums over blocks of eight, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), the remainder added
// one by one -- for rows of fewer than 128 elements; rows of fewer than 8 are added left to right.
#include <math.h>
#include <string.h>

#include "ub_common.cuh"

namespace ub {

static double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r[8];
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] += a[i + k];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}

// np.trapz over one row of m points with precomputed dx = diff(x) [m - 1]; tmp: m - 1 doubles
static double np_trapz(const double* y, const double* dx, int m, double* tmp) {
  for (int i = 0; i + 1 < m; ++i) tmp[i] = dx[i] * (y[i + 1] + y[i]) / 2.0;
  return np_pairwise_sum(tmp, m - 1);
}

// Python's max(iterable) over floats: the first element, replaced whenever a later one compares greater
template <typename T>
static T py_max(const T* a, int n) {
  T best = a[0];
  for (int i = 1; i < n; ++i)
    if (a[i] > best) best = a[i];
  return best;
}

}  // namespace ub

extern "C" {

int ub_score_tail_host(const double* packed, int32_t num_views, const int64_t* num_pixels, int32_t channels,
                       const int64_t* cuts, int32_t num_cuts, const double* ratio_steps, const double* z_values, int32_t num_z,
                       const double* one_minus_alpha, const double* alpha_steps, double* out_by_unc,
                       double* out_oracle64, float* out_oracle32, int32_t* out_oracle_is64, double* out_ause,
                       float* out_scalars, double* out_auce_curves, double* out_auc) {
  using namespace ub;
  UB_REQUIRE(packed && num_pixels && cuts && ratio_steps && z_values && one_minus_alpha && alpha_steps, UB_ERR_BAD_ARG,
             "score_tail: NULL input");
  UB_REQUIRE(out_by_unc && out_oracle64 && out_oracle32 && out_oracle_is64 && out_ause && out_scalars &&
                 out_auce_curves && out_auc,
             UB_ERR_BAD_ARG, "score_tail: NULL output");
  UB_REQUIRE(num_views >= 1 && num_cuts >= 2 && num_cuts <= 128 && num_z >= 2 && num_z <= 127 && channels >= 1,
             UB_ERR_BAD_ARG, "score_tail: bad sizes");
  const int B = num_views, NC = num_cuts, NZ = num_z;
  const double* sums = packed;                              // [B][4][NC]: ae by var, se by var, ae ascending, se ascending
  const double* psums = packed + (size_t)B * 4 * NC;        // [B][5]: se, ae, var, nll, sigma
  const double* hist_bits = psums + (size_t)B * 5;          // [B][NZ + 1] int64 bit patterns
  double tmp[128];
  for (int b = 0; b < B; ++b) {
    const int64_t n = num_pixels[b];                  // pixels of this view (ragged for the masked depth modality)
    const int64_t* vcuts = cuts + (size_t)b * NC;     // its slice lengths int((1 - r) n)
    const double* s4 = sums + (size_t)b * 4 * NC;
    // rows (mae, mse, rmse): oracle = (ae, se, se) ascending, by-uncertainty = (ae, se, se) by var
    const int by_row[3] = {0, 1, 1}, or_row[3] = {2, 3, 3};
    for (int e = 0; e < 3; ++e) {
      float ora[128], byu32[128];
      for (int k = 0; k < NC; ++k) {
        const int64_t c = vcuts[k];
        // float32 value of err_sorted[:c].mean(); an empty slice gives NaN like torch
        float o = c > 0 ? (float)(s4[or_row[e] * NC + k] / (double)c) : NAN;
        float u = c > 0 ? (float)(s4[by_row[e] * NC + k] / (double)c) : NAN;
        if (e == 2) {  // torch.sqrt of the float32 mean
          o = sqrtf(o);
          u = sqrtf(u);
        }
        ora[k] = o;
        byu32[k] = u;
      }
      double byu[128];
      for (int k = 0; k < NC; ++k) byu[k] = (double)byu32[k];
      const float a = py_max(ora, NC);
      const double bmax = py_max(byu, NC);
      const bool b_wins = bmax > (double)a;
      const double max64 = b_wins ? bmax : (double)a;
      double* o64 = out_oracle64 + ((size_t)b * 3 + e) * NC;
      float* o32 = out_oracle32 + ((size_t)b * 3 + e) * NC;
      double* bu = out_by_unc + ((size_t)b * 3 + e) * NC;
      double gap[128];
      for (int k = 0; k < NC; ++k) {
        o32[k] = ora[k] / a;                     // float32 / np.float32
        o64[k] = (double)ora[k] / max64;         // float32 / np.float64 -> float64
        bu[k] = byu[k] / max64;
        gap[k] = bu[k] - (b_wins ? o64[k] : (double)o32[k]);
      }
      out_oracle_is64[b * 3 + e] = b_wins ? 1 : 0;
      out_ause[b * 3 + e] = np_trapz(gap, ratio_steps, NC, tmp);
    }
    const double* ps = psums + (size_t)b * 5;
    out_scalars[b * 3 + 0] = (float)(ps[3] / (double)(n * channels));  // nll
    out_scalars[b * 3 + 1] = (float)(ps[2] / (double)n);               // avg_var
    out_scalars[b * 3 + 2] = (float)(ps[0] / (double)n);               // mse_mean
    // AUCE from the interval histogram: coverage_k = #{elements satisfying more than k thresholds} / (n c)
    int64_t hist[128];
    memcpy(hist, hist_bits + (size_t)b * (NZ + 1), sizeof(int64_t) * (NZ + 1));
    const double nel = (double)(n * channels);
    const double mean_sigma = (ps[4] * (double)channels) / nel;
    double* cov = out_auce_curves + (size_t)b * 5 * NZ;
    double* len = cov + NZ;
    double* err = len + NZ;
    double* abs_err = err + NZ;
    double* neg_err = abs_err + NZ;
    int64_t inside = 0;
    for (int k = NZ; k >= 1; --k) {  // reversed cumulative sum: inside_k = sum_{c > k-1} hist[c]
      inside += hist[k];
      cov[k - 1] = (double)inside / nel;
    }
    for (int k = 0; k < NZ; ++k) {
      len[k] = (2.0 * z_values[k]) * mean_sigma;
      err[k] = cov[k] - one_minus_alpha[k];
      abs_err[k] = fabs(err[k]);
      neg_err[k] = (abs_err[k] - err[k]) / 2.0;
    }
    out_auc[b * 3 + 0] = np_trapz(abs_err, alpha_steps, NZ, tmp);
    out_auc[b * 3 + 1] = np_trapz(len, alpha_steps, NZ, tmp);
    out_auc[b * 3 + 2] = np_trapz(neg_err, alpha_steps, NZ, tmp);
  }
  return UB_OK;
}

}  // extern "C"
