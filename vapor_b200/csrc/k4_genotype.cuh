// Kernel 4: per-SV summary and genotype.
//
// Restates result_organize_ins (vapor_vali/Simple_function.pyx:1219-1231) and
// gt_estimate_log_likelihood / log_likelihood_calcu (:2054-2077).  One thread per SV: the
// reference's sums are order-dependent, so the sequential order is kept, including numpy's
// pairwise summation inside np.mean (:1226).
#pragma once
#include "common.cuh"

namespace vb {

// numpy's pairwise_sum for contiguous float64 (numpy/_core/src/umath/loops_utils.h.src)
__device__ double k4_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return k4_pairwise_sum(a, n2) + k4_pairwise_sum(a + n2, n - n2);
}

// log_likelihood_calcu(k, l, m=2, g, err=0.05), Simple_function.pyx:2071-2077
__device__ double k4_loglik(int k, int l, int g) {
    const double m = 2.0, err = 0.05;
    double out = (double)(-k) * log(m);
    const double t1 = log((m - g) * err + g * (1.0 - err));
    const double t2 = log((m - g) * (1.0 - err) + g * err);
    for (int j = 0; j < l; ++j) out += t1;
    for (int j = 0; j < k - l; ++j) out += t2;
    return out;
}

__global__ void k4_genotype(const int64_t* __restrict__ sv_off, int n_sv,
                            const double* __restrict__ task_score, const uint8_t* __restrict__ task_status,
                            double* __restrict__ pos_scratch,        // [n_task]
                            double* sv_qs, double* sv_gs, double* sv_gq, uint8_t* sv_gt, int32_t* sv_nscore)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sv) return;
    const int64_t t0 = sv_off[s], t1 = sv_off[s + 1];
    double* pos = pos_scratch + t0;
    int k = 0, npos = 0, l = 0;
    for (int64_t t = t0; t < t1; ++t) {
        if (task_status[t] != 1) continue;
        const double sc = task_score[t];
        ++k;
        if (sc > 0) pos[npos++] = sc;                    // float(i) > 0, Simple_function.pyx:1222
        // gt_estimate re-parses Rec = str(round(score, 2)) (:2055): "not i > 0" there means the
        // score rounds to <= 0.00 at 2 decimals, i.e. score < 0.005 (the double literal).
        if (!(sc >= 0.005)) ++l;
    }
    sv_nscore[s] = k;
    if (k == 0) { sv_qs[s] = 0; sv_gs[s] = 0; sv_gq[s] = 0; sv_gt[s] = 255; return; }   // 'NA' row
    const double qs = npos ? k4_pairwise_sum(pos, npos) / (double)npos : 0.0;
    const double gs = (double)npos / (double)k;
    const double ll[3] = { k4_loglik(k, l, 2), k4_loglik(k, l, 1), k4_loglik(k, l, 0) };   // 0/0, 0/1, 1/1
    double mx = ll[0]; int arg = 0;
    for (int i = 1; i < 3; ++i) if (ll[i] > mx) { mx = ll[i]; arg = i; }
    double ori[3], sum = 0.0;
    for (int i = 0; i < 3; ++i) { ori[i] = exp(ll[i] - mx); sum += ori[i]; }
    double nrm[3];
    for (int i = 0; i < 3; ++i) nrm[i] = ori[i] / sum;
    // np.median of three values = the middle one
    double median = nrm[0];
    if ((nrm[0] <= nrm[1] && nrm[1] <= nrm[2]) || (nrm[2] <= nrm[1] && nrm[1] <= nrm[0])) median = nrm[1];
    else if ((nrm[0] <= nrm[2] && nrm[2] <= nrm[1]) || (nrm[1] <= nrm[2] && nrm[2] <= nrm[0])) median = nrm[2];
    const double gq = -log(median) / log(10.0);
    if (arg == 0 && gs > 0.15) arg = 1;                  // Simple_function.pyx:2068
    sv_qs[s] = qs; sv_gs[s] = gs; sv_gq[s] = gq; sv_gt[s] = (uint8_t)arg;
}

}  // namespace vb
