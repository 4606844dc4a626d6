// On-device binary classification metrics of the reference's evaluation loop (eval.py:79-94 calculate_eer, eval.py:174-192:
// accuracy / AUC / EER / precision / recall / F1 / average precision / confusion matrix, all scikit-learn calls on host lists
// gathered with a .cpu() per batch).  Here the per-video scores stay on the GPU (one NCCL all_gather brings every rank's
// scores together) and ONE single-CTA kernel reduces them to the 12 numbers, following scikit-learn's definitions step by step:
//   _binary_clf_curve : sort by score (descending), one point per DISTINCT score: tps = cumsum(y)[idx], fps = 1 + idx - tps
//   roc_curve         : drop_intermediate=True removes points where the second differences of fps AND tps vanish (first and
//                       last kept); a (0, 0) point with threshold +inf is prepended; fpr = fps / fps[-1], tpr = tps / tps[-1]
//   roc_auc_score     : trapezoid rule over (fpr, tpr)
//   calculate_eer     : idx = nanargmin |(1 - tpr) - fpr| (first minimum), eer = fpr[idx], threshold = thresholds[idx]
//   average_precision : sum over distinct thresholds (descending) of (R_k - R_{k-1}) * P_k, P = tps / (tps + fps), R = tps / tps[-1]
//   threshold 0.5     : pred = score >= 0.5 -> confusion matrix, accuracy, precision, recall, F1 (0 when the denominator is 0)
// Work is tiny (n = number of videos, <= 8192): a bitonic sort in shared memory by 1024 threads, then one thread walks the
// sorted list in double precision (the arithmetic type numpy uses there).  No atomics, fixed order: bit-reproducible.
#include "ewvit_common.cuh"

namespace {

constexpr int kMaxN = 8192;

__global__ void __launch_bounds__(1024) binary_metrics_kernel(const float *__restrict__ scores, const int *__restrict__ labels, int n,
                                                              float *__restrict__ out) {
    __shared__ float s_key[kMaxN];
    __shared__ unsigned char s_lab[kMaxN];
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        s_key[i] = i < n ? scores[i] : -INFINITY;      // padding sorts to the end (descending order)
        s_lab[i] = i < n ? (labels[i] != 0) : 0;
    }
    __syncthreads();
    // bitonic sort, descending by score
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool desc = (i & k) == 0;
                    const float a = s_key[i], b = s_key[p];
                    if (desc ? (a < b) : (a > b)) {
                        s_key[i] = b; s_key[p] = a;
                        const unsigned char t = s_lab[i]; s_lab[i] = s_lab[p]; s_lab[p] = t;
                    }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x != 0) return;

    // ---- threshold 0.5 metrics
    long long tp = 0, fp = 0, tn = 0, fn = 0;
    for (int i = 0; i < n; ++i) {
        const bool pred = s_key[i] >= 0.5f, y = s_lab[i] != 0;
        tp += pred && y; fp += pred && !y; fn += !pred && y; tn += !pred && !y;
    }
    const double P = (double)(tp + fn), N = (double)(fp + tn);
    out[3] = (float)((double)(tp + tn) / (double)n);
    out[4] = tp + fp > 0 ? (float)((double)tp / (double)(tp + fp)) : 0.f;
    out[5] = tp + fn > 0 ? (float)((double)tp / (double)(tp + fn)) : 0.f;
    out[6] = 2 * tp + fp + fn > 0 ? (float)(2.0 * (double)tp / (double)(2 * tp + fp + fn)) : 0.f;
    out[8] = (float)tn; out[9] = (float)fp; out[10] = (float)fn; out[11] = (float)tp;

    // ---- curves over distinct thresholds (descending).  Single-class inputs have no ROC curve / AP: NaN, as an error marker.
    if (P == 0.0 || N == 0.0) {
        const float nanv = __int_as_float(0x7fc00000);
        out[0] = out[1] = out[2] = out[7] = nanv;
        return;
    }
    // pass 1: AP and the trapezoid AUC (dropping collinear points does not change the area)
    double auc = 0.0, ap = 0.0, prev_fpr = 0.0, prev_tpr = 0.0, prev_rec = 0.0;
    long long ctp = 0;
    for (int i = 0; i < n; ++i) {
        ctp += s_lab[i];
        if (i == n - 1 || s_key[i] != s_key[i + 1]) {
            const double tps = (double)ctp, fps = (double)(1 + i) - tps;
            const double fpr = fps / N, tpr = tps / P;
            auc += (fpr - prev_fpr) * (tpr + prev_tpr) * 0.5;
            ap += (tpr - prev_rec) * (tps / (tps + fps));
            prev_fpr = fpr; prev_tpr = tpr; prev_rec = tpr;
        }
    }
    out[0] = (float)auc;
    out[7] = (float)ap;
    // pass 2: EER on the roc_curve points that survive drop_intermediate.  Point k is kept iff it is the first or last distinct
    // threshold or (fps[k-1] - 2 fps[k] + fps[k+1]) != 0 or the same for tps; the prepended (0, 0) point has threshold +inf.
    double best = fabs(1.0 - 0.0 - 0.0);            // the (fpr, tpr) = (0, 0) point: |fnr - fpr| = 1
    double eer = 0.0;
    float thr = INFINITY;
    long long ndist = 0;
    for (int i = 0; i < n; ++i) ndist += (i == n - 1 || s_key[i] != s_key[i + 1]);
    double f_prev = 0.0, t_prev = 0.0, f_cur = 0.0, t_cur = 0.0;      // (fps, tps) of the previous / current distinct point
    float thr_cur = 0.f;
    long long k = -1;                                // index of the current distinct point
    ctp = 0;
    for (int i = 0; i <= n; ++i) {
        bool emit = false;
        double f_next = 0.0, t_next = 0.0;
        float thr_next = 0.f;
        if (i < n) {
            ctp += s_lab[i];
            if (i == n - 1 || s_key[i] != s_key[i + 1]) {
                t_next = (double)ctp; f_next = (double)(1 + i) - t_next; thr_next = s_key[i];
                emit = true;
            }
        } else {
            emit = true;                             // flush the last point
        }
        if (!emit) continue;
        if (k >= 0) {                                // decide about the CURRENT point now that its successor is known
            bool keep = (k == 0) || (i == n) || ndist <= 2;
            if (!keep) keep = (f_prev - 2.0 * f_cur + f_next) != 0.0 || (t_prev - 2.0 * t_cur + t_next) != 0.0;
            if (keep) {
                const double fpr = f_cur / N, fnr = 1.0 - t_cur / P;
                const double d = fabs(fnr - fpr);
                if (d < best) { best = d; eer = fpr; thr = thr_cur; }
            }
        }
        if (i < n) {
            f_prev = f_cur; t_prev = t_cur;
            f_cur = f_next; t_cur = t_next; thr_cur = thr_next;
            ++k;
        }
    }
    out[1] = (float)eer;
    out[2] = thr;
}

}  // namespace

extern "C" int ewvit_binary_metrics_fwd(const float *scores, const int *labels, int n, float *out, void *stream) {
    EWVIT_REQUIRE(n > 0 && n <= kMaxN, EWVIT_ERR_UNSUPPORTED, "ewvit_binary_metrics_fwd: needs 1 <= n <= %d samples (got %d)", kMaxN, n);
    EWVIT_REQUIRE(scores && labels && out, EWVIT_ERR_INVALID_ARG, "ewvit_binary_metrics_fwd: NULL pointer");
    int rc = ewvit_check_device();
    if (rc != EWVIT_OK) return rc;
    binary_metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scores, labels, n, out);
    EWVIT_LAUNCH_OK();
    return EWVIT_OK;
}
