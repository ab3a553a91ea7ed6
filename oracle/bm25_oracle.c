/* CPU oracle for the BM25 retrieval hot path, plain C -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
 * load this library.  The product (probing_rag_b200) never links or calls it.
 *
 * PARITY UNPINNED: restates the published algorithm of the third-party packages the
 * reference calls at /root/reference/exp_rag.py:242,426,428,492 (llama-index BM25Retriever
 * -> bm25s); those packages are absent here.
 *
 *   oracle_bm25_score      bm25s `_compute_relevance_from_scores` (App. A.5): dense f32
 *                          accumulator, postings added per query token in query order.
 *   oracle_bm25_topk       canonical order of SURVEY 8c: score desc, doc id asc, zero-score
 *                          tail filled with the lowest doc ids.
 *   oracle_bm25_retrieve   both, for a CSR batch of queries [q_lo, q_hi).
 *
 * Build: see oracle/Makefile (gcc -O2, no -ffast-math: every f32 add must round once).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int oracle_bm25_score(const int64_t *indptr, const int32_t *indices, const float *data,
                      int32_t n_terms, const int32_t *q_terms, int64_t n_q, float *scores)
{
    for (int64_t j = 0; j < n_q; ++j) {
        int32_t t = q_terms[j];
        if (t < 0 || t >= n_terms) return -1;
        for (int64_t p = indptr[t]; p < indptr[t + 1]; ++p) {
            scores[indices[p]] = scores[indices[p]] + data[p];   /* one rounded f32 add (SSE) */
        }
    }
    return 0;
}

static int beats(float s, int32_t d, float s2, int32_t d2)
{
    return s > s2 || (s == s2 && d < d2);
}

void oracle_bm25_topk(const float *scores, int32_t n_docs, int32_t k, int32_t doc_id_base,
                      float *out_s, int32_t *out_d)
{
    int32_t cnt = 0;
    for (int32_t d = 0; d < n_docs; ++d) {
        float s = scores[d];
        if (cnt == k && !beats(s, d, out_s[k - 1], out_d[k - 1])) continue;
        int32_t p = cnt < k ? cnt : k - 1;
        while (p > 0 && beats(s, d, out_s[p - 1], out_d[p - 1])) {
            out_s[p] = out_s[p - 1];
            out_d[p] = out_d[p - 1];
            --p;
        }
        out_s[p] = s;
        out_d[p] = d;
        if (cnt < k) ++cnt;
    }
    for (int32_t i = 0; i < cnt; ++i) out_d[i] += doc_id_base;
}

int oracle_bm25_retrieve(const int64_t *indptr, const int32_t *indices, const float *data,
                         int32_t n_docs, int32_t n_terms, int32_t doc_id_base,
                         const int64_t *q_indptr, const int32_t *q_terms,
                         int32_t q_lo, int32_t q_hi, int32_t k,
                         float *out_scores, int32_t *out_ids)
{
    if (k > n_docs) return -2;
    float *scores = (float *)malloc(sizeof(float) * (size_t)(n_docs > 0 ? n_docs : 1));
    if (!scores) return -3;
    int rc = 0;
    for (int32_t q = q_lo; q < q_hi && rc == 0; ++q) {
        memset(scores, 0, sizeof(float) * (size_t)n_docs);
        rc = oracle_bm25_score(indptr, indices, data, n_terms, q_terms + q_indptr[q],
                               q_indptr[q + 1] - q_indptr[q], scores);
        if (rc == 0)
            oracle_bm25_topk(scores, n_docs, k, doc_id_base, out_scores + (size_t)q * k,
                             out_ids + (size_t)q * k);
    }
    free(scores);
    return rc;
}
