/*
 * blp_oracle.c -- plain-C CPU restatement of the reference's similarity.py
 * (TEST INFRASTRUCTURE ONLY; pinned to the reference's own code through Oracle A and the fixtures
 * that code wrote, see similarity_oracle.py / ref_runner.py).  A third, independent statement of the same algorithm,
 * fast enough to check the CUDA path at the full BASELINE.json sizes.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product never does.
 *
 * What follows what (file:line relative to /root/reference):
 *   graph build (sort, unique)        <- snap.LoadEdgeList(PUNGraph, f, 0, 1)     similarity.py:16
 *   "in graph" = degree >= 1          <- node list membership                     similarity.py:22,52,59-60
 *   mark_hop2()                       <- GetNodesAtHop(G, x, 2, ...) into a Set   similarity.py:24-33,67-78
 *   walk of N(y)                      <- GetNodesAtHop(G, y, 1, ...)              similarity.py:36-45,81-89
 *   cn / union / jaccard              <- common_neighbors, jaccard                similarity.py:108-114
 *   aa                                <- adamic_adar (1/ln(deg), deg>1 only)      similarity.py:116-126
 *   pa                                <- degree product                           "Link prediction.R":400-415
 *
 * Indices are local: users 0..n_users-1, businesses 0..n_biz-1; a negative / out-of-range index
 * or a node of degree 0 means "id not in graph" and zeroes every output of the pair.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_rows;
    int64_t* off; /* n_rows + 1 */
    int* adj;     /* ascending inside each row, no duplicates */
} csr_t;

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static int build(int n_users, int n_biz, int64_t m, const int* eu, const int* eb, csr_t* ub,
                 csr_t* bu) {
    uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(m ? m : 1));
    if (!key) return -1;
    for (int64_t i = 0; i < m; ++i) {
        if (eu[i] < 0 || eu[i] >= n_users || eb[i] < 0 || eb[i] >= n_biz) {
            free(key);
            return -2;
        }
        key[i] = ((uint64_t)(uint32_t)eu[i] << 32) | (uint32_t)eb[i];
    }
    qsort(key, (size_t)m, sizeof(uint64_t), cmp_u64);
    int64_t k = 0;
    for (int64_t i = 0; i < m; ++i)      /* duplicate review lines collapse (TUNGraph) */
        if (i == 0 || key[i] != key[i - 1]) key[k++] = key[i];
    ub->n_rows = n_users;
    bu->n_rows = n_biz;
    ub->off = (int64_t*)calloc((size_t)n_users + 1, sizeof(int64_t));
    bu->off = (int64_t*)calloc((size_t)n_biz + 1, sizeof(int64_t));
    ub->adj = (int*)malloc(sizeof(int) * (size_t)(k ? k : 1));
    bu->adj = (int*)malloc(sizeof(int) * (size_t)(k ? k : 1));
    if (!ub->off || !bu->off || !ub->adj || !bu->adj) return -1;
    for (int64_t i = 0; i < k; ++i) {
        ub->off[(key[i] >> 32) + 1]++;
        bu->off[(uint32_t)key[i] + 1]++;
    }
    for (int i = 0; i < n_users; ++i) ub->off[i + 1] += ub->off[i];
    for (int i = 0; i < n_biz; ++i) bu->off[i + 1] += bu->off[i];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_biz);
    if (!cur) return -1;
    memcpy(cur, bu->off, sizeof(int64_t) * (size_t)n_biz);
    for (int64_t i = 0; i < k; ++i) {    /* keys ascend by (u,b): both directions come out sorted */
        int u = (int)(key[i] >> 32), b = (int)(uint32_t)key[i];
        ub->adj[i] = b;
        bu->adj[cur[b]++] = u;
    }
    free(cur);
    free(key);
    return 0;
}

static int deg(const csr_t* g, int x) { return (int)(g->off[x + 1] - g->off[x]); }

/* stamp[] marks hop2(x): nodes of x's own side that share >= 1 neighbour with x, x excluded.
 * Returns |hop2(x)|. */
static int mark_hop2(const csr_t* gx, const csr_t* gm, int x, int* stamp, int tag) {
    int count = 0;
    for (int64_t i = gx->off[x]; i < gx->off[x + 1]; ++i) {
        int m = gx->adj[i];
        for (int64_t j = gm->off[m]; j < gm->off[m + 1]; ++j) {
            int y = gm->adj[j];
            if (y != x && stamp[y] != tag) {
                stamp[y] = tag;
                ++count;
            }
        }
    }
    return count;
}

typedef struct {
    int key;
    int64_t idx;
} order_t;

static int cmp_order(const void* a, const void* b) {
    const order_t* x = (const order_t*)a;
    const order_t* y = (const order_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* One side.  gx: rows of the grouping side (x -> middle nodes); gm: rows of the middle side. */
static int side(const csr_t* gx, const csr_t* gm, int64_t n, const int* px, const int* py,
                int* cn, int* uni, double* jac, double* aa) {
    int* stamp = (int*)calloc((size_t)gx->n_rows, sizeof(int));
    order_t* ord = (order_t*)malloc(sizeof(order_t) * (size_t)(n ? n : 1));
    double* w = (double*)malloc(sizeof(double) * (size_t)gx->n_rows);
    if (!stamp || !ord || !w) return -1;
    for (int i = 0; i < gx->n_rows; ++i) {
        int d = deg(gx, i);
        w[i] = d > 1 ? 1.0 / log((double)d) : 0.0;     /* similarity.py:121-125 */
    }
    for (int64_t i = 0; i < n; ++i) {
        ord[i].key = px[i];
        ord[i].idx = i;
    }
    qsort(ord, (size_t)n, sizeof(order_t), cmp_order);  /* one hop-2 set per distinct node */
    int tag = 0, cur = -1, hop2 = 0;
    for (int64_t t = 0; t < n; ++t) {
        int64_t i = ord[t].idx;
        int x = px[i], y = py[i];
        cn[i] = 0;
        uni[i] = 0;
        jac[i] = 0.0;
        aa[i] = 0.0;
        if (x < 0 || x >= gx->n_rows || y < 0 || y >= gm->n_rows) continue;
        if (deg(gx, x) == 0 || deg(gm, y) == 0) continue; /* id not in graph -> 0 (:59-60) */
        if (x != cur) {
            cur = x;
            hop2 = mark_hop2(gx, gm, x, stamp, ++tag);
        }
        int c = 0;
        double s = 0.0;
        for (int64_t j = gm->off[y]; j < gm->off[y + 1]; ++j) {
            int z = gm->adj[j];
            if (stamp[z] == tag) {
                ++c;
                s += w[z];
            }
        }
        cn[i] = c;
        uni[i] = hop2 + deg(gm, y) - c;                  /* |a| + |b| - |a & b| */
        jac[i] = (double)c / (double)uni[i];
        aa[i] = s;
    }
    free(stamp);
    free(ord);
    free(w);
    return 0;
}

int blp_oracle_score(int n_users, int n_biz, int64_t n_edges, const int* edge_u,
                     const int* edge_b, int64_t n_pairs, const int* pair_u, const int* pair_b,
                     int* u_cn, int* u_union, double* u_jac, double* u_aa, int* b_cn,
                     int* b_union, double* b_jac, double* b_aa, int64_t* pa) {
    csr_t ub, bu;
    memset(&ub, 0, sizeof(ub));
    memset(&bu, 0, sizeof(bu));
    int rc = build(n_users, n_biz, n_edges, edge_u, edge_b, &ub, &bu);
    if (rc == 0) rc = side(&ub, &bu, n_pairs, pair_u, pair_b, u_cn, u_union, u_jac, u_aa);
    if (rc == 0) rc = side(&bu, &ub, n_pairs, pair_b, pair_u, b_cn, b_union, b_jac, b_aa);
    if (rc == 0) {
        for (int64_t i = 0; i < n_pairs; ++i) {
            int u = pair_u[i], v = pair_b[i];
            int ok = u >= 0 && u < n_users && v >= 0 && v < n_biz && deg(&ub, u) > 0 &&
                     deg(&bu, v) > 0;
            pa[i] = ok ? (int64_t)deg(&ub, u) * (int64_t)deg(&bu, v) : 0;
        }
    }
    free(ub.off);
    free(ub.adj);
    free(bu.off);
    free(bu.adj);
    return rc;
}
