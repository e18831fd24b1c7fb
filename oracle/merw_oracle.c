/*
 * oracle/merw_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference MERW path
 * sampler (/root/reference/preprocess/gen_merw.cpp and gen_epoch_merw.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library, and only as the checker or
 * the timed CPU baseline — the product (libpathnet_b200.so) never links,
 * loads or calls it.
 *
 * Parity pinning: this restatement is checked byte-for-byte against the
 * UNMODIFIED reference binary (oracle/_ref/gen_merw, built by `make ref`
 * from the sources where they lie, run under oracle/faketime.c) in
 * tests/test_oracle_pinning.py, and against the committed fixtures in
 * tests/golden/ that were produced by that binary (tests/golden/make_golden.py).
 *
 * The RNG is glibc's srand()/rand() (TYPE_3 additive feedback generator);
 * glibc is a third-party dependency not vendored in /root/reference
 * (call sites: gen_merw.cpp:88-89 and :161).  Its published algorithm is
 * restated in glibc_srand()/glibc_rand() below and pinned against the libc
 * in this image in tests/test_oracle_pinning.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ RNG -- */
/* glibc stdlib/random_r.c, TYPE_3: degree 31, separation 3.               */
typedef struct {
    uint32_t r[31];
    int f, b; /* front / rear indices */
} glibc_rng;

void glibc_srand(glibc_rng *g, unsigned int seed)
{
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    g->r[0] = (uint32_t)word;
    for (int i = 1; i < 31; i++) {
        /* 16807 * word % 2147483647 without overflow (Schrage) */
        long hi = word / 127773;
        long lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        g->r[i] = (uint32_t)word;
    }
    g->f = 3;
    g->b = 0;
    for (int i = 0; i < 310; i++) {
        g->r[g->f] += g->r[g->b];
        if (++g->f == 31) g->f = 0;
        if (++g->b == 31) g->b = 0;
    }
}

int glibc_rand(glibc_rng *g)
{
    uint32_t v = (g->r[g->f] += g->r[g->b]);
    if (++g->f == 31) g->f = 0;
    if (++g->b == 31) g->b = 0;
    return (int)(v >> 1);
}

/* exported for the pinning test: first `count` outputs after srand(seed) */
void merw_oracle_rand_stream(unsigned int seed, long count, int *out)
{
    glibc_rng g;
    glibc_srand(&g, seed);
    for (long i = 0; i < count; i++) out[i] = glibc_rand(&g);
}

/* ---------------------------------------------------------------- graph -- */
typedef struct {
    int n;
    long m;
    long *off; /* [n+1] rows per node, file row order (gen_merw.cpp:95-99) */
    int *adj;  /* [m]   E[u]                                               */
    /* alias cells, CSR aligned with the rows (gen_merw.cpp:23-79) */
    int *A, *B;
    double *S;
    int uniform; /* 1: the uniform-random-walk variant preprocess/gen.cpp (one rand() per roll) */
} merw_graph;

/* AliasTable::init, gen_merw.cpp:23-79, for one node's rows.
 * ids/p: the node's deg rows; out cells appended at A/B/S; returns #cells. */
static long alias_init(int deg, const int *ids, const double *p_in, int *A,
                       int *B, double *S, int *qa_id, double *qa_p, int *qb_id,
                       double *qb_p, double *p)
{
    long cells = 0;
    long ah = 0, at = 0, bh = 0, bt = 0; /* FIFO heads/tails */
    for (int i = 0; i < deg; i++) p[i] = p_in[i] * deg;         /* :27 */
    for (int i = 0; i < deg; i++) {                               /* :28-35 */
        if (p[i] > 1.0) { qa_id[at] = ids[i]; qa_p[at++] = p[i]; }
        else            { qb_id[bt] = ids[i]; qb_p[bt++] = p[i]; }
    }
    while (ah < at && bh < bt) {                                  /* :36 */
        int idA = qa_id[ah]; double probA = qa_p[ah++];
        int idB = qb_id[bh]; double probB = qb_p[bh++];
        A[cells] = idA; B[cells] = idB; S[cells] = probB; cells++; /* :42-44 */
        double res = probA - (1.0 - probB);                       /* :46 */
        if (fabs(res - 1.0) < 1e-5) {                             /* :48-53 */
            A[cells] = idA; B[cells] = idA; S[cells] = res; cells++;
            continue;
        }
        if (res > 1.0) { qa_id[at] = idA; qa_p[at++] = res; }     /* :55-61 */
        else           { qb_id[bt] = idA; qb_p[bt++] = res; }
    }
    while (ah < at) {                                             /* :64-70 */
        int id = qa_id[ah++];
        A[cells] = id; B[cells] = id; S[cells] = 1.0; cells++;
    }
    while (bh < bt) {                                             /* :72-78 */
        int id = qb_id[bh++];
        A[cells] = id; B[cells] = id; S[cells] = 1.0; cells++;
    }
    return cells;
}

/* parse + link (gen_merw.cpp:162-172, :95-99) from in-memory edge rows,
 * then per-node alias build (:174-176).  Returns NULL on failure. */
merw_graph *merw_oracle_graph(int n, long m, const int *u, const int *v,
                              const double *p)
{
    merw_graph *g = (merw_graph *)calloc(1, sizeof(*g));
    g->n = n; g->m = m;
    g->off = (long *)calloc((size_t)n + 1, sizeof(long));
    g->adj = (int *)malloc(sizeof(int) * (size_t)(m ? m : 1));
    double *pr = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1));
    for (long e = 0; e < m; e++) g->off[u[e] + 1]++;
    for (int i = 0; i < n; i++) g->off[i + 1] += g->off[i];
    long *cur = (long *)malloc(sizeof(long) * (size_t)(n ? n : 1));
    memcpy(cur, g->off, sizeof(long) * (size_t)n);
    for (long e = 0; e < m; e++) { /* stable: keeps file row order per node */
        long k = cur[u[e]]++;
        g->adj[k] = v[e];
        pr[k] = p[e];
    }
    free(cur);
    g->A = (int *)malloc(sizeof(int) * (size_t)(m ? m : 1));
    g->B = (int *)malloc(sizeof(int) * (size_t)(m ? m : 1));
    g->S = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1));
    int maxdeg = 0;
    for (int i = 0; i < n; i++) {
        long d = g->off[i + 1] - g->off[i];
        if (d > maxdeg) maxdeg = (int)d;
    }
    size_t q = (size_t)2 * maxdeg + 2;
    int *qa_id = (int *)malloc(sizeof(int) * q), *qb_id = (int *)malloc(sizeof(int) * q);
    double *qa_p = (double *)malloc(sizeof(double) * q), *qb_p = (double *)malloc(sizeof(double) * q);
    double *ptmp = (double *)malloc(sizeof(double) * (size_t)(maxdeg + 1));
    int ok = 1;
    for (int i = 0; i < n; i++) {
        long o = g->off[i];
        int d = (int)(g->off[i + 1] - o);
        long c = alias_init(d, g->adj + o, pr + o, g->A + o, g->B + o, g->S + o,
                            qa_id, qa_p, qb_id, qb_p, ptmp);
        if (c != d) ok = 0; /* cells == rows always (pop 2 / emit 1 / push 1) */
    }
    free(qa_id); free(qb_id); free(qa_p); free(qb_p); free(ptmp); free(pr);
    if (!ok) { fprintf(stderr, "merw_oracle: cell count != row count\n"); }
    return g;
}

/* Adjacency of the uniform-random-walk sampler, preprocess/gen.cpp:83-94: a self loop for every
 * node first (link(i,i)), then for every input row (u,v) with u != v: link(u,v) and link(v,u).
 * Rows are 2-column (`<name>_nsl.in`).  A walk step is `u = E[u][rand() % E[u].size()]` (:112-114). */
merw_graph *merw_oracle_graph_uniform(int n, long m, const int *u, const int *v)
{
    merw_graph *g = (merw_graph *)calloc(1, sizeof(*g));
    g->n = n; g->uniform = 1;
    long rows = n;
    for (long e = 0; e < m; e++) if (u[e] != v[e]) rows += 2;
    g->m = rows;
    g->off = (long *)calloc((size_t)n + 1, sizeof(long));
    g->adj = (int *)malloc(sizeof(int) * (size_t)rows);
    for (int i = 0; i < n; i++) g->off[i + 1] = 1;
    for (long e = 0; e < m; e++) if (u[e] != v[e]) { g->off[u[e] + 1]++; g->off[v[e] + 1]++; }
    for (int i = 0; i < n; i++) g->off[i + 1] += g->off[i];
    long *cur = (long *)malloc(sizeof(long) * (size_t)n);
    memcpy(cur, g->off, sizeof(long) * (size_t)n);
    for (int i = 0; i < n; i++) g->adj[cur[i]++] = i;
    for (long e = 0; e < m; e++) if (u[e] != v[e]) { g->adj[cur[u[e]]++] = v[e]; g->adj[cur[v[e]]++] = u[e]; }
    free(cur);
    g->A = (int *)malloc(sizeof(int) * (size_t)rows);
    g->B = (int *)malloc(sizeof(int) * (size_t)rows);
    g->S = (double *)malloc(sizeof(double) * (size_t)rows);
    for (long k = 0; k < rows; k++) { g->A[k] = g->B[k] = g->adj[k]; g->S[k] = 1.0; }
    return g;
}

void merw_oracle_graph_free(merw_graph *g)
{
    if (!g) return;
    free(g->off); free(g->adj); free(g->A); free(g->B); free(g->S); free(g);
}

/* raw alias cells, for pinning the product's host-side alias build */
void merw_oracle_graph_cells(const merw_graph *g, long *off, int *A, int *B,
                             double *S)
{
    memcpy(off, g->off, sizeof(long) * ((size_t)g->n + 1));
    memcpy(A, g->A, sizeof(int) * (size_t)g->m);
    memcpy(B, g->B, sizeof(int) * (size_t)g->m);
    memcpy(S, g->S, sizeof(double) * (size_t)g->m);
}

/* AliasTable::roll, gen_merw.cpp:81-91.  Returns -1 for a dead-end node
 * (the reference prints to stderr and exit(0)s there, :84-87). */
static int alias_roll(const merw_graph *g, int u, glibc_rng *rng)
{
    long o = g->off[u];
    int cells = (int)(g->off[u + 1] - o);
    if (cells == 0) return -1;
    if (g->uniform) return g->adj[o + glibc_rand(rng) % cells]; /* gen.cpp:112-114: one draw */
    int x = glibc_rand(rng) % cells;
    double p = 1.0 * glibc_rand(rng) / 2147483647; /* RAND_MAX */
    return p > g->S[o + x] ? g->A[o + x] : g->B[o + x];
}

/* bfs(), gen_merw.cpp:101-123, into a per-root row `dis` (n bytes, 1-based,
 * 0 = unseen) instead of the reference's dis[N][N] table; `touched` lets the
 * caller reset the row cheaply.  queue: scratch of n ints. */
static long bfs_row(const merw_graph *g, int S, int seq_len, unsigned char *dis,
                    int *queue)
{
    long qh = 0, qt = 0;
    queue[qt++] = S;
    dis[S] = 1;
    while (qh < qt) {
        int u = queue[qh++];
        if (dis[u] > seq_len) break; /* :110-111 `return` */
        for (long i = g->off[u]; i < g->off[u + 1]; i++) {
            int v = g->adj[i];
            if (dis[v] == 0) {
                dis[v] = (unsigned char)(dis[u] + 1);
                queue[qt++] = v;
            }
        }
    }
    return qt; /* number of touched nodes = queue[0..qt) */
}

/*
 * The walk loop, gen_merw.cpp:182-209: epochs [0, epoch_end) are generated
 * in the reference's order (epoch-major, start node, walk index) from
 * srand(seed); epochs < epoch_begin only advance the RNG.  Outputs for
 * epochs [epoch_begin, epoch_end):
 *   nodes[e][st][i][t]  int32   walk node at position t
 *   types[e][st][i][t]  uint8   dis[st][u]-1 (hop distance to the start node)
 * Labels are filled in a second pass (one BFS per start node) which yields
 * exactly dis[st][u] without the 10 GB table.  Returns 0, or -1 if a
 * dead-end node was rolled (reference would exit(0)).
 */
int merw_oracle_sample_range(const merw_graph *g, int nw, int L, unsigned int seed,
                             long epoch_begin, long epoch_end, int node_begin, int node_end,
                             int32_t *nodes, uint8_t *types)
{
    /* Same as the reference loop, but only start nodes [node_begin, node_end)
     * are materialised: nodes[e][st - node_begin][i][t].  Walks of other start
     * nodes only advance the RNG (2 draws per roll, L rolls, whatever happens). */
    glibc_rng rng;
    glibc_srand(&rng, seed);
    const int n = g->n;
    const long np = node_end - node_begin;
    for (long epoch = 0; epoch < epoch_end; epoch++) {
        if (epoch < epoch_begin) {
            long draws = (long)n * nw * L * (g->uniform ? 1 : 2);
            for (long k = 0; k < draws; k++) (void)glibc_rand(&rng);
            continue;
        }
        for (int st = 0; st < n; st++) {
            if (st < node_begin || st >= node_end) {
                long draws = (long)nw * L * (g->uniform ? 1 : 2);
                for (long k = 0; k < draws; k++) (void)glibc_rand(&rng);
                continue;
            }
            for (int i = 0; i < nw; i++) {
                int32_t *o = nodes + ((((epoch - epoch_begin) * np + (st - node_begin)) * nw + i) * L);
                int u = st;
                for (int t = 0; t < L; t++) {
                    o[t] = u;
                    u = alias_roll(g, u, &rng); /* L rolls, last one unused */
                    if (u < 0) return -1;
                }
            }
        }
    }
    if (!types) return 0;
    unsigned char *dis = (unsigned char *)calloc((size_t)(n ? n : 1), 1);
    int *queue = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
    long E = epoch_end - epoch_begin;
    for (int st = node_begin; st < node_end; st++) {
        long touched = bfs_row(g, st, L, dis, queue);
        for (long e = 0; e < E; e++)
            for (int i = 0; i < nw; i++) {
                long base = ((e * np + (st - node_begin)) * nw + i) * L;
                for (int t = 0; t < L; t++)
                    types[base + t] = (uint8_t)(dis[nodes[base + t]] - 1);
            }
        for (long k = 0; k < touched; k++) dis[queue[k]] = 0;
    }
    free(dis); free(queue);
    return 0;
}

int merw_oracle_sample(const merw_graph *g, int nw, int L, unsigned int seed,
                       long epoch_begin, long epoch_end, int32_t *nodes,
                       uint8_t *types)
{
    return merw_oracle_sample_range(g, nw, L, seed, epoch_begin, epoch_end, 0, g->n, nodes, types);
}

/* ---------------------------------------------------------------------------
 * rng_mode = philox (an extension of this repo, SURVEY 8b: `rng_mode{glibc_compat, philox}`; the reference has only
 * the sequential rand() stream).  Philox4x32-10, Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
 * restated from the paper; pinned by the Random123 known-answer vectors in tests/test_oracle_pinning.py.
 * Walk g = (epoch * n + st) * nw + i (the reference's walk order) owns the counter stream {lo32(g), hi32(g), b, 0}
 * under key {lo32(seed), hi32(seed)}; roll t (t = 0 .. L-2, the wasted L-th roll of the reference is not drawn) takes
 * words 2(t & 1), 2(t & 1) + 1 of block b = t / 2, each >> 1 (31 bits, like rand()):
 *     x = r1 % cells(u);  next = r2 > thr[x] ? A[x] : B[x]   with  r2 > thr  <=>  1.0 * r2 / RAND_MAX > S  (gen_merw.cpp:88-90)
 * Every walk is independent of every other: any sharding of epochs / start nodes / walks gives the same walks. */
void merw_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

int merw_oracle_sample_philox(const merw_graph *g, int nw, int L, unsigned long long seed,
                              long epoch_begin, long epoch_end, int node_begin, int node_end,
                              int32_t *nodes, uint8_t *types)
{
    const int n = g->n;
    const long np = node_end - node_begin;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (long epoch = epoch_begin; epoch < epoch_end; epoch++)
        for (int st = node_begin; st < node_end; st++)
            for (int i = 0; i < nw; i++) {
                int32_t *o = nodes + ((((epoch - epoch_begin) * np + (st - node_begin)) * nw + i) * L);
                const uint64_t gw = ((uint64_t)epoch * (uint64_t)n + (uint64_t)st) * (uint64_t)nw + (uint64_t)i;
                uint32_t blk[4] = {0, 0, 0, 0};
                int u = st;
                for (int t = 0; t < L; t++) {
                    o[t] = u;
                    if (t + 1 == L) break;
                    if ((t & 1) == 0) {
                        const uint32_t ctr[4] = {(uint32_t)gw, (uint32_t)(gw >> 32), (uint32_t)(t >> 1), 0u};
                        merw_oracle_philox4x32_10(ctr, key, blk);
                    }
                    const uint32_t r1 = blk[2 * (t & 1)] >> 1, r2 = blk[2 * (t & 1) + 1] >> 1;
                    const long off = g->off[u];
                    const int cells = (int)(g->off[u + 1] - off);
                    if (cells == 0) return -1;
                    const int x = (int)(r1 % (uint32_t)cells);
                    if (g->uniform) u = g->adj[off + x];
                    else {
                        const double p = 1.0 * (double)r2 / 2147483647;
                        u = p > g->S[off + x] ? g->A[off + x] : g->B[off + x];
                    }
                }
            }
    if (!types) return 0;
    unsigned char *dis = (unsigned char *)calloc((size_t)(n ? n : 1), 1);
    int *queue = (int *)malloc(sizeof(int) * (size_t)(n ? n : 1));
    long E = epoch_end - epoch_begin;
    for (int st = node_begin; st < node_end; st++) {
        long touched = bfs_row(g, st, L, dis, queue);
        for (long e = 0; e < E; e++)
            for (int i = 0; i < nw; i++) {
                long base = ((e * np + (st - node_begin)) * nw + i) * L;
                for (int t = 0; t < L; t++)
                    types[base + t] = (uint8_t)(dis[nodes[base + t]] - 1);
            }
        for (long k = 0; k < touched; k++) dis[queue[k]] = 0;
    }
    free(dis); free(queue);
    return 0;
}

/* Text sink, gen_merw.cpp:188-206: "[n0, n1, ..., o0, o1, ...]\n" per walk.
 * Returns bytes written (or needed, when buf is NULL). */
long merw_oracle_format(const int32_t *nodes, const uint8_t *types, long walks,
                        int L, char *buf)
{
    long pos = 0;
    char tmp[16];
    for (long w = 0; w < walks; w++) {
        if (buf) buf[pos] = '['; pos++;
        for (int t = 0; t < 2 * L; t++) {
            int val = t < L ? nodes[w * L + t] : (int)types[w * L + (t - L)];
            int k = snprintf(tmp, sizeof tmp, "%d", val);
            if (buf) memcpy(buf + pos, tmp, (size_t)k);
            pos += k;
            if (t != 2 * L - 1) { if (buf) { buf[pos] = ','; buf[pos + 1] = ' '; } pos += 2; }
        }
        if (buf) { buf[pos] = ']'; buf[pos + 1] = '\n'; } pos += 2;
    }
    return pos;
}
