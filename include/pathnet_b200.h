/*
 * pathnet_b200.h — C ABI of libpathnet_b200.so (sm_100a).
 *
 * The drop-in boundary for PathNet's one hot path: the MERW path sampler and
 * the PAGG path-aggregation operator (forward + backward).  The reference has
 * no FFI layer at all — its C++ sampler and its Python model talk through a
 * text file — so each entry point below cites the reference code it replaces
 * (paths relative to the reference repo root).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative pn_status otherwise;
 *     pn_last_error() returns a thread-local message for the last failure.
 *   - "device" pointers are CUDA device pointers on the graph's / call's
 *     current device; "host" pointers are ordinary host memory.
 *   - no ownership transfer: the caller allocates every output and the
 *     workspace; the library keeps no global state besides the error TLS.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all device work is enqueued on it and nothing synchronises unless said.
 */
#ifndef PATHNET_B200_H
#define PATHNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_VERSION 200 /* 0.2.0 */

typedef enum {
    PN_OK = 0,
    PN_ERR_INVALID = -1,   /* bad argument */
    PN_ERR_CUDA = -2,      /* CUDA runtime failure (message has the detail) */
    PN_ERR_DEAD_END = -3,  /* a node with no out-rows: gen_merw.cpp:84-87 exit(0)s there */
    PN_ERR_NOMEM = -4,
    PN_ERR_UNSUPPORTED = -5
} pn_status;

int pn_version(void);
const char *pn_last_error(void);

/* ------------------------------------------------------------------ *
 *  MERW sampler  (replaces preprocess/gen_merw.cpp, gen_epoch_merw.cpp)
 * ------------------------------------------------------------------ */

typedef struct pn_graph pn_graph;

/* Replaces the input parse + link() + AliasTable::init loop
 * (gen_merw.cpp:162-176, :95-99, :23-79).  Rows (u[e], v[e], p[e]) are the
 * `<name>.in` rows in FILE ORDER (host pointers); duplicates are kept.  Builds,
 * on the host, the CSR in file-row order and one alias cell per row exactly as
 * the reference's two-queue construction does in fp64, reduces each cell's
 * fp64 split S to the int32 threshold thr with  (r/RAND_MAX > S) == (r > thr)
 * for every rand() output r, and uploads everything to `device`. */
int pn_graph_from_edges(int32_t n, int64_t m, const int32_t *u, const int32_t *v,
                        const double *p, int device, pn_graph **out);
/* Uniform-random-walk variant of the sampler (preprocess/gen.cpp, the "RW-PathNet" ablation): rows are the
 * 2-column `<name>_nsl.in` rows; adjacency = self loop per node, then both directions of every row with
 * u != v (gen.cpp:83-94); a roll draws ONE rand() (`rand() % deg`, :112-114).  pn_merw_sample on such a
 * graph reproduces gen.cpp's output for srand(seed). */
int pn_graph_from_edges_uniform(int32_t n, int64_t m, const int32_t *u, const int32_t *v, int device,
                                pn_graph **out);
int pn_graph_free(pn_graph *g);

/* n, rows, max out-degree, and the first node without out-rows (-1 if none). */
int pn_graph_info(const pn_graph *g, int32_t *n, int64_t *m, int32_t *max_deg,
                  int32_t *dead_end);

/* Host copies of the alias cells, CSR-aligned with the rows: off[n+1],
 * A[m], B[m], S[m] (fp64 split as the reference stores it) and thr[m].
 * Any pointer may be NULL.  Used by the parity tests. */
int pn_graph_cells_host(const pn_graph *g, int64_t *off, int32_t *A, int32_t *B,
                        double *S, int32_t *thr);

/* Replaces the all-sources bfs() + dis[N][N] table (gen_merw.cpp:101-123,
 * :10, :178-179) for walks of length L: builds on the device the hop-distance
 * table the labels are read from (dense u8 [n][n] saturated at L-1; exact for
 * every distance a walk node can have).  Cached in the graph per L; called
 * implicitly by pn_merw_sample when labels are requested.  Returns
 * PN_ERR_UNSUPPORTED if n*n bytes exceed `max_bytes` (0 = default 24 GiB). */
int pn_graph_prepare_labels(pn_graph *g, int32_t L, size_t max_bytes, void *stream);

typedef enum {
    PN_RNG_GLIBC = 0, /* bit-exact srand(seed)/rand() stream of the reference (seed: low 32 bits) */
    PN_RNG_PHILOX = 1 /* counter-based Philox4x32-10, one stream per walk: counter {lo32(g), hi32(g), t/2, 0}, key = the 64-bit
                       * seed, g = (epoch*n + st)*nw + i; roll t uses words 2(t&1), 2(t&1)+1 of its block, each >> 1.  Not the
                       * reference's stream (nothing but rand() can be), but the same walk law, order-free and shard-invariant:
                       * the mode for multi-GPU / per-epoch on-device resampling where bit-compatibility with a reference
                       * path file is not needed.  CPU restatement: oracle/merw_oracle.c:merw_oracle_sample_philox. */
} pn_rng_mode;

/* Replaces the walk loop (gen_merw.cpp:182-209 / gen_epoch_merw.cpp:164-206).
 * Draws, for epochs [epoch_begin, epoch_end) and start nodes
 * [node_begin, node_end), nw walks of L nodes per start node, reproducing the
 * reference's output for srand(seed) bit for bit: walk (epoch, st, i) consumes
 * rand() outputs [2L*g, 2L*(g+1)) with g = (epoch*n + st)*nw + i, obtained by
 * polynomial jump-ahead of the additive-feedback generator.
 *   nodes  device int32 [E][n'][nw][L]   walk node at position t
 *   types  device uint8 [E][n'][nw][L]   hop distance to the start node
 *                                        (`dis[st][u]-1`, gen_merw.cpp:193,201);
 *                                        NULL = skip labels
 * E = epoch_end-epoch_begin, n' = node_end-node_begin. */
int pn_merw_sample(pn_graph *g, int32_t nw, int32_t L, int64_t epoch_begin,
                   int64_t epoch_end, int32_t node_begin, int32_t node_end,
                   int32_t rng_mode, uint64_t seed, int32_t *nodes, uint8_t *types,
                   void *stream);

/* Replaces compute_merw() (preprocess/compute_merw.py:109-121) whose output init_rw.py:81-86 writes
 * as the p column of `<name>.in`: MERW transition probabilities p[e] = w[e] psi[col[e]] / (lambda psi[row])
 * from the dominant eigenpair of the symmetric adjacency A given in CSR (device pointers; w = NULL means
 * all-ones).  Shifted power iteration in fp64, run per connected component (on a connected graph this
 * is the reference's P; on a disconnected one every component gets a proper stochastic P instead of the
 * reference's negative / zero entries).  Outputs (device): p[m], psi[n] (unit norm per component),
 * lambda[n] (valid at the component's label node), component[n] (smallest node id of the component).
 * Stops when max |psi_new - psi_old| < tol or after max_iter iterations (*iters_out = iterations run). */
int pn_merw_probabilities(int32_t n, const int64_t *rowptr, const int32_t *col, const double *w, double *p,
                          double *psi, double *lambda, int32_t *component, double tol, int32_t max_iter,
                          int32_t *iters_out, void *stream);

/* Text sink / source of the path hand-off format (writer gen_merw.cpp:188-206,
 * reader PathNet_run.py:325-334, :419-423): "[n0, .., nL-1, o0, .., oL-1]\n".
 * Host buffers.  format: returns bytes written, or the bytes needed when
 * out == NULL; < 0 on error.  parse: returns walks parsed, < 0 on error. */
int64_t pn_paths_format_txt(const int32_t *nodes, const uint8_t *types, int64_t walks,
                            int32_t L, char *out, int64_t cap);
int64_t pn_paths_parse_txt(const char *txt, int64_t bytes, int32_t L, int32_t *nodes,
                           uint8_t *types, int64_t max_walks);

/* Reader of the sampler's input files (what gen_merw.cpp:162-172 / gen.cpp:79-94 read with fscanf): header "n m",
 * then m rows "u v p" (cols = 3) or "u v" (cols = 2).  Host buffers; the doubles are correctly rounded like strtod's.
 * With u == NULL only the header is read (*n_out, *m_out) and 0 is returned; otherwise returns m (rows stored; rows
 * beyond m are ignored as in the reference), < 0 on error (fewer than m rows, m > max_rows, malformed header). */
int64_t pn_edges_parse_txt(const char *txt, int64_t bytes, int32_t cols, int64_t max_rows, int32_t *n_out,
                           int64_t *m_out, int32_t *u, int32_t *v, double *p);

/* ------------------------------------------------------------------ *
 *  PAGG operator  (replaces PathNet.forward / PathNet_homo.forward and
 *  their autograd backward, PathNet_run.py:172-211, :239-278, :351)
 * ------------------------------------------------------------------ */

typedef enum { PN_VARIANT_HETERO = 0, /* class PathNet      (:150) */
               PN_VARIANT_HOMO = 1,   /* class PathNet_homo (:214) */
               PN_VARIANT_PLAIN = 2   /* natural path order, no ReLUs, no attention partner: the older `PAGG` class
                                         (baseline/GPRGNN/src/copy.py:299-359) and the pooling ablations */ } pn_variant;

/* SURVEY 8f-4 variants (fp32 math mode; the tensor-core kernels implement the LSTM cell only):
 *   cell: LSTM (PathNet_run.py:164), the tanh RNN of copy.py:307 (then w_ih / w_hh are [H,H], b_ih / b_hh [H]), or a GRU
 *         (the "GRU" ablation of performance.jpg; torch nn.GRU semantics and layout: w_ih / w_hh [3H,H] in r, z, n order,
 *         b_ih / b_hh [3H]; n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) * n + z * h)
 *   pool: the variant's own attention (:199-204 / :269-273), or mean (copy.py:350) / sum / max over the path_num
 *         final hidden states (attw_* are then unused and their gradients zero) */
typedef enum { PN_CELL_LSTM = 0, PN_CELL_RNN = 1, PN_CELL_GRU = 2 } pn_cell;
typedef enum { PN_POOL_ATTENTION = 0, PN_POOL_MEAN = 1, PN_POOL_SUM = 2, PN_POOL_MAX = 3 } pn_pool;

typedef enum { PN_MATH_FP32 = 0, /* CUDA-core fp32 FMA, strict parity         */
               PN_MATH_TC = 1,   /* tcgen05 tensor cores, fp16 operands/fp32 acc */
               PN_MATH_TC3 = 2   /* tcgen05 tensor cores at fp32-grade accuracy: the LSTM's operands travel as two fp16
                                    terms (hi + lo, 22 bits) and every product is three tensor-core passes into one fp32
                                    accumulator; ex2/rcp gates; projections on the fp32 kernels.  Same parity bar as
                                    PN_MATH_FP32 (hidden 128, path_len <= 8, LSTM cell)                               */ } pn_math;

typedef struct {
    int32_t N, F, H, C;      /* nodes, input features, hidden, classes       */
    int32_t S, nw, L;        /* output rows (sum(indices)), path_num, path_len */
    int32_t variant;         /* pn_variant                                   */
    int32_t math_mode;       /* pn_math                                      */
    int32_t training;        /* F.dropout(training=...) (:194,:209,:264,:276) */
    float dropout_p;         /* the reference's module-global `dropout` (:71) */
    uint64_t dropout_seed;   /* counter-based mask stream    */
    int32_t no_backward;     /* 1: inference only — forward skips saving activations / gradient operands;
                                pn_pagg_backward on such a workspace is an error (PN_ERR_INVALID)       */
    int32_t row_chunk;       /* output rows processed per pass over the LSTM / pooling kernels.  0 = all S rows at
                                once (the workspace then scales with S*nw sequences).  > 0: the per-sequence part of
                                the workspace is sized for row_chunk rows; forward loops over chunks sharing one Z / T
                                table, backward RECOMPUTES each chunk's forward before its backward (activation
                                checkpointing by row chunk), so 10^8-path calls run in a bounded workspace.          */
    const uint64_t *dropout_seed_dev; /* optional device word ADDED to dropout_seed inside the kernels (NULL = none):
                                lets a captured CUDA graph draw fresh masks on every replay                         */
    const void *x16;         /* optional device copy of X as fp16 [N][F16] row-major, F16 = F rounded up to 8 (zero
                                padded), e.g. made once by pn_pagg_pack_x16: the tensor-core mode then TMA-loads it
                                instead of re-reading and re-converting the fp32 X on every call (X is constant over
                                a training run, PathNet_run.py:300).  NULL = convert per call.                       */
    int32_t cell;            /* a pn_cell value: 0 = LSTM, the reference's                                                   */
    int32_t pool;            /* a pn_pool value: 0 = the variant's attention                               */
} pn_pagg_cfg;

/* Parameter (or gradient) block, fp32 device pointers, PyTorch layouts
 * (state_dict keys of PathNet_run.py:155-170):
 *   fc0_w[H,F] fc0_b[H]  nets_w[L,H,H] nets_b[L,H] (nets.{d}.weight/bias stacked)
 *   w_ih[4H,H] w_hh[4H,H] b_ih[4H] b_hh[4H] (LSTM.*_l0, gate order i,f,g,o)
 *   attw_w[2H] attw_b[1]  fc2_w[C,2H] fc2_b[C] */
typedef struct {
    float *fc0_w, *fc0_b, *nets_w, *nets_b, *w_ih, *w_hh, *b_ih, *b_hh, *attw_w, *attw_b,
        *fc2_w, *fc2_b;
} pn_pagg_params;

typedef enum { PN_IDX_I64 = 0, PN_IDX_I32 = 1, PN_IDX_U8 = 2 } pn_index_dtype;

/* Integer-only canonicalisation of the reference's index tensors into the
 * form one fused kernel serves for both variants (DESIGN.md §PAGG canonical
 * form).  neis[S, nw*L] and layer_type[S, nw, L] are the index tensors the
 * reference's forward() receives (device copies): int64 as the reference
 * passes them, or the sampler's own int32 node ids / uint8 labels
 * (`*_dtype` = pn_index_dtype; node ids: i64 or i32, labels: i64, i32 or u8).
 * Outputs, device:
 *   seq_node[S][nw][L] int32, seq_type[S][nw][L] uint8 : the L (type,node)
 *       table entries fed to the LSTM for the w-th sequence pooled into row s
 *   partner[S][nw] int32 : node whose embedding joins h_n in the attention
 * hetero reproduces the flip/reshape/view re-interpretations of
 * PathNet_run.py:179-199 exactly.
 * Range check (the reference raises IndexError / a device assert on X[neis],
 * nei[indxx, layer_type]): a node id outside [0,N) or a label outside [0,L)
 * sets bit 0 / bit 1 of *status and is replaced by 0, so no later kernel ever
 * indexes out of bounds.  `status` is one int32 the kernel can write — device
 * memory or pinned host memory (readable by the host without a
 * synchronisation once the kernel has run); NULL skips the report (the
 * sanitisation still happens).  The caller zeroes it. */
int pn_pagg_canonicalize(int32_t variant, int32_t N, int32_t S, int32_t nw, int32_t L,
                         const void *neis, int32_t neis_dtype, const void *layer_type,
                         int32_t type_dtype, int32_t *seq_node, uint8_t *seq_type,
                         int32_t *partner, int32_t *status, void *stream);

/* fp16 copy of X for pn_pagg_cfg.x16: [N][F16] row-major, F16 = F rounded up to a multiple of 8, zero padded
 * (pn_pagg_x16_bytes(N, F) bytes, 16-byte aligned).  X does not change during training (PathNet_run.py:300), so a
 * caller makes this once and the tensor-core mode TMA-loads it on every step instead of re-reading fp32 X. */
size_t pn_pagg_x16_bytes(int32_t N, int32_t F);
int pn_pagg_pack_x16(int32_t N, int32_t F, const float *X, void *x16, void *stream);

/* Bytes of device workspace forward/backward need for this cfg (saved
 * activations + scratch).  The same workspace must be passed to forward and
 * to the matching backward, untouched in between. */
size_t pn_pagg_workspace_bytes(const pn_pagg_cfg *cfg);

/* forward: logits[S,C] = PAGG(X[N,F]; params) for the S rows whose node ids are
 * rows[S] (= nonzero(indices)).  Replaces Module.forward. */
int pn_pagg_forward(const pn_pagg_cfg *cfg, const pn_pagg_params *params, const float *X,
                    const int32_t *seq_node, const uint8_t *seq_type, const int32_t *partner,
                    const int32_t *rows, void *workspace, float *logits, void *stream);

/* backward: given dlogits[S,C], OVERWRITES every array of `grads` with the
 * parameter gradients (the reference never needs dX).  Replaces autograd. */
int pn_pagg_backward(const pn_pagg_cfg *cfg, const pn_pagg_params *params, const float *X,
                     const int32_t *seq_node, const uint8_t *seq_type, const int32_t *partner,
                     const int32_t *rows, void *workspace, const float *dlogits,
                     const pn_pagg_params *grads, void *stream);

/* ------------------------------------------------------------------ *
 *  Rest of the training step on the device (PathNet_run.py:345-352), so that
 *  forward -> loss -> backward -> optimizer is one stream of device work
 *  (capturable as one CUDA graph)
 * ------------------------------------------------------------------ */

/* Per-step device state: `step` (Adam's t) and `seed_add` (added to
 * cfg.dropout_seed inside the kernels when cfg.dropout_seed_dev points at it).
 * pn_step_advance increments both on the device, so a captured graph draws
 * fresh dropout masks and uses the right bias correction on every replay. */
typedef struct { uint64_t seed_add; int64_t step; } pn_step_state;
int pn_step_advance(pn_step_state *state /* device */, void *stream);

/* Replaces lossfunc(y_hat, Y[train_indices]) and its autograd (torch.nn.CrossEntropyLoss, :297,:348,:351):
 *   *loss += inv_count * sum_s [ logsumexp(logits[s]) - logits[s][labels[s]] ]   (device float, caller zeroes it)
 *   dlogits[s][c] = inv_count * (softmax(logits[s])[c] - [c == labels[s]])        (may be NULL: loss only)
 * inv_count = 1 / (number of rows of the GLOBAL batch), so shards of a data-parallel step add up to the mean.
 * labels: device int64 or int32 (`label_dtype` = pn_index_dtype), one per row. */
int pn_ce_loss_grad(int64_t S, int32_t C, const float *logits, const void *labels, int32_t label_dtype,
                    float inv_count, float *loss, float *dlogits, void *stream);

/* Replaces optimizer.step() of torch.optim.Adam(params, lr, weight_decay) (:295-296,:352) over flat fp32 buffers
 * (params, grads, exp_avg, exp_avg_sq: device, n elements each); t = state->step (device). */
typedef struct { float lr, beta1, beta2, eps, weight_decay; } pn_adam_cfg;
int pn_adam_step(int64_t n, float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                 const pn_adam_cfg *cfg, const pn_step_state *state /* device */, void *stream);

/* The 16-bit dropout hash the kernels use for element `idx` (dropped iff
 * hash < floor(p*65536)) of mask
 * stream `stream_id` (0 = LSTM input [L,P,H], 1 = concat [S,2H]); host
 * function so tests can feed the oracle the identical masks. */
uint32_t pn_dropout_hash(uint64_t seed, uint32_t stream_id, uint64_t idx);

/* Debug/unit-test entry: C[M,N] = act(A[M,K] B[N,K]^T + bias) on the tcgen05
 * path (fp16 operands, fp32 accumulate); device pointers, fp32, row-major. */
int pn_debug_gemm_tc(const float *A, const float *B, float *C, const float *bias, int64_t M,
                     int64_t N, int64_t K, int relu, void *stream);
/* same with operand-layout flags: bit0 A is [K][M], bit1 B is [K][N], bit2 bf16 operands, bit4 split precision (hi + lo fp16 terms, 3 passes),
 * bit3 accumulate into C (atomics); splitk >= 1 splits K across CTAs (C pre-zeroed). */
int pn_debug_gemm_tc2(const float *A, const float *B, float *C, int64_t M, int64_t N, int64_t K,
                      int flags, int splitk, void *stream);

/* Unit-test entry of the TMA-fed GEMM (gemm_tma.cu): C[M,N] (fp32, may be NULL) and C16[M,N] (fp16 of the value times
 * c16_scale, may be NULL) = act(A[M,K] B[N,K]^T + bias); fp32 device inputs are converted to fp16 in `scratch`
 * ((M + N) * roundup8(K) halves) first. */
int pn_debug_gemm_tma(const float *A, const float *B, float *C, void *C16, const float *bias, int64_t M, int64_t N,
                      int64_t K, int relu, float c16_scale, void *scratch, void *stream);

/* In-stream kernel timing for bench.py's roofline: when enabled, the library
 * brackets its dominant kernels with CUDA events on the launching stream.
 * pn_profile_read synchronises those events and returns, for slot `name`
 * ("lstm_fwd", "lstm_bwd", "merw_walk", "wgrad", ...), the summed milliseconds
 * and the number of launches since the last pn_profile_enable(1). */
int pn_profile_enable(int on);
int pn_profile_read(const char *name, double *ms, int64_t *launches);

/* Number of kernels this library has launched in this process (all entry
 * points; bench.py reports it as gpu_launches). */
uint64_t pn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PATHNET_B200_H */
