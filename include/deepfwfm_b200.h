/*
 * deepfwfm_b200.h -- C ABI of libdeepfwfm_sm100a.so
 *
 * B200-native (sm_100a) implementation of the DeepFwFM / DeepLight forward hot path:
 * the body of `DeepFMs.forward(Xi, Xv)` of ShanningLiu/xsDeepFwFM_deprecated
 * (reference: model/DeepFMs.py:285-469 and model/QREmbeddingBag.py:156-174).
 *
 * The reference has no native interface for this path -- it is a chain of ~130 eager aten
 * calls inside one Python method.  The entry points below are what a binding for that
 * method binds instead; each one cites the reference lines it replaces.  The Python host
 * (xsdeepfwfm_deprecated_b200/model/DeepFMs.py) loads this library with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a raw device pointer unless named *_host
 *   - sizes are int64_t / int32_t, `stream` is a cudaStream_t passed as void*
 *   - every function returns int: 0 ok, <0 argument/ABI error (DFW_E_*), >0 a cudaError_t;
 *     nothing throws across the boundary; dfw_last_error_string() describes the last failure
 *     of the calling thread
 *   - all launches are asynchronous on `stream`; no hidden synchronisation except in the
 *     *_host entry points, which say so
 *   - kernels read the model's parameters IN PLACE through the pointers below (no copies), so
 *     in-place edits of the embedding tables and the fp32 MLP weights (the reference's pruner,
 *     model/DeepFMs.py:660-666) are visible to the next call; derived data are the shallow image
 *     (dfw_pack_shallow: field_cov, fwfm_linear, fm_1st, descriptors), the bf16 MLP image
 *     (dfw_pack_mlp_bf16) and the CSR image (dfw_csr_*)
 */
#ifndef DEEPFWFM_B200_H
#define DEEPFWFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFW_ABI_VERSION 3
#define DFW_MAX_DEPTH 8
#define DFW_MAX_FIELDS 64
#define DFW_MAX_K 32
#define DFW_MAX_RANKS 8

/* error codes (negative) */
#define DFW_E_ARG (-1)       /* null pointer / bad size                         */
#define DFW_E_UNSUPPORTED (-2) /* shape outside what the kernels are built for  */
#define DFW_E_ABI (-3)       /* struct size / version mismatch                  */
#define DFW_E_NODEVICE (-4)  /* no sm_100 device: there is NO CPU fallback      */
#define DFW_E_WORKSPACE (-5) /* workspace too small                             */

/* dfw_model.flags */
#define DFW_USE_FWFM   (1u << 0) /* field_cov weights; else FM (all pairs weight 1)  model/DeepFMs.py:353-367 */
#define DFW_USE_FWLW   (1u << 1) /* first order = <E_f, fwfm_linear_f>              model/DeepFMs.py:338-347 */
#define DFW_USE_LW     (1u << 2) /* first order projected by fm_1st                 model/DeepFMs.py:445-450 */
#define DFW_USE_DEEP   (1u << 3) /* MLP term                                        model/DeepFMs.py:395-436 */
#define DFW_CHECK_INDEX (1u << 8) /* out-of-range Xi sets *err_word (and reads row 0) instead of UB */
#define DFW_XI_INT32   (1u << 9) /* every Xi buffer of this model (device and host) holds int32 elements, not int64 -- the packed input
                                    format of SURVEY 8(f): all Criteo / Twitter cardinalities fit, and indices are 80 % of the bytes a
                                    host batch sends over PCIe.  Strides stay in elements; the `const int64_t*` parameters are then
                                    plain addresses of int32 data */

#define DFW_HINT_THROUGHPUT (1u << 10) /* several forwards of this model are kept in flight (a serving loop, bench.py's streams): the
                                    fused kernel then gives each CTA pair two tiles, so that the second tile's gather runs under the
                                    first one's MLP -- 16 % less SM time per batch, 60 % more latency of a launch that runs alone.
                                    Results are bit-identical either way. */

/* qr_op of a table (model/QREmbeddingBag.py:167-172) */
#define DFW_TABLE_PLAIN 0
#define DFW_TABLE_QR_MULT 1
#define DFW_TABLE_QR_ADD 2

/* precision selector of dfw_forward */
#define DFW_PREC_FP32 0      /* CUDA-core fp32 MLP: the 1e-5 parity path                  */
#define DFW_PREC_BF16 1      /* tcgen05 bf16 MLP, fp32 accumulate: the looser-bound path  */
#define DFW_PREC_FP32_CSR 2  /* fp32 CSR MLP for magnitude-pruned weights                 */
#define DFW_PREC_BF16X3 3    /* tcgen05 MLP on split operands x = hi + lo (two bf16): W_hi X_hi + W_hi X_lo + W_lo X_hi,
                                fp32 accumulate -- inside the fp32 parity bound (1e-5 * max|logit|) at tensor-core speed;
                                shapes outside the fused kernel's limits run the CUDA-core fp32 MLP instead             */

/* One embedding field (reference: one entry of fm_2nd_embeddings / fm_1st_embeddings,
 * model/DeepFMs.py:197-210, 1066-1091).  An array of F of these lives in DEVICE memory.
 * Rank-sharded tables use w2_shard[]: row i lives on rank i % n_ranks at local
 * row i / n_ranks; entries are peer-mapped device pointers (dfw_ipc_*).                      */
typedef struct dfw_field_desc {
    const float* w2;      /* (rows, K) second-order table, or quotient table (ceil(rows/c), K) */
    const float* w2_r;    /* (c, K) remainder table for QR, else NULL                          */
    const float* w1;      /* (rows, 1) first-order table / quotient table, NULL with USE_FWLW  */
    const float* w1_r;    /* (c, 1) first-order remainder table for QR, else NULL              */
    int64_t rows;         /* number of categories n_f: valid indices are [0, rows)            */
    int32_t collisions;   /* QR c (>= 1); 1 for a plain table                                 */
    int32_t qr_op;        /* DFW_TABLE_* of w2                                                */
    int32_t qr1_op;       /* DFW_TABLE_* of w1                                                */
    int32_t n_ranks;      /* 0/1: w2 is local; >1: use w2_shard[rank]                          */
    const float* w2_shard[DFW_MAX_RANKS];
} dfw_field_desc;

/* CSR image of one pruned Linear layer (row = output neuron).  Built by dfw_csr_build. */
typedef struct dfw_csr {
    const int32_t* row_ptr; /* (out+1)                 */
    const int32_t* col;     /* (nnz) input index       */
    const float* val;       /* (nnz)                   */
    int32_t nnz;
    int32_t max_row_nnz;
} dfw_csr;

/* Everything `forward` reads, by pointer.  Filled by the host, passed by pointer (host memory). */
typedef struct dfw_model {
    uint32_t struct_bytes;          /* sizeof(dfw_model), checked                            */
    uint32_t abi_version;           /* DFW_ABI_VERSION                                       */
    uint32_t flags;                 /* DFW_USE_*                                             */
    int32_t field_size;             /* F                                                     */
    int32_t numerical;              /* num: fields [0,num) are one-row tables scaled by Xv   */
    int32_t embedding_size;         /* K                                                     */
    int32_t depth;                  /* h_depth                                               */
    int32_t widths[DFW_MAX_DEPTH];  /* hidden widths                                         */
    const dfw_field_desc* fields;   /* DEVICE array [F]                                      */
    const float* fwfm_linear;       /* (F, K)  fwfm_linear.weight        or NULL             */
    const float* fm_1st;            /* (F)     fm_1st.weight             or NULL             */
    const float* field_cov;         /* (F, F)  field_cov.weight          or NULL (FM)        */
    const float* bias;              /* (1)     bias                                          */
    const float* W[DFW_MAX_DEPTH];  /* net_1_linear_l.weight (out, in) row-major fp32        */
    const float* b[DFW_MAX_DEPTH];  /* net_1_linear_l.bias                                   */
    const float* fc;                /* (N) net_1_fc.weight                                   */
    const void* Wbf16[DFW_MAX_DEPTH]; /* dfw_pack_mlp_bf16 image of W[l] (= hi part), or NULL */
    const void* Wbf16_lo[DFW_MAX_DEPTH]; /* lo part bf16(W - hi) from dfw_pack_mlp_bf16_split (DFW_PREC_BF16X3), or NULL */
    dfw_csr csr[DFW_MAX_DEPTH];     /* dfw_csr_build image of W[l] (row_ptr NULL if absent)  */
    const void* shallow_image;      /* dfw_pack_shallow image (DEVICE), or NULL: dfw_forward then packs per call */
    const float* field_cov_host;    /* HOST copy of field_cov.weight (F, F) taken when the plan was built, or NULL.  Derived
                                       data like the shallow image: with it the fused kernel receives the symmetrised field
                                       matrix as a kernel parameter (constant bank) instead of reading it from shared memory */
} dfw_model;

/* ---- library ---------------------------------------------------------------------------- */
int dfw_version(void);
const char* dfw_last_error_string(void);
/* 0 if device `ordinal` is sm_100; DFW_E_NODEVICE otherwise (callers must fail loudly). */
int dfw_check_device(int ordinal);
/* sizeof of the ABI structs as the library was compiled: 0 dfw_model, 1 dfw_field_desc, 2 dfw_csr. */
size_t dfw_struct_bytes(int which);
/* Number of kernels this library has launched in the calling process (bench's gpu_launches). */
int64_t dfw_launch_count(void);

/* ---- stage 1: gather + Xv scale + first order + FM/FwFM second order ---------------------
 * Replaces model/DeepFMs.py:297-367 (and 445-450, the use_lw projection) and
 * model/QREmbeddingBag.py:156-174.
 *   xi   (B, F-num, 1) int64, element strides given (non-contiguous views accepted)
 *   xv   (B, num) fp32, element strides given
 *   E_out      (B, ldE) fp32 or NULL : E[b, f*K+k], field-major == torch.cat(list, 1) (:398); columns
 *              [F*K, ldE) are zero-filled
 *   E_bf16_out (B, ldEb) bf16 or NULL: same values rounded to bf16 (operand of the tcgen05 MLP)
 *   shallow_out (B) fp32: sum(first) + sum(second) + bias   (the non-deep part of :458/:463)
 *   err_word   int32 device word, set to 1+field on an out-of-range index when DFW_CHECK_INDEX
 */
int dfw_embed_fwfm(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                   const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B,
                   float* E_out, int64_t ldE, void* E_bf16_out, int64_t ldEb,
                   float* shallow_out, int32_t* err_word, void* stream);

/* Shallow image: the batch-independent operands of stage 1 in the layout the kernel copies to shared memory --
 * U = strict upper triangle of (field_cov + field_cov^T)/2 (model/DeepFMs.py:364; ones for FM), the compacted
 * list of its non-zero pairs (pruned R, model/DeepFMs.py:667-673), fwfm_linear with fm_1st folded in, and the
 * field descriptors.  Derived data: rebuild after field_cov / fwfm_linear / fm_1st / table pointers change.
 * Runs on the device, no host synchronisation.  `image`: dfw_shallow_image_bytes(m) bytes, 16-byte aligned. */
size_t dfw_shallow_image_bytes(const dfw_model* m);
int dfw_pack_shallow(const dfw_model* m, void* image, void* stream);

/* ---- stage 2: deep MLP + total + optional sigmoid -----------------------------------------
 * Replaces model/DeepFMs.py:408-436 and the sum at :458; `prob_out` fuses the sigmoid every
 * reference caller applies next (model/DeepFMs.py:777, 861, 872).
 *   X (B, ldX) fp32 (= E_out), shallow (B) or NULL, logits_out (B), prob_out (B) or NULL
 *   workspace: dfw_mlp_workspace_bytes(m, B, precision) bytes of device memory, 256-byte aligned
 */
size_t dfw_mlp_workspace_bytes(const dfw_model* m, int64_t B, int precision);
int dfw_mlp_fp32(const dfw_model* m, const float* X, int64_t ldX, int64_t B, const float* shallow,
                 void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out,
                 void* stream);
int dfw_mlp_csr(const dfw_model* m, const float* X, int64_t ldX, int64_t B, const float* shallow,
                void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out,
                void* stream);
/* tcgen05 / TMA path.  Xb (B_pad, ldXb) bf16 where B_pad = B rounded up to 128 rows of readable
 * memory; weights from dfw_pack_mlp_bf16. */
int dfw_mlp_bf16(const dfw_model* m, const void* Xb, int64_t ldXb, int64_t B, const float* shallow,
                 void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out,
                 void* stream);
/* no-deep models: logits = shallow, prob = sigmoid(shallow) */
int dfw_finish_shallow(const float* shallow, int64_t B, float* logits_out, float* prob_out, void* stream);

/* ---- derived weight images ---------------------------------------------------------------- */
/* bf16 image of one Linear weight (out, in) fp32 -> (out_pad, in_pad) bf16, zero padded; sizes via
 * dfw_pack_mlp_bf16_bytes.  Runs on the device (no host sync). */
size_t dfw_pack_mlp_bf16_bytes(int32_t out_dim, int32_t in_dim);
int dfw_pack_mlp_bf16(const float* W, int32_t out_dim, int32_t in_dim, void* dst, void* stream);
/* hi = bf16(W) and lo = bf16(W - hi) images (same layout as dfw_pack_mlp_bf16; either destination may be NULL). */
int dfw_pack_mlp_bf16_split(const float* W, int32_t out_dim, int32_t in_dim, void* dst_hi, void* dst_lo, void* stream);
/* CSR image of one pruned Linear weight.  Two device passes: count (fills row_ptr, returns nnz through
 * a device word the caller reads) and fill.  `row_ptr` has out_dim+1 entries. */
int dfw_csr_count(const float* W, int32_t out_dim, int32_t in_dim, int32_t* row_ptr, void* stream);
int dfw_csr_fill(const float* W, int32_t out_dim, int32_t in_dim, const int32_t* row_ptr,
                 int32_t* col, float* val, void* stream);

/* ---- whole forward -------------------------------------------------------------------------
 * Replaces the body of DeepFMs.forward (model/DeepFMs.py:285-469).  Device inputs.
 *   workspace: dfw_forward_workspace_bytes(m, B, precision) bytes
 */
size_t dfw_forward_workspace_bytes(const dfw_model* m, int64_t B, int precision);
int dfw_forward(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B, int precision,
                void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out,
                int32_t* err_word, void* stream);
/* The single-kernel form of the forward (gather + FwFM on the CUDA cores, MLP on tcgen05, logit/sigmoid epilogue;
 * DFW_PREC_BF16 or DFW_PREC_BF16X3).  dfw_forward takes this route by itself whenever dfw_fused_supported(m, precision);
 * needs m->shallow_image and the bf16 weight images.  No workspace. */
int dfw_fused_supported(const dfw_model* m, int precision);
int dfw_forward_fused(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                      const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B, int precision,
                      float* logits_out, float* prob_out, int32_t* err_word, void* stream);
/* Same with HOST inputs/outputs (what eval_by_batch / predict_proba do around forward,
 * model/DeepFMs.py:771-777): H2D of xi/xv (contiguous) into the staging area at the start of the
 * workspace, forward, sigmoid, D2H of `prob_host` and/or `logits_host`, then SYNCHRONISES `stream`.
 * Host buffers should be pinned.  workspace: dfw_forward_host_workspace_bytes. */
size_t dfw_forward_host_workspace_bytes(const dfw_model* m, int64_t B, int precision);
int dfw_forward_host(const dfw_model* m, const int64_t* xi_host, const float* xv_host, int64_t B,
                     int precision, void* workspace, size_t workspace_bytes, float* logits_host,
                     float* prob_host, void* stream);

/* N samples in (pinned) host memory as batches of `batch` samples, round-robin over internal streams so that the H2D copy of
 * batch i+1 overlaps the kernels of batch i and the D2H of batch i-1; ONE host synchronisation at the end.  The streamed
 * form of eval_by_batch / predict_proba (model/DeepFMs.py:750-784, 864-873).  Results are identical to N/batch calls of
 * dfw_forward_host.  workspace: dfw_forward_host_stream_workspace_bytes(m, batch, precision). */
size_t dfw_forward_host_stream_workspace_bytes(const dfw_model* m, int64_t batch, int precision);
int dfw_forward_host_stream(const dfw_model* m, const int64_t* xi_host, const float* xv_host, int64_t N, int64_t batch,
                            int precision, void* workspace, size_t workspace_bytes, float* logits_host, float* prob_host,
                            void* stream);
/* Transport dfw_forward_host_stream will use for these buffers.  1 = "mapped": every buffer is pinned host memory the device
 * can address (cudaHostAlloc / cudaHostRegister under unified addressing) and the fused kernel takes the model.  Xi / Xv then
 * travel in CHUNKS of up to 8 batches per copy-engine transfer (SM-issued loads from host memory saturate at ~34 GB/s on B200
 * boxes, the copy engine reaches ~55 GB/s once a transfer is a few MB) into rotating staging slots, the fused kernels of a
 * chunk's batches wait on the chunk's event, and their epilogues store logits / probabilities straight into the pinned host
 * buffers -- one launch per batch, no D2H copies.  0 = "staged": cudaMemcpyAsync H2D -> kernels -> D2H per batch (pageable
 * buffers, models outside the fused kernel's shapes).  Results are bit-identical. */
int dfw_host_transport_is_mapped(const dfw_model* m, int precision, const void* xi_host, const void* xv_host,
                                 const void* logits_host, const void* prob_host);

/* ---- one-shot magnitude pruning on the device -------------------------------
 * Replaces the pruning block of fit (model/DeepFMs.py:647-673) applied once at the target rates, and its
 * binary_search_threshold (model/DeepFMs.py:807-823), which costs up to 101 kernel + .item() round trips per tensor.  A set
 * of device tensors (at most 2 * DFW_MAX_FIELDS) is treated as one concatenated array (the reference torch.cat's the
 * fm_2nd_embeddings, :651-655). */
typedef struct dfw_prune_span {
    float* ptr;       /* device tensor, contiguous fp32 */
    int64_t count;    /* elements                       */
} dfw_prune_span;
/* Bisection on t in [0, 100] until count(|w| < (float)t) / total is within 1e-4 of `target`, at most 101 probes -- the same
 * fp64 steps as the Python loop, so *threshold_out (DEVICE double) equals the reference's return value bit for bit.  One
 * cooperative launch, no host synchronisation.  sym_F > 0: `spans` is one (sym_F, sym_F) matrix R and the values are
 * |0.5 (R + R^T)| (the field_cov recipe, :667-670).  probes_out (DEVICE int32) may be NULL.
 * workspace: dfw_prune_workspace_bytes() bytes of device memory, 8-byte aligned. */
size_t dfw_prune_workspace_bytes(void);
int dfw_prune_threshold(const dfw_prune_span* spans, int n_spans, int sym_F, double target, int64_t total,
                        void* workspace, size_t workspace_bytes, double* threshold_out, int32_t* probes_out, void* stream);
/* w = 0 where |w| < (float)*threshold (DEVICE double) over every span (:660-666); sym_F > 0: R[i][j] = R[j][i] = 0 where
 * |0.5 (R + R^T)[i][j]| < threshold (:671-673).  zeroed_out (DEVICE uint64, may be NULL) is incremented by the number of
 * elements under the threshold.  Derived images (shallow image, bf16 / CSR weight images) must be rebuilt afterwards. */
int dfw_prune_apply(const dfw_prune_span* spans, int n_spans, int sym_F, const double* threshold, uint64_t* zeroed_out,
                    void* stream);

/* ---- multi-GPU: row-sharded tables ------------------------------------------- */
/* cudaMalloc'ed shard storage that can be exported to peers of the same node. */
int dfw_shard_alloc(size_t bytes, void** dev_ptr);
int dfw_shard_free(void* dev_ptr);
int dfw_ipc_export(const void* dev_ptr, uint8_t handle_out[64]);
int dfw_ipc_import(const uint8_t handle[64], void** peer_ptr);
int dfw_ipc_close(void* peer_ptr);
/* Build this rank's shard of a table: dst[r, :] = src[r * n_ranks + rank, :]. */
int dfw_shard_rows(const float* src, int64_t rows, int32_t width, int32_t rank, int32_t n_ranks,
                   float* dst, void* stream);
/* NCCL-exchange variant: owner-side gather of requested rows into a send buffer.
 *   req (n) int64 global row ids all owned by this rank -> out (n, width) */
int dfw_gather_rows(const float* shard, int64_t shard_rows, int32_t width, const int64_t* req, int64_t n,
                    int32_t n_ranks, float* out, void* stream);

/* Exchange step as its own kernel ("p2p_pull"): for every sample of the batch and every field listed in sharded_fields (HOST
 * array of field numbers whose descriptors in m->fields carry n_ranks > 1 and the peer-mapped w2_shard pointers), copy the row
 * from the owning GPU -- direct peer loads over NVLink -- into
 *     staged_out (n_sharded, B, K) fp32   (the quotient row for a QR table)
 * and write xi2_out (B, C), a contiguous copy of xi (same element type) whose sharded columns hold b * c + idx mod c, so that
 * a dfw_model whose descriptors point those fields at staged_out (rows = B * c, n_ranks = 0) reads the staging buffer as an
 * ordinary table.  Small co-resident CTAs: launched on another stream it runs one batch ahead of, and under, the fused kernel.
 * Rows are copied, never combined: bit-identical to one GPU.  Out-of-range indices behave as in dfw_embed_fwfm. */
int dfw_pull_rows(const dfw_model* m, const int32_t* sharded_fields, int32_t n_sharded, const int64_t* xi,
                  int64_t xi_stride_b, int64_t xi_stride_c, int64_t B, float* staged_out, void* xi2_out, int32_t* err_word,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPFWFM_B200_H */
