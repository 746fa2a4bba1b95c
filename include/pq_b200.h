/*
 * pq_b200.h -- C ABI of the B200-native tensor-contraction backend for PicoQuant.
 *
 * One `pq_handle` == one GPU == one CUDA stream == one device tensor store keyed
 * by label, i.e. the device-side twin of PicoQuant's `InteractiveBackend{T}`
 * (reference: src/backends/interactive.jl:10-23).  Every entry point below
 * replaces one method of the reference's backend interface
 * (src/backends.jl:62-66, forwarded at src/layer3.jl:124-134); the Julia-side
 * `ccall` binding is shown in INTEGRATION.md.
 *
 * Conventions (same as the reference): labels are NUL-terminated strings (Julia
 * Symbols), tensors are dense COLUMN-MAJOR, and every index / axis / range value
 * crossing this interface is 1-BASED.  Functions return 0 on success or a negative
 * `pq_status`; nothing throws across the ABI; `pq_last_error` gives the message.
 * All device work is asynchronous on the handle's stream except `pq_load_tensor`
 * and `pq_sync`.  A handle is not thread-safe.  There is no CPU fallback: creating
 * a handle without a usable CUDA device fails.
 */
#ifndef PQ_B200_H
#define PQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pq_handle pq_handle;
typedef struct pq_program pq_program;

typedef enum {
  PQ_OK = 0,
  PQ_ERR_INVALID = -1,   /* bad argument (rank, axis, permutation, label list ...) */
  PQ_ERR_NOT_FOUND = -2, /* label absent (Julia: KeyError / `nothing`)             */
  PQ_ERR_SHAPE = -3,     /* extent mismatch (Julia: DimensionMismatch)             */
  PQ_ERR_CUDA = -4,      /* CUDA runtime / launch failure                          */
  PQ_ERR_NCCL = -5,      /* NCCL unavailable or failed                             */
  PQ_ERR_PARSE = -6,     /* malformed .tl command stream                           */
  PQ_ERR_UNSUPPORTED = -7
} pq_status;

/* element type the backend computes and stores in (the `T` of InteractiveBackend{T}) */
typedef enum { PQ_C64 = 0, PQ_C128 = 1 } pq_dtype;
/* element type of a host buffer handed to save/load */
typedef enum { PQ_HOST_F32 = 0, PQ_HOST_F64 = 1, PQ_HOST_C64 = 2, PQ_HOST_C128 = 3 } pq_host_dtype;

#define PQ_MAX_RANK 64

/* ---- lifetime -------------------------------------------------------------- */

/* InteractiveBackend{T}() -- src/backends/interactive.jl:14-22. */
int pq_create(int device, int dtype, pq_handle** out);
int pq_destroy(pq_handle* h);
const char* pq_last_error(const pq_handle* h);
const char* pq_version(void);

/* ---- the backend interface ---------------------------------------------------- */

/* save_tensor_data(backend, label, data) -- interactive.jl:32-36.
 * Copies `host` (column-major, `rank` extents in `dims`) to the device, converting to
 * the backend dtype (`convert(T, data)`); replaces an existing tensor of that label. */
int pq_save_tensor(pq_handle* h, const char* label, int rank, const int64_t* dims,
                   const void* host, int host_dtype);

/* The same for a whole network's worth of tensors in one call: save_tensor_data is called once
 * per node while a network is built (src/layer3.jl:195,226,277,308) and again on every rank of
 * the sliced flow (examples/dist_slicing_example.jl:22-27) -- O(#gates) uploads of <= 16 elements.
 * `dims_flat` holds the extents of tensor 0, then of tensor 1, ... (sum of ranks entries).  One
 * pinned staging block, one host->device copy, one scatter launch; per-tensor semantics are
 * exactly pq_save_tensor's. */
int pq_save_tensors(pq_handle* h, int n, const char* const* labels, const int* ranks,
                    const int64_t* dims_flat, const void* const* hosts, const int* host_dtypes);

/* size(load_tensor_data(...)) without the copy; PQ_ERR_NOT_FOUND mirrors `nothing`
 * (interactive.jl:44-49).  `dims` must hold PQ_MAX_RANK entries. */
int pq_tensor_info(pq_handle* h, const char* label, int* rank, int64_t* dims);

/* load_tensor_data(backend, label) -- interactive.jl:44-49.  Synchronises the stream
 * and copies the tensor into `host_out` (column-major) converted to `host_dtype`
 * (PQ_HOST_C64 or PQ_HOST_C128). */
int pq_load_tensor(pq_handle* h, const char* label, void* host_out, int host_dtype);

/* contract_tensors(backend, A, A_ncon_indices, B, B_ncon_indices, C)
 * -- interactive.jl:60-75 -> src/layer1.jl:85-92 (TensorOperations.tensorcontract).
 * ncon labels: a value present in both lists is contracted, the rest are open;
 * C's axes are A's open axes in A order followed by B's open axes in B order.
 * Stores C, then deletes A and B. */
int pq_contract(pq_handle* h, const char* A, const int32_t* a_idx, int na,
                const char* B, const int32_t* b_idx, int nb, const char* C);

/* permute_tensor(backend, tensor, axes) -- interactive.jl:111-115, layer1.jl:111-114
 * (`permutedims`): size(out, k) == size(in, axes[k]); axes are 1-based. */
int pq_permute(pq_handle* h, const char* label, const int32_t* axes, int n);

/* reshape_tensor(backend, tensor, groups) -- interactive.jl:97-102: new extent k is
 * the product of the old extents at the (1-based) axis positions of group k.
 * `groups_flat` holds the concatenated groups, `group_sizes[k]` their lengths.
 * Metadata only, no kernel. */
int pq_reshape(pq_handle* h, const char* label, const int32_t* groups_flat,
               const int32_t* group_sizes, int ngroups);

/* view_tensor!(backend, view, node, bond_idx, bond_range) -- interactive.jl:169-172,
 * layer1.jl:191-194: an independent COPY of `src` with axis `axis` (1-based)
 * restricted to the 1-based positions idx[0..nidx); the axis is kept. */
int pq_view(pq_handle* h, const char* view, const char* src, int axis,
            const int32_t* idx, int nidx);

/* decompose_tensor!(backend, tensor, left_positions, right_positions; threshold, max_rank,
 * left_label, right_label) -- interactive.jl:130-152 -> src/layer1.jl:146-184.
 * Permutes `tensor` to [left | right] (1-based axis positions), views it as a matrix, takes
 * its SVD (one-sided Jacobi on the device), keeps the chi singular values with
 * S / norm(S) > max(threshold, sqrt(eps(real(T)))) -- at most `max_rank` when max_rank > 0 --
 * and stores B = U * sqrt(S) under `left_label` with extents (left..., chi) and
 * C = sqrt(S) * V^H under `right_label` with extents (chi, right...); deletes `tensor`.
 * `*chi_out` receives chi.  Synchronous (it returns a value computed on the device). */
int pq_decompose(pq_handle* h, const char* tensor, const int32_t* left_positions, int nleft,
                 const int32_t* right_positions, int nright, double threshold, int max_rank,
                 const char* left_label, const char* right_label, int* chi_out);

/* delete_tensor!(backend, label) -- interactive.jl:159-161; a missing label is OK. */
int pq_delete(pq_handle* h, const char* label);

/* save_output(backend, node, name) -- interactive.jl:84-88: alias, no copy. */
int pq_save_output(pq_handle* h, const char* node, const char* name);

int pq_sync(pq_handle* h);

/* ---- sliced contraction (examples/dist_slicing_example.jl:28-30) ------------------ */

/* dst += src elementwise (dst is created as a copy of src when absent): the local
 * accumulation of slice partials that precedes the single reduction. */
int pq_accumulate(pq_handle* h, const char* dst, const char* src);

/* NCCL communicator over the GPUs of one box; replaces MPI.Init / MPI.Reduce!
 * (dist_slicing_example.jl:5-10,30).  `pq_comm_unique_id` fills 128 bytes on one rank;
 * the caller broadcasts them (torch.distributed / MPI / files) and every rank calls
 * `pq_comm_init`.  `pq_allreduce_sum` is one ncclAllReduce(sum) on the stream. */
int pq_comm_unique_id(void* id128);
int pq_comm_init(pq_handle* h, const void* id128, int rank, int nranks);
int pq_allreduce_sum(pq_handle* h, const char* label);

/* ---- .tl programs: execute_dsl_file on the device (src/layer1.jl:211-315) ---------- */

/* Compiles a PicoQuant DSL command stream (grammar: src/backends/dsl.jl:65-214) against
 * the tensors currently stored in the handle: `tensor <name> <key>` binds <name> to the
 * stored tensor <key> (the handle's store stands in for the HDF5 file) without
 * consuming it.  Shapes are propagated, every intermediate gets a fixed offset in one
 * arena, and the launch sequence is captured into a CUDA graph on first run. */
int pq_program_compile(pq_handle* h, const char* tl_text, pq_program** out);
/* Number of `view` commands (slice parameters) in the program. */
int pq_program_num_views(const pq_program* p);
/* Runs the program.  `view_starts` (may be NULL) overrides, per `view` command in
 * stream order, the 1-based start of its index range -- how one compiled plan is
 * replayed for every slice.  `save <t> <file> <key>` stores <t> under <key> in the
 * handle; when `accumulate_into` is non-NULL the saved tensor is also added to it. */
int pq_program_run(pq_handle* h, pq_program* p, const int32_t* view_starts, int nviews,
                   const char* accumulate_into);
/* The slice loop of examples/dist_slicing_example.jl:14-28 (one rank's share of it) in one
 * call: for s in [0, nslices) run the program with view_starts[s * nviews ...] and add the
 * saved result to `accumulate_into`, in slice order -- the sum is bit-identical to nslices
 * pq_program_run calls.  Up to `nlanes` (1..8) slices are kept in flight on private
 * arenas / streams / graph instances; accumulations are ordered on the handle's stream. */
int pq_program_run_slices(pq_handle* h, pq_program* p, const int32_t* view_starts, int nslices,
                          int nviews, const char* accumulate_into, int nlanes);
int pq_program_destroy(pq_handle* h, pq_program* p);
/* Slice-invariant hoisting.  Steps of a program whose inputs do not depend on any `view`
 * (slice) parameter compute the same tensors for every slice of a sliced contraction.  With
 * hoisting enabled, `pq_program_prepare` executes that invariant part once (call it once per
 * amplitude, and again whenever the bound tensors are re-saved) and `pq_program_run` replays
 * only the slice-dependent part.  Off by default: every run then executes the whole stream,
 * exactly like one execute_dsl_file call per slice in the reference flow. */
int pq_program_set_hoist(pq_program* p, int on);
int pq_program_prepare(pq_handle* h, pq_program* p);
int pq_program_hoist_stats(const pq_program* p, int64_t* macs_invariant, int64_t* macs_dependent,
                           int64_t* launches_invariant, int64_t* launches_dependent);
/* arena bytes, number of kernel launches per run, complex MACs per run (data extents) */
int pq_program_stats(const pq_program* p, int64_t* arena_bytes, int64_t* launches,
                     int64_t* macs);

/* ---- instrumentation ----------------------------------------------------------------- */

/* Metrics measured on DATA extents (unlike the graph-side Metrics of backends.jl:8-57,
 * which go stale after slicing): #contract calls, complex MACs, largest tensor. */
int pq_get_counters(pq_handle* h, int64_t* n_contract, int64_t* macs, int64_t* max_elems,
                    int64_t* kernel_launches);
int pq_reset_counters(pq_handle* h);

/* Per-kernel-class CUDA-event timing on the handle's stream (eager mode only).
 * Classes: see pq_kernel_class_name.  Each record accumulates device time, launch
 * count, algorithmic bytes and real flops. */
#define PQ_NUM_KERNEL_CLASSES 15
int pq_profile_enable(pq_handle* h, int on);
int pq_profile_read(pq_handle* h, double* ms, int64_t* launches, double* bytes,
                    double* flops); /* arrays of PQ_NUM_KERNEL_CLASSES */
const char* pq_kernel_class_name(int cls);

/* The same per-class figures measured INSIDE the graph replays of pq_program_run_slices
 * (same lanes, same parallel branches as the timed run): a profiling instance of every
 * (lane, round) graph carries event-record nodes around each kernel node.  Per class:
 * `busy_ms` = length of the union of its kernels' [start, end] intervals (concurrent kernels
 * are not counted twice, so busy_ms <= wall_ms), `sum_ms` = plain sum of durations, launches,
 * algorithmic bytes and real flops; `wall_ms` = first kernel start to last kernel end.
 * nslices <= 16 * nlanes.  Results are NOT accumulated anywhere (measurement only). */
int pq_program_profile_slices(pq_handle* h, pq_program* p, const int32_t* view_starts, int nslices,
                              int nviews, int nlanes, double* busy_ms, double* sum_ms,
                              int64_t* launches, double* bytes, double* flops, double* wall_ms);

/* Device-side stopwatch on the handle's stream (CUDA events): `pq_timer_begin` records
 * the start event; `pq_timer_end` records the stop event, waits for it and returns the
 * elapsed milliseconds.  This is how bench.py times the stream the kernels run on. */
int pq_timer_begin(pq_handle* h);
int pq_timer_end(pq_handle* h, double* ms);

/* Behaviour knobs for A/B checks (0 is always the default / automatic choice):
 *   "gemm"          1 SIMT GEMM, 2 tensor-core GEMM, 3 direct kernel everywhere
 *   "permute"       1 generic permute kernel, 2 tiled bit-permutation
 *   "fused"         1 disables the fused small-operand / dot kernels and the fused-TTGT GEMMs
 *   "graph"         1 eager replay of programs, 2 single-chain CUDA graph (default: DAG graph)
 *   "chain"         1 one launch per tiny contraction in programs (default: batched chains)
 *   "prio"          1 no launch priorities on graph nodes
 *   "zgemm_cfg"     fused ZGEMM tile configuration: 1 64x64, 2 64x32, 3 128x8, 4 64x32 with 3M
 *                   products, 7 persistent skinny kernel where eligible
 *   "zgemm_skinny"  1 disables the persistent skinny ZGEMM
 *   "zgemm_thin"    ComplexF64 steps with N <= 16 and K <= 64: default k_zgemm_thin (DMMA fragments
 *                   loaded straight from the un-permuted operand) when a contracted axis is the
 *                   fastest axis of A, else the 128x8 tiled kernel; the same kernel takes K <= 8
 *                   with 16 < N <= 64 (output-bound).  1 never, 2 / 3 always on N <= 16 with
 *                   k-first / rows-first loads, 4 as default plus K <= 16 on the wide shapes
 *   "zgemm_3m"      1 four DMMAs per complex product instead of three (3M)
 *   "ozaki_auto"    default 1: GEMM-shaped steps with K <= 64, N <= 64 and M >= 4096 run on the INT8
 *                   tensor-core kernel k_ozaki_t (tcgen05.mma kind::i8, Ozaki-scheme slicing into
 *                   int8 digits, exact int32 accumulation; rel-L2 ~3e-13 per ComplexF64 and ~3e-8
 *                   per ComplexF32 contraction) where it is the faster kernel: ComplexF64 with
 *                   K >= 32 and N >= 32, ComplexF32 with N > 16 or K > 16.  0 keeps those steps
 *                   on DMMA / K1 + tcgen05 3xTF32
 *   "zgemm_ozaki"   6 forces every eligible ComplexF64 step (K <= 64, N <= 64) onto k_ozaki_t
 *   "cgemm_ozaki"   4 forces every eligible ComplexF32 step onto k_ozaki_t (gather fused, no K1 pass)
 *   "ozaki_tsw"     2: the ComplexF64 k_ozaki_t keeps digit planes 0..3 of its resident operand in tensor
 *                   memory (TS form of tcgen05.mma) instead of double-buffering two accumulator groups;
 *                   bit-identical results, measured equal or slower, kept for A/B runs
 *   "zgemm_kfirst"  1 row-first gather order only
 *   "zgemm_stagger" ns of start delay per resident-CTA slot in the first wave (tile-per-CTA ZGEMM)
 * Every alternative computes the same contraction; the tests run them against each other. */
int pq_set_option(pq_handle* h, const char* key, int value);

/* Stand-alone micro-benchmarks used by bench.py for roofline denominators measured on
 * the same box: device-to-device copy GB/s and the FP64 DMMA / FP64 FMA issue peaks. */
int pq_microbench(pq_handle* h, const char* what, double* result);

#ifdef __cplusplus
}
#endif
#endif /* PQ_B200_H */
