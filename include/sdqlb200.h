/* sdqlpy-b200 C ABI -- the boundary every generated query module (<script>_compiled.so) exports.
 *
 * What it replaces in the reference (edin-dal/sdqlpy):
 *   - the generated CPython method  <fn>_compiled(db)        sdql_compiler.py:601-672, 749-777
 *     (numpy column pointers in: PyArray_DATA casts           sdql_compiler.py:644-668)
 *   - the result boxing into a FastDict / PyFloat / PyLong    sdql_ir_cpp_generator_par.py:866-877
 * The reference has no plain C ABI (its C-API function *is* the ABI); this header is
 * the interface a maintainer binds instead (ctypes stub in INTEGRATION.md; sdqlpy_b200/runtime.py is that stub).
 *
 * Conventions: plain pointers and sizes only, no exceptions cross the boundary, every function returns 0 on
 * success and a negative SDQLB200_E_* code otherwise (text via sdqlb200_last_error()).
 */
#ifndef SDQLB200_H
#define SDQLB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SDQLB200_I32 = 0, SDQLB200_F64 = 1, SDQLB200_CODE = 2, SDQLB200_BYTES = 3 };
enum { SDQLB200_OK = 0, SDQLB200_E_WORKSPACE = -1, SDQLB200_E_CUDA = -2, SDQLB200_E_ARG = -3, SDQLB200_E_NOQUERY = -4 };
enum { SDQLB200_F_NOFETCH = 1, SDQLB200_F_KERNEL_TIMES = 2, SDQLB200_F_TRACE = 4 /* per-step wall-clock times on stderr */ };
enum { SDQLB200_COL_PARTKEY = 1 };                       /* sdqlb200_col.flags: the relation is range partitioned on this column */
enum { SDQLB200_SUM_F64 = 0, SDQLB200_SUM_I64 = 1, SDQLB200_MIN_I32 = 2, SDQLB200_MERGE_TABLE = 3, SDQLB200_MERGE_DIRECT = 4 }; /* merge ops */
/* multi-GPU: called (stream ordered) after a kernel over a partitioned relation for every partial buffer that has to
 * be combined across ranks.  SUM / MIN: the callee all-reduces `count` elements at workspace + offset in place (NCCL).
 * MERGE_DIRECT: `workspace_offset` carries a HOST pointer to a sdqlb200_table with keys == NULL: a direct-indexed partial
 * dictionary of `cap` slots (rep[slot] >= 0: present on this rank).  The callee may merge it SPARSELY -- the ranks exchange
 * their occupied slots, every rank adds the others' fields into its own arrays and marks the entries a lower rank also holds
 * (or that it does not hold itself) rep = -2 -- and return 0, or decline with +1 when the table is dense enough that
 * all-reducing the whole arrays is cheaper (the module then does that with MIN_I32 + SUM); the decision is taken from
 * all-gathered counts, alike on every rank.  Q17's 20 M-slot table holds 20 K parts: 400 MB per dense merge.
 * MERGE_TABLE: `workspace_offset` carries a HOST pointer to a sdqlb200_table describing a hashed partial dictionary;
 * the callee merges it across ranks with the sdqlb200_table_* helpers below (hash all-to-all, combine at the
 * destination, all-gather, write back).  The reference's counterpart is the serial AddMap merge of thread-local
 * phmap tables (map_helper.h:2-23, sdql_ir_cpp_generator_par.py:436-438). */
typedef int (*sdqlb200_merge_fn)(void* ctx, uint64_t workspace_offset, uint64_t count, int32_t op);
/* The product implementation is sdqlb200_comm_merge of include/sdqlb200_comm.h (NVLink peer-memory all-reduce for small
 * partials, NCCL for large ones, hash all-to-all for hashed tables; plain C, no host synchronisation for SUM / MIN). */

/* one hashed device dictionary: open addressing, linear probing, keys[slot] == ~0 means free */
typedef struct {
    uint64_t* keys;     /* DEVICE, cap slots                                                         */
    int32_t* rep;       /* DEVICE, cap slots: >= 0 owned here, -1 free, -2 present but owned elsewhere */
    int64_t cap;        /* power of two                                                              */
    int32_t nfields;    /* aggregate fields (<= 16), one 8-byte array of cap slots each              */
    uint32_t f64_mask;  /* bit j set: field j is fp64, else int64                                    */
    void* agg[16];      /* DEVICE                                                                    */
} sdqlb200_table;

/* one device-resident column (replaces the borrowed numpy buffer of sdql_compiler.py:653-668) */
typedef struct {
    const void* data; /* DEVICE pointer, 16-byte aligned                                         */
    int64_t rows;
    int64_t min, max; /* value range of int/date columns; [0, dictionary size - 1] for codes    */
    int32_t width;    /* element bytes: I32 4, F64 8, CODE 1 or 4, BYTES = fixed string width   */
    int32_t kind;     /* SDQLB200_I32 | F64 | CODE | BYTES                                       */
    int32_t flags;    /* SDQLB200_COL_*                                                          */
    int32_t stride;   /* 0 = no statement.  (log2 B << 16) | K: every value v of this int column satisfies
                       * (v - min) mod B < K (B a power of two) -- dbgen order keys use 8 of every 32.  Tables keyed by such a
                       * column pack the holes away: slot (d / B) * K + d mod B, d = v - min (4x fewer slots for order keys) */
} sdqlb200_col;

/* result rows in SoA form; every field is one 8-byte slot per row (int64 / fp64 bits / string reference) */
typedef struct {
    int64_t count;
    int32_t nfields;
    int32_t reserved;
    int64_t* cols[32]; /* HOST buffers owned by the library until sdqlb200_result_free */
} sdqlb200_result;

typedef struct {
    const sdqlb200_col* cols; /* in the order of the query manifest's "inputs"                 */
    int32_t ncols;
    int32_t nargs;
    const int64_t* nrows;     /* rows of each relation argument, in query-argument order        */
    const int64_t* consts;    /* resolved constants, in the order of the manifest's "consts"   */
    int32_t nconsts;
    int32_t flags;            /* SDQLB200_F_*                                                   */
    void* workspace;          /* DEVICE scratch (tables, partials, result buffers)              */
    uint64_t workspace_bytes;
    uint64_t workspace_needed; /* out: set when SDQLB200_E_WORKSPACE is returned                */
    void* stream;             /* cudaStream_t                                                   */
    float device_ms;          /* out: CUDA-event time of all kernels + memsets of the query     */
    int32_t launches;         /* out: kernels launched                                          */
    int32_t tier;             /* out: aggregation tier of the last group-by kernel (0/1/2)      */
    int32_t reserved;
    sdqlb200_result result;   /* out                                                            */
    float kernel_ms[24];      /* out (SDQLB200_F_KERNEL_TIMES): CUDA-event time of each launch   */
    sdqlb200_merge_fn merge;  /* NULL on a single GPU                                           */
    void* merge_ctx;
    uint32_t part_mask;       /* bit i set: relation argument i holds only this rank's partition */
    int32_t result_partial;   /* out: 1 = result rows are this rank's share (concatenate ranks) */
    int32_t rank;             /* this process' rank among the GPUs (0 on a single GPU)          */
    int32_t world;            /* number of ranks (0 or 1 on a single GPU)                       */
    const int64_t* nrows_global; /* multi-GPU: rows of each relation argument over ALL ranks (NULL on a single GPU).
                               * Tables that are merged across ranks are planned (direct vs hashed, slots) from these, so
                               * every rank takes the same decisions and a hashed table holds the union of the ranks' keys */
} sdqlb200_args;

int sdqlb200_num_queries(void);
const char* sdqlb200_query_name(int i);
/* JSON: per query its arguments, required device inputs (arg, column, representation), constants to resolve,
 * result schema and algorithmic bytes per scanned row */
const char* sdqlb200_manifest(void);
/* run one query (replaces <fn>_compiled(db)).  Call with workspace == NULL to obtain workspace_needed. */
int sdqlb200_run(const char* query, sdqlb200_args* args);
void sdqlb200_result_free(sdqlb200_result* r);
const char* sdqlb200_last_error(void);

/* Diagnostics (no reference counterpart): counters of the data-dependent memory operations executed since the previous
 * call -- the inputs of the "bytes-moved" roofline of join-heavy queries.  out[0..7]: presence-bit tests, table probes,
 * slots touched by probes, insert-or-find operations, slots they touched, global aggregate atomics, gather loads,
 * (unused); out[8]: bytes of table arrays initialised (counted in every build).  Returns 1 when the module is a
 * counting build (-DSDQLB200_STATS; such a build is never timed), 0 when out[0..7] are not collected, < 0 on error. */
int sdqlb200_stats(uint64_t* out, int32_t n);

/* building blocks of SDQLB200_MERGE_TABLE (all stream ordered; records are 2 + nfields 8-byte words:
 * key, owner rank, field values).  Records destined to rank d = hash(key) mod world.
 *   count : d_counts[d] += occupied slots destined to d                      (d_counts: world zeroed uint64)
 *   pack  : writes the records grouped by destination; d_offsets[d] = first record of d's run, d_cursor zeroed;
 *           owner word = rank, or d_own[slot] when d_own != NULL
 *   absorb: mode 0 -- insert / add the records into t, d_own[slot] = min(owner)  (combine at the destination)
 *           mode 1 -- insert / overwrite fields; slots whose owner != rank (or that are new here) get rep = -2 */
int sdqlb200_table_count(const sdqlb200_table* t, int32_t world, uint64_t* d_counts, void* stream);
int sdqlb200_table_pack(const sdqlb200_table* t, int32_t world, int32_t rank, const int32_t* d_own,
                        const uint64_t* d_offsets, uint64_t* d_cursor, int64_t* d_records, void* stream);
int sdqlb200_table_absorb(const sdqlb200_table* t, const int64_t* d_records, int64_t n, int32_t mode, int32_t rank,
                          int32_t* d_own, void* stream);

#ifdef __cplusplus
}
#endif
#endif
