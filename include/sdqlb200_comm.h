/* sdqlpy-b200 multi-GPU exchange library (libsdqlb200_comm.so) -- what combines the ranks' partial dictionaries.
 *
 * What it replaces in the reference (edin-dal/sdqlpy): the merge of the thread-local partial results of a parallel
 * sum -- tbb::parallel_reduce's join (sdql_ir_cpp_generator_par.py:258-291) for scalars / records, and the SERIAL
 * key-wise AddMap of thread-local phmap tables (map_helper.h:2-23, sdql_ir_cpp_generator_par.py:436-438) for
 * dictionaries.  Here a "thread" is a GPU and the merge runs on the devices:
 *   - small partials (scalars, records, direct-indexed tables of a few thousand words): ONE kernel per rank that stores
 *     the rank's partial into every peer's mailbox over NVLink peer memory, raises a flag, waits for the peers' flags and
 *     sums the W partials in rank order -- no NCCL launch, no host involvement, bit-identical sums on every rank;
 *   - larger dense partials: ncclAllReduce enqueued on the query's stream;
 *   - hashed partial dictionaries: hash all-to-all (grouped ncclSend / ncclRecv), combine at the destination,
 *     broadcast of the combined runs, write back.
 * A generated query module calls all of this through ONE C function pointer, sdqlb200_comm_merge (the
 * sdqlb200_merge_fn of include/sdqlb200.h): no Python, no GIL, nothing but stream-ordered work in the steady state.
 *
 * Conventions as in sdqlb200.h: plain pointers and sizes, 0 on success / negative SDQLB200_E_* otherwise, text through
 * sdqlb200_comm_last_error().  NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries --
 * torch's -- or the system one), so the library loads on a machine without it; every entry point that needs it fails
 * loudly.  There is no CPU path.
 */
#ifndef SDQLB200_COMM_H
#define SDQLB200_COMM_H
#include <stdint.h>

#include "sdqlb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdqlb200_comm sdqlb200_comm; /* one per (process, GPU) */

enum { SDQLB200_COMM_ID_BYTES = 128, SDQLB200_COMM_IPC_BYTES = 64 };
/* largest all-reduce (8-byte words) that takes the one-kernel peer-memory path */
enum { SDQLB200_COMM_P2P_MAX_WORDS = 4096 };

/* ---- bootstrap: one process per GPU (torchrun).  Rank 0 makes the id, the host side broadcasts it (plumbing), every
 * rank creates its communicator (collective), then the ranks exchange the IPC handles of their mailboxes (64 bytes
 * each, all-gathered by the host side) and map the peers'.  Without open_peers every all-reduce goes through NCCL. */
int sdqlb200_comm_unique_id(void* out128);
int sdqlb200_comm_create(const void* id128, int32_t rank, int32_t world, sdqlb200_comm** out);
int sdqlb200_comm_ipc_handle(sdqlb200_comm* c, void* out64);
int sdqlb200_comm_open_peers(sdqlb200_comm* c, const void* handles /* world x 64 bytes, rank order */);
/* ---- bootstrap: ONE process driving n GPUs (sdqlpy_init(mode, n) -- the reference's threads_count, sdql_lib.py:372):
 * n communicators, mailboxes mapped through plain peer access.  out[i] belongs to devices[i] and to the host thread
 * that drives it. */
int sdqlb200_comm_create_all(int32_t n, const int32_t* devices, sdqlb200_comm** out);
int sdqlb200_comm_destroy(sdqlb200_comm* c);
int sdqlb200_comm_rank(const sdqlb200_comm* c);
int sdqlb200_comm_world(const sdqlb200_comm* c);
int sdqlb200_comm_p2p(const sdqlb200_comm* c); /* 1 = peers' mailboxes are mapped */

/* ---- data path (stream ordered) */
/* in-place all-reduce of `count` 8-byte (SUM_F64, SUM_I64) or 4-byte (MIN_I32) elements of device memory.
 * Returns 1 (not an error) when the call took the one-kernel peer-memory path, 0 when it went through NCCL. */
int sdqlb200_comm_allreduce(sdqlb200_comm* c, void* d_buf, uint64_t count, int32_t op, void* stream);
/* SDQLB200_MERGE_TABLE: all-reduce of a hashed partial dictionary.  Synchronises the stream twice (run lengths).
 * Fails on EVERY rank alike when the union of the keys does not fit t->cap / 2. */
int sdqlb200_comm_merge_table(sdqlb200_comm* c, const sdqlb200_table* t, void* stream);
/* SDQLB200_MERGE_DIRECT: sparse merge of a direct-indexed partial dictionary (t->keys == NULL): 0 = merged, 1 = declined
 * (the table is dense enough for an all-reduce of its arrays, alike on every rank), < 0 = error.  Synchronises the stream once. */
int sdqlb200_comm_merge_direct(sdqlb200_comm* c, const sdqlb200_table* t, void* stream);
/* concatenation of the ranks' result rows: every rank contributes `count` rows of `nfields` 8-byte columns (column j at
 * cols[j], in host OR device memory -- the copy direction is inferred from the pointer); on return h_total = rows of all ranks and h_out[j] (malloc'ed here, caller frees) holds
 * column j of all ranks in rank order.  Synchronises the stream. */
int sdqlb200_comm_gather_rows(sdqlb200_comm* c, const int64_t* const* cols, int32_t nfields, int64_t count,
                              int64_t** h_out, int64_t* h_total, void* stream);
/* host-side helpers the bootstrap needs (small; synchronise): max over ranks of n int64 values, in place (HOST buffer) */
int sdqlb200_comm_host_max(sdqlb200_comm* c, int64_t* h_values, int32_t n, void* stream);
int sdqlb200_comm_barrier(sdqlb200_comm* c, void* stream);

/* the sdqlb200_merge_fn to put into sdqlb200_args.merge; sdqlb200_args.merge_ctx = a sdqlb200_comm_ctx */
typedef struct {
    sdqlb200_comm* comm;
    void* workspace; /* = sdqlb200_args.workspace: merge offsets are relative to it */
    void* stream;
    int64_t merges, table_merges, p2p_merges; /* out: counters */
} sdqlb200_comm_ctx;
int sdqlb200_comm_merge(void* ctx, uint64_t workspace_offset, uint64_t count, int32_t op);

const char* sdqlb200_comm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
