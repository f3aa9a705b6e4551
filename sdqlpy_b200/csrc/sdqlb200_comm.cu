// libsdqlb200_comm.so -- cross-GPU merge of partial dictionaries (include/sdqlb200_comm.h).
//
// Replaces the reference's merge of thread-local partial results: tbb::parallel_reduce's join
// (sdql_ir_cpp_generator_par.py:258-291) and the serial AddMap of thread-local phmap tables (map_helper.h:2-23,
// sdql_ir_cpp_generator_par.py:436-438).  A "thread" is a GPU here; the partials never leave the devices.
//
//   small partials   k_p2p_allreduce: one CTA per rank.  Every rank stores its words into slot [rank] of EVERY peer's
//                    mailbox (plain stores into NVLink peer memory), publishes a flag (release.sys), waits for the W flags
//                    in its own mailbox (acquire.sys) and reduces the W slots in rank order.  Two mailbox buffers alternate
//                    by call parity: a rank can run at most one call ahead of the slowest peer, because call e+1 cannot
//                    complete without every peer's flag for e+1, which a peer raises only after it has finished e.
//   large partials   ncclAllReduce on the query's stream.
//   hashed tables    count -> all-gather of the counts -> pack by destination -> grouped ncclSend/ncclRecv -> combine in a
//                    scratch table -> broadcast of the combined runs -> write back (SURVEY.md 8e "hash all-to-all").
//
// NCCL is bound with dlopen/dlsym at first use: the process usually carries torch's libnccl.so.2 already.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "sdqlb200_comm.h"
#include "sdqlb200_host.h"

using sdqlrt::i64;
using sdqlrt::u64;

namespace {

constexpr int kMaxWorld = 8;
constexpr int kP2PWords = SDQLB200_COMM_P2P_MAX_WORDS;

#define COMM_FAIL(...) sdqlhost::fail(SDQLB200_E_ARG, __VA_ARGS__)

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time
// ---------------------------------------------------------------------------------------------
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* (*GetLastError)(ncclComm_t) = nullptr;
};
Nccl g_nccl;

int nccl_bind() {
    if (g_nccl.h) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already carries (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return COMM_FAIL("libnccl.so.2 not found: %s", dlerror());
#define BIND(name)                                                                                   \
    *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name);                                                \
    if (!g_nccl.name) return COMM_FAIL("libnccl: symbol nccl" #name " missing")
    BIND(GetUniqueId); BIND(CommInitRank); BIND(CommInitAll); BIND(CommDestroy); BIND(AllReduce); BIND(AllGather);
    BIND(Broadcast); BIND(Send); BIND(Recv); BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
    *(void**)(&g_nccl.GetLastError) = dlsym(h, "ncclGetLastError");
    g_nccl.h = h;
    return 0;
}

#define SDQL_NCCL(x)                                                                                                 \
    do {                                                                                                             \
        ncclResult_t r_ = (x);                                                                                       \
        if (r_ != ncclSuccess)                                                                                       \
            return sdqlhost::fail(SDQLB200_E_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #x, g_nccl.GetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// mailboxes
// ---------------------------------------------------------------------------------------------
struct Mailbox {
    u64 flag[2][kMaxWorld];            // flag[b][r]: last call number whose words rank r has stored into buffer b
    u64 pad[16];
    u64 data[2][kMaxWorld][kP2PWords];
};
struct Peers { Mailbox* p[kMaxWorld]; };

}  // namespace

struct sdqlb200_comm {
    int rank, world, dev;
    ncclComm_t nccl;
    Mailbox* local;         // cudaMalloc'ed on this device
    Peers peers;            // every rank's mailbox as seen from this device (peers.p[rank] == local)
    bool p2p, ipc_opened[kMaxWorld];
    u64 call;               // number of p2p all-reduces issued so far (the same on every rank)
    volatile u64* h_err;    // pinned, mapped: set by a kernel that gave up waiting for a peer
    u64* d_err;
    u64* d_scratch;         // 4 KB of device scratch (counts, small staging)
    u64* h_scratch;         // pinned staging of the same size
};

namespace {

__device__ __forceinline__ void st_release_sys(u64* p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// op: SDQLB200_SUM_F64 | SUM_I64 | MIN_I32 (two int32 per word)
__global__ void __launch_bounds__(1024) k_p2p_allreduce(const __grid_constant__ Peers P, int rank, int world, u64* buf, int nwords, int op, u64 call, u64* err) {
    const int b = (int)(call & 1);
    for (int t = threadIdx.x; t < nwords; t += blockDim.x) {
        const u64 v = buf[t];
        for (int d = 0; d < world; ++d) P.p[d]->data[b][rank][t] = v;  // NVLink peer stores (d == rank: local)
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(&P.p[threadIdx.x]->flag[b][rank], call);
        const u64* f = &P.p[rank]->flag[b][threadIdx.x];
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < call) {
            if (clock64() - t0 > 6000000000ll) { *err = call; break; }  // ~3 s: a peer never arrived; do not hang the GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    const Mailbox* m = P.p[rank];
    for (int t = threadIdx.x; t < nwords; t += blockDim.x) {
        u64 acc = __ldcg(&m->data[b][0][t]);  // L2 is the point of coherence for the peers' stores
        for (int r = 1; r < world; ++r) {
            const u64 v = __ldcg(&m->data[b][r][t]);
            if (op == SDQLB200_SUM_F64) acc = (u64)__double_as_longlong(__longlong_as_double((i64)acc) + __longlong_as_double((i64)v));
            else if (op == SDQLB200_SUM_I64) acc += v;
            else {
                const int lo = min((int)(unsigned)acc, (int)(unsigned)v), hi = min((int)(acc >> 32), (int)(v >> 32));
                acc = ((u64)(unsigned)hi << 32) | (u64)(unsigned)lo;
            }
        }
        buf[t] = acc;
    }
}

int comm_alloc(sdqlb200_comm* c) {
    SDQL_CUDA(cudaMalloc((void**)&c->local, sizeof(Mailbox)));
    SDQL_CUDA(cudaMemset(c->local, 0, sizeof(Mailbox)));
    SDQL_CUDA(cudaMalloc((void**)&c->d_scratch, 4096));
    SDQL_CUDA(cudaHostAlloc((void**)&c->h_scratch, 4096, cudaHostAllocDefault));
    SDQL_CUDA(cudaHostAlloc((void**)&c->h_err, 64, cudaHostAllocMapped));
    *c->h_err = 0;
    SDQL_CUDA(cudaHostGetDevicePointer((void**)&c->d_err, (void*)c->h_err, 0));
    cudaMemPool_t pool;  // scratch of the table merges comes from the stream-ordered allocator: keep freed blocks cached
    SDQL_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->dev));
    unsigned long long keep = ~0ull;
    SDQL_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    SDQL_CUDA(cudaDeviceSynchronize());
    for (int r = 0; r < kMaxWorld; ++r) { c->peers.p[r] = nullptr; c->ipc_opened[r] = false; }
    c->peers.p[c->rank] = c->local;
    c->p2p = c->world == 1;
    c->call = 0;
    return 0;
}

}  // namespace

extern "C" {

const char* sdqlb200_comm_last_error(void) { return sdqlhost::g_err; }

int sdqlb200_comm_unique_id(void* out128) {
    if (int rc = nccl_bind()) return rc;
    static_assert(sizeof(ncclUniqueId) == SDQLB200_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    SDQL_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return 0;
}

int sdqlb200_comm_create(const void* id128, int32_t rank, int32_t world, sdqlb200_comm** out) {
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return COMM_FAIL("comm_create: rank %d of %d", rank, world);
    if (int rc = nccl_bind()) return rc;
    sdqlb200_comm* c = new sdqlb200_comm();
    c->rank = rank; c->world = world; c->nccl = nullptr;
    SDQL_CUDA(cudaGetDevice(&c->dev));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    SDQL_NCCL(g_nccl.CommInitRank(&c->nccl, world, id, rank));
    if (int rc = comm_alloc(c)) return rc;
    *out = c;
    return 0;
}

int sdqlb200_comm_ipc_handle(sdqlb200_comm* c, void* out64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == SDQLB200_COMM_IPC_BYTES, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t h;
    SDQL_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(out64, &h, sizeof h);
    return 0;
}

int sdqlb200_comm_open_peers(sdqlb200_comm* c, const void* handles) {
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof h, sizeof h);
        void* p = nullptr;
        SDQL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peers.p[r] = (Mailbox*)p;
        c->ipc_opened[r] = true;
    }
    c->p2p = true;
    return 0;
}

int sdqlb200_comm_create_all(int32_t n, const int32_t* devices, sdqlb200_comm** out) {
    if (n < 1 || n > kMaxWorld) return COMM_FAIL("comm_create_all: %d devices (1..%d)", n, kMaxWorld);
    if (int rc = nccl_bind()) return rc;
    int prev = 0;
    SDQL_CUDA(cudaGetDevice(&prev));
    ncclComm_t comms[kMaxWorld];
    int devs[kMaxWorld];
    for (int i = 0; i < n; ++i) devs[i] = devices[i];
    SDQL_NCCL(g_nccl.CommInitAll(comms, n, devs));
    bool all_peer = true;
    for (int i = 0; i < n; ++i) {
        sdqlb200_comm* c = new sdqlb200_comm();
        c->rank = i; c->world = n; c->dev = devs[i]; c->nccl = comms[i];
        SDQL_CUDA(cudaSetDevice(devs[i]));
        if (int rc = comm_alloc(c)) return rc;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            int can = 0;
            SDQL_CUDA(cudaDeviceCanAccessPeer(&can, devs[i], devs[j]));
            if (!can) { all_peer = false; continue; }
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return sdqlhost::fail(SDQLB200_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devs[i], devs[j], cudaGetErrorString(e));
        }
        out[i] = c;
    }
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) out[i]->peers.p[j] = out[j]->local;
        out[i]->p2p = all_peer;
    }
    SDQL_CUDA(cudaSetDevice(prev));
    return 0;
}

int sdqlb200_comm_destroy(sdqlb200_comm* c) {
    if (!c) return 0;
    for (int r = 0; r < c->world; ++r)
        if (c->ipc_opened[r]) cudaIpcCloseMemHandle(c->peers.p[r]);
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    cudaFree(c->local);
    cudaFree(c->d_scratch);
    cudaFreeHost(c->h_scratch);
    cudaFreeHost((void*)c->h_err);
    delete c;
    return 0;
}

int sdqlb200_comm_rank(const sdqlb200_comm* c) { return c->rank; }
int sdqlb200_comm_world(const sdqlb200_comm* c) { return c->world; }
int sdqlb200_comm_p2p(const sdqlb200_comm* c) { return c->p2p ? 1 : 0; }

int sdqlb200_comm_allreduce(sdqlb200_comm* c, void* d_buf, uint64_t count, int32_t op, void* stream) {
    if (op != SDQLB200_SUM_F64 && op != SDQLB200_SUM_I64 && op != SDQLB200_MIN_I32) return COMM_FAIL("comm_allreduce: op %d", op);
    if (*c->h_err) return sdqlhost::fail(SDQLB200_E_CUDA, "comm: rank %d gave up waiting for a peer in all-reduce #%llu", c->rank, (unsigned long long)*c->h_err);
    if (count == 0 || c->world == 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const u64 words = op == SDQLB200_MIN_I32 ? (count + 1) / 2 : count;
    if (c->p2p && words <= (u64)kP2PWords && (op != SDQLB200_MIN_I32 || count % 2 == 0)) {
        const u64 call = ++c->call;
        const int threads = words <= 256 ? 256 : words <= 512 ? 512 : 1024;
        k_p2p_allreduce<<<1, threads, 0, st>>>(c->peers, c->rank, c->world, (u64*)d_buf, (int)words, op, call, c->d_err);
        SDQL_CUDA(cudaGetLastError());
        return 1;  // > 0: took the peer-memory path (callers count it)
    }
    if (op == SDQLB200_SUM_F64) SDQL_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, c->nccl, st));
    else if (op == SDQLB200_SUM_I64) SDQL_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclInt64, ncclSum, c->nccl, st));
    else SDQL_NCCL(g_nccl.AllReduce(d_buf, d_buf, count, ncclInt32, ncclMin, c->nccl, st));
    return 0;
}

int sdqlb200_comm_host_max(sdqlb200_comm* c, int64_t* h_values, int32_t n, void* stream) {
    if (n < 0 || n > 256) return COMM_FAIL("comm_host_max: %d values (<= 256)", n);
    if (c->world == 1 || n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    memcpy(c->h_scratch, h_values, (size_t)n * 8);
    SDQL_CUDA(cudaMemcpyAsync(c->d_scratch, c->h_scratch, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    SDQL_NCCL(g_nccl.AllReduce(c->d_scratch, c->d_scratch, (size_t)n, ncclInt64, ncclMax, c->nccl, st));
    SDQL_CUDA(cudaMemcpyAsync(c->h_scratch, c->d_scratch, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    memcpy(h_values, c->h_scratch, (size_t)n * 8);
    return 0;
}

int sdqlb200_comm_barrier(sdqlb200_comm* c, void* stream) {
    int64_t v = 0;
    if (int rc = sdqlb200_comm_host_max(c, &v, 1, stream)) return rc;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// hashed partial dictionary: hash all-to-all + combine + broadcast + write back
// ---------------------------------------------------------------------------------------------
static int tbl_io(const sdqlb200_table* t, sdqlrt::TblIO* io) {
    if (!t || !t->keys || t->cap < 1 || (t->cap & (t->cap - 1)) || t->nfields < 0 || t->nfields > 16)
        return COMM_FAIL("comm_merge_table: bad table descriptor");
    memset(io, 0, sizeof *io);
    io->keys = (u64*)t->keys; io->rep = t->rep; io->cap = t->cap; io->nf = t->nfields; io->f64_mask = t->f64_mask;
    for (int j = 0; j < t->nfields; ++j) io->agg[j] = (u64*)t->agg[j];
    return 0;
}

int sdqlb200_comm_merge_table(sdqlb200_comm* c, const sdqlb200_table* t, void* stream) {
    sdqlrt::TblIO io;
    if (int rc = tbl_io(t, &io)) return rc;
    if (c->world == 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int W = c->world, R = 2 + io.nf, sms = 148;
    u64* d_cnt = c->d_scratch;                 // [W] my records per destination
    u64* d_all = c->d_scratch + 16;            // [W x W] all ranks' counts
    u64* d_off = c->d_scratch + 16 + 64;       // [W] send offsets
    u64* d_cur = c->d_scratch + 16 + 64 + 16;  // [W] pack cursors
    u64* d_m = c->d_scratch + 16 + 64 + 32;    // [W] combined entries per rank
    SDQL_CUDA(cudaMemsetAsync(c->d_scratch, 0, 4096, st));
    const int g = sdqlhost::grid_for(io.cap, 8, sms);
    sdqlrt::k_tbl_count<<<g, sdqlrt::kBlock, 0, st>>>(io, W, d_cnt);
    SDQL_NCCL(g_nccl.AllGather(d_cnt, d_all, (size_t)W, ncclUint64, c->nccl, st));
    SDQL_CUDA(cudaMemcpyAsync(c->h_scratch, d_all, (size_t)W * W * 8, cudaMemcpyDeviceToHost, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    u64 cnt[kMaxWorld], rcnt[kMaxWorld], off[kMaxWorld], roff[kMaxWorld], nsend = 0, nrecv = 0;
    for (int d = 0; d < W; ++d) {
        cnt[d] = c->h_scratch[c->rank * W + d];
        rcnt[d] = c->h_scratch[d * W + c->rank];
        off[d] = nsend; roff[d] = nrecv;
        nsend += cnt[d]; nrecv += rcnt[d];
    }
    i64 *d_send = nullptr, *d_recv = nullptr;
    SDQL_CUDA(cudaMallocAsync((void**)&d_send, (nsend + 1) * R * 8, st));
    SDQL_CUDA(cudaMallocAsync((void**)&d_recv, (nrecv + 1) * R * 8, st));
    memcpy(c->h_scratch, off, sizeof off);
    SDQL_CUDA(cudaMemcpyAsync(d_off, c->h_scratch, (size_t)W * 8, cudaMemcpyHostToDevice, st));
    sdqlrt::k_tbl_pack<<<g, sdqlrt::kBlock, 0, st>>>(io, W, c->rank, nullptr, d_off, d_cur, d_send);
    SDQL_NCCL(g_nccl.GroupStart());
    for (int d = 0; d < W; ++d) {
        if (cnt[d]) SDQL_NCCL(g_nccl.Send(d_send + off[d] * R, cnt[d] * R, ncclInt64, d, c->nccl, st));
        if (rcnt[d]) SDQL_NCCL(g_nccl.Recv(d_recv + roff[d] * R, rcnt[d] * R, ncclInt64, d, c->nccl, st));
    }
    SDQL_NCCL(g_nccl.GroupEnd());
    // combine what arrived in a scratch table: fields summed, owner = the lowest rank that saw the key
    i64 cap2 = 1024;
    while (cap2 < 2 * (i64)nrecv) cap2 <<= 1;
    sdqlrt::TblIO t2 = io;
    int* own2 = nullptr;
    char* blk = nullptr;
    const size_t b_keys = (size_t)cap2 * 8, b_own = (size_t)cap2 * 4, b_agg = (size_t)cap2 * 8 * (io.nf ? io.nf : 1);
    SDQL_CUDA(cudaMallocAsync((void**)&blk, b_keys + b_own + b_agg, st));
    t2.keys = (u64*)blk; own2 = (int*)(blk + b_keys); t2.rep = own2; t2.cap = cap2;
    for (int j = 0; j < io.nf; ++j) t2.agg[j] = (u64*)(blk + b_keys + b_own) + (size_t)j * cap2;
    SDQL_CUDA(cudaMemsetAsync(blk, 0xFF, b_keys, st));
    SDQL_CUDA(cudaMemsetAsync(own2, 0x7f, b_own, st));  // 0x7f7f7f7f: larger than any rank
    SDQL_CUDA(cudaMemsetAsync(blk + b_keys + b_own, 0, b_agg, st));
    if (nrecv) sdqlrt::k_tbl_absorb<<<sdqlhost::grid_for((i64)nrecv, 8, sms), sdqlrt::kBlock, 0, st>>>(t2, d_recv, (i64)nrecv, 0, c->rank, own2);
    // my combined entries, packed; their number goes to everybody
    const int g2 = sdqlhost::grid_for(cap2, 8, sms);
    SDQL_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
    sdqlrt::k_tbl_count<<<g2, sdqlrt::kBlock, 0, st>>>(t2, 1, d_cnt);
    SDQL_NCCL(g_nccl.AllGather(d_cnt, d_m, 1, ncclUint64, c->nccl, st));
    SDQL_CUDA(cudaMemcpyAsync(c->h_scratch, d_m, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    u64 m[kMaxWorld], moff[kMaxWorld], total = 0;
    for (int r = 0; r < W; ++r) { m[r] = c->h_scratch[r]; moff[r] = total; total += m[r]; }
    int rc = 0;
    if (2 * total > (u64)io.cap)  // known to every rank alike: all ranks fail together, nobody waits in a collective
        rc = COMM_FAIL("merged dictionary has %llu entries, the table was sized for %lld slots", (unsigned long long)total, (long long)io.cap);
    i64* d_allrec = nullptr;
    if (!rc) {
        SDQL_CUDA(cudaMallocAsync((void**)&d_allrec, (total + 1) * R * 8, st));
        SDQL_CUDA(cudaMemsetAsync(d_off, 0, 16, st));
        if (m[c->rank]) {
            SDQL_CUDA(cudaMemsetAsync(d_cur, 0, 8, st));
            sdqlrt::k_tbl_pack<<<g2, sdqlrt::kBlock, 0, st>>>(t2, 1, c->rank, own2, d_off, d_cur, d_allrec + moff[c->rank] * R);
        }
        SDQL_NCCL(g_nccl.GroupStart());
        for (int r = 0; r < W; ++r)
            if (m[r]) SDQL_NCCL(g_nccl.Broadcast(d_allrec + moff[r] * R, d_allrec + moff[r] * R, m[r] * R, ncclInt64, r, c->nccl, st));
        SDQL_NCCL(g_nccl.GroupEnd());
        if (total) sdqlrt::k_tbl_absorb<<<sdqlhost::grid_for((i64)total, 8, sms), sdqlrt::kBlock, 0, st>>>(io, d_allrec, (i64)total, 1, c->rank, nullptr);
        SDQL_CUDA(cudaGetLastError());
        SDQL_CUDA(cudaFreeAsync(d_allrec, st));
    }
    SDQL_CUDA(cudaFreeAsync(blk, st));
    SDQL_CUDA(cudaFreeAsync(d_send, st));
    SDQL_CUDA(cudaFreeAsync(d_recv, st));
    return rc;
}

// SDQLB200_MERGE_DIRECT: -> 0 merged sparsely, 1 declined (dense: the caller all-reduces the arrays), < 0 error
int sdqlb200_comm_merge_direct(sdqlb200_comm* c, const sdqlb200_table* t, void* stream) {
    if (!t || t->keys || !t->rep || t->cap < 1 || t->nfields < 0 || t->nfields > 16) return COMM_FAIL("comm_merge_direct: bad table descriptor");
    if (c->world == 1) return 0;
    sdqlrt::TblIO io;
    memset(&io, 0, sizeof io);
    io.rep = t->rep; io.cap = t->cap; io.nf = t->nfields; io.f64_mask = t->f64_mask;
    for (int j = 0; j < t->nfields; ++j) io.agg[j] = (u64*)t->agg[j];
    cudaStream_t st = (cudaStream_t)stream;
    const int W = c->world, R = 2 + io.nf, sms = 148;
    u64* d_cnt = c->d_scratch;        // [1] my occupied slots
    u64* d_all = c->d_scratch + 16;   // [W]
    u64* d_cur = c->d_scratch + 32;   // pack cursor
    SDQL_CUDA(cudaMemsetAsync(c->d_scratch, 0, 512, st));
    const int g = sdqlhost::grid_for(io.cap, 8, sms);
    sdqlrt::k_dtbl_count<<<g, sdqlrt::kBlock, 0, st>>>(io, d_cnt);
    SDQL_NCCL(g_nccl.AllGather(d_cnt, d_all, 1, ncclUint64, c->nccl, st));
    SDQL_CUDA(cudaMemcpyAsync(c->h_scratch, d_all, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    u64 m[kMaxWorld], moff[kMaxWorld], total = 0;
    for (int r = 0; r < W; ++r) { m[r] = c->h_scratch[r]; moff[r] = total; total += m[r]; }
    // the same numbers on every rank: all decline or none does.  Dense moves the whole arrays through an all-reduce
    // (rep as int32 MIN + nf 8-byte SUMs), sparse the occupied slots of every rank to every rank + one atomic per field
    const double dense = (double)io.cap * (4.0 + 8.0 * io.nf), sparse = (double)total * 8.0 * R;
    if (2.0 * sparse >= dense) return 1;
    i64* d_allrec = nullptr;
    SDQL_CUDA(cudaMallocAsync((void**)&d_allrec, (total + 1) * R * 8, st));
    if (m[c->rank]) sdqlrt::k_dtbl_pack<<<g, sdqlrt::kBlock, 0, st>>>(io, c->rank, d_cur, d_allrec + moff[c->rank] * R);
    SDQL_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < W; ++r)
        if (m[r]) SDQL_NCCL(g_nccl.Broadcast(d_allrec + moff[r] * R, d_allrec + moff[r] * R, m[r] * R, ncclInt64, r, c->nccl, st));
    SDQL_NCCL(g_nccl.GroupEnd());
    if (total) sdqlrt::k_dtbl_absorb<<<sdqlhost::grid_for((i64)total, 8, sms), sdqlrt::kBlock, 0, st>>>(
        io, d_allrec, (i64)total, (i64)moff[c->rank], (i64)(moff[c->rank] + m[c->rank]), c->rank);
    SDQL_CUDA(cudaGetLastError());
    SDQL_CUDA(cudaFreeAsync(d_allrec, st));
    return 0;
}

int sdqlb200_comm_gather_rows(sdqlb200_comm* c, const int64_t* const* cols, int32_t nfields, int64_t count,
                              int64_t** h_out, int64_t* h_total, void* stream) {
    if (nfields < 0 || nfields > 32 || count < 0) return COMM_FAIL("comm_gather_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int W = c->world;
    u64* d_cnt = c->d_scratch;
    u64* d_all = c->d_scratch + 16;
    c->h_scratch[0] = (u64)count;
    SDQL_CUDA(cudaMemcpyAsync(d_cnt, c->h_scratch, 8, cudaMemcpyHostToDevice, st));
    SDQL_NCCL(g_nccl.AllGather(d_cnt, d_all, 1, ncclUint64, c->nccl, st));
    SDQL_CUDA(cudaMemcpyAsync(c->h_scratch, d_all, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    u64 n[kMaxWorld], off[kMaxWorld], total = 0;
    for (int r = 0; r < W; ++r) { n[r] = c->h_scratch[r]; off[r] = total; total += n[r]; }
    *h_total = (int64_t)total;
    for (int j = 0; j < nfields; ++j) h_out[j] = nullptr;
    if (!total || !nfields) return 0;
    i64* d_buf = nullptr;
    SDQL_CUDA(cudaMallocAsync((void**)&d_buf, total * (size_t)nfields * 8, st));
    for (int j = 0; j < nfields; ++j)
        if (count) SDQL_CUDA(cudaMemcpyAsync(d_buf + (size_t)j * total + off[c->rank], cols[j], (size_t)count * 8, cudaMemcpyDefault, st));  // host or device source (UVA)
    SDQL_NCCL(g_nccl.GroupStart());
    for (int j = 0; j < nfields; ++j)
        for (int r = 0; r < W; ++r)
            if (n[r]) {
                i64* p = d_buf + (size_t)j * total + off[r];
                SDQL_NCCL(g_nccl.Broadcast(p, p, n[r], ncclInt64, r, c->nccl, st));
            }
    SDQL_NCCL(g_nccl.GroupEnd());
    for (int j = 0; j < nfields; ++j) {
        h_out[j] = (int64_t*)malloc(total * 8);
        SDQL_CUDA(cudaMemcpyAsync(h_out[j], d_buf + (size_t)j * total, total * 8, cudaMemcpyDeviceToHost, st));
    }
    SDQL_CUDA(cudaFreeAsync(d_buf, st));
    SDQL_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int sdqlb200_comm_merge(void* vctx, uint64_t workspace_offset, uint64_t count, int32_t op) {
    sdqlb200_comm_ctx* x = (sdqlb200_comm_ctx*)vctx;
    if (!x || !x->comm) return COMM_FAIL("comm_merge: no communicator");
    int rc;
    if (op == SDQLB200_MERGE_TABLE) {
        rc = sdqlb200_comm_merge_table(x->comm, (const sdqlb200_table*)(uintptr_t)workspace_offset, x->stream);
        x->table_merges += rc == 0;
    } else if (op == SDQLB200_MERGE_DIRECT) {
        rc = sdqlb200_comm_merge_direct(x->comm, (const sdqlb200_table*)(uintptr_t)workspace_offset, x->stream);
        if (rc == 1) return 1;  // declined: the module all-reduces the arrays (counted there)
        x->table_merges += rc == 0;
    } else {
        rc = sdqlb200_comm_allreduce(x->comm, (char*)x->workspace + workspace_offset, count, op, x->stream);
        if (rc > 0) { x->p2p_merges += 1; rc = 0; }
    }
    x->merges += rc == 0;
    return rc;
}

}  // extern "C"
