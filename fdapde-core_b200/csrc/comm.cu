// Multi-GPU plumbing of the solve: one process per GPU, each owning a block of rows of the CSR matrix
// (SURVEY.md section 8e).  The reference has no distributed code at all (single process, single thread); this is
// the B200 side only.  NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a torchrun rank this is the
// copy torch.distributed already loaded), so the library has no link-time NCCL dependency and single-GPU use
// never touches it.
//
//   halo exchange : before every SpMV the owned entries other ranks need are packed and sent with
//                   ncclSend/ncclRecv in one group; received values land directly in the halo tail of the vector.
//   dot products  : per-block partials are collapsed to one value per rank and summed with ncclAllReduce(fp64).
#include <dlfcn.h>

#include <cstring>

#include "solve_common.cuh"

namespace fdb {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclDouble = 8 };  // ncclRedOp_t / ncclDataType_t values, stable across NCCL 2.x

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.handle) return FDB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    FDB_CHECK(h != nullptr, FDB_ERR_UNSUPPORTED, std::string("cannot load libnccl.so.2: ") + dlerror());
#define FDB_SYM(field, name)                                                         \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));         \
    FDB_CHECK(g_nccl.field != nullptr, FDB_ERR_UNSUPPORTED, std::string("libnccl lacks ") + name)
    FDB_SYM(GetUniqueId, "ncclGetUniqueId");
    FDB_SYM(CommInitRank, "ncclCommInitRank");
    FDB_SYM(CommDestroy, "ncclCommDestroy");
    FDB_SYM(Send, "ncclSend");
    FDB_SYM(Recv, "ncclRecv");
    FDB_SYM(AllReduce, "ncclAllReduce");
    FDB_SYM(GroupStart, "ncclGroupStart");
    FDB_SYM(GroupEnd, "ncclGroupEnd");
    FDB_SYM(GetErrorString, "ncclGetErrorString");
#undef FDB_SYM
    g_nccl.handle = h;
    return FDB_OK;
}

#define FDB_NCCL(call)                                                                                   \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != 0) {                                                                                   \
            set_error(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
            return FDB_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

}  // namespace fdb

struct fdb_comm {
    fdb::NcclComm comm = nullptr;
    int rank = 0, world = 1;
};

namespace fdb {

__global__ void k_pack(int n, const int32_t* __restrict__ idx, const double* __restrict__ v, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v[idx[i]];
}

int halo_exchange(fdb_matrix* A, double* vec) {
    Partition* P = A->part;
    if (!P || P->nbr.empty()) return FDB_OK;
    cudaStream_t st = A->space->stream;
    if (P->n_send > 0) {
        k_pack<<<(P->n_send + 255) / 256, 256, 0, st>>>(P->n_send, P->send_idx.p, vec, P->sendbuf.p);
        FDB_CUDA(cudaGetLastError());
    }
    FDB_NCCL(g_nccl.GroupStart());
    for (size_t i = 0; i < P->nbr.size(); ++i) {
        const int sc = P->send_off[i + 1] - P->send_off[i], rc = P->recv_off[i + 1] - P->recv_off[i];
        if (sc > 0) FDB_NCCL(g_nccl.Send(P->sendbuf.p + P->send_off[i], (size_t)sc, kNcclDouble, P->nbr[i], P->comm->comm, st));
        if (rc > 0)
            FDB_NCCL(g_nccl.Recv(vec + P->n_owned + P->recv_off[i], (size_t)rc, kNcclDouble, P->nbr[i], P->comm->comm, st));
    }
    FDB_NCCL(g_nccl.GroupEnd());
    return FDB_OK;
}

int allreduce_sum(fdb_matrix* A, const double* in, double* out, int count) {
    Partition* P = A->part;
    FDB_NCCL(g_nccl.AllReduce(in, out, (size_t)count, kNcclDouble, kNcclSum, P->comm->comm, A->space->stream));
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" {

int fdb_comm_unique_id(void* id128) {
    FDB_CHECK(id128, FDB_ERR_ARG, "null argument");
    FDB_TRY(load_nccl());
    NcclUniqueId id;
    FDB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return FDB_OK;
}

int fdb_comm_create(fdb_comm** out, int rank, int world_size, const void* id128) {
    FDB_CHECK(out && id128 && world_size >= 1 && rank >= 0 && rank < world_size, FDB_ERR_ARG, "bad argument");
    FDB_TRY(load_nccl());
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    fdb_comm* c = new fdb_comm();
    c->rank = rank;
    c->world = world_size;
    int r = g_nccl.CommInitRank(&c->comm, world_size, id, rank);
    if (r != 0) {
        set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
        delete c;
        return FDB_ERR_CUDA;
    }
    *out = c;
    return FDB_OK;
}

void fdb_comm_destroy(fdb_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

int fdb_matrix_set_partition(fdb_matrix* A, fdb_comm* comm, int n_owned, int n_neighbors, const int32_t* neighbor_ranks,
                             const int32_t* send_counts, const int32_t* send_idx, const int32_t* recv_counts) {
    FDB_CHECK(A && A->space && comm, FDB_ERR_ARG, "null argument");
    FDB_CHECK(n_owned >= 0 && n_owned <= A->space->n_dofs && n_neighbors >= 0, FDB_ERR_ARG, "bad partition sizes");
    FDB_CHECK(n_neighbors == 0 || (neighbor_ranks && send_counts && recv_counts), FDB_ERR_ARG, "null neighbour arrays");
    delete A->part;
    Partition* P = new Partition();
    A->part = P;
    P->comm = comm;
    P->n_owned = n_owned;
    P->send_off.assign(1, 0);
    P->recv_off.assign(1, 0);
    for (int i = 0; i < n_neighbors; ++i) {
        FDB_CHECK(neighbor_ranks[i] >= 0 && neighbor_ranks[i] < comm->world && neighbor_ranks[i] != comm->rank, FDB_ERR_ARG,
                  "bad neighbour rank");
        P->nbr.push_back(neighbor_ranks[i]);
        P->send_off.push_back(P->send_off.back() + send_counts[i]);
        P->recv_off.push_back(P->recv_off.back() + recv_counts[i]);
    }
    P->n_send = P->send_off.back();
    P->n_halo = P->recv_off.back();
    FDB_CHECK(n_owned + P->n_halo <= A->space->n_dofs, FDB_ERR_ARG, "owned + halo dofs exceed the local space");
    if (P->n_send > 0) {
        FDB_CHECK(send_idx != nullptr, FDB_ERR_ARG, "null send index list");
        for (int i = 0; i < P->n_send; ++i) FDB_CHECK(send_idx[i] >= 0 && send_idx[i] < n_owned, FDB_ERR_ARG, "send index is not an owned dof");
        FDB_TRY(P->send_idx.alloc(P->n_send));
        FDB_TRY(P->sendbuf.alloc(P->n_send));
        FDB_CUDA(cudaMemcpy(P->send_idx.p, send_idx, sizeof(int32_t) * P->n_send, cudaMemcpyHostToDevice));
    }
    FDB_TRY(P->stage.alloc(32));
    FDB_CUDA(cudaMemset(P->stage.p, 0, sizeof(double) * 32));
    {   // global number of rows: every rank must run the same default iteration budget (collective call)
        const double mine = (double)n_owned;
        double all = 0;
        cudaStream_t st = A->space->stream;
        FDB_CUDA(cudaMemcpyAsync(P->stage.p, &mine, sizeof(double), cudaMemcpyHostToDevice, st));
        FDB_TRY(allreduce_sum(A, P->stage.p, P->stage.p + 16, 1));
        FDB_CUDA(cudaMemcpyAsync(&all, P->stage.p + 16, sizeof(double), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        FDB_CUDA(cudaMemsetAsync(P->stage.p, 0, sizeof(double) * 32, st));
        P->n_global = (long long)(all + 0.5);
    }
    return FDB_OK;
}

int fdb_matrix_peer_export(fdb_matrix* A, void* ipc_handle64) {
    FDB_CHECK(A && A->part && ipc_handle64, FDB_ERR_ARG, "fdb_matrix_set_partition must be called first");
    Partition* P = A->part;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!P->peer_buf) {
        const PeerLayout L = PeerLayout::of((size_t)P->n_halo, P->comm->world);
        FDB_CUDA(cudaMalloc(&P->peer_buf, L.bytes));  // a dedicated allocation: IPC handles map whole allocations
        FDB_CUDA(cudaMemset(P->peer_buf, 0, L.bytes));
    }
    cudaIpcMemHandle_t h;
    FDB_CUDA(cudaIpcGetMemHandle(&h, P->peer_buf));
    memcpy(ipc_handle64, &h, sizeof(h));
    return FDB_OK;
}

int fdb_matrix_peer_connect(fdb_matrix* A, const void* handles64, const int64_t* peer_n_halo,
                            const int32_t* nbr_recv_offset) {
    FDB_CHECK(A && A->part && A->part->peer_buf && handles64 && peer_n_halo, FDB_ERR_ARG,
              "fdb_matrix_peer_export must be called first");
    Partition* P = A->part;
    const int world = P->comm->world, rank = P->comm->rank;
    const int n_nbr = (int)P->nbr.size();
    FDB_CHECK(world <= 8 && n_nbr <= 8, FDB_ERR_UNSUPPORTED, "peer-memory plan supports up to 8 ranks");
    FDB_CHECK(n_nbr == 0 || nbr_recv_offset, FDB_ERR_ARG, "null neighbour array");
    FDB_CHECK(peer_n_halo[rank] == P->n_halo, FDB_ERR_ARG, "peer_n_halo[rank] differs from this rank's halo size");
    std::vector<char*> base(world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == rank) { base[r] = static_cast<char*>(P->peer_buf); continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles64) + 64 * (size_t)r, sizeof(h));
        void* ptr = nullptr;
        FDB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        P->peer_mapped.push_back(ptr);
        base[r] = static_cast<char*>(ptr);
    }
    PeerView* pv = new PeerView();
    memset(pv, 0, sizeof(*pv));
    pv->world = world; pv->rank = rank; pv->n_nbr = n_nbr; pv->n_owned = P->n_owned; pv->n_halo = P->n_halo;
    pv->send_idx = P->send_idx.p;
    for (int i = 0; i <= n_nbr; ++i) pv->send_off[i] = P->send_off[i];
    for (int r = 0; r < world; ++r) {
        const PeerLayout L = PeerLayout::of((size_t)peer_n_halo[r], world);
        pv->red_of[r] = reinterpret_cast<LLWord*>(base[r] + L.off_red);
    }
    const PeerLayout Lme = PeerLayout::of((size_t)P->n_halo, world);
    pv->my_red = pv->red_of[rank];
    pv->my_halo = reinterpret_cast<LLWord*>(base[rank] + Lme.off_halo);
    pv->error = reinterpret_cast<int*>(base[rank] + Lme.off_error);
    for (int i = 0; i < n_nbr; ++i) {
        const int q = P->nbr[i];
        FDB_CHECK(nbr_recv_offset[i] >= 0 && nbr_recv_offset[i] + (P->send_off[i + 1] - P->send_off[i]) <= peer_n_halo[q],
                  FDB_ERR_ARG, "bad neighbour receive offset");
        pv->nbr_halo[i] = reinterpret_cast<LLWord*>(base[q]) + nbr_recv_offset[i];
        pv->nbr_n_halo[i] = (long long)peer_n_halo[q];
    }
    delete static_cast<PeerView*>(P->peer_view);
    P->peer_view = pv;
    P->peer_ready = true;
    return FDB_OK;
}

}  // extern "C"

namespace fdb {
Partition::~Partition() {
    cudaDeviceSynchronize();
    for (void* m : peer_mapped) cudaIpcCloseMemHandle(m);
    if (peer_buf) cudaFree(peer_buf);
    delete static_cast<PeerView*>(peer_view);
}
}  // namespace fdb
