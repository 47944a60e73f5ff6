// Multi-GPU plumbing of the Krylov solver: halo exchange of ghost-element coefficients and all-reduce of dot products
// over NCCL (NVLink 5 / NVSwitch).  The reference is a single-process program (no counterpart); SURVEY.md section 8(e).
//
// NCCL is NOT linked: its entry points are resolved at run time from the library the process has already loaded
// (PyTorch's bundled libnccl.so.2), or from an explicit path, so that exactly one NCCL lives in the process.
// The exported functions reyna_b200_nccl_halo / reyna_b200_nccl_allreduce have the rb_halo_fn / rb_allreduce_fn
// signatures of include/reyna_b200.h; `user` is the rb_dist handle.
#include "common.cuh"
#include <dlfcn.h>

namespace rb {

// the slice of the NCCL API used here (ABI-stable since NCCL 2.7)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl(const char* path) {
    if (g_nccl.handle) return 0;
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load NCCL: %s", dlerror()); return RB_E_ARG; }
    NcclApi a;
    a.handle = h;
#define RB_SYM(field, name)                                                             \
    *(void**)(&a.field) = dlsym(h, name);                                               \
    if (!a.field) { set_error("NCCL symbol %s not found", name); return RB_E_ARG; }
    RB_SYM(GetUniqueId, "ncclGetUniqueId")
    RB_SYM(CommInitRank, "ncclCommInitRank")
    RB_SYM(CommDestroy, "ncclCommDestroy")
    RB_SYM(GroupStart, "ncclGroupStart")
    RB_SYM(GroupEnd, "ncclGroupEnd")
    RB_SYM(Send, "ncclSend")
    RB_SYM(Recv, "ncclRecv")
    RB_SYM(AllReduce, "ncclAllReduce")
    RB_SYM(Broadcast, "ncclBroadcast")
    RB_SYM(GetErrorString, "ncclGetErrorString")
#undef RB_SYM
    g_nccl = a;
    return 0;
}

#define RB_NCCL(call)                                                                   \
    do {                                                                                \
        ncclResult_t _r = (call);                                                       \
        if (_r != 0) { set_error("%s: %s", #call, g_nccl.GetErrorString(_r)); return 1000 + _r; } \
    } while (0)

// pack[i*d + k] = x[idx[i]*d + k]
__global__ void halo_pack(int n_send, int d, const int32_t* __restrict__ idx, const double* __restrict__ x,
                          double* __restrict__ pack) {
    const int total = n_send * d;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int i = t / d, k = t - i * d;
        pack[t] = x[(size_t)idx[i] * d + k];
    }
}

}  // namespace rb

struct rb_dist {
    rb::ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int d = 1, n_rows = 0;                 // coefficients per element, owned elements
    int n_peers = 0;
    int peer[64];
    int send_off[65], recv_off[65];        // element offsets per peer (send list / ghost block)
    const int32_t* send_idx = nullptr;     // device: owned element ids to send, concatenated per peer
    double* pack = nullptr;                // device: send staging (owned by the handle)
    unsigned long long exchanges = 0, allreduces = 0;
};

using namespace rb;

extern "C" int reyna_b200_nccl_unique_id(const char* nccl_path, char out[128]) {
    int rc = load_nccl(nccl_path);
    if (rc) return rc;
    ncclUniqueId id;
    RB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return 0;
}

// One NCCL communicator per process and world (creation costs 0.3 ... several seconds): created once, shared by every
// halo plan / solve.
extern "C" int reyna_b200_comm_create(const char* nccl_path, const char id[128], int rank, int world, void** comm_out) {
    RB_REQUIRE(id && comm_out, RB_E_ARG, "reyna_b200_comm_create: bad argument");
    int rc = load_nccl(nccl_path);
    if (rc) return rc;
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    ncclComm_t comm = nullptr;
    RB_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
    *comm_out = (void*)comm;
    return 0;
}

extern "C" void reyna_b200_comm_destroy(void* comm) {
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)comm);
}

// plan arrays are HOST arrays except send_idx (device); offsets are in elements; ghosts received from peer[i] occupy the
// ghost block [recv_off[i], recv_off[i+1]) (ghosts are sorted by global id, peers own contiguous id ranges).
extern "C" int reyna_b200_dist_create(void* comm, int rank, int world, int d, int n_rows, int n_peers, const int* peer,
                                      const int* send_off, const int* recv_off, const int32_t* send_idx_dev,
                                      rb_dist** out) {
    RB_REQUIRE(comm && out && n_peers >= 0 && n_peers <= 64, RB_E_ARG, "reyna_b200_dist_create: bad argument");
    rb_dist* h = new rb_dist();
    h->comm = (ncclComm_t)comm;
    h->rank = rank; h->world = world; h->d = d; h->n_rows = n_rows; h->n_peers = n_peers;
    for (int i = 0; i < n_peers; ++i) h->peer[i] = peer[i];
    for (int i = 0; i <= n_peers; ++i) { h->send_off[i] = send_off[i]; h->recv_off[i] = recv_off[i]; }
    h->send_idx = send_idx_dev;
    const size_t n_send = n_peers ? (size_t)send_off[n_peers] : 0;
    if (cudaMalloc((void**)&h->pack, sizeof(double) * (n_send * d + 1)) != cudaSuccess) {
        set_error("reyna_b200_dist_create: cudaMalloc failed");
        delete h;
        return RB_E_WORKSPACE;
    }
    *out = h;
    return 0;
}

extern "C" void reyna_b200_dist_destroy(rb_dist* h) {
    if (!h) return;
    if (h->pack) cudaFree(h->pack);
    delete h;
}

extern "C" int reyna_b200_dist_counters(const rb_dist* h, unsigned long long* exchanges, unsigned long long* allreduces) {
    RB_REQUIRE(h, RB_E_ARG, "reyna_b200_dist_counters: null handle");
    if (exchanges) *exchanges = h->exchanges;
    if (allreduces) *allreduces = h->allreduces;
    return 0;
}

namespace rb {

// refresh the ghost block of x holding `d` coefficients per element (d <= the handle's d) -- one grouped send/recv
int dist_halo_d(rb_dist* h, double* x, int d, cudaStream_t st) {
    if (!h || h->n_peers == 0) return 0;
    const int n_send = h->send_off[h->n_peers];
    if (n_send > 0) {
        halo_pack<<<grid_for((int64_t)n_send * d, 256, kNumSMs * 4), 256, 0, st>>>(n_send, d, h->send_idx, x, h->pack);
        RB_LAUNCH_CHECK("halo_pack");
    }
    RB_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < h->n_peers; ++i) {
        const int ns = h->send_off[i + 1] - h->send_off[i], nr = h->recv_off[i + 1] - h->recv_off[i];
        if (ns > 0) RB_NCCL(g_nccl.Send(h->pack + (size_t)h->send_off[i] * d, (size_t)ns * d, ncclFloat64, h->peer[i], h->comm, st));
        if (nr > 0)
            RB_NCCL(g_nccl.Recv(x + ((size_t)h->n_rows + h->recv_off[i]) * d, (size_t)nr * d, ncclFloat64, h->peer[i], h->comm, st));
    }
    RB_NCCL(g_nccl.GroupEnd());
    ++h->exchanges;
    return 0;
}

// all ranks obtain the concatenation of every rank's `local` slice (rank r contributes global[bounds[r] .. bounds[r+1]))
int dist_allgather_rows(rb_dist* h, const double* local, double* global, const int32_t* bounds, cudaStream_t st) {
    if (!h) return RB_E_ARG;
    RB_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < h->world; ++r) {
        const size_t cnt = (size_t)(bounds[r + 1] - bounds[r]);
        if (cnt == 0) continue;
        RB_NCCL(g_nccl.Broadcast(r == h->rank ? (const void*)local : (const void*)(global + bounds[r]), global + bounds[r], cnt,
                                 ncclFloat64, r, h->comm, st));
    }
    RB_NCCL(g_nccl.GroupEnd());
    ++h->allreduces;
    return 0;
}

int dist_rank(const rb_dist* h) { return h ? h->rank : 0; }
int dist_world(const rb_dist* h) { return h ? h->world : 1; }

}  // namespace rb

// rb_halo_fn: refresh the ghost block of x (owned part [0, n_rows*d), ghosts behind it)
extern "C" int reyna_b200_nccl_halo(void* user, double* x, void* stream) {
    rb_dist* h = (rb_dist*)user;
    if (!h) return 0;
    return dist_halo_d(h, x, h->d, (cudaStream_t)stream);
}

// rb_allreduce_fn: in-place sum of n doubles over all ranks
extern "C" int reyna_b200_nccl_allreduce(void* user, double* vals, int n, void* stream) {
    rb_dist* h = (rb_dist*)user;
    if (!h || h->world == 1) return 0;
    RB_NCCL(g_nccl.AllReduce(vals, vals, (size_t)n, ncclFloat64, ncclSum, h->comm, (cudaStream_t)stream));
    ++h->allreduces;
    return 0;
}
