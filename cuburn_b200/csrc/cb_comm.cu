// Multi-GPU exchange steps of the render path behind the C ABI: the histogram
// (all-)reduce of a still split over GPUs and the gather of filtered row bands.
//
// The reference has no multi-GPU data path (its job farm ships whole frames between
// processes, distribute.py:131-248); these entry points are what SURVEY 8(b)/(e) ask
// of a drop-in: `hist_reduce(comm, root)` on the caller's stream.  NCCL is bound at
// run time (dlopen of libnccl.so.2), so the library still loads on a box without it.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

// NCCL is bound with dlopen at run time, so its header is not needed to build: when the
// development package is absent the handful of ABI types used here (stable since NCCL 2.0)
// are declared locally.
#if defined(__has_include) && __has_include(<nccl.h>)
#include <nccl.h>
#else
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
typedef enum { ncclFloat32 = 7, ncclFloat = 7 } ncclDataType_t;
}
#endif

#include "cb_common.h"

namespace {

struct nccl_api {
    void *lib = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int,
                           ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t,
                              ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t,
                         cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
nccl_api g_nccl;

template <typename F>
bool bind(F &fn, const char *name) {
    fn = reinterpret_cast<F>(dlsym(g_nccl.lib, name));
    return fn != nullptr;
}

int load_nccl() {
    if (g_nccl.lib) return CB_OK;
    // One process must not end up with two NCCLs under the same soname: take the copy
    // that is already loaded (PyTorch's, if the host imported it first), else the one
    // named by CUBURN_B200_NCCL, else the system's.
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char *path = getenv("CUBURN_B200_NCCL");
    if (!lib && path && *path) lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        cb_set_error("NCCL is not available: %s", dlerror());
        return CB_ERR_INVALID;
    }
    g_nccl.lib = lib;
    bool ok = bind(g_nccl.GetVersion, "ncclGetVersion") &&
              bind(g_nccl.GetUniqueId, "ncclGetUniqueId") &&
              bind(g_nccl.CommInitRank, "ncclCommInitRank") &&
              bind(g_nccl.CommDestroy, "ncclCommDestroy") &&
              bind(g_nccl.Reduce, "ncclReduce") && bind(g_nccl.AllReduce, "ncclAllReduce") &&
              bind(g_nccl.Send, "ncclSend") && bind(g_nccl.Recv, "ncclRecv") &&
              bind(g_nccl.GroupStart, "ncclGroupStart") &&
              bind(g_nccl.GroupEnd, "ncclGroupEnd") &&
              bind(g_nccl.GetErrorString, "ncclGetErrorString");
    if (!ok) {
        g_nccl.lib = nullptr;
        cb_set_error("libnccl lacks a required symbol");
        return CB_ERR_INVALID;
    }
    return CB_OK;
}

}  // namespace

#define CB_NCCL(expr)                                                              \
    do {                                                                           \
        ncclResult_t r__ = (expr);                                                 \
        if (r__ != ncclSuccess) {                                                  \
            cb_set_error("%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), \
                         __FILE__, __LINE__);                                      \
            return CB_ERR_CUDA;                                                    \
        }                                                                          \
    } while (0)

struct cb_comm_s {
    ncclComm_t comm;
    int rank, world;
};

static_assert(sizeof(ncclUniqueId) == CB_COMM_ID_BYTES, "unique id size");

extern "C" {

int cb_comm_version(int *version) {
    CB_REQUIRE(version, "null argument");
    int rc = load_nccl();
    if (rc != CB_OK) return rc;
    CB_NCCL(g_nccl.GetVersion(version));
    return CB_OK;
}

int cb_comm_unique_id(uint8_t id[CB_COMM_ID_BYTES]) {
    CB_REQUIRE(id, "null argument");
    int rc = load_nccl();
    if (rc != CB_OK) return rc;
    ncclUniqueId uid;
    CB_NCCL(g_nccl.GetUniqueId(&uid));
    memcpy(id, &uid, CB_COMM_ID_BYTES);
    return CB_OK;
}

int cb_comm_create(const uint8_t id[CB_COMM_ID_BYTES], int rank, int world, cb_comm *comm) {
    CB_REQUIRE(id && comm, "null argument");
    CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank outside the job");
    int rc = load_nccl();
    if (rc != CB_OK) return rc;
    ncclUniqueId uid;
    memcpy(&uid, id, CB_COMM_ID_BYTES);
    ncclComm_t c;
    CB_NCCL(g_nccl.CommInitRank(&c, world, uid, rank));
    cb_comm_s *out = new cb_comm_s;
    out->comm = c;
    out->rank = rank;
    out->world = world;
    *comm = out;
    return CB_OK;
}

int cb_comm_destroy(cb_comm comm) {
    if (!comm) return CB_OK;
    ncclResult_t r = g_nccl.CommDestroy(comm->comm);
    delete comm;
    if (r != ncclSuccess) {
        cb_set_error("ncclCommDestroy failed: %s", g_nccl.GetErrorString(r));
        return CB_ERR_CUDA;
    }
    return CB_OK;
}

int cb_hist_reduce(cb_comm comm, cb_dptr hist4, const cb_dims *dim, int root, cb_stream s) {
    CB_REQUIRE(comm && hist4 && dim, "null argument");
    CB_REQUIRE(root >= -1 && root < comm->world, "root outside the job");
    const size_t count = 4 * (size_t)dim->aheight * (size_t)dim->astride;
    float *p = cb_ptr<float>(hist4);
    if (root < 0)
        CB_NCCL(g_nccl.AllReduce(p, p, count, ncclFloat, ncclSum, comm->comm, cb_cs(s)));
    else
        CB_NCCL(g_nccl.Reduce(p, p, count, ncclFloat, ncclSum, root, comm->comm, cb_cs(s)));
    return CB_OK;
}

int cb_band_rows(int aheight, int rank, int world, int *row0, int *row1) {
    CB_REQUIRE(row0 && row1, "null argument");
    CB_REQUIRE(aheight > 0 && aheight % 16 == 0, "aheight must come from cb_calc_dim");
    CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank outside the job");
    const int blocks = aheight / 16;
    int band = 16 * ((blocks + world - 1) / world);
    if (band > aheight) band = aheight;
    int r0 = rank * band;
    if (r0 > aheight - band) r0 = aheight - band;
    *row0 = r0;
    *row1 = r0 + band;
    return CB_OK;
}

int cb_band_gather(cb_comm comm, cb_dptr frame4, const cb_dims *dim, int root, cb_stream s) {
    CB_REQUIRE(comm && frame4 && dim, "null argument");
    CB_REQUIRE(root >= 0 && root < comm->world, "root outside the job");
    if (comm->world == 1) return CB_OK;
    const size_t row_floats = 4 * (size_t)dim->astride;
    float *p = cb_ptr<float>(frame4);
    CB_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < comm->world; r++) {
        if (r == root) continue;
        if (comm->rank != root && comm->rank != r) continue;
        int r0, r1;
        int rc = cb_band_rows(dim->aheight, r, comm->world, &r0, &r1);
        if (rc != CB_OK) {
            g_nccl.GroupEnd();
            return rc;
        }
        float *band = p + (size_t)r0 * row_floats;
        const size_t count = (size_t)(r1 - r0) * row_floats;
        ncclResult_t res = comm->rank == root
            ? g_nccl.Recv(band, count, ncclFloat, r, comm->comm, cb_cs(s))
            : g_nccl.Send(band, count, ncclFloat, root, comm->comm, cb_cs(s));
        if (res != ncclSuccess) {
            g_nccl.GroupEnd();
            cb_set_error("band exchange with rank %d failed: %s", r, g_nccl.GetErrorString(res));
            return CB_ERR_CUDA;
        }
    }
    CB_NCCL(g_nccl.GroupEnd());
    return CB_OK;
}

}  // extern "C"
