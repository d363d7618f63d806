// comm.cu — the multi-GPU data plane of the library: the reference's content-based partition scheme
// (LocalitySensitivePartitioner, src/main/scala/mclab/utils/Partitioner.scala:27-64; one sub-index store per partition,
// src/main/java/mclab/mapdb/RandomDrawTreeMap.java:1430-1459; findStepWiseSubIndexIDs :613-621) mapped onto the GPUs of
// one box, one handle (= one rank) per GPU, NCCL over NVLink / NVSwitch between them.
//
//   build   every rank holds the vectors (the re-rank needs them; SURVEY 8e option A) but hashes only its slice of the
//           ids; ONE all-gather of the slices' keys + sub-index ids (5 L bytes per vector) gives every rank all keys,
//           from which it builds the sub-forests it owns.  Hashing, the dominant build stage, scales with the ranks.
//   query   queries are replicated; every rank probes the sub-indexes it owns, re-ranks its candidates and writes its
//           top k straight into its segment of the gather buffer; ONE in-place all-gather (12 bytes per result entry)
//           and one merge kernel give every rank the global top k.  No host synchronisation in between.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded),
// so a single-GPU user of the library needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace dpf {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (api.lib) return api;
    // a copy that is already in the process (PyTorch's, for one) is the one to use: loading a second NCCL of another
    // version under the same soname would break whoever loads theirs later
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
        if (api.lib) break;
    }
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        if (api.lib) break;
        api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    DPF_REQUIRE(api.lib, DPF_ERR_STATE, "libnccl.so.2 not found: the multi-GPU plane needs NCCL");
    auto sym = [&](const char* n) {
        void* p = dlsym(api.lib, n);
        if (!p) { api.lib = nullptr; throw Error{DPF_ERR_STATE, std::string("NCCL symbol missing: ") + n}; }
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}

#define DPF_NCCL(expr)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r__ = (expr);                                                                              \
        if (r__ != ncclSuccess) throw ::dpf::Error{DPF_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(r__)}; \
    } while (0)

static_assert(sizeof(ncclUniqueId) <= DPF_COMM_ID_BYTES, "unique id does not fit the ABI's buffer");

void comm_unique_id(uint8_t* out) {
    ncclUniqueId id;
    DPF_NCCL(nccl().GetUniqueId(&id));
    memset(out, 0, DPF_COMM_ID_BYTES);
    memcpy(out, &id, sizeof(id));
}

void comm_init(dpf_index* h, const uint8_t* id_bytes) {
    DPF_REQUIRE(!h->comm, DPF_ERR_STATE, "communicator already initialised");
    const int world = h->cfg.world > 1 ? h->cfg.world : 1, rank = h->cfg.world > 1 ? h->cfg.rank : 0;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t c = nullptr;
    DPF_NCCL(nccl().CommInitRank(&c, world, id, rank));
    h->comm = c;
    // NCCL connects its channels at the first collective (hundreds of ms): do that here, not inside the first fit
    h->comm_stage.reserve(64 * (size_t)world);
    DPF_NCCL(nccl().AllGather(h->comm_stage.p + 64 * rank, h->comm_stage.p, 64, ncclChar, c, h->stream));
    DPF_CUDA(cudaStreamSynchronize(h->stream));
}

void comm_destroy(dpf_index* h) {
    if (!h->comm) return;
    nccl().CommDestroy(static_cast<ncclComm_t>(h->comm));
    h->comm = nullptr;
}

// ---- build: hash a slice, all-gather the keys ------------------------------------------------------------------------
// segment of rank g: keys of its slice, table-major with leading dimension `slice` (L x slice int32), then the sub-index
// ids the same way (L x slice bytes, padded to 16)
__global__ void __launch_bounds__(256)
k_unpack_gathered_keys(const unsigned char* __restrict__ stage, size_t seg_bytes, int64_t slice, int64_t n, int L, int64_t n_old,
                       int64_t ld, int32_t* __restrict__ keys, uint8_t* __restrict__ pids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // new vector
    const int t = blockIdx.y;
    if (i >= n) return;
    const int64_t g = i / slice, j = i - g * slice;
    const unsigned char* seg = stage + (size_t)g * seg_bytes;
    keys[(int64_t)t * ld + n_old + i] = reinterpret_cast<const int32_t*>(seg)[(int64_t)t * slice + j];
    pids[(int64_t)t * ld + n_old + i] = seg[(size_t)L * slice * 4 + (size_t)t * slice + j];
}

void hash_dense_sharded(dpf_index* h, const double* Xnew, int64_t n, int64_t n_old) {
    DPF_REQUIRE(h->comm, DPF_ERR_STATE, "dpf_comm_init has not been called");
    const int world = h->cfg.world > 1 ? h->cfg.world : 1, rank = h->cfg.world > 1 ? h->cfg.rank : 0;
    const int L = h->cfg.L, d = h->cfg.d;
    const int64_t slice = (n + world - 1) / world;
    const size_t seg_bytes = ((size_t)L * slice * 5 + 15) / 16 * 16;
    h->comm_stage.reserve(seg_bytes * world);
    unsigned char* seg = reinterpret_cast<unsigned char*>(h->comm_stage.p) + seg_bytes * rank;
    const int64_t r0 = std::min<int64_t>(n, slice * rank), m = std::min<int64_t>(n, r0 + slice) - r0;
    if (m > 0) {
        if (h->dbg[DPF_DBG_HASH_EXACT] == 1 && h->cfg.family_kind == DPF_FAMILY_ANGLE)
            hash_dense_device_exact(h, Xnew + r0 * d, m, reinterpret_cast<int32_t*>(seg), seg + (size_t)L * slice * 4, slice);
        else
            hash_dense_device(h, Xnew + r0 * d, m, reinterpret_cast<int32_t*>(seg), seg + (size_t)L * slice * 4, slice);
    }
    {
        StageTimer tm(h, DPF_T_COMM);
        DPF_NCCL(nccl().AllGather(seg, h->comm_stage.p, seg_bytes, ncclChar, static_cast<ncclComm_t>(h->comm), h->stream));
        const dim3 grid((unsigned)((n + 255) / 256), L);
        k_unpack_gathered_keys<<<grid, 256, 0, h->stream>>>(reinterpret_cast<unsigned char*>(h->comm_stage.p), seg_bytes, slice, n, L,
                                                            n_old, h->key_ld, h->keys.p, h->pids.p); DPF_LAUNCHED();
        DPF_CUDA(cudaGetLastError());
    }
}

// ---- query: per-rank top k -> in-place all-gather -> merge -------------------------------------------------------------
// segment of rank g: nq x topk int32 ids (padded to 8 bytes), then nq x topk FP64 scores
size_t comm_topk_segment_bytes(int64_t nq, int topk) { return ((size_t)nq * topk * 4 + 7) / 8 * 8 + (size_t)nq * topk * 8; }

void comm_topk_buffers(dpf_index* h, int64_t nq, int topk, int32_t** ids_seg, double** sc_seg) {
    DPF_REQUIRE(h->comm, DPF_ERR_STATE, "dpf_comm_init has not been called");
    const int world = h->cfg.world > 1 ? h->cfg.world : 1, rank = h->cfg.world > 1 ? h->cfg.rank : 0;
    const size_t seg = comm_topk_segment_bytes(nq, topk);
    h->comm_topk.reserve(seg * world);
    char* mine = h->comm_topk.p + seg * rank;
    *ids_seg = reinterpret_cast<int32_t*>(mine);
    *sc_seg = reinterpret_cast<double*>(mine + ((size_t)nq * topk * 4 + 7) / 8 * 8);
}

void comm_gather_merge_topk(dpf_index* h, int64_t nq, int topk, int metric, int32_t* ids_out, double* score_out) {
    const int world = h->cfg.world > 1 ? h->cfg.world : 1, rank = h->cfg.world > 1 ? h->cfg.rank : 0;
    const size_t seg = comm_topk_segment_bytes(nq, topk);
    StageTimer tm(h, DPF_T_COMM);
    DPF_NCCL(nccl().AllGather(h->comm_topk.p + seg * rank, h->comm_topk.p, seg, ncclChar, static_cast<ncclComm_t>(h->comm), h->stream));
    const int32_t* gids = reinterpret_cast<const int32_t*>(h->comm_topk.p);
    const double* gsc = reinterpret_cast<const double*>(h->comm_topk.p + ((size_t)nq * topk * 4 + 7) / 8 * 8);
    merge_topk_strided(h, gids, gsc, (int64_t)(seg / 4), (int64_t)(seg / 8), world, nq, topk, metric, ids_out, score_out);
}

}  // namespace dpf
