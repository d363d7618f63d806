// capi.cu — the extern "C" boundary declared in include/dpf.h.  Every entry point takes the handle's lock,
// selects the handle's device, runs the device pipeline on the handle's stream and converts failures into
// DPF_ERR_* codes (no exception leaves the library, nothing aborts the process).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

using namespace dpf;

namespace dpf {
unsigned long long g_launches = 0;
thread_local cudaStream_t tl_stream = nullptr;
}

namespace {

int ilog2_exact(int v) {
    int r = 0;
    while ((1 << r) < v) r++;
    return ((1 << r) == v) ? r : -1;
}

template <class F>
int guarded(dpf_handle h, F&& f) {
    if (!h) return DPF_ERR_INVALID;
    std::lock_guard<std::mutex> lk(h->mu);
    try {
        DPF_CUDA(cudaSetDevice(h->cfg.device));
        tl_stream = h->stream;
        f();
        return DPF_OK;
    } catch (const Error& e) {
        h->last_error = e.msg;
        cudaGetLastError();   // clear a sticky-less error state
        return e.code;
    } catch (const std::bad_alloc&) {
        h->last_error = "host allocation failed";
        return DPF_ERR_NOMEM;
    } catch (const std::exception& e) {
        h->last_error = e.what();
        return DPF_ERR_INVALID;
    }
}

void begin_profile(dpf_index* h) {
    if (!h->profiling) return;
    h->ev_used = 0;
    for (int i = 0; i < DPF_T_COUNT; ++i) h->stage_ms[i] = 0.f;
}
void end_profile(dpf_index* h) {
    if (!h->profiling || h->ev_used == 0) return;
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < DPF_T_COUNT; ++i) h->stage_ms[i] = 0.f;
    for (size_t s = 0; s < h->ev_used; ++s) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_pool[2 * s], h->ev_pool[2 * s + 1]) == cudaSuccess)
            h->stage_ms[h->ev_stage[s]] += ms;
    }
    h->ev_used = 0;
}

template <class T>
void h2d(dpf_index* h, T* dst, const T* src, size_t n) {
    if (n) DPF_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
}
template <class T>
void d2h(dpf_index* h, T* dst, const T* src, size_t n) {
    if (n) DPF_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
}

// the ownership table on the device follows the host copy (pageable source: the copy is staged before the call returns)
void upload_ownership(dpf_index* h) {
    h->own_dev.reserve(h->own_cells.size());
    h2d(h, h->own_dev.p, h->own_cells.data(), h->own_cells.size());
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    h->own.cells = h->own_dev.p;
}

// clears the per-batch device counters (one memset); every query entry point starts with it
void begin_batch(dpf_index* h) {
    DPF_CUDA(cudaMemsetAsync(h->counters.p + CTR_BATCH_FIRST, 0, (CTR_BATCH_END - CTR_BATCH_FIRST) * sizeof(int32_t), h->stream));
}

void require_ready(dpf_index* h, bool need_fit) {
    DPF_REQUIRE(h->family_set, DPF_ERR_STATE, "dpf_set_family has not been called");
    DPF_REQUIRE(h->part_set || h->cfg.pb == 0, DPF_ERR_STATE, "dpf_set_partitioners has not been called");
    if (need_fit) DPF_REQUIRE(h->fitted && h->n > 0, DPF_ERR_STATE, "need to fit the data first");
}

// make room for n more keys (table-major, leading dimension key_ld), preserving what is there
void grow_keys(dpf_index* h, int64_t add) {
    const int L = h->cfg.L;
    const int64_t need = h->n + add;
    if (need <= h->key_ld) return;
    const int64_t nld = std::max<int64_t>(need, h->key_ld + h->key_ld / 2);
    DevBuf<int32_t> nk;
    DevBuf<uint8_t> np;
    nk.reserve((size_t)L * nld);
    np.reserve((size_t)L * nld);
    if (h->n > 0) {
        DPF_CUDA(cudaMemcpy2DAsync(nk.p, nld * sizeof(int32_t), h->keys.p, h->key_ld * sizeof(int32_t),
                                   h->n * sizeof(int32_t), L, cudaMemcpyDeviceToDevice, h->stream));
        DPF_CUDA(cudaMemcpy2DAsync(np.p, nld, h->pids.p, h->key_ld, h->n, L, cudaMemcpyDeviceToDevice, h->stream));
        DPF_CUDA(cudaStreamSynchronize(h->stream));
    }
    std::swap(h->keys.p, nk.p); std::swap(h->keys.cap, nk.cap);
    std::swap(h->pids.p, np.p); std::swap(h->pids.cap, np.cap);
    h->key_ld = nld;
}


void hash_dense_any(dpf_index* h, const double* Xd, int64_t n, int32_t* keys, uint8_t* pids, int64_t ld) {
    if (h->dbg[DPF_DBG_HASH_EXACT] == 1 && h->cfg.family_kind == DPF_FAMILY_ANGLE) hash_dense_device_exact(h, Xd, n, keys, pids, ld);
    else hash_dense_device(h, Xd, n, keys, pids, ld);
}

// widen uint8 pids to the ABI's int32 on the host
void pids_to_host(dpf_index* h, const uint8_t* dev, int64_t count, int32_t* out) {
    std::vector<uint8_t> tmp((size_t)count);
    d2h(h, tmp.data(), dev, (size_t)count);
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t i = 0; i < count; ++i) out[i] = tmp[(size_t)i];
}

// ids of candidate scratch per chunk of queries (2 GiB); DPF_CAND_BUDGET overrides it (tests force many small chunks)
static int64_t cand_budget(const dpf_index* h) {
    const int64_t v = h->dbg[DPF_DBG_CAND_BUDGET];
    return v > 0 ? v : (1LL << 29);
}
#define kCandBudget cand_budget(h)

// allocate the candidate scratch once for the largest chunk (a reallocation between chunks would serialise the
// stream on cudaFree)
void reserve_candidate_scratch(dpf_index* h, const std::vector<int64_t>& ub) {
    const int64_t nq = (int64_t)ub.size() - 1;
    int64_t mx = 1;
    for (int64_t q0 = 0; q0 < nq;) {
        const int64_t q1 = next_chunk_end(ub, q0, kCandBudget);
        mx = std::max(mx, ub[(size_t)q1] - ub[(size_t)q0]);
        q0 = q1;
    }
    h->cand.reserve((size_t)mx);
}

// shared body of the candidate queries: probe, then per memory-bounded chunk expand -> sort -> copy out (CSR)
void run_candidates(dpf_index* h, const QueryKeys& qk, int steps, int probe_mode, int64_t* offsets_out, int32_t* ids_out,
                    int64_t cap, int64_t* total_out) {
    const int64_t nq = qk.nq;
    std::vector<int64_t> ub;
    begin_batch(h);
    probe_count_all(h, qk, steps, probe_mode, ub);
    reserve_candidate_scratch(h, ub);
    DevBuf<int64_t> off;
    off.reserve(nq + 1);
    int64_t total = 0;
    for (int64_t q0 = 0; q0 < nq;) {
        const int64_t q1 = next_chunk_end(ub, q0, kCandBudget);
        const int64_t base = ub[(size_t)q0];
        expand_range(h, qk, steps, probe_mode, q0, q1, base, ub[(size_t)q1] - base);
        const int64_t tc = finalize_candidates_sorted(h, q0, q1, base, off.p);
        if (ids_out && total + tc <= cap) d2h(h, ids_out + total, h->cand.p, (size_t)tc);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        total += tc;
        q0 = q1;
    }
    std::vector<int32_t> cnt((size_t)nq);
    d2h(h, cnt.data(), h->q_cnt.p, (size_t)nq);
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    offsets_out[0] = 0;
    for (int64_t i = 0; i < nq; ++i) offsets_out[i + 1] = offsets_out[i] + cnt[(size_t)i];
    if (total_out) *total_out = total;
    DPF_REQUIRE(total <= cap, DPF_ERR_CAPACITY, "candidate buffer too small");
}

void upload_queries_dense(dpf_index* h, const double* Q, int64_t nq, const int32_t* qids) {
    const int d = h->cfg.d, L = h->cfg.L;
    h->qbuf.reserve((size_t)nq * d);
    h2d(h, h->qbuf.p, Q, (size_t)nq * d);
    h->qidbuf.reserve(nq + 1);
    if (qids) h2d(h, h->qidbuf.p, qids, (size_t)nq);
    h->qkeys.reserve((size_t)L * nq);
    h->qpids.reserve((size_t)L * nq);
}

}  // namespace

extern "C" {

const char* dpf_strerror(int code) {
    switch (code) {
        case DPF_OK: return "ok";
        case DPF_ERR_INVALID: return "invalid argument";
        case DPF_ERR_STATE: return "invalid state / call order";
        case DPF_ERR_CUDA: return "CUDA failure";
        case DPF_ERR_NOMEM: return "out of memory";
        case DPF_ERR_CAPACITY: return "output buffer too small";
        default: return "unknown";
    }
}

int dpf_destroy(dpf_handle h);

int dpf_create(const dpf_config* cfg, dpf_handle* out) {
    if (!cfg || !out) return DPF_ERR_INVALID;
    *out = nullptr;
    if (cfg->abi_version != DPF_ABI_VERSION) return DPF_ERR_INVALID;
    if (cfg->d <= 0 || cfg->L <= 0 || cfg->L > kMaxTables || cfg->k <= 0 || cfg->k > kMaxChain || cfg->pb < 0 ||
        cfg->pb > kMaxPb || cfg->bucket_bits < 1 || cfg->bucket_bits > 32 || cfg->bucket_overflow < 1)
        return DPF_ERR_INVALID;
    const int nb = ilog2_exact(cfg->dir_node_size);
    if (nb < 1 || nb > 8) return DPF_ERR_INVALID;
    const int seg_bits = 32 - cfg->bucket_bits;
    const int maxl = (cfg->k - seg_bits) / nb - 1;            // RandomDrawTreeMap.java:456
    if (maxl < 0 || seg_bits > 12) return DPF_ERR_INVALID;
    if (cfg->world > 1 && (cfg->rank < 0 || cfg->rank >= cfg->world)) return DPF_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return DPF_ERR_CUDA;   // no CPU fallback
    }
    dpf_index* h = nullptr;
    try {
        h = new dpf_index();
        h->cfg = *cfg;
        h->tp = TreeParams{1 << seg_bits, seg_bits, nb, 1 << nb, maxl, cfg->bucket_bits, cfg->pb,
                           (1 << cfg->pb) * (1 << seg_bits), cfg->bucket_overflow};
        DPF_CUDA(cudaSetDevice(cfg->device));
        DPF_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        {   // lowest priority: its (large) kernels fill the gaps of the main stream's chain of small ones, not the reverse
            int lo = 0, hi = 0;
            DPF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            DPF_CUDA(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, lo));
        }
        DPF_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        DPF_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        tl_stream = h->stream;
        {   // keep freed blocks in the pool instead of returning them to the driver at every synchronisation
            cudaMemPool_t pool = nullptr;
            DPF_CUDA(cudaDeviceGetDefaultMemPool(&pool, cfg->device));
            unsigned long long keep = ~0ULL;
            DPF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        DPF_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device));
        h->counters.reserve(CTR_COUNT);
        DPF_CUDA(cudaMemsetAsync(h->counters.p, 0, CTR_COUNT * sizeof(int32_t), h->stream));
        h->occupancy.assign(1 << cfg->pb, 0.0);
        h->own_cells.assign((size_t)cfg->L * kOwnWords, 0u);
        for (int t = 0; t < cfg->L; ++t)
            for (int p = 0; p < (1 << cfg->pb); ++p)  // Partitioner scheme on G GPUs: sub-index p lives on GPU p mod G
                if (cfg->world <= 1 || p % cfg->world == cfg->rank) h->own_cells[(size_t)t * kOwnWords + (p >> 5)] |= 1u << (p & 31);
        upload_ownership(h);
    } catch (const Error& e) {
        dpf_destroy(h);
        return e.code;
    } catch (...) {
        dpf_destroy(h);
        return DPF_ERR_NOMEM;
    }
    *out = h;
    return DPF_OK;
}

int dpf_destroy(dpf_handle h) {
    if (!h) return DPF_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);   // a caller stream that is already gone only returns an error
    cudaGetLastError();
    cudaStream_t own = h->own_stream ? h->own_stream : h->stream;
    try { comm_destroy(h); } catch (...) {}
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    tl_stream = own;        // the buffers go back to the pool in the order of the handle's own stream
    delete h;
    if (own) {
        cudaStreamSynchronize(own);
        cudaStreamDestroy(own);
    }
    tl_stream = nullptr;
    return DPF_OK;
}

const char* dpf_last_error(dpf_handle h) { return h ? h->last_error.c_str() : "null handle"; }

int dpf_sync(dpf_handle h) {
    return guarded(h, [&] { DPF_CUDA(cudaStreamSynchronize(h->stream)); });
}

int dpf_set_stream(dpf_handle h, void* cuda_stream) {
    return guarded(h, [&] {
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        if (cuda_stream) {
            if (!h->own_stream) h->own_stream = h->stream;
            h->stream = (cudaStream_t)cuda_stream;
        } else if (h->own_stream) {
            h->stream = h->own_stream;
            h->own_stream = nullptr;
        }
    });
}

int dpf_set_family(dpf_handle h, const double* A, int32_t P, const int32_t* chain_idx, const double* b, const int32_t* w) {
    return guarded(h, [&] {
        DPF_REQUIRE(A && chain_idx && P > 0, DPF_ERR_INVALID, "null family");
        DPF_REQUIRE(h->n == 0, DPF_ERR_STATE, "hash functions cannot change once vectors are indexed");
        const int d = h->cfg.d, L = h->cfg.L, k = h->cfg.k;
        for (int64_t i = 0; i < (int64_t)L * k; ++i)
            DPF_REQUIRE(chain_idx[i] >= 0 && chain_idx[i] < P, DPF_ERR_INVALID, "chain_idx out of range");
        if (h->cfg.family_kind == DPF_FAMILY_PSTABLE) {
            DPF_REQUIRE(b && w, DPF_ERR_INVALID, "pStable family needs b and w");
            for (int p = 0; p < P; ++p) DPF_REQUIRE(w[p] != 0, DPF_ERR_INVALID, "pStable w must be non-zero");
        }
        h->P = P;
        h->PW = (P + 31) / 32;
        h->hA.assign(A, A + (size_t)P * d);
        h->A.reserve((size_t)P * d);
        h2d(h, h->A.p, A, (size_t)P * d);
        h->chain.reserve((size_t)L * k);
        h2d(h, h->chain.p, chain_idx, (size_t)L * k);
        if (h->cfg.family_kind == DPF_FAMILY_PSTABLE) {
            h->fb.reserve(P);
            h->fw.reserve(P);
            h2d(h, h->fb.p, b, (size_t)P);
            h2d(h, h->fw.p, w, (size_t)P);
        }
        h->At.release();
        if (h->cfg.pb == 0) {
            h->Ap.reserve(1);
            h->part_set = true;
        }
        prepare_family(h);
        h->family_set = true;
    });
}

int dpf_set_partitioners(dpf_handle h, const double* Ap) {
    return guarded(h, [&] {
        DPF_REQUIRE(h->n == 0, DPF_ERR_STATE, "partitioners cannot change once vectors are indexed");
        // the reference builds the partitioner's LSH from the main family's configuration (DensevectorRDFInit.scala:63-70):
        // a pStable index has pStable partitioner chains, which need their b and w
        DPF_REQUIRE(h->cfg.family_kind != DPF_FAMILY_PSTABLE || h->cfg.pb == 0, DPF_ERR_INVALID,
                    "a pStable index takes its partitioners through dpf_set_partitioners_pstable");
        const size_t cnt = (size_t)h->cfg.L * h->cfg.pb * 32;
        DPF_REQUIRE(Ap || cnt == 0, DPF_ERR_INVALID, "null partitioner functions");
        h->Ap.reserve(std::max<size_t>(cnt, 1));
        h2d(h, h->Ap.p, Ap, cnt);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        h->part_pstable = false;
        h->part_set = true;
    });
}

int dpf_set_partitioners_pstable(dpf_handle h, const double* Ap, const double* b, const int32_t* w) {
    return guarded(h, [&] {
        DPF_REQUIRE(h->n == 0, DPF_ERR_STATE, "partitioners cannot change once vectors are indexed");
        DPF_REQUIRE(h->cfg.family_kind == DPF_FAMILY_PSTABLE, DPF_ERR_INVALID,
                    "pStable partitioner chains belong to a pStable index (the angle family uses dpf_set_partitioners)");
        const size_t cnt = (size_t)h->cfg.L * h->cfg.pb * 32, nf = (size_t)h->cfg.L * h->cfg.pb;
        DPF_REQUIRE((Ap && b && w) || cnt == 0, DPF_ERR_INVALID, "null partitioner functions");
        for (size_t i = 0; i < nf; ++i) DPF_REQUIRE(w[i] != 0, DPF_ERR_INVALID, "partitioner w must be non-zero");
        h->Ap.reserve(std::max<size_t>(cnt, 1));
        h->Apb.reserve(std::max<size_t>(nf, 1));
        h->Apw.reserve(std::max<size_t>(nf, 1));
        h2d(h, h->Ap.p, Ap, cnt);
        h2d(h, h->Apb.p, b, nf);
        h2d(h, h->Apw.p, w, nf);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        h->part_pstable = nf > 0;
        h->part_set = true;
    });
}

int dpf_hash_dense(dpf_handle h, const double* X, int64_t n, int32_t* keys_out, int32_t* pids_out) {
    return guarded(h, [&] {
        require_ready(h, false);
        DPF_REQUIRE(n >= 0 && (n == 0 || (X && keys_out)), DPF_ERR_INVALID, "null buffer");
        if (n == 0) return;
        begin_profile(h);
        const int d = h->cfg.d, L = h->cfg.L;
        DevBuf<double> Xd;
        DevBuf<int32_t> kd;
        DevBuf<uint8_t> pd;
        Xd.reserve((size_t)n * d);
        kd.reserve((size_t)L * n);
        pd.reserve((size_t)L * n);
        h2d(h, Xd.p, X, (size_t)n * d);
        hash_dense_any(h, Xd.p, n, kd.p, pd.p, n);
        d2h(h, keys_out, kd.p, (size_t)L * n);
        if (pids_out) pids_to_host(h, pd.p, (int64_t)L * n, pids_out);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        end_profile(h);
    });
}

int dpf_hash_csr(dpf_handle h, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n,
                 int32_t* keys_out, int32_t* pids_out) {
    return guarded(h, [&] {
        require_ready(h, false);
        DPF_REQUIRE(n >= 0 && (n == 0 || (indptr && keys_out)), DPF_ERR_INVALID, "null buffer");
        if (n == 0) return;
        begin_profile(h);
        const int L = h->cfg.L;
        const int64_t nnz = indptr[n] - indptr[0];
        DPF_REQUIRE(indptr[0] == 0, DPF_ERR_INVALID, "indptr must start at 0");
        for (int64_t i = 0; i < nnz; ++i)
            DPF_REQUIRE(indices[i] >= 0 && indices[i] < h->cfg.d, DPF_ERR_INVALID, "CSR index out of range");
        DevBuf<int64_t> pd_;
        DevBuf<int32_t> id_, kd;
        DevBuf<double> vd;
        DevBuf<uint8_t> pd;
        pd_.reserve(n + 1); id_.reserve(std::max<int64_t>(nnz, 1)); vd.reserve(std::max<int64_t>(nnz, 1));
        kd.reserve((size_t)L * n); pd.reserve((size_t)L * n);
        h2d(h, pd_.p, indptr, (size_t)n + 1);
        h2d(h, id_.p, indices, (size_t)nnz);
        h2d(h, vd.p, values, (size_t)nnz);
        hash_csr_device(h, pd_.p, id_.p, vd.p, n, kd.p, pd.p, n);
        d2h(h, keys_out, kd.p, (size_t)L * n);
        if (pids_out) pids_to_host(h, pd.p, (int64_t)L * n, pids_out);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        end_profile(h);
    });
}

// DPF_TRACE=1: host wall-clock per phase of a fit (each phase ends with a stream synchronisation), to stderr
struct PhaseTrace {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t;
    PhaseTrace(const dpf_index* h, cudaStream_t s) : on(h->dbg[DPF_DBG_TRACE] == 1), st(s), t(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[dpf] %-24s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// dpf_set_balanced_partition: at the first fit, the (table, sub-index) cells are dealt to the GPUs largest first, each to
// the GPU with the fewest ids so far.  A cell is one of the reference's per-partition stores
// (RandomDrawTreeMap.java:1430-1459), so this only chooses where each store lives: with 8 sub-indexes whose sizes differ
// by 3.5x (configs[1]) whole sub-indexes cannot be spread evenly over 8 GPUs, their 240 cells can.  Vectors and hash
// functions are replicated, so every rank computes the same assignment from the same pids; it is then kept for the life
// of the handle.
__global__ void __launch_bounds__(256) k_pid_histogram(const uint8_t* __restrict__ pids, int64_t n, int64_t ld, int L,
                                                       unsigned long long* __restrict__ hist /* L x 256 */) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int t = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[pids[(int64_t)t * ld + i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[(size_t)t * 256 + threadIdx.x], (unsigned long long)sh[threadIdx.x]);
    (void)L;
}

static void assign_balanced_partition(dpf_index* h, int64_t n_new) {
    if (!h->balance_partition || h->own_fixed || h->cfg.world <= 1) return;
    const int np = 1 << h->cfg.pb, G = h->cfg.world;
    const int L = h->cfg.L;
    DevBuf<unsigned long long> hist;
    hist.reserve((size_t)L * 256);
    DPF_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)L * 256 * sizeof(unsigned long long), h->stream));
    const dim3 grid((unsigned)std::min<int64_t>((n_new + 255) / 256, 1024), L);
    k_pid_histogram<<<grid, 256, 0, h->stream>>>(h->pids.p, n_new, h->key_ld, L, hist.p); DPF_LAUNCHED();
    std::vector<unsigned long long> occ((size_t)L * 256);
    DPF_CUDA(cudaMemcpyAsync(occ.data(), hist.p, occ.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<int> order;                      // cell = t * 256 + p
    for (int t = 0; t < L; ++t)
        for (int p = 0; p < np; ++p) order.push_back(t * 256 + p);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return occ[a] > occ[b]; });
    std::vector<unsigned long long> load(G, 0);
    std::fill(h->own_cells.begin(), h->own_cells.end(), 0u);
    for (int cell : order) {
        int best = 0;
        for (int g = 1; g < G; ++g)
            if (load[g] < load[best]) best = g;
        load[best] += occ[cell];
        const int t = cell >> 8, p = cell & 255;
        if (best == h->cfg.rank) h->own_cells[(size_t)t * kOwnWords + (p >> 5)] |= 1u << (p & 31);
    }
    upload_ownership(h);
    h->own_fixed = true;
}

constexpr int kFitCopyChunks = 8;

static void fit_dense_common(dpf_index* h, const double* X, int64_t n, bool on_device, bool sharded_hash = false) {
    require_ready(h, false);
    DPF_REQUIRE(n > 0 && X, DPF_ERR_INVALID, "empty fit");
    DPF_REQUIRE(h->n == 0 || h->dense, DPF_ERR_STATE, "index holds sparse vectors");
    DPF_REQUIRE(h->n + n < (1LL << 31), DPF_ERR_INVALID, "ids are int32");
    begin_profile(h);
    PhaseTrace tr(h, h->stream);
    const int d = h->cfg.d;
    h->dense = true;
    const double* Xnew;
    int copy_chunks = 1;
    if (on_device) {
        DPF_REQUIRE(h->n == 0, DPF_ERR_STATE, "dpf_fit_dense_dev borrows the buffer and cannot append");
        h->Xdev = X;
        h->X_borrowed = true;
        Xnew = X;
    } else {
        DPF_REQUIRE(!h->X_borrowed, DPF_ERR_STATE, "cannot append to a borrowed device buffer");
        h->X.grow_keep((size_t)(h->n + n) * d, (size_t)h->n * d, h->stream);
        tr.mark("store: allocate");
        h->Xdev = h->X.p;
        Xnew = h->X.p + (size_t)h->n * d;
        // A large fit is copied in pieces on the second stream while the main stream hashes the pieces that have landed:
        // the copy (1 GB at configs[1]: 18.5 of the 24 ms of a fit through the host API) hides the hashing
        copy_chunks = (!sharded_hash && (size_t)n * d * sizeof(double) >= (64u << 20) && h->dbg[DPF_DBG_TRACE] == 0) ? kFitCopyChunks : 1;
        if (copy_chunks == 1) {
            h2d(h, h->X.p + (size_t)h->n * d, X, (size_t)n * d);
            tr.mark("store: host -> device");
        }
    }
    if (h->store_mode != DPF_STORE_F64_ONLY && h->Xc_kind != DPF_STORE_KIND_F32) {
        // room for a byte copy of the rows, taken before the build's scratch buffers carve up the pool's free blocks
        // (allocated last it regularly cost a pool growth of several ms); released again if the data are not bytes
        try {
            h->Xc.grow_keep((size_t)(h->n + n) * (size_t)((h->cfg.d + 15) / 16 * 16), 0, h->stream);
        } catch (const Error& err) {
            if (err.code != DPF_ERR_NOMEM) throw;
            cudaGetLastError();
        }
    }
    grow_keys(h, n);
    tr.mark("keys: allocate");
    if (sharded_hash) hash_dense_sharded(h, Xnew, n, h->n);      // this rank's slice + one all-gather of the keys
    else if (copy_chunks > 1) {
        cudaEvent_t ev[kFitCopyChunks] = {};
        try {
            DPF_CUDA(cudaEventRecord(h->ev_fork, h->stream));                 // the (re)allocated store is ready
            DPF_CUDA(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
            const int64_t piece = ((n + copy_chunks - 1) / copy_chunks + 63) / 64 * 64;
            // copies run one piece ahead of the hashing (whose fix-up step waits for the device once per piece)
            auto copy_piece = [&](int c) {
                const int64_t r0 = (int64_t)c * piece;
                if (r0 >= n) return;
                const int64_t m = std::min(piece, n - r0);
                DPF_CUDA(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
                DPF_CUDA(cudaMemcpyAsync(const_cast<double*>(Xnew) + r0 * d, X + r0 * d, (size_t)m * d * sizeof(double), cudaMemcpyHostToDevice,
                                         h->aux_stream));
                DPF_CUDA(cudaEventRecord(ev[c], h->aux_stream));
            };
            copy_piece(0);
            copy_piece(1);
            for (int c = 0; (int64_t)c * piece < n; ++c) {
                const int64_t r0 = (int64_t)c * piece, m = std::min(piece, n - r0);
                DPF_CUDA(cudaStreamWaitEvent(h->stream, ev[c], 0));
                hash_dense_any(h, Xnew + r0 * d, m, h->keys.p + h->n + r0, h->pids.p + h->n + r0, h->key_ld);
                copy_piece(c + 2);
            }
        } catch (...) {
            cudaStreamSynchronize(h->aux_stream);
            for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
            throw;
        }
        for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
    } else hash_dense_any(h, Xnew, n, h->keys.p + h->n, h->pids.p + h->n, h->key_ld);
    tr.mark("hash");
    if (h->n == 0) assign_balanced_partition(h, n);
    // the new vectors count only once the forest and the store have been rebuilt: a failure in between (out of memory,
    // node capacity) leaves the handle un-fitted with its old size instead of fitted with stale arrays
    // A small batch on top of a fitted index is PUT into the existing forest (O(batch): RandomDrawTreeMap.put one id
    // after the other, incremental.cu); anything else — or a put that runs out of head-room — rebuilds from the keys.
    const int64_t n_old = h->n;
    const int64_t mode = h->dbg[DPF_DBG_APPEND];
    bool incremental = h->fitted && n_old > 0 && mode != 1 && (mode == 2 || n <= std::max<int64_t>(n_old / 256, 64));
    h->fitted = false;
    h->n += n;
    try {
        if (incremental) {
            incremental = forest_insert_incremental(h, n_old, n);
            if (incremental) { rebuild_leaf_table(h); update_occupancy(h); }
            tr.mark("forest (incremental put)");
        }
        if (!incremental) {
            DPF_REQUIRE(h->n_removed == 0 || mode != 3, DPF_ERR_STATE, "rebuild after remove refused (debug option)");
            build_forest(h);
            tr.mark("forest");
        }
        h->fitted = true;
        if (!(incremental && append_compact_store(h, n_old, n))) build_compact_store(h);
        tr.mark("compact store");
    } catch (...) {
        h->n -= n;
        h->fitted = false;
        throw;
    }
    h->stats[DPF_STAT_SIZE] = h->n;
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    end_profile(h);
}

int dpf_set_balanced_partition(dpf_handle h, int32_t enable) {
    return guarded(h, [&] {
        DPF_REQUIRE(h->n == 0, DPF_ERR_STATE, "dpf_set_balanced_partition must be called before fit");
        h->balance_partition = enable != 0;
    });
}

int dpf_owned_subindexes(dpf_handle h, uint8_t* owned_out) {
    return guarded(h, [&] {
        DPF_REQUIRE(owned_out, DPF_ERR_INVALID, "null buffer");
        const int np = 1 << h->cfg.pb;
        for (int t = 0; t < h->cfg.L; ++t)
            for (int p = 0; p < np; ++p) owned_out[(size_t)t * np + p] = h->owns(t, p) ? 1 : 0;
    });
}

int dpf_set_store_mode(dpf_handle h, int32_t mode) {
    return guarded(h, [&] {
        DPF_REQUIRE(mode == DPF_STORE_AUTO || mode == DPF_STORE_F64_ONLY || mode == DPF_STORE_NARROWEST, DPF_ERR_INVALID, "bad store mode");
        DPF_REQUIRE(h->n == 0, DPF_ERR_STATE, "dpf_set_store_mode must be called before fit");
        h->store_mode = mode;
    });
}

int dpf_fit_dense(dpf_handle h, const double* X, int64_t n) {
    return guarded(h, [&] { fit_dense_common(h, X, n, false); });
}
int dpf_fit_dense_dev(dpf_handle h, const double* X_dev, int64_t n) {
    return guarded(h, [&] { fit_dense_common(h, X_dev, n, true); });
}

int dpf_fit_dense_sharded(dpf_handle h, const double* X, int64_t n) {
    return guarded(h, [&] { fit_dense_common(h, X, n, false, true); });
}
int dpf_fit_dense_sharded_dev(dpf_handle h, const double* X_dev, int64_t n) {
    return guarded(h, [&] { fit_dense_common(h, X_dev, n, true, true); });
}

int dpf_comm_unique_id(uint8_t* id_out) {
    if (!id_out) return DPF_ERR_INVALID;
    try {
        comm_unique_id(id_out);
        return DPF_OK;
    } catch (const Error& e) {
        return e.code;
    }
}
int dpf_comm_init(dpf_handle h, const uint8_t* id) {
    return guarded(h, [&] {
        DPF_REQUIRE(id, DPF_ERR_INVALID, "null id");
        comm_init(h, id);
    });
}
int dpf_comm_destroy(dpf_handle h) {
    return guarded(h, [&] {
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        comm_destroy(h);
    });
}

// RandomDrawTreeMap.remove (RandomDrawTreeMap.java:1817-1932) for a batch of ids, in every table
int dpf_remove(dpf_handle h, const int32_t* ids, int64_t m, int64_t* removed_entries_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(m >= 0 && (m == 0 || ids), DPF_ERR_INVALID, "null buffer");
        if (removed_entries_out) *removed_entries_out = 0;
        if (m == 0) return;
        const int64_t gone = forest_remove(h, ids, m);
        rebuild_leaf_table(h);
        // a later rebuild (large append) must not bring the ids back
        if (h->removed.cap < (size_t)h->key_ld) {
            DevBuf<uint8_t> nr;
            nr.reserve((size_t)h->key_ld);
            DPF_CUDA(cudaMemsetAsync(nr.p, 0, (size_t)h->key_ld, h->stream));
            if (h->removed.p) DPF_CUDA(cudaMemcpyAsync(nr.p, h->removed.p, h->removed.cap, cudaMemcpyDeviceToDevice, h->stream));
            DPF_CUDA(cudaStreamSynchronize(h->stream));
            std::swap(h->removed.p, nr.p);
            std::swap(h->removed.cap, nr.cap);
        }
        const uint8_t one = 1;
        for (int64_t j = 0; j < m; ++j)
            if (ids[j] >= 0 && ids[j] < h->n)
                DPF_CUDA(cudaMemcpyAsync(h->removed.p + ids[j], &one, 1, cudaMemcpyHostToDevice, h->stream));
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        h->n_removed += m;
        update_occupancy(h);
        if (removed_entries_out) *removed_entries_out = gone;
    });
}

int dpf_fit_csr(dpf_handle h, const int64_t* indptr, const int32_t* indices, const double* values, int64_t n) {
    return guarded(h, [&] {
        require_ready(h, false);
        DPF_REQUIRE(n > 0 && indptr && indices && values, DPF_ERR_INVALID, "empty fit");
        DPF_REQUIRE(h->n == 0 || !h->dense, DPF_ERR_STATE, "index holds dense vectors");
        DPF_REQUIRE(indptr[0] == 0, DPF_ERR_INVALID, "indptr must start at 0");
        DPF_REQUIRE(h->n + n < (1LL << 31), DPF_ERR_INVALID, "ids are int32");
        begin_profile(h);
        h->dense = false;
        const int64_t nnz = indptr[n];
        for (int64_t i = 0; i < nnz; ++i)
            DPF_REQUIRE(indices[i] >= 0 && indices[i] < h->cfg.d, DPF_ERR_INVALID, "CSR index out of range");
        // append to the CSR store (row pointers rebased)
        h->sp_ptr.grow_keep((size_t)(h->n + n + 1), (size_t)(h->n ? h->n + 1 : 0), h->stream);
        h->sp_idx.grow_keep((size_t)(h->sp_nnz + nnz) + 1, (size_t)h->sp_nnz, h->stream);
        h->sp_val.grow_keep((size_t)(h->sp_nnz + nnz) + 1, (size_t)h->sp_nnz, h->stream);
        std::vector<int64_t> rebased((size_t)n + 1);
        for (int64_t i = 0; i <= n; ++i) rebased[(size_t)i] = indptr[i] + h->sp_nnz;
        h2d(h, h->sp_ptr.p + h->n, rebased.data(), (size_t)n + 1);
        h2d(h, h->sp_idx.p + h->sp_nnz, indices, (size_t)nnz);
        h2d(h, h->sp_val.p + h->sp_nnz, values, (size_t)nnz);
        grow_keys(h, n);
        hash_csr_device(h, h->sp_ptr.p + h->n, h->sp_idx.p, h->sp_val.p, n, h->keys.p + h->n, h->pids.p + h->n, h->key_ld);
        DPF_CUDA(cudaStreamSynchronize(h->stream));   // `rebased` must outlive the copy
        h->fitted = false;
        h->sp_nnz += nnz;
        h->n += n;
        try {
            build_forest(h);
        } catch (...) {
            h->sp_nnz -= nnz;
            h->n -= n;
            throw;
        }
        h->stats[DPF_STAT_SIZE] = h->n;
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        end_profile(h);
    });
}

int64_t dpf_size(dpf_handle h) { return h ? h->n : -1; }

int dpf_query_candidates_dense(dpf_handle h, const double* Q, int64_t nq, const int32_t* qids, int32_t steps,
                               int32_t probe_mode, int64_t* offsets_out, int32_t* ids_out, int64_t cap,
                               int64_t* total_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && Q && offsets_out, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(steps >= 0 && (probe_mode == DPF_PROBE_NONE || probe_mode == DPF_PROBE_DENSE), DPF_ERR_INVALID, "bad steps/probe_mode");
        begin_profile(h);
        upload_queries_dense(h, Q, nq, qids);
        hash_dense_any(h, h->qbuf.p, nq, h->qkeys.p, h->qpids.p, nq);
        run_candidates(h, QueryKeys{h->qkeys.p, nq, nq, qids ? h->qidbuf.p : nullptr}, steps, probe_mode, offsets_out,
                       ids_out, cap, total_out);
        end_profile(h);
    });
}

int dpf_query_candidates_csr(dpf_handle h, const int64_t* indptr, const int32_t* indices, const double* values,
                             int64_t nq, const int32_t* qids, int32_t steps, int64_t* offsets_out, int32_t* ids_out,
                             int64_t cap, int64_t* total_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && indptr && offsets_out, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(indptr[0] == 0 && steps >= 0, DPF_ERR_INVALID, "indptr must start at 0");
        begin_profile(h);
        const int L = h->cfg.L;
        const int64_t nnz = indptr[nq];
        for (int64_t i = 0; i < nnz; ++i)
            DPF_REQUIRE(indices[i] >= 0 && indices[i] < h->cfg.d, DPF_ERR_INVALID, "CSR index out of range");
        DevBuf<int64_t> pd_;
        DevBuf<int32_t> id_;
        DevBuf<double> vd;
        pd_.reserve(nq + 1); id_.reserve(std::max<int64_t>(nnz, 1)); vd.reserve(std::max<int64_t>(nnz, 1));
        h2d(h, pd_.p, indptr, (size_t)nq + 1);
        h2d(h, id_.p, indices, (size_t)nnz);
        h2d(h, vd.p, values, (size_t)nnz);
        h->qidbuf.reserve(nq + 1);
        if (qids) h2d(h, h->qidbuf.p, qids, (size_t)nq);
        h->qkeys.reserve((size_t)L * nq);
        h->qpids.reserve((size_t)L * nq);
        hash_csr_device(h, pd_.p, id_.p, vd.p, nq, h->qkeys.p, h->qpids.p, nq);
        // the sparse overload has steps but no probes (RandomDrawTreeMap.java:686-732, quirk Q5)
        run_candidates(h, QueryKeys{h->qkeys.p, nq, nq, qids ? h->qidbuf.p : nullptr}, steps, DPF_PROBE_NONE, offsets_out,
                       ids_out, cap, total_out);
        end_profile(h);
    });
}

int dpf_query_candidates_by_id(dpf_handle h, const int32_t* qids, int64_t nq, int32_t steps, int64_t* offsets_out,
                               int32_t* ids_out, int64_t cap, int64_t* total_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && qids && offsets_out && steps >= 0, DPF_ERR_INVALID, "null buffer");
        begin_profile(h);
        const int L = h->cfg.L;
        h->qidbuf.reserve(nq + 1);
        h2d(h, h->qidbuf.p, qids, (size_t)nq);
        h->qkeys.reserve((size_t)L * nq);
        h->qpids.reserve((size_t)L * nq);
        gather_query_keys(h, h->qidbuf.p, nq);
        run_candidates(h, QueryKeys{h->qkeys.p, nq, nq, h->qidbuf.p}, steps, DPF_PROBE_NONE, offsets_out, ids_out, cap,
                       total_out);
        end_profile(h);
    });
}

static void topk_device(dpf_index* h, const double* Qd, int64_t nq, const int32_t* qids_dev, int steps, int probe_mode,
                        int topk, int metric, int32_t* ids_out_dev, double* score_out_dev) {
    const int L = h->cfg.L;
    h->qkeys.reserve((size_t)L * nq);
    h->qpids.reserve((size_t)L * nq);
    begin_batch(h);
    hash_dense_any(h, Qd, nq, h->qkeys.p, h->qpids.p, nq);
    const QueryKeys qk{h->qkeys.p, nq, nq, qids_dev};
    h->Q8_valid = false;
    const bool aligned = (reinterpret_cast<uintptr_t>(Qd) & 15) == 0;   // query rows are moved by 16-byte bulk copies / LDG.128
    if (score_u8_usable(h) && aligned) prepare_queries_u8(h, Qd, nq, metric == DPF_METRIC_L2);
    if (aligned && bucket_major_supported(h, metric, topk)) {
        // bucket-major: chunks of queries whose worst-case pair count fits the scratch; nothing is read back
        int cap = 1;
        const int64_t chunk = std::min(bm_chunk_queries(h, steps, probe_mode, &cap), std::max<int64_t>(1, kMaxPool / bm_pool_per_query(h, topk, steps)));
        for (int64_t q0 = 0; q0 < nq; q0 += chunk)
            topk_bucket_major(h, Qd, qk, steps, probe_mode, q0, std::min(nq, q0 + chunk), cap, topk, metric, ids_out_dev, score_out_dev);
        return;
    }
    // row-major (d > 128, squared L2 on real-valued data, k > 256): candidate lists, memory-bounded chunks of queries
    std::vector<int64_t> ub;
    probe_count_all(h, qk, steps, probe_mode, ub);
    reserve_candidate_scratch(h, ub);
    for (int64_t q0 = 0; q0 < nq;) {
        const int64_t q1 = next_chunk_end(ub, q0, kCandBudget);
        const int64_t base = ub[(size_t)q0];
        expand_range(h, qk, steps, probe_mode, q0, q1, base, ub[(size_t)q1] - base);
        int64_t max_cnt = 1;
        for (int64_t q = q0; q < q1; ++q) max_cnt = std::max(max_cnt, ub[(size_t)q + 1] - ub[(size_t)q]);
        rerank_topk(h, Qd, q0, q1, base, h->q_off.p, h->q_cnt.p, h->cand.p, max_cnt, ub[(size_t)q1] - base, topk, metric,
                    ids_out_dev, score_out_dev);
        q0 = q1;
    }
}

int dpf_query_topk_dense(dpf_handle h, const double* Q, int64_t nq, const int32_t* qids, int32_t steps,
                         int32_t probe_mode, int32_t topk, int32_t metric, int32_t* ids_out, double* score_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && Q && ids_out && score_out, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(steps >= 0 && metric >= 0 && metric <= 2, DPF_ERR_INVALID, "bad steps/metric");
        DPF_REQUIRE(topk >= 1 && topk <= 256, DPF_ERR_INVALID, "topk must be in 1..256");
        begin_profile(h);
        upload_queries_dense(h, Q, nq, qids);
        h->out_ids.reserve((size_t)nq * topk);
        h->out_scores.reserve((size_t)nq * topk);
        topk_device(h, h->qbuf.p, nq, qids ? h->qidbuf.p : nullptr, steps, probe_mode, topk, metric, h->out_ids.p,
                    h->out_scores.p);
        d2h(h, ids_out, h->out_ids.p, (size_t)nq * topk);
        d2h(h, score_out, h->out_scores.p, (size_t)nq * topk);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        end_profile(h);
    });
}

int dpf_query_topk_dense_dev(dpf_handle h, const double* Q_dev, int64_t nq, const int32_t* qids_dev, int32_t steps,
                             int32_t probe_mode, int32_t topk, int32_t metric, int32_t* ids_out_dev,
                             double* score_out_dev) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && Q_dev && ids_out_dev && score_out_dev, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(steps >= 0 && metric >= 0 && metric <= 2, DPF_ERR_INVALID, "bad steps/metric");
        DPF_REQUIRE(topk >= 1 && topk <= 256, DPF_ERR_INVALID, "topk must be in 1..256");
        begin_profile(h);
        topk_device(h, Q_dev, nq, qids_dev, steps, probe_mode, topk, metric, ids_out_dev, score_out_dev);
        end_profile(h);
    });
}

// collective over the ranks of the handle's communicator: every rank passes the same queries and gets the global top k
int dpf_query_topk_dense_all_dev(dpf_handle h, const double* Q_dev, int64_t nq, const int32_t* qids_dev, int32_t steps,
                                 int32_t probe_mode, int32_t topk, int32_t metric, int32_t* ids_out_dev, double* score_out_dev) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && Q_dev && ids_out_dev && score_out_dev, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(steps >= 0 && metric >= 0 && metric <= 2, DPF_ERR_INVALID, "bad steps/metric");
        DPF_REQUIRE(topk >= 1 && topk <= 256, DPF_ERR_INVALID, "topk must be in 1..256");
        begin_profile(h);
        int32_t* ids_seg;
        double* sc_seg;
        comm_topk_buffers(h, nq, topk, &ids_seg, &sc_seg);
        topk_device(h, Q_dev, nq, qids_dev, steps, probe_mode, topk, metric, ids_seg, sc_seg);
        comm_gather_merge_topk(h, nq, topk, metric, ids_out_dev, score_out_dev);
        end_profile(h);
    });
}

int dpf_query_topk_dense_all(dpf_handle h, const double* Q, int64_t nq, const int32_t* qids, int32_t steps, int32_t probe_mode,
                             int32_t topk, int32_t metric, int32_t* ids_out, double* score_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && Q && ids_out && score_out, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(steps >= 0 && metric >= 0 && metric <= 2, DPF_ERR_INVALID, "bad steps/metric");
        DPF_REQUIRE(topk >= 1 && topk <= 256, DPF_ERR_INVALID, "topk must be in 1..256");
        begin_profile(h);
        upload_queries_dense(h, Q, nq, qids);
        h->out_ids.reserve((size_t)nq * topk);
        h->out_scores.reserve((size_t)nq * topk);
        int32_t* ids_seg;
        double* sc_seg;
        comm_topk_buffers(h, nq, topk, &ids_seg, &sc_seg);
        topk_device(h, h->qbuf.p, nq, qids ? h->qidbuf.p : nullptr, steps, probe_mode, topk, metric, ids_seg, sc_seg);
        comm_gather_merge_topk(h, nq, topk, metric, h->out_ids.p, h->out_scores.p);
        d2h(h, ids_out, h->out_ids.p, (size_t)nq * topk);
        d2h(h, score_out, h->out_scores.p, (size_t)nq * topk);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        end_profile(h);
    });
}

int dpf_rerank_dense(dpf_handle h, const double* Q, int64_t nq, const int64_t* offsets, const int32_t* cand,
                     int32_t topk, int32_t metric, int32_t* ids_out, double* score_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nq > 0 && Q && offsets && ids_out && score_out, DPF_ERR_INVALID, "null buffer");
        DPF_REQUIRE(metric >= 0 && metric <= 2, DPF_ERR_INVALID, "bad metric");
        DPF_REQUIRE(topk >= 1 && topk <= 256, DPF_ERR_INVALID, "topk must be in 1..256");
        DPF_REQUIRE(offsets[0] == 0, DPF_ERR_INVALID, "offsets must start at 0");
        for (int64_t q = 0; q < nq; ++q) DPF_REQUIRE(offsets[q + 1] >= offsets[q], DPF_ERR_INVALID, "offsets must be non-decreasing");
        const int64_t total = offsets[nq];
        DPF_REQUIRE(total == 0 || cand, DPF_ERR_INVALID, "null candidate buffer");
        for (int64_t i = 0; i < total; ++i)
            DPF_REQUIRE(cand[i] >= 0 && cand[i] < h->n, DPF_ERR_INVALID, "candidate id out of range");
        begin_profile(h);
        begin_batch(h);
        const int d = h->cfg.d;
        h->qbuf.reserve((size_t)nq * d);
        h2d(h, h->qbuf.p, Q, (size_t)nq * d);
        DevBuf<int64_t> off;
        DevBuf<int32_t> cnt, cd;
        off.reserve(nq + 1); cnt.reserve(nq); cd.reserve(std::max<int64_t>(total, 1));
        h2d(h, off.p, offsets, (size_t)nq + 1);
        h2d(h, cd.p, cand, (size_t)total);
        counts_from_offsets(h, off.p, nq, cnt.p);
        h->out_ids.reserve((size_t)nq * topk);
        h->out_scores.reserve((size_t)nq * topk);
        int64_t max_cnt = 1;
        for (int64_t q = 0; q < nq; ++q) max_cnt = std::max(max_cnt, offsets[q + 1] - offsets[q]);
        rerank_topk(h, h->qbuf.p, 0, nq, 0, off.p, cnt.p, cd.p, max_cnt, total, topk, metric, h->out_ids.p, h->out_scores.p);
        d2h(h, ids_out, h->out_ids.p, (size_t)nq * topk);
        d2h(h, score_out, h->out_scores.p, (size_t)nq * topk);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        end_profile(h);
    });
}

// ---- persist / reload -------------------------------------------------------------------------------------------
// File = "DPFIDX02" | dpf_config | P, dense, has_bw, n, nnz | shard (store mode, balanced flag, ownership mask, removed
// count) | A | chain | [b, w] | Ap | store (X or CSR) | keys | pids | [removed flags].
// A file holds the shard of the rank that saved it: on G GPUs every rank saves and loads its own file, and gets back the
// sub-indexes it owned (also when they were dealt by occupancy) and the store mode it was fitted with.  The forest
// arrays are written as they stand (nodes, bucket ranges, the arena of buckets moved by incremental puts): after puts
// and removes the tree is no longer a function of the keys alone, and a reload must be the same tree.
// The flat forest is not written: it is a pure function of (keys, pids, ids in ascending order) and build_forest
// re-creates it at > 150M vectors/s, bit for bit (the same call an append makes) — what would be expensive to redo,
// hashing the vectors, is what the file keeps.
}  // extern "C"
namespace {
struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
};
template <class T>
void dev_to_file(dpf_index* h, FILE* f, const T* dev, size_t count) {
    const size_t chunk = (64u << 20) / sizeof(T);
    std::vector<T> buf(std::min(chunk, std::max<size_t>(count, 1)));
    for (size_t at = 0; at < count; at += chunk) {
        const size_t m = std::min(chunk, count - at);
        DPF_CUDA(cudaMemcpyAsync(buf.data(), dev + at, m * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        DPF_REQUIRE(fwrite(buf.data(), sizeof(T), m, f) == m, DPF_ERR_INVALID, "dpf_save: short write");
    }
}
template <class T>
void file_to_dev(dpf_index* h, FILE* f, T* dev, size_t count) {
    const size_t chunk = (64u << 20) / sizeof(T);
    std::vector<T> buf(std::min(chunk, std::max<size_t>(count, 1)));
    for (size_t at = 0; at < count; at += chunk) {
        const size_t m = std::min(chunk, count - at);
        DPF_REQUIRE(fread(buf.data(), sizeof(T), m, f) == m, DPF_ERR_INVALID, "dpf_load: truncated file");
        DPF_CUDA(cudaMemcpyAsync(dev + at, buf.data(), m * sizeof(T), cudaMemcpyHostToDevice, h->stream));
        DPF_CUDA(cudaStreamSynchronize(h->stream));
    }
}
template <class T>
void put(FILE* f, const T& v) { DPF_REQUIRE(fwrite(&v, sizeof(T), 1, f) == 1, DPF_ERR_INVALID, "dpf_save: short write"); }
template <class T>
void get(FILE* f, T& v) { DPF_REQUIRE(fread(&v, sizeof(T), 1, f) == 1, DPF_ERR_INVALID, "dpf_load: truncated file"); }
const char kMagic[8] = {'D', 'P', 'F', 'I', 'D', 'X', '0', '3'};
struct ShardHeader {
    int32_t store_mode, balance_partition, own_fixed, reserved;
    int64_t n_removed;
};
}  // namespace
extern "C" {

int dpf_save(dpf_handle h, const char* path) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(path, DPF_ERR_INVALID, "null path");
        File file;
        file.f = fopen(path, "wb");
        DPF_REQUIRE(file.f, DPF_ERR_INVALID, std::string("dpf_save: cannot open ") + path);
        FILE* f = file.f;
        const int d = h->cfg.d, L = h->cfg.L, k = h->cfg.k;
        DPF_REQUIRE(fwrite(kMagic, 1, 8, f) == 8, DPF_ERR_INVALID, "dpf_save: short write");
        put(f, h->cfg);
        const int32_t P = h->P, dense = h->dense ? 1 : 0, has_bw = h->cfg.family_kind == DPF_FAMILY_PSTABLE ? 1 : 0;
        put(f, P); put(f, dense); put(f, has_bw); put(f, h->n); put(f, h->sp_nnz);
        ShardHeader sh{};
        sh.store_mode = h->store_mode;
        sh.balance_partition = h->balance_partition ? 1 : 0;
        sh.own_fixed = h->own_fixed ? 1 : 0;
        sh.n_removed = h->n_removed;
        put(f, sh);
        DPF_REQUIRE(fwrite(h->own_cells.data(), sizeof(uint32_t), h->own_cells.size(), f) == h->own_cells.size(), DPF_ERR_INVALID,
                    "dpf_save: short write");
        dev_to_file(h, f, h->A.p, (size_t)P * d);
        dev_to_file(h, f, h->chain.p, (size_t)L * k);
        if (has_bw) { dev_to_file(h, f, h->fb.p, (size_t)P); dev_to_file(h, f, h->fw.p, (size_t)P); }
        dev_to_file(h, f, h->Ap.p, (size_t)L * h->cfg.pb * 32);
        put(f, (int32_t)(h->part_pstable ? 1 : 0));
        if (h->part_pstable) { dev_to_file(h, f, h->Apb.p, (size_t)L * h->cfg.pb); dev_to_file(h, f, h->Apw.p, (size_t)L * h->cfg.pb); }
        if (dense) {
            dev_to_file(h, f, h->Xdev, (size_t)h->n * d);
        } else {
            dev_to_file(h, f, h->sp_ptr.p, (size_t)h->n + 1);
            dev_to_file(h, f, h->sp_idx.p, (size_t)h->sp_nnz);
            dev_to_file(h, f, h->sp_val.p, (size_t)h->sp_nnz);
        }
        for (int t = 0; t < L; ++t) dev_to_file(h, f, h->keys.p + (size_t)t * h->key_ld, (size_t)h->n);
        for (int t = 0; t < L; ++t) dev_to_file(h, f, h->pids.p + (size_t)t * h->key_ld, (size_t)h->n);
        if (h->n_removed > 0) dev_to_file(h, f, h->removed.p, (size_t)h->n);
        // the forest itself, as it stands (after incremental puts and removes it is not a function of the keys alone)
        const int64_t slots = (int64_t)h->num_nodes * h->tp.W;
        put(f, h->num_nodes); put(f, h->arena_used);
        for (int t = 0; t <= L; ++t) put(f, h->h_table_base[(size_t)t]);
        for (int i : {DPF_STAT_SINGLETON_SPLITS, DPF_STAT_SPLITS}) put(f, h->stats[i]);
        for (double v : h->occupancy) put(f, v);
        dev_to_file(h, f, h->child_ptr.p, (size_t)slots);
        dev_to_file(h, f, h->child_cnt.p, (size_t)slots);
        dev_to_file(h, f, h->node_table.p, (size_t)h->num_nodes);
        dev_to_file(h, f, h->ids_sorted.p, (size_t)h->arena_used);
        DPF_REQUIRE(fflush(f) == 0, DPF_ERR_INVALID, "dpf_save: flush failed");
    });
}

int dpf_load(const char* path, int32_t device, dpf_handle* out) {
    if (!path || !out) return DPF_ERR_INVALID;
    *out = nullptr;
    File file;
    file.f = fopen(path, "rb");
    if (!file.f) return DPF_ERR_INVALID;
    FILE* f = file.f;
    char magic[8];
    dpf_config cfg;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kMagic, 8) != 0 || fread(&cfg, sizeof(cfg), 1, f) != 1) return DPF_ERR_INVALID;
    cfg.device = device;
    dpf_handle h = nullptr;
    int rc = dpf_create(&cfg, &h);
    if (rc != DPF_OK) return rc;
    rc = guarded(h, [&] {
        const int d = cfg.d, L = cfg.L, k = cfg.k;
        int32_t P = 0, dense = 0, has_bw = 0;
        int64_t n = 0, nnz = 0;
        get(f, P); get(f, dense); get(f, has_bw); get(f, n); get(f, nnz);
        DPF_REQUIRE(P > 0 && n > 0 && nnz >= 0 && has_bw == (cfg.family_kind == DPF_FAMILY_PSTABLE ? 1 : 0), DPF_ERR_INVALID,
                    "dpf_load: bad header");
        ShardHeader sh{};
        get(f, sh);
        h->store_mode = sh.store_mode;
        h->balance_partition = sh.balance_partition != 0;
        h->own_fixed = sh.own_fixed != 0;
        DPF_REQUIRE(fread(h->own_cells.data(), sizeof(uint32_t), h->own_cells.size(), f) == h->own_cells.size(), DPF_ERR_INVALID,
                    "dpf_load: truncated file");
        upload_ownership(h);
        h->P = P;
        h->PW = (P + 31) / 32;
        h->hA.resize((size_t)P * d);
        DPF_REQUIRE(fread(h->hA.data(), sizeof(double), h->hA.size(), f) == h->hA.size(), DPF_ERR_INVALID, "dpf_load: truncated file");
        h->A.reserve(h->hA.size());
        h2d(h, h->A.p, h->hA.data(), h->hA.size());
        h->chain.reserve((size_t)L * k);
        file_to_dev(h, f, h->chain.p, (size_t)L * k);
        if (has_bw) {
            h->fb.reserve(P); h->fw.reserve(P);
            file_to_dev(h, f, h->fb.p, (size_t)P);
            file_to_dev(h, f, h->fw.p, (size_t)P);
        }
        h->Ap.reserve(std::max<size_t>((size_t)L * cfg.pb * 32, 1));
        file_to_dev(h, f, h->Ap.p, (size_t)L * cfg.pb * 32);
        int32_t part_pstable = 0;
        get(f, part_pstable);
        h->part_pstable = part_pstable != 0;
        if (h->part_pstable) {
            h->Apb.reserve((size_t)L * cfg.pb); h->Apw.reserve((size_t)L * cfg.pb);
            file_to_dev(h, f, h->Apb.p, (size_t)L * cfg.pb);
            file_to_dev(h, f, h->Apw.p, (size_t)L * cfg.pb);
        }
        prepare_family(h);
        h->family_set = h->part_set = true;
        h->dense = dense != 0;
        if (h->dense) {
            h->X.reserve((size_t)n * d);
            file_to_dev(h, f, h->X.p, (size_t)n * d);
            h->Xdev = h->X.p;
        } else {
            h->sp_ptr.reserve((size_t)n + 1);
            h->sp_idx.reserve((size_t)nnz + 1);
            h->sp_val.reserve((size_t)nnz + 1);
            file_to_dev(h, f, h->sp_ptr.p, (size_t)n + 1);
            file_to_dev(h, f, h->sp_idx.p, (size_t)nnz);
            file_to_dev(h, f, h->sp_val.p, (size_t)nnz);
            h->sp_nnz = nnz;
        }
        grow_keys(h, n);
        for (int t = 0; t < L; ++t) file_to_dev(h, f, h->keys.p + (size_t)t * h->key_ld, (size_t)n);
        for (int t = 0; t < L; ++t) file_to_dev(h, f, h->pids.p + (size_t)t * h->key_ld, (size_t)n);
        if (sh.n_removed > 0) {
            h->removed.reserve((size_t)h->key_ld);
            DPF_CUDA(cudaMemsetAsync(h->removed.p, 0, (size_t)h->key_ld, h->stream));
            file_to_dev(h, f, h->removed.p, (size_t)n);
            h->n_removed = sh.n_removed;
        }
        h->n = n;
        {   // the forest as it was saved, with the head-room a build leaves for incremental puts
            int32_t num_nodes = 0;
            int64_t arena_used = 0;
            get(f, num_nodes); get(f, arena_used);
            DPF_REQUIRE(num_nodes > 0 && arena_used >= 0, DPF_ERR_INVALID, "dpf_load: bad forest header");
            h->h_table_base.assign((size_t)L + 1, 0);
            for (int t = 0; t <= L; ++t) get(f, h->h_table_base[(size_t)t]);
            for (int i : {DPF_STAT_SINGLETON_SPLITS, DPF_STAT_SPLITS}) get(f, h->stats[i]);
            for (double& v : h->occupancy) get(f, v);
            h->num_nodes = num_nodes;
            h->node_cap = num_nodes + num_nodes / 4 + 4096;
            const int64_t slots = (int64_t)num_nodes * h->tp.W;
            h->child_ptr.reserve((size_t)h->node_cap * h->tp.W);
            h->child_cnt.reserve((size_t)h->node_cap * h->tp.W);
            h->node_table.reserve((size_t)h->node_cap);
            h->ids_sorted.reserve((size_t)arena_used + (size_t)(arena_used / 4) + (1 << 20) + 64);
            file_to_dev(h, f, h->child_ptr.p, (size_t)slots);
            file_to_dev(h, f, h->child_cnt.p, (size_t)slots);
            file_to_dev(h, f, h->node_table.p, (size_t)num_nodes);
            file_to_dev(h, f, h->ids_sorted.p, (size_t)arena_used);
            h->arena_used = arena_used;
            h->table_base.reserve((size_t)L + 1);
            h2d(h, h->table_base.p, h->h_table_base.data(), (size_t)L + 1);
            DPF_CUDA(cudaStreamSynchronize(h->stream));
            h->stats[DPF_STAT_DIR_NODES] = num_nodes;
            rebuild_leaf_table(h);
            h->fitted = true;
        }
        if (h->dense) build_compact_store(h);
        h->stats[DPF_STAT_SIZE] = h->n;
        DPF_CUDA(cudaStreamSynchronize(h->stream));
    });
    if (rc != DPF_OK) {
        dpf_destroy(h);
        return rc;
    }
    *out = h;
    return DPF_OK;
}

int dpf_merge_topk_dev(dpf_handle h, const int32_t* gathered_ids_dev, const double* gathered_scores_dev, int32_t G,
                       int64_t nq, int32_t topk, int32_t metric, int32_t* ids_out_dev, double* score_out_dev) {
    return guarded(h, [&] {
        DPF_REQUIRE(gathered_ids_dev && gathered_scores_dev && ids_out_dev && score_out_dev, DPF_ERR_INVALID, "null buffer");
        merge_topk(h, gathered_ids_dev, gathered_scores_dev, G, nq, topk, metric, ids_out_dev, score_out_dev);
    });
}

int dpf_dump_buckets(dpf_handle h, int32_t table, int64_t* nbuckets_out, int64_t* nids_out, int32_t* desc_out,
                     int64_t* off_out, int32_t* ids_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(table >= 0 && table < h->cfg.L, DPF_ERR_INVALID, "table out of range");
        const TreeParams& tp = h->tp;
        const int64_t slots = (int64_t)h->num_nodes * tp.W;
        std::vector<int32_t> cp((size_t)slots), cc((size_t)slots);
        d2h(h, cp.data(), h->child_ptr.p, (size_t)slots);
        d2h(h, cc.data(), h->child_cnt.p, (size_t)slots);
        // the table's own range, and behind all tables the buckets an incremental put has moved (incremental.cu)
        const int64_t tb = h->h_table_base[table], tn = h->arena_used - tb;
        std::vector<int32_t> ids((size_t)tn);
        d2h(h, ids.data(), h->ids_sorted.p + tb, (size_t)tn);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        int64_t nbk = 0, nid = 0;
        struct Frame { int32_t node; int level; int64_t path; int next; };
        for (int r = 0; r < tp.R; ++r) {
            std::vector<Frame> st;
            st.push_back(Frame{table * tp.R + r, tp.MAXL, 0, 0});
            while (!st.empty()) {
                Frame& f = st.back();
                if (f.next >= tp.W) { st.pop_back(); continue; }
                const int slot = f.next++;
                const int64_t idx = (int64_t)f.node * tp.W + slot;
                const int32_t c = cc[(size_t)idx], p = cp[(size_t)idx];
                if (c == 0) continue;
                const int64_t path = (f.path << tp.nb) | slot;
                if (c < 0) { const Frame nf{p, f.level - 1, path, 0}; st.push_back(nf); continue; }
                if (desc_out) { desc_out[nbk * 3] = r; desc_out[nbk * 3 + 1] = f.level; desc_out[nbk * 3 + 2] = (int32_t)path; }
                if (off_out) off_out[nbk] = nid;
                if (ids_out) std::memcpy(ids_out + nid, ids.data() + p, (size_t)c * sizeof(int32_t));
                nid += c;
                nbk++;
            }
        }
        if (off_out) off_out[nbk] = nid;
        if (nbuckets_out) *nbuckets_out = nbk;
        if (nids_out) *nids_out = nid;
    });
}

int dpf_debug_leaf_pairs(dpf_handle h, int64_t* nleaves_out, uint32_t* pair_off_out, int32_t* leaf_len_out) {
    return guarded(h, [&] {
        require_ready(h, true);
        DPF_REQUIRE(nleaves_out, DPF_ERR_INVALID, "null buffer");
        *nleaves_out = h->num_leaves;
        if (!h->leaf_table || h->num_leaves == 0) return;
        if (pair_off_out) d2h(h, pair_off_out, h->leaf_off.p, (size_t)h->num_leaves + 1);
        if (leaf_len_out) d2h(h, leaf_len_out, h->leaf_len.p, (size_t)h->num_leaves);
        DPF_CUDA(cudaStreamSynchronize(h->stream));
    });
}

int dpf_debug_tc_diag(dpf_handle h, uint64_t* out8) {
    return guarded(h, [&] {
        DPF_REQUIRE(out8, DPF_ERR_INVALID, "null buffer");
        DPF_CUDA(cudaStreamSynchronize(h->stream));
        unsigned long long v[24];
        tc_diag_read(v);
        for (int i = 0; i < 24; ++i) out8[i] = v[i];
    });
}

int dpf_stats(dpf_handle h, int64_t* stats_out, double* occupancy_out) {
    return guarded(h, [&] {
        DPF_REQUIRE(stats_out, DPF_ERR_INVALID, "null buffer");
        h->stats[DPF_STAT_SIZE] = h->n;
        {   // the counters of the last batch live on the device (a batch never synchronises to report them)
            int32_t c[CTR_COUNT];
            DPF_CUDA(cudaMemcpyAsync(c, h->counters.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
            DPF_CUDA(cudaStreamSynchronize(h->stream));
            auto u64 = [&](int slot) { unsigned long long v; memcpy(&v, c + slot, sizeof(v)); return (int64_t)v; };
            h->stats[DPF_STAT_NLZ_GT28] = u64(CTR_STAT_NLZ);
            h->stats[DPF_STAT_LAST_CANDIDATES] = u64(CTR_STAT_UNIQUE);
            h->stats[DPF_STAT_LAST_CAND_WITH_DUPS] = u64(CTR_ENTRIES);
            h->stats[DPF_STAT_BM_PAIRS] = u64(CTR_BM_PAIRS_TOTAL);
            h->stats[DPF_STAT_BM_RUNS] = u64(CTR_BM_STAT);
            h->stats[DPF_STAT_BM_ROWS_STAGED] = u64(CTR_BM_STAT + 2);
            h->stats[DPF_STAT_BM_SURVIVORS] = u64(CTR_BM_STAT + 4);
            h->stats[DPF_STAT_BM_DIRECT] = c[CTR_DIRECT];
            h->stats[DPF_STAT_NEAR_ZERO_FIXUPS] = u64(CTR_FIX_TOTAL);
        }
        h->stats[DPF_STAT_KERNEL_LAUNCHES] = (int64_t)g_launches;
        for (int i = 0; i < DPF_STAT_COUNT; ++i) stats_out[i] = h->stats[i];
        if (occupancy_out)
            for (size_t i = 0; i < h->occupancy.size(); ++i) occupancy_out[i] = h->occupancy[i];
    });
}

int dpf_set_profiling(dpf_handle h, int32_t enable) {
    return guarded(h, [&] { h->profiling = enable != 0; });
}

int dpf_set_debug_option(dpf_handle h, int32_t option, int64_t value) {
    return guarded(h, [&] {
        DPF_REQUIRE(option >= 0 && option < DPF_DBG_COUNT, DPF_ERR_INVALID, "unknown debug option");
        h->dbg[option] = value;
    });
}

int dpf_stage_times_ms(dpf_handle h, float* ms_out) {
    return guarded(h, [&] {
        DPF_REQUIRE(ms_out, DPF_ERR_INVALID, "null buffer");
        end_profile(h);
        for (int i = 0; i < DPF_T_COUNT; ++i) ms_out[i] = h->stage_ms[i];
    });
}

}  // extern "C"
