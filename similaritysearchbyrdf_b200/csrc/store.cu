// store.cu — compact vector store.
//
// The reference keeps every vector as a double[] in its dataTable (vectorIdToVector, DensevectorRDFInit.scala:35-36,
// RandomDrawTreeMap.java:871-924) and the re-rank reads d doubles per candidate (DensevectorRDFInit.scala:480-490).
// On the GPU the re-rank is bound by the bytes a candidate row costs in HBM, and the datasets this index is built
// for are not FP64 at the source: SIFT descriptors are bytes, GIST/Deep/GloVe are float32 (fvecs).  After a dense
// fit one pass over the store checks whether EVERY value survives the round trip double -> T -> double for
// T = uint8 and T = float; if one does, a copy of the store in the narrowest such T is kept (row pitch padded to 16
// bytes with zeros for the bulk-copy engine) and the re-rank kernels read that copy, widening each value back to the
// identical double in registers.  Nothing lossy is ever stored: a single value that does not round-trip (including
// -0.0 for uint8 and any NaN) keeps the store in FP64.
#include "common.cuh"

namespace dpf {

// flags bit 0: some value is not a uint8; bit 1: some value is not a float
__global__ void __launch_bounds__(256)
k_scan_representable(const double* __restrict__ X, int64_t total, int* __restrict__ flags) {
    int f = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int it = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride, ++it) {
        const double v = X[i];
        const long long bits = __double_as_longlong(v);
        const double r8 = (double)(unsigned char)min(max(v, 0.0), 255.0);     // NaN -> 0 by the min/max
        if (__double_as_longlong(r8) != bits) f |= 1;                         // bit compare: -0.0 and NaN fail
        const double r32 = (double)(float)v;
        if (__double_as_longlong(r32) != bits) f |= 2;
        if ((it & 15) == 15 && f == 3) break;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// one thread per group of 4 columns of one row; columns >= d of the padded row are written as 0
template <class T>
__global__ void __launch_bounds__(256)
k_narrow_rows(const double* __restrict__ X, int64_t n, int d, int groups /* padded columns / 4 */, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * groups) return;
    const int64_t row = i / groups;
    const int c0 = (int)(i % groups) * 4;
    const double* x = X + row * d;
    T v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = c0 + e < d ? (T)x[c0 + e] : (T)0;
    T* o = out + (row * groups + i % groups) * 4;
    if constexpr (sizeof(T) == 1) {
        *reinterpret_cast<uchar4*>(o) = make_uchar4(v[0], v[1], v[2], v[3]);
    } else {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// The byte copy written WHILE the check runs (one pass over the FP64 rows instead of two: 0.27 + 0.16 ms -> 0.2 ms per 1M x 128);
// flags as in k_scan_representable.  A CTA that starts after another one has seen a value that is not a byte only keeps
// checking (the copy is going to be dropped).
__global__ void __launch_bounds__(256)
k_narrow_check_u8(const double* __restrict__ X, int64_t n, int d, int groups /* padded columns / 4 */,
                  unsigned char* __restrict__ out, int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int f = 0;
    if (i < n * groups) {
        const int64_t row = i / groups;
        const int c0 = (int)(i % groups) * 4;
        const double* x = X + row * d;
        unsigned char v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = 0;
            if (c0 + e < d) {
                const double val = x[c0 + e];
                const long long bits = __double_as_longlong(val);
                v[e] = (unsigned char)min(max(val, 0.0), 255.0);                  // NaN -> 0 by the min/max
                if (__double_as_longlong((double)v[e]) != bits) f |= 1;           // bit compare: -0.0 and NaN fail
                if (__double_as_longlong((double)(float)val) != bits) f |= 2;
            }
        }
        *reinterpret_cast<uchar4*>(out + (row * groups + i % groups) * 4) = make_uchar4(v[0], v[1], v[2], v[3]);
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// rows [n_old, n_old + m) join the compact store when they fit its element type (one pass over the NEW rows only);
// false = the caller rebuilds the store (the new rows need a wider type, or there is no room)
bool append_compact_store(dpf_index* h, int64_t n_old, int64_t m) {
    if (h->Xc_kind == DPF_STORE_KIND_F64) return h->store_mode == DPF_STORE_F64_ONLY || h->dbg[DPF_DBG_STORE] == 1 || n_old > 0;
    const int d = h->cfg.d;
    const int64_t row_bytes = h->Xc_row_bytes;
    if ((size_t)((n_old + m) * row_bytes) > h->Xc.cap) return false;
    cudaStream_t st = h->stream;
    int* flags = h->counters.p + CTR_STORE_FLAGS;
    DPF_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), st));
    const int64_t total = m * d;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)h->num_sms * 16);
    k_scan_representable<<<grid, 256, 0, st>>>(h->Xdev + n_old * d, total, flags); DPF_LAUNCHED();
    int f = 3;
    DPF_CUDA(cudaMemcpyAsync(&f, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    const int bad = h->Xc_kind == DPF_STORE_KIND_U8 ? 1 : 2;
    if (f & bad) return false;
    const int sz = h->Xc_kind == DPF_STORE_KIND_U8 ? 1 : 4;
    const int groups = (int)(row_bytes / sz / 4);
    const unsigned g2 = (unsigned)((m * groups + 255) / 256);
    if (h->Xc_kind == DPF_STORE_KIND_U8)
        k_narrow_rows<unsigned char><<<g2, 256, 0, st>>>(h->Xdev + n_old * d, m, d, groups, h->Xc.p + n_old * row_bytes);
    else
        k_narrow_rows<float><<<g2, 256, 0, st>>>(h->Xdev + n_old * d, m, d, groups, reinterpret_cast<float*>(h->Xc.p + n_old * row_bytes));
    DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
    return true;
}

void build_compact_store(dpf_index* h) {
    h->Xc_kind = DPF_STORE_KIND_F64;
    h->Xc_row_bytes = (int64_t)h->cfg.d * 8;
    h->stats[DPF_STAT_STORE_KIND] = DPF_STORE_KIND_F64;
    h->stats[DPF_STAT_STORE_ROW_BYTES] = h->Xc_row_bytes;
    if (h->store_mode == DPF_STORE_F64_ONLY || h->dbg[DPF_DBG_STORE] == 1 || !h->dense || !h->Xdev || h->n == 0) {
        h->Xc.release();
        return;
    }
    StageTimer tm(h, DPF_T_NARROW);
    cudaStream_t st = h->stream;
    const int d = h->cfg.d;
    const int64_t total = h->n * d;
    int* flags = h->counters.p + 40;
    DPF_CUDA(cudaMemsetAsync(flags, 0, sizeof(int), st));
    const bool force32 = h->dbg[DPF_DBG_STORE] == 2;
    // room for the byte copy is normally there already (taken at the start of the fit): write it while checking
    const int cols8 = (d + 15) / 16 * 16;
    const bool fused = !force32 && h->Xc.cap >= (size_t)(h->n * (int64_t)cols8);
    if (fused) {
        const int64_t threads = h->n * (cols8 / 4);
        k_narrow_check_u8<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(h->Xdev, h->n, d, cols8 / 4, h->Xc.p, flags); DPF_LAUNCHED();
    } else {
        const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)h->num_sms * 16);
        k_scan_representable<<<grid, 256, 0, st>>>(h->Xdev, total, flags); DPF_LAUNCHED();
    }
    int f = 3;
    DPF_CUDA(cudaMemcpyAsync(&f, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    DPF_CUDA(cudaStreamSynchronize(st));
    // bytes pay (k_score_u8*: 1/8 of the traffic, integer tensor pipe).  Floats are only kept on request
    // (DPF_STORE_NARROWEST, or DPF_STORE=f32 which also skips the byte check): the FP64 tensor-pipe kernels are bound
    // by instruction issue, not by the bytes of a row, and the extra float -> double conversions make them slower
    // (measured: 11.2 ms against 10.2 ms per 10k queries at d = 128).
    int kind = DPF_STORE_KIND_F64;
    if (!(f & 1) && !force32) kind = DPF_STORE_KIND_U8;
    else if (!(f & 2) && (force32 || h->store_mode == DPF_STORE_NARROWEST)) kind = DPF_STORE_KIND_F32;
    if (kind == DPF_STORE_KIND_F64) {
        h->Xc.release();
        return;
    }
    const int sz = kind == DPF_STORE_KIND_U8 ? 1 : 4;
    const int cols = (d * sz + 15) / 16 * 16 / sz;               // padded columns: the row pitch is a multiple of 16 B
    const int64_t row_bytes = (int64_t)cols * sz;
    try {
        h->Xc.reserve((size_t)(h->n * row_bytes));
    } catch (const Error& err) {
        if (err.code != DPF_ERR_NOMEM) throw;
        cudaGetLastError();                                      // no room for the copy: stay on the FP64 rows
        return;
    }
    const int groups = cols / 4;
    const int64_t threads = h->n * groups;
    const unsigned g2 = (unsigned)((threads + 255) / 256);
    if (kind == DPF_STORE_KIND_U8) {
        if (!fused) { k_narrow_rows<unsigned char><<<g2, 256, 0, st>>>(h->Xdev, h->n, d, groups, h->Xc.p); DPF_LAUNCHED(); }
    } else {
        k_narrow_rows<float><<<g2, 256, 0, st>>>(h->Xdev, h->n, d, groups, reinterpret_cast<float*>(h->Xc.p)); DPF_LAUNCHED();
    }
    DPF_CUDA(cudaGetLastError());
    h->Xc_kind = kind;
    h->Xc_row_bytes = row_bytes;
    h->stats[DPF_STAT_STORE_KIND] = kind;
    h->stats[DPF_STAT_STORE_ROW_BYTES] = row_bytes;
}

}  // namespace dpf
