// hash.cu — K1/K1s/K2: batched LSH compound-key evaluation and partition ids.
//
// Replaces, for a whole batch at once (reference file:line, relative to /root/reference):
//   SimilarityCalculator.fastCalculateSimilarity          src/main/scala/mclab/lsh/vector/SimilarityCalculator.scala:9-49
//   AngleHashChain.compute / PStableHashChain.compute     .../hashFamilies/AngleHashFamily.scala:184-219, PStableHashFamily.scala:122-177
//   LSH.calculateIndex (+ typeOfIndex transforms)         src/main/scala/mclab/lsh/LSH.scala:93-166, Sampling.scala:32-39,
//                                                         significantBits.scala:11-67,113-127
//   LocalitySensitivePartitioner.getPartition             src/main/scala/mclab/utils/Partitioner.scala:40-64
//
// Data flow (dense, angle family):
//   X[n x d] (HBM, FP64 row-major)  --k_project_dmma-->  sign bits S[n x PW] (+ list of near-zero projections)
//                                   --k_fixup_exact-->   near-zero projections recomputed in the reference's
//                                                        summation order (sequential, unfused) => signs are exact
//   S, chain[L x k], Ap[L x pb x 32] --k_pack_keys-->    keys[L x n] int32, pids[L x n] uint8
// The FP64 contraction runs on the tensor pipe (mma.sync m8n8k4 f64 = DMMA.8x8x4 on sm_100a; tcgen05 has no
// FP64 kind).  The sign/quantise step is fused into the GEMM epilogue: projections are never written to HBM.
#include "common.cuh"

namespace dpf {

// ---------------------------------------------------------------------------------------------------------
// host-side key-transform tables (Sampling.scala:6-11 needs java.util.Random; significantBits.newMethod
// depends only on popcount(key & 0x0fffffff))
// ---------------------------------------------------------------------------------------------------------
struct JRandom {  // java.util.Random's documented 48-bit LCG
    uint64_t s;
    explicit JRandom(int64_t seed) : s(((uint64_t)seed ^ 0x5DEECE66DULL) & 0xFFFFFFFFFFFFULL) {}
    int32_t next31() {
        s = (s * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
        return (int32_t)(s >> 17);
    }
    int32_t next_int(int32_t bound) {
        if ((bound & (bound - 1)) == 0) return (int32_t)(((int64_t)bound * next31()) >> 31);
        for (;;) {
            int32_t r = next31();
            int32_t m = r % bound;
            if ((int64_t)r - m + (bound - 1) < 2147483648LL) return m;
        }
    }
};

struct KeyTransformTables {
    int32_t sigma[32];      // sampling bit order
    int32_t new_index[29];  // angleNewMethod label by popcount of the low 28 bits
};

static KeyTransformTables make_transform_tables() {
    KeyTransformTables t;
    for (int i = 0; i < 32; ++i) t.sigma[i] = i;
    JRandom rnd(88387);  // `new Sampling(88387)` (LSH.scala:21)
    for (int n = 32; n >= 2; --n) {  // scala.util.Random.shuffle (2.10): swap(n-1, nextInt(n))
        int j = rnd.next_int(n);
        int32_t tmp = t.sigma[n - 1];
        t.sigma[n - 1] = t.sigma[j];
        t.sigma[j] = tmp;
    }
    const double metric[9] = {16.0, 25.0, 33.0, 39.0, 46.0, 52.0, 58.0, 66.0, 72.0};
    for (int pc = 0; pc <= 28; ++pc) {
        // angleDistance (significantBits.scala:100-110): acos(v1.base / (|base| |v1|)) in degrees
        double ang = acos((double)pc / (sqrt(28.0) * sqrt((double)pc))) * 360 / 2 / M_PI;
        int idx = 0;
        while (idx < 9 && ang > metric[idx]) idx++;
        t.new_index[pc] = idx;
    }
    return t;
}

__constant__ KeyTransformTables c_tt;

static void upload_transform_tables() {
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && done[dev]) return;
    KeyTransformTables t = make_transform_tables();
    DPF_CUDA(cudaMemcpyToSymbol(c_tt, &t, sizeof(t)));
    if (dev < 64) done[dev] = true;
}

__device__ __forceinline__ int32_t apply_key_transform(int32_t key, int kind) {
    const uint32_t u = (uint32_t)key;
    if (kind == DPF_KEY_SAMPLING) {
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) r |= ((u >> c_tt.sigma[j]) & 1u) << (31 - j);
        return (int32_t)r;
    }
    if (kind == DPF_KEY_CONTINUE_BITS) {
        // run-length recoding of the low 28 bits with thresholds 6,4,2,1 (significantBits.scala:11-67)
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int count = 0;
        for (int i = 0; i < 28; ++i) {
            const bool one = (u >> i) & 1u;
            if (one) count++;
            if (!one || i == 27) {
                if (count >= 6) { a0++; a1++; a2++; a3++; }
                else if (count >= 4) { a1++; a2++; a3++; }
                else if (count >= 2) { a2++; a3++; }
                else if (count >= 1) { a3++; }
                count = 0;
            }
        }
        return (int32_t)((a3 << 21) + (a2 << 14) + (a1 << 7) + a0 + ((u >> 28) << 28));
    }
    if (kind == DPF_KEY_ANGLE_NEW) {
        const uint32_t label = (uint32_t)c_tt.new_index[__popc(u & 0x0FFFFFFFu)];
        return (int32_t)((u & 0x7Fu) + (((u >> 7) & 0x7Fu) << 7) + (label << 14) + (((u >> 21) & 0x7Fu) << 21) +
                         (((u >> 28) & 0x7Fu) << 28));
    }
    return key;
}

// ---------------------------------------------------------------------------------------------------------
// K1 (tensor pipe): sign bits of X . A^T with DMMA, near-zero projections listed for exact recomputation
// ---------------------------------------------------------------------------------------------------------
constexpr int BM = 64;          // vectors per CTA
constexpr int BN = 64;          // hash functions per CTA (two sign words)
constexpr int BK = 32;          // reduction slice
constexpr int LDS_ = BK + 4;    // smem row pitch in doubles: (4*g + t) mod 16 distinct per half-warp => conflict-free LDS.64
constexpr int K1_THREADS = 256; // 8 warps: 2 (rows) x 4 (functions); warp tile 32 x 16

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(K1_THREADS, 2)
k_project_dmma(const double* __restrict__ X, const double* __restrict__ A, const double* __restrict__ Anorm,
               int64_t n, int d, int P, int PW, int nct, double coef, uint32_t* __restrict__ S,
               int2* __restrict__ fix_list, int* __restrict__ fix_count, int fix_cap) {
    extern __shared__ double smem[];
    double* Xs = smem;                       // [2][BM][LDS_]
    double* As = smem + 2 * BM * LDS_;       // [2][BN][LDS_]
    __shared__ double rown[BM];
    __shared__ uint32_t sbits[BM][BN / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;            // warp tile origin: rows wm*32, cols wn*16
    // 1-D grid, column tile fastest: the CTAs that share an X tile are co-resident, so X is read from HBM once
    const int64_t row0 = (int64_t)(blockIdx.x / nct) * BM;
    const int col0 = (int)(blockIdx.x % nct) * BN;

    if (tid < BM * (BN / 32)) (&sbits[0][0])[tid] = 0u;

    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double nsq[BM / 8];
#pragma unroll
    for (int i = 0; i < BM / 8; ++i) nsq[i] = 0.0;

    // register-staged global loads: warp w loads rows i*8+w, lane = k within the slice (256 B per warp load)
    double xr[BM / 8], ar[BN / 8];
    auto gload = [&](int k0) {
        const int k = k0 + lane;
        const bool kin = k < d;
#pragma unroll
        for (int i = 0; i < BM / 8; ++i) {
            const int64_t r = row0 + i * 8 + warp;
            xr[i] = (kin && r < n) ? __ldg(X + r * d + k) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
            const int p = col0 + i * 8 + warp;
            ar[i] = (kin && p < P) ? __ldg(A + (int64_t)p * d + k) : 0.0;
        }
    };
    auto sstore = [&](int buf) {
        double* xs = Xs + buf * BM * LDS_;
        double* as = As + buf * BN * LDS_;
#pragma unroll
        for (int i = 0; i < BM / 8; ++i) {
            xs[(i * 8 + warp) * LDS_ + lane] = xr[i];
            nsq[i] = fma(xr[i], xr[i], nsq[i]);
        }
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) as[(i * 8 + warp) * LDS_ + lane] = ar[i];
    };

    const int nk = (d + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
        const double* xs = Xs + buf * BM * LDS_ + (wm * 32 + g) * LDS_ + t;
        const double* as = As + buf * BN * LDS_ + (wn * 16 + g) * LDS_ + t;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[i * 8 * LDS_ + ks];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = as[j * 8 * LDS_ + ks];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }

    // row norms: lane-partial sums of squares -> ||x_r||
#pragma unroll
    for (int i = 0; i < BM / 8; ++i) {
        double v = nsq[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) rown[i * 8 + warp] = sqrt(v);
    }
    __syncthreads();

    // epilogue: sign + near-zero test, fused; thread holds rows wm*32+i*8+g, cols wn*16+j*8+2t+{0,1}
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = wm * 32 + i * 8 + g;
        const int64_t r = row0 + rl;
        const double rn = rown[rl];
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cl = wn * 16 + j * 8 + 2 * t + e;
                const int c = col0 + cl;
                const double v = acc[i][j][e];
                if (r < n && c < P) {
                    if (!(v <= 0.0)) bits |= 1u << (cl & 31);     // sign(x) = if (x <= 0) 0 else 1
                    const double thr = coef * rn * __ldg(Anorm + c);
                    // not provably on one side of zero (or norms degenerate): recompute in reference order
                    if (!(fabs(v) > thr) || !(rn > 1e-140)) {
                        const int idx = atomicAdd(fix_count, 1);
                        if (idx < fix_cap) fix_list[idx] = make_int2((int)r, c);
                    }
                }
            }
        bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
        if (t == 0 && bits) atomicOr(&sbits[rl][(wn * 16) >> 5], bits);
    }
    __syncthreads();
    for (int i = tid; i < BM * (BN / 32); i += K1_THREADS) {
        const int rl = i / (BN / 32), w = i % (BN / 32);
        const int64_t r = row0 + rl;
        const int gw = col0 / 32 + w;
        if (r < n && gw < PW) S[r * PW + gw] = (&sbits[0][0])[i];
    }
}

// exact recomputation of listed (vector, function) pairs in the reference's order:
// s = s + a[j]*x[j], j ascending, product and sum rounded separately (SimilarityCalculator.scala:45-47)
__global__ void k_fixup_exact(const double* __restrict__ X, const double* __restrict__ A, int d, int PW,
                              const int2* __restrict__ fix_list, const int* __restrict__ fix_count, int fix_cap,
                              uint32_t* __restrict__ S, unsigned long long* __restrict__ total) {
    const int m = min(*fix_count, fix_cap);
    if (m > 0 && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(total, (unsigned long long)m);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const int2 e = fix_list[i];
        const double* x = X + (int64_t)e.x * d;
        const double* a = A + (int64_t)e.y * d;
        double s = 0.0;
        for (int j = 0; j < d; ++j) s = __dadd_rn(s, __dmul_rn(a[j], x[j]));
        uint32_t* w = S + (int64_t)e.x * PW + (e.y >> 5);
        const uint32_t m1 = 1u << (e.y & 31);
        if (!(s <= 0.0)) atomicOr(w, m1);
        else atomicAnd(w, ~m1);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K1 (CUDA cores, exact by construction): every projection accumulated in reference order.  Used for the
// pStable family (quantised value needed, not just the sign) and as the cross-check of the DMMA path.
// ---------------------------------------------------------------------------------------------------------
constexpr int EX_T = 64, EX_K = 16;
template <bool PSTABLE>
__global__ void __launch_bounds__(256)
k_project_exact(const double* __restrict__ X, const double* __restrict__ A, int64_t n, int d, int P, int PW, int nct,
                const double* __restrict__ fb, const int32_t* __restrict__ fw, uint32_t* __restrict__ S,
                int32_t* __restrict__ PQ) {
    __shared__ double Xs[EX_T][EX_K + 1];
    __shared__ double As[EX_T][EX_K + 1];
    __shared__ uint32_t sbits[EX_T][EX_T / 32];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)(blockIdx.x / nct) * EX_T;
    const int col0 = (int)(blockIdx.x % nct) * EX_T;
    if (tid < EX_T * (EX_T / 32)) (&sbits[0][0])[tid] = 0u;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < d; k0 += EX_K) {
        for (int i = tid; i < EX_T * EX_K; i += 256) {
            const int r = i / EX_K, k = i % EX_K;
            const int64_t gr = row0 + r;
            const int gp = col0 + r;
            Xs[r][k] = (gr < n && k0 + k < d) ? X[gr * d + k0 + k] : 0.0;
            As[r][k] = (gp < P && k0 + k < d) ? A[(int64_t)gp * d + k0 + k] : 0.0;
        }
        __syncthreads();
        const int kmax = min(EX_K, d - k0);
        for (int k = 0; k < kmax; ++k) {   // ascending j, one rounding per product and per sum
            double xv[4], av[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = Xs[ty + 16 * i][k];
#pragma unroll
            for (int j = 0; j < 4; ++j) av[j] = As[tx + 16 * j][k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(av[j], xv[i]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = ty + 16 * i;
        const int64_t r = row0 + rl;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cl = tx + 16 * j;
            const int c = col0 + cl;
            if (r < n && c < P) {
                if (PSTABLE) {
                    // ((sum + b) / w).toInt : truncation toward zero, saturating, NaN -> 0 (PStableHashFamily.scala:130)
                    PQ[r * P + c] = __double2int_rz(__ddiv_rn(__dadd_rn(acc[i][j], fb[c]), (double)fw[c]));
                } else if (!(acc[i][j] <= 0.0)) {
                    atomicOr(&sbits[rl][cl >> 5], 1u << (cl & 31));
                }
            }
        }
    }
    if (!PSTABLE) {
        __syncthreads();
        for (int i = tid; i < EX_T * (EX_T / 32); i += 256) {
            const int rl = i / (EX_T / 32), w = i % (EX_T / 32);
            const int64_t r = row0 + rl;
            const int gw = col0 / 32 + w;
            if (r < n && gw < PW) S[r * PW + gw] = (&sbits[0][0])[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K1s: CSR variant.  One warp per sparse vector, lanes across hash functions (feature-major At[D x ld] so a
// non-zero's P coefficients are one contiguous, coalesced read); non-zeros visited in ascending index order
// with unfused mul/add = the reference's BitSet-intersection order (SimilarityCalculator.scala:9-27).
// ---------------------------------------------------------------------------------------------------------
template <int PPL, bool PSTABLE>  // PPL = functions per lane (P <= 32*PPL)
__global__ void __launch_bounds__(256)
k_project_csr(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const double* __restrict__ val,
              const double* __restrict__ At, int ld, int64_t n, int P, int PW, const double* __restrict__ fb,
              const int32_t* __restrict__ fw, uint32_t* __restrict__ S, int32_t* __restrict__ PQ) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    double acc[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) acc[i] = 0.0;
    const int64_t b = ptr[row], e = ptr[row + 1];
    for (int64_t base = b; base < e; base += 32) {
        const int cnt = (int)min((int64_t)32, e - base);
        int32_t myi = 0;
        double myv = 0.0;
        if (lane < cnt) { myi = __ldg(idx + base + lane); myv = __ldg(val + base + lane); }
        for (int q = 0; q < cnt; ++q) {
            const int32_t j = __shfl_sync(0xffffffffu, myi, q);
            const double v = __shfl_sync(0xffffffffu, myv, q);
            const double* arow = At + (int64_t)j * ld;
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
                const int p = i * 32 + lane;
                if (p < P) {
                    const double a = __ldg(arow + p);
                    // a zero coefficient is outside the function's support (AngleHashFamily.scala:48)
                    if (a != 0.0) acc[i] = __dadd_rn(acc[i], __dmul_rn(a, v));
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
        const int p = i * 32 + lane;
        if (PSTABLE) {
            if (p < P) PQ[row * P + p] = __double2int_rz(__ddiv_rn(__dadd_rn(acc[i], fb[p]), (double)fw[p]));
        } else {
            const uint32_t word = __ballot_sync(0xffffffffu, p < P && !(acc[i] <= 0.0));
            if (lane == 0 && i < PW) S[row * PW + i] = word;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// key packing (AngleHashFamily.scala:187-195 / PStableHashFamily.scala:122-143), key transform
// (LSH.scala:152-161) and K2 partition id (Partitioner.scala:40-64), one thread per (vector, table)
// ---------------------------------------------------------------------------------------------------------
// With the pStable family the partitioner's chain is a pStable chain too (its LSH is built from the main configuration,
// DensevectorRDFInit.scala:63-70): the pb sums are quantised ((sum + b) / w).toInt and packed by Arrays.hashCode
// (PStableHashFamily.scala:155-177); pbb / pbw are that chain's b and w, null for the angle family.
__device__ __forceinline__ int32_t partition_id_dev(int32_t key, const double* ap /* pb x 32, smem */, int pb,
                                                    int transform, const double* pbb = nullptr, const int32_t* pbw = nullptr) {
    if (pb == 0) return 0;
    uint32_t r = pbb ? 1u : 0u;
    for (int j = 0; j < pb; ++j) {
        double s = 0.0;
#pragma unroll 8
        for (int i = 0; i < 32; ++i)   // a * 1.0 for the set bits, ascending i; a clear bit adds +0.0, which leaves
            s = __dadd_rn(s, (((uint32_t)key >> i) & 1u) ? ap[j * 32 + i] : 0.0);   // the value (and s <= 0) unchanged
        if (pbb) {
            const int32_t q = __double2int_rz(__ddiv_rn(__dadd_rn(s, pbb[j]), (double)pbw[j]));
#pragma unroll
            for (int sh = 24; sh >= 0; sh -= 8) r = 31u * r + (uint32_t)(int32_t)(int8_t)((uint32_t)q >> sh);
        } else {
            r = (r << 1) | (!(s <= 0.0) ? 1u : 0u);
        }
    }
    int32_t pk = pbb ? (int32_t)r : (int32_t)(r << (32 - pb));
    pk = apply_key_transform(pk, transform);
    return (int32_t)((uint32_t)pk >> (32 - pb));
}

template <bool PSTABLE>
__global__ void __launch_bounds__(256)
k_pack_keys(const uint32_t* __restrict__ S, const int32_t* __restrict__ PQ, const int32_t* __restrict__ chain,
            const double* __restrict__ Ap, const double* __restrict__ Apb, const int32_t* __restrict__ Apw, int64_t n, int P,
            int PW, int k, int pb, int transform, int32_t* __restrict__ keys, uint8_t* __restrict__ pids, int64_t ld) {
    __shared__ int32_t ch[kMaxChain];
    __shared__ double ap[kMaxPb * 32];
    __shared__ double pbb[kMaxPb];
    __shared__ int32_t pbw[kMaxPb];
    const int t = blockIdx.y;
    if (threadIdx.x < k) ch[threadIdx.x] = chain[t * k + threadIdx.x];
    for (int i = threadIdx.x; i < pb * 32; i += blockDim.x) ap[i] = Ap[(int64_t)t * pb * 32 + i];
    if (Apb && threadIdx.x < pb) {
        pbb[threadIdx.x] = Apb[t * pb + threadIdx.x];
        pbw[threadIdx.x] = Apw[t * pb + threadIdx.x];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key;
    if (PSTABLE) {
        key = 1u;   // java.util.Arrays.hashCode(byte[]) over the big-endian bytes of the k quantised values
        for (int b = 0; b < k; ++b) {
            const int32_t q = PQ[i * P + ch[b]];
#pragma unroll
            for (int s = 24; s >= 0; s -= 8) key = 31u * key + (uint32_t)(int32_t)(int8_t)((uint32_t)q >> s);
        }
    } else {
        const uint32_t* srow = S + i * PW;
        key = 0u;
        if (PW <= 4) {   // the usual case (P <= 128 distinct functions): the row's sign words live in registers
            const uint32_t w0 = srow[0], w1 = PW > 1 ? srow[1] : 0u, w2 = PW > 2 ? srow[2] : 0u, w3 = PW > 3 ? srow[3] : 0u;
            for (int b = 0; b < k; ++b) {
                const int c = ch[b], wi = c >> 5;
                const uint32_t w = wi == 0 ? w0 : (wi == 1 ? w1 : (wi == 2 ? w2 : w3));
                key = (key << 1) | ((w >> (c & 31)) & 1u);
            }
        } else {
            for (int b = 0; b < k; ++b) {
                const int c = ch[b];
                key = (key << 1) | ((srow[c >> 5] >> (c & 31)) & 1u);
            }
        }
        if (k < 32) key <<= (32 - k);
    }
    const int32_t fk = apply_key_transform((int32_t)key, transform);
    keys[(int64_t)t * ld + i] = fk;
    if (pids) pids[(int64_t)t * ld + i] = (uint8_t)partition_id_dev(fk, ap, pb, transform, Apb ? pbb : nullptr, pbw);
}

// The same, table-driven, for the angle family with P <= 128 sign bits per vector.  k_pack_keys spends ~750 instructions
// per key (a 4-way select per chain bit, a select + FP64 add per partitioner coefficient) and is issue-bound; here a CTA
// first turns its table's chain into 16 byte-indexed lookup tables (sign byte s, value v -> the key bits that byte
// supplies) and its partitioner rows into 4 byte-indexed tables of FP32 partial sums, then packs PACK_KPT keys per
// thread with 16 + 4 shared-memory lookups each.
//
// The partition bit is the SIGN of a sum the reference adds in FP64 in ascending bit order (Partitioner.scala:40-64).
// The FP32 table sum f differs from the real sum T by at most 5 * 2^-24 * A (A = sum |coefficient|: rounding of four
// partials to FP32 and three FP32 adds) and the reference's own sum by at most 31 * 2^-53 * A, so |f| > 8 * 2^-24 * A
// fixes the sign of both; the few keys per million below that redo the row in the reference's order.
constexpr int PACK_KPT = 32;

template <int PBV /* float4 per table entry: 1 for pb <= 4, 2 for pb <= 8 */>
__global__ void __launch_bounds__(256, 4)
k_pack_keys_tab(const uint32_t* __restrict__ S, const int32_t* __restrict__ chain, const double* __restrict__ Ap, int64_t n,
                int PW, int k, int pb, int transform, int32_t* __restrict__ keys, uint8_t* __restrict__ pids, int64_t ld) {
    extern __shared__ __align__(16) unsigned char pack_smem[];
    uint32_t* keytab = reinterpret_cast<uint32_t*>(pack_smem);                       // [PW*4][256]
    float4* ptab = reinterpret_cast<float4*>(pack_smem + (size_t)PW * 4 * 256 * 4);  // [4][256][PBV]
    __shared__ double ap[kMaxPb * 32];
    __shared__ float bound[kMaxPb];
    __shared__ unsigned char item_cnt[16], item[16][kMaxChain];   // per sign byte: (bit in byte) | (chain position << 3)
    const int t = blockIdx.y, tid = threadIdx.x;
    if (tid < 16) item_cnt[tid] = 0;
    for (int i = tid; i < pb * 32; i += 256) ap[i] = Ap[(int64_t)t * pb * 32 + i];
    __syncthreads();
    if (tid == 0)
        for (int b = 0; b < k; ++b) {
            const int c = chain[t * k + b], sb = c >> 3;
            item[sb][item_cnt[sb]++] = (unsigned char)((c & 7) | (b << 3));
        }
    if (tid >= 32 && tid < 32 + pb) {
        const int j = tid - 32;
        double a = 0.0;
        for (int i = 0; i < 32; ++i) a += fabs(ap[j * 32 + i]);
        // (not finite, or beyond what an FP32 partial can hold: no fast path for this row)
        bound[j] = (a < 1e37) ? (float)(a * (8.0 / 16777216.0)) * 1.0001f : __int_as_float(0x7f800000);
    }
    __syncthreads();
#pragma unroll 1
    for (int e = tid; e < PW * 4 * 256; e += 256) {
        const int sb = e >> 8, v = e & 255, cnt = item_cnt[sb];
        uint32_t r = 0;
#pragma unroll 1
        for (int q = 0; q < cnt; ++q) {
            const int it = item[sb][q];
            r |= ((uint32_t)(v >> (it & 7)) & 1u) << (31 - (it >> 3));
        }
        keytab[e] = r;
    }
    {
        float* pf = reinterpret_cast<float*>(ptab);
#pragma unroll 1
        for (int e = tid; e < 4 * 256 * 4 * PBV; e += 256) {   // one (byte, value, partitioner row) per pass
            const int j = e % (4 * PBV), bv = e / (4 * PBV), byte = bv >> 8, v = bv & 255;
            double sum = 0.0;
            if (j < pb) {
#pragma unroll 1
                for (int i = 0; i < 8; ++i)
                    if ((v >> i) & 1) sum = __dadd_rn(sum, ap[j * 32 + byte * 8 + i]);
            }
            pf[e] = (float)sum;
        }
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < PACK_KPT; ++r) {
        const int64_t i = ((int64_t)blockIdx.x * PACK_KPT + r) * 256 + tid;
        if (i >= n) break;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (PW == 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(S + i * 4);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            for (int q = 0; q < PW; ++q) w[q] = S[i * PW + q];
        }
        uint32_t key = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < PW) {
#pragma unroll
                for (int b = 0; b < 4; ++b) key |= keytab[(q * 4 + b) * 256 + ((w[q] >> (8 * b)) & 255u)];
            }
        const int32_t fk = apply_key_transform((int32_t)key, transform);
        keys[(int64_t)t * ld + i] = fk;
        if (pids) {
            float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int e = (b * 256 + (int)(((uint32_t)fk >> (8 * b)) & 255u)) * PBV;
                const float4 x = ptab[e];
                f[0] += x.x; f[1] += x.y; f[2] += x.z; f[3] += x.w;
                if (PBV > 1) {
                    const float4 y = ptab[e + 1];
                    f[4] += y.x; f[5] += y.y; f[6] += y.z; f[7] += y.w;
                }
            }
            uint32_t pr = 0;
#pragma unroll
            for (int j = 0; j < 4 * PBV; ++j)
                if (j < pb) {
                    bool pos = f[j] > 0.f;
                    if (!(fabsf(f[j]) > bound[j])) {
                        double sum = 0.0;
#pragma unroll 1
                        for (int q = 0; q < 32; ++q)
                            sum = __dadd_rn(sum, (((uint32_t)fk >> q) & 1u) ? ap[j * 32 + q] : 0.0);
                        pos = !(sum <= 0.0);
                    }
                    pr = (pr << 1) | (pos ? 1u : 0u);
                }
            int32_t pk = (int32_t)(pr << (32 - pb));
            pk = apply_key_transform(pk, transform);
            pids[(int64_t)t * ld + i] = (uint8_t)((uint32_t)pk >> (32 - pb));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------
void prepare_family(dpf_index* h) {
    upload_transform_tables();
    const int P = h->P, d = h->cfg.d;
    std::vector<double> nrm(P);
    for (int p = 0; p < P; ++p) {
        double s = 0;
        for (int j = 0; j < d; ++j) s += h->hA[(size_t)p * d + j] * h->hA[(size_t)p * d + j];
        nrm[p] = sqrt(s);
    }
    h->Anorm.reserve(P);
    DPF_CUDA(cudaMemcpyAsync(h->Anorm.p, nrm.data(), P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    DPF_CUDA(cudaStreamSynchronize(h->stream));
    {   // function attributes are per device: set for this handle's device (dpf_set_family runs once per handle)
        const int smem = (2 * BM + 2 * BN) * LDS_ * (int)sizeof(double);
        DPF_CUDA(cudaFuncSetAttribute(k_project_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
}

static int64_t hash_chunk_rows(const dpf_index* h) {
    // bound the sign / quantised scratch to ~1 GiB
    const int64_t per_row = (int64_t)h->PW * 4 + (h->cfg.family_kind == DPF_FAMILY_PSTABLE ? (int64_t)h->P * 4 : 0);
    int64_t rows = (1LL << 30) / (per_row > 0 ? per_row : 1);
    if (rows > (1LL << 22)) rows = 1LL << 22;
    if (rows < 1024) rows = 1024;
    return rows;
}

static void pack_launch(dpf_index* h, int64_t n, int32_t* keys_out, uint8_t* pids_out, int64_t ld) {
    StageTimer tm(h, DPF_T_PACK);
    const dim3 grid((unsigned)((n + 255) / 256), h->cfg.L);
    // (a query batch of a few thousand keys is latency, not throughput: there the tables cost more than they save)
    if (h->cfg.family_kind != DPF_FAMILY_PSTABLE && h->PW <= 4 && h->cfg.pb >= 1 && pids_out && n >= 65536 &&
        h->dbg[DPF_DBG_HASH_EXACT] != 2) {
        const dim3 g2((unsigned)((n + 256 * PACK_KPT - 1) / (256 * PACK_KPT)), h->cfg.L);
        const int pbv = h->cfg.pb > 4 ? 2 : 1;
        const size_t smem = (size_t)h->PW * 4 * 256 * 4 + (size_t)4 * 256 * 16 * pbv;
        if (pbv == 1)
            k_pack_keys_tab<1><<<g2, 256, smem, h->stream>>>(h->signs.p, h->chain.p, h->Ap.p, n, h->PW, h->cfg.k, h->cfg.pb,
                                                            h->cfg.key_transform, keys_out, pids_out, ld);
        else
            k_pack_keys_tab<2><<<g2, 256, smem, h->stream>>>(h->signs.p, h->chain.p, h->Ap.p, n, h->PW, h->cfg.k, h->cfg.pb,
                                                            h->cfg.key_transform, keys_out, pids_out, ld);
    } else if (h->cfg.family_kind == DPF_FAMILY_PSTABLE)
        k_pack_keys<true><<<grid, 256, 0, h->stream>>>(nullptr, h->pq.p, h->chain.p, h->Ap.p, h->part_pstable ? h->Apb.p : nullptr,
                                                       h->Apw.p, n, h->P, h->PW, h->cfg.k, h->cfg.pb, h->cfg.key_transform,
                                                       keys_out, pids_out, ld);
    else
        k_pack_keys<false><<<grid, 256, 0, h->stream>>>(h->signs.p, nullptr, h->chain.p, h->Ap.p, nullptr, nullptr, n, h->P,
                                                        h->PW, h->cfg.k, h->cfg.pb, h->cfg.key_transform, keys_out,
                                                        pids_out, ld);
    DPF_LAUNCHED();
    DPF_CUDA(cudaGetLastError());
}

void hash_dense_device(dpf_index* h, const double* Xd, int64_t n, int32_t* keys_out, uint8_t* pids_out, int64_t ld) {
    if (n <= 0) return;
    const int d = h->cfg.d, P = h->P, PW = h->PW;
    const bool pst = h->cfg.family_kind == DPF_FAMILY_PSTABLE;
    const int64_t chunk = hash_chunk_rows(h);
    h->signs.reserve((size_t)std::min(n, chunk) * PW);
    if (pst) h->pq.reserve((size_t)std::min(n, chunk) * P);
    h->counters.reserve(CTR_COUNT);
    if (h->fix_list.cap == 0) h->fix_list.reserve(4 << 20);
    unsigned long long* fix_total = reinterpret_cast<unsigned long long*>(h->counters.p + CTR_FIX_TOTAL);
    const double coef = 4.0 * (double)(d + 8) * 1.1102230246251565e-16;   // 4 (d+8) 2^-53 >= 2 * 2 gamma_d
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t m = std::min(chunk, n - r0);
        const double* Xc = Xd + r0 * d;
        if (pst) {
            StageTimer tm(h, DPF_T_HASH);
            const int nct = (P + EX_T - 1) / EX_T;
            const unsigned grid = (unsigned)(((m + EX_T - 1) / EX_T) * nct);
            k_project_exact<true><<<grid, 256, 0, h->stream>>>(Xc, h->A.p, m, d, P, PW, nct, h->fb.p, h->fw.p, nullptr,
                                                               h->pq.p); DPF_LAUNCHED();
            DPF_CUDA(cudaGetLastError());
        } else {
            for (;;) {
                {
                    StageTimer tm(h, DPF_T_HASH);
                    DPF_CUDA(cudaMemsetAsync(h->counters.p, 0, sizeof(int32_t), h->stream));
                    const int nct = (P + BN - 1) / BN;
                    const unsigned grid = (unsigned)(((m + BM - 1) / BM) * nct);
                    const int smem = (2 * BM + 2 * BN) * LDS_ * (int)sizeof(double);
                    k_project_dmma<<<grid, K1_THREADS, smem, h->stream>>>(Xc, h->A.p, h->Anorm.p, m, d, P, PW, nct, coef,
                                                                          h->signs.p, h->fix_list.p, h->counters.p,
                                                                          (int)h->fix_list.cap); DPF_LAUNCHED();
                    DPF_CUDA(cudaGetLastError());
                }
                if ((size_t)m * (size_t)P <= h->fix_list.cap) {
                    // the list cannot overflow (a query batch, a small fit): no need to read its length back — the
                    // fix-up kernel takes it from the device and returns at once when it is empty
                    StageTimer tm(h, DPF_T_FIXUP);
                    k_fixup_exact<<<32, 128, 0, h->stream>>>(Xc, h->A.p, d, PW, h->fix_list.p, h->counters.p,
                                                            (int)h->fix_list.cap, h->signs.p, fix_total); DPF_LAUNCHED();
                    DPF_CUDA(cudaGetLastError());
                    break;
                }
                int32_t cnt = 0;
                DPF_CUDA(cudaMemcpyAsync(&cnt, h->counters.p, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
                DPF_CUDA(cudaStreamSynchronize(h->stream));
                if ((size_t)cnt > h->fix_list.cap) {   // list overflowed: grow and redo this chunk
                    h->fix_list.reserve((size_t)cnt + 1024);
                    continue;
                }
                if (cnt > 0) {
                    StageTimer tm(h, DPF_T_FIXUP);
                    k_fixup_exact<<<std::min(1024, (cnt + 127) / 128), 128, 0, h->stream>>>(
                        Xc, h->A.p, d, PW, h->fix_list.p, h->counters.p, (int)h->fix_list.cap, h->signs.p, fix_total); DPF_LAUNCHED();
                    DPF_CUDA(cudaGetLastError());
                }
                break;
            }
        }
        pack_launch(h, m, keys_out + r0, pids_out ? pids_out + r0 : nullptr, ld);
    }
}

// exact CUDA-core path for the angle family (cross-check hook; selected with DPF_HASH_EXACT=1)
void hash_dense_device_exact(dpf_index* h, const double* Xd, int64_t n, int32_t* keys_out, uint8_t* pids_out,
                             int64_t ld) {
    const int d = h->cfg.d, P = h->P, PW = h->PW;
    const int64_t chunk = hash_chunk_rows(h);
    h->signs.reserve((size_t)std::min(n, chunk) * PW);
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t m = std::min(chunk, n - r0);
        {
            StageTimer tm(h, DPF_T_HASH);
            const int nct = (P + EX_T - 1) / EX_T;
            const unsigned grid = (unsigned)(((m + EX_T - 1) / EX_T) * nct);
            k_project_exact<false><<<grid, 256, 0, h->stream>>>(Xd + r0 * d, h->A.p, m, d, P, PW, nct, nullptr, nullptr,
                                                                h->signs.p, nullptr); DPF_LAUNCHED();
            DPF_CUDA(cudaGetLastError());
        }
        pack_launch(h, m, keys_out + r0, pids_out ? pids_out + r0 : nullptr, ld);
    }
}

template <int PPL>
static void csr_launch(dpf_index* h, const int64_t* ptr, const int32_t* idx, const double* val, int64_t m, bool pst) {
    const int wpb = 8;
    const unsigned grid = (unsigned)((m + wpb - 1) / wpb);
    if (pst)
        k_project_csr<PPL, true><<<grid, wpb * 32, 0, h->stream>>>(ptr, idx, val, h->At.p, h->At_ld, m, h->P, h->PW,
                                                                    h->fb.p, h->fw.p, nullptr, h->pq.p);
    else
        k_project_csr<PPL, false><<<grid, wpb * 32, 0, h->stream>>>(ptr, idx, val, h->At.p, h->At_ld, m, h->P, h->PW,
                                                                     nullptr, nullptr, h->signs.p, nullptr);
    DPF_LAUNCHED();
}

void hash_csr_device(dpf_index* h, const int64_t* ptr, const int32_t* idx, const double* val, int64_t n,
                     int32_t* keys_out, uint8_t* pids_out, int64_t ld) {
    if (n <= 0) return;
    const int D = h->cfg.d, P = h->P, PW = h->PW;
    const bool pst = h->cfg.family_kind == DPF_FAMILY_PSTABLE;
    if (h->At.cap == 0) {   // feature-major copy of the functions, padded to a multiple of 32 columns
        h->At_ld = (P + 31) / 32 * 32;
        std::vector<double> at((size_t)D * h->At_ld, 0.0);
        for (int p = 0; p < P; ++p)
            for (int j = 0; j < D; ++j) at[(size_t)j * h->At_ld + p] = h->hA[(size_t)p * D + j];
        h->At.reserve(at.size());
        DPF_CUDA(cudaMemcpyAsync(h->At.p, at.data(), at.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        DPF_CUDA(cudaStreamSynchronize(h->stream));
    }
    DPF_REQUIRE(P <= 32 * 16, DPF_ERR_INVALID, "CSR hashing supports at most 512 distinct functions");
    const int64_t chunk = hash_chunk_rows(h);
    h->signs.reserve((size_t)std::min(n, chunk) * PW);
    if (pst) h->pq.reserve((size_t)std::min(n, chunk) * P);
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
        const int64_t m = std::min(chunk, n - r0);
        {
            StageTimer tm(h, DPF_T_HASH);
            const int ppl = (P + 31) / 32;
            if (ppl <= 1) csr_launch<1>(h, ptr + r0, idx, val, m, pst);
            else if (ppl <= 2) csr_launch<2>(h, ptr + r0, idx, val, m, pst);
            else if (ppl <= 4) csr_launch<4>(h, ptr + r0, idx, val, m, pst);
            else if (ppl <= 8) csr_launch<8>(h, ptr + r0, idx, val, m, pst);
            else if (ppl <= 10) csr_launch<10>(h, ptr + r0, idx, val, m, pst);
            else csr_launch<16>(h, ptr + r0, idx, val, m, pst);
            DPF_CUDA(cudaGetLastError());
        }
        pack_launch(h, m, keys_out + r0, pids_out ? pids_out + r0 : nullptr, ld);
    }
}

}  // namespace dpf
