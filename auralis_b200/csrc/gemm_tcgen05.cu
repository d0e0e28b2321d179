// bf16 tensor-core GEMM for sm_100a:  out[M,N] = epi(A[M,K] . W[N,K]^T + bias)
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 4-stage smem ring
//   * tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) issued by one thread, fp32 accumulators in TMEM
//   * epilogue warps read TMEM with tcgen05.ld.32x32b, fuse bias / gelu_new / residual, store fp32 or bf16
// Warp roles: warp0 = TMA producer, warp1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue (one TMEM
// lane quarter each).  Replaces cuBLAS behind vLLM's GPT2 linears (SURVEY.md §2.4 K4,K6,K7,K9).
//
// Every mbarrier wait carries a clock watchdog that traps instead of hanging the GPU.
#include <cuda.h>
#include "kernels.h"

namespace xtts {
namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
// ring depth per tile width
// (BN <= 64 rings stay <= 100 KB so that, under PDL, the next GEMM's CTA can sit on the same SM and prefetch its weight tiles)
__host__ __device__ constexpr int stages_for(int bn) { return bn >= 128 ? 6 : (bn >= 64 ? 4 : 5); }
constexpr int kThreads = 192;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {      // ~2 s at 2 GHz: a protocol bug, not a slow tile
            printf("gemm_tcgen05: mbarrier watchdog (tag %d, block %d,%d, thread %d)\n", tag, blockIdx.x, blockIdx.y,
                   threadIdx.x);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SW128)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows x 128 B between core-matrix groups
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ bias, const float* resid, void* out, int M, int N, int K, int flags,
                    int a_box_rows) {
    // (bias / out / flags are re-pointed below for split-K launches)
    constexpr int STAGES = stages_for(BN);
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    // split-K: gridDim.z CTAs share one output tile, each owning a contiguous range of k-blocks and writing its raw
    // fp32 partial tile to out + z*M*N (bias/activation/residual are applied by the consumer, deterministically)
    const int total_kb = K / BK;
    const int kb_per = (total_kb + gridDim.z - 1) / gridDim.z;
    const int kb_begin = blockIdx.z * kb_per;
    const int num_kb = max(0, min(total_kb, kb_begin + kb_per) - kb_begin);
    if (gridDim.z > 1) {
        out = reinterpret_cast<float*>(out) + (size_t)blockIdx.z * M * N;
        bias = nullptr; flags = 0;
    }

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_smem)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    pdl_trigger();                                   // dependents may start their prologue now
    if (warp == 0) {
        if (lane == 0) {
            // weights never depend on the previous kernel: their tiles for the first ring pass are requested BEFORE the
            // dependency wait, the activation tiles right after it
            const int pre = min(num_kb, STAGES);
            for (int kb = 0; kb < pre; ++kb) {
                mbar_expect_tx(&full_bar[kb], (uint32_t)a_box_rows * BK * 2 + B_BYTES);
                tma_load_2d(sB + kb * B_BYTES, &tmB, &full_bar[kb], (kb_begin + kb) * BK, n0);
            }
            pdl_wait();
            for (int kb = 0; kb < pre; ++kb)
                tma_load_2d(sA + kb * A_BYTES, &tmA, &full_bar[kb], (kb_begin + kb) * BK, m0);
            for (int kb = pre; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1, 1);
                mbar_expect_tx(&full_bar[s], (uint32_t)a_box_rows * BK * 2 + B_BYTES);
                tma_load_2d(sA + s * A_BYTES, &tmA, &full_bar[s], (kb_begin + kb) * BK, m0);
                tma_load_2d(sB + s * B_BYTES, &tmB, &full_bar[s], (kb_begin + kb) * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[s], ph, 2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(sA + s * A_BYTES), b_addr = smem_u32(sB + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t ad = make_sw128_desc(a_addr + k * UMMA_K * 2);
                    const uint64_t bd = make_sw128_desc(b_addr + k * UMMA_K * 2);
                    umma_bf16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // frees the smem slot once the MMAs have read it
            }
            umma_commit(&tmem_full_bar);             // accumulator complete
        }
    } else {
        // ---------------- epilogue: warp (2..5) owns TMEM lanes [32*(warp%4), +32) = tile rows
        const int q = warp & 3;
        pdl_wait();                                  // this warp reads `resid` and overwrites `out`
        mbar_wait(&tmem_full_bar, 0, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int nb = n0 + c * 32;
            if (row < M && nb < N) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x = __uint_as_float(r[i]);
                    if (bias) x += __ldg(bias + nb + i);
                    if (flags & GEMM_GELU) x = gelu_new(x);
                    v[i] = x;
                }
                if (flags & GEMM_RESID) {
                    const float4* rp = reinterpret_cast<const float4*>(resid + (size_t)row * N + nb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = rp[i];
                        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                    }
                }
                if (flags & GEMM_OUT_BF16) {
                    __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * N + nb;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 pk;
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]);
                        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
                        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        reinterpret_cast<uint4*>(op)[i] = pk;
                    }
                } else {
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * N + nb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// rows of the A box: the whole 128-row tile, or (single m-tile problems) just the valid rows rounded up to 8 —
// the rest of the smem tile keeps stale data whose accumulator rows are never stored
inline int a_box_rows_for(int M) { return M >= BM ? BM : ((M + 7) / 8) * 8; }

void encode_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, box_rows);
        throw CudaError(buf);
    }
}

template <int BN>
void launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, const float* resid, void* out, int M,
               int N, int K, int flags, cudaStream_t st, int splits, int a_box_rows, bool pdl) {
    constexpr int STAGES = stages_for(BN);
    constexpr size_t smem = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    ProfScope ps(KF_GEMM_TC, st, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + ((flags & GEMM_OUT_BF16) ? 2.0 : 4.0) * M * N);
    launch_k(gemm_bf16_tc_kernel<BN>, grid, dim3(kThreads), smem, st, pdl, tmA, tmB, bias, resid, out, M, N, K, flags, a_box_rows);
    COUNT_LAUNCH(); KERNEL_CHECK();
}

}  // namespace

bool gemm_tc_init(std::string* err) {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        if (err) *err = std::string("cuTensorMapEncodeTiled not available: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
        return false;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    // opt every instantiation into its dynamic shared memory now (never inside a stream capture)
    auto smem_of = [](int bn) { return (int)(stages_for(bn) * (BM * BK * 2 + bn * BK * 2) + 1024); };
    cudaFuncSetAttribute(gemm_bf16_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(128));
    cudaFuncSetAttribute(gemm_bf16_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(64));
    cudaFuncSetAttribute(gemm_bf16_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(32));
    (void)cudaGetLastError();
    return true;
}

void launch_gemm_bf16_tc(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, const float* resid,
                         void* out, int M, int N, int K, int flags, cudaStream_t st, bool pdl) {
    if (M <= 0 || N <= 0) return;
    if (K % BK != 0 || N % 32 != 0) throw CudaError("gemm_bf16_tc: need K % 64 == 0 and N % 32 == 0");
    if (!g_encode) {
        std::string err;
        if (!gemm_tc_init(&err)) throw CudaError(err);
    }
    // tile-N heuristic: the largest tile that still gives >= 1 CTA per SM; decode-shaped problems
    // (small M) are weight-streaming, so more, narrower CTAs pull more HBM bandwidth.
    const int mt = ceil_div(M, BM);
    int bn = 128;
    if (N % 128 != 0 || mt * (N / 128) < 148) bn = 64;
    if (bn == 64 && (N % 64 != 0 || mt * (N / 64) < 148)) bn = 32;
    CUtensorMap tmA, tmB;
    const int abox = a_box_rows_for(M);
    encode_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint32_t)abox);
    encode_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint32_t)bn);
    if (bn == 128) launch_bn<128>(tmA, tmB, bias, resid, out, M, N, K, flags, st, 1, abox, pdl);
    else if (bn == 64) launch_bn<64>(tmA, tmB, bias, resid, out, M, N, K, flags, st, 1, abox, pdl);
    else launch_bn<32>(tmA, tmB, bias, resid, out, M, N, K, flags, st, 1, abox, pdl);
}


// split-K variant for the skinny decode GEMMs (N = hidden): partials[z][M][N] = A[:, kz] . W[:, kz]^T, fp32, no epilogue
void launch_gemm_bf16_tc_splitk(const __nv_bfloat16* A, const __nv_bfloat16* W, float* partials, int M, int N, int K,
                                int splits, cudaStream_t st, bool pdl) {
    if (M <= 0 || N <= 0) return;
    if (K % BK != 0 || N % 32 != 0 || splits < 1 || (K / BK) % splits != 0) throw CudaError("gemm_bf16_tc_splitk: bad shape");
    if (!g_encode) { std::string err; if (!gemm_tc_init(&err)) throw CudaError(err); }
    const int mt = ceil_div(M, BM);
    int bn = 128;
    if (N % 128 != 0 || mt * (N / 128) * splits < 148) bn = 64;
    if (bn == 64 && (N % 64 != 0 || mt * (N / 64) * splits < 148)) bn = 32;
    CUtensorMap tmA, tmB;
    const int abox = a_box_rows_for(M);
    encode_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint32_t)abox);
    encode_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint32_t)bn);
    if (bn == 128) launch_bn<128>(tmA, tmB, nullptr, nullptr, partials, M, N, K, 0, st, splits, abox, pdl);
    else if (bn == 64) launch_bn<64>(tmA, tmB, nullptr, nullptr, partials, M, N, K, 0, st, splits, abox, pdl);
    else launch_bn<32>(tmA, tmB, nullptr, nullptr, partials, M, N, K, 0, st, splits, abox, pdl);
}

}  // namespace xtts
