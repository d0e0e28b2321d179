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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
    int polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 1023) == 0 && clock64() - t0 > 4000000000LL) {      // ~2 s at 2 GHz: a protocol bug, not a slow tile
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

// STAGES = 0: the default ring (stages_for(BN), two CTAs of the narrow tiles fit on an SM); otherwise an explicit depth — the
// "deep" instantiations (engine option "gemm_deep_ring") fill the SM's shared memory with ONE CTA's ring: a decode-shaped GEMM
// is a chain of ring passes that each pay the load latency, and its 16 k-blocks then need 1.6 passes instead of 3.2
template <int BN, int STAGES_ = 0>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ bias, const float* resid, void* out, int M, int N, int K, int flags,
                    int a_box_rows, const DepFlag dep) {
    // (bias / out / flags are re-pointed below for split-K launches)
    constexpr int STAGES = STAGES_ > 0 ? STAGES_ : stages_for(BN);
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

    trace_pt(TR_GEMM, 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    // split-K: gridDim.z CTAs share one output tile, each owning a contiguous range of k-blocks and writing its raw
    // fp32 partial tile to out + z*M*N (bias/activation/residual are applied by the consumer, deterministically)
    const int total_kb = K / BK;
    const int kb_per = (total_kb + gridDim.z - 1) / gridDim.z;
    const int kb_begin = blockIdx.z * kb_per;
    const int num_kb = max(0, min(total_kb, kb_begin + kb_per) - kb_begin);
    const int flags_in = flags;
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
            // ... and the weight tiles of the later passes are pulled into L2 meanwhile: a decode-shaped GEMM is a chain of
            // ring passes each paying the full HBM latency (~1.6 us under load); from L2 a pass costs ~0.7 us
            if (!(flags_in & GEMM_NO_L2PF))
            for (int kb = pre; kb < num_kb; ++kb)
                asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                             ::"l"(&tmB), "r"((kb_begin + kb) * BK), "r"(n0) : "memory");
            dep_wait(dep, 21);                       // the activations exist (whole predecessor grid, or its counter)
            if (dep.wait) asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores -> TMA (async proxy) reads
            trace_pt(TR_GEMM, 1);
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
            // operand format bits of the instruction descriptor: bf16 (1) by default, IEEE fp16 (0) with GEMM_F16
            const uint32_t idesc = (flags_in & GEMM_F16) ? (make_idesc(BN) & ~((1u << 7) | (1u << 10))) : make_idesc(BN);
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
        if (dep.wait == nullptr) pdl_wait();         // this warp reads `resid` and overwrites `out` (with a counter: ordered behind
                                                     // the producer's wait through tmem_full_bar; such launches carry no resid)
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
                    const bool f16 = (flags & GEMM_F16) != 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 pk;
                        pk.x = pack16(v[8 * i], v[8 * i + 1], f16); pk.y = pack16(v[8 * i + 2], v[8 * i + 3], f16);
                        pk.z = pack16(v[8 * i + 4], v[8 * i + 5], f16); pk.w = pack16(v[8 * i + 6], v[8 * i + 7], f16);
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
    dep_arrive(dep, threadIdx.x == 0);               // (fence + CTA barrier inside) this CTA's tile is in global memory
    __syncthreads();
    trace_pt(TR_GEMM, 2);
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

constexpr int deep_stages_for(int bn) { return bn >= 128 ? 6 : (bn >= 64 ? 8 : 10); }

template <int BN, int STAGES_>
int launch_bn_st(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, const float* resid, void* out, int M,
                 int N, int K, int flags, cudaStream_t st, int splits, int a_box_rows, bool pdl, const DepFlag& dep) {
    constexpr int STAGES = STAGES_ > 0 ? STAGES_ : stages_for(BN);
    constexpr size_t smem = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024;
    static bool attr_set[64] = {};
    if (first_on_device(attr_set))
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, STAGES_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
    ProfScope ps(KF_GEMM_TC, st, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + ((flags & GEMM_OUT_BF16) ? 2.0 : 4.0) * M * N);
    if (dep.wait && (flags & GEMM_RESID)) throw CudaError("gemm_bf16_tc: a counter dependency cannot order a residual read");
    launch_k(gemm_bf16_tc_kernel<BN, STAGES_>, grid, dim3(kThreads), smem, st, pdl, tmA, tmB, bias, resid, out, M, N, K,
             flags | (g_gemm_l2_prefetch ? 0 : GEMM_NO_L2PF), a_box_rows, dep);
    COUNT_LAUNCH(); KERNEL_CHECK();
    return (int)(grid.x * grid.y * grid.z);
}
template <int BN>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, const float* resid, void* out, int M,
              int N, int K, int flags, cudaStream_t st, int splits, int a_box_rows, bool pdl, const DepFlag& dep) {
    // deep ring: only where the whole K range is one CTA's (no split-K: those CTAs have 2-4 k-blocks) and the shape is decode's
    if (g_gemm_deep_ring && BN < 128 && M <= 256 && splits == 1 && K / BK > stages_for(BN))
        return launch_bn_st<BN, deep_stages_for(BN)>(tmA, tmB, bias, resid, out, M, N, K, flags, st, splits, a_box_rows, pdl, dep);
    return launch_bn_st<BN, 0>(tmA, tmB, bias, resid, out, M, N, K, flags, st, splits, a_box_rows, pdl, dep);
}


// =================================================================================================================
// Fused decode "chain" kernel: everything between two attention calls of the decode step in ONE persistent launch
//
//   proj (split-K) -> residual + LayerNorm2 -> fc + gelu -> fc2 (split-K) -> residual + LayerNorm1(next) -> qkv(next)
//
// At decode shapes (M <= 256 rows) every one of those six launches is latency-bound (a 16-k-block main loop between a
// launch, a TMEM/barrier prologue and an epilogue); here the CTAs stay resident, keep their barriers, smem ring and TMEM
// accumulator, and separate the phases with a device-wide barrier (one atomic + one acquire-poll per CTA).
// One CTA per SM, 192 threads: GEMM phases use the roles of gemm_bf16_tc_kernel (warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-5 epilogue); the reduce + LayerNorm phases use all 6 warps, one row per warp.
// Data written by generic stores in one phase is read by TMA (async proxy) in the next: the barrier is followed by a
// cross-proxy fence.  Co-residency: the grid is <= the SM count, one CTA per SM, and griddepcontrol.launch_dependents is
// only issued after the last device-wide barrier, so no dependent grid can take an SM one of our CTAs still needs.
// =================================================================================================================
constexpr int CH_BN = 64, CH_STAGES = 6, CH_SPLITS = 4;
enum { EPI_PARTIAL = 0, EPI_BIAS_F32 = 1, EPI_BIAS_GELU_BF16 = 2 };

struct ChainParams {
    int phases, M, H, FF, abox;
    float eps;
    const float* proj_bias; const float* ln2_w; const float* ln2_b;
    const float* fc_bias;
    const float* fc2_bias; const float* ln1_w; const float* ln1_b;     // LayerNorm1 of the NEXT layer (null: residual only)
    const float* qkv_bias;                                              // next layer
    float* X; __nv_bfloat16* Xn; __nv_bfloat16* FFb; float* QKV; float* PART;
    unsigned* sync;                                                     // [0] arrival count, [32] generation (64 words)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// self-resetting device-wide barrier (all CTAs of the grid are resident: grid <= #SMs, 1 CTA/SM).
// sync[0] = arrival count, sync[32] = generation: separate 128-byte lines, so the pollers do not queue behind the arrivals.
__device__ __forceinline__ void chain_grid_sync(unsigned* sync) {
    asm volatile("fence.proxy.async;" ::: "memory");          // generic writes of this phase -> later async-proxy (TMA) reads
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* gen_p = sync + 32;
        const unsigned gen = ld_acquire_u32(gen_p);
        __threadfence();
        if (atomicAdd(sync, 1u) == gridDim.x - 1) {
            sync[0] = 0u;
            __threadfence();
            atomicAdd(gen_p, 1u);
        } else {
            const long long t0 = clock64();
            int polls = 0;
            while (ld_acquire_u32(gen_p) == gen) {
                __nanosleep(40);
                if ((++polls & 1023) == 0 && clock64() - t0 > 4000000000LL) {
                    printf("decode_chain: grid barrier watchdog (block %d)\n", blockIdx.x);
                    __trap();
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
    asm volatile("fence.proxy.async;" ::: "memory");
}

struct ChainSmem {
    uint8_t* sA; uint8_t* sB;
    uint64_t* full_bar; uint64_t* empty_bar; uint64_t* tmem_full; uint64_t* tmem_empty;
    uint32_t tmem_base;
};

// one GEMM phase: out = epi(A[M,K] . W[N,K]^T); tiles (n-tile, m-tile, k-split) strided over the grid
__device__ __forceinline__ void chain_gemm(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, int splits,
                                           int epi, const float* __restrict__ bias, void* out, int abox, const ChainSmem& S,
                                           int warp, int lane, int& it, int& lt) {
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = CH_BN * BK * 2;
    const int tiles_n = N / CH_BN, mt = (M + BM - 1) / BM;
    const int kb_per = (K / BK) / splits;
    const int total = tiles_n * mt * splits;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int nt = tile % tiles_n, mi = (tile / tiles_n) % mt, z = tile / (tiles_n * mt);
        const int n0 = nt * CH_BN, m0 = mi * BM, kb0 = z * kb_per;
        if (warp == 0) {
            if (lane == 0) {
                for (int kb = 0; kb < kb_per; ++kb, ++it) {
                    const int s = it % CH_STAGES;
                    mbar_wait(&S.empty_bar[s], (uint32_t)(((it / CH_STAGES) & 1) ^ 1), 11);
                    mbar_expect_tx(&S.full_bar[s], (uint32_t)abox * BK * 2 + B_BYTES);
                    tma_load_2d(S.sB + s * B_BYTES, tmB, &S.full_bar[s], (kb0 + kb) * BK, n0);
                    tma_load_2d(S.sA + s * A_BYTES, tmA, &S.full_bar[s], (kb0 + kb) * BK, m0);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc(CH_BN);
                mbar_wait(S.tmem_empty, (uint32_t)((lt & 1) ^ 1), 16);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < kb_per; ++kb, ++it) {
                    const int s = it % CH_STAGES;
                    mbar_wait(&S.full_bar[s], (uint32_t)((it / CH_STAGES) & 1), 12);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(S.sA + s * A_BYTES), b_addr = smem_u32(S.sB + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        umma_bf16(S.tmem_base, make_sw128_desc(a_addr + k * UMMA_K * 2), make_sw128_desc(b_addr + k * UMMA_K * 2),
                                  idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&S.empty_bar[s]);
                }
                umma_commit(S.tmem_full);
            }
            ++lt;
        } else {
            const int q = warp & 3;
            mbar_wait(S.tmem_full, (uint32_t)(lt & 1), 13);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < CH_BN / 32; ++c) {
                uint32_t r[32];
                const uint32_t taddr = S.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
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
                if (row < M) {
                    if (epi == EPI_BIAS_GELU_BF16) {
                        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * N + nb;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = gelu_new(__uint_as_float(r[8 * i + e]) + __ldg(bias + nb + 8 * i + e));
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            reinterpret_cast<uint4*>(op)[i] = pk;
                        }
                    } else {
                        float* base = reinterpret_cast<float*>(out) + (epi == EPI_PARTIAL ? (size_t)z * M * N : (size_t)0);
                        float4* op = reinterpret_cast<float4*>(base + (size_t)row * N + nb);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                                   __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                            if (epi == EPI_BIAS_F32) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb) + i);
                                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                            }
                            op[i] = v;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(S.tmem_empty);
            ++lt;
        }
    }
}

// X[r] += bias + sum_z PART[z][r] (fixed order: deterministic); optional LayerNorm -> bf16 Xn[r].  One row per warp.
__device__ __forceinline__ void chain_reduce_ln(const ChainParams& P, const float* __restrict__ bias, const float* __restrict__ w,
                                                const float* __restrict__ b, int warp, int lane) {
    const int H = P.H, nj = H / 128;                 // H <= 1024: up to 8 float4 per lane
    const size_t zs = (size_t)P.M * H;
    for (int r = blockIdx.x * 6 + warp; r < P.M; r += gridDim.x * 6) {
        float4 v[8];
        float* x = P.X + (size_t)r * H;
        const float* p = P.PART + (size_t)r * H;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nj) {
                const int c = j * 128 + lane * 4;
                float4 acc = __ldcg(reinterpret_cast<const float4*>(x + c));
                float4 pz[CH_SPLITS];
#pragma unroll
                for (int z = 0; z < CH_SPLITS; ++z) pz[z] = __ldcg(reinterpret_cast<const float4*>(p + (size_t)z * zs + c));
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
                acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
#pragma unroll
                for (int z = 0; z < CH_SPLITS; ++z) { acc.x += pz[z].x; acc.y += pz[z].y; acc.z += pz[z].z; acc.w += pz[z].w; }
                v[j] = acc;
                *reinterpret_cast<float4*>(x + c) = acc;
            }
        }
        if (w == nullptr) continue;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nj) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        const float mean = warp_sum(s) / (float)H;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nj) {
                const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
                q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
            }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + P.eps);
        __nv_bfloat16* y = P.Xn + (size_t)r * H;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nj) {
                const int c = j * 128 + lane * 4;
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c)), b4 = __ldg(reinterpret_cast<const float4*>(b + c));
                __nv_bfloat162 h0 = __floats2bfloat162_rn((v[j].x - mean) * rstd * w4.x + b4.x, (v[j].y - mean) * rstd * w4.y + b4.y);
                __nv_bfloat162 h1 = __floats2bfloat162_rn((v[j].z - mean) * rstd * w4.z + b4.z, (v[j].w - mean) * rstd * w4.w + b4.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(y + c) = pk;
            }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
decode_chain_kernel(const __grid_constant__ CUtensorMap tmATT, const __grid_constant__ CUtensorMap tmWo,
                    const __grid_constant__ CUtensorMap tmXn, const __grid_constant__ CUtensorMap tmWfc,
                    const __grid_constant__ CUtensorMap tmFF, const __grid_constant__ CUtensorMap tmWproj,
                    const __grid_constant__ CUtensorMap tmWqkv, const ChainParams P) {
    constexpr uint32_t A_BYTES = BM * BK * 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[CH_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[CH_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar, tmem_empty_bar;
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmATT) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXn) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWfc) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmFF) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWproj) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWqkv) : "memory");
        for (int s = 0; s < CH_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        mbar_init(&tmem_empty_bar, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_smem)), "r"((uint32_t)CH_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    ChainSmem S;
    S.sA = smem; S.sB = smem + CH_STAGES * A_BYTES;
    S.full_bar = full_bar; S.empty_bar = empty_bar; S.tmem_full = &tmem_full_bar; S.tmem_empty = &tmem_empty_bar;
    S.tmem_base = tmem_base_smem;

    pdl_wait();                                    // the attention output (and everything before it) is complete
    int it = 0, lt = 0;
    const int last = 31 - __clz(P.phases);         // index of the last enabled phase
    auto boundary = [&](int bit) {                 // device-wide barrier after every phase but the last; dependents may
        if ((1 << last) == bit) return;            // start launching once the final barrier is behind us
        chain_grid_sync(P.sync);
        if ((P.phases & ~(2 * bit - 1)) == (1 << last)) pdl_trigger();
    };
    if (P.phases == (1 << last)) pdl_trigger();    // single-phase launch: nothing to wait for
    if (P.phases & DC_PROJ) {
        chain_gemm(&tmATT, &tmWo, P.M, P.H, P.H, CH_SPLITS, EPI_PARTIAL, nullptr, P.PART, P.abox, S, warp, lane, it, lt);
        boundary(DC_PROJ);
    }
    if (P.phases & DC_LN2) { chain_reduce_ln(P, P.proj_bias, P.ln2_w, P.ln2_b, warp, lane); boundary(DC_LN2); }
    if (P.phases & DC_FC) {
        chain_gemm(&tmXn, &tmWfc, P.M, P.FF, P.H, 1, EPI_BIAS_GELU_BF16, P.fc_bias, P.FFb, P.abox, S, warp, lane, it, lt);
        boundary(DC_FC);
    }
    if (P.phases & DC_FC2) {
        chain_gemm(&tmFF, &tmWproj, P.M, P.H, P.FF, CH_SPLITS, EPI_PARTIAL, nullptr, P.PART, P.abox, S, warp, lane, it, lt);
        boundary(DC_FC2);
    }
    if (P.phases & DC_LN1) { chain_reduce_ln(P, P.fc2_bias, P.ln1_w, P.ln1_b, warp, lane); boundary(DC_LN1); }
    if (P.phases & DC_QKV)
        chain_gemm(&tmXn, &tmWqkv, P.M, 3 * P.H, P.H, 1, EPI_BIAS_F32, P.qkv_bias, P.QKV, P.abox, S, warp, lane, it, lt);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(S.tmem_base), "r"((uint32_t)CH_BN) : "memory");
    }
}

}  // namespace

XTTS_TRACE_SETTER(trace_set_gemm)

bool gemm_tc_init(std::string* err) {
    static bool dev_done[64] = {};
    auto per_device = [&] {
        // opt every instantiation into its dynamic shared memory now (never inside a stream capture); per device
        if (!first_on_device(dev_done)) return;
        auto smem_of = [](int bn) { return (int)(stages_for(bn) * (BM * BK * 2 + bn * BK * 2) + 1024); };
        cudaFuncSetAttribute(gemm_bf16_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(128));
        cudaFuncSetAttribute(gemm_bf16_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(64));
        cudaFuncSetAttribute(gemm_bf16_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(32));
        auto deep_of = [](int bn) { return (int)(deep_stages_for(bn) * (BM * BK * 2 + bn * BK * 2) + 1024); };
        cudaFuncSetAttribute(gemm_bf16_tc_kernel<64, deep_stages_for(64)>, cudaFuncAttributeMaxDynamicSharedMemorySize, deep_of(64));
        cudaFuncSetAttribute(gemm_bf16_tc_kernel<32, deep_stages_for(32)>, cudaFuncAttributeMaxDynamicSharedMemorySize, deep_of(32));
        (void)cudaGetLastError();
    };
    if (g_encode) { per_device(); return true; }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        if (err) *err = std::string("cuTensorMapEncodeTiled not available: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
        return false;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    per_device();
    return true;
}

int launch_gemm_bf16_tc(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, const float* resid,
                        void* out, int M, int N, int K, int flags, cudaStream_t st, bool pdl, DepFlag dep) {
    if (M <= 0 || N <= 0) return 0;
    if (K % BK != 0 || N % 32 != 0) throw CudaError("gemm_bf16_tc: need K % 64 == 0 and N % 32 == 0");
    if (g_gemm_2cta && !pdl && !dep.wait && !dep.arrive && gemm_2cta_supported(M, N, K)) {       // prefill-shaped: persistent CTA pairs
        launch_gemm_bf16_2cta(A, W, bias, resid, out, M, N, K, flags, st);
        return 0;
    }
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
    if (g_gemm_decode_bn && M <= 256 && N % g_gemm_decode_bn == 0) bn = g_gemm_decode_bn;
    CUtensorMap tmA, tmB;
    const int abox = a_box_rows_for(M);
    encode_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint32_t)abox);
    encode_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint32_t)bn);
    if (bn == 128) return launch_bn<128>(tmA, tmB, bias, resid, out, M, N, K, flags, st, 1, abox, pdl, dep);
    if (bn == 64) return launch_bn<64>(tmA, tmB, bias, resid, out, M, N, K, flags, st, 1, abox, pdl, dep);
    return launch_bn<32>(tmA, tmB, bias, resid, out, M, N, K, flags, st, 1, abox, pdl, dep);
}


// split-K variant for the skinny decode GEMMs (N = hidden): partials[z][M][N] = A[:, kz] . W[:, kz]^T, fp32, no epilogue
int launch_gemm_bf16_tc_splitk(const __nv_bfloat16* A, const __nv_bfloat16* W, float* partials, int M, int N, int K,
                               int splits, cudaStream_t st, bool pdl, DepFlag dep, int flags) {
    flags &= GEMM_F16;
    if (M <= 0 || N <= 0) return 0;
    if (K % BK != 0 || N % 32 != 0 || splits < 1 || (K / BK) % splits != 0) throw CudaError("gemm_bf16_tc_splitk: bad shape");
    if (!g_encode) { std::string err; if (!gemm_tc_init(&err)) throw CudaError(err); }
    const int mt = ceil_div(M, BM);
    int bn = 128;
    if (N % 128 != 0 || mt * (N / 128) * splits < 148) bn = 64;
    if (bn == 64 && (N % 64 != 0 || mt * (N / 64) * splits < 148)) bn = 32;
    if (g_gemm_decode_bn && M <= 256 && N % g_gemm_decode_bn == 0) bn = g_gemm_decode_bn;
    CUtensorMap tmA, tmB;
    const int abox = a_box_rows_for(M);
    encode_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint32_t)abox);
    encode_2d(&tmB, W, (uint64_t)N, (uint64_t)K, (uint32_t)bn);
    if (bn == 128) return launch_bn<128>(tmA, tmB, nullptr, nullptr, partials, M, N, K, flags, st, splits, abox, pdl, dep);
    if (bn == 64) return launch_bn<64>(tmA, tmB, nullptr, nullptr, partials, M, N, K, flags, st, splits, abox, pdl, dep);
    return launch_bn<32>(tmA, tmB, nullptr, nullptr, partials, M, N, K, flags, st, splits, abox, pdl, dep);
}


bool decode_chain_supported(int M, int H, int FF) {
    return M >= 1 && H % 128 == 0 && H <= 1024 && (H / BK) % CH_SPLITS == 0 && FF % CH_BN == 0 && (FF / BK) % CH_SPLITS == 0;
}

// one launch = the phases of `a.phases` (DC_* bits) for one layer boundary of the decode step
void launch_decode_chain(const DecodeChainArgs& a, cudaStream_t st, bool pdl) {
    if (a.M <= 0 || a.phases == 0) return;
    if (!decode_chain_supported(a.M, a.H, a.FF)) throw CudaError("decode_chain: unsupported geometry");
    if (!g_encode) { std::string err; if (!gemm_tc_init(&err)) throw CudaError(err); }
    static int n_sm = 0;
    static bool attr_set[64] = {};
    constexpr size_t smem = CH_STAGES * (BM * BK * 2 + CH_BN * BK * 2) + 1024;
    if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    if (first_on_device(attr_set))
        CUDA_CHECK(cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int abox = a_box_rows_for(a.M);
    const int H = a.H, FF = a.FF;
    CUtensorMap tATT, tWo, tXn, tWfc, tFF, tWproj, tWqkv;
    encode_2d(&tATT, a.ATT, (uint64_t)a.M, (uint64_t)H, (uint32_t)abox);
    encode_2d(&tXn, a.Xn, (uint64_t)a.M, (uint64_t)H, (uint32_t)abox);
    encode_2d(&tFF, a.FFb, (uint64_t)a.M, (uint64_t)FF, (uint32_t)abox);
    encode_2d(&tWo, a.Wo, (uint64_t)H, (uint64_t)H, CH_BN);
    encode_2d(&tWfc, a.Wfc, (uint64_t)FF, (uint64_t)H, CH_BN);
    encode_2d(&tWproj, a.Wproj, (uint64_t)H, (uint64_t)FF, CH_BN);
    encode_2d(&tWqkv, a.Wqkv ? a.Wqkv : a.Wo, (uint64_t)(a.Wqkv ? 3 * H : H), (uint64_t)H, CH_BN);
    ChainParams P{};
    P.phases = a.phases; P.M = a.M; P.H = H; P.FF = FF; P.abox = abox; P.eps = a.eps;
    P.proj_bias = a.proj_bias; P.ln2_w = a.ln2_w; P.ln2_b = a.ln2_b; P.fc_bias = a.fc_bias; P.fc2_bias = a.fc2_bias;
    P.ln1_w = a.ln1_w; P.ln1_b = a.ln1_b; P.qkv_bias = a.qkv_bias;
    P.X = a.X; P.Xn = a.Xn; P.FFb = a.FFb; P.QKV = a.QKV; P.PART = a.PART; P.sync = a.sync;
    double fl = 0, by = 0;
    const double M = a.M;
    if (a.phases & DC_PROJ) { fl += 2.0 * M * H * H; by += 2.0 * (M * H + (double)H * H) + 4.0 * CH_SPLITS * M * H; }
    if (a.phases & DC_LN2) by += (8.0 + 4.0 * CH_SPLITS + 2.0) * M * H;
    if (a.phases & DC_FC) { fl += 2.0 * M * H * FF; by += 2.0 * (M * H + (double)H * FF) + 2.0 * M * FF; }
    if (a.phases & DC_FC2) { fl += 2.0 * M * H * FF; by += 2.0 * (M * FF + (double)H * FF) + 4.0 * CH_SPLITS * M * H; }
    if (a.phases & DC_LN1) by += (8.0 + 4.0 * CH_SPLITS + 2.0) * M * H;
    if (a.phases & DC_QKV) { fl += 6.0 * M * H * H; by += 2.0 * (M * H + 3.0 * H * H) + 12.0 * M * H; }
    ProfScope ps(KF_DECODE_CHAIN, st, fl, by);
    launch_k(decode_chain_kernel, dim3(n_sm), dim3(kThreads), smem, st, pdl, tATT, tWo, tXn, tWfc, tFF, tWproj, tWqkv, P);
    COUNT_LAUNCH(); KERNEL_CHECK();
}

}  // namespace xtts
