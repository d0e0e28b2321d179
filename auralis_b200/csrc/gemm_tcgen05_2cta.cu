// bf16 tensor-core GEMM for prefill-shaped problems (M >= 256):  out[M,N] = epi(A[M,K] . W[N,K]^T + bias)
//
// The one-tile-per-CTA kernel of gemm_tcgen05.cu tops out at ~0.25 of the tensor peak on large shapes: every CTA pays its
// own prologue, epilogue and pipeline ramp, and a 128 x 128 tile streams 2 operand bytes per 128 flops from shared memory.
// This kernel is the large-shape path:
//   * CTA PAIRS (thread-block cluster 2x1, tcgen05 cta_group::2): one UMMA covers 256 (M) x 256 (N) x 16; each CTA stages
//     its own 128 rows of A and HALF of the B tile (128 of the 256 weight rows), the instruction reads the other half from
//     the peer's shared memory — B traffic per CTA halves, the MMA runs at its full-rate shape;
//   * PERSISTENT: one cluster per SM pair walks the tile list; barriers, the 6-stage TMA ring and TMEM live across tiles;
//   * DOUBLE-BUFFERED ACCUMULATORS: two 256-column TMEM buffers (all 512 columns): the epilogue of tile i drains one
//     while the MMAs of tile i + 1 fill the other.
// Roles per CTA (320 threads): warp 0 TMA producer (both CTAs load; transaction bytes land on the LEADER's barrier),
// warp 1 TMEM allocation + (leader CTA only) MMA issue, warps 2..9 epilogue (TMEM lane quarter = warp % 4, two warpgroups
// on alternate column chunks: four warps alone could not drain a 128 x 256 fp32 tile in the time its MMAs take).
// Same watchdog discipline as the other tcgen05 kernels: every mbarrier wait traps instead of hanging the GPU.
#include <cuda.h>
#include "kernels.h"

namespace xtts {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
// epilogue warps EW = 8 or 16 (two or four warpgroups; warps w, w + 4, ... share a TMEM lane quarter and take alternate
// 32-column chunks).  A warp's chunk is a dependent chain (tcgen05.ld -> wait for its previous TMA store to have read the box ->
// shared stores -> fence -> TMA store) of ~2 us; with 8 warps a 128 x 256 fp32 tile takes ~8.7 us to drain — longer than its
// MMAs at K = 1024 (5.8 us), so short-K shapes were epilogue-bound (run 9: 8.7 us per tile at K = 1024, 22.9 us = the MMA
// time at K = 4096).  16 warps halve the chain per tile.
constexpr int threads_for(int ew) { return 64 + 32 * ew; }
constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2;      // per CTA and 64-wide k-block: 16 KB + 16 KB
// a ring stage holds KPS k-blocks (KPS = 2: half the barrier round trips per byte)
__host__ __device__ constexpr size_t smem_for(int stages, int kps, int ew) {
    return (size_t)stages * kps * (A_BYTES + B_BYTES) + (size_t)ew * 4096 + 1024;     // ring + epilogue boxes + alignment
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
    const long long t0 = clock64();
    int polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 1023) == 0 && clock64() - t0 > 4000000000LL) {
            printf("gemm_2cta: mbarrier watchdog (tag %d, block %d, thread %d)\n", tag, blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: the bytes land in THIS CTA's shared memory, the transaction count on the LEADER CTA's barrier (the barrier
// operand is a shared::cluster address; clearing the peer bit selects CTA 0 of the pair — cute::SM100_TMA_2SM_LOAD_2D)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}
// K-major SWIZZLE_128B descriptor (as gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D = f32, A = B = bf16, both K-major, M = 256 (the pair), N = BN
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// completion of every MMA issued so far -> arrive on `bar` in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)0b11) : "memory");
}

template <int STAGES, int KPS, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads_for(EW), 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, const float* resid, int M, int N,
                      int K, int flags) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];       // leader's copy is the live one (both CTAs' bytes land there)
    __shared__ __align__(8) uint64_t empty_bar[STAGES];      // one per CTA: the MMA commit is multicast to the pair
    __shared__ __align__(8) uint64_t tmem_full[2];           // one per CTA (multicast commit)
    __shared__ __align__(8) uint64_t tmem_empty[2];          // leader's copy: 4 epilogue warps x 2 CTAs arrive
    __shared__ uint32_t tmem_base_smem;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * KPS * A_BYTES;
    uint8_t* sC = smem + STAGES * KPS * (A_BYTES + B_BYTES);          // EW x 4 KB: one 32 x 32 output box per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * EW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // the same warp of both CTAs allocates all 512 columns for the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                               // the peer's barriers are initialised before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN;
    const int total = tiles_m * tiles_n, n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;
    const int num_kb = (K / BK + KPS - 1) / KPS;      // ring stages per tile (the last may be half filled)
    const int kb_total = K / BK;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int tile = cid; tile < total; tile += n_clusters) {
                const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM, n0 = (tile % tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], (uint32_t)(((it / STAGES) & 1) ^ 1), 1);
                    const int nk = min(KPS, kb_total - kb * KPS);
                    if (leader) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)nk * (A_BYTES + B_BYTES));
#pragma unroll
                    for (int j = 0; j < KPS; ++j)
                        if (j < nk) {
                            tma_load_2d_2sm(sA + (s * KPS + j) * A_BYTES, &tmA, &full_bar[s], (kb * KPS + j) * BK, m0);
                            tma_load_2d_2sm(sB + (s * KPS + j) * B_BYTES, &tmB, &full_bar[s], (kb * KPS + j) * BK, n0);
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            int it = 0, lt = 0;
            const uint32_t idesc = (flags & GEMM_F16) ? (kIdesc & ~((1u << 7) | (1u << 10))) : kIdesc;      // fp16 / bf16 operands
            for (int tile = cid; tile < total; tile += n_clusters, ++lt) {
                const int ab = lt & 1;
                mbar_wait(&tmem_empty[ab], (uint32_t)(((lt >> 1) & 1) ^ 1), 2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(ab * BN);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (uint32_t)((it / STAGES) & 1), 3);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const int nk = min(KPS, kb_total - kb * KPS);
#pragma unroll
                    for (int j = 0; j < KPS; ++j) {
                        if (j >= nk) break;
                        const uint32_t a_addr = smem_u32(sA + (s * KPS + j) * A_BYTES), b_addr = smem_u32(sB + (s * KPS + j) * B_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16_2cta(acc, make_sw128_desc(a_addr + k * UMMA_K * 2), make_sw128_desc(b_addr + k * UMMA_K * 2),
                                           idesc, (kb | j | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_pair(&empty_bar[s]);        // the slot is free in both CTAs once these MMAs have read it
                }
                umma_commit_pair(&tmem_full[ab]);           // accumulators of both CTAs complete
            }
        }
    } else {
        // ---------------- epilogue: warp w owns TMEM lanes [32*(w%4), +32) = this CTA's tile rows; the two warpgroups take
        // alternate 32-column chunks of the 256-column accumulator
        const int q = warp & 3, grp = (warp - 2) >> 2;
        int lt = 0;
        for (int tile = cid; tile < total; tile += n_clusters, ++lt) {
            const int ab = lt & 1;
            const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM, n0 = (tile % tiles_n) * BN;
            mbar_wait(&tmem_full[ab], (uint32_t)((lt >> 1) & 1), 4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
#pragma unroll 1
            for (int c = grp; c < BN / 32; c += EW / 4) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
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
                if (c + EW / 4 >= BN / 32) {
                    // this warp has read the last of the buffer: hand it back to the (leader's) MMA thread before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&tmem_empty[ab], 0);
                }
                // ---- the chunk leaves through shared memory and a TMA store: a lane holds 32 consecutive columns of ONE row,
                // so written directly every store instruction touches 32 different 128-byte lines (16 B each) and the LSU —
                // not HBM — bounds the epilogue (measured: 12 us per 128 x 256 fp32 tile, twice the tile's MMA time at
                // K = 1024).  Staged as a 128B-swizzled 32 x 32 box (conflict-free 16-byte shared stores), one
                // cp.async.bulk.tensor store writes whole rows and clips the tile tails itself.
                const int nb = n0 + c * 32;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x = __uint_as_float(r[i]);
                    if (bias && nb + i < N) x += __ldg(bias + nb + i);
                    if (flags & GEMM_GELU) x = gelu_new(x);
                    v[i] = x;
                }
                if ((flags & GEMM_RESID) && row < M && nb < N) {
                    const float4* rp = reinterpret_cast<const float4*>(resid + (size_t)row * N + nb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = rp[i];
                        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                    }
                }
                uint8_t* box = sC + (warp - 2) * 4096;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the box
                __syncwarp();
                if (flags & GEMM_OUT_BF16) {
                    // 32 rows x 64 B, SWIZZLE_64B: 16-byte chunk j of row r sits at chunk j ^ ((r >> 1) & 3)
                    const bool f16 = (flags & GEMM_F16) != 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        pk.x = pack16(v[8 * j], v[8 * j + 1], f16); pk.y = pack16(v[8 * j + 2], v[8 * j + 3], f16);
                        pk.z = pack16(v[8 * j + 4], v[8 * j + 5], f16); pk.w = pack16(v[8 * j + 6], v[8 * j + 7], f16);
                        *reinterpret_cast<uint4*>(box + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
                    }
                } else {
                    // 32 rows x 128 B, SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> the TMA engine's reads
                __syncwarp();
                if (lane == 0 && nb < N && m0 + q * 32 < M) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmC), "r"(smem_u32(box)), "r"(nb), "r"(m0 + q * 32) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store of this warp has landed
    }
    // nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory or signal its barriers
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode2 = nullptr;

void encode_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = g_encode2(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("gemm_2cta: cuTensorMapEncodeTiled failed");
}

}  // namespace

bool gemm_2cta_supported(int M, int N, int K) { return M >= 256 && N >= 256 && N % 32 == 0 && K % BK == 0 && K >= BK; }

void launch_gemm_bf16_2cta(const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias, const float* resid, void* out,
                           int M, int N, int K, int flags, cudaStream_t st) {
    if (!gemm_2cta_supported(M, N, K)) throw CudaError("gemm_2cta: unsupported shape");
    if (!g_encode2) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { (void)cudaGetLastError(); throw CudaError("cuTensorMapEncodeTiled not available"); }
        g_encode2 = reinterpret_cast<EncodeTiledFn>(fn);
    }
    static bool attr[64] = {};
    if (first_on_device(attr)) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<3, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(3, 2, 8)));
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<4, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(4, 1, 8)));
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<6, 1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(6, 1, 8)));
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<2, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(2, 2, 8)));
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<2, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(2, 2, 16)));
        CUDA_CHECK(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<5, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(5, 1, 16)));
    }
    static int n_sm = 0;
    if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    CUtensorMap tmA, tmB, tmC;
    encode_2d(&tmA, A, (uint64_t)M, (uint64_t)K, BM);
    encode_2d(&tmB, W, (uint64_t)N, (uint64_t)K, BN / 2);
    {   // output [M, N] row-major, 32 x 32 boxes; the swizzle spans one box row (128 B fp32, 64 B bf16)
        const bool h = (flags & GEMM_OUT_BF16) != 0;
        const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)N * (h ? 2 : 4)};
        const cuuint32_t box[2] = {32, 32};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = g_encode2(&tmC, h ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides,
                                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, h ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) throw CudaError("gemm_2cta: cuTensorMapEncodeTiled (output) failed");
    }
    const int total = ceil_div(M, 2 * BM) * ceil_div(N, BN);
    const int clusters = std::min(total, n_sm / 2);
    ProfScope ps(KF_GEMM_TC, st, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + ((flags & GEMM_OUT_BF16) ? 2.0 : 4.0) * M * N);
    // g_gemm_2cta selects the variant (ring stages x k-blocks per stage, epilogue warps):
    //   1 (default): K < 2048 -> 5 x 1 ring, 16 epilogue warps (short K is epilogue-bound); else 3 x 2 ring, 8 warps (run 8)
    //   2: 4 x 1, 8    3: 6 x 1, 8    4: 2 x 2, 8    5: 3 x 2, 8 for every K    6: 2 x 2, 16    7: 5 x 1, 16 for every K
    const dim3 grid(2 * clusters);
    int v = g_gemm_2cta;
    if (v == 1) v = (K < 2048) ? 7 : 5;
#define XTTS_2CTA(ST, KP, E) gemm_bf16_2cta_kernel<ST, KP, E><<<grid, dim3(threads_for(E)), smem_for(ST, KP, E), st>>>(tmA, tmB, tmC, bias, resid, M, N, K, flags)
    switch (v) {
        case 2: XTTS_2CTA(4, 1, 8); break;
        case 3: XTTS_2CTA(6, 1, 8); break;
        case 4: XTTS_2CTA(2, 2, 8); break;
        case 6: XTTS_2CTA(2, 2, 16); break;
        case 7: XTTS_2CTA(5, 1, 16); break;
        default: XTTS_2CTA(3, 2, 8); break;
    }
#undef XTTS_2CTA
    COUNT_LAUNCH(); KERNEL_CHECK();
}

}  // namespace xtts
