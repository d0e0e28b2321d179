// Dilated Conv1d ("same" padding) as an implicit GEMM on the 5th-gen tensor cores (fast mode of the vocoder).
//
//   out[co][t] (=|+=) bias[co] + cbias[co] + resid[co][t] + sum_j sum_ci W_j[co][ci] * act(s * x[ci][t + (j-c)*d])
//
// Mapping (per CTA):  D[M = 128*NACC time steps, N = C_out tile] += A[M, K=16 ci] * B[N, K=16 ci]^T per tap and ci-step
//   * A (activations): loader warps read fp32 x[ci][t] coalesced along t, apply scale + leaky-relu, convert to fp16 and
//     store 16-byte atoms into the UMMA *no-swizzle K-major* layout  [ci/8 plane][time row][8 ci]  (SBO = 128 B, so
//     rows are linear in time and a tap is just a start-address shift of j*d rows — no im2col copy);
//     staged per 64-channel chunk in a ring, reused by every tap.
//   * B (weights): pre-packed on the host into the same atom layout per (ci-chunk, tap); each tile is ONE contiguous
//     cp.async.bulk (TMA 1-D) into a second ring.
//   * tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in TMEM (NACC accumulators of N columns share each B tile),
//     epilogue warps read TMEM (lane = time step) and store channel-major fp32, coalesced along t.
// Warp roles: warps 0-3 A-loaders then epilogue (TMEM lane quarter = warp), warp 4 TMA producer, warp 5 MMA issuer.
//
// Replaces the cuDNN fp16-autocast Conv1d calls of HifiganGenerator.forward / ResBlock1.forward
// (hifigan_decoder.py:76-91,241-259; the reference runs them in fp16 under torch.amp.autocast on GPU, App. B.8).
#include "kernels.h"

namespace xtts {
namespace {

constexpr int kThreadsTC = 192;
constexpr int SA = 2, SB = 3;          // ring depths (activation chunks, weight tiles)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(s_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity, int tag) {
    const long long t0 = clock64();
    while (!bar_try(b, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("conv1d_tc: mbarrier watchdog (tag %d, block %d,%d,%d thread %d)\n", tag, blockIdx.x, blockIdx.y, blockIdx.z,
                   threadIdx.x);
            __trap();
        }
    }
}
// no-swizzle K-major descriptor: rows 16 B apart (SBO = 128 B per 8 rows), the two 8-element K halves `lbo` bytes apart
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell); layout_type = 0 (SWIZZLE_NONE)
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(b)) : "memory");
}
__device__ __forceinline__ float lrelu_s(float v, float slope) { return v > 0.f ? v : v * slope; }

struct ConvTcParams {
    const float* x; const __half* wblob; const float* bias; const float* cbias; const float* resid; float* out;
    int Cin, Cout, L, K, dil, mode;
    float in_scale, slope;
    int N;          // output channels per CTA (<= 256, multiple of 16)
    int CK;         // input channels per chunk (<= 64, multiple of 16)
    int rows;       // time rows staged per chunk = 128*NACC + (K-1)*dil
    int cbias_bs;   // elements between the speaker-bias vectors of consecutive batch items
};

template <int NACC>
__global__ void __launch_bounds__(kThreadsTC, 1)
conv1d_tc_kernel(const ConvTcParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], tmem_full;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int halo = (P.K - 1) / 2 * P.dil;
    const int T0 = blockIdx.x * (128 * NACC);
    const int n0 = blockIdx.y * P.N;
    const int planes = P.CK / 8, ksteps = P.CK / 16, nch = P.Cin / P.CK;
    const uint32_t a_plane = (uint32_t)P.rows * 16u;               // bytes per ci-plane of a staged chunk
    const uint32_t a_stage = a_plane * planes;
    const uint32_t b_plane = (uint32_t)P.N * 16u;
    const uint32_t b_stage = b_plane * planes;
    uint8_t* sA = smem;
    uint8_t* sB = smem + ((SA * a_stage + 127) & ~127u);
    const float* x = P.x + (size_t)blockIdx.z * P.Cin * P.L;
    float* out = P.out + (size_t)blockIdx.z * P.Cout * P.L;
    const float* resid = P.resid ? P.resid + (size_t)blockIdx.z * P.Cout * P.L : nullptr;
    const float* cbias = P.cbias ? P.cbias + (size_t)blockIdx.z * P.cbias_bs : nullptr;
    constexpr uint32_t TM_COLS_MAX = 512;
    uint32_t tm_cols = 32;
    while (tm_cols < (uint32_t)(NACC * P.N)) tm_cols <<= 1;
    if (tm_cols > TM_COLS_MAX) tm_cols = TM_COLS_MAX;

    if (threadIdx.x == 0) {
        for (int i = 0; i < SA; ++i) { bar_init(&a_full[i], 128); bar_init(&a_empty[i], 1); }
        for (int i = 0; i < SB; ++i) { bar_init(&b_full[i], 1); bar_init(&b_empty[i], 1); }
        bar_init(&tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_s)), "r"(tm_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
        // ------------------------------------------------ activation loaders (128 threads)
        const int tid = threadIdx.x;
        const int items = planes * P.rows;
        for (int c = 0; c < nch; ++c) {
            const int s = c % SA;
            bar_wait(&a_empty[s], ((c / SA) & 1) ^ 1, 1);
            uint8_t* dst = sA + (size_t)s * a_stage;
            for (int it = tid; it < items; it += 128) {
                const int p = it / P.rows, r = it - p * P.rows;
                const int gt = T0 - halo + r;
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (gt >= 0 && gt < P.L) {
                    const float* src = x + (size_t)(c * P.CK + p * 8) * P.L + gt;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = lrelu_s(P.in_scale * __ldg(src + (size_t)e * P.L), P.slope);
                    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
                    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
                    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                }
                *reinterpret_cast<uint4*>(dst + (size_t)p * a_plane + (size_t)r * 16) = pk;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy (UMMA)
            bar_arrive(&a_full[s]);
        }
        // ------------------------------------------------ epilogue: warp q owns TMEM lanes [32q, 32q+32) = time rows
        const int q = warp;
        bar_wait(&tmem_full, 0, 4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int a = 0; a < NACC; ++a) {
            const int t = T0 + a * 128 + q * 32 + lane;
#pragma unroll 1
            for (int nc = 0; nc < P.N / 32; ++nc) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * P.N + nc * 32);
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
                if (t < P.L) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int co = n0 + nc * 32 + i;
                        if (co < P.Cout) {
                            const size_t o = (size_t)co * P.L + t;
                            float v = __uint_as_float(r[i]) + (P.bias ? __ldg(P.bias + co) : 0.f) + (cbias ? __ldg(cbias + co) : 0.f);
                            if (resid) v += resid[o];
                            if (P.mode == CONV_ACCUM) v += out[o];
                            out[o] = v;
                        }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------ weight tiles: one bulk copy per (ci-chunk, tap)
        if (lane == 0) {
            const __half* wsrc = P.wblob + (size_t)blockIdx.y * nch * P.K * (b_stage / 2);
            int it = 0;
            for (int c = 0; c < nch; ++c)
                for (int j = 0; j < P.K; ++j, ++it) {
                    const int s = it % SB;
                    bar_wait(&b_empty[s], ((it / SB) & 1) ^ 1, 2);
                    bar_expect_tx(&b_full[s], b_stage);
                    const __half* src = wsrc + (size_t)it * (b_stage / 2);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s_u32(sB + (size_t)s * b_stage)), "l"(src), "r"(b_stage), "r"(s_u32(&b_full[s])) : "memory");
                }
        }
    } else {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
            int it = 0;
            for (int c = 0; c < nch; ++c) {
                const int sa = c % SA;
                bar_wait(&a_full[sa], (c / SA) & 1, 3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = s_u32(sA + (size_t)sa * a_stage);
                for (int j = 0; j < P.K; ++j, ++it) {
                    const int sb = it % SB;
                    bar_wait(&b_full[sb], (it / SB) & 1, 5);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t b_addr = s_u32(sB + (size_t)sb * b_stage);
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t bd = desc_nosw(b_addr + (uint32_t)(2 * k) * b_plane, b_plane);
#pragma unroll
                        for (int a = 0; a < NACC; ++a) {
                            const uint64_t ad = desc_nosw(a_addr + (uint32_t)(2 * k) * a_plane + (uint32_t)(a * 128 + j * P.dil) * 16u, a_plane);
                            mma_f16(tmem_base + (uint32_t)(a * P.N), ad, bd, idesc, (c | j | k) != 0 ? 1u : 0u);
                        }
                    }
                    mma_commit(&b_empty[sb]);
                }
                mma_commit(&a_empty[sa]);
            }
            mma_commit(&tmem_full);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tm_cols) : "memory");
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
ConvTcPlan conv1d_tc_plan(int Cin, int Cout, int K) {
    ConvTcPlan pl{};
    pl.N = Cout > 256 ? 256 : Cout;
    pl.CK = Cin >= 64 ? 64 : Cin;
    pl.n_tiles = ceil_div(Cout, pl.N);
    pl.nacc = pl.N >= 256 ? 2 : 4;
    pl.ok = (Cin % pl.CK == 0) && (pl.CK % 16 == 0) && (pl.N % 32 == 0) && (Cout % pl.N == 0) && (K % 2 == 1);
    pl.tile_halves = (size_t)(pl.CK / 8) * pl.N * 8;
    pl.blob_halves = (size_t)pl.n_tiles * (Cin / pl.CK) * K * pl.tile_halves;
    return pl;
}

// w: folded Conv1d weight [Cout][Cin][K] fp32 -> blob [n_tile][ci-chunk][tap][plane][co][8] fp16
void conv1d_tc_pack(const float* w, int Cin, int Cout, int K, const ConvTcPlan& pl, __half* blob) {
    const int nch = Cin / pl.CK, planes = pl.CK / 8;
    for (int nt = 0; nt < pl.n_tiles; ++nt)
        for (int c = 0; c < nch; ++c)
            for (int j = 0; j < K; ++j) {
                __half* tile = blob + (((size_t)nt * nch + c) * K + j) * pl.tile_halves;
                for (int p = 0; p < planes; ++p)
                    for (int n = 0; n < pl.N; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int co = nt * pl.N + n, ci = c * pl.CK + p * 8 + e;
                            tile[((size_t)p * pl.N + n) * 8 + e] = __float2half_rn(w[((size_t)co * Cin + ci) * K + j]);
                        }
            }
}

void launch_conv1d_tc(const float* x, const __half* wblob, const ConvTcPlan& pl, const float* bias, const float* cbias,
                      const float* resid, float* out, int Cin, int Cout, int L, int K, int dil, float in_scale, float slope,
                      int mode, int batch, int cbias_batch_stride, cudaStream_t st) {
    if (L <= 0 || batch <= 0) return;
    if (!pl.ok) throw CudaError("conv1d_tc: unsupported channel geometry");
    ConvTcParams P{};
    P.x = x; P.wblob = wblob; P.bias = bias; P.cbias = cbias; P.resid = resid; P.out = out;
    P.Cin = Cin; P.Cout = Cout; P.L = L; P.K = K; P.dil = dil; P.mode = mode; P.in_scale = in_scale; P.slope = slope;
    P.N = pl.N; P.CK = pl.CK;
    P.rows = 128 * pl.nacc + (K - 1) * dil;
    P.cbias_bs = cbias_batch_stride;
    const size_t a_stage = (size_t)P.rows * 16 * (pl.CK / 8), b_stage = (size_t)pl.N * 16 * (pl.CK / 8);
    const size_t smem = ((SA * a_stage + 127) & ~(size_t)127) + SB * b_stage + 128;
    constexpr int kMaxDyn = 227 * 1024 - 2048;      // opt-in limit minus this kernel's static shared memory
    if (smem > (size_t)kMaxDyn) throw CudaError("conv1d_tc: shared memory budget exceeded");
    dim3 grid(ceil_div(L, 128 * pl.nacc), pl.n_tiles, batch);
    ProfScope ps(KF_CONV1D_TC, st, 2.0 * Cin * Cout * K * (double)L * batch,
                 batch * 4.0 * ((double)L * (Cin + Cout * (1 + (resid ? 1 : 0) + (mode == CONV_ACCUM ? 1 : 0)))) + 2.0 * Cin * Cout * K);
    static bool attr2 = false, attr4 = false;
    if (pl.nacc == 2) {
        if (!attr2) { CUDA_CHECK(cudaFuncSetAttribute(conv1d_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn)); attr2 = true; }
        conv1d_tc_kernel<2><<<grid, kThreadsTC, smem, st>>>(P);
    } else {
        if (!attr4) { CUDA_CHECK(cudaFuncSetAttribute(conv1d_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDyn)); attr4 = true; }
        conv1d_tc_kernel<4><<<grid, kThreadsTC, smem, st>>>(P);
    }
    COUNT_LAUNCH(); KERNEL_CHECK();
}

}  // namespace xtts
