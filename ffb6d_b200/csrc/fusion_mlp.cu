// fusion_mlp.cu -- the 1x1 fusion MLP of FFB6D on 5th-generation tensor cores (sm_100a).
//
// Reference: pt_utils.Conv2d(in, out, kernel_size=(1,1), bn=True) = conv(bias=False) ->
// BatchNorm2d -> ReLU (models/pytorch_utils.py:75-129, 168-201), 28 instances built at
// models/ffb6d.py:55-80, 104-129 and applied to torch.cat((a, b), dim=1) (:246-262, 282-298).
// With frozen (eval) statistics the layer is, per frame,
//     out[co, p] = relu( scale[co] * sum_ci W[co, ci] * X[ci, p] + shift[co] ),   X = [X1; X2]
// i.e. a dense GEMM D[Co x P] = W[Co x Ci] * X[Ci x P] with a per-row affine + ReLU epilogue.
//
// Kernels in this file:
//   fusion_mlp_packed_kernel  the product path: weights pre-split once (ffb6d_fusion_mlp_pack) and
//                             fetched by bulk-async copies, warp-specialised mbarrier pipeline,
//                             16 staging warps (K > 128) or the 3-CTA/SM DIRECT variant (K <= 128)
//
// All of them fuse the concat (two K ranges read from two tensors), the GEMM, BN and ReLU:
//   * tcgen05.mma kind::tf32, M = N = 128 per CTA, accumulators in TMEM (128 lanes x 128 columns)
//   * fp32 fidelity through 3xTF32: every fp32 operand is split into hi = tf32(x) and
//     lo = tf32(x - hi) while it is staged, and D += Ahi*Bhi + Alo*Bhi + Ahi*Blo (the dropped lo*lo
//     term is 2^-22 relative): results agree with the fp32 cuDNN path to ~1e-6 relative, inside the
//     1e-5 contract of BASELINE.json
//   * operands are staged by the CTA's threads (the split needs a register pass anyway) into the
//     canonical K-major no-swizzle UMMA layout: 8-row x 16-byte core matrices, 128 B between 8-row
//     groups (SBO), 2048 B between 16-byte K chunks (LBO); X is transposed on the fly (it is
//     point-major in memory)
//   * two stages: the MMAs of stage s run asynchronously (completion -> mbarrier via
//     tcgen05.commit) while the threads stage s+1
//   * the tensor core adds into its fp32 accumulator with truncation, which drifts by ~2^-24 per
//     MMA (measured 1.2e-5 relative at K = 2048); K is therefore accumulated in chunks of 128 into
//     two alternating TMEM accumulators that the warps drain into round-to-nearest fp32 register
//     sums while the next chunk is being multiplied
//   * epilogue: tcgen05.ld 32x32b -> scale/shift/ReLU in registers -> 128-bit stores along the
//     point axis (NCHW output, no transposition needed because M = output channel = TMEM lane).
#include "common.cuh"

#include <algorithm>
#include <stdlib.h>
#include <type_traits>

namespace ffb6d {

constexpr int TM = 128, TN = 128, TK = 32;            // CTA tile
constexpr int CHUNK_BYTES = TM * 16;                  // one 16-byte K chunk of all 128 rows
constexpr int TILE_BYTES = (TK / 4) * CHUNK_BYTES;    // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int CH = 4;                                 // k-tiles per accumulation chunk (K = 128)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// round-to-nearest (ties away) to TF32's 10-bit mantissa with two full-rate integer ops; identical to
// cvt.rna.tf32.f32 for finite inputs (that conversion runs on the quarter-rate conversion pipe and was
// the staging bottleneck)
__device__ __forceinline__ float to_tf32(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

#ifndef FFB6D_MLP_ROUND_LO
#define FFB6D_MLP_ROUND_LO 1   // 0: leave the low part unrounded (the tensor core truncates it): 2 ops fewer per element
#endif
__device__ __forceinline__ float lo_part(float d) { return FFB6D_MLP_ROUND_LO ? to_tf32(d) : d; }

__device__ __forceinline__ void split4(const float4 v, float4 &hi, float4 &lo)
{
    hi.x = to_tf32(v.x); lo.x = lo_part(v.x - hi.x);
    hi.y = to_tf32(v.y); lo.y = lo_part(v.y - hi.y);
    hi.z = to_tf32(v.z); lo.z = lo_part(v.z - hi.z);
    hi.w = to_tf32(v.w); lo.w = lo_part(v.w - hi.w);
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(CHUNK_BYTES >> 4) << 16)   // LBO: next K chunk
           | ((uint64_t)(128 >> 4) << 32)                                              // SBO: next 8 rows
           | (1ull << 46);                                                             // version 1 (sm100)
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128 (InstrDescriptor bit layout)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(a), "l"(b), "r"(kIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}\n"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1 << 26)) __trap();   // never hang the GPU on a protocol bug
    }
}

// ===================================================================== packed-weight kernel
// The weights are constants at inference: their hi/lo split is computed once
// (fusion_mlp_pack_kernel) and written in exactly the shared-memory image the A operand needs, one
// contiguous 32 KB block per (128-row tile, 32-column k-tile): [A_hi | A_lo], each 8 chunks x 128
// rows x 16 B.  The main kernel is warp-specialised:
//   warps 0-7  stage the activations (global -> registers -> split -> UMMA layout), drain the
//              accumulator chunks into round-to-nearest register sums, run the epilogue
//   warp  8    producer: one cp.async.bulk (TMA, no tensor map needed for a flat block) per k-tile
//              brings the packed A block straight into shared memory
//   warp  9    issues the tcgen05.mma's as soon as both halves of a stage have landed
// Three stages, mbarriers only (no CTA-wide barrier in the main loop): full_a (TMA transaction
// bytes), full_b (one arrival per staging warp), empty (tcgen05.commit), chunk (as above).
// Layers with K <= 128 (one accumulation chunk; the P = 76800 / 19200 image-map layers, which are
// bound by their activation traffic, not by the MMAs) use the DIRECT variant: a single stage and no
// register sums, so that two CTAs fit on an SM and one's prologue / epilogue overlaps the other's
// main loop.
constexpr int PACK_BLOCK_BYTES = 2 * TILE_BYTES;                 // A_hi + A_lo of one k-tile

// Optional extras of the epilogue (all null / 0 for the plain layer):
//   addend [B, NA, Co] (channels-last) + idx [B, P]: out = act(scale * (W x + addend[b, idx[b, p], :]) + shift)
//   out_nsc: the result is stored channels-last, [B, P, Co]
struct MlpEpilogue {
    const float *addend;
    const void *idx;
    int NA;
    int idx_is_i64;
    int out_nsc;
};
constexpr int mlp2_threads(int nsw) { return (nsw + 2) * 32; }   // staging warps + producer + issuer
constexpr int mlp2_stage(int mt) { return (2 * mt + 2) * TILE_BYTES; }           // MT x [A_hi | A_lo], B_hi, B_lo
constexpr int mlp2_smem(int nst, int mt = 1) { return nst * mlp2_stage(mt) + 128; }

__global__ void __launch_bounds__(256)
fusion_mlp_pack_kernel(const float *__restrict__ w, int Co, int Ci, int nk, int nblocks, float4 *__restrict__ packed)
{
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one (block, chunk, row) each
    if (o >= (long long)nblocks * 1024) return;
    const int blk = (int)(o >> 10), r = (int)(o & 1023), c = r >> 7, m = r & 127;
    const int gm = (blk / nk) * TM + m, gk = (blk % nk) * TK + 4 * c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gm < Co) {
        const float *src = w + (size_t)gm * Ci + gk;
        if (gk + 0 < Ci) v.x = __ldg(src + 0);
        if (gk + 1 < Ci) v.y = __ldg(src + 1);
        if (gk + 2 < Ci) v.z = __ldg(src + 2);
        if (gk + 3 < Ci) v.w = __ldg(src + 3);
    }
    float4 hi, lo;
    split4(v, hi, lo);
    float4 *dst = packed + (size_t)blk * (PACK_BLOCK_BYTES / 16);
    dst[r] = hi;
    dst[TILE_BYTES / 16 + r] = lo;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// MT = 128-row tiles of W per CTA (1 or 2).  With MT = 2 a CTA multiplies ONE staged activation tile with two weight
// tiles (256 output channels): the activation split + staging -- the instruction-bound part of this kernel, repeated
// by every row tile of a layer -- halves per FLOP; all 512 TMEM columns are in use (2 row tiles x 2 chunk buffers).
template <int NST, bool DIRECT, int NSW, int PD, int MT = 1>
__global__ void __launch_bounds__(mlp2_threads(NSW), DIRECT ? 3 : 1)
fusion_mlp_packed_kernel(const float *__restrict__ x1, int C1, const float *__restrict__ x2, int C2,
                         const unsigned char *__restrict__ wpack, const float *__restrict__ scale,
                         const float *__restrict__ shift, float *__restrict__ out, int Co, int P, int act,
                         float slope, MlpEpilogue ep)
{
    static_assert(MT == 1 || !DIRECT, "the single-chunk variant keeps one row tile");
    constexpr int STAGE = mlp2_stage(MT);
    constexpr int TMEM_COLS = DIRECT ? TN : 2 * TN * MT;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + NST * STAGE);
    uint64_t *full_a = bars, *full_b = bars + NST, *empty = bars + 2 * NST, *chunk = bars + 3 * NST;   // 11 barriers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + NST * STAGE + 96);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, m0 = blockIdx.y * TM * MT, n0 = blockIdx.x * TN;
    const int mt_valid = min(MT, (Co - m0 + TM - 1) / TM);   // row tiles of this CTA that exist
    const int Ci = C1 + C2;
    const int nk = (Ci + TK - 1) / TK;

    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        for (int i = 0; i < NST; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(full_a + i)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(full_b + i)), "r"(NSW));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(empty + i)));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(chunk + 0)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(chunk + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (wid == NSW) {
        // ---------------- A producer
        if (lane == 0) {
            const unsigned char *src = wpack + (size_t)blockIdx.y * MT * nk * PACK_BLOCK_BYTES;
            for (int kt = 0; kt < nk; ++kt) {
                const int st = kt % NST, n = kt / NST;
                if (n >= 1) mbar_wait(smem_u32(empty + st), (uint32_t)((n - 1) & 1));
                const uint32_t bar = smem_u32(full_a + st);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(mt_valid * PACK_BLOCK_BYTES) : "memory");
                for (int mt = 0; mt < mt_valid; ++mt)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem_u32(smem + st * STAGE + mt * PACK_BLOCK_BYTES)),
                                    "l"(src + ((size_t)mt * nk + kt) * PACK_BLOCK_BYTES), "r"(PACK_BLOCK_BYTES), "r"(bar) : "memory");
            }
        }
        __syncwarp();
    } else if (wid == NSW + 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            for (int kt = 0; kt < nk; ++kt) {
                const int st = kt % NST, n = kt / NST;
                mbar_wait(smem_u32(full_a + st), (uint32_t)(n & 1));
                mbar_wait(smem_u32(full_b + st), (uint32_t)(n & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // this lane shares its scheduler with four staging warps: keep its per-tile instruction
                // count small (one descriptor per operand tile, the k-steps are +256 in the address field)
                const uint32_t s0 = smem_u32(smem + st * STAGE);
                const uint64_t db_hi = umma_desc(s0 + 2 * MT * TILE_BYTES), db_lo = umma_desc(s0 + (2 * MT + 1) * TILE_BYTES);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    if (mt < mt_valid) {
                        const uint64_t da_hi = umma_desc(s0 + mt * PACK_BLOCK_BYTES), da_lo = umma_desc(s0 + mt * PACK_BLOCK_BYTES + TILE_BYTES);
                        const uint32_t d_buf = tmem_d + (uint32_t)((mt * 2 + ((kt / CH) & 1)) * TN);
#pragma unroll
                        for (int j = 0; j < TK / 8; ++j) {
                            const uint64_t off = (uint64_t)(j * ((2 * CHUNK_BYTES) >> 4));
                            umma_tf32(d_buf, da_hi + off, db_hi + off, (kt % CH != 0 || j > 0) ? 1u : 0u);
                            umma_tf32(d_buf, da_lo + off, db_hi + off, 1u);
                            umma_tf32(d_buf, da_hi + off, db_lo + off, 1u);
                        }
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             :: "r"(smem_u32(empty + st)) : "memory");
                if (kt % CH == CH - 1 || kt == nk - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 :: "r"(smem_u32(chunk + ((kt / CH) & 1))) : "memory");
            }
        }
        __syncwarp();
    } else {
        // ---------------- activation staging, accumulator drain, epilogue
        const float *xb1 = x1 + (size_t)b * C1 * P;
        const float *xb2 = x2 ? x2 + (size_t)b * C2 * P : nullptr;
        const bool vec = ((P & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb1) & 15) == 0) &&
                         (!xb2 || (reinterpret_cast<uintptr_t>(xb2) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        constexpr int COLS = 4 * TN / NSW, NI = COLS / 32;   // accumulator columns per thread: 64 or 32
        const int q = wid & 3, cs = wid >> 2;                 // TMEM lane quarter (fixed by the warp id), column slice
        float acc[MT][DIRECT ? 1 : COLS];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int i = 0; i < (DIRECT ? 1 : COLS); ++i) acc[mt][i] = 0.f;
        // columns COLS*cs + 32*i .. +31 of accumulator `buf`, this thread's row
        auto tmem_load32 = [&](int buf, int i, uint32_t (&v)[32]) {
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TN + COLS * cs + 32 * i);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        };
        auto drain = [&](int c) {
            if constexpr (!DIRECT) {
                const int buf = c & 1;
                mbar_wait(smem_u32(chunk + buf), (uint32_t)((c >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int i = 0; i < COLS / 16; ++i) {   // 16 columns at a time keeps the temporaries small
                    uint32_t v[16];
                    const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)((mt * 2 + buf) * TN + COLS * cs + 16 * i);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[mt][16 * i + j] = __fadd_rn(acc[mt][16 * i + j], __uint_as_float(v[j]));
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        };
        // X tile (32 k x 128 n): warp -> K chunk kg (4 k rows); with 8 staging warps a lane takes 4 n
        // (128-bit loads), with 16 the two warps of a chunk take 64 n each and a lane 2 n (64-bit loads).
        // Either way a thread ends up with whole 16-byte (n; k..k+3) slots of the K-major UMMA layout.
        constexpr int NPL = (NSW == 8) ? 4 : 2;                 // n per lane
        typedef typename std::conditional<NSW == 8, float4, float2>::type ldt;
        const int kg = wid & 7, nl = (NSW == 8) ? 4 * lane : 64 * (wid >> 3) + 2 * lane;
        // fast path: whole tile inside the tensors, vector loads, and a 4-row K chunk never straddles the
        // concat boundary -> one base pointer per tile, four loads at a row stride
        const bool fast = vec && ((C1 & 3) == 0) && (n0 + TN <= P);
        auto load_b = [&](int kt, ldt (&rb)[4]) {
            const int gk0 = kt * TK + 4 * kg;
            if (fast && gk0 + 3 < Ci) {
                const float *p0 = ((gk0 < C1) ? xb1 + (size_t)gk0 * P : xb2 + (size_t)(gk0 - C1) * P) + (n0 + nl);
#pragma unroll
                for (int j = 0; j < 4; ++j) rb[j] = __ldg(reinterpret_cast<const ldt *>(p0 + (size_t)j * P));
                return;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gk = kt * TK + 4 * kg + j, gn = n0 + nl;
                float e[4] = {0.f, 0.f, 0.f, 0.f};
                if (gk < Ci) {
                    const float *row = (gk < C1) ? xb1 + (size_t)gk * P : xb2 + (size_t)(gk - C1) * P;
                    if (vec && gn + NPL - 1 < P) {
                        const ldt v = __ldg(reinterpret_cast<const ldt *>(row + gn));
                        e[0] = v.x;
                        e[1] = v.y;
                        if constexpr (NSW == 8) {
                            e[2] = v.z;
                            e[3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < NPL; ++u)
                            if (gn + u < P) e[u] = __ldg(row + gn + u);
                    }
                }
                if constexpr (NSW == 8) rb[j] = make_float4(e[0], e[1], e[2], e[3]);
                else rb[j] = make_float2(e[0], e[1]);
            }
        };
        auto store_b = [&](int st, const ldt (&rb)[4]) {
            unsigned char *sB_hi = smem + st * STAGE + 2 * MT * TILE_BYTES, *sB_lo = sB_hi + TILE_BYTES;
            unsigned char *bh = sB_hi + kg * CHUNK_BYTES + nl * 16, *bl = sB_lo + kg * CHUNK_BYTES + nl * 16;
            float4 hi, lo;
            if constexpr (NSW == 8) {
                const float4 t0 = make_float4(rb[0].x, rb[1].x, rb[2].x, rb[3].x);   // n = nl + 0, k = 4kg .. 4kg+3
                const float4 t1 = make_float4(rb[0].y, rb[1].y, rb[2].y, rb[3].y);
                const float4 t2 = make_float4(rb[0].z, rb[1].z, rb[2].z, rb[3].z);
                const float4 t3 = make_float4(rb[0].w, rb[1].w, rb[2].w, rb[3].w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {   // rotated slots: every quarter-warp covers all 32 banks
                    const int x = (i + (lane >> 1)) & 3;
                    const float4 tx = (x == 0) ? t0 : (x == 1) ? t1 : (x == 2) ? t2 : t3;
                    split4(tx, hi, lo);
                    *reinterpret_cast<float4 *>(bh + 16 * x) = hi;
                    *reinterpret_cast<float4 *>(bl + 16 * x) = lo;
                }
            } else {
                const float4 t0 = make_float4(rb[0].x, rb[1].x, rb[2].x, rb[3].x);
                const float4 t1 = make_float4(rb[0].y, rb[1].y, rb[2].y, rb[3].y);
#pragma unroll
                for (int i = 0; i < 2; ++i) {   // 32-byte lane stride: lanes l and l+4 swap slots
                    const int x = (i + (lane >> 2)) & 1;
                    split4(x == 0 ? t0 : t1, hi, lo);
                    *reinterpret_cast<float4 *>(bh + 16 * x) = hi;
                    *reinterpret_cast<float4 *>(bl + 16 * x) = lo;
                }
            }
        };
        // register ring of PD tiles: tile kt lives in ring[kt % PD]; as soon as it has been written to
        // shared memory the slot is refilled with tile kt + PD, so PD - 1 .. PD tiles of global-memory
        // latency are in flight per thread (one tile was not enough: the staging warps sat on the
        // scoreboard of their own loads)
        ldt ring[PD][4];
#pragma unroll
        for (int d = 0; d < PD; ++d)
            if (d < nk) load_b(d, ring[d]);
        for (int kt0 = 0; kt0 < nk; kt0 += PD) {
#pragma unroll
            for (int d = 0; d < PD; ++d) {
                const int kt = kt0 + d;
                if (kt < nk) {
                    const int st = kt % NST, n = kt / NST;
                    if (n >= 1) mbar_wait(smem_u32(empty + st), (uint32_t)((n - 1) & 1));   // MMAs that read this stage are done
                    store_b(st, ring[d]);
                    if (kt + PD < nk) load_b(kt + PD, ring[d]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(full_b + st));
                    // the previous chunk is drained while the tensor core works on this one
                    if (kt % CH == 0 && kt > 0) drain(kt / CH - 1);
                }
            }
        }
        if constexpr (DIRECT) {   // single chunk: the epilogue reads the accumulator itself
            mbar_wait(smem_u32(chunk), 0u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        } else {
            drain((nk - 1) / CH);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
        const int gm = m0 + mt * TM + 32 * q + lane;
        const bool row_ok = gm < Co;
        const float sc = row_ok ? __ldg(scale + gm) : 0.f, sh = row_ok ? __ldg(shift + gm) : 0.f;
        float *orow = out + ((size_t)b * Co + (row_ok ? gm : 0)) * P;
        // optional epilogue extras (MlpEpilogue): a gathered addend in front of the affine -- the second half
        // of a concat-GEMM restructured as W1*x + (W2*p)[idx] -- and a channels-last store
        const float *add_row = ep.addend ? ep.addend + (size_t)b * ep.NA * Co + (row_ok ? gm : 0) : nullptr;
        const int *idx32 = (ep.addend && !ep.idx_is_i64) ? reinterpret_cast<const int *>(ep.idx) + (size_t)b * P : nullptr;
        const long long *idx64 = (ep.addend && ep.idx_is_i64) ? reinterpret_cast<const long long *>(ep.idx) + (size_t)b * P : nullptr;
        float *ocol = ep.out_nsc ? out + (size_t)b * P * Co + (row_ok ? gm : 0) : nullptr;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            uint32_t v[32];
            if constexpr (DIRECT) {
                tmem_load32(0, i, v);   // warp-collective: rows beyond Co take part, they just do not store
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(acc[mt][32 * i + j]);
            }
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int gn = n0 + COLS * cs + 32 * i + j;
                    float y[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) y[u] = __uint_as_float(v[j + u]);
                    if (add_row) {   // the column's index is the same for the whole warp (one broadcast load),
                                     // the addend row is channels-last: 32 lanes read 128 contiguous bytes
                        int a[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            a[u] = (gn + u < P) ? (idx32 ? __ldg(idx32 + gn + u) : (int)__ldg(idx64 + gn + u)) : 0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) y[u] = __fadd_rn(y[u], __ldg(add_row + (size_t)a[u] * Co));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        y[u] = __fadd_rn(__fmul_rn(y[u], sc), sh);
                        if (act == 1) y[u] = fmaxf(y[u], 0.f);
                        else if (act == 2) y[u] = (y[u] > 0.f) ? y[u] : __fmul_rn(y[u], slope);
                    }
                    if (ocol) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (gn + u < P) ocol[(size_t)(gn + u) * Co] = y[u];
                    } else if (vec && gn + 3 < P) {
                        *reinterpret_cast<float4 *>(orow + gn) = make_float4(y[0], y[1], y[2], y[3]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (gn + u < P) orow[gn + u] = y[u];
                    }
                }
            }
        }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(DIRECT ? TN : 2 * TN * MT));
}


// ===================================================================== weight gradient (training)
// dW[co, ci] = sum_b sum_p dZ[b, co, p] * X[b, ci, p],  X = [X1; X2]  (backward of the 1x1 conv w.r.t. its
// weight: a GEMM with M = Co, N = Ci and the long axis K = B * P).  Both operands are K-contiguous in memory
// (dZ and X are NCHW: the point axis is the fast one), so every thread loads 16 bytes along K and writes one
// whole (row; k..k+3) slot of the K-major UMMA layout -- no transposition.  Same numerics as the forward
// layer: 3xTF32 split of both operands while they are staged, K accumulated in chunks of 128 into two
// alternating TMEM accumulators that are drained into round-to-nearest fp32 register sums.  Split-K over
// the grid's z axis: every CTA owns a run of (frame, 32-point) k-tiles and adds its 128x128 partial to dW
// with fp32 atomics (dW is zero-filled by the entry point), like cuDNN's own wgrad: summation order is not
// deterministic, values agree to fp32 round-off.
constexpr int WG_SMEM = 2 * STAGE_BYTES + 64;

__global__ void __launch_bounds__(256, 1)
fusion_wgrad_kernel(const float *__restrict__ dz, const float *__restrict__ x1, int C1, const float *__restrict__ x2, int C2,
                    float *__restrict__ dw, int Co, int P, int tiles_per_frame, int total_tiles, int tiles_per_split)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE_BYTES);   // [0],[1]: stage free; [2],[3]: chunk done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * STAGE_BYTES + 48);
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int Ci = C1 + C2;
    const int kt_begin = blockIdx.z * tiles_per_split, kt_end = min(total_tiles, kt_begin + tiles_per_split);
    const int nk = kt_end - kt_begin;
    if (nk <= 0) return;   // CTA-uniform

    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(2 * TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        for (int i = 0; i < 4; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    const bool vec = ((P & 3) == 0) && ((reinterpret_cast<uintptr_t>(dz) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(x1) & 15) == 0) && (!x2 || (reinterpret_cast<uintptr_t>(x2) & 15) == 0);
    const int q = wid & 3, half = wid >> 2;   // this thread's accumulator row = 32q + lane, columns 64*half ..
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    auto drain = [&](int c) {
        const int buf = c & 1;
        mbar_wait(smem_u32(bars + 2 + buf), (uint32_t)((c >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * TN + 64 * half + 16 * i);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[16 * i + j] = __fadd_rn(acc[16 * i + j], __uint_as_float(v[j]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };
    // global -> registers: warp w stages K chunk w (4 points) of both operands; lanes take rows lane + 32*i
    auto load_rows = [&](const float *base, int rows, int row0, int p0, float4 (&r)[4]) {
        // base: [rows_total, P] of this frame; element (row, p0 + 4*wid .. +3)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gr = row0 + lane + 32 * i, gp = p0 + 4 * wid;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < rows && gp < P) {
                const float *src = base + (size_t)gr * P + gp;
                if (vec && gp + 3 < P) {
                    v = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    v.x = __ldg(src);
                    if (gp + 1 < P) v.y = __ldg(src + 1);
                    if (gp + 2 < P) v.z = __ldg(src + 2);
                    if (gp + 3 < P) v.w = __ldg(src + 3);
                }
            }
            r[i] = v;
        }
    };
    auto load_tile = [&](int kt, float4 (&ra)[4], float4 (&rb)[4]) {
        const int b = kt / tiles_per_frame, p0 = (kt % tiles_per_frame) * TK;
        load_rows(dz + (size_t)b * Co * P, Co, m0, p0, ra);
        // B rows = input channels n0 .. n0+127 of cat(x1, x2); a 128-row tile may straddle the concat
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gn = n0 + lane + 32 * i, gp = p0 + 4 * wid;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gn < Ci && gp < P) {
                const float *src = (gn < C1) ? x1 + ((size_t)b * C1 + gn) * P + gp : x2 + ((size_t)b * C2 + (gn - C1)) * P + gp;
                if (vec && gp + 3 < P) {
                    v = __ldg(reinterpret_cast<const float4 *>(src));
                } else {
                    v.x = __ldg(src);
                    if (gp + 1 < P) v.y = __ldg(src + 1);
                    if (gp + 2 < P) v.z = __ldg(src + 2);
                    if (gp + 3 < P) v.w = __ldg(src + 3);
                }
            }
            rb[i] = v;
        }
    };
    auto store_tile = [&](int st, const float4 (&ra)[4], const float4 (&rb)[4]) {
        unsigned char *sA_hi = smem + st * STAGE_BYTES, *sA_lo = sA_hi + TILE_BYTES;
        unsigned char *sB_hi = sA_lo + TILE_BYTES, *sB_lo = sB_hi + TILE_BYTES;
        float4 hi, lo;
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // a warp's 16-byte stores are contiguous: bank-conflict free
            const int m = lane + 32 * i;
            split4(ra[i], hi, lo);
            *reinterpret_cast<float4 *>(sA_hi + wid * CHUNK_BYTES + m * 16) = hi;
            *reinterpret_cast<float4 *>(sA_lo + wid * CHUNK_BYTES + m * 16) = lo;
            split4(rb[i], hi, lo);
            *reinterpret_cast<float4 *>(sB_hi + wid * CHUNK_BYTES + m * 16) = hi;
            *reinterpret_cast<float4 *>(sB_lo + wid * CHUNK_BYTES + m * 16) = lo;
        }
    };

    float4 ra[4], rb[4], na[4], nb[4];
    load_tile(kt_begin, ra, rb);
    for (int kt = 0; kt < nk; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < nk) load_tile(kt_begin + kt + 1, na, nb);
        if (kt >= 2) mbar_wait(smem_u32(bars + st), ((kt >> 1) - 1) & 1);   // MMAs that read this stage are done
        store_tile(st, ra, rb);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(smem + st * STAGE_BYTES);
            const uint64_t da_hi = umma_desc(a_hi), da_lo = umma_desc(a_hi + TILE_BYTES);
            const uint64_t db_hi = umma_desc(a_hi + 2 * TILE_BYTES), db_lo = umma_desc(a_hi + 3 * TILE_BYTES);
            const uint32_t d_buf = tmem_d + (uint32_t)(((kt / CH) & 1) * TN);
#pragma unroll
            for (int j = 0; j < TK / 8; ++j) {
                const uint64_t off = (uint64_t)(j * ((2 * CHUNK_BYTES) >> 4));
                umma_tf32(d_buf, da_hi + off, db_hi + off, (kt % CH != 0 || j > 0) ? 1u : 0u);
                umma_tf32(d_buf, da_lo + off, db_hi + off, 1u);
                umma_tf32(d_buf, da_hi + off, db_lo + off, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         :: "r"(smem_u32(bars + st)) : "memory");
            if (kt % CH == CH - 1 || kt == nk - 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             :: "r"(smem_u32(bars + 2 + ((kt / CH) & 1))) : "memory");
        }
        if (kt % CH == 0 && kt > 0) drain(kt / CH - 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ra[i] = na[i];
            rb[i] = nb[i];
        }
    }
    drain((nk - 1) / CH);
    {
        const int gm = m0 + 32 * q + lane;
        if (gm < Co) {
            float *row = dw + (size_t)gm * Ci;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int gn = n0 + 64 * half + j;
                if (gn < Ci) atomicAdd(row + gn, acc[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(2 * TN));
}

}  // namespace ffb6d

using namespace ffb6d;

static int mlp_check(const void *x1, int64_t C1, const void *x2, int64_t C2, const void *weight, const void *scale,
                     const void *shift, int64_t B, int64_t Co, int64_t P, int act, const void *out)
{
    FFB6D_CHECK_ARG(B >= 0 && C1 >= 1 && C2 >= 0 && Co >= 1 && P >= 0, "fusion_mlp_fwd: bad size");
    FFB6D_CHECK_ARG(act >= 0 && act <= 2, "fusion_mlp_fwd: act=%d (0 none, 1 ReLU, 2 LeakyReLU)", act);
    FFB6D_CHECK_ARG(B < 65536 && Co <= 65535ll * TM && P < (1ll << 31) && C1 + C2 < (1ll << 31),
                    "fusion_mlp_fwd: size too large");
    if (B == 0 || P == 0) return 1;   // nothing to do
    FFB6D_CHECK_ARG(x1 && weight && scale && shift && out && (C2 == 0 || x2), "fusion_mlp_fwd: null pointer");
    return FFB6D_OK;
}

extern "C" size_t ffb6d_fusion_mlp_pack_bytes(int64_t Co, int64_t Ci)
{
    if (Co < 1 || Ci < 1) return 0;
    return (size_t)ceil_div(Co, TM) * (size_t)ceil_div(Ci, TK) * PACK_BLOCK_BYTES;
}

extern "C" int ffb6d_fusion_mlp_pack(const float *weight, int64_t Co, int64_t Ci, void *packed, size_t packed_bytes,
                                     ffb6d_stream_t stream)
{
    FFB6D_CHECK_ARG(weight && packed && Co >= 1 && Ci >= 1 && Ci < (1ll << 31) && Co <= 65535ll * TM,
                    "fusion_mlp_pack: bad argument");
    const size_t need = ffb6d_fusion_mlp_pack_bytes(Co, Ci);
    if (packed_bytes < need) {
        set_error("fusion_mlp_pack: %zu bytes required, %zu given", need, packed_bytes);
        return FFB6D_ERR_WORKSPACE;
    }
    FFB6D_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "fusion_mlp_pack: packed must be 16-byte aligned");
    const int nk = (int)ceil_div(Ci, TK), nblocks = (int)(ceil_div(Co, TM) * nk);
    fusion_mlp_pack_kernel<<<(unsigned)ceil_div((int64_t)nblocks * 1024, 256), 256, 0, (cudaStream_t)stream>>>(
        weight, (int)Co, (int)Ci, nk, nblocks, (float4 *)packed);
    FFB6D_LAUNCH_OK("fusion_mlp_pack_kernel");
    return FFB6D_OK;
}

extern "C" int ffb6d_fusion_mlp_fwd_ex(const float *x1, int64_t C1, const float *x2, int64_t C2, const void *packed,
                                       const float *scale, const float *shift, int64_t B, int64_t Co, int64_t P,
                                       int act, float negative_slope, const float *addend, const void *add_idx,
                                       int add_idx_is_i64, int64_t NA, int out_layout, float *out, ffb6d_stream_t stream)
{
    const int rc = mlp_check(x1, C1, x2, C2, packed, scale, shift, B, Co, P, act, out);
    if (rc != FFB6D_OK) return rc > 0 ? FFB6D_OK : rc;
    FFB6D_CHECK_ARG(out_layout == FFB6D_LAYOUT_NCS || out_layout == FFB6D_LAYOUT_NSC, "fusion_mlp_fwd: unknown out_layout %d",
                    out_layout);
    FFB6D_CHECK_ARG((addend == nullptr) == (add_idx == nullptr), "fusion_mlp_fwd: addend and add_idx come together");
    FFB6D_CHECK_ARG(!addend || (NA >= 1 && NA < (1ll << 31)), "fusion_mlp_fwd: bad addend length");
    {
        auto k_big = fusion_mlp_packed_kernel<3, false, 16, 3>;
        auto k_big2 = fusion_mlp_packed_kernel<2, false, 16, 2, 2>;
        auto k_direct = fusion_mlp_packed_kernel<1, true, 8, 2>;
        FFB6D_OPTIN_SMEM(k_big, mlp2_smem(3));
        FFB6D_OPTIN_SMEM(k_big2, mlp2_smem(2, 2));
        FFB6D_OPTIN_SMEM(k_direct, mlp2_smem(1));
    }
    MlpEpilogue ep;
    ep.addend = addend;
    ep.idx = add_idx;
    ep.NA = (int)NA;
    ep.idx_is_i64 = add_idx_is_i64;
    ep.out_nsc = out_layout == FFB6D_LAYOUT_NSC;
    dim3 grid((unsigned)ceil_div(P, TN), (unsigned)ceil_div(Co, TM), (unsigned)B);
    if (ceil_div(C1 + C2, TK) <= CH && !env().mlp_no_direct)
        fusion_mlp_packed_kernel<1, true, 8, 2><<<grid, mlp2_threads(8), mlp2_smem(1), (cudaStream_t)stream>>>(
            x1, (int)C1, C2 ? x2 : nullptr, (int)C2, (const unsigned char *)packed, scale, shift, out, (int)Co, (int)P,
            act, negative_slope, ep);
    else if (Co > TM && !env().mlp_no_pair_tiles) {   // two row tiles per CTA: one staged activation tile feeds 256 output channels
        dim3 grid2(grid.x, (unsigned)ceil_div(ceil_div(Co, TM), 2), grid.z);
        fusion_mlp_packed_kernel<2, false, 16, 2, 2><<<grid2, mlp2_threads(16), mlp2_smem(2, 2), (cudaStream_t)stream>>>(
            x1, (int)C1, C2 ? x2 : nullptr, (int)C2, (const unsigned char *)packed, scale, shift, out, (int)Co, (int)P,
            act, negative_slope, ep);
    } else
        fusion_mlp_packed_kernel<3, false, 16, 3><<<grid, mlp2_threads(16), mlp2_smem(3), (cudaStream_t)stream>>>(
            x1, (int)C1, C2 ? x2 : nullptr, (int)C2, (const unsigned char *)packed, scale, shift, out, (int)Co, (int)P,
            act, negative_slope, ep);
    FFB6D_LAUNCH_OK("fusion_mlp_packed_kernel");
    return FFB6D_OK;
}

extern "C" int ffb6d_fusion_mlp_fwd_packed(const float *x1, int64_t C1, const float *x2, int64_t C2, const void *packed,
                                           const float *scale, const float *shift, int64_t B, int64_t Co, int64_t P,
                                           int act, float negative_slope, float *out, ffb6d_stream_t stream)
{
    return ffb6d_fusion_mlp_fwd_ex(x1, C1, x2, C2, packed, scale, shift, B, Co, P, act, negative_slope, nullptr, nullptr, 0,
                                   0, FFB6D_LAYOUT_NCS, out, stream);
}

extern "C" int ffb6d_fusion_mlp_fwd(const float *x1, int64_t C1, const float *x2, int64_t C2, const float *weight,
                                    const float *scale, const float *shift, int64_t B, int64_t Co, int64_t P,
                                    int act, float negative_slope, float *out, ffb6d_stream_t stream)
{
    const int rc = mlp_check(x1, C1, x2, C2, weight, scale, shift, B, Co, P, act, out);
    if (rc != FFB6D_OK) return rc > 0 ? FFB6D_OK : rc;
    // raw weights: split them into a stream-ordered scratch block first (callers that keep their
    // weights should pack once with ffb6d_fusion_mlp_pack and call ffb6d_fusion_mlp_fwd_packed)
    const size_t bytes = ffb6d_fusion_mlp_pack_bytes(Co, C1 + C2);
    static std::atomic<unsigned long long> pool_set_mask{0};
    const int dev_bit = current_device() & (kMaxDevices - 1);
    if (!((pool_set_mask.load(std::memory_order_relaxed) >> dev_bit) & 1ull)) {   // keep freed scratch blocks in the stream-ordered pool instead of returning them to the OS
        int dev = 0;
        cudaMemPool_t pool;
        unsigned long long keep = ~0ull;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pool_set_mask.fetch_or(1ull << dev_bit, std::memory_order_relaxed);
    }
    void *scratch = nullptr;
    FFB6D_CUDA(cudaMallocAsync(&scratch, bytes, (cudaStream_t)stream));
    int r = ffb6d_fusion_mlp_pack(weight, Co, C1 + C2, scratch, bytes, stream);
    if (r == FFB6D_OK)
        r = ffb6d_fusion_mlp_fwd_packed(x1, C1, x2, C2, scratch, scale, shift, B, Co, P, act, negative_slope, out, stream);
    cudaFreeAsync(scratch, (cudaStream_t)stream);
    return r;
}

extern "C" int ffb6d_fusion_mlp_wgrad(const float *grad_z, const float *x1, int64_t C1, const float *x2, int64_t C2,
                                      int64_t B, int64_t Co, int64_t P, float *grad_w, ffb6d_stream_t stream)
{
    FFB6D_CHECK_ARG(B >= 0 && C1 >= 1 && C2 >= 0 && Co >= 1 && P >= 0, "fusion_mlp_wgrad: bad size");
    FFB6D_CHECK_ARG(B < 65536 && Co <= 65535ll * TM && C1 + C2 <= 65535ll * TN && P < (1ll << 31),
                    "fusion_mlp_wgrad: size too large");
    FFB6D_CHECK_ARG(grad_w, "fusion_mlp_wgrad: null grad_w");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Ci = C1 + C2;
    FFB6D_CUDA(cudaMemsetAsync(grad_w, 0, (size_t)Co * Ci * sizeof(float), st));
    if (B == 0 || P == 0) return FFB6D_OK;
    FFB6D_CHECK_ARG(grad_z && x1 && (C2 == 0 || x2), "fusion_mlp_wgrad: null pointer");
    if (Co <= 64 && Ci <= 64)   // narrow layer: a 128x128 tensor-core tile would be mostly padding (train.cu)
        return wgrad_small_launch(grad_z, x1, C1, x2, C2, B, Co, P, grad_w, st);
    FFB6D_OPTIN_SMEM(fusion_wgrad_kernel, WG_SMEM);
    const int64_t tpf = ceil_div(P, TK), total = B * tpf;
    FFB6D_CHECK_ARG(total < (1ll << 31), "fusion_mlp_wgrad: B * P too large");
    const int64_t tiles = ceil_div(Ci, TN) * ceil_div(Co, TM);
    // split K so that the grid fills the SMs about twice, but keep at least 8 k-tiles (one chunk pair) per CTA
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div(2 * (int64_t)num_sms(), tiles), ceil_div(total, 8)));
    if (splits > 65535) splits = 65535;
    const int64_t per = ceil_div(total, splits);
    splits = ceil_div(total, per);
    dim3 grid((unsigned)ceil_div(Ci, TN), (unsigned)ceil_div(Co, TM), (unsigned)splits);
    fusion_wgrad_kernel<<<grid, 256, WG_SMEM, st>>>(grad_z, x1, (int)C1, C2 ? x2 : nullptr, (int)C2, grad_w, (int)Co, (int)P,
                                                    (int)tpf, (int)total, (int)per);
    FFB6D_LAUNCH_OK("fusion_wgrad_kernel");
    return FFB6D_OK;
}
