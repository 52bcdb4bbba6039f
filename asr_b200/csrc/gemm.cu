// asr_b200 -- TN GEMM on the 5th-gen tensor cores:  C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N])
//
// Both operands are K-major fp32 in HBM (activations [rows, features], weights [out, in] -- the
// layouts torch.nn.Linear / GRU / LSTM already use, asr_deepspeech/modules/blocks.py:76-78,
// deepspeech.py:105), consumed as TF32 by tcgen05.mma.kind::tf32 with fp32 accumulation in TMEM.
//
// Kernel shape (persistent, warp-specialised, one CTA per SM):
//   warp 0     : TMA producer  -- cp.async.bulk.tensor 128x32 (A) and BNx32 (B) fp32 boxes, 128B swizzle,
//                kStages-deep mbarrier ring
//   warp 1     : MMA issuer    -- one thread issues 4 x tcgen05.mma (K=8 each) per 32-wide K block into a
//                double-buffered TMEM accumulator (2 x BN columns)
//   warps 2..5 : epilogue      -- tcgen05.ld 32x32b, + bias / accumulate, vectorised stores; overlaps the
//                next tile's main loop through the tmem_full/tmem_empty barriers
// Out-of-range rows/columns/K are zero-filled by TMA and masked in the epilogue, so any M, N and
// any K (row strides must be multiples of 16 bytes) are accepted.
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace asrb {

unsigned g_debug_flags = 0;
int g_gemm_force_bn = 0, g_gemm_bn256_gain = 118, g_gemm_tma_store = 1, g_gemm_cta_limit = 0;

// ------------------------------------------------------------------------------------------------
// host: tensor map encoding
// ------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return ASRB_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return ASRB_ERR_ALIGNMENT;
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) {
            if (strides_bytes[i - 1] % 16 != 0) return ASRB_ERR_ALIGNMENT;
            gstr[i - 1] = strides_bytes[i - 1];
        }
    }
    if (swizzle128 && box[0] * esize != 128) return ASRB_ERR_BAD_ARG;
    CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ASRB_ERR_TENSORMAP;
}

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128, bool as_tf32) {
    // TFLOAT32 makes the TMA unit round fp32 -> tf32 (round-to-nearest) on the way into shared
    // memory, instead of the tensor core truncating the low 13 mantissa bits.
    return make_tmap(out, as_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rank,
                     dims, strides_bytes, box, swizzle128);
}
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rank, dims, strides_bytes, box, true);
}

// ------------------------------------------------------------------------------------------------
// device: tcgen05 GEMM
// ------------------------------------------------------------------------------------------------
constexpr int kBM = 128;
constexpr int kBK = 32;  // fp32 elements = 128 bytes = one swizzle row
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
    static constexpr int kStageBytesA = kBM * kBK * 4;
    static constexpr int kStageBytesB = BN * kBK * 4;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kStages = (BN <= 64) ? 8 : (BN <= 128 ? 6 : 4);
    static constexpr int kTmemCols = 2 * BN;  // double-buffered accumulator
    static constexpr int kStoreBytes = 4 * 2 * 4096;  // per epilogue warp: two 32x32 fp32 staging tiles for TMA stores
    static constexpr int kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// BF16 = false: fp32 operands in HBM, tf32 MMA (TMA rounds fp32 -> tf32 on the way into shared memory);
// BF16 = true : bf16 operands (64 elements per 128-byte stage row), kind::f16 MMA.  Accumulation and C are fp32.
template <int BN, bool BF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, float* __restrict__ C, int ldc,
                    const float* __restrict__ bias, int M, int N, int K, int flags, int tma_store) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kStages * Cfg::kStageBytesA;
    uint8_t* smem_c = smem + Cfg::kStages * Cfg::kStageBytes;          // 1024-aligned (stage sizes are multiples of 4 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + Cfg::kStoreBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + Cfg::kStages;
    uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = ceil_div(M, kBM), num_n = ceil_div(N, BN);
    const int num_tiles = num_m * num_n;
    constexpr int kKE = BF16 ? 2 * kBK : kBK;      // operand elements per 128-byte stage row
    const int num_kb = ceil_div(K, kKE);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (tma_store) tma_prefetch_desc(&tmC);
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the (warp-uniform) loop and ONE elected lane issues, so the loop state lives in uniform
        // registers -- which is what TMA / tcgen05 instructions take.
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / num_n) * kBM, n0 = (tile % num_n) * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                    tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tmA, &full_bar[stage], kb * kKE, m0);
                    tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tmB, &full_bar[stage], kb * kKE, n0);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        constexpr uint32_t idesc = umma_idesc(BF16 ? kFmtBF16 : kFmtTF32, kBM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after_sync();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * Cfg::kStageBytesB));
#pragma unroll
                    for (int k = 0; k < kBK / 8; ++k) {  // UMMA_K = 8 tf32 / 16 bf16: advance 32 B inside the swizzle row
                        if constexpr (BF16) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        else                umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quad = warp & 3;  // TMEM lane quarter this warp may access
        const bool accumulate = flags & ASRB_GEMM_ACCUMULATE;
        const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        if (tma_store) {
            /* Each lane owns one accumulator row, so direct stores touch 32 different rows per instruction (32 half-used
             * sectors; the LSU then bounds tiles with short K).  Instead the warp lays its 32x32 sub-tile out in shared
             * memory in the 128-byte-swizzle pattern (16-byte chunk j of row r at chunk j ^ (r & 7): conflict-free for
             * the row-per-lane writes) and one lane hands it to the TMA unit, which writes whole 128-byte rows, clips at
             * M and N, and in accumulate mode adds in L2 (cp.reduce .add.f32).  Two staging tiles per warp. */
            uint8_t* stg = smem_c + quad * 8192;
            int cbuf = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / num_n) * kBM, n0 = (tile % num_n) * BN;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after_sync();
                const int row0 = m0 + quad * 32;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int col0 = n0 + c * 32;
                    if (row0 >= M || col0 >= N) break;             // warp-uniform
                    float v[32];
                    tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + c * 32, v);
                    tmem_ld_wait();
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < N) v[j] += __ldg(bias + col0 + j);
                    }
                    if (lane == 0) bulk_wait_group_read<1>();      // the store that last read this staging tile is done
                    __syncwarp();
                    uint8_t* buf = stg + cbuf * 4096;
                    const uint32_t rowaddr = smem_u32(buf) + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_f32x4(rowaddr + ((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (accumulate) tma_reduce_add_2d(&tmC, buf, col0, row0);
                        else            tma_store_2d(&tmC, buf, col0, row0);
                        bulk_commit_group();
                    }
                    cbuf ^= 1;
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (lane == 0) bulk_wait_group<0>();
        } else {
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / num_n) * kBM, n0 = (tile % num_n) * BN;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after_sync();
            const int row = m0 + quad * 32 + lane;
            float* crow = C + (size_t)row * ldc;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BN + c * 32, v);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (row < M && col0 < N) {
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < N) v[j] += __ldg(bias + col0 + j);
                    }
                    if (vec_ok && col0 + 32 <= N) {
                        float4* dst = reinterpret_cast<float4*>(crow + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            if (accumulate) {
                                float4 p = dst[j];
                                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                            }
                            dst[j] = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < N) crow[col0 + j] = accumulate ? crow[col0 + j] + v[j] : v[j];
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

template <int BN, bool BF16>
static int launch_gemm_tc(const void* A, int lda, const void* B, int ldb, float* C, int ldc, const float* bias,
                          int M, int N, int K, int flags, asrb_stream_t stream) {
    using Cfg = GemmCfg<BN>;
    constexpr int ES = BF16 ? 2 : 4;
    CUtensorMap tmA, tmB;
    uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)lda * ES};
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)ldb * ES};
    uint32_t bA[2] = {128 / ES, kBM}, bB[2] = {128 / ES, (uint32_t)BN};
    int rc = BF16 ? make_tmap_bf16(&tmA, A, 2, dA, sA, bA) : make_tmap_f32(&tmA, A, 2, dA, sA, bA);
    if (rc) return rc;
    rc = BF16 ? make_tmap_bf16(&tmB, B, 2, dB, sB, bB) : make_tmap_f32(&tmB, B, 2, dB, sB, bB);
    if (rc) return rc;
    // TMA-store epilogue when C qualifies for a tensor map (16-byte aligned base and row stride)
    CUtensorMap tmC = tmA;
    int tma_store = 0;
    if (g_gemm_tma_store && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && M >= 32 && N >= 32) {
        uint64_t dC[2] = {(uint64_t)N, (uint64_t)M}, sC[1] = {(uint64_t)ldc * 4};
        uint32_t bC[2] = {32, 32};
        rc = make_tmap_f32(&tmC, C, 2, dC, sC, bC, true, false);
        if (rc) return rc;
        tma_store = 1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        ASRB_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tf32_kernel<BN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::kSmemBytes));
        attr_set = true;
    }
    const int tiles = ceil_div(M, kBM) * ceil_div(N, BN);
    int grid = tiles < kNumSMs ? tiles : kNumSMs;
    if (g_gemm_cta_limit > 0 && grid > g_gemm_cta_limit) grid = g_gemm_cta_limit;   // persistent tile loop: any grid works
    gemm_tn_tf32_kernel<BN, BF16><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmC, C, ldc, bias, M, N, K,
                                                                                    flags, tma_store);
    ASRB_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device: plain fp32 SIMT GEMM -- DEBUG cross-check only (asrb_set_debug_flags bit 0), never the
// product path.  64x64 tile, 16-wide K slab, 4x4 micro-tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_tn_simt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C,
                    int ldc, const float* __restrict__ bias, int M, int N, int K, int flags) {
    __shared__ float As[16][64 + 1];
    __shared__ float Bs[16][64 + 1];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, k = i & 15;
            As[k][r] = (m0 + r < M && k0 + k < K) ? A[(size_t)(m0 + r) * lda + k0 + k] : 0.f;
            Bs[k][r] = (n0 + r < N && k0 + k < K) ? B[(size_t)(n0 + r) * ldb + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
            if (r < M && c < N) {
                float v = acc[i][j] + (bias ? bias[c] : 0.f);
                float* p = C + (size_t)r * ldc + c;
                *p = (flags & ASRB_GEMM_ACCUMULATE) ? *p + v : v;
            }
        }
}

// ------------------------------------------------------------------------------------------------
// device: tiled transpose  out[c, r] = in[r, c]
// ------------------------------------------------------------------------------------------------
// 64 x 64 tiles, 256 threads, 16-byte loads along the input rows and 16-byte stores along the output rows when the
// strides and base pointers allow (VEC); tile edges fall back to scalars.
__device__ __forceinline__ void store4(float* dst, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(__nv_bfloat16* dst, float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&lo);
    u.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst) = u;
}
__device__ __forceinline__ void store1(float* dst, float a) { *dst = a; }
__device__ __forceinline__ void store1(__nv_bfloat16* dst, float a) { *dst = __float2bfloat16_rn(a); }

template <bool VEC, typename OutT>
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int rows, int cols, long long ld_in, OutT* __restrict__ out,
                 long long ld_out) {
    __shared__ float tile[64][65];
    const long long r0 = (long long)blockIdx.x * 64;   // rows on grid.x (up to 2^31 blocks)
    const int c0 = blockIdx.y * 64;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k, r = idx >> 4, c = (idx & 15) * 4;
        if (r0 + r < rows) {
            const float* src = in + (r0 + r) * ld_in + c0 + c;
            if (VEC && c0 + c + 3 < cols) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src));
                tile[r][c] = v.x; tile[r][c + 1] = v.y; tile[r][c + 2] = v.z; tile[r][c + 3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) tile[r][c + e] = (c0 + c + e < cols) ? src[e] : 0.f;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k, c = idx >> 4, r = (idx & 15) * 4;
        if (c0 + c < cols) {
            OutT* dst = out + (size_t)(c0 + c) * ld_out + r0 + r;
            if (VEC && r0 + r + 3 < rows) {
                store4(dst, tile[r][c], tile[r + 1][c], tile[r + 2][c], tile[r + 3][c]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (r0 + r + e < rows) store1(dst + e, tile[r + e][c]);
            }
        }
    }
}

// out = [A | A - tf32(A) | A]  (mode 0)   or   [B | B | B - tf32(B)]  (mode 1), row-major [rows, 3*cols]
// Feeding these to the TN GEMM gives the 3xTF32 product a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (~fp32 accuracy).
__global__ void split3_kernel(const float* __restrict__ in, long long rows, int cols, int ld_in, float* __restrict__ out,
                              int ld_out, int mode) {
    const long long n = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i % cols);
        const float v = in[r * ld_in + c];
        uint32_t u = __float_as_uint(v);
        // round-to-nearest tf32 (matches the TMA TFLOAT32 conversion closely enough: the residual is
        // exact whatever rounding the hardware applies to `hi`, because lo is itself re-rounded)
        const float hi = __uint_as_float((u + 0x1000u) & 0xFFFFE000u);
        const float lo = v - hi;
        float* o = out + r * ld_out;
        if (mode == 0) { o[c] = hi; o[cols + c] = lo; o[2 * cols + c] = hi; }
        else           { o[c] = hi; o[cols + c] = hi; o[2 * cols + c] = lo; }
    }
}

}  // namespace asrb

using namespace asrb;

extern "C" {

int asrb_version(void) { return 100; }

int asrb_set_debug_flags(unsigned flags) {
    g_debug_flags = flags;
    return 0;
}

const char* asrb_strerror(int code) {
    switch (code) {
        case 0: return "ok";
        case ASRB_ERR_BAD_ARG: return "asr_b200: bad argument";
        case ASRB_ERR_ALIGNMENT: return "asr_b200: pointer or stride not 16-byte aligned";
        case ASRB_ERR_UNSUPPORTED: return "asr_b200: unsupported shape for this kernel";
        case ASRB_ERR_WORKSPACE: return "asr_b200: workspace too small";
        case ASRB_ERR_DRIVER: return "asr_b200: cuTensorMapEncodeTiled not available from the driver";
        case ASRB_ERR_TENSORMAP: return "asr_b200: tensor map encoding failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "asr_b200: unknown error";
}

namespace asrb {
// 128 x 256 tiles read 1.5x fewer operand bytes from shared memory per FLOP than 128 x 128 (the SMEM port, not the tensor
// pipe, is the limit: 8 KB per 64-cycle MMA at N=128), but they quantise worse; pick the tile with the smaller estimated
// time = waves x tile width / efficiency.
static int gemm_pick_bn(int M, int N) {
    if (N <= 64) return 64;
    int bn = 128;
    if (N > 128) {
        const long long w128 = ceil_div64((long long)ceil_div(M, kBM) * ceil_div(N, 128), kNumSMs);
        const long long w256 = ceil_div64((long long)ceil_div(M, kBM) * ceil_div(N, 256), kNumSMs);
        if (w256 * 256 * 100 < w128 * 128 * g_gemm_bn256_gain) bn = 256;
    }
    if (g_gemm_force_bn) bn = g_gemm_force_bn;
    return bn;
}
}  // namespace asrb

int asrb_gemm_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias, int M, int N,
                 int K, int flags, asrb_stream_t stream) {
    ASRB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(lda >= K && ldb >= K && ldc >= N, ASRB_ERR_BAD_ARG);
    if (g_debug_flags & ASRB_DEBUG_SIMT_GEMM) {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
        gemm_tn_simt_kernel<<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags);
        ASRB_LAUNCH_OK();
        return 0;
    }
    ASRB_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, ASRB_ERR_ALIGNMENT);
    const int bn = gemm_pick_bn(M, N);
    if (bn == 64) return launch_gemm_tc<64, false>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags, stream);
    if (bn == 256) return launch_gemm_tc<256, false>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags, stream);
    return launch_gemm_tc<128, false>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags, stream);
}

/* Same product with bf16 operands (A [M, lda], B [N, ldb] bf16, lda/ldb multiples of 8), fp32 accumulate and C. */
int asrb_gemm_tn_bf16(const void* A, int lda, const void* B, int ldb, float* C, int ldc, const float* bias, int M, int N,
                      int K, int flags, asrb_stream_t stream) {
    ASRB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(lda >= K && ldb >= K && ldc >= N, ASRB_ERR_BAD_ARG);
    ASRB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, ASRB_ERR_ALIGNMENT);
    const int bn = gemm_pick_bn(M, N);
    if (bn == 64) return launch_gemm_tc<64, true>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags, stream);
    if (bn == 256) return launch_gemm_tc<256, true>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags, stream);
    return launch_gemm_tc<128, true>(A, lda, B, ldb, C, ldc, bias, M, N, K, flags, stream);
}

/* DEBUG / tuning: force the N tile (0 = automatic, 128, 256); gain = assumed speed of the 256-wide tile relative to
 * the 128-wide one in percent (default 118: measured 1.10-1.18 on the configs[1] shapes) */
/* Caps the number of persistent CTAs of the GEMM launches that follow (0 = no cap; returns the old value): work issued
 * on a second stream next to a kernel that needs its own SMs (the recurrent chain) must not take all of them. */
int asrb_gemm_cta_limit(int n) {
    const int old = g_gemm_cta_limit;
    if (n >= 0) g_gemm_cta_limit = n;
    return old;
}

/* debug/tuning: 1 (default) epilogue through shared memory + TMA stores, 0 direct row-per-lane stores; v < 0 queries */
int asrb_debug_gemm_tma_store(int v) {
    const int old = g_gemm_tma_store;
    if (v >= 0) g_gemm_tma_store = v ? 1 : 0;
    return old;
}

int asrb_debug_gemm_tile(int force_bn, int gain_pct) {
    ASRB_REQUIRE(force_bn == 0 || force_bn == 128 || force_bn == 256, ASRB_ERR_BAD_ARG);
    g_gemm_force_bn = force_bn;
    if (gain_pct > 0) g_gemm_bn256_gain = gain_pct;
    return 0;
}

int asrb_transpose(const float* in, long long rows, int cols, int ld_in, float* out, int ld_out, asrb_stream_t stream) {
    ASRB_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, ASRB_ERR_BAD_ARG);
    dim3 grid((unsigned)ceil_div64(rows, 64), ceil_div(cols, 64));
    ASRB_REQUIRE(grid.y <= 65535u && rows < (1LL << 31), ASRB_ERR_UNSUPPORTED);
    const bool vec = ld_in % 4 == 0 && ld_out % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec) transpose_kernel<true, float><<<grid, 256, 0, stream>>>(in, (int)rows, cols, ld_in, out, ld_out);
    else     transpose_kernel<false, float><<<grid, 256, 0, stream>>>(in, (int)rows, cols, ld_in, out, ld_out);
    ASRB_LAUNCH_OK();
    return 0;
}

/* out[c, r] = bf16(in[r, c]) : the bf16 K-major operands of the backward GEMMs (ld_out multiple of 8) */
int asrb_transpose_bf16(const float* in, long long rows, int cols, int ld_in, void* out, int ld_out, asrb_stream_t stream) {
    ASRB_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, ASRB_ERR_BAD_ARG);
    dim3 grid((unsigned)ceil_div64(rows, 64), ceil_div(cols, 64));
    ASRB_REQUIRE(grid.y <= 65535u && rows < (1LL << 31), ASRB_ERR_UNSUPPORTED);
    const bool vec = ld_in % 4 == 0 && ld_out % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (vec) transpose_kernel<true, __nv_bfloat16><<<grid, 256, 0, stream>>>(in, (int)rows, cols, ld_in, o, ld_out);
    else     transpose_kernel<false, __nv_bfloat16><<<grid, 256, 0, stream>>>(in, (int)rows, cols, ld_in, o, ld_out);
    ASRB_LAUNCH_OK();
    return 0;
}

int asrb_split3(const float* in, long long rows, int cols, int ld_in, float* out, int ld_out, int mode,
                asrb_stream_t stream) {
    ASRB_REQUIRE(in && out && rows > 0 && cols > 0 && ld_out >= 3 * cols && (mode == 0 || mode == 1), ASRB_ERR_BAD_ARG);
    long long n = rows * cols;
    int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    split3_kernel<<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, out, ld_out, mode);
    ASRB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
