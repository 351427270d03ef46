// michigan_b200 — weight-gradient implicit GEMM on tcgen05 (sm_100a).
//
//   dW[co, tap, ci] = sum_{pixels} dY[pix, co] * X[pix (+) tap, ci]
//
// GEMM view per filter tap: D[M = 128 output channels, N = BN input channels] accumulated over
// K = pixels.  Both operands are NHWC activations, i.e. contiguous along their M/N index and strided
// along K, so they are fed to tcgen05.mma as MN-major operands: every TMA box [32 channels x 64
// pixels] lands as eight 1024 B swizzle atoms (32 ch x 8 pixels); M and N span several boxes (LBO).
// The tap shift and the conv zero padding come from the box start coordinate + TMA OOB zero fill,
// exactly as in the forward kernel.  Pixels are split across CTAs (split-K); partial tiles are
// reduced into the packed gradient with vector fp32 atomics.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "mg_ptx.cuh"
#include "mg_internal.h"

namespace mg {

// warp 0: TMA producer, warp 1: MMA issuer 0, warps 2-5: epilogue, warp 6: MMA issuer 1 (filter taps are split between the
// two issuers when KW >= 2: one thread cannot issue more than one small-N tcgen05.mma per ~100-120 cycles)
constexpr int kWgThreads = 224;
constexpr int kWgStagesMax = 6;

struct WgradParams {
    int N, OH, OW, Cout, Cin, KH, KW, stride, pad;
    int TW, TH, TN, tiles_w, tiles_h, tiles_n, pix_tiles;
    int BN, m_tiles, n_tiles, splits, stages;
    int issuers;               // 1 or 2 MMA-issuing threads (taps kw = j, j + issuers, ... belong to issuer j)
    int pix_tile, box_bytes;   // K (pixels) per pipeline stage: 32 or 64; bytes of one [pix_tile x 128 B] box
    int f16;                   // 0: fp32 storage read as TF32 (32 channels per 128 B row, K = 8 pixels per MMA);
                               // 2: bf16 operands (64 channels per row, K = 16 pixels per MMA)
    int kelem, mboxes, kmma;   // channels per box, boxes per 128-row M tile, pixels per MMA
    int halo;                  // bf16, stride 1, KW >= 2: ONE input patch [(TW + KW - 1) x TH pixels] per stage instead of KW shifted
                               // tiles; tap kw reads it through an operand descriptor that starts kw pixel rows (128 B) later
    int xbox_bytes, x_tx_bytes;// smem slot of one input box (= box_bytes unless halo) and the bytes TMA writes into it
    uint32_t idesc, tmem_cols;
    float* dw;  // [Cout][KH*KW*Cin]
};

// MN-major TF32 operands must use the "128B swizzle with 32B atoms" layout (UMMA layout type 1,
// TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 32 contiguous fp32 (128 B), 4 rows (K) per
// 512 B swizzle atom.  LBO = byte distance between consecutive 32-element blocks along M/N,
// SBO = distance between 4-row K groups (512 B for a dense box); one K=8 MMA spans two groups.
// 16-bit MN-major operands use the plain 128B swizzle: rows of 64 contiguous elements (128 B), 8 rows (K) per 1024 B atom,
// SBO = 1024 B between 8-row K groups (one K = 16 MMA spans two), LBO = distance between 64-element blocks along M/N.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128_16(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
    return d;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // One unit = one filter ROW (kh): the dY tile is loaded once per stage and multiplied with the KW shifted
    // input tiles (accumulators kw*BN.. in TMEM), so dY is streamed KH (not KH*KW) times from HBM/L2.
    const int kBoxBytes = p.box_bytes;
    const int a_bytes = p.mboxes * kBoxBytes;           // M = 128 -> 4 boxes of 32 fp32 channels / 2 boxes of 64 bf16 channels
    const int b1_bytes = (p.BN / p.kelem) * p.xbox_bytes;  // one tap's input tile (halo: the shared patch)
    const int stage_bytes = a_bytes + (p.halo ? 1 : p.KW) * b1_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgStagesMax;
    uint64_t* done_bar = bars + 2 * kWgStagesMax;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStagesMax + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // unit decode: blockIdx.x = split * units + (kh * m_tiles + mt) * n_tiles + nt
    // units are the FAST index: the CTAs of one wave work on the same pixel range, so dY / X come from DRAM once per wave
    const int n_units = p.KH * p.m_tiles * p.n_tiles;
    int u = blockIdx.x % n_units;
    const int split = blockIdx.x / n_units;
    const int nt = u % p.n_tiles; u /= p.n_tiles;
    const int mt = u % p.m_tiles;
    const int kh = u / p.m_tiles;
    const int t_begin = (int)((long long)p.pix_tiles * split / p.splits);
    const int t_end = (int)((long long)p.pix_tiles * (split + 1) / p.splits);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], p.issuers); }
        mbar_init(done_bar, p.issuers);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // TMA producer WARP: a stage is 4 + KW*(BN/32) boxes of [pix_tile x 32 channels] (MN-major operands cannot use wider
        // boxes: one 128 B swizzle row = 32 fp32), up to 16 bulk-tensor instructions.  Issued by one thread they cost more than
        // the stage's MMAs; here lane j issues box j, so a stage goes out in one pass.
        int st = 0; uint32_t ph = 0;
        const int nb_x = p.BN / p.kelem;
        // bytes the TMA unit delivers per stage (a halo patch box may be shorter than its atom-aligned slot)
        const uint32_t tx = (uint32_t)(a_bytes + (p.halo ? nb_x * p.x_tx_bytes : p.KW * b1_bytes));
        const int n_boxes = p.mboxes + (p.halo ? 1 : p.KW) * nb_x;
        for (int t = t_begin; t < t_end; ++t) {
            const int tw = t % p.tiles_w;
            const int th = (t / p.tiles_w) % p.tiles_h;
            const int tn = t / (p.tiles_w * p.tiles_h);
            const int ow0 = tw * p.TW, oh0 = th * p.TH, n0 = tn * p.TN;
            if (lane == 0) {
                mbar_wait(&empty_bar[st], ph ^ 1);
                mbar_arrive_expect_tx(&full_bar[st], tx);
            }
            __syncwarp();
            uint8_t* sa = smem + (size_t)st * stage_bytes;
            for (int b = lane; b < n_boxes; b += 32) {
                if (b < p.mboxes) {
                    tma_load_4d(sa + b * kBoxBytes, &tmDY, &full_bar[st], mt * 128 + b * p.kelem, ow0, oh0, n0);
                } else {
                    const int kw = (b - p.mboxes) / nb_x, j = (b - p.mboxes) - kw * nb_x;
                    tma_load_4d(sa + a_bytes + kw * b1_bytes + j * p.xbox_bytes, &tmX, &full_bar[st], nt * p.BN + j * p.kelem,
                                ow0 * p.stride - p.pad + kw, oh0 * p.stride - p.pad + kh, n0);
                }
            }
            __syncwarp();
            if (++st == p.stages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1 || warp == 6) {
        const int issuer = warp == 6 ? 1 : 0;
        if (lane == 0 && issuer < p.issuers) {
            int st = 0; uint32_t ph = 0;
            uint32_t first = 1;
            for (int t = t_begin; t < t_end; ++t) {
                mbar_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
                for (int kw = issuer; kw < p.KW; kw += p.issuers)
                    for (int k = 0; k < p.pix_tile / p.kmma; ++k) {
                        const uint32_t koff = (uint32_t)(k * p.kmma * 128);      // kmma pixel rows of 128 B
                        if (p.halo) {
                            // K step k = tile row k (TW = 16 pixels = two contiguous 8-row swizzle groups); the input row starts
                            // kw pixels into row k of the (TW + KW - 1)-wide patch
                            const uint64_t da = umma_desc_mnmajor_sw128_16(sa + koff, kBoxBytes);
                            const uint64_t db = umma_desc_mnmajor_sw128_16(sa + a_bytes + (uint32_t)((k * (p.TW + p.KW - 1) + kw) * 128), p.xbox_bytes);
                            umma_f16(tmem_base + (uint32_t)(kw * p.BN), da, db, p.idesc, (first && k == 0) ? 0u : 1u);
                        } else if (p.f16) {
                            const uint64_t da = umma_desc_mnmajor_sw128_16(sa + koff, kBoxBytes);
                            const uint64_t db = umma_desc_mnmajor_sw128_16(sa + a_bytes + kw * b1_bytes + koff, kBoxBytes);
                            umma_f16(tmem_base + (uint32_t)(kw * p.BN), da, db, p.idesc, (first && k == 0) ? 0u : 1u);
                        } else {
                            const uint64_t da = umma_desc_mnmajor_sw128(sa + koff, kBoxBytes);
                            const uint64_t db = umma_desc_mnmajor_sw128(sa + a_bytes + kw * b1_bytes + koff, kBoxBytes);
                            umma_tf32(tmem_base + (uint32_t)(kw * p.BN), da, db, p.idesc, (first && k == 0) ? 0u : 1u);
                        }
                    }
                first = 0;
                umma_commit(&empty_bar[st]);
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
            umma_commit(done_bar);
        }
    } else if (warp >= 2 && warp <= 5) {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;  // output channel within the M tile
        const int co = mt * 128 + row;
        if (t_end > t_begin) {
            mbar_wait(done_bar, 0);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
            for (int kw = 0; kw < p.KW; ++kw) {
                float* dst = p.dw + (size_t)co * (p.KH * p.KW * p.Cin) + (size_t)(kh * p.KW + kw) * p.Cin + nt * p.BN;
                for (int j0 = 0; j0 < p.BN; j0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(t_row + kw * p.BN + j0, v);
                    tmem_ld_wait();
                    if (co < p.Cout) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            float4 val = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                                                     __uint_as_float(v[i + 3]));
                            if (p.splits > 1) atomicAdd(reinterpret_cast<float4*>(dst + j0 + i), val);
                            else *reinterpret_cast<float4*>(dst + j0 + i) = val;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

static int np2(int v) { int r = 1; while (r < v) r <<= 1; return r; }

}  // namespace mg

using namespace mg;

// dw: [Cout][KH*KW*Cin] fp32 (tap-major K, the layout of mg_pack_weight); zeroed here when split-K > 1.
// fmt 0: dy / x fp32 (read as TF32); fmt 2: dy / x bf16.
static int wgrad_launch(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                        int KH, int KW, int stride, int pad, int fmt, cudaStream_t stream) {
    if (!dy || !x || !dw) return set_error(-1, "mg_conv_wgrad: null pointer");
    if (fmt != 0 && fmt != 2) return set_error(-5, "mg_conv_wgrad: operand format must be 0 (tf32) or 2 (bf16)");
    const int kelem = fmt ? 64 : 32;
    if (Cin % kelem != 0 || Cout % kelem != 0) return set_error(-2, "mg_conv_wgrad: channels must be multiples of %d", kelem);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.OH = OH; p.OW = OW; p.Cout = Cout; p.Cin = Cin; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
    p.f16 = fmt; p.kelem = kelem; p.mboxes = 128 / kelem; p.kmma = fmt ? 16 : 8;
    int BN = Cin >= 128 ? 128 : Cin;
    if (Cin % BN != 0) BN = kelem;
    while (KW * BN > 512) BN /= 2;                       // KW accumulators of BN columns must fit TMEM
    if (BN < kelem || Cin % BN != 0) return set_error(-3, "mg_conv_wgrad: KW %d x Cin %d does not fit TMEM", KW, Cin);
    p.BN = BN;
    // K (pixels) per stage: 64 when at least 3 stages fit in shared memory, else 32
    int pix_tile = 64;
    if ((200 * 1024) / ((p.mboxes + KW * (BN / kelem)) * 64 * 128) < 3) pix_tile = 32;
    // halo schedule: tile 16 x 4 pixels (one K = 16 MMA per tile row), input patch (16 + KW - 1) x 4 pixels
    p.halo = (fmt == 2 && stride == 1 && KW >= 2 && KW <= 4 && OW >= 16 && OH >= 4 && tune(TK_WGRAD_HALO)) ? 1 : 0;
    if (p.halo) pix_tile = 64;
    p.pix_tile = pix_tile; p.box_bytes = pix_tile * 128;
    p.TW = p.halo ? 16 : (np2(OW) < 8 ? np2(OW) : 8);
    int th = pix_tile / p.TW;
    p.TH = np2(OH) < th ? np2(OH) : th;
    p.TN = pix_tile / (p.TW * p.TH);
    p.tiles_w = (OW + p.TW - 1) / p.TW; p.tiles_h = (OH + p.TH - 1) / p.TH; p.tiles_n = (N + p.TN - 1) / p.TN;
    p.pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.m_tiles = (Cout + 127) / 128;
    p.n_tiles = Cin / BN;
    const int dual_env = tune(TK_WGRAD_DUAL);
    p.issuers = (KW >= 2 && dual_env) ? 2 : 1;
    const int units = KH * p.m_tiles * p.n_tiles;
    // split-K so that units * splits fills whole waves of one CTA per SM: round DOWN (296 / 6 units = 49 -> 294 CTAs = 2 waves;
    // rounding up gave 300 CTAs = a third wave with 4 CTAs)
    int splits = (2 * num_sms()) / units;
    if (splits < 1) splits = 1;
    if (splits > p.pix_tiles) splits = p.pix_tiles;
    p.splits = splits;
    // halo: the patch box is rounded up to whole 1024 B swizzle atoms so that every box starts on an atom boundary
    p.x_tx_bytes = p.halo ? (p.TW + KW - 1) * p.TH * p.TN * 128 : p.box_bytes;
    p.xbox_bytes = (p.x_tx_bytes + 1023) / 1024 * 1024;
    const int stage_bytes = p.mboxes * p.box_bytes + (p.halo ? 1 : KW) * (BN / kelem) * p.xbox_bytes;
    int stages = (200 * 1024) / stage_bytes;
    if (stages < 1) return set_error(-4, "mg_conv_wgrad: stage of %d bytes does not fit shared memory", stage_bytes);
    if (stages > kWgStagesMax) stages = kWgStagesMax;
    p.stages = stages;
    // A and B both MN-major (bits 15, 16), M = 128, N = BN, fp32 accumulate
    p.idesc = (fmt ? umma_idesc_16(128, BN, 2) : umma_idesc_tf32(128, BN)) | (1u << 15) | (1u << 16);
    int tc = np2(KW * BN); p.tmem_cols = tc < 32 ? 32 : tc;
    p.dw = dw;

    CUtensorMap tmDY, tmX;
    const int esz = fmt ? 2 : 4;
    const CUtensorMapDataType dt = fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle sw = fmt ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cout * esz, (cuuint64_t)OW * Cout * esz, (cuuint64_t)OH * OW * Cout * esz};
        cuuint32_t box[4] = {(cuuint32_t)kelem, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
        cuuint32_t es[4] = {1, 1, 1, 1};
        int rc = encode_tensor_map(&tmDY, (void*)dy, dt, 4, dims, strides, box, es, sw);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * esz, (cuuint64_t)W * Cin * esz, (cuuint64_t)H * W * Cin * esz};
        cuuint32_t box[4] = {(cuuint32_t)kelem, (cuuint32_t)(p.halo ? p.TW + KW - 1 : p.TW * stride), (cuuint32_t)(p.TH * stride), (cuuint32_t)p.TN};
        cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        int rc = encode_tensor_map(&tmX, (void*)x, dt, 4, dims, strides, box, es, sw);
        if (rc) return rc;
    }
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)Cout * KH * KW * Cin * sizeof(float), stream);
        if (e != cudaSuccess) return set_error((int)e, "wgrad memset: %s", cudaGetErrorString(e));
    }
    static thread_local int attr_dev = -1;
    int dev = 0; cudaGetDevice(&dev);
    if (attr_dev != dev) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error((int)e, "wgrad attr: %s", cudaGetErrorString(e));
        attr_dev = dev;
    }
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 256;
    wgrad_tf32_kernel<<<units * splits, kWgThreads, smem_bytes, stream>>>(tmDY, tmX, p);
    return check_launch("mg_conv_wgrad");
}

extern "C" int mg_conv_wgrad(const float* dy, const float* x, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                             int KH, int KW, int stride, int pad, void* stream_) {
    return wgrad_launch(dy, x, dw, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, 0, reinterpret_cast<cudaStream_t>(stream_));
}

// Same with 16-bit (bf16) operands: dy [N,OH,OW,Cout] and x [N,H,W,Cin] bf16, fp32 accumulation and output.
extern "C" int mg_conv_wgrad16(const void* dy16, const void* x16, float* dw, int N, int H, int W, int Cin, int OH, int OW, int Cout,
                               int KH, int KW, int stride, int pad, void* stream_) {
    return wgrad_launch(dy16, x16, dw, N, H, W, Cin, OH, OW, Cout, KH, KW, stride, pad, 2, reinterpret_cast<cudaStream_t>(stream_));
}
