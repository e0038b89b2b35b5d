// Regional pooling: every (image, channel) plane of an NCHW feature map is read from HBM once and
// pooled over R rectangular regions (GeM / MAC / SPoC) in the same pass.
//
// Replaces the region loop of Rpool.roipool (cirtorch/modules/pools.py:126-167: rpool applied to
// x.narrow(2, i, wl).narrow(3, j, wl) for every region of the L-level grid) and the two max-pools
// of RMAC.forward (pools.py:64-113), which read the map once per region (1 + 1 + 4 + 9 = 15 times
// at L = 3) through non-contiguous views.
//
// One warp per plane: the plane is loaded with coalesced 128-bit loads (8 in flight per lane),
// transformed once (max(x, eps)^p for GeM) and kept in shared memory; each region is then
// reduced from shared memory with lanes = columns (column sums over the region's rows, one
// masked warp reduction per region; regions of one row band share the column sums).  Output is
// [N][R][C] (row = one region descriptor, C contiguous), which is what the L2N / whitening
// kernels take next.  Algorithmic bytes: N*C*H*W*4 read + N*R*C*4 written.
#include "common.cuh"

#include <stdlib.h>

namespace cir {

constexpr int REGION_MAX = 64;
constexpr int REGION_WARPS = 8;

struct RegionBox { short i0, j0, h, w; };

struct RegionParams {
    const float* x;
    float* out;
    const float* p;
    int p_stride;
    float eps;
    int pool_mode;
    int C, H, W, R;
    long long planes;
    RegionBox box[REGION_MAX];
};

__device__ __forceinline__ float gem_pow(float t, float p, int ip) {
    switch (ip) {
        case 1: return t;
        case 2: return t * t;
        case 3: return t * t * t;
        case 4: { const float t2 = t * t; return t2 * t2; }
        default: return powf(t, p);
    }
}

__global__ void __launch_bounds__(REGION_WARPS * 32)
region_pool_kernel(const __grid_constant__ RegionParams P) {
    extern __shared__ __align__(16) float region_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = P.H * P.W, W = P.W;
    float* s = region_smem + (size_t)warp * ((HW + 3) & ~3);
    const long long plane = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (plane >= P.planes) return;                      // warps are independent
    const int c = (int)(plane % P.C);
    const long long n = plane / P.C;
    const bool gem = P.pool_mode == CIR_POOL_GEM;
    const float pr = gem ? __ldg(P.p + (size_t)c * P.p_stride) : 1.0f;
    const int ip = (pr == 1.0f) ? 1 : (pr == 2.0f) ? 2 : (pr == 3.0f) ? 3 : (pr == 4.0f) ? 4 : 0;
    const float* src = P.x + plane * HW;
    // ---- the plane: HBM -> (transform) -> shared memory, 8 x 128-bit loads in flight per lane
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        for (int t0 = 0; t0 < HW; t0 += 8 * 128) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * 128 + lane * 4;
                if (t < HW) v[u] = ld_stream_f4(reinterpret_cast<const float4*>(src + t));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * 128 + lane * 4;
                if (t < HW) {
                    if (gem) {
                        v[u].x = gem_pow(fmaxf(v[u].x, P.eps), pr, ip); v[u].y = gem_pow(fmaxf(v[u].y, P.eps), pr, ip);
                        v[u].z = gem_pow(fmaxf(v[u].z, P.eps), pr, ip); v[u].w = gem_pow(fmaxf(v[u].w, P.eps), pr, ip);
                    }
                    *reinterpret_cast<float4*>(s + t) = v[u];
                }
            }
        }
    } else {
        for (int t = lane; t < HW; t += 32) {
            const float v = ld_stream_f1(src + t);
            s[t] = gem ? gem_pow(fmaxf(v, P.eps), pr, ip) : v;
        }
    }
    __syncwarp();
    // ---- the regions.  Lanes = columns: a lane reduces its column over the rows of the region (conflict-free shared
    // memory reads, all lanes busy), then one masked warp reduction per region.  Consecutive regions of the same row
    // band (the grid emits them row by row) reuse the column values.  Lane r keeps the result of region r (mod 32).
    const bool is_max = P.pool_mode == CIR_POOL_MAC;
    const float ident = is_max ? -INFINITY : 0.0f;
    float colv = ident, mine = 0.0f;
    int cur_i0 = -1, cur_h = -1;
    for (int r = 0; r < P.R; ++r) {
        const RegionBox b = P.box[r];
        float v = ident;
        if (W <= 32) {
            if (b.i0 != cur_i0 || b.h != cur_h) {
                colv = ident;
                if (lane < W) {
                    const float* col = s + b.i0 * W + lane;
                    int rr = 0;
                    for (; rr + 4 <= b.h; rr += 4) {
                        const float a0 = col[rr * W], a1 = col[(rr + 1) * W], a2 = col[(rr + 2) * W], a3 = col[(rr + 3) * W];
                        colv = is_max ? fmaxf(fmaxf(colv, a0), fmaxf(fmaxf(a1, a2), a3)) : colv + ((a0 + a1) + (a2 + a3));
                    }
                    for (; rr < b.h; ++rr) colv = is_max ? fmaxf(colv, col[rr * W]) : colv + col[rr * W];
                }
                cur_i0 = b.i0; cur_h = b.h;
            }
            v = (lane >= b.j0 && lane < b.j0 + b.w) ? colv : ident;
        } else {
            for (int cc = b.j0 + lane; cc < b.j0 + b.w; cc += 32) {
                const float* col = s + b.i0 * W + cc;
                for (int rr = 0; rr < b.h; ++rr) v = is_max ? fmaxf(v, col[rr * W]) : v + col[rr * W];
            }
        }
        const float acc = is_max ? warp_max(v) : warp_sum(v);
        if ((r & 31) == lane) mine = is_max ? acc : acc / (float)(b.h * b.w);
        if ((r & 31) == 31 || r == P.R - 1) {             // flush up to 32 results: finalise in parallel, one store each
            const int rbase = r & ~31;
            if (rbase + lane <= r) {
                float y = mine;
                if (gem && ip != 1) y = powf(y, 1.0f / pr);
                P.out[((size_t)n * P.R + rbase + lane) * P.C + c] = y;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Sum-type pooling (GeM, SPoC) over regions of a map at most 32 columns wide: prefix sums instead of one pass per band.
// Lanes = columns.  A lane walks down its column ONCE, keeping the running sum P of f(x) = max(x, eps)^p, and parks P in
// shared memory whenever it crosses one of the (<= 16) distinct row boundaries of the region list; a region is then
// P[bottom] - P[top] per column, masked to its columns, and one warp reduction.  ~550 warp instructions per 32 x 32
// plane with 15 regions instead of ~1,400 (transform to shared memory + one column pass per row band), no staging of the
// plane, 16 coalesced row loads in flight per lane.
// ---------------------------------------------------------------------------------------
constexpr int REGION_MAX_BND = 16;

struct RegionPrefixParams {
    const float* x;
    float* out;
    const float* p;
    int p_stride;
    float eps;
    int pool_mode;
    int C, H, W, R, nb;
    long long planes;
    short bnd[REGION_MAX_BND];            // sorted distinct row boundaries (region tops and bottoms)
    unsigned char top[REGION_MAX], bot[REGION_MAX];      // indices into bnd
    short j0[REGION_MAX], w[REGION_MAX];
    float inv_area[REGION_MAX];
};

// t^p for t >= eps > 0 as ex2(p * lg2 t) on the MUFU pipe (two instructions, ~1e-6 relative: the same as the tail kernel)
__device__ __forceinline__ float approx_pow(float t, float p) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(t));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p * l));
    return r;
}

// one lane = one column: running sum of f(x) down the column, parked in sv[k][lane] at the k-th row boundary.
// ip: 1..4 = x^ip by multiplication, 0 = general exponent, 5 = no transform (SPoC).  ONE instantiation with the class decided
// per element by selects: dispatching to per-class instantiations outside the loop took the kernel from 40 to 114-128
// registers (2 blocks per SM instead of 6) and from 262 to 820 us; as a noinline function 382 us.
__device__ __forceinline__ void prefix_rows(const RegionPrefixParams& P, const float* src, bool col, float pr,
                                            float (*sv)[32], int lane, int ip) {
    float run = 0.0f;
    int k = 0;                                           // boundaries crossed so far
    int next = P.bnd[0];
    const int last = P.bnd[P.nb - 1];
    const int W = P.W;
    if (next == 0) { sv[0][lane] = 0.0f; k = 1; next = P.nb > 1 ? P.bnd[1] : -1; }
    for (int r0 = 0; r0 < last; r0 += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (col && r0 + u < last) ? ld_stream_f1(src + (long long)(r0 + u) * W) : 0.0f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            float t = v[u];
            if (ip != 5) {
                t = fmaxf(t, P.eps);
                t = ip == 3 ? t * t * t : ip == 2 ? t * t : ip == 1 ? t : ip == 4 ? (t * t) * (t * t) : approx_pow(t, pr);
            }
            run += (col && r0 + u < last) ? t : 0.0f;
            if (r0 + u + 1 == next) {                    // warp-uniform, rare
                sv[k][lane] = run;
                ++k;
                next = k < P.nb ? P.bnd[k] : -1;
            }
        }
    }
}

__global__ void __launch_bounds__(REGION_WARPS * 32)
region_pool_prefix_kernel(const __grid_constant__ RegionPrefixParams P) {
    __shared__ float sv[REGION_WARPS][REGION_MAX_BND][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long plane = (long long)blockIdx.x * REGION_WARPS + warp;
    if (plane >= P.planes) return;                      // warps are independent
    const int c = (int)(plane % P.C);
    const long long n = plane / P.C;
    const bool gem = P.pool_mode == CIR_POOL_GEM;
    const float pr = gem ? __ldg(P.p + (size_t)c * P.p_stride) : 1.0f;
    const int ip = (pr == 1.0f) ? 1 : (pr == 2.0f) ? 2 : (pr == 3.0f) ? 3 : (pr == 4.0f) ? 4 : 0;
    const int W = P.W;
    const bool col = lane < W;
    const float* src = P.x + plane * (long long)P.H * W + lane;
    prefix_rows(P, src, col, pr, sv[warp], lane, gem ? ip : 5);
    __syncwarp();
    float mine = 0.0f;
    for (int r = 0; r < P.R; ++r) {
        const float a = sv[warp][P.bot[r]][lane] - sv[warp][P.top[r]][lane];
        const int j0 = P.j0[r];
        const float acc = warp_sum((lane >= j0 && lane < j0 + P.w[r]) ? a : 0.0f);
        if ((r & 31) == lane) mine = acc * P.inv_area[r];
        if ((r & 31) == 31 || r == P.R - 1) {            // flush up to 32 results: finalise in parallel, one store each
            const int rbase = r & ~31;
            if (rbase + lane <= r) {
                float y = mine;
                if (gem && ip != 1) y = powf(y, 1.0f / pr);
                P.out[((size_t)n * P.R + rbase + lane) * P.C + c] = y;
            }
        }
    }
}

}  // namespace cir

using namespace cir;

extern "C" int cir_region_pool(const float* x, int N, int C, int H, int W, const int32_t* regions, int R, const float* p,
                               int p_stride, float eps, int pool_mode, float* out, void* stream) {
    CIR_REQUIRE(x && out && regions, CIR_ERR_INVALID_ARG, "cir_region_pool: null pointer");
    CIR_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, CIR_ERR_INVALID_ARG, "cir_region_pool: bad shape %dx%dx%dx%d", N, C, H, W);
    CIR_REQUIRE(R >= 1 && R <= REGION_MAX, CIR_ERR_UNSUPPORTED, "cir_region_pool: R=%d outside [1, %d]", R, REGION_MAX);
    CIR_REQUIRE(pool_mode == CIR_POOL_GEM || pool_mode == CIR_POOL_MAC || pool_mode == CIR_POOL_SPOC, CIR_ERR_INVALID_ARG,
                "cir_region_pool: unknown pool_mode %d", pool_mode);
    CIR_REQUIRE(pool_mode != CIR_POOL_GEM || p, CIR_ERR_INVALID_ARG, "cir_region_pool: GeM needs the exponent p");
    CIR_REQUIRE(H <= 32767 && W <= 32767, CIR_ERR_UNSUPPORTED, "cir_region_pool: map too large");
    if (N == 0) return CIR_OK;
    RegionParams P{};
    for (int r = 0; r < R; ++r) {
        const int32_t* b = regions + 4 * r;
        CIR_REQUIRE(b[0] >= 0 && b[1] >= 0 && b[2] >= 1 && b[3] >= 1 && b[0] + b[2] <= H && b[1] + b[3] <= W,
                    CIR_ERR_INVALID_ARG, "cir_region_pool: region %d = (%d, %d, %d, %d) outside the %d x %d map", r, b[0], b[1],
                    b[2], b[3], H, W);
        P.box[r] = RegionBox{(short)b[0], (short)b[1], (short)b[2], (short)b[3]};
    }
    if (pool_mode != CIR_POOL_MAC && W <= 32) {
        // sum-type pooling on a narrow map: the prefix-sum kernel, if the region list has few distinct row boundaries
        static const char* dbg = getenv("CIR_DEBUG_REGIONS");           // experiments: "band" forces the staging kernel
        RegionPrefixParams Q{};
        int nb = 0;
        bool ok = !(dbg && dbg[0] == 'b');
        auto add_bnd = [&](int row) {
            for (int i = 0; i < nb; ++i) if (Q.bnd[i] == row) return;
            if (nb == REGION_MAX_BND) { ok = false; return; }
            int i = nb++;
            while (i > 0 && Q.bnd[i - 1] > row) { Q.bnd[i] = Q.bnd[i - 1]; --i; }
            Q.bnd[i] = (short)row;
        };
        for (int r = 0; r < R && ok; ++r) { add_bnd(regions[4 * r]); add_bnd(regions[4 * r] + regions[4 * r + 2]); }
        if (ok) {
            for (int r = 0; r < R; ++r) {
                const int32_t* b = regions + 4 * r;
                for (int i = 0; i < nb; ++i) {
                    if (Q.bnd[i] == b[0]) Q.top[r] = (unsigned char)i;
                    if (Q.bnd[i] == b[0] + b[2]) Q.bot[r] = (unsigned char)i;
                }
                Q.j0[r] = (short)b[1]; Q.w[r] = (short)b[3];
                Q.inv_area[r] = 1.0f / (float)(b[2] * b[3]);
            }
            Q.x = x; Q.out = out; Q.p = p; Q.p_stride = p_stride; Q.eps = eps; Q.pool_mode = pool_mode;
            Q.C = C; Q.H = H; Q.W = W; Q.R = R; Q.nb = nb;
            Q.planes = (long long)N * C;
            const long long blocks = (Q.planes + REGION_WARPS - 1) / REGION_WARPS;
            CIR_REQUIRE(blocks <= 0x7fffffffll, CIR_ERR_UNSUPPORTED, "cir_region_pool: too many planes");
            region_pool_prefix_kernel<<<(unsigned)blocks, REGION_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(Q);
            CIR_CHECK_CUDA(cudaGetLastError());
            count_launch();
            return CIR_OK;
        }
    }
    const DeviceInfo& dev = device_info();
    const size_t plane_bytes = (((size_t)H * W + 3) & ~(size_t)3) * 4;
    CIR_REQUIRE(plane_bytes <= (size_t)dev.max_smem_optin, CIR_ERR_UNSUPPORTED,
                "cir_region_pool: a %d x %d plane does not fit the shared-memory staging (%zu B)", H, W, plane_bytes);
    int warps = REGION_WARPS;                           // planes per block: as many as fit, at most 8
    while ((size_t)warps * plane_bytes > (size_t)dev.max_smem_optin) warps >>= 1;
    const size_t smem = (size_t)warps * plane_bytes;
    if (smem > 48 * 1024)
        CIR_CHECK_CUDA(cudaFuncSetAttribute(region_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    P.x = x; P.out = out; P.p = p; P.p_stride = p_stride; P.eps = eps; P.pool_mode = pool_mode;
    P.C = C; P.H = H; P.W = W; P.R = R;
    P.planes = (long long)N * C;
    const long long blocks = (P.planes + warps - 1) / warps;
    CIR_REQUIRE(blocks <= 0x7fffffffll, CIR_ERR_UNSUPPORTED, "cir_region_pool: too many planes");
    region_pool_kernel<<<(unsigned)blocks, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(P);
    CIR_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return CIR_OK;
}
