/*
 * rv_hist_lut.cuh -- the two passes in front of k_chain: per-tile luminance histograms (+ optional gray min/max for the
 * low-contrast gate) and clip / redistribute / prefix-sum -> LUT -> quad tables (cv2.createCLAHE(...).apply, A.3;
 * reference call sites: src/preprocess/ops/clahe_dehaze.py:19,24,29 and src/preprocess/pipeline.py:24-30).
 */
#pragma once
#include "rv_colour.cuh"

#ifndef RV_HIST_DP4A
#define RV_HIST_DP4A 1
#endif
#ifndef RV_HIST_LAB16
#define RV_HIST_LAB16 0                // LAB histograms also take the 16-pixel vector path, table addresses from the packed pixel word (IDP.4A)
#endif

namespace rv {

// ---------------------------------------------------------------------------------------------
// K1: luminance + per-tile histograms
// grid = (tiles, slices, frames), block = 256.  With more than one slice per tile hist must be zeroed (the slices
// accumulate with atomics); with one slice the CTA stores its bins.
// ---------------------------------------------------------------------------------------------
constexpr int HIST_THREADS = 256;
constexpr int HIST_WARPS = HIST_THREADS / 32;

// EXTRA = false is the production instantiation (no luma plane, no gray min/max: nothing but the histogram).
template <int SPACE, bool EXTRA>
__global__ void __launch_bounds__(HIST_THREADS)
k_luma_hist(const uint8_t *__restrict__ src, size_t pitch, size_t fstride, Geo g, int rows_per_slice,
            int32_t *__restrict__ hist, uint8_t *__restrict__ luma_arg, int32_t *__restrict__ gray_arg)
{
    uint8_t *const luma = EXTRA ? luma_arg : nullptr;
    int32_t *const gray_minmax = EXTRA ? gray_arg : nullptr;
    __shared__ uint32_t wh[HIST_WARPS][256];
    __shared__ __align__(16) unsigned char tab_raw[SPACE == 1 ? sizeof(LabHistTabs) : 16];
    LabHistTabs *tabs = reinterpret_cast<LabHistTabs *>(tab_raw);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int tile = blockIdx.x, f = blockIdx.z;
    const int ty = tile / g.grid, tx = tile - ty * g.grid;
    for (int i = tid; i < HIST_WARPS * 256; i += HIST_THREADS) (&wh[0][0])[i] = 0;
    if (SPACE == 1) {
        const uint4 *ts = reinterpret_cast<const uint4 *>(&g_labh);
        uint4 *td = reinterpret_cast<uint4 *>(tabs);
        for (int i = tid; i < (int)(sizeof(LabHistTabs) / 16); i += HIST_THREADS) td[i] = __ldg(ts + i);
    }
    __syncthreads();

    const uint8_t *frame = src + (size_t)f * fstride;
    const int x0 = tx * g.tw, y0 = ty * g.th + blockIdx.y * rows_per_slice;
    const int y1 = min(y0 + rows_per_slice, (ty + 1) * g.th);
    const int nrows = y1 - y0;
    uint32_t *myh = wh[warp];
    int gmin = 255, gmax = 0;
    const bool want_gray = gray_minmax != nullptr;

    auto one = [&](int B, int G, int R) -> int {
        int v;
        if (SPACE == 1) v = lab_L_fast(tabs, B, G, R);
        else v = (4899 * R + 9617 * G + 1868 * B + 8192) >> 14;
        RV_CHECK_IDX(v, 256, "histogram bin");
        atomicAdd(&myh[v], 1u);
        return v;
    };

    const bool interior = (x0 + g.tw <= g.W) && (y1 <= g.H);
    const bool vec_ok = interior && (g.tw % 4 == 0) && (pitch % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(frame) & 3) == 0);
    const bool vec16_ok = vec_ok && (SPACE == 0 || RV_HIST_LAB16) && !EXTRA && RV_HIST_DP4A && (g.tw % 16 == 0) && (pitch % 16 == 0) &&
                          ((reinterpret_cast<uintptr_t>(frame) & 15) == 0);
    [[maybe_unused]] const uint32_t tabs_s = smem_u32(tabs);
    if (nrows > 0 && vec16_ok) {
        // 16 pixels = 48 bytes = three 16-byte loads per group, two groups in flight per thread; Y straight from the packed
        // words with byte dot products (no unpacking), one shared-memory atomic per pixel
        const int gpr = g.tw >> 4;
        const int total = nrows * gpr;
        const float inv_gpr = 1.0f / (float)gpr;
        auto load = [&](int idx, uint4 (&w)[3]) {
            int r = __float2int_rz(__int2float_rn(idx) * inv_gpr);
            int gx = idx - r * gpr;
            if (gx < 0) { gx += gpr; --r; }
            if (gx >= gpr) { gx -= gpr; ++r; }
            const uint4 *p = reinterpret_cast<const uint4 *>(frame + (size_t)(y0 + r) * pitch + 3 * (x0 + 16 * gx));
            w[0] = __ldg(p); w[1] = __ldg(p + 1); w[2] = __ldg(p + 2);
        };
        auto quad = [&](uint32_t w0, uint32_t w1, uint32_t w2) {
            const uint32_t pp[4] = {w0, __funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 16), w2 >> 8};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (SPACE == 1) atomicAdd(&myh[lab_L_px(tabs_s, pp[j])], 1u);
                else atomicAdd(&myh[luma_y(pp[j])], 1u);
            }
        };
        auto consume = [&](const uint4 (&w)[3]) {
            quad(w[0].x, w[0].y, w[0].z);
            quad(w[0].w, w[1].x, w[1].y);
            quad(w[1].z, w[1].w, w[2].x);
            quad(w[2].y, w[2].z, w[2].w);
        };
        int idx = tid;
        for (; idx + HIST_THREADS < total; idx += 2 * HIST_THREADS) {
            uint4 wa[3], wb[3];
            load(idx, wa);
            load(idx + HIST_THREADS, wb);
            consume(wa);
            consume(wb);
        }
        if (idx < total) {
            uint4 wa[3];
            load(idx, wa);
            consume(wa);
        }
    } else if (nrows > 0 && vec_ok) {
        const int gpr = g.tw >> 2;                 // 4-pixel groups per tile row
        const int total = nrows * gpr;
        const float inv_gpr = 1.0f / (float)gpr;
        auto locate = [&](int idx, int &y, int &x) {
            int r = __float2int_rz(__int2float_rn(idx) * inv_gpr);
            int gx = idx - r * gpr;
            if (gx < 0) { gx += gpr; --r; }
            if (gx >= gpr) { gx -= gpr; ++r; }
            y = y0 + r; x = x0 + 4 * gx;
        };
        auto process = [&](uint32_t w0, uint32_t w1, uint32_t w2, int y, int x) {
#if RV_HIST_DP4A
            if (SPACE == 0 && !EXTRA) {
                // Y straight from the packed BGRx word (luma_y), no per-channel unpacking
                const uint32_t p0 = w0, p1 = __funnelshift_r(w0, w1, 24), p2 = __funnelshift_r(w1, w2, 16), p3 = w2 >> 8;
                const uint32_t pp[4] = {p0, p1, p2, p3};
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(&myh[luma_y(pp[j])], 1u);
                return;
            }
#endif
            const int B0 = w0 & 255, G0 = (w0 >> 8) & 255, R0 = (w0 >> 16) & 255;
            const int B1 = w0 >> 24, G1 = w1 & 255, R1 = (w1 >> 8) & 255;
            const int B2 = (w1 >> 16) & 255, G2 = w1 >> 24, R2 = w2 & 255;
            const int B3 = (w2 >> 8) & 255, G3 = (w2 >> 16) & 255, R3 = w2 >> 24;
            const int v0 = one(B0, G0, R0), v1 = one(B1, G1, R1), v2 = one(B2, G2, R2), v3 = one(B3, G3, R3);
            if (luma) {
                uint8_t *lp = luma + ((size_t)f * g.H + y) * g.W + x;
                lp[0] = (uint8_t)v0; lp[1] = (uint8_t)v1; lp[2] = (uint8_t)v2; lp[3] = (uint8_t)v3;
            }
            if (want_gray) {
                const int a0 = gray_of(B0, G0, R0), a1 = gray_of(B1, G1, R1), a2 = gray_of(B2, G2, R2), a3 = gray_of(B3, G3, R3);
                gmin = min(gmin, min(min(a0, a1), min(a2, a3)));
                gmax = max(gmax, max(max(a0, a1), max(a2, a3)));
            }
        };
        // four 12-byte groups (48 bytes) in flight per thread: the pass is DRAM-latency bound otherwise
        constexpr int UNR = 4;
        int idx = tid;
        for (; idx + (UNR - 1) * HIST_THREADS < total; idx += UNR * HIST_THREADS) {
            uint32_t w[UNR][3];
            int yy[UNR], xx[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                locate(idx + u * HIST_THREADS, yy[u], xx[u]);
                const uint32_t *p = reinterpret_cast<const uint32_t *>(frame + (size_t)yy[u] * pitch + 3 * xx[u]);
                w[u][0] = __ldg(p); w[u][1] = __ldg(p + 1); w[u][2] = __ldg(p + 2);
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) process(w[u][0], w[u][1], w[u][2], yy[u], xx[u]);
        }
        for (; idx < total; idx += HIST_THREADS) {
            int y, x;
            locate(idx, y, x);
            const uint32_t *p = reinterpret_cast<const uint32_t *>(frame + (size_t)y * pitch + 3 * x);
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            process(w0, w1, w2, y, x);
        }
    } else if (nrows > 0) {
        // generic path: ragged tiles (REFLECT_101 padding), odd tile widths, unaligned buffers
        const int total = nrows * g.tw;
        for (int idx = tid; idx < total; idx += HIST_THREADS) {
            const int r = idx / g.tw, cx = idx - r * g.tw;
            const int ey = y0 + r, ex = x0 + cx;
            const int y = reflect101(ey, g.H), x = reflect101(ex, g.W);
            const uint8_t *p = frame + (size_t)y * pitch + 3 * x;
            const int B = p[0], G = p[1], R = p[2];
            const int v = one(B, G, R);
            if (ey < g.H && ex < g.W) {            // each real pixel is visited exactly once un-reflected
                if (luma) luma[((size_t)f * g.H + y) * g.W + x] = (uint8_t)v;
                if (want_gray) { const int a = gray_of(B, G, R); gmin = min(gmin, a); gmax = max(gmax, a); }
            }
        }
    }
    __syncthreads();
    {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < HIST_WARPS; ++w) s += wh[w][tid];
        int32_t *bin = &hist[((size_t)f * g.grid * g.grid + tile) * 256 + tid];
        if (gridDim.y == 1) *bin = (int)s;          // one CTA per tile: plain store, the buffer need not be zeroed
        else if (s) atomicAdd(bin, (int)s);
    }
    if (want_gray) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            gmin = min(gmin, __shfl_xor_sync(0xffffffffu, gmin, o));
            gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        }
        if ((tid & 31) == 0 && gmin <= gmax) {
            atomicMin(&gray_minmax[2 * f], gmin);
            atomicMax(&gray_minmax[2 * f + 1], gmax);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: LUT + quad tables.  grid = ((grid+1)^2 quads, frames), block = 128 (4 warps = the 4 tiles
// of the quad; each warp rebuilds its tile's LUT -- 256 bins, 8 per lane, shuffle scans).
// quad q=(qy,qx): tiles ty in {max(qy-1,0), min(qy,grid-1)}, tx likewise (A.3 tx1/tx2 clamping).
// quads[f][q][v] = lut[ty1][tx1][v] | lut[ty1][tx2][v]<<8 | lut[ty2][tx1][v]<<16 | lut[ty2][tx2][v]<<24
// ---------------------------------------------------------------------------------------------
// one warp: the 256-entry LUT of one tile from its histogram (clip, redistribute, prefix sum, scale; A.3), 8 bins per lane,
// returned as eight packed bytes (lo = bins 8*lane .. +3, hi = +4 .. +7)
__device__ __forceinline__ uint2 tile_lut_warp(const int32_t *__restrict__ hist_tile, int clip, float lut_scale, int lane)
{
    int h[8];
    {
        const int4 *hp = reinterpret_cast<const int4 *>(hist_tile + lane * 8);
        const int4 a = __ldg(hp), b = __ldg(hp + 1);
        h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
    }
    if (clip > 0) {
        int clipped = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { clipped += max(h[i] - clip, 0); h[i] = min(h[i], clip); }
#pragma unroll
        for (int o = 16; o; o >>= 1) clipped += __shfl_xor_sync(0xffffffffu, clipped, o);
        const int batch = clipped >> 8, residual = clipped & 255;
        const int step = residual ? max(256 / residual, 1) : 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int bin = lane * 8 + i;
            const int qd = bin / step;
            h[i] += batch + ((residual && bin - qd * step == 0 && qd < residual) ? 1 : 0);
        }
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) h[i] += h[i - 1];
    int run = h[7];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += t;
    }
    const int base = run - h[7];
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float v = __fmul_rn(__int2float_rn(base + h[i]), lut_scale);
        const uint32_t u = (uint32_t)sat8(__float2int_rn(v));
        if (i < 4) lo |= u << (8 * i); else hi |= u << (8 * (i - 4));
    }
    return make_uint2(lo, hi);
}

__global__ void __launch_bounds__(128)
k_build_lut(const int32_t *__restrict__ hist, int grid, int clip, float lut_scale,
            uint8_t *__restrict__ lut, uint32_t *__restrict__ quads)
{
    __shared__ __align__(16) uint8_t sl[4][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x, f = blockIdx.y;
    const int nq1 = grid + 1;
    const int qy = q / nq1, qx = q - qy * nq1;
    const int ty = (w >> 1) ? min(qy, grid - 1) : max(qy - 1, 0);
    const int tx = (w & 1) ? min(qx, grid - 1) : max(qx - 1, 0);
    const size_t tile = (size_t)f * grid * grid + ty * grid + tx;
    const uint2 l8 = tile_lut_warp(hist + tile * 256, clip, lut_scale, lane);
    *reinterpret_cast<uint2 *>(&sl[w][lane * 8]) = l8;
    if (lut != nullptr && w == 3 && qy < grid && qx < grid)      // warp 3 of quad (ty,tx) owns tile (ty,tx)
        *reinterpret_cast<uint2 *>(lut + tile * 256 + lane * 8) = l8;
    __syncthreads();
    uint32_t *qo = quads + ((size_t)f * nq1 * nq1 + q) * 256;
    for (int v = threadIdx.x; v < 256; v += 128)
        qo[v] = (uint32_t)sl[0][v] | ((uint32_t)sl[1][v] << 8) | ((uint32_t)sl[2][v] << 16) | ((uint32_t)sl[3][v] << 24);
}

// Same result, one CTA per (row of quads, frame) for grids up to 16: warp w = (r, tx) builds the LUT of tile
// (r ? min(qy, grid-1) : max(qy-1, 0), tx) once for all grid+1 quads of the row -- every tile LUT is built twice per frame
// instead of four times and the whole pass is a single wave of CTAs.  block = 64 * grid threads.
constexpr int LUT_ROWS_MAX_GRID = 16;
__global__ void __launch_bounds__(64 * LUT_ROWS_MAX_GRID)
k_build_lut_rows(const int32_t *__restrict__ hist, int grid, int clip, float lut_scale,
                 uint8_t *__restrict__ lut, uint32_t *__restrict__ quads)
{
    __shared__ __align__(16) uint8_t sl[2 * LUT_ROWS_MAX_GRID][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int qy = blockIdx.x, f = blockIdx.y;
    const int nq1 = grid + 1;
    const int r = w >= grid ? 1 : 0, tx = w - r * grid;
    const int ty = r ? min(qy, grid - 1) : max(qy - 1, 0);
    const size_t tile = (size_t)f * grid * grid + ty * grid + tx;
    const uint2 l8 = tile_lut_warp(hist + tile * 256, clip, lut_scale, lane);
    RV_CHECK_IDX(w, 2 * grid, "sl (LUT rows)");
    *reinterpret_cast<uint2 *>(&sl[w][lane * 8]) = l8;
    if (lut != nullptr && r == 1 && qy < grid)                   // the lower tile row of quad row qy = tile row qy: written once
        *reinterpret_cast<uint2 *>(lut + tile * 256 + lane * 8) = l8;
    __syncthreads();
    uint32_t *qo = quads + ((size_t)f * nq1 * nq1 + (size_t)qy * nq1) * 256;
    for (int i = threadIdx.x; i < nq1 * 256; i += blockDim.x) {
        const int qx = i >> 8, v = i & 255;
        const int t1 = max(qx - 1, 0), t2 = min(qx, grid - 1);
        RV_CHECK_IDX(grid + t2, 2 * LUT_ROWS_MAX_GRID, "sl (quad packing)");
        qo[i] = (uint32_t)sl[t1][v] | ((uint32_t)sl[t2][v] << 8) | ((uint32_t)sl[grid + t1][v] << 16) |
                ((uint32_t)sl[grid + t2][v] << 24);
    }
}

// per frame: flag = (max - min < thresh)  (pipeline.py:24-30); thresh is ceil() of the configured float, so the integer
// comparison decides exactly like the reference's int-against-double comparison
__global__ void k_gate_flags(const int32_t *__restrict__ gray_minmax, int n, int thresh, int32_t *__restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (gray_minmax[2 * i + 1] - gray_minmax[2 * i] < thresh) ? 1 : 0;
}

__global__ void k_init_minmax(int32_t *mm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mm[2 * i] = 255; mm[2 * i + 1] = 0; }
}

// frames the gate skipped are passed through unchanged
__global__ void k_gate_copy(const uint8_t *__restrict__ src, size_t spitch, size_t sfstride,
                            uint8_t *__restrict__ dst, size_t dpitch, size_t dfstride,
                            int H, int rowbytes, const int32_t *__restrict__ flags)
{
    const int f = blockIdx.z;
    if (flags[f]) return;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        const uint8_t *s = src + (size_t)f * sfstride + (size_t)y * spitch;
        uint8_t *d = dst + (size_t)f * dfstride + (size_t)y * dpitch;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowbytes; i += gridDim.x * blockDim.x) d[i] = s[i];
    }
}

}  // namespace rv
