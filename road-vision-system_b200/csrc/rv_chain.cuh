/*
 * rv_chain.cuh -- k_chain (fused CLAHE apply + inverse colour + k x k median, optional detector-input letterbox) and the
 * stand-alone letterbox kernel.  Reference call sites: src/preprocess/ops/clahe_dehaze.py:19-30,
 * src/preprocess/ops/median_derain.py:14.
 */
#pragma once
#include "rv_colour.cuh"

// Median compare-exchange forms.  Plane values are kept as 0x00vv per 16-bit lane: as unsigned integers they order like v
// (VIMNMX.U16x2, ALU pipe) and as IEEE halves they are the subnormals v * 2^-24, exactly representable together with every
// difference and sum used below (all multiples of 2^-24 below 2^-14; HFMA2/HADD2 on the FMA pipe do not flush them).  (Round 1
// kept 0x6400|v = 1024 + v instead, switch RV_PLANE_BIASED: the bias cost one more instruction per packed word when the planes
// are written; without it the headline runs 3.2 % faster, profiles/r2_l_plane_bias.txt.)  On sm_100a both
// pipes issue 64 lanes/clk/SM (tools/ubench_minmax.cu), so a fraction RV_FMA_NUM/RV_FMA_DEN of the
// compare-exchanges runs on the FMA pipe:  s = relu(b - a);  max = a + s;  min = b - s.
#ifndef RV_FMA_NUM
#define RV_FMA_NUM 0
#endif
#ifndef RV_FMA_DEN
#define RV_FMA_DEN 1
#endif
// RV_PLANE_BIASED = 0 keeps the plane values as plain 0x00vv per lane instead: as IEEE halves they are subnormals v * 2^-24, on which
// the same three operations are still exact (every difference and sum stays a multiple of 2^-24 below 2^-14; no flush-to-zero in
// fma.rn.relu.f16x2 / add.f16x2), the saturation of two packed values is ONE VIMNMX.S16x2.RELU and no bias has to be attached.
#ifndef RV_PLANE_BIASED
#define RV_PLANE_BIASED 0
#endif
#define RV_PLANE_BIAS (RV_PLANE_BIASED ? 0x64006400u : 0u)
#ifndef RV_MEDIAN5_2ROW
#define RV_MEDIAN5_2ROW 1
#endif
#ifndef RV_MEDIAN3_2ROW
#define RV_MEDIAN3_2ROW 1
#endif
// k = 7, 9: the same two-row scheme from the generic hierarchical builder (tools/gen_median_net.py: build_hier_2rows): 143 / 290
// packed operations per output instead of 222 / 605 -> 1080p k7 26.1 k -> 37.2 k fps, k9 11.1 k -> 19.8 k (profiles/r2_y_median79_two_rows.txt)
#ifndef RV_MEDIAN7_2ROW
#define RV_MEDIAN7_2ROW 1
#endif
#ifndef RV_MEDIAN9_2ROW
#define RV_MEDIAN9_2ROW 1
#endif
__device__ __forceinline__ void rv_ce_fma(uint32_t a, uint32_t b, uint32_t &lo, uint32_t &hi)
{
    const __half2 x = *reinterpret_cast<const __half2 *>(&a), y = *reinterpret_cast<const __half2 *>(&b);
    const uint32_t m1bits = 0xBC00BC00u;                       // (-1, -1)
    const __half2 m1 = *reinterpret_cast<const __half2 *>(&m1bits);
    const __half2 s = __hfma2_relu(x, m1, y);                  // relu(y - x)
    const __half2 h = __hadd2(x, s), l = __hsub2(y, s);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
#define RV_CEX_MIX(num, den, n, lo, hi, a, b)                      \
    uint32_t lo, hi;                                               \
    if constexpr (((n) % (den)) < (num)) rv_ce_fma(a, b, lo, hi);  \
    else { lo = __vminu2(a, b); hi = __vmaxu2(a, b); }
#define RV_CEX(n, lo, hi, a, b) RV_CEX_MIX(RV_FMA_NUM, RV_FMA_DEN, n, lo, hi, a, b)
// the 3x3 networks have few full compare-exchanges (64 of 212 ops in the two-row one) and their kernel's CLAHE phase is
// ALU-heavy, so more of them go to the FMA pipe: 2/3 measured best (1/2: -1.1 %, 3/5: -0.5 %, 1/1: -4.6 %)
#ifndef RV_FMA3_NUM
#define RV_FMA3_NUM 2
#define RV_FMA3_DEN 3
#endif
#define RV_CEX3(n, lo, hi, a, b) RV_CEX_MIX(RV_FMA3_NUM, RV_FMA3_DEN, n, lo, hi, a, b)
#define RV_CEX3X2 RV_CEX3              // the production 3x3 network (two rows per task) takes the k3 mix
// the 7x7 / 9x9 two-row networks: their kernels are almost nothing but the median, and 3/7 of the compare-exchanges on the FMA pipe is
// close to where the ALU-pipe, FMA-pipe and issue limits of a pure compare-exchange stream meet (f = 0.4).  Measured against 1/2:
// k7 +1.6 %, k9 +3.5 % (1/3: -2.5 % / +3.4 %, 2/5: +1.0 % / +3.1 %, 3/5: -5 % / -9 %; profiles/r2_aa_median79_mix.txt)
#ifndef RV_FMA79_NUM
#define RV_FMA79_NUM 3
#define RV_FMA79_DEN 7
#endif
#define RV_CEX7X2(n, lo, hi, a, b) RV_CEX_MIX(RV_FMA79_NUM, RV_FMA79_DEN, n, lo, hi, a, b)
#define RV_CEX9X2(n, lo, hi, a, b) RV_CEX_MIX(RV_FMA79_NUM, RV_FMA79_DEN, n, lo, hi, a, b)

#ifndef RV_MEDIAN_NET_FILE
#define RV_MEDIAN_NET_FILE "rv_median_net.h"      // tools/exp_median_order.py builds the kernel against alternative emissions
#endif
#include RV_MEDIAN_NET_FILE

namespace rv {

// ---------------------------------------------------------------------------------------------
// K3+K4: fused apply (+ inverse colour) + median.
// ---------------------------------------------------------------------------------------------
constexpr int TILE_W = 120;            // output pixels per tile row
constexpr int BOX_W = 128;             // staged pixels per row: 4 left + 120 + 4 right
constexpr int LPAD = 4;
#ifndef RV_TILE_H
#define RV_TILE_H 48
#endif
constexpr int TILE_H = RV_TILE_H;
constexpr int HALF = TILE_H / 2;       // u16x2 lanes of the median hold rows (s, s + HALF)
#ifndef RV_CHAIN_THREADS
#define RV_CHAIN_THREADS 256
#endif
constexpr int CHAIN_THREADS = RV_CHAIN_THREADS;
constexpr int CHAIN_WARPS = CHAIN_THREADS / 32;
constexpr int A_STRIDE = BOX_W * 3 + 16; // bytes per staged BGR row: the box starts at the 16-byte boundary at or below
                                       // pixel x0-LPAD (a TMA box must start 16-byte aligned), so up to 12 bytes of slack
constexpr int P_STRIDE = BOX_W;        // words per plane row (one u16x2 word per pixel)
constexpr int O_STRIDE = TILE_W * 3;   // bytes per output staging row
constexpr int MAXQ = 6;                // quad tables kept in shared memory per CTA
#ifndef RV_PLANE_SKEW
#define RV_PLANE_SKEW 24
#endif
#ifndef RV_DP2A_INDEX
#define RV_DP2A_INDEX 1                // YCrCb: Y is consumed as the upper half-word of the scaled luminance sum by IDP.2A.HI (table addresses
#endif                                 // base +- 4 Y in one FMA-pipe instruction, no shift); 0 = shift Y out and add
#ifndef RV_LAB_DP2A
#define RV_LAB_DP2A 1                  // LAB: every table index is the upper (or lower) half-word of a scaled sum, consumed by IDP.2A (FMA pipe)
#endif
#ifndef RV_I2F_BLEND
#define RV_I2F_BLEND 2                 // bit 0: YCrCb kernels, bit 1: LAB kernels take the I2F form of the blend's byte -> float step
#endif
#ifndef RV_DP4A_ADDR
#define RV_DP4A_ADDR 1                 // table addresses from the packed pixel word with IDP.4A (FMA pipe); 0 = byte extraction + add
#endif

struct ChainArgs {
    const uint8_t *src; size_t spitch, sfstride;
    uint8_t *dst; size_t dpitch, dfstride;
    Geo g;
    const uint32_t *quads;             // [frames][(grid+1)^2][256]
    const int32_t *flags;              // optional per-frame gate flags (0 = skip frame)
    int use_tma;                       // stage the box with one cp.async.bulk.tensor per CTA (aligned buffers)
    int prefetch_dist;                 // > 0: also prefetch into L2 the box of the CTA this many linear block ids ahead (the
                                       // one that takes this CTA's place on the SM), so its staging wait is an L2 hit
    const float *colp;                 // per 4-pixel box group: xa[4], xa1[4], -2^23 xa[4], -2^23 xa1[4], quad column[4]
    // optional fused detector-input stage (integer down-scale letterbox, see k_letterbox): 0 = off
    uint16_t *lb_out;                  // [frames][3][lb_S][lb_S] halves, RGB planes, value/255
    int lb_scale, lb_S, lb_top, lb_left;
    int write_full;                    // also store the full-resolution BGR result to dst
};

template <int K> struct MedianCfg;
template <> struct MedianCfg<0> { static constexpr int M = 4; };
template <> struct MedianCfg<3> { static constexpr int M = RV_MEDIAN3_M; };
template <> struct MedianCfg<5> { static constexpr int M = RV_MEDIAN5_M; };
template <> struct MedianCfg<7> { static constexpr int M = RV_MEDIAN7_M; };
template <> struct MedianCfg<9> { static constexpr int M = RV_MEDIAN9_M; };

template <int K, int NC, int M>
__device__ __forceinline__ void median_net(const uint32_t (&v)[NC][K], uint32_t (&out)[M])
{
    if constexpr (K == 3) rv_median3_net(v, out);
    else if constexpr (K == 5) rv_median5_net(v, out);
    else if constexpr (K == 7) rv_median7_net(v, out);
    else rv_median9_net(v, out);
}

template <int MODE, int K>
struct ChainSmem {
    static constexpr int R = K / 2;
    static constexpr int BOX_H = TILE_H + 2 * R;
    static constexpr int NSLOT = HALF + 2 * R;
    // words between the channel planes.  A skew of 24 makes the plane offset = 24 (mod 32 banks): consecutive median tasks
    // (group m fastest, then channel) then step through the banks by 6 words ACROSS the change of channel as well, so a half-warp that
    // straddles two channels' tasks reads 16 distinct even banks with its 8-byte loads instead of colliding two-way (measured on
    // k_chain<YCrCb,5>: bank-conflict wavefronts 40.8 M -> 28.0 M per 64 frames, +0.2 % fps; <YCrCb,3>: +1.1 %).  Not for LAB: the
    // 192 extra bytes push k_chain<LAB,3> from three CTAs per SM to two (-7 %).
    static constexpr int SKEW = (MODE == 1) ? 0 : RV_PLANE_SKEW;
    static constexpr int PLANE = NSLOT * P_STRIDE + SKEW;
    static constexpr size_t a_bytes = (size_t)BOX_H * A_STRIDE;
    static constexpr size_t p_bytes = K > 0 ? (size_t)(3 * PLANE - SKEW) * 4 : (size_t)TILE_H * O_STRIDE;  // K==0: output staging
    static constexpr size_t row_bytes = (size_t)BOX_H * 16;
    static constexpr size_t q_bytes = MODE == 2 ? 0 : (size_t)MAXQ * 256 * 4;
    static constexpr size_t t_bytes = MODE == 1 ? sizeof(LabTabs) : MODE == 0 ? sizeof(YccTabs) : 0;
    static constexpr size_t off_a = 0;
    static constexpr size_t off_p = (a_bytes + 15) & ~(size_t)15;
    static constexpr size_t off_row = off_p + ((p_bytes + 15) & ~(size_t)15);
    static constexpr size_t off_q = off_row + row_bytes;
    static constexpr size_t off_t = off_q + q_bytes;
    static constexpr size_t total = off_t + t_bytes;
};

// MODE: 0 = CLAHE in YCrCb, 1 = CLAHE in LAB, 2 = no CLAHE (median only).  K: 0 (no median), 3, 5, 7, 9.
#ifndef RV_CHAIN_MIN_CTAS
#define RV_CHAIN_MIN_CTAS 2
#endif
#ifndef RV_CHAIN_MIN_CTAS_BIG
#define RV_CHAIN_MIN_CTAS_BIG 2        // k = 7, 9: a 128-register cap keeps two CTAs per SM (the k9 network spills ~36 bytes; one CTA per SM
#endif                                 // with 153 registers and no spill is 15 % slower)
template <int MODE, int K>
__global__ void __launch_bounds__(CHAIN_THREADS, (K <= 5 ? RV_CHAIN_MIN_CTAS : RV_CHAIN_MIN_CTAS_BIG))
k_chain(const ChainArgs a, const __grid_constant__ CUtensorMap tmap)
{
    using S = ChainSmem<MODE, K>;
    constexpr int R = S::R, BOX_H = S::BOX_H;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t tma_bar;
    uint8_t *A = smem + S::off_a;
    uint32_t *P = reinterpret_cast<uint32_t *>(smem + S::off_p);
    float4 *rowp = reinterpret_cast<float4 *>(smem + S::off_row);
    uint32_t *Qs = reinterpret_cast<uint32_t *>(smem + S::off_q);
    const LabTabs *tabs = reinterpret_cast<const LabTabs *>(smem + S::off_t);
    uint8_t *O = K > 0 ? A : reinterpret_cast<uint8_t *>(P);       // output staging

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.z;
    if (a.flags != nullptr && a.flags[f] == 0) return;            // gated-off frame: k_gate_copy handles it
    const Geo g = a.g;
    const int x0 = blockIdx.x * TILE_W, y0 = blockIdx.y * TILE_H;
    const uint8_t *frame = a.src + (size_t)f * a.sfstride;

    // ---- phase 0: stage the BGR box: rows y0-R .. y0-R+BOX_H-1 (rows outside the frame are never read:
    // compute_row clamps the row index = BORDER_REPLICATE of the later median); bytes from the 16-byte boundary
    // at or below 3*(x0-LPAD) (`aoff` bytes of slack, 4 or 12), A_STRIDE bytes per row.
    const int aoff = (3 * (x0 - LPAD)) & 15;
    const int bx0 = 3 * (x0 - LPAD) - aoff;                       // first staged byte of each row (may be < 0)
    if (a.use_tma) {
        // one TMA box per CTA: 100 u32 x BOX_H rows of frame f; out-of-range parts are zero-filled by the hardware
        if (tid == 0) mbar_init(&tma_bar, 1);
        __syncthreads();
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&tma_bar, (uint32_t)(BOX_H * A_STRIDE));
            tma_load_3d(A, &tmap, &tma_bar, bx0 / 4, y0 - R, f);
            if (a.prefetch_dist > 0) {
                const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) + (unsigned)a.prefetch_dist;
                const unsigned per_frame = gridDim.x * gridDim.y;
                const unsigned nf = lin / per_frame, rem = lin - nf * per_frame;
                const unsigned ny = rem / gridDim.x, nx = rem - ny * gridDim.x;
                if (nf < gridDim.z) {
                    const int nbx = 3 * ((int)nx * TILE_W - LPAD);
                    tma_prefetch_3d(&tmap, (nbx - (nbx & 15)) / 4, (int)ny * TILE_H - R, (int)nf);
                }
            }
        }
    } else {
        const int rowbytes = 3 * g.W;
        const bool al4 = ((reinterpret_cast<uintptr_t>(frame) & 3) == 0) && (a.spitch % 4 == 0);
        constexpr int WPR = A_STRIDE / 4;                         // 100 words per staged row
        const bool full = al4 && bx0 >= 0 && bx0 + A_STRIDE <= rowbytes;
        for (int ry = warp; ry < BOX_H; ry += CHAIN_WARPS) {
            const int gy = y0 - R + ry;
            if (gy < 0 || gy >= g.H) continue;
            const uint8_t *rp = frame + (size_t)gy * a.spitch + bx0;
            uint32_t *ar = reinterpret_cast<uint32_t *>(A + ry * A_STRIDE);
            RV_CHECK_IDX(ry * A_STRIDE + A_STRIDE - 1, S::a_bytes, "A (staging store)");
            if (full) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(rp);
                const uint32_t v0 = __ldg(rw + lane), v1 = __ldg(rw + lane + 32), v2 = __ldg(rw + lane + 64);
                ar[lane] = v0; ar[lane + 32] = v1; ar[lane + 64] = v2;
                if (lane + 96 < WPR) ar[lane + 96] = __ldg(rw + lane + 96);
            } else {
                for (int wx = lane; wx < WPR; wx += 32) {
                    const int b = bx0 + 4 * wx;
                    uint32_t v = 0;
                    if (al4 && b >= 0 && b + 4 <= rowbytes) {
                        v = __ldg(reinterpret_cast<const uint32_t *>(rp + 4 * wx));
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (b + k >= 0 && b + k < rowbytes) v |= (uint32_t)rp[4 * wx + k] << (8 * k);
                    }
                    ar[wx] = v;
                }
            }
        }
    }
    // interpolation terms of this lane's four pixels (A.3), evaluated at the clamped coordinate.  They depend only on
    // the column, so the host builds them once per (W, tile width) with the same IEEE single-precision operations
    // (rv_b200.cu: build_colparams) and each lane fetches its five 16-byte records.
    float xa[4], xa1[4], cxa[4], cxa1[4];
    int qxl[4], qcol[4];
    int qx_lo = 0, nqx = 1, qy_lo = 0;
    bool q_smem = true;
    const bool lane_inside = (x0 - LPAD + 4 * lane >= 0) && (x0 - LPAD + 4 * lane + 3 < g.W);
    if (MODE != 2) {
        auto qof = [](int p, float inv) { return (int)floorf(__fsub_rn(__fmul_rn((float)p, inv), 0.5f)) + 1; };
        const int cy_first = min(max(y0 - R, 0), g.H - 1), cy_last = min(max(y0 - R + BOX_H - 1, 0), g.H - 1);
        const int gbase = (TILE_W / 4) * blockIdx.x;              // record of lane 0 (box group x0/4 - 1, stored at +1)
        {
            const float4 *cp = reinterpret_cast<const float4 *>(a.colp) + 5 * (gbase + lane);
            const float4 r0 = __ldg(cp), r1 = __ldg(cp + 1), r2 = __ldg(cp + 2), r3 = __ldg(cp + 3);
            const int4 r4 = __ldg(reinterpret_cast<const int4 *>(cp + 4));
            xa[0] = r0.x; xa[1] = r0.y; xa[2] = r0.z; xa[3] = r0.w;
            xa1[0] = r1.x; xa1[1] = r1.y; xa1[2] = r1.z; xa1[3] = r1.w;
            cxa[0] = r2.x; cxa[1] = r2.y; cxa[2] = r2.z; cxa[3] = r2.w;
            cxa1[0] = r3.x; cxa1[1] = r3.y; cxa1[2] = r3.z; cxa1[3] = r3.w;
            qxl[0] = r4.x; qxl[1] = r4.y; qxl[2] = r4.z; qxl[3] = r4.w;
        }
        qx_lo = __ldg(reinterpret_cast<const int *>(a.colp) + 20 * gbase + 16);
        nqx = __ldg(reinterpret_cast<const int *>(a.colp) + 20 * (gbase + 31) + 19) - qx_lo + 1;
        qy_lo = qof(cy_first, g.inv_th);
        const int nqy = qof(cy_last, g.inv_th) - qy_lo + 1;
        const int nq = nqx * nqy;
        q_smem = nq <= MAXQ;
        const uint32_t *qf = a.quads + (size_t)f * (g.grid + 1) * (g.grid + 1) * 256;
        if (q_smem) {
            // lq / nqx without an integer division: lq < nq <= MAXQ = 6, so (lq * ceil(256 / nqx)) >> 8 is exact
            const int rcp_b = (int)((0x2B3440568000ull >> (8 * (nqx - 1))) & 0xFF);   // ceil(256 / nqx) for nqx = 2..6; 0 for 1
            const int rcp = rcp_b ? rcp_b : 256;
            for (int i = tid; i < nq * 64; i += CHAIN_THREADS) {   // 64 x 16 bytes per quad table
                const int lq = i >> 6, v4 = i & 63;
                const int dq = (lq * rcp) >> 8;
                const int qy = qy_lo + dq, qx = qx_lo + lq - dq * nqx;
                RV_CHECK_IDX(16 * i + 15, S::q_bytes, "Qs (quad table fill)");
                reinterpret_cast<uint4 *>(Qs)[i] = __ldg(reinterpret_cast<const uint4 *>(qf + ((size_t)qy * (g.grid + 1) + qx) * 256) + v4);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            qcol[j] = (qxl[j] - qx_lo) << 10;      // byte offset of the quad column's table; only this form stays live
        }
        for (int ry = tid; ry < BOX_H; ry += CHAIN_THREADS) {
            const int gy = min(max(y0 - R + ry, 0), g.H - 1);
            const float tyf = __fsub_rn(__fmul_rn((float)gy, g.inv_th), 0.5f);
            const float fl = floorf(tyf);
            const float ya = __fsub_rn(tyf, fl);
            const int qy = (int)fl + 1;
            RV_CHECK_IDX(ry, BOX_H, "rowp");
            // .z: shared-memory byte address of the row's first quad table (tables in shared memory), or the global quad row
            rowp[ry] = make_float4(ya, __fsub_rn(1.0f, ya), __int_as_float(q_smem ? (int)(smem_u32(Qs) + (((qy - qy_lo) * nqx) << 10)) : qy),
                                   __int_as_float((gy - (y0 - R)) * A_STRIDE));
        }
        if (MODE == 1) copy_lab_tabs(const_cast<LabTabs *>(tabs));
        if (MODE == 0) {
            const uint4 *ys = reinterpret_cast<const uint4 *>(&g_ycc);
            uint4 *yd = reinterpret_cast<uint4 *>(smem + S::off_t);
            for (int i = tid; i < (int)(sizeof(YccTabs) / 16); i += CHAIN_THREADS) yd[i] = __ldg(ys + i);
        }
    }
    if (a.use_tma && warp == 0) mbar_wait(&tma_bar, 0);     // one warp polls the mbarrier; the others sleep in the barrier below
    __syncthreads();

    // ---- phase 1: CLAHE on the luminance of every staged pixel (or plain unpack when MODE == 2)
    const uint32_t *qglob = (MODE != 2) ? a.quads + (size_t)f * (g.grid + 1) * (g.grid + 1) * 256 : nullptr;
    // YCrCb results are produced UNCLAMPED (K > 0): the saturation to [0,255] happens on the packed u16x2 plane words
    // (two values per VIMNMX.S16x2.RELU) instead of per value; LAB / passthrough values are exact.
    constexpr bool RAW = (MODE == 0) && (K > 0);
    const uint32_t ycc_s = smem_u32(smem + S::off_t);            // shared-memory address of the chroma tables (MODE 0)
    [[maybe_unused]] const uint32_t lab_s = ycc_s;              // ... of the LAB tables (MODE 1): same slot
    // phase 1 is instantiated twice (quad tables in shared memory / fetched from global) and the CTA-uniform choice is
    // made once, outside: a predicated dual path costs issue slots for every masked-off address instruction.
    auto phase1 = [&](auto QS) {
    constexpr bool q_in_smem = decltype(QS)::value;
    auto compute_row = [&](int ry, int (&o)[12]) {
        int Bv[4], Gv[4], Rv[4];
        float4 rp;
        const uint8_t *ar;
        if (MODE == 2) {
            ar = A + (min(max(y0 - R + ry, 0), g.H - 1) - (y0 - R)) * A_STRIDE;
        } else {
            rp = rowp[ry];
            ar = A + __float_as_int(rp.w);                 // staged row of the clamped image row
        }
        uint32_t px[4];                                    // (B, G, R, x) of each pixel in one word
        RV_CHECK_IDX(ar - A, S::a_bytes - A_STRIDE + 1, "A (row base)");
        if (lane_inside) {
            RV_CHECK_IDX((ar - A) + aoff + 12 * lane + 11, S::a_bytes, "A (packed pixel load)");
            const uint32_t *p = reinterpret_cast<const uint32_t *>(ar + aoff + 12 * lane);
            const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
            px[0] = w0; px[1] = __funnelshift_r(w0, w1, 24); px[2] = __funnelshift_r(w1, w2, 16); px[3] = w2 >> 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                Bv[j] = px[j] & 255;
                Gv[j] = __byte_perm(px[j], 0, 0x4441);
                Rv[j] = __byte_perm(px[j], 0, 0x4442);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cx = min(max(x0 - LPAD + 4 * lane + j, 0), g.W - 1);
                const uint8_t *p = ar + aoff + 3 * (cx - (x0 - LPAD));
                RV_CHECK_IDX(p - A, S::a_bytes - 2, "A (edge pixel load)");
                Bv[j] = p[0]; Gv[j] = p[1]; Rv[j] = p[2];
                px[j] = (uint32_t)Bv[j] | ((uint32_t)Gv[j] << 8) | ((uint32_t)Rv[j] << 16);
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { o[j] = Bv[j]; o[4 + j] = Gv[j]; o[8 + j] = Rv[j]; }
            return;
        }
        const float ya = rp.x, ya1 = rp.y;
        int ly[4], lx[4], lz[4];                                 // LAB: y and the two XZ arguments of the four pixels
        const int qrow = __float_as_int(rp.z);                    // byte address of the row's quad tables in shared memory, or the global row
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int L, c1 = 0, c2 = 0;
            [[maybe_unused]] uint32_t eB = 0, eR = 0, acc = 0;
            if (MODE == 1) {
#if RV_LAB_DP2A
                lab_fwd_px2(lab_s, px[j], acc, c1, c2);
                L = (int)(acc >> 16);                             // only the global-table path and the debug checks use it
#elif RV_DP4A_ADDR
                lab_fwd_px(tabs, lab_s, px[j], L, c1, c2);
#else
                lab_fwd(tabs, Bv[j], Gv[j], Rv[j], L, c1, c2);
#endif
            } else {
                // A.1 forward: Y from the packed pixel word; the chroma round trip comes from the tables (see YccTabs):
                // entries of d = B - Y and d = R - Y
                // one shared term (table base - 4 Y) for both look-ups; the empty asm keeps the compiler from re-associating it
                // into a subtraction per channel.  The tables are constant after the barrier above and the address depends on
                // this pixel, so a plain (non-volatile) shared load is safe.
#if RV_DP2A_INDEX
                // Y stays inside the scaled sum (its upper half-word): base - 4 Y is one IDP.2A.HI with the byte -4, no shift
                acc = luma_acc16(px[j]);
                L = (int)(acc >> 16);                             // only the global-table path and the debug checks use it
                uint32_t yrow = (uint32_t)__dp2a_hi((int)acc, (int)0xFC000000u, (int)(ycc_s + 4u * 255u));
#else
                L = (int)luma_y(px[j]);
                uint32_t yrow = ycc_s + 4u * 255u - 4u * (uint32_t)L;
#endif
                asm("" : "+r"(yrow));
                // table addresses yrow + 4 B and yrow + 4 R straight from the packed pixel word with one byte dot product each
                // (IDP.4A, FMA pipe: selects the byte, scales it by the entry size and adds the base) instead of a byte extraction
                // and a shift-and-add on the ALU pipe, which is the busier one in this kernel
#if RV_DP4A_ADDR
                asm("ld.shared.u32 %0, [%1];" : "=r"(eB) : "r"(__dp4a(px[j], 0x00000004u, yrow)));
                asm("ld.shared.u32 %0, [%1+2048];" : "=r"(eR) : "r"(__dp4a(px[j], 0x00040000u, yrow)));
#else
                asm("ld.shared.u32 %0, [%1];" : "=r"(eB) : "r"(yrow + 4u * (uint32_t)Bv[j]));
                asm("ld.shared.u32 %0, [%1+2048];" : "=r"(eR) : "r"(yrow + 4u * (uint32_t)Rv[j]));
#endif
            }
            uint32_t q;
            if constexpr (q_in_smem) RV_CHECK_IDX(((qrow + qcol[j]) - (int)smem_u32(Qs)) / 4 + L, MAXQ * 256, "Qs (LUT quad load)");
            if constexpr (MODE == 0) RV_CHECK_IDX(255 - L + (int)(px[j] & 255), 512, "ycc table (B - Y)");
            if constexpr (MODE == 0) RV_CHECK_IDX(255 - L + (int)((px[j] >> 16) & 255), 512, "ycc table (R - Y)");
            if constexpr (q_in_smem) {
                // entry L of the quad table at byte address qrow + qcol[j]
#if RV_DP2A_INDEX
                if constexpr (MODE == 0 || (MODE == 1 && RV_LAB_DP2A))
                    asm("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(__dp2a_hi(acc, 0x04000000u, (uint32_t)(qrow + qcol[j]))));
                else asm("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"((uint32_t)(qrow + qcol[j]) + 4u * (uint32_t)L));
#else
                asm("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"((uint32_t)(qrow + qcol[j]) + 4u * (uint32_t)L));
#endif
            } else {
                q = __ldg(qglob + (((size_t)qrow * (g.grid + 1) + ((qcol[j] >> 10) + qx_lo)) << 8) + L);
            }
            // 0x4B0000vv = 2^23 + vv ; fma(2^23 + v, w, -2^23 * w) == v * w rounded once (A.3: no FMA contraction
            // between the products and the sums -- each step below is individually rounded)
            float p00, p01, p10, p11;
            if constexpr (MODE == 1 ? (RV_I2F_BLEND & 2) != 0 : (RV_I2F_BLEND & 1) != 0) {
                // bytes -> floats with I2F.U8 (byte selector; conversion unit, off the ALU pipe) and plain multiplies.  Pays in the LAB
                // kernels, whose ALU pipe is the binding one (+0.5 % k3, +1.2 % k5); costs 1.1 % in k_chain<YCrCb,5>.
                p00 = __fmul_rn((float)(q & 255u), xa1[j]);
                p01 = __fmul_rn((float)((q >> 8) & 255u), xa[j]);
                p10 = __fmul_rn((float)((q >> 16) & 255u), xa1[j]);
                p11 = __fmul_rn((float)(q >> 24), xa[j]);
            } else {
                const float m00 = __uint_as_float(__byte_perm(q, 0x4B000000u, 0x7440));
                const float m01 = __uint_as_float(__byte_perm(q, 0x4B000000u, 0x7441));
                const float m10 = __uint_as_float(__byte_perm(q, 0x4B000000u, 0x7442));
                const float m11 = __uint_as_float(__byte_perm(q, 0x4B000000u, 0x7443));
                p00 = __fmaf_rn(m00, xa1[j], cxa1[j]);
                p01 = __fmaf_rn(m01, xa[j], cxa[j]);
                p10 = __fmaf_rn(m10, xa1[j], cxa1[j]);
                p11 = __fmaf_rn(m11, xa[j], cxa[j]);
            }
            const float top = __fmul_rn(__fadd_rn(p00, p01), ya1);
            const float bot = __fmul_rn(__fadd_rn(p10, p11), ya);
            const float res = __fadd_rn(top, bot);
            // round-half-even via the 1.5*2^23 trick; res <= 255*(1 + 1e-6), so the result is already in [0,255]
            // RAW: the float's upper bits are left in place (with RV_PLANE_BIASED the bias 0x6400 also rides along in the magic constant); only the
            // low 16 bits of the sums below are ever used (pack2 keeps the low halves), so no masking is needed.
            // YCrCb: - 256 because the tables' f fields and the G sum carry + 256.
            constexpr float MAGIC = 12582912.0f + ((RAW && RV_PLANE_BIASED) ? 25600.0f : 0.0f) - (MODE == 0 ? 256.0f : 0.0f);
            const int Lw = __float_as_int(__fadd_rn(res, MAGIC));
            if (MODE == 1) {
#if RV_LAB_DP2A
                lab_inv_args_w(lab_s, (uint32_t)Lw, c1, c2, ly[j], lx[j], lz[j]);
#else
                lab_inv_args(tabs, Lw & 0x1FF, c1, c2, ly[j], lx[j], lz[j]);
#endif
            } else {
                // A.1 inverse: B' = Y' + fB, G' = Y' + ((tB + tR + 8192) >> 14), R' = Y' + fR  (table layout: rv_colour.cuh, YccTabs)
#if RV_YCC16
                // the f terms sit in the low half-words, so plain adds give B' and R' in their low 16 bits (the upper bits are never
                // used: pack2 keeps the low halves); the G term is the top nine bits of the sum of the two entries
                const int bb = Lw + (int)eB;
                const int gg = Lw + (int)((eB + eR) >> 23);
                const int rr = Lw + (int)eR;
                if (RAW) { o[j] = bb; o[4 + j] = gg; o[8 + j] = rr; }
                else { o[j] = sat8((int)(short)bb); o[4 + j] = sat8((int)(short)gg); o[8 + j] = sat8((int)(short)rr); }
#else
                const int L2 = RAW ? Lw : (Lw << 16) >> 16;          // non-RAW: sign-extended Y' - 256
                const int bb = L2 + (int)(eB >> 22);
                const int gg = L2 + (int)(((eB + eR) << 10) >> 23);
                const int rr = L2 + (int)(eR >> 22);
                if (RAW) { o[j] = bb; o[4 + j] = gg; o[8 + j] = rr; }
                else { o[j] = sat8(bb); o[4 + j] = sat8(gg); o[8 + j] = sat8(rr); }
#endif
            }
        }
        if (MODE == 1) {
            const int lowest = min(min(min(lx[0], lz[0]), min(lx[1], lz[1])), min(min(lx[2], lz[2]), min(lx[3], lz[3])));
            if (__any_sync(0xffffffffu, lowest <= 3390)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) lab_inv_tail<false>(tabs, ly[j], lx[j], lz[j], o[j], o[4 + j], o[8 + j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) lab_inv_tail<true>(tabs, ly[j], lx[j], lz[j], o[j], o[4 + j], o[8 + j]);
            }
        }
    };
    // two rows' values of one pixel/channel -> one plane word (low half = first row), saturated
    auto pack2 = [&](int lo, int hi) -> uint32_t {
        const uint32_t w = __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
#if RV_PLANE_BIASED
        if (RAW) return __vmins2(__vmaxs2(w, RV_PLANE_BIAS), RV_PLANE_BIAS | 0x00FF00FFu);
        return w | RV_PLANE_BIAS;
#else
        if (RAW) return __vimin_s16x2_relu(w, 0x00FF00FFu);       // max(min(v, 255), 0) on both halves at once
        return w;
#endif
    };

    if constexpr (K == 0) {
        // no median: write the interleaved result straight to the staging tile (lanes 0 and 31 hold halo only)
        for (int ry = warp; ry < TILE_H; ry += CHAIN_WARPS) {
            if (y0 + ry >= g.H) break;
            int o[12];
            compute_row(ry, o);                      // all 32 lanes: compute_row votes across the warp (LAB inverse)
            if (lane >= 1 && lane <= 30) {
                RV_CHECK_IDX(ry * O_STRIDE + 12 * (lane - 1) + 11, S::p_bytes, "O (K = 0 staging)");
                uint32_t *op = reinterpret_cast<uint32_t *>(O + ry * O_STRIDE + 12 * (lane - 1));
                op[0] = (uint32_t)o[0] | ((uint32_t)o[4] << 8) | ((uint32_t)o[8] << 16) | ((uint32_t)o[1] << 24);
                op[1] = (uint32_t)o[5] | ((uint32_t)o[9] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[6] << 24);
                op[2] = (uint32_t)o[10] | ((uint32_t)o[3] << 8) | ((uint32_t)o[7] << 16) | ((uint32_t)o[11] << 24);
            }
        }
    } else {
        // planes P[c][slot][px]: low half = row `slot`, high half = row `slot + HALF` of the box
        for (int s = warp; s < HALF; s += CHAIN_WARPS) {
            int o0[12], o1[12];
            compute_row(s, o0);
            compute_row(s + HALF, o1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint4 w;
                w.x = pack2(o0[4 * c + 0], o1[4 * c + 0]);
                w.y = pack2(o0[4 * c + 1], o1[4 * c + 1]);
                w.z = pack2(o0[4 * c + 2], o1[4 * c + 2]);
                w.w = pack2(o0[4 * c + 3], o1[4 * c + 3]);
                RV_CHECK_IDX(4 * (c * S::PLANE + s * P_STRIDE + 4 * lane) + 15, S::p_bytes, "P (plane store)");
                *reinterpret_cast<uint4 *>(P + c * S::PLANE + s * P_STRIDE + 4 * lane) = w;
            }
            if (s < 2 * R) {       // rows [HALF, HALF+2R) are also the low half of slots [HALF, HALF+2R)
                compute_row(s + TILE_H, o0);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    uint4 w;
                    w.x = pack2(o1[4 * c + 0], o0[4 * c + 0]);
                    w.y = pack2(o1[4 * c + 1], o0[4 * c + 1]);
                    w.z = pack2(o1[4 * c + 2], o0[4 * c + 2]);
                    w.w = pack2(o1[4 * c + 3], o0[4 * c + 3]);
                    RV_CHECK_IDX(4 * (c * S::PLANE + (s + HALF) * P_STRIDE + 4 * lane) + 15, S::p_bytes, "P (tail plane store)");
                    *reinterpret_cast<uint4 *>(P + c * S::PLANE + (s + HALF) * P_STRIDE + 4 * lane) = w;
                }
            }
        }
    }
    };   // phase1
    if (q_smem) phase1(std::true_type{}); else phase1(std::false_type{});
    __syncthreads();

    // ---- phase 2: k x k median per channel plane; lanes of each u16x2 are output rows (s, s+HALF)
    constexpr bool TWO_ROW = (K == 5 && RV_MEDIAN5_2ROW) || (K == 3 && RV_MEDIAN3_2ROW) || (K == 7 && RV_MEDIAN7_2ROW) || (K == 9 && RV_MEDIAN9_2ROW);
    if constexpr (TWO_ROW) {
        // two vertically adjacent output rows per task (slots s, s+1): the K-1 middle window rows are shared
        constexpr int M = (K == 5) ? RV_MEDIAN5X2_M : (K == 3) ? RV_MEDIAN3X2_M : (K == 7) ? RV_MEDIAN7X2_M : RV_MEDIAN9X2_M;
        constexpr int NG = TILE_W / M;
        constexpr int NC = M + K - 1;
        constexpr int NR = K + 1;                    // plane rows per task
        constexpr int C0 = LPAD - R;                 // first needed plane word, relative to the group's first output
        static_assert((M == 4 || M == 6) && TILE_W % M == 0 && HALF % 2 == 0, "two-row median layout");
        // task = (row pair sp, channel c, group m), m fastest; a thread's next task is CHAIN_THREADS further on.  The three
        // coordinates are carried incrementally (one division per thread instead of four per task).
        constexpr int DM = CHAIN_THREADS % NG, DT = CHAIN_THREADS / NG, DC = DT % 3, DSP = DT / 3;
        int m = tid % NG, c = (tid / NG) % 3, sp = (tid / NG) / 3;
        for (; sp < HALF / 2; ) {
            const int s = 2 * sp;
            const bool live = (x0 + M * m < g.W) && (y0 + s < g.H);
            const int mo = m, co = c;
            // advance to this thread's next task
            m += DM; c += DC; sp += DSP;
            if (m >= NG) { m -= NG; ++c; }
            if (c >= 3) { c -= 3; ++sp; }
            if constexpr (DC + 1 >= 3) { if (c >= 3) { c -= 3; ++sp; } }
            if (!live) continue;
            uint32_t v[NC][NR];
            const uint32_t *pc = P + co * S::PLANE + s * P_STRIDE;
            RV_CHECK_IDX(4 * ((pc - P) + (NR - 1) * P_STRIDE + M * mo + C0 + NC - 1) + 3, S::p_bytes, "P (two-row median load)");
            RV_CHECK_IDX((pc - P) + M * mo + (C0 & ~1), S::p_bytes / 4, "P (two-row median first word)");
            if constexpr (M == 4) {
#pragma unroll
                for (int d = 0; d < NR; ++d) {
                    const uint4 *p4 = reinterpret_cast<const uint4 *>(pc + d * P_STRIDE + 4 * mo);
                    const uint4 q0 = p4[0], q1 = p4[1], q2 = p4[2];
                    const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[cc][d] = w[C0 + cc];
                }
            } else if constexpr (M % 2 == 0) {
                // 8-byte loads from the even word at or below the first needed column (k = 3: one unused leading word);
                // with M = 6 the 16 lanes of a phase hit 16 distinct even banks (6m + 2 mod 32): conflict-free
                constexpr int C0E = C0 & ~1, SKIP = C0 - C0E, NW = (SKIP + NC + 1) / 2;
#pragma unroll
                for (int d = 0; d < NR; ++d) {
                    const uint2 *p2 = reinterpret_cast<const uint2 *>(pc + d * P_STRIDE + M * mo + C0E);
                    uint32_t w[2 * NW];
#pragma unroll
                    for (int cc = 0; cc < NW; ++cc) {
                        const uint2 q = p2[cc];
                        w[2 * cc] = q.x; w[2 * cc + 1] = q.y;
                    }
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[cc][d] = w[SKIP + cc];
                }
            } else {
#pragma unroll
                for (int d = 0; d < NR; ++d)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[cc][d] = pc[d * P_STRIDE + M * mo + C0 + cc];
            }
            uint32_t out[2][M];
            if constexpr (K == 5) rv_median5x2_net(v, out);
            else if constexpr (K == 7) rv_median7x2_net(v, out);
            else if constexpr (K == 9) rv_median9x2_net(v, out);
            else rv_median3x2_net(v, out);
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                uint8_t *o0 = O + (s + hrow) * O_STRIDE + 3 * M * mo + co;
                uint8_t *o1 = o0 + HALF * O_STRIDE;
                RV_CHECK_IDX((o1 - O) + 3 * (M - 1), S::a_bytes, "O (two-row median store)");
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    o0[3 * j] = (uint8_t)(out[hrow][j] & 255);
                    o1[3 * j] = (uint8_t)((out[hrow][j] >> 16) & 255);
                }
            }
        }
        __syncthreads();
    } else if constexpr (K > 0) {
        constexpr int M = MedianCfg<K>::M;
        constexpr int NG = TILE_W / M;               // output groups per row
        constexpr int NC = M + K - 1;                // pixel columns per group
        for (int task = tid; task < 3 * NG * HALF; task += CHAIN_THREADS) {
            const int m = task % NG;
            const int t2 = task / NG;
            const int c = t2 % 3, s = t2 / 3;
            if (x0 + M * m >= g.W) continue;
            if (y0 + s >= g.H) continue;
            uint32_t v[NC][K];
            const uint32_t *pc = P + c * S::PLANE + s * P_STRIDE;
            constexpr int C0 = LPAD - R;                 // first needed plane word, relative to the group's first output
            RV_CHECK_IDX(4 * ((pc - P) + (K - 1) * P_STRIDE + M * m + C0 + NC - 1) + 3, S::p_bytes, "P (median load)");
            RV_CHECK_IDX((pc - P) + M * m + C0, S::p_bytes / 4, "P (median first word)");
            if constexpr (M == 4) {
                // 16-byte loads of words 4m .. 4m+11 (conflict-free: 8 lanes x 16 B per phase)
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    const uint4 *p4 = reinterpret_cast<const uint4 *>(pc + d * P_STRIDE + 4 * m);
                    const uint4 q0 = p4[0], q1 = p4[1], q2 = p4[2];
                    const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[cc][d] = w[C0 + cc];
                }
            } else if constexpr (M % 2 == 0) {
                // 8-byte loads from the even word at or below the first needed column; with M = 6 the 16 lanes of a phase hit 16
                // distinct even banks (6m+2 mod 32)
                constexpr int C0E = C0 & ~1, SKIP = C0 - C0E, NW = (SKIP + NC + 1) / 2;
                static_assert(TILE_W - M + C0E + 2 * NW <= P_STRIDE, "median loads stay inside the plane row");
#pragma unroll
                for (int d = 0; d < K; ++d) {
                    const uint2 *p2 = reinterpret_cast<const uint2 *>(pc + d * P_STRIDE + M * m + C0E);
                    uint32_t w[2 * NW];
#pragma unroll
                    for (int cc = 0; cc < NW; ++cc) {
                        const uint2 q = p2[cc];
                        w[2 * cc] = q.x; w[2 * cc + 1] = q.y;
                    }
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[cc][d] = w[SKIP + cc];
                }
            } else {
#pragma unroll
                for (int d = 0; d < K; ++d)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) v[cc][d] = pc[d * P_STRIDE + M * m + C0 + cc];
            }
            uint32_t out[M];
            median_net<K, NC, M>(v, out);
            uint8_t *o0 = O + s * O_STRIDE + 3 * M * m + c;
            uint8_t *o1 = o0 + HALF * O_STRIDE;
            RV_CHECK_IDX((o1 - O) + 3 * (M - 1), S::a_bytes, "O (median store)");
#pragma unroll
            for (int j = 0; j < M; ++j) {
                o0[3 * j] = (uint8_t)(out[j] & 255);
                o1[3 * j] = (uint8_t)((out[j] >> 16) & 255);
            }
        }
        __syncthreads();
    } else {
        __syncthreads();
    }

    // ---- phase 3: coalesced store of the staging tile (one warp per row, three words per lane)
    if (a.write_full) {
        uint8_t *dframe = a.dst + (size_t)f * a.dfstride;
        const int nb = min(3 * TILE_W, 3 * (g.W - x0));          // valid bytes per row
        const bool al4 = ((reinterpret_cast<uintptr_t>(dframe) & 3) == 0) && (a.dpitch % 4 == 0);
        const int rows = min(TILE_H, g.H - y0);
        constexpr int WPR = O_STRIDE / 4;                        // 90 words per row
        const bool al8 = ((reinterpret_cast<uintptr_t>(dframe) & 7) == 0) && (a.dpitch % 8 == 0);
        if (al8 && nb == 3 * TILE_W && rows == TILE_H) {
            // full tile, 8-byte aligned rows (3*x0 = 360*bx): 45 double words per row, four rows per warp, fully unrolled
            uint8_t *dp = dframe + (size_t)(y0 + warp) * a.dpitch + 3 * (size_t)x0 + 8 * lane;
            const uint8_t *op = O + warp * O_STRIDE + 8 * lane;
            RV_CHECK_IDX((TILE_H - CHAIN_WARPS + warp) * O_STRIDE + 8 * (lane < 13 ? lane + 32 : lane) + 7, (K > 0 ? S::a_bytes : S::p_bytes), "O (tile store)");
            const size_t dstep = (size_t)CHAIN_WARPS * a.dpitch;
#pragma unroll
            for (int k = 0; k < TILE_H / CHAIN_WARPS; ++k) {
                const uint2 v0 = *reinterpret_cast<const uint2 *>(op);
                *reinterpret_cast<uint2 *>(dp) = v0;
                if (lane < 45 - 32) *reinterpret_cast<uint2 *>(dp + 256) = *reinterpret_cast<const uint2 *>(op + 256);
                dp += dstep;
                op += CHAIN_WARPS * O_STRIDE;
            }
        } else
        for (int ry = warp; ry < rows; ry += CHAIN_WARPS) {
            uint8_t *dp = dframe + (size_t)(y0 + ry) * a.dpitch + 3 * (size_t)x0;
            const uint32_t *orow = reinterpret_cast<const uint32_t *>(O + ry * O_STRIDE);
            if (al4 && nb == 3 * TILE_W) {
                uint32_t *dw = reinterpret_cast<uint32_t *>(dp);
                const uint32_t v0 = orow[lane], v1 = orow[lane + 32];
                dw[lane] = v0; dw[lane + 32] = v1;
                if (lane + 64 < WPR) dw[lane + 64] = orow[lane + 64];
            } else {
                for (int wx = lane; wx < WPR; wx += 32) {
                    const int b = 4 * wx;
                    if (b >= nb) break;
                    const uint32_t v = orow[wx];
                    if (al4 && b + 4 <= nb) *reinterpret_cast<uint32_t *>(dp + b) = v;
                    else for (int k = 0; k < 4 && b + k < nb; ++k) dp[b + k] = (uint8_t)(v >> (8 * k));
                }
            }
        }
    }
    // ---- phase 3b (optional): detector input.  Integer down-scale s: cv2.resize(INTER_LINEAR) degenerates to the
    // centre pixel (odd s) or the rounded mean of the centre 2x2 block (even s; = the fixed-point formula with both
    // weights 1024), which never straddles a tile because tile sizes are even and multiples of s are tile aligned in x.
    if (a.lb_out != nullptr) {
        const int sc = a.lb_scale, S = a.lb_S;
        const int off = (sc - 1) >> 1;                           // first source pixel of output d is sc*d + off
        const int nw = g.W / sc, nh = g.H / sc;
        const int dx0 = (x0 - off + sc - 1) / sc;                // outputs whose first source column is in this tile
        const int dx1 = min((x0 + TILE_W - 1 - off) / sc, nw - 1);
        const int dy0 = (y0 - off + sc - 1) / sc;
        const int dy1 = min((y0 + TILE_H - 1 - off) / sc, nh - 1);
        const int ncol = dx1 - dx0 + 1, nrow = dy1 - dy0 + 1;
        if (ncol > 0 && nrow > 0) {
            uint16_t *ob = a.lb_out + (size_t)f * 3 * S * S;
            const bool even = (sc & 1) == 0;
            for (int i = tid; i < ncol * nrow; i += CHAIN_THREADS) {
                const int ry = i / ncol, rx = i - ry * ncol;
                const int dy = dy0 + ry, dx = dx0 + rx;
                const uint8_t *p = O + (sc * dy + off - y0) * O_STRIDE + 3 * (sc * dx + off - x0);
                RV_CHECK_IDX((p - O) + (even ? O_STRIDE + 3 : 0) + 2, TILE_H * O_STRIDE, "O (fused letterbox read)");
                int v[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (even) v[c] = (p[c] + p[c + 3] + p[c + O_STRIDE] + p[c + O_STRIDE + 3] + 2) >> 2;
                    else v[c] = p[c];
                }
                uint16_t *o = ob + (size_t)(a.lb_top + dy) * S + a.lb_left + dx;
#pragma unroll
                for (int c = 0; c < 3; ++c)                      // planes are R, G, B; staging is B, G, R
                    o[(size_t)(2 - c) * S * S] = __half_as_ushort(__float2half_rn(__fdiv_rn((float)v[c], 255.0f)));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Detector-input stage, general form: letterbox to S x S with cv2.resize(INTER_LINEAR) 8-bit
// fixed-point arithmetic (coefficients scaled by 2048, tables built on the host exactly as OpenCV builds them),
// constant padding, BGR->RGB, HWC->CHW, /255, fp16.  grid (ceil(S/32), ceil(S/8), frames), block (32, 8).
// only_pad = 1 writes just the padding (the image rectangle is produced by k_chain's fused phase 3b).
// ---------------------------------------------------------------------------------------------
struct LbArgs {
    const uint8_t *src; size_t spitch, sfstride;
    uint16_t *out;
    int H, W, S, nw, nh, top, left, pad, only_pad;
    const int32_t *xofs;     // [nw] first source column (already clamped)
    const int16_t *xa;       // [nw][2]
    const int32_t *yofs;     // [nh][2] both source rows (clamped)
    const int16_t *ya;       // [nh][2]
};

__global__ void __launch_bounds__(256) k_letterbox(const LbArgs a)
{
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (dx >= a.S || dy >= a.S) return;
    uint16_t *o = a.out + (size_t)f * 3 * a.S * a.S + (size_t)dy * a.S + dx;
    const int ix = dx - a.left, iy = dy - a.top;
    const size_t plane = (size_t)a.S * a.S;
    if (ix < 0 || ix >= a.nw || iy < 0 || iy >= a.nh) {
        const uint16_t pv = __half_as_ushort(__float2half_rn(__fdiv_rn((float)a.pad, 255.0f)));
        o[0] = pv; o[plane] = pv; o[2 * plane] = pv;
        return;
    }
    if (a.only_pad) return;
    const int sx = a.xofs[ix], sx1 = min(sx + 1, a.W - 1);
    const int a0 = a.xa[2 * ix], a1 = a.xa[2 * ix + 1];
    const int b0 = a.ya[2 * iy], b1 = a.ya[2 * iy + 1];
    const uint8_t *r0 = a.src + (size_t)f * a.sfstride + (size_t)a.yofs[2 * iy] * a.spitch;
    const uint8_t *r1 = a.src + (size_t)f * a.sfstride + (size_t)a.yofs[2 * iy + 1] * a.spitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = r0[3 * sx + c] * a0 + r0[3 * sx1 + c] * a1;       // horizontal pass, scaled by 2048
        const int h1 = r1[3 * sx + c] * a0 + r1[3 * sx1 + c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;   // VResizeLinear<uchar,int,short>
        o[(size_t)(2 - c) * plane] = __half_as_ushort(__float2half_rn(__fdiv_rn((float)min(max(v, 0), 255), 255.0f)));
    }
}

}  // namespace rv
