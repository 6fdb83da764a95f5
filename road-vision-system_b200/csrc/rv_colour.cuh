/*
 * rv_colour.cuh -- OpenCV's 8-bit fixed-point colour arithmetic and the device tables
 * that carry it: LAB gamma / cube-root / inverse tables, the YCrCb chroma round-trip table of k_chain, the folded LAB
 * luminance tables of the histogram pass.  The reference only calls cv2.cvtColor: src/preprocess/ops/clahe_dehaze.py:22-30.
 */
#pragma once
#include "rv_common.cuh"
#include "rv_lab_tables.h"

#include <stddef.h>

namespace rv {


// ---------------------------------------------------------------------------------------------
// LAB tables in global memory (copied into shared memory by the kernels that need them)
// ---------------------------------------------------------------------------------------------
// RV_LAB_FT_BIAS: constant carried by every entry of LabTabs::ft (0 = OpenCV's table as it is; 10484 = bdiv's constant, see lab_inv_args_w)
#ifndef RV_LAB_FT_BIAS
#define RV_LAB_FT_BIAS 10484           // measured +1.1 % on the LAB kernels (profiles/r2_ak_lab_ft_bias.txt)
#endif
struct LabTabs {
    uint16_t g8[256];
    uint16_t yt[256];
    uint16_t ft[256];
    uint16_t cb[2048];     // entries 0..2040 reachable; [2041..2047] padding
    uint8_t ig[4096];
};
static_assert(sizeof(LabTabs) % 16 == 0, "LabTabs must be 16-byte granular");
__device__ LabTabs g_lab;   // filled once per context from rv_lab_tables.h

// YCrCb chroma round trip (A.1) as two 511-entry tables indexed by d + 255, d = B - Y (first 512 words) or R - Y (next 512).
// Cb and Cr only pass through the CLAHE, so B' = Y' + fB(d_B), R' = Y' + fR(d_R) and G' = Y' + ((tB + tR + 8192) >> 14) with
//   fB = ((Cb - 128) * 29049 + 8192) >> 14,  tB = (Cb - 128) * -5636   (Cb = sat8((d * 9241 + (128 << 14) + 8192) >> 14)),
//   fR = ((Cr - 128) * 22987 + 8192) >> 14,  tR = (Cr - 128) * -11698  (Cr likewise with 11682).
// RV_YCC16 = 1 (round 2).  Entry: low half-word = f + 256 (in [29, 481]: the two low halves add without a carry), high half-word =
// the 16-bit G term of rv_ycc_g.h (seven fractional bits; (gB + gR) >> 7 == 256 + ((tB + tR + 8192) >> 14) EXACTLY for all 65,536
// chroma pairs, tools/gen_ycc_g_tables.py).  With s = eB + eR:  G' = Y' + (s >> 23) - 256 is one LEA.HI, and B' = Y' + eB,
// R' = Y' + eR are plain adds whose LOW 16 BITS are the result (+ 256): three adds and one shift-add instead of six instructions,
// and only one of them on the ALU pipe instead of four.
// RV_YCC16 = 0 (round 1).  Entry: bits 22..31 = f + 256, bits 0..21 = t / 2 modulo 2^22 (+ 4096 + 2^21 in the R table):
// ((eB + eR) << 10) >> 23 == 256 + ((tB + tR + 8192) >> 14), e >> 22 == f + 256.
// Built on the host (rv_b200.cu: build_ycc_table) with the same integer formulas; checked over all 2^24 colours by the CPU suite.
#ifndef RV_YCC16
#define RV_YCC16 1
#endif
struct YccTabs { uint32_t e[1024]; };
__device__ YccTabs g_ycc;


// A.1 luminance straight from the packed pixel word (B, G, R, x): Y = (1868 B + 9617 G + 4899 R + 8192) >> 14 as two
// chained 16-bit x 8-bit dot products (IDP.2A.LO takes bytes 0,1, IDP.2A.HI bytes 2,3; the x byte meets a zero coefficient)
__device__ __forceinline__ uint32_t luma_y(uint32_t px)
{
    constexpr uint32_t CBG = 1868u | (9617u << 16), CR0 = 4899u;
    return __dp2a_hi(CR0, px, __dp2a_lo(CBG, px, 8192u)) >> 14;
}

// The same sum scaled by four, not shifted: acc = 4 (1868 B + 9617 G + 4899 R) + 32768, so that Y is exactly the UPPER HALF-WORD
// of acc ((x + 8192) >> 14 == (4 x + 32768) >> 16; acc < 2^24).  A consumer can then take Y straight out of acc with another
// 16-bit x 8-bit dot product (IDP.2A.HI multiplies the upper half-word by a byte and adds a base): table addresses that are
// base + 4 Y or base - 4 Y cost one FMA-pipe instruction and the shift that would isolate Y never happens.
__device__ __forceinline__ uint32_t luma_acc16(uint32_t px)
{
    constexpr uint32_t CBG = 7472u | (38468u << 16), CR0 = 19596u;
    return __dp2a_hi(CR0, px, __dp2a_lo(CBG, px, 32768u));
}

// A.1 forward
__device__ __forceinline__ void ycrcb_fwd(int B, int G, int R, int &Y, int &Cr, int &Cb)
{
    Y = (4899 * R + 9617 * G + 1868 * B + 8192) >> 14;
    Cr = sat8(((R - Y) * 11682 + ((128 << 14) + 8192)) >> 14);
    Cb = sat8(((B - Y) * 9241 + ((128 << 14) + 8192)) >> 14);
}
// A.1 inverse
__device__ __forceinline__ void ycrcb_inv(int Y, int Cr, int Cb, int &B, int &G, int &R)
{
    const int cb = Cb - 128, cr = Cr - 128;
    B = sat8(Y + ((cb * 29049 + 8192) >> 14));
    G = sat8(Y + ((cb * -5636 + cr * -11698 + 8192) >> 14));
    R = sat8(Y + ((cr * 22987 + 8192) >> 14));
}
// A.5
__device__ __forceinline__ int gray_of(int B, int G, int R) { return (3735 * B + 19235 * G + 9798 * R + 16384) >> 15; }

// A.2 forward, luminance only, for the histogram pass: the three gamma look-ups and the Y row of the matrix folded into
// pre-multiplied tables (pm[0][R] = 871 g8[R], pm[1][G] = 2929 g8[G], pm[2][B] = 296 g8[B] + 2048) and the cube-root
// look-up folded with the L formula (lq[i] = (296 cb[i] - 1336934 + 16384) >> 15).  Same integers as lab_fwd's L; built on the
// host (rv_b200.cu: build_lab_hist_table) and checked over all 2^24 colours by the luminance-plane test.
struct LabHistTabs {
    uint32_t pm[3][256];
    uint8_t lq[2048];
};
static_assert(sizeof(LabHistTabs) % 16 == 0, "LabHistTabs must be 16-byte granular");
__device__ LabHistTabs g_labh;
__device__ __forceinline__ int lab_L_fast(const LabHistTabs *t, int B, int G, int R)
{
    return t->lq[(t->pm[0][R] + t->pm[1][G] + t->pm[2][B]) >> 12];
}

// The same from the packed pixel word (B, G, R, x): the three table addresses pm[c] + 4 v come from one byte dot product each (IDP.4A,
// FMA pipe) and the last look-up takes base + (sum >> 12) as one shift-and-add.  `tabs_s` = shared-memory address of the LabHistTabs.
__device__ __forceinline__ uint32_t lab_L_px(uint32_t tabs_s, uint32_t px)
{
    uint32_t r, g, b, l;
    asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(__dp4a(px, 0x00040000u, tabs_s)));                 // pm[0][R]
    asm("ld.shared.u32 %0, [%1+1024];" : "=r"(g) : "r"(__dp4a(px, 0x00000400u, tabs_s)));            // pm[1][G]
    asm("ld.shared.u32 %0, [%1+2048];" : "=r"(b) : "r"(__dp4a(px, 0x00000004u, tabs_s)));            // pm[2][B]
    asm("ld.shared.u8 %0, [%1+3072];" : "=r"(l) : "r"(tabs_s + ((r + g + b) >> 12)));                // lq[]
    return l;
}
static_assert(offsetof(LabHistTabs, pm) == 0 && offsetof(LabHistTabs, lq) == 3072, "lab_L_px addresses the tables by offset");

// Right shifts of the LAB fixed-point code.  RV_LAB_MULHI = 1 expresses them as the high half of a multiplication by 2^(32-n)
// (IMAD.HI, FMA pipe) instead of SHF (ALU pipe): in k_chain<LAB,*> the ALU pipe is the busier one.  floor semantics = arithmetic
// shift for negative values.  Experiment switch; see DESIGN.md for the measurement.
#ifndef RV_LAB_AB_IMAD
#define RV_LAB_AB_IMAD 0
#endif
#ifndef RV_LAB_MULHI
#define RV_LAB_MULHI 0
#endif
template <int N>
__device__ __forceinline__ int lab_shr(int x)
{
#if RV_LAB_MULHI
    return __mulhi(x, 1 << (32 - N));
#else
    return x >> N;
#endif
}

// A.2 forward, all three
__device__ __forceinline__ void lab_fwd(const LabTabs *t, int B, int G, int R, int &L, int &a, int &bb)
{
    const int r = t->g8[R], g = t->g8[G], b = t->g8[B];
    const int fX = t->cb[lab_shr<12>(1777 * r + 1541 * g + 778 * b + 2048)];
    const int fY = t->cb[lab_shr<12>(871 * r + 2929 * g + 296 * b + 2048)];
    const int fZ = t->cb[lab_shr<12>(73 * r + 448 * g + 3575 * b + 2048)];
    // over all 2^24 colours a stays in [42,226] and b in [20,223] (tests/test_oracle.py::test_lab_forward_ranges): no saturation needed
    L = lab_shr<15>(296 * fY - 1336934 + 16384);
    a = lab_shr<15>(500 * (fX - fY) + ((128 << 15) + 16384));
    bb = lab_shr<15>(200 * (fY - fZ) + ((128 << 15) + 16384));
}
// The same from the packed pixel word (B, G, R, x): the three gamma look-ups take their shared-memory addresses g8 + 2 B / 2 G / 2 R
// from one byte dot product each (IDP.4A, FMA pipe) instead of a byte extraction plus an add on the ALU pipe.  `tabs_s` is the
// shared-memory address of `t` (LabTabs::g8 sits at offset 0).
__device__ __forceinline__ void lab_fwd_px(const LabTabs *t, uint32_t tabs_s, uint32_t px, int &L, int &a, int &bb)
{
    uint32_t r, g, b;
    asm("ld.shared.u16 %0, [%1];" : "=r"(b) : "r"(__dp4a(px, 0x00000002u, tabs_s)));
    asm("ld.shared.u16 %0, [%1];" : "=r"(g) : "r"(__dp4a(px, 0x00000200u, tabs_s)));
    asm("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(__dp4a(px, 0x00020000u, tabs_s)));
    const int fX = t->cb[lab_shr<12>(1777 * (int)r + 1541 * (int)g + 778 * (int)b + 2048)];
    const int fY = t->cb[lab_shr<12>(871 * (int)r + 2929 * (int)g + 296 * (int)b + 2048)];
    const int fZ = t->cb[lab_shr<12>(73 * (int)r + 448 * (int)g + 3575 * (int)b + 2048)];
    L = lab_shr<15>(296 * fY - 1336934 + 16384);
    a = lab_shr<15>(500 * (fX - fY) + ((128 << 15) + 16384));
    bb = lab_shr<15>(200 * (fY - fZ) + ((128 << 15) + 16384));
}
static_assert(offsetof(LabTabs, g8) == 0, "lab_fwd_px addresses g8 at the start of LabTabs");

// Same results with every table index taken as the UPPER HALF-WORD of a scaled sum by IDP.2A.HI (FMA pipe) -- no shift and no address
// add on the ALU pipe, which is the binding one in k_chain<LAB,*>: the matrix rows are scaled by 16 (index = (16 (row . rgb + 2048)) >> 16,
// coefficients <= 57200 still fit 16 bits... they are ordinary IMAD immediates here; sums < 2^28), the L formula by 2.  L itself is returned
// inside accL (L = accL >> 16) so that the caller's quad-table address can be formed the same way.
__device__ __forceinline__ void lab_fwd_px2(uint32_t tabs_s, uint32_t px, uint32_t &accL, int &a, int &bb)
{
    uint32_t r, g, b, fX, fY, fZ;
    asm("ld.shared.u16 %0, [%1];" : "=r"(b) : "r"(__dp4a(px, 0x00000002u, tabs_s)));
    asm("ld.shared.u16 %0, [%1];" : "=r"(g) : "r"(__dp4a(px, 0x00000200u, tabs_s)));
    asm("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(__dp4a(px, 0x00020000u, tabs_s)));
    const uint32_t cb_s = tabs_s + (uint32_t)offsetof(LabTabs, cb);
    const uint32_t sx = 28432u * r + 24656u * g + 12448u * b + 32768u;        // 16 x (1777 r + 1541 g + 778 b + 2048)
    const uint32_t sy = 13936u * r + 46864u * g + 4736u * b + 32768u;         // 16 x (871 r + 2929 g + 296 b + 2048)
    const uint32_t sz = 1168u * r + 7168u * g + 57200u * b + 32768u;          // 16 x (73 r + 448 g + 3575 b + 2048)
    asm("ld.shared.u16 %0, [%1];" : "=r"(fX) : "r"(__dp2a_hi(sx, 0x02000000u, cb_s)));
    asm("ld.shared.u16 %0, [%1];" : "=r"(fY) : "r"(__dp2a_hi(sy, 0x02000000u, cb_s)));
    asm("ld.shared.u16 %0, [%1];" : "=r"(fZ) : "r"(__dp2a_hi(sz, 0x02000000u, cb_s)));
    accL = 592u * fY - 2641100u;                                              // 2 x (296 fY - 1336934 + 16384) >= 0
#if RV_LAB_AB_IMAD
    // the differences as two multiply-adds each (FMA pipe) instead of a subtraction (ALU pipe, the busier one) and a multiply-add
    a = lab_shr<15>((int)fY * -500 + ((int)fX * 500 + ((128 << 15) + 16384)));
    bb = lab_shr<15>((int)fZ * -200 + ((int)fY * 200 + ((128 << 15) + 16384)));
#else
    a = lab_shr<15>(500 * ((int)fX - (int)fY) + ((128 << 15) + 16384));
    bb = lab_shr<15>(200 * ((int)fY - (int)fZ) + ((128 << 15) + 16384));
#endif
}

// A.2 inverse arguments from the rounded blend result as it leaves the float unit: Lw = bits of (res + 1.5 * 2^23), i.e. the new L in
// the low half-word under a constant upper half; the two look-ups take base + 2 L from one IDP.2A.LO each.
__device__ __forceinline__ void lab_inv_args_w(uint32_t tabs_s, uint32_t Lw, int a, int b, int &y, int &ix, int &iz)
{
    uint32_t yy, fy;
    asm("ld.shared.u16 %0, [%1];" : "=r"(yy) : "r"(__dp2a_lo(Lw, 0x00000002u, tabs_s + (uint32_t)offsetof(LabTabs, yt))));
    asm("ld.shared.u16 %0, [%1];" : "=r"(fy) : "r"(__dp2a_lo(Lw, 0x00000002u, tabs_s + (uint32_t)offsetof(LabTabs, ft))));
    y = (int)yy;
#if RV_LAB_FT_BIAS
    // the ft table carries + 10484, the constants of adiv and bdiv ride inside the products and -floor(u / 512) is taken as
    // floor((511 - u) / 512): each argument is one IMAD and one shift-and-add (LEA.HI) instead of IMAD, shift and two adds
    ix = (int)fy + lab_shr<13>(a * 268435 + (128 - (4194 + RV_LAB_FT_BIAS) * 8192));
    iz = (int)fy + lab_shr<9>(b * -41943 + (511 - 16));
#else
    const int adiv = lab_shr<13>(5 * a * 53687 + 128) - 4194;
    const int bdiv = lab_shr<9>(b * 41943 + 16) - 10485 + 1;
    ix = (int)fy + adiv;
    iz = (int)fy - bdiv;
#endif
}

__device__ __forceinline__ int lab_xz(int i)
{
    const int lin = (i * 108) / 841 - 290;            // truncating division, as in C
    const int cub = (((i * i) >> 14) * i) >> 14;
    return i <= 3390 ? lin : cub;
}
// A.2 inverse.  lab_inv_args gives the two XZ arguments; when every argument in the warp is above 3390 (any pixel that is not
// nearly black) the caller uses CUBIC_ONLY = true and the linear branch of XZ with its division is never evaluated.
__device__ __forceinline__ void lab_inv_args(const LabTabs *t, int L, int a, int b, int &y, int &ix, int &iz)
{
    y = t->yt[L];
    const int fy = t->ft[L] - RV_LAB_FT_BIAS;
    const int adiv = lab_shr<13>(5 * a * 53687 + 128) - 4194;
    const int bdiv = lab_shr<9>(b * 41943 + 16) - 10485 + 1;
    ix = fy + adiv;
    iz = fy - bdiv;
}
#ifndef RV_LAB_LEA_IG
#define RV_LAB_LEA_IG 1                // measured: LAB k3 +3.2 %, LAB k5 +2.4 %, 4K LAB +3.4 % (profiles/r2_aj_lab_lea.txt)
#endif
template <bool CUBIC_ONLY>
__device__ __forceinline__ void lab_inv_tail(const LabTabs *t, int y, int ix, int iz, int &B, int &G, int &R)
{
    const int x = CUBIC_ONLY ? lab_shr<14>(lab_shr<14>(ix * ix) * ix) : lab_xz(ix);
    const int z = CUBIC_ONLY ? lab_shr<14>(lab_shr<14>(iz * iz) * iz) : lab_xz(iz);
#if RV_LAB_LEA_IG
    // clamp BEFORE the shift -- clamp(s >> 14, 0, 4095) == clamp(s, 0, 4095 << 14 | 0x3fff) >> 14 -- so that the shift and the table's
    // base address fuse into one LEA.HI: VIMNMX.RELU + LEA.HI per channel instead of SHF + VIMNMX.RELU + IADD
    constexpr int TOP = (4095 << 14) | 0x3fff;
    const uint32_t ro = (uint32_t)__vimin_s32_relu(12615 * x - 6296 * y - 2223 * z + 8192, TOP);
    const uint32_t go = (uint32_t)__vimin_s32_relu(-3773 * x + 7684 * y + 185 * z + 8192, TOP);
    const uint32_t bo = (uint32_t)__vimin_s32_relu(217 * x - 836 * y + 4715 * z + 8192, TOP);
    const uint32_t ig_s = smem_u32(t) + (uint32_t)offsetof(LabTabs, ig);
    uint32_t b8, g8v, r8;
    asm("ld.shared.u8 %0, [%1];" : "=r"(b8) : "r"(ig_s + (bo >> 14)));
    asm("ld.shared.u8 %0, [%1];" : "=r"(g8v) : "r"(ig_s + (go >> 14)));
    asm("ld.shared.u8 %0, [%1];" : "=r"(r8) : "r"(ig_s + (ro >> 14)));
    B = (int)b8; G = (int)g8v; R = (int)r8;
#else
    int ro = lab_shr<14>(12615 * x - 6296 * y - 2223 * z + 8192);
    int go = lab_shr<14>(-3773 * x + 7684 * y + 185 * z + 8192);
    int bo = lab_shr<14>(217 * x - 836 * y + 4715 * z + 8192);
    ro = __vimin_s32_relu(ro, 4095);                  // clamp to [0, 4095] in one VIMNMX.RELU each
    go = __vimin_s32_relu(go, 4095);
    bo = __vimin_s32_relu(bo, 4095);
    B = t->ig[bo];
    G = t->ig[go];
    R = t->ig[ro];
#endif
}

__device__ __forceinline__ void copy_lab_tabs(LabTabs *dst)
{
    const uint4 *s = reinterpret_cast<const uint4 *>(&g_lab);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(LabTabs) / 16); i += blockDim.x) d[i] = s[i];
}

}  // namespace rv
