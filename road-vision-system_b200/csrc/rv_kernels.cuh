/*
 * rv_kernels.cuh -- sm_100a kernels of the preprocessing chain (CLAHEDehaze -> MedianDerain); umbrella over
 * rv_common.cuh (geometry, TMA helpers), rv_colour.cuh (colour arithmetic + tables), rv_hist_lut.cuh (k_luma_hist,
 * k_build_lut*), rv_chain.cuh (k_chain, k_letterbox).
 *
 * Arithmetic follows OpenCV's 8-bit fixed-point paths exactly; the
 * reference only *calls* them: /root/reference/src/preprocess/ops/clahe_dehaze.py:19-30 and
 * ops/median_derain.py:14.  Kernels:
 *
 *   k_luma_hist   BGR -> luminance (Y of YCrCb via two byte dot products on the packed pixel word / L of LAB) +
 *                 per-tile 256-bin histograms (per-warp private sub-histograms in shared memory, four 12-byte
 *                 groups in flight per thread), optional gray min/max (low-contrast gate) and luma plane (tests)
 *   k_build_lut   clip + redistribute + prefix sum -> u8 LUT per tile (one warp per tile, shuffle scans) and the
 *                 four-LUT "quad" tables used by the interpolation
 *   k_chain       per 120x48 output tile: TMA box load (cp.async.bulk.tensor + mbarrier), forward colour conversion,
 *                 bilinear four-LUT blend in float32 without FMA contraction, inverse colour conversion, packed
 *                 saturation into u16x2 planes, k x k median (generated selection networks, half the compare-
 *                 exchanges on the FMA pipe, two output rows per task for 5x5), coalesced store; optionally the
 *                 detector-input letterbox for integer down-scales
 *   k_letterbox   detector-input stage, general form (cv2.resize INTER_LINEAR arithmetic, RGB fp16 NCHW)
 */
#pragma once
#include "rv_common.cuh"
#include "rv_colour.cuh"
#include "rv_hist_lut.cuh"
#include "rv_chain.cuh"
