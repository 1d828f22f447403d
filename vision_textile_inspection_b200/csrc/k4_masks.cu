// K4 -- mask-coefficient x prototype contraction fused with sigmoid, box crop, 4x bilinear upsample, > 0.5 threshold
// AND the nearest-resize mask statistics of the measure stage, so full-resolution masks are never materialised.
//
// Replaces: ultralytics ops.process_mask / crop_mask / F.interpolate / gt_(0.5)
// (variant A), reached from /root/reference/measurement.py:208-210 and consumed at :74-75, and then
//   measurement.py:78-81   cv2.resize(mask, (w,h), INTER_NEAREST) > 0        -> per-axis multiplicity LUTs
//   measurement.py:160-185 union of fabric masks + lower envelope            -> atomicMax per letterbox column
//   (Utils/check_stitch_distance.py:238-251 upper envelope                   -> atomicMin)
//   measurement.py:304-314 cv2.moments m00/m10/m01 + occupied column range   -> exact int64 sums
// Oracle: oracle/ultra_ref.py process_mask (real torch ops) + oracle/measure_port.py; parity bar IoU >= 0.999.
//
// v4: the work is a flat list of UNITS written by K3, one per (kept detection, VTI_K4_UR x VTI_K4_UC block of
// interpolation cells inside the detection's crop window).  The box crop makes the contraction sparse -- cfg2 needs
// ~1 % of the dense N x 32 x ph x pw products -- so the kernel is organised around the crop, not around a dense GEMM:
// one warp per unit, grid-stride over the list, no block barriers, no shared-memory staging of the prototypes
// (each unit reads its (UR+1) x (UC+1) x 32 prototype corner values straight from global/L2, 32 independent
// coalesced loads per lane).  Per unit:
//   (1) one lane per corner pixel: 32-term contraction + sigmoid + crop                       -> per-warp s_c
//   (2) one lane per interpolation cell (the 4x4 output pixels between four prototype pixels).  A cell whose four
//       corners are all > 0.5 (+margin) is entirely set, all < 0.5 (-margin) entirely clear -- bilinear weights are a
//       convex combination -- and its contribution to m00/m10/m01, the column extent and the envelope is closed-form
//       from prefix sums of the multiplicity tables; only boundary cells evaluate their 16 pixels, and those are
//       shared out over the warp two cells at a time, one pixel per lane (the mask edge crosses a unit as a line, so
//       few lanes hold one).
//   (3) warp reduction (three REDUX on 32-bit per-unit sums), one set of 64-bit global atomics per unit.
// Optional bit-packed mask export: cells OR their bits into the (pre-zeroed) global mask words.
// Mask variant B (vti_params.mask_variant = 1, newer Ultralytics: SURVEY 8a U6) is the same kernel in another threshold
// domain (template parameter VB): the corner values are the raw logits (no sigmoid), cropped-out corners are 0, the
// threshold is > 0.0; K5 then drops the detections whose letterbox mask stayed empty (flag VTI_F_LB_MASK never set).
// Bound: the prototype read (128 B per prototype pixel touched); algorithmic bytes 128*ph*pw per frame.
#include <climits>
#include <cstdlib>

#include <cuda.h>

#include "vti_internal.h"

namespace {

#ifndef VTI_K4_THREADS
#define VTI_K4_THREADS 256
#endif
constexpr int K4_THREADS = VTI_K4_THREADS;
constexpr int NWARP = K4_THREADS / 32;
constexpr int UR = VTI_K4_UR, UC = VTI_K4_UC;
constexpr int SCW = UC + 1;                 // corner columns per unit
constexpr float MARGIN = 1e-5f;             // >> fp32 rounding of the lerp; cells within the margin are evaluated per pixel
constexpr float MARGIN_B = 1e-3f;
constexpr int MAX_DENSE_DETS = 1024;         // = the max_det ceiling of vti_create           // variant B: logits (|values| up to ~30) instead of probabilities

struct K4Args {
    const float* proto;         // [B][32][ph][pw]
    vti_det* dets;              // [B][max_det]
    const float* det_coef;      // [B][max_det][32]
    const uint4* units;
    const int32_t* unit_count;
    AxisLut ly, lx;
    int32_t* env;               // [B][LW] frame-row envelope per letterbox column
    uint32_t* masks;            // optional [B][max_det][LH][LW/32]
    int LH, LW, ph, pw, max_det;
    int upper;                  // envelope mode
};

__global__ void k4_zero_masks_kernel(uint32_t* masks, const int32_t* counts, int max_det, size_t words_per_mask) {
    const int k = blockIdx.x, b = blockIdx.y;
    if (k >= counts[b]) return;
    uint4* p = reinterpret_cast<uint4*>(masks + ((size_t)b * max_det + k) * words_per_mask);
    const size_t n4 = words_per_mask / 4;
    for (size_t i = threadIdx.x; i < n4; i += blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Phases 2 and 3 of a unit: classify the interpolation cells from the corner values `sc` (pitch SCW), accumulate the
// nearest-resize statistics, flush the unit's fabric envelope and reduce across the warp.
template <bool EXPORT, bool VB>
__device__ __forceinline__ void unit_cells(const K4Args& a, int b, int k, vti_det* __restrict__ det, bool fabric, int R0,
                                           int C0, int nr, int ncw, const float* sc, int* envw, int lane) {
    const int ex0 = 4 * C0 - 2;                            // first output column of the unit
    constexpr float THR = VB ? 0.0f : 0.5f, MRG = VB ? MARGIN_B : MARGIN;
    bool any = false;                                      // a letterbox pixel of this detection is set
    // (2) cells
    const int ncc = ncw - 1;
    const float inv_ncc = 1.0f / (float)ncc;
    // per-unit sums fit 32 bits (<= 1024 pixels x multiplicity^2 x frame size; vti_create refuses geometries that do not),
    // so the warp reduction is three REDUX instructions instead of thirty 64-bit shuffle-adds
    unsigned m00 = 0u, m10 = 0u, m01 = 0u;
    int cmin = INT_MAX, cmax = -1;
    uint32_t* __restrict__ mrow = EXPORT ? a.masks + ((size_t)b * a.max_det + k) * a.LH * (a.LW / 32) : nullptr;
    const int ncells = (nr - 1) * ncc;
    for (int i0 = 0; i0 < ncells; i0 += 32) {              // warp-uniform trip count: the boundary cells are shared out below
        const int i = i0 + lane;
        const int r = (int)(((float)i + 0.5f) * inv_ncc), c = i - r * ncc;
        // cell (R, C) covers output rows 4R-2 .. 4R+1 and cols 4C-2 .. 4C+1, clipped to the image
        const int cy_first = 4 * (R0 + r) - 2, cx_first = 4 * (C0 + c) - 2;
        const int ya = max(cy_first, 0), yb = min(cy_first + 3, a.LH - 1);
        const int xa = max(cx_first, 0), xb = min(cx_first + 3, a.LW - 1);
        const bool valid = i < ncells && ya <= yb && xa <= xb;
        float c00 = 0.0f, c01 = 0.0f, c10 = 0.0f, c11 = 0.0f;
        if (valid) {
            c00 = sc[r * SCW + c]; c01 = sc[r * SCW + c + 1];
            c10 = sc[(r + 1) * SCW + c]; c11 = sc[(r + 1) * SCW + c + 1];
        }
        const float vmin = fminf(fminf(c00, c01), fminf(c10, c11));
        const float vmax = fmaxf(fmaxf(c00, c01), fmaxf(c10, c11));
        const bool full = valid && vmin > THR + MRG;
        if (full) {
            any = true;
            // fully set: closed form from the prefix sums of the nearest-resize multiplicity tables
            const int f_cy = a.ly.pc[yb + 1] - a.ly.pc[ya], f_sy = a.ly.ps[yb + 1] - a.ly.ps[ya];
            const int f_cx = a.lx.pc[xb + 1] - a.lx.pc[xa], f_sx = a.lx.ps[xb + 1] - a.lx.ps[xa];
            m00 += f_cy * f_cx; m10 += f_cy * f_sx; m01 += f_sy * f_cx;
            if (f_cy > 0 && f_cx > 0) {
                cmin = min(cmin, a.lx.next_first[xa]); cmax = max(cmax, a.lx.prev_last[xb]);
                if (fabric) {
                    const int f_env = a.upper ? a.ly.next_first[ya] : a.ly.prev_last[yb];
                    for (int X = xa; X <= xb; ++X)
                        if (a.lx.cnt[X] > 0) {
                            if (a.upper) atomicMin(&envw[X - ex0], f_env);
                            else atomicMax(&envw[X - ex0], f_env);
                        }
                }
            }
            if (EXPORT) {
                const unsigned long long bits = ((1ull << (xb - xa + 1)) - 1ull) << (xa & 31);
                for (int Y = ya; Y <= yb; ++Y) {
                    uint32_t* wp = mrow + (size_t)Y * (a.LW / 32) + (xa >> 5);
                    atomicOr(wp, (unsigned)bits);
                    if ((unsigned)(bits >> 32)) atomicOr(wp + 1, (unsigned)(bits >> 32));
                }
            }
        }
        // Boundary cells evaluate their 16 pixels (torch upsample_bilinear2d, align_corners=False, scale 1/4).  Few lanes
        // of a warp hold one -- the mask edge crosses a unit as a line -- so the warp takes them two at a time, one
        // pixel per lane, instead of every holder looping over its 16 pixels with the other lanes idle.
        unsigned bm = __ballot_sync(0xffffffffu, valid && !full && vmax >= THR - MRG);
        while (bm) {
            const int s0 = __ffs(bm) - 1;
            bm &= bm - 1;
            const bool pair = bm != 0u;
            const int s1 = pair ? __ffs(bm) - 1 : s0;
            bm &= bm - 1;
            const int src = lane < 16 ? s0 : s1;
            const float q00 = __shfl_sync(0xffffffffu, c00, src), q01 = __shfl_sync(0xffffffffu, c01, src);
            const float q10 = __shfl_sync(0xffffffffu, c10, src), q11 = __shfl_sync(0xffffffffu, c11, src);
            const int rc = __shfl_sync(0xffffffffu, (r << 16) | c, src);
            const int py = (lane >> 2) & 3, px = lane & 3;
            const int Y = 4 * (R0 + (rc >> 16)) - 2 + py, X = 4 * (C0 + (rc & 0xFFFF)) - 2 + px;
            if ((lane < 16 || pair) && Y >= 0 && Y < a.LH && X >= 0 && X < a.LW) {
                const float ly1 = (float)(2 * py + 1) * 0.125f, lx1 = (float)(2 * px + 1) * 0.125f;
                const float top = (1.0f - lx1) * q00 + lx1 * q01;
                const float bot = (1.0f - lx1) * q10 + lx1 * q11;
                const float v = (1.0f - ly1) * top + ly1 * bot;
                if (v > THR) {
                    any = true;
                    const int cY = a.ly.cnt[Y], sY = a.ly.sum[Y], cX = a.lx.cnt[X];
                    m00 += cY * cX; m10 += cY * a.lx.sum[X]; m01 += sY * cX;
                    if (cY > 0 && cX > 0) {
                        cmin = min(cmin, a.lx.first[X]);
                        cmax = max(cmax, a.lx.last[X]);
                        if (fabric) {
                            if (a.upper) atomicMin(&envw[X - ex0], a.ly.first[Y]);
                            else atomicMax(&envw[X - ex0], a.ly.last[Y]);
                        }
                    }
                    if (EXPORT) atomicOr(mrow + (size_t)Y * (a.LW / 32) + (X >> 5), 1u << (X & 31));
                }
            }
        }
    }
    if (fabric) {                                       // one RED per touched column and unit
        __syncwarp();
        int32_t* __restrict__ env = a.env + (size_t)b * a.LW;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int ev = envw[lane + 32 * hh], X = ex0 + lane + 32 * hh;
            if (a.upper) { if (ev != INT_MAX) atomicMin(env + X, ev); }
            else { if (ev >= 0) atomicMax(env + X, ev); }
        }
    }
    // (3) warp reduction, one set of global atomics per unit
    if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr(&det->flags, VTI_F_LB_MASK);
    m00 = __reduce_add_sync(0xffffffffu, m00);
    if (m00 > 0u) {
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        cmin = __reduce_min_sync(0xffffffffu, cmin);
        cmax = __reduce_max_sync(0xffffffffu, cmax);
        if (lane == 0) {
            atomicAdd((unsigned long long*)&det->m00, (unsigned long long)m00);
            atomicAdd((unsigned long long*)&det->m10, (unsigned long long)m10);
            atomicAdd((unsigned long long*)&det->m01, (unsigned long long)m01);
            atomicMin(&det->col_min, cmin);
            atomicMax(&det->col_max, cmax);
        }
    }
}

template <bool EXPORT, bool VB>
__global__ void __launch_bounds__(K4_THREADS, 1024 / K4_THREADS) k4_units_kernel(const K4Args a) {
    __shared__ float s_c[NWARP][(UR + 1) * SCW];
    __shared__ __align__(16) float s_coef[NWARP][VTI_NM];   // registers go to the 32 in-flight prototype loads
    __shared__ int s_envw[NWARP][4 * UC];                   // fabric envelope of the unit's columns (one RED each)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = *a.unit_count;
    const int nwarps = gridDim.x * NWARP;
    float* sc = s_c[warp];
    const size_t plane = (size_t)a.ph * a.pw;
    const unsigned long long plane_bytes = plane * sizeof(float);

    for (int u = blockIdx.x * NWARP + warp; u < total; u += nwarps) {
        const uint4 un = __ldg(a.units + u);
        const int b = (int)(un.x & 0x7FFFu), k = (int)(un.x >> 16), br = (int)(un.y & 0xFFFFu), bc = (int)(un.y >> 16);
        const bool fabric = (un.x & 0x8000u) != 0u;
        vti_det* __restrict__ det = a.dets + (size_t)b * a.max_det + k;
        VtiWindow w;                                        // the crop window travels in the unit record
        w.cx_lo = (int)(un.z & 0xFFFFu); w.cy_lo = (int)(un.z >> 16);
        w.cx_hi = (int)(un.w & 0xFFFFu); w.cy_hi = (int)(un.w >> 16);
        // cells of this unit (global cell coordinates) and their corner rectangle
        const int R0 = w.cy_lo + br * UR, R1 = min(R0 + UR - 1, w.cy_hi + 1);
        const int C0 = w.cx_lo + bc * UC, C1 = min(C0 + UC - 1, w.cx_hi + 1);
        const int nr = R1 - R0 + 2, ncw = C1 - C0 + 2;
        const float inv_ncw = 1.0f / (float)ncw;
        __syncwarp();
        s_coef[warp][lane] = __ldg(a.det_coef + ((size_t)b * a.max_det + k) * VTI_NM + lane);
        if (fabric) { s_envw[warp][lane] = a.upper ? INT_MAX : -1; s_envw[warp][lane + 32] = a.upper ? INT_MAX : -1; }
        const float* __restrict__ proto = a.proto + (size_t)b * VTI_NM * plane;
        __syncwarp();
        // (1) logits -> sigmoid -> crop over the corner rectangle (replicate-clamped at the plane border)
        for (int i = lane; i < nr * ncw; i += 32) {
            const int r = (int)(((float)i + 0.5f) * inv_ncw), c = i - r * ncw;
            const int py = min(max(R0 - 1 + r, 0), a.ph - 1), px = min(max(C0 - 1 + c, 0), a.pw - 1);
            float v = 0.0f;
            if (py >= w.cy_lo && py <= w.cy_hi && px >= w.cx_lo && px <= w.cx_hi) {
                // walk the 32 channel planes with one 64-bit add per load (the compiler's own q * plane indexing costs
                // a wide multiply + two LEAs per load)
                const char* src = reinterpret_cast<const char*>(proto + (size_t)py * a.pw + px);
                float pv[VTI_NM];
#pragma unroll
                for (int q = 0; q < VTI_NM; ++q) {
                    pv[q] = __ldg(reinterpret_cast<const float*>(src));
                    asm volatile("add.u64 %0, %0, %1;" : "+l"(src) : "l"(plane_bytes));
                }
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < VTI_NM; q += 4) {
                    const float4 cf = *reinterpret_cast<const float4*>(&s_coef[warp][q]);
                    acc = fmaf(cf.x, pv[q], acc); acc = fmaf(cf.y, pv[q + 1], acc);
                    acc = fmaf(cf.z, pv[q + 2], acc); acc = fmaf(cf.w, pv[q + 3], acc);
                }
                v = VB ? acc : 1.0f / (1.0f + expf(-acc));
            }
            sc[r * SCW + c] = v;
        }
        __syncwarp();
        unit_cells<EXPORT, VB>(a, b, k, det, fabric, R0, C0, nr, ncw, sc, s_envw[warp], lane);
    }
}


// ====================================================================================================================
// TMA form of the unit kernel (opt-in, kept as a measured experiment: see vti_launch_k4).  A unit's prototype corner block is a 3-D box {TB_W columns, UR+1 rows, 32 channels} of
// the [B*32][ph][pw] prototype tensor: ONE cp.async.bulk.tensor (TMA) brings it into shared memory, where the LDG
// form needs 32 loads + 64 address instructions per lane and iteration.  Each warp double-buffers: the box of its next
// unit is in flight (mbarrier with the byte count) while it contracts, classifies and reduces the current one, so the
// L2/HBM latency that bounded the LDG form (issue slots 47-60 % used) is off the critical path.  Out-of-plane parts of
// a box are zero-filled by the TMA unit and never read: the replicate clamp of the upsample is applied to the
// shared-memory index.
// ====================================================================================================================
constexpr int TB_W = 24;                           // box columns: the box starts at the 16-byte-aligned column at or below
                                                   // C0 - 1 (TMA needs 16-byte-aligned row starts), so 3 + UC + 1 = 20 <= 24
constexpr int TB_H = UR + 1;
constexpr int TB_FLOATS = VTI_NM * TB_H * TB_W;
constexpr int TB_BYTES = TB_FLOATS * 4;            // 15 360 with UR = 4
constexpr int T_WARPS = 3;                         // warps per CTA: 3 x 2 buffers x 15 KB = 90 KB, two CTAs per SM
static_assert(TB_BYTES % 128 == 0, "TMA destination alignment");

__device__ __forceinline__ unsigned k4_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct UnitGeo {
    int b, k, R0, C0, nr, ncw;
    bool fabric;
    VtiWindow w;
};
__device__ __forceinline__ UnitGeo unit_geo(const uint4 un) {
    UnitGeo g;
    g.b = (int)(un.x & 0x7FFFu); g.k = (int)(un.x >> 16);
    g.fabric = (un.x & 0x8000u) != 0u;
    const int br = (int)(un.y & 0xFFFFu), bc = (int)(un.y >> 16);
    g.w.cx_lo = (int)(un.z & 0xFFFFu); g.w.cy_lo = (int)(un.z >> 16);
    g.w.cx_hi = (int)(un.w & 0xFFFFu); g.w.cy_hi = (int)(un.w >> 16);
    g.w.empty = false;
    g.R0 = g.w.cy_lo + br * UR; g.C0 = g.w.cx_lo + bc * UC;
    const int R1 = min(g.R0 + UR - 1, g.w.cy_hi + 1), C1 = min(g.C0 + UC - 1, g.w.cx_hi + 1);
    g.nr = R1 - g.R0 + 2; g.ncw = C1 - g.C0 + 2;
    return g;
}

template <bool EXPORT>
__global__ void __launch_bounds__(T_WARPS * 32, 2) k4_tma_kernel(const __grid_constant__ CUtensorMap tmap, const K4Args a) {
    extern __shared__ __align__(128) float s_box[];                     // [T_WARPS][2][TB_FLOATS]
    __shared__ __align__(8) unsigned long long s_bar[T_WARPS][2];
    __shared__ float s_c[T_WARPS][(UR + 1) * SCW];
    __shared__ __align__(16) float s_coef[T_WARPS][VTI_NM];
    __shared__ int s_envw[T_WARPS][4 * UC];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = *a.unit_count;
    const int nwarps = gridDim.x * T_WARPS;
    int u = blockIdx.x * T_WARPS + warp;
    if (u >= total) return;                                             // (no block-wide barrier anywhere below)
    float* buf0 = s_box + (size_t)warp * 2 * TB_FLOATS;
    float* sc = s_c[warp];
    const unsigned long long tmap_addr = reinterpret_cast<unsigned long long>(&tmap);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"(k4_smem_u32(&s_bar[warp][0])) : "memory");
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"(k4_smem_u32(&s_bar[warp][1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](const UnitGeo& g, int slot) {                      // lane 0: arm the barrier, start the box copy
        const unsigned bar = k4_smem_u32(&s_bar[warp][slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((unsigned)TB_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(k4_smem_u32(buf0 + (size_t)slot * TB_FLOATS)), "l"(tmap_addr), "r"((g.C0 - 1) & ~3), "r"(g.R0 - 1),
                        "r"(g.b * VTI_NM), "r"(bar) : "memory");
    };
    UnitGeo g = unit_geo(__ldg(a.units + u));
    if (lane == 0) issue(g, 0);
    unsigned parity0 = 0u, parity1 = 0u;
    for (int it = 0;; ++it) {
        const int cur = it & 1;
        const int un = u + nwarps;
        const bool more = un < total;
        UnitGeo gn = g;
        if (more) {
            gn = unit_geo(__ldg(a.units + un));
            if (lane == 0) issue(gn, cur ^ 1);                          // that buffer was consumed one iteration ago
        }
        // this unit's coefficients and envelope scratch while its box lands
        vti_det* __restrict__ det = a.dets + (size_t)g.b * a.max_det + g.k;
        s_coef[warp][lane] = __ldg(a.det_coef + ((size_t)g.b * a.max_det + g.k) * VTI_NM + lane);
        if (g.fabric) { s_envw[warp][lane] = a.upper ? INT_MAX : -1; s_envw[warp][lane + 32] = a.upper ? INT_MAX : -1; }
        {
            const unsigned bar = k4_smem_u32(&s_bar[warp][cur]), par = cur ? parity1 : parity0;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "W_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\t"
                "bra W_%=;\n\t"
                "D_%=:\n\t}"
                :: "r"(bar), "r"(par) : "memory");
            if (cur) parity1 ^= 1u; else parity0 ^= 1u;
        }
        __syncwarp();
        // (1) logits -> sigmoid -> crop over the corner rectangle, prototype values from the box in shared memory
        const float* box = buf0 + (size_t)cur * TB_FLOATS;
        const float inv_ncw = 1.0f / (float)g.ncw;
        for (int i = lane; i < g.nr * g.ncw; i += 32) {
            const int r = (int)(((float)i + 0.5f) * inv_ncw), c = i - r * g.ncw;
            const int py = min(max(g.R0 - 1 + r, 0), a.ph - 1), px = min(max(g.C0 - 1 + c, 0), a.pw - 1);
            float v = 0.0f;
            if (py >= g.w.cy_lo && py <= g.w.cy_hi && px >= g.w.cx_lo && px <= g.w.cx_hi) {
                const float* src = box + (py - (g.R0 - 1)) * TB_W + (px - ((g.C0 - 1) & ~3));   // replicate clamp on the index
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < VTI_NM; q += 4) {
                    const float4 cf = *reinterpret_cast<const float4*>(&s_coef[warp][q]);
                    acc = fmaf(cf.x, src[(q + 0) * TB_H * TB_W], acc); acc = fmaf(cf.y, src[(q + 1) * TB_H * TB_W], acc);
                    acc = fmaf(cf.z, src[(q + 2) * TB_H * TB_W], acc); acc = fmaf(cf.w, src[(q + 3) * TB_H * TB_W], acc);
                }
                v = 1.0f / (1.0f + expf(-acc));
            }
            sc[r * SCW + c] = v;
        }
        __syncwarp();
        unit_cells<EXPORT, false>(a, g.b, g.k, det, g.fabric, g.R0, g.C0, g.nr, g.ncw, sc, s_envw[warp], lane);
        __syncwarp();                                                   // everyone is done with box[cur], s_c, s_coef
        if (!more) break;
        g = gn; u = un;
    }
}


// ====================================================================================================================
// Dense tile form on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// The north star names the N x 32 @ 32 x (ph*pw) contraction as the one tensor-core candidate.  On the configured
// workloads the box crop makes it 0.1-0.8 % dense and the unit form above wins by a wide margin; this form
// is for scenes where MANY kept boxes overlap the same prototype pixels (dispatch: vti_params.k4_dense).  Work item =
// (frame, tile of DT_R x DT_C interpolation cells = 7 x 17 = 119 corner prototype pixels, padded to M = 128):
//   * the tile's detections (those whose crop window touches the tile's cells) are listed,
//   * A = the tile's prototype values [128 px][32 ch], B = the listed detections' coefficients [N][32], both K-major
//     and split into tf32 hi + lo parts in shared memory: L = A_hi B_hi + A_lo B_hi + A_hi B_lo, three
//     accumulating passes of 4 k-steps each (fp32 accumulate in TMEM) -- a single tf32 pass flips mask pixels on ~2 % of
//     the instances, the three-term split is at fp32 rounding level,
//   * tcgen05.ld brings D[128 px][DN dets] back, thread t owning TMEM lane t = corner pixel t; sigmoid (or the raw logit,
//     variant B) + box crop go to shared memory as one corner array per detection,
//   * the cells of every detection are then classified and reduced by exactly the code of the unit form (unit_cells).
// The prototype tile is read ONCE per tile instead of once per (detection, unit).
// ====================================================================================================================
constexpr int DT_R = 6, DT_C = UC;                  // cells per tile: 6 x 16  (corner rows x cols: 7 x 17 = 119 <= 128)
constexpr int DT_CR = DT_R + 1, DT_CC = DT_C + 1;
constexpr int DT_M = 128;                           // MMA M (TMEM lanes); rows >= 119 are zero padding
constexpr int DN = 64;                              // detections per MMA pass = MMA N = TMEM columns
constexpr int DENSE_THREADS = 256;
constexpr int DA_BYTES = DT_M * VTI_NM * 4;         // 16 KB per A part
constexpr int DB_BYTES = DN * VTI_NM * 4;           // 8 KB per B part
constexpr size_t K4_DENSE_SMEM = 2 * DA_BYTES + 2 * DB_BYTES + (size_t)DN * DT_M * 4 + 1024;   // + s_c, + alignment slack
static_assert(DT_CR * DT_CC <= DT_M, "tile corners must fit the MMA M");

__device__ __forceinline__ unsigned tf32_rna(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (all >> 4), version 1 (sm_100)
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Both operands K-major, no swizzle: element (row, k) at (row / 8) * 1024 + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4
// [core matrix = 8 rows x 4 k = 128 contiguous bytes; the two core matrices of a k-step lie LBO = 128 B apart, 8-row groups
// SBO = 1024 B apart].  (tools/probe/tcgen05_probe.cu: with this layout kind::tf32 reproduces an integer test product
// exactly; an MN-major A -- the prototype plane's native order -- returned zeros on this toolkit, hence the transposing
// staging pass, which the hi / lo split needs anyway.)
__device__ __forceinline__ int da_off(int m, int k) { return (m >> 3) * 1024 + (k >> 2) * 128 + (m & 7) * 16 + (k & 3) * 4; }
__device__ __forceinline__ int db_off(int n, int k) { return (n >> 3) * 1024 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4; }

template <bool EXPORT, bool VB>
__global__ void __launch_bounds__(DENSE_THREADS, 2) k4_dense_kernel(const K4Args a, const int32_t* __restrict__ counts,
                                                                    const int32_t* __restrict__ dense_flag, int all_dets) {
    extern __shared__ __align__(16) unsigned char s_raw8[];
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ unsigned s_tmem;
    __shared__ short s_list[MAX_DENSE_DETS];
    __shared__ int s_nt;
    __shared__ int s_envw[DENSE_THREADS / 32][4 * UC];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    if (dense_flag && !dense_flag[b]) return;
    const int ntx = (a.pw + 1 + DT_C - 1) / DT_C;
    const int Ty = blockIdx.x / ntx, Tx = blockIdx.x - Ty * ntx;
    const int R0 = Ty * DT_R, C0 = Tx * DT_C;                              // first cell of the tile
    const int R1 = min(R0 + DT_R - 1, a.ph), C1 = min(C0 + DT_C - 1, a.pw); // last cell (cells exist for 0..ph / 0..pw)
    const int nr = R1 - R0 + 2, ncw = C1 - C0 + 2;                          // corner rows / cols actually used
    const int n = min(counts[b], MAX_DENSE_DETS);
    vti_det* __restrict__ dets = a.dets + (size_t)b * a.max_det;

    // ---- the tile's detections
    if (tid == 0) s_nt = 0;
    __syncthreads();
    for (int k = tid; k < n; k += DENSE_THREADS) {
        const unsigned f = dets[k].flags;
        const bool wanted = all_dets || ((f & VTI_F_IN_ROI) && (f & (VTI_F_STITCH | VTI_F_FABRIC)));
        const VtiWindow w = vti_det_window(dets[k].box_lb, a.ph, a.pw);
        // the detection's non-zero cells are rows [cy_lo, cy_hi + 1], cols [cx_lo, cx_hi + 1]
        if (wanted && !w.empty && w.cy_lo <= R1 && w.cy_hi + 1 >= R0 && w.cx_lo <= C1 && w.cx_hi + 1 >= C0)
            s_list[atomicAdd(&s_nt, 1)] = (short)k;
    }
    __syncthreads();
    const int nt = s_nt;
    if (nt == 0) return;

    unsigned char* sA_hi = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_raw8) + 1023) & ~uintptr_t(1023));
    unsigned char* sA_lo = sA_hi + DA_BYTES;
    unsigned char* sB_hi = sA_lo + DA_BYTES;
    unsigned char* sB_lo = sB_hi + DB_BYTES;
    float* s_c = reinterpret_cast<float*>(sB_lo + DB_BYTES);               // [DN][DT_M] corner values per detection

    // ---- TMEM allocation (one warp), barrier init
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(k4_smem_u32(&s_tmem)), "n"(DN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(k4_smem_u32(&s_mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- A: the tile's prototype corner pixels (replicate-clamped like the unit form), tf32 hi + lo
    const float* __restrict__ proto = a.proto + (size_t)b * VTI_NM * a.ph * a.pw;
    const size_t plane = (size_t)a.ph * a.pw;
    for (int i = tid; i < DT_M * VTI_NM; i += DENSE_THREADS) {
        const int q = i / DT_M, m = i - q * DT_M;                          // consecutive threads: consecutive pixels
        const int r = m / DT_CC, c = m - r * DT_CC;
        float v = 0.0f;
        if (r < nr && c < ncw) {
            const int py = min(max(R0 - 1 + r, 0), a.ph - 1), px = min(max(C0 - 1 + c, 0), a.pw - 1);
            v = __ldg(proto + q * plane + (size_t)py * a.pw + px);
        }
        const unsigned hi = tf32_rna(v);
        const unsigned lo = tf32_rna(__fsub_rn(v, __uint_as_float(hi)));
        *reinterpret_cast<unsigned*>(sA_hi + da_off(m, q)) = hi;
        *reinterpret_cast<unsigned*>(sA_lo + da_off(m, q)) = lo;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = s_tmem;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = DN, M = 128
    constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((unsigned)(DN >> 3) << 17) |
                               ((unsigned)(DT_M >> 4) << 24);
    unsigned phase = 0u;
    const int my_m = tid & 127;                                            // TMEM lane = corner pixel of this thread
    const int mr = my_m / DT_CC, mc = my_m - mr * DT_CC;
    const int mpy = R0 - 1 + mr, mpx = C0 - 1 + mc;                        // un-clamped prototype coordinates (crop test)

    for (int d0 = 0; d0 < nt; d0 += DN) {
        const int nd = min(DN, nt - d0);
        // ---- B: coefficients of this pass's detections, tf32 hi + lo (rows >= nd zero)
        for (int i = tid; i < DN * VTI_NM; i += DENSE_THREADS) {
            const int j = i >> 5, q = i & 31;
            float v = 0.0f;
            if (j < nd) v = __ldg(a.det_coef + ((size_t)b * a.max_det + s_list[d0 + j]) * VTI_NM + q);
            const unsigned hi = tf32_rna(v);
            const unsigned lo = tf32_rna(__fsub_rn(v, __uint_as_float(hi)));
            *reinterpret_cast<unsigned*>(sB_hi + db_off(j, q)) = hi;
            *reinterpret_cast<unsigned*>(sB_lo + db_off(j, q)) = lo;
        }
        // generic-proxy writes -> visible to the tensor core's async-proxy reads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned aH = k4_smem_u32(sA_hi), aL = k4_smem_u32(sA_lo), bH = k4_smem_u32(sB_hi), bL = k4_smem_u32(sB_lo);
            bool acc = false;
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const unsigned aa = term == 1 ? aL : aH, bb = term == 2 ? bL : bH;
#pragma unroll
                for (int ks = 0; ks < VTI_NM / 8; ++ks) {
                    // a k-step = two 4-channel core matrices, 128 B apart (LBO); 8-row groups 1024 B apart (SBO)
                    const unsigned long long da = umma_desc(aa + ks * 256, 128, 1024);
                    const unsigned long long db = umma_desc(bb + ks * 256, 128, 1024);
                    const unsigned en = acc ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 :: "r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(en) : "memory");
                    acc = true;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         :: "r"(k4_smem_u32(&s_mbar)) : "memory");
        }
        // ---- wait for the accumulator
        {
            const unsigned bar = k4_smem_u32(&s_mbar);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "W_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\t"
                "bra W_%=;\n\t"
                "D_%=:\n\t}"
                :: "r"(bar), "r"(phase) : "memory");
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue 1: D[lane = pixel][column = detection] -> sigmoid + crop -> s_c[det][pixel]
        {
            const int col0 = (warp >> 2) * (DN / 2);                       // warps 0-3: columns 0..31, warps 4-7: 32..63
            const unsigned taddr = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)col0;
#pragma unroll
            for (int cc = 0; cc < DN / 2; cc += 16) {
                unsigned v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr + (unsigned)cc));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int dj = col0 + cc + j;
                    float val = 0.0f;
                    if (dj < nd && mr < nr && mc < ncw) {
                        const VtiWindow w = vti_det_window(dets[s_list[d0 + dj]].box_lb, a.ph, a.pw);
                        const int py = min(max(mpy, 0), a.ph - 1), px = min(max(mpx, 0), a.pw - 1);
                        if (py >= w.cy_lo && py <= w.cy_hi && px >= w.cx_lo && px <= w.cx_hi) {
                            const float acc = __uint_as_float(v[j]);
                            val = VB ? acc : 1.0f / (1.0f + expf(-acc));
                        }
                    }
                    s_c[dj * DT_M + my_m] = val;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- epilogue 2: cells of every detection of this pass, one warp per detection (the unit form's code)
        for (int j = warp; j < nd; j += DENSE_THREADS / 32) {
            const int k = s_list[d0 + j];
            vti_det* __restrict__ det = dets + k;
            const unsigned f = det->flags;
            const bool fabric = (f & VTI_F_FABRIC) && (f & VTI_F_IN_ROI);
            if (fabric) { s_envw[warp][lane] = a.upper ? INT_MAX : -1; s_envw[warp][lane + 32] = a.upper ? INT_MAX : -1; }
            __syncwarp();
            unit_cells<EXPORT, VB>(a, b, k, det, fabric, R0, C0, nr, ncw, s_c + j * DT_M, s_envw[warp], lane);
            __syncwarp();
        }
        __syncthreads();                                                   // s_c, sB are rewritten by the next pass
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(DN) : "memory");
    }
}


// Host-buffer path: copy, per frame and channel, the union rectangle of the crop windows (the only prototype values the
// unit kernel reads) from device-mapped pinned host memory into the device prototype buffer.  Rows are widened to
// 16-byte boundaries for float4 accesses; blockIdx = (channel, frame), warps take rows.
__global__ void __launch_bounds__(256) k4_fetch_proto_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             const int4* __restrict__ bbox, int ph, int pw) {
    const int q = blockIdx.x, b = blockIdx.y;
    const int4 bb = bbox[b];
    if (bb.z < bb.x || bb.w < bb.y) return;
    const int x0 = bb.x & ~3, x1 = min((bb.z + 4) & ~3, pw);               // [x0, x1) float4-aligned (pw % 4 == 0)
    const int n4 = (x1 - x0) >> 2;
    const size_t plane = (size_t)ph * pw, base = ((size_t)b * VTI_NM + q) * plane;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int y = bb.y + warp; y <= bb.w; y += 8) {
        const float4* s4 = reinterpret_cast<const float4*>(src + base + (size_t)y * pw + x0);
        float4* d4 = reinterpret_cast<float4*>(dst + base + (size_t)y * pw + x0);
        for (int i = lane; i < n4; i += 32) d4[i] = s4[i];
    }
}

}  // namespace

typedef CUresult (*vti_encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static vti_encode_tiled_t g_encode_tiled = nullptr;
constexpr size_t K4_TMA_SMEM = (size_t)T_WARPS * 2 * TB_BYTES;

// 3-D float32 tensor map (innermost dimension first) with box {b0, b1, b2}; false if the driver entry point is missing
bool vti_encode_tmap_f32_3d(CUtensorMap* m, const void* base, unsigned long long d0, unsigned long long d1,
                            unsigned long long d2, unsigned b0, unsigned b1, unsigned b2) {
    if (!g_encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode_tiled = reinterpret_cast<vti_encode_tiled_t>(fn);
        else
            cudaGetLastError();
        if (!g_encode_tiled) return false;
    }
    const cuuint64_t gdim[3] = {d0, d1, d2};
    const cuuint64_t gstr[2] = {d0 * 4, d0 * d1 * 4};
    const cuuint32_t box[3] = {b0, b1, b2}, estr[3] = {1, 1, 1};
    return g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int vti_k4_prepare() {
    // cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda
    if (!g_encode_tiled) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode_tiled = reinterpret_cast<vti_encode_tiled_t>(fn);
        else
            cudaGetLastError();
    }
    int rc;
    if ((rc = vti_raise_dyn_smem((const void*)k4_dense_kernel<true, true>, K4_DENSE_SMEM)) ||
        (rc = vti_raise_dyn_smem((const void*)k4_dense_kernel<true, false>, K4_DENSE_SMEM)) ||
        (rc = vti_raise_dyn_smem((const void*)k4_dense_kernel<false, true>, K4_DENSE_SMEM)) ||
        (rc = vti_raise_dyn_smem((const void*)k4_dense_kernel<false, false>, K4_DENSE_SMEM))) return rc;
    VTI_CUDA(cudaFuncSetAttribute(k4_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K4_TMA_SMEM));
    VTI_CUDA(cudaFuncSetAttribute(k4_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K4_TMA_SMEM));
    return VTI_OK;
}

int vti_launch_k4(vti_handle* h, const float* proto, int B, vti_det* dets, const int32_t* counts, uint32_t* masks,
                  cudaStream_t s) {
    K4Args a;
    a.proto = proto;
    a.dets = dets;
    a.det_coef = h->d_det_coef;
    a.units = h->d_units;
    a.unit_count = h->d_cand_count + h->p.max_batch;
    a.ly = h->lutY; a.lx = h->lutX;
    a.env = h->d_env;
    a.masks = masks;
    a.LH = h->g.LH; a.LW = h->g.LW; a.ph = h->g.ph; a.pw = h->g.pw; a.max_det = h->p.max_det;
    a.upper = (h->p.variant == 1);
    if (masks) {
        const size_t wpm = (size_t)a.LH * (a.LW / 32);
        k4_zero_masks_kernel<<<dim3(a.max_det, B), 256, 0, s>>>(masks, counts, a.max_det, wpm);
        h->launches++;
    }
    // TMA form: the prototype tensor as [B*32][ph][pw] float32, boxes of {TB_W, UR+1, 32}
    CUtensorMap tmap;
    // OPT-IN (VTI_K4_TMA=1): measured on B200 the TMA form is 2x SLOWER than the LDG form (137 vs 70 us per 64 frames):
    // a unit's box is 128 rows of 96 bytes, and the TMA unit's per-row request rate, not bandwidth or latency, bounds it.
    const bool vb = h->p.mask_variant == 1;
    bool tma = g_encode_tiled && getenv("VTI_K4_TMA") && !vb && (reinterpret_cast<uintptr_t>(proto) & 15) == 0 &&
               (a.pw % 4) == 0;
    if (tma) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a.pw, (cuuint64_t)a.ph, (cuuint64_t)VTI_NM * B};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.pw * 4, (cuuint64_t)a.pw * a.ph * 4};
        const cuuint32_t box[3] = {TB_W, TB_H, VTI_NM}, estr[3] = {1, 1, 1};
        tma = g_encode_tiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(proto), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (tma) {
        const int grid = 2 * h->num_sms;                   // two 4-warp CTAs per SM, grid-stride over the unit list
        if (masks) k4_tma_kernel<true><<<grid, T_WARPS * 32, K4_TMA_SMEM, s>>>(tmap, a);
        else k4_tma_kernel<false><<<grid, T_WARPS * 32, K4_TMA_SMEM, s>>>(tmap, a);
    } else {
        // LDG form.  Alone, K4 is fastest with four 8-warp CTAs per SM (66 vs 88 us on cfg2), but those fill every
        // register file and K1 -- which runs beside it on the pre stream -- then waits for K4 to finish.  Two CTAs per
        // SM leave half of each SM to K1: the K1 || K2-K5 step is 2.3 % faster on cfg2 / cfg3.  When K1 has little to do
        // (no remap, no resize: cfg4) K4 is on the critical path and keeps the four.  VTI_K4_GRID overrides.
        const bool k1_heavy = h->p.undistort || h->p.frame_w != h->g.new_w || h->p.frame_h != h->g.new_h;
        const int grid = (getenv("VTI_K4_GRID") ? atoi(getenv("VTI_K4_GRID")) : (k1_heavy ? 2 : 4)) * h->num_sms;
        if (h->p.k4_dense == 2) { /* every frame goes through the tile form below: K3 emitted no units */ }
        else if (masks && vb) k4_units_kernel<true, true><<<grid, K4_THREADS, 0, s>>>(a);
        else if (masks) k4_units_kernel<true, false><<<grid, K4_THREADS, 0, s>>>(a);
        else if (vb) k4_units_kernel<false, true><<<grid, K4_THREADS, 0, s>>>(a);
        else k4_units_kernel<false, false><<<grid, K4_THREADS, 0, s>>>(a);
    }
    if (h->p.k4_dense != 2) h->launches++;
    if (h->p.k4_dense) {
        // tcgen05 tile form for the frames K3 flagged: one CTA per (tile, frame); CTAs of other frames / empty tiles exit
        const dim3 dgrid(((a.pw + 1 + DT_C - 1) / DT_C) * ((a.ph + 1 + DT_R - 1) / DT_R), B);
        const int all_dets = masks != nullptr || h->p.mask_variant == 1 || getenv("VTI_ALL_DETS") != nullptr;
        if (masks && vb) k4_dense_kernel<true, true><<<dgrid, DENSE_THREADS, K4_DENSE_SMEM, s>>>(a, counts, h->d_dense, all_dets);
        else if (masks) k4_dense_kernel<true, false><<<dgrid, DENSE_THREADS, K4_DENSE_SMEM, s>>>(a, counts, h->d_dense, all_dets);
        else if (vb) k4_dense_kernel<false, true><<<dgrid, DENSE_THREADS, K4_DENSE_SMEM, s>>>(a, counts, h->d_dense, all_dets);
        else k4_dense_kernel<false, false><<<dgrid, DENSE_THREADS, K4_DENSE_SMEM, s>>>(a, counts, h->d_dense, all_dets);
        h->launches++;
    }
    VTI_CUDA(cudaGetLastError());
    return VTI_OK;
}

int vti_launch_fetch_proto(vti_handle* h, const float* host_proto_mapped, float* dev_proto, int B, cudaStream_t s) {
    if (h->g.pw % 4) {                                     // unaligned rows: plain copy of everything
        VTI_CUDA(cudaMemcpyAsync(dev_proto, host_proto_mapped, sizeof(float) * (size_t)B * VTI_NM * h->g.ph * h->g.pw,
                                 cudaMemcpyDefault, s));
        return VTI_OK;
    }
    k4_fetch_proto_kernel<<<dim3(VTI_NM, B), 256, 0, s>>>(host_proto_mapped, dev_proto, h->d_proto_bbox, h->g.ph, h->g.pw);
    h->launches++;
    VTI_CUDA(cudaGetLastError());
    return VTI_OK;
}
