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
// One CTA per (frame, 64x64 letterbox-pixel tile).  The CTA lists the detections whose non-zero mask region can
// touch the tile (cropped prototype window dilated by the bilinear support) and exits at once when there is none;
// otherwise it stages the tile's 18x18x32 prototype footprint in shared memory once and, per detection:
//   (1) one thread per footprint pixel: 32-term contraction + sigmoid + crop                    -> s_c[18][18]
//   (2) one thread per 4x4-pixel interpolation cell (the output pixels between four prototype pixels).  A cell
//       whose four corners are all > 0.5 (+margin) is entirely set, all < 0.5 (-margin) entirely clear -- bilinear
//       weights are a convex combination -- and its contribution to m00/m10/m01, the column extent and the envelope
//       is closed-form from prefix sums of the multiplicity tables; only boundary cells evaluate their 16 pixels.
//   (3) warp + shared reduction, a handful of 64-bit global atomics per (detection, tile).
// Optional bit-packed mask export: cells OR their bits into a shared 64x2-word tile that is then written out.
#include <climits>

#include "vti_internal.h"

namespace {

constexpr int K4_THREADS = 384;
constexpr int TW = 64;      // tile width  (letterbox px)
constexpr int TH = 64;      // tile height
constexpr int FR = 18;      // footprint rows  (TH/4 + 2)
constexpr int FC = 18;      // footprint cols  (TW/4 + 2)
constexpr int FP = FR * FC; // 324
constexpr int NCELL = 17 * 17;
constexpr int MAX_LIST = 1024;
constexpr float MARGIN = 1e-5f;   // >> fp32 rounding of the lerp; cells within the margin are evaluated per pixel

struct K4Args {
    const float* proto;         // [B][32][ph][pw]
    vti_det* dets;              // [B][max_det]
    const int32_t* counts;
    const float* det_coef;      // [B][max_det][32]
    AxisLut ly, lx;
    int32_t* env;               // [B][LW] frame-row envelope per letterbox column
    uint32_t* masks;            // optional [B][max_det][LH][LW/32]
    int LH, LW, ph, pw, max_det;
    int upper;                  // envelope mode
};

struct Window {                 // where a detection's mask can be non-zero
    int cx_lo, cx_hi, cy_lo, cy_hi;   // cropped prototype window (inclusive)
    int ox_lo, ox_hi, oy_lo, oy_hi;   // letterbox-pixel window (inclusive)
    bool empty;
};

__device__ __forceinline__ Window det_window(const float* box, int LH, int LW, int ph, int pw) {
    // crop_mask: keep prototype pixel (Y,X) iff X >= x1/4 && X < x2/4 && Y >= y1/4 && Y < y2/4 (float compares)
    const float dx1 = box[0] * 0.25f, dy1 = box[1] * 0.25f, dx2 = box[2] * 0.25f, dy2 = box[3] * 0.25f;
    Window w;
    w.cx_lo = max((int)ceilf(dx1), 0);
    w.cy_lo = max((int)ceilf(dy1), 0);
    w.cx_hi = min((int)ceilf(dx2) - 1, pw - 1);
    w.cy_hi = min((int)ceilf(dy2) - 1, ph - 1);
    w.empty = (w.cx_lo > w.cx_hi) || (w.cy_lo > w.cy_hi) || !(dx1 == dx1) || !(dx2 == dx2) || !(dy1 == dy1) || !(dy2 == dy2);
    w.ox_lo = max(4 * w.cx_lo - 2, 0);
    w.ox_hi = min(4 * w.cx_hi + 5, LW - 1);
    w.oy_lo = max(4 * w.cy_lo - 2, 0);
    w.oy_hi = min(4 * w.cy_hi + 5, LH - 1);
    return w;
}

__global__ void k4_zero_masks_kernel(uint32_t* masks, const int32_t* counts, int max_det, size_t words_per_mask) {
    const int k = blockIdx.x, b = blockIdx.y;
    if (k >= counts[b]) return;
    uint4* p = reinterpret_cast<uint4*>(masks + ((size_t)b * max_det + k) * words_per_mask);
    const size_t n4 = words_per_mask / 4;
    for (size_t i = threadIdx.x; i < n4; i += blockDim.x) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

template <bool EXPORT>
__global__ void __launch_bounds__(K4_THREADS) k4_masks_kernel(const K4Args a) {
    __shared__ float s_proto[VTI_NM][FP];
    __shared__ float s_c[FP];
    __shared__ unsigned short s_list[MAX_LIST];
    __shared__ int s_nlist;
    __shared__ unsigned long long s_m00, s_m10, s_m01;
    __shared__ int s_cmin, s_cmax;
    __shared__ int s_env[TW];
    __shared__ unsigned s_mask[TH][2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.z;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int n = a.counts[b];
    vti_det* __restrict__ dets = a.dets + (size_t)b * a.max_det;

    if (tid == 0) { s_nlist = 0; s_m00 = 0; s_m10 = 0; s_m01 = 0; s_cmin = INT_MAX; s_cmax = -1; }
    if (tid < TW) s_env[tid] = a.upper ? INT_MAX : -1;
    if (EXPORT && tid < TH * 2) (&s_mask[0][0])[tid] = 0u;
    __syncthreads();
    for (int k = tid; k < n; k += K4_THREADS) {
        const unsigned f = dets[k].flags;
        const bool wanted = EXPORT || ((f & VTI_F_IN_ROI) && (f & (VTI_F_STITCH | VTI_F_FABRIC)));
        if (!wanted) continue;
        const Window w = det_window(dets[k].box_lb, a.LH, a.LW, a.ph, a.pw);
        if (w.empty || w.ox_hi < X0 || w.ox_lo > X0 + TW - 1 || w.oy_hi < Y0 || w.oy_lo > Y0 + TH - 1) continue;
        s_list[atomicAdd(&s_nlist, 1)] = (unsigned short)k;
    }
    __syncthreads();
    const int nlist = s_nlist;
    if (nlist == 0) return;

    // ---- stage the prototype footprint (replicate-clamped at the plane border): one thread per footprint pixel,
    //      looping over the 32 channels, so the index arithmetic is done once per thread
    const int pr0 = (Y0 >> 2) - 1, pc0 = (X0 >> 2) - 1;
    const float* __restrict__ proto = a.proto + (size_t)b * VTI_NM * a.ph * a.pw;
    if (tid < FP) {
        const int fr = tid / FC, fc = tid - fr * FC;
        const int py = min(max(pr0 + fr, 0), a.ph - 1), px = min(max(pc0 + fc, 0), a.pw - 1);
        const float* __restrict__ src = proto + (size_t)py * a.pw + px;
        const size_t plane = (size_t)a.ph * a.pw;
#pragma unroll 8
        for (int k = 0; k < VTI_NM; ++k) s_proto[k][tid] = __ldg(src + k * plane);
    }

    // this thread's footprint pixel (phase 1) and interpolation cell (phase 2)
    const int p_fr = tid / FC, p_fc = tid - p_fr * FC;
    const int p_py = min(max(pr0 + p_fr, 0), a.ph - 1), p_px = min(max(pc0 + p_fc, 0), a.pw - 1);
    const int c_r = tid / 17, c_c = tid - c_r * 17;
    // cell (r,c) covers output rows Y0-2+4r .. Y0+1+4r and cols X0-2+4c .. X0+1+4c, clipped to the tile and the image
    const int cy_first = Y0 - 2 + 4 * c_r, cx_first = X0 - 2 + 4 * c_c;
    const int ya = max(cy_first, Y0), yb = min(cy_first + 3, min(Y0 + TH - 1, a.LH - 1));
    const int xa = max(cx_first, X0), xb = min(cx_first + 3, min(X0 + TW - 1, a.LW - 1));
    const bool cell_ok = (tid < NCELL) && (ya <= yb) && (xa <= xb);
    // closed-form contribution of a fully set cell
    int f_cy = 0, f_sy = 0, f_cx = 0, f_sx = 0, f_cmin = INT_MAX, f_cmax = -1, f_env = 0;
    if (cell_ok) {
        f_cy = a.ly.pc[yb + 1] - a.ly.pc[ya];
        f_sy = a.ly.ps[yb + 1] - a.ly.ps[ya];
        f_cx = a.lx.pc[xb + 1] - a.lx.pc[xa];
        f_sx = a.lx.ps[xb + 1] - a.lx.ps[xa];
        if (f_cy > 0 && f_cx > 0) { f_cmin = a.lx.next_first[xa]; f_cmax = a.lx.prev_last[xb]; }
        f_env = a.upper ? a.ly.next_first[ya] : a.ly.prev_last[yb];
    }

    for (int li = 0; li < nlist; ++li) {
        const int k = s_list[li];
        const Window w = det_window(dets[k].box_lb, a.LH, a.LW, a.ph, a.pw);
        __syncthreads();                 // (first iteration) s_proto visible; previous flush finished
        // (1) logits -> sigmoid -> crop, over the footprint
        if (tid < FP) {
            float c = 0.0f;
            if (p_py >= w.cy_lo && p_py <= w.cy_hi && p_px >= w.cx_lo && p_px <= w.cx_hi) {
                const float* __restrict__ coef = a.det_coef + ((size_t)b * a.max_det + k) * VTI_NM;
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < VTI_NM; ++q) acc = fmaf(__ldg(coef + q), s_proto[q][tid], acc);
                c = 1.0f / (1.0f + expf(-acc));
            }
            s_c[tid] = c;
        }
        __syncthreads();
        // (2) cells
        const unsigned f = dets[k].flags;
        const bool fabric = (f & VTI_F_FABRIC) && (f & VTI_F_IN_ROI);
        int m00 = 0, m10 = 0, m01 = 0, cmin = INT_MAX, cmax = -1;
        if (cell_ok) {
            const float c00 = s_c[c_r * FC + c_c], c01 = s_c[c_r * FC + c_c + 1];
            const float c10 = s_c[(c_r + 1) * FC + c_c], c11 = s_c[(c_r + 1) * FC + c_c + 1];
            const float vmin = fminf(fminf(c00, c01), fminf(c10, c11));
            const float vmax = fmaxf(fmaxf(c00, c01), fmaxf(c10, c11));
            if (vmin > 0.5f + MARGIN) {
                m00 = f_cy * f_cx; m10 = f_cy * f_sx; m01 = f_sy * f_cx;
                cmin = f_cmin; cmax = f_cmax;
                if (fabric && f_cy > 0) {
                    for (int X = xa; X <= xb; ++X)
                        if (a.lx.cnt[X] > 0) {
                            if (a.upper) atomicMin(&s_env[X - X0], f_env); else atomicMax(&s_env[X - X0], f_env);
                        }
                }
                if (EXPORT) {
                    const unsigned long long bits = ((1ull << (xb - xa + 1)) - 1ull) << (xa - X0);
                    for (int Y = ya; Y <= yb; ++Y) {
                        if ((unsigned)bits) atomicOr(&s_mask[Y - Y0][0], (unsigned)bits);
                        if ((unsigned)(bits >> 32)) atomicOr(&s_mask[Y - Y0][1], (unsigned)(bits >> 32));
                    }
                }
            } else if (vmax >= 0.5f - MARGIN) {
                // boundary cell: evaluate its pixels (torch upsample_bilinear2d, align_corners=False, scale 1/4)
                for (int Y = ya; Y <= yb; ++Y) {
                    const float ly1 = (float)(2 * (Y - cy_first) + 1) * 0.125f;
                    const int cY = a.ly.cnt[Y], sY = a.ly.sum[Y];
                    unsigned long long rowbits = 0ull;
                    for (int X = xa; X <= xb; ++X) {
                        const float lx1 = (float)(2 * (X - cx_first) + 1) * 0.125f;
                        const float top = (1.0f - lx1) * c00 + lx1 * c01;
                        const float bot = (1.0f - lx1) * c10 + lx1 * c11;
                        const float v = (1.0f - ly1) * top + ly1 * bot;
                        if (v > 0.5f) {
                            rowbits |= 1ull << (X - X0);
                            const int cX = a.lx.cnt[X];
                            m00 += cY * cX; m10 += cY * a.lx.sum[X]; m01 += sY * cX;
                            if (cY > 0 && cX > 0) {
                                cmin = min(cmin, a.lx.first[X]);
                                cmax = max(cmax, a.lx.last[X]);
                                if (fabric) {
                                    if (a.upper) atomicMin(&s_env[X - X0], a.ly.first[Y]);
                                    else atomicMax(&s_env[X - X0], a.ly.last[Y]);
                                }
                            }
                        }
                    }
                    if (EXPORT) {
                        if ((unsigned)rowbits) atomicOr(&s_mask[Y - Y0][0], (unsigned)rowbits);
                        if ((unsigned)(rowbits >> 32)) atomicOr(&s_mask[Y - Y0][1], (unsigned)(rowbits >> 32));
                    }
                }
            }
        }
        // (3) per-warp reduction, shared atomics
        m00 = __reduce_add_sync(0xffffffffu, m00);
        if (m00 > 0) {                                   // warp-uniform
            m10 = __reduce_add_sync(0xffffffffu, m10);
            m01 = __reduce_add_sync(0xffffffffu, m01);
            cmin = __reduce_min_sync(0xffffffffu, cmin);
            cmax = __reduce_max_sync(0xffffffffu, cmax);
            if (lane == 0) {
                atomicAdd(&s_m00, (unsigned long long)m00);
                atomicAdd(&s_m10, (unsigned long long)m10);
                atomicAdd(&s_m01, (unsigned long long)m01);
                atomicMin(&s_cmin, cmin);
                atomicMax(&s_cmax, cmax);
            }
        }
        __syncthreads();
        // flush (the threads that read also reset; the next iteration's first barrier orders it)
        if (tid == 0) {
            if (s_m00 > 0) {
                atomicAdd((unsigned long long*)&dets[k].m00, s_m00);
                atomicAdd((unsigned long long*)&dets[k].m10, s_m10);
                atomicAdd((unsigned long long*)&dets[k].m01, s_m01);
                atomicMin(&dets[k].col_min, s_cmin);
                atomicMax(&dets[k].col_max, s_cmax);
            }
            s_m00 = 0; s_m10 = 0; s_m01 = 0; s_cmin = INT_MAX; s_cmax = -1;
        }
        if (fabric && tid >= 32 && tid < 32 + TW) {
            const int t = tid - 32;
            const int ev = s_env[t];
            if (X0 + t < a.LW) {
                if (a.upper) { if (ev != INT_MAX) atomicMin(&a.env[(size_t)b * a.LW + X0 + t], ev); }
                else { if (ev >= 0) atomicMax(&a.env[(size_t)b * a.LW + X0 + t], ev); }
            }
            s_env[t] = a.upper ? INT_MAX : -1;
        }
        if (EXPORT && tid >= 128 && tid < 128 + TH * 2) {
            const int t = tid - 128, row = t >> 1, wd = t & 1;
            const int Y = Y0 + row, word = (X0 >> 5) + wd;
            if (Y < a.LH && word < a.LW / 32)
                a.masks[(((size_t)b * a.max_det + k) * a.LH + Y) * (a.LW / 32) + word] = s_mask[row][wd];
            s_mask[row][wd] = 0u;
        }
    }
}

}  // namespace

int vti_launch_k4(vti_handle* h, const float* proto, int B, vti_det* dets, const int32_t* counts, uint32_t* masks,
                  cudaStream_t s) {
    K4Args a;
    a.proto = proto;
    a.dets = dets;
    a.counts = counts;
    a.det_coef = h->d_det_coef;
    a.ly = h->lutY; a.lx = h->lutX;
    a.env = h->d_env;
    a.masks = masks;
    a.LH = h->g.LH; a.LW = h->g.LW; a.ph = h->g.ph; a.pw = h->g.pw; a.max_det = h->p.max_det;
    a.upper = (h->p.variant == 1);
    dim3 grid((a.LW + TW - 1) / TW, (a.LH + TH - 1) / TH, B);
    if (masks) {
        const size_t wpm = (size_t)a.LH * (a.LW / 32);
        k4_zero_masks_kernel<<<dim3(a.max_det, B), 256, 0, s>>>(masks, counts, a.max_det, wpm);
        h->launches++;
        k4_masks_kernel<true><<<grid, K4_THREADS, 0, s>>>(a);
    } else {
        k4_masks_kernel<false><<<grid, K4_THREADS, 0, s>>>(a);
    }
    h->launches++;
    VTI_CUDA(cudaGetLastError());
    return VTI_OK;
}
