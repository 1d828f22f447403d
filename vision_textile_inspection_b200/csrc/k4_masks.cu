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

// Work unit of one warp: one detection x one band of interpolation-cell rows of the tile.
constexpr int NWARP = K4_THREADS / 32;
constexpr int BAND = 6;                 // cell rows per unit (17 cell rows -> 3 bands)
constexpr int MAX_UNITS = 3 * MAX_LIST;

template <bool EXPORT>
__global__ void __launch_bounds__(K4_THREADS) k4_masks_kernel(const K4Args a) {
    extern __shared__ __align__(16) float s_dyn[];         // s_proto[VTI_NM][FP]
    float (*s_proto)[FP] = reinterpret_cast<float (*)[FP]>(s_dyn);
    __shared__ float s_c[NWARP][(BAND + 1) * FC];          // per-warp sigmoid/crop values of the unit's corner rows
    __shared__ unsigned short s_units[MAX_UNITS];          // det | band << 10   (max_det <= 1024)
    __shared__ int s_nunits, s_next;
    __shared__ int s_env[TW];
    __shared__ unsigned s_mask[EXPORT ? NWARP : 1][BAND * 4][2];
    // the tile's slice of the nearest-resize multiplicity tables (rows Y0.., cols X0..)
    __shared__ int t_ypc[TH + 1], t_yps[TH + 1], t_ycnt[TH], t_ysum[TH], t_yfirst[TH], t_ylast[TH], t_ynf[TH], t_ypl[TH];
    __shared__ int t_xpc[TW + 1], t_xps[TW + 1], t_xcnt[TW], t_xsum[TW], t_xfirst[TW], t_xlast[TW], t_xnf[TW], t_xpl[TW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int n = a.counts[b];
    vti_det* __restrict__ dets = a.dets + (size_t)b * a.max_det;
    const int pr0 = (Y0 >> 2) - 1, pc0 = (X0 >> 2) - 1;    // prototype coordinates of footprint (0, 0)

    if (tid == 0) { s_nunits = 0; s_next = 0; }
    if (tid < TW) s_env[tid] = a.upper ? INT_MAX : -1;
    __syncthreads();
    // ---- unit list: detections whose non-zero mask region can touch the tile, split into cell-row bands
    for (int k = tid; k < n; k += K4_THREADS) {
        const unsigned f = dets[k].flags;
        const bool wanted = EXPORT || ((f & VTI_F_IN_ROI) && (f & (VTI_F_STITCH | VTI_F_FABRIC)));
        if (!wanted) continue;
        const Window w = det_window(dets[k].box_lb, a.LH, a.LW, a.ph, a.pw);
        if (w.empty || w.ox_hi < X0 || w.ox_lo > X0 + TW - 1 || w.oy_hi < Y0 || w.oy_lo > Y0 + TH - 1) continue;
        // cell (r, c) has corners at footprint rows r, r+1: it can be non-zero iff one of them is inside the crop window
        const int cr_lo = max(w.cy_lo - pr0 - 1, 0), cr_hi = min(w.cy_hi - pr0, 16);
        for (int bi = cr_lo / BAND; bi <= cr_hi / BAND; ++bi) {
            const int u = atomicAdd(&s_nunits, 1);
            if (u < MAX_UNITS) s_units[u] = (unsigned short)(k | (bi << 10));
        }
    }
    __syncthreads();
    const int nunits = min(s_nunits, MAX_UNITS);
    if (nunits == 0) return;

    // ---- stage the prototype footprint (replicate-clamped at the plane border): one thread per footprint pixel
    const float* __restrict__ proto = a.proto + (size_t)b * VTI_NM * a.ph * a.pw;
    if (tid < FP) {
        const int fr = tid / FC, fc = tid - fr * FC;
        const int py = min(max(pr0 + fr, 0), a.ph - 1), px = min(max(pc0 + fc, 0), a.pw - 1);
        const float* __restrict__ src = proto + (size_t)py * a.pw + px;
        const size_t plane = (size_t)a.ph * a.pw;
#pragma unroll 16
        for (int k = 0; k < VTI_NM; ++k) s_proto[k][tid] = __ldg(src + k * plane);
    }
    // (threads 0..128 also fetch the table slices; they are tiny and L2-resident)
    if (tid <= TH) {
        const int Y = min(Y0 + tid, a.LH), X = min(X0 + tid, a.LW);          // prefix arrays have n + 1 entries
        t_ypc[tid] = a.ly.pc[Y]; t_yps[tid] = a.ly.ps[Y];
        t_xpc[tid] = a.lx.pc[X]; t_xps[tid] = a.lx.ps[X];
    }
    if (tid < TH) {
        const int Y = min(Y0 + tid, a.LH - 1);
        t_ycnt[tid] = a.ly.cnt[Y]; t_ysum[tid] = a.ly.sum[Y]; t_yfirst[tid] = a.ly.first[Y]; t_ylast[tid] = a.ly.last[Y];
        t_ynf[tid] = a.ly.next_first[Y]; t_ypl[tid] = a.ly.prev_last[Y];
    } else if (tid < TH + TW) {
        const int i = tid - TH, X = min(X0 + i, a.LW - 1);
        t_xcnt[i] = a.lx.cnt[X]; t_xsum[i] = a.lx.sum[X]; t_xfirst[i] = a.lx.first[X]; t_xlast[i] = a.lx.last[X];
        t_xnf[i] = a.lx.next_first[X]; t_xpl[i] = a.lx.prev_last[X];
    }
    __syncthreads();

    float* sc = s_c[warp];
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(&s_next, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nunits) break;
        const unsigned un = s_units[u];
        const int k = (int)(un & 0x3FFu), bi = (int)(un >> 10);
        const Window w = det_window(dets[k].box_lb, a.LH, a.LW, a.ph, a.pw);
        const unsigned f = dets[k].flags;
        const bool fabric = (f & VTI_F_FABRIC) && (f & VTI_F_IN_ROI);
        // crop window in footprint coordinates (may extend past the footprint)
        const int wy_lo = w.cy_lo - pr0, wy_hi = w.cy_hi - pr0, wx_lo = w.cx_lo - pc0, wx_hi = w.cx_hi - pc0;
        // cells of this unit: rows [cr_lo, cr_hi], cols [cc_lo, cc_hi] (inside the tile's 17 x 17 cells)
        const int cr_lo = max(max(wy_lo - 1, 0), bi * BAND), cr_hi = min(min(wy_hi, 16), bi * BAND + BAND - 1);
        const int cc_lo = max(wx_lo - 1, 0), cc_hi = min(wx_hi, 16);
        const int nr = cr_hi - cr_lo + 2, ncw = cc_hi - cc_lo + 2;          // corner rows / cols
        const float inv_ncw = 1.0f / (float)ncw;
        const float* __restrict__ coef = a.det_coef + ((size_t)b * a.max_det + k) * VTI_NM;
        __syncwarp();
        // (1) logits -> sigmoid -> crop over the corner rectangle
        for (int i = lane; i < nr * ncw; i += 32) {
            const int r = (int)(((float)i + 0.5f) * inv_ncw), c = i - r * ncw;
            const int fr = cr_lo + r, fc = cc_lo + c;
            float v = 0.0f;
            if (fr >= wy_lo && fr <= wy_hi && fc >= wx_lo && fc <= wx_hi) {
                const int p = fr * FC + fc;
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < VTI_NM; ++q) acc = fmaf(__ldg(coef + q), s_proto[q][p], acc);
                v = 1.0f / (1.0f + expf(-acc));
            }
            sc[r * FC + c] = v;
        }
        if (EXPORT) for (int i = lane; i < BAND * 4 * 2; i += 32) (&s_mask[EXPORT ? warp : 0][0][0])[i] = 0u;
        __syncwarp();
        // (2) cells
        const int ncc = ncw - 1;
        const float inv_ncc = 1.0f / (float)ncc;
        long long m00 = 0, m10 = 0, m01 = 0;
        int cmin = INT_MAX, cmax = -1;
        for (int i = lane; i < (nr - 1) * ncc; i += 32) {
            const int r = (int)(((float)i + 0.5f) * inv_ncc), c = i - r * ncc;
            const int c_r = cr_lo + r, c_c = cc_lo + c;
            // cell (c_r, c_c) covers output rows Y0-2+4 c_r .. +3 and cols X0-2+4 c_c .. +3, clipped to tile and image
            const int cy_first = Y0 - 2 + 4 * c_r, cx_first = X0 - 2 + 4 * c_c;
            const int ya = max(cy_first, Y0), yb = min(cy_first + 3, min(Y0 + TH - 1, a.LH - 1));
            const int xa = max(cx_first, X0), xb = min(cx_first + 3, min(X0 + TW - 1, a.LW - 1));
            if (ya > yb || xa > xb) continue;
            const float c00 = sc[r * FC + c], c01 = sc[r * FC + c + 1];
            const float c10 = sc[(r + 1) * FC + c], c11 = sc[(r + 1) * FC + c + 1];
            const float vmin = fminf(fminf(c00, c01), fminf(c10, c11));
            const float vmax = fmaxf(fmaxf(c00, c01), fmaxf(c10, c11));
            if (vmin > 0.5f + MARGIN) {
                // fully set: closed form from the prefix sums of the nearest-resize multiplicity tables
                const int f_cy = t_ypc[yb + 1 - Y0] - t_ypc[ya - Y0], f_sy = t_yps[yb + 1 - Y0] - t_yps[ya - Y0];
                const int f_cx = t_xpc[xb + 1 - X0] - t_xpc[xa - X0], f_sx = t_xps[xb + 1 - X0] - t_xps[xa - X0];
                m00 += f_cy * f_cx; m10 += f_cy * f_sx; m01 += f_sy * f_cx;
                if (f_cy > 0 && f_cx > 0) {
                    cmin = min(cmin, t_xnf[xa - X0]); cmax = max(cmax, t_xpl[xb - X0]);
                    if (fabric) {
                        const int f_env = a.upper ? t_ynf[ya - Y0] : t_ypl[yb - Y0];
                        for (int X = xa; X <= xb; ++X)
                            if (t_xcnt[X - X0] > 0) {
                                if (a.upper) atomicMin(&s_env[X - X0], f_env); else atomicMax(&s_env[X - X0], f_env);
                            }
                    }
                }
                if (EXPORT) {
                    const unsigned long long bits = ((1ull << (xb - xa + 1)) - 1ull) << (xa - X0);
                    for (int Y = ya; Y <= yb; ++Y) {
                        const int row = Y - (Y0 - 2 + 4 * bi * BAND);
                        if ((unsigned)bits) atomicOr(&s_mask[EXPORT ? warp : 0][row][0], (unsigned)bits);
                        if ((unsigned)(bits >> 32)) atomicOr(&s_mask[EXPORT ? warp : 0][row][1], (unsigned)(bits >> 32));
                    }
                }
            } else if (vmax >= 0.5f - MARGIN) {
                // boundary cell: evaluate its pixels (torch upsample_bilinear2d, align_corners=False, scale 1/4)
                for (int Y = ya; Y <= yb; ++Y) {
                    const float ly1 = (float)(2 * (Y - cy_first) + 1) * 0.125f;
                    const int cY = t_ycnt[Y - Y0], sY = t_ysum[Y - Y0];
                    unsigned long long rowbits = 0ull;
                    for (int X = xa; X <= xb; ++X) {
                        const float lx1 = (float)(2 * (X - cx_first) + 1) * 0.125f;
                        const float top = (1.0f - lx1) * c00 + lx1 * c01;
                        const float bot = (1.0f - lx1) * c10 + lx1 * c11;
                        const float v = (1.0f - ly1) * top + ly1 * bot;
                        if (v > 0.5f) {
                            rowbits |= 1ull << (X - X0);
                            const int cX = t_xcnt[X - X0];
                            m00 += cY * cX; m10 += cY * t_xsum[X - X0]; m01 += sY * cX;
                            if (cY > 0 && cX > 0) {
                                cmin = min(cmin, t_xfirst[X - X0]);
                                cmax = max(cmax, t_xlast[X - X0]);
                                if (fabric) {
                                    if (a.upper) atomicMin(&s_env[X - X0], t_yfirst[Y - Y0]);
                                    else atomicMax(&s_env[X - X0], t_ylast[Y - Y0]);
                                }
                            }
                        }
                    }
                    if (EXPORT) {
                        const int row = Y - (Y0 - 2 + 4 * bi * BAND);
                        if ((unsigned)rowbits) atomicOr(&s_mask[EXPORT ? warp : 0][row][0], (unsigned)rowbits);
                        if ((unsigned)(rowbits >> 32)) atomicOr(&s_mask[EXPORT ? warp : 0][row][1], (unsigned)(rowbits >> 32));
                    }
                }
            }
        }
        // (3) warp reduction, one set of global atomics per unit
        const bool any = __any_sync(0xffffffffu, m00 > 0);
        if (any) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                m00 += __shfl_xor_sync(0xffffffffu, m00, o);
                m10 += __shfl_xor_sync(0xffffffffu, m10, o);
                m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            }
            cmin = __reduce_min_sync(0xffffffffu, cmin);
            cmax = __reduce_max_sync(0xffffffffu, cmax);
            if (lane == 0) {
                atomicAdd((unsigned long long*)&dets[k].m00, (unsigned long long)m00);
                atomicAdd((unsigned long long*)&dets[k].m10, (unsigned long long)m10);
                atomicAdd((unsigned long long*)&dets[k].m01, (unsigned long long)m01);
                atomicMin(&dets[k].col_min, cmin);
                atomicMax(&dets[k].col_max, cmax);
            }
        }
        if (EXPORT) {
            __syncwarp();
            // the unit owns output rows Y0-2+4*bi*BAND .. +4*BAND-1 of this tile's two mask words: plain stores
            for (int i = lane; i < BAND * 4 * 2; i += 32) {
                const int row = i >> 1, wd = i & 1;
                const int Y = Y0 - 2 + 4 * bi * BAND + row, word = (X0 >> 5) + wd;
                const unsigned bits = s_mask[EXPORT ? warp : 0][row][wd];
                if (bits && Y >= Y0 && Y < min(Y0 + TH, a.LH) && word < a.LW / 32)
                    atomicOr(&a.masks[(((size_t)b * a.max_det + k) * a.LH + Y) * (a.LW / 32) + word], bits);
            }
        }
    }
    __syncthreads();
    if (tid < TW && X0 + tid < a.LW) {
        const int ev = s_env[tid];
        if (a.upper) { if (ev != INT_MAX) atomicMin(&a.env[(size_t)b * a.LW + X0 + tid], ev); }
        else { if (ev >= 0) atomicMax(&a.env[(size_t)b * a.LW + X0 + tid], ev); }
    }
}

}  // namespace

constexpr size_t K4_DYN_SMEM = sizeof(float) * VTI_NM * FP;

int vti_k4_prepare() {
    VTI_CUDA(cudaFuncSetAttribute(k4_masks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K4_DYN_SMEM));
    VTI_CUDA(cudaFuncSetAttribute(k4_masks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K4_DYN_SMEM));
    return VTI_OK;
}

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
        k4_masks_kernel<true><<<grid, K4_THREADS, K4_DYN_SMEM, s>>>(a);
    } else {
        k4_masks_kernel<false><<<grid, K4_THREADS, K4_DYN_SMEM, s>>>(a);
    }
    h->launches++;
    VTI_CUDA(cudaGetLastError());
    return VTI_OK;
}
