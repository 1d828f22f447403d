// K1 -- fused [cv2.undistort] + Ultralytics LetterBox (cv2.resize INTER_LINEAR + 114 border) + HWC->CHW + /255.
//
// Replaces: cv2.undistort (north-star addition), ultralytics LetterBox.__call__ and
// BasePredictor.preprocess, reached from /root/reference/measurement.py:205-210.
// Spec: oracle/cv_fixed.py (bit-exact integer formulas, pinned against cv2 in tests/test_oracle_cv.py).
//
// One CTA produces a TY x TX tile of the letterboxed output for all three planes.
//   phase A  stage the source footprint of the tile in shared memory as planar uint8; with undistort on, every
//            staged pixel is the 4-tap fixed-point remap of the raw frame (the intermediate uint8 rounding of
//            cv2.undistort is preserved -- a single composed bilinear sample would NOT be bit-exact)
//   phase B  11-bit H pass + V pass of cv2.resize from shared memory, /255 through a 256-entry table
//            (true division, not *1/255), 128-bit coalesced stores of the fp32 planes; pad = 114/255.
// HBM traffic: frame read once (tile halos hit L2) + 12*LH*LW written once.
#include "vti_internal.h"

namespace {

constexpr int TX = 128;
constexpr int TY = 16;
constexpr int K1_THREADS = 256;
constexpr int MAXROWS = 2 * TY + 2;
constexpr int MAXCOLS = 2 * TX + 2;
constexpr int PITCH = 272;   // >= MAXCOLS, multiple of 16

struct K1Args {
    const uint8_t* frames;
    float* out;
    const int32_t* tap_x_idx;   // [new_w]
    const int16_t* tap_x_a;     // [new_w][2]
    const int32_t* tap_y_i;     // [new_h][2]
    const int16_t* tap_y_b;     // [new_h][2]
    const int32_t* und_lut;     // [h*w] (dy16 << 16) | (dx16 & 0xffff)
    int h, w, new_h, new_w, top, left, LH, LW;
    int mode;                   // 1 bilinear (also used for the identity copy), 2 exact-2x area
    int flip;
};

__device__ __forceinline__ int remap_px(const uint8_t* __restrict__ f, int h, int w, int sy, int sx, int lut, int ch) {
    const int dx = (int)(short)(lut & 0xffff);
    const int dy = lut >> 16;
    const int ix = sx * 32 + dx, iy = sy * 32 + dy;
    const int x = ix >> 5, y = iy >> 5, fx = ix & 31, fy = iy & 31;
    const bool x0 = (unsigned)x < (unsigned)w, x1 = (unsigned)(x + 1) < (unsigned)w;
    const bool y0 = (unsigned)y < (unsigned)h, y1 = (unsigned)(y + 1) < (unsigned)h;
    const uint8_t* r0 = f + ((size_t)y * w + x) * 3 + ch;
    const uint8_t* r1 = r0 + (size_t)w * 3;
    const int p00 = (y0 && x0) ? r0[0] : 0;
    const int p01 = (y0 && x1) ? r0[3] : 0;
    const int p10 = (y1 && x0) ? r1[0] : 0;
    const int p11 = (y1 && x1) ? r1[3] : 0;
    const int acc = p00 * ((32 - fx) * (32 - fy)) + p01 * (fx * (32 - fy)) + p10 * ((32 - fx) * fy) + p11 * (fx * fy);
    return (acc + 512) >> 10;   // == (sum(w15 * p) + 2^14) >> 15 because every 15-bit weight is 32 * (10-bit weight)
}

template <bool UNDISTORT>
__global__ void __launch_bounds__(K1_THREADS) k1_letterbox_kernel(const K1Args a) {
    __shared__ __align__(16) uint8_t s_px[3][MAXROWS][PITCH];
    __shared__ float s_div[256];
    __shared__ int s_rowsrc[MAXROWS];
    __shared__ int s_colsrc[MAXCOLS];

    const int tid = threadIdx.x;
    const int X0 = blockIdx.x * TX, Y0 = blockIdx.y * TY;
    const int b = blockIdx.z;
    const uint8_t* __restrict__ frame = a.frames + (size_t)b * a.h * a.w * 3;
    float* __restrict__ out = a.out + (size_t)b * 3 * a.LH * a.LW;

    s_div[tid] = __fdiv_rn((float)tid, 255.0f);

    // resized-image rows / cols covered by this tile
    const int ry_lo = max(Y0 - a.top, 0), ry_hi = min(Y0 + TY - 1 - a.top, a.new_h - 1);
    const int rx_lo = max(X0 - a.left, 0), rx_hi = min(X0 + TX - 1 - a.left, a.new_w - 1);
    const bool any = (ry_lo <= ry_hi) && (rx_lo <= rx_hi);
    int nrows = 0, ncols = 0, r_lo = 0, c_lo = 0;
    bool row_contig = true, col_contig = true;
    if (any) {
        r_lo = a.tap_y_i[2 * ry_lo];
        const int r_hi = a.tap_y_i[2 * ry_hi + 1];
        row_contig = (r_hi - r_lo + 1) <= MAXROWS;
        nrows = row_contig ? (r_hi - r_lo + 1) : 2 * (ry_hi - ry_lo + 1);
        c_lo = a.tap_x_idx[rx_lo];
        const int c_hi = min(a.tap_x_idx[rx_hi] + 1, a.w - 1);
        col_contig = (c_hi - c_lo + 1) <= MAXCOLS;
        ncols = col_contig ? (c_hi - c_lo + 1) : 2 * (rx_hi - rx_lo + 1);
        for (int i = tid; i < nrows; i += K1_THREADS)
            s_rowsrc[i] = row_contig ? (r_lo + i) : a.tap_y_i[2 * ry_lo + i];
        for (int i = tid; i < ncols; i += K1_THREADS)
            s_colsrc[i] = col_contig ? (c_lo + i) : min(a.tap_x_idx[rx_lo + (i >> 1)] + (i & 1), a.w - 1);
    }
    __syncthreads();

    // ---- phase A: stage source pixels
    if (any) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < nrows; r += K1_THREADS / 32) {
            const int sy = s_rowsrc[r];
            for (int c = lane; c < ncols; c += 32) {
                const int sx = s_colsrc[c];
                if (UNDISTORT) {
                    const int lut = __ldg(a.und_lut + (size_t)sy * a.w + sx);
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        s_px[ch][r][c] = (uint8_t)remap_px(frame, a.h, a.w, sy, sx, lut, a.flip ? 2 - ch : ch);
                } else {
                    const uint8_t* p = frame + ((size_t)sy * a.w + sx) * 3;
                    const uint8_t v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
                    s_px[0][r][c] = a.flip ? v2 : v0;
                    s_px[1][r][c] = v1;
                    s_px[2][r][c] = a.flip ? v0 : v2;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B: resize + normalise + store.  item = (plane, tile row, group of 4 columns)
    const float pad = s_div[114];
    for (int item = tid; item < 3 * TY * (TX / 4); item += K1_THREADS) {
        const int xq = item % (TX / 4);
        const int j = (item / (TX / 4)) % TY;
        const int ch = item / (TY * (TX / 4));
        const int Y = Y0 + j, X = X0 + xq * 4;
        if (Y >= a.LH || X >= a.LW) continue;
        const int ry = Y - a.top;
        float4 o = make_float4(pad, pad, pad, pad);
        if (ry >= 0 && ry < a.new_h) {
            int rs0, rs1;
            if (row_contig) {
                rs0 = a.tap_y_i[2 * ry] - r_lo;
                rs1 = a.tap_y_i[2 * ry + 1] - r_lo;
            } else {
                rs0 = 2 * (ry - ry_lo);
                rs1 = rs0 + 1;
            }
            const int b0 = a.tap_y_b[2 * ry], b1 = a.tap_y_b[2 * ry + 1];
            const uint8_t* row0 = s_px[ch][rs0];
            const uint8_t* row1 = s_px[ch][rs1];
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rx = X + k - a.left;
                v[k] = pad;
                if (rx >= 0 && rx < a.new_w) {
                    int cs0, cs1;
                    if (col_contig) {
                        const int sx = a.tap_x_idx[rx];
                        cs0 = sx - c_lo;
                        cs1 = min(sx + 1, a.w - 1) - c_lo;
                    } else {
                        cs0 = 2 * (rx - rx_lo);
                        cs1 = cs0 + 1;
                    }
                    const int p00 = row0[cs0], p01 = row0[cs1], p10 = row1[cs0], p11 = row1[cs1];
                    int q;
                    if (a.mode == 2) {
                        q = (p00 + p01 + p10 + p11 + 2) >> 2;
                    } else {
                        const int a0 = a.tap_x_a[2 * rx], a1 = a.tap_x_a[2 * rx + 1];
                        const int S0 = p00 * a0 + p01 * a1, S1 = p10 * a0 + p11 * a1;
                        q = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
                        q = min(max(q, 0), 255);
                    }
                    v[k] = s_div[q];
                }
            }
            o = make_float4(v[0], v[1], v[2], v[3]);
        }
        *reinterpret_cast<float4*>(out + ((size_t)ch * a.LH + Y) * a.LW + X) = o;
    }
}

}  // namespace

int vti_launch_k1(vti_handle* h, const uint8_t* frames, int B, float* net_in, cudaStream_t s) {
    K1Args a;
    a.frames = frames;
    a.out = net_in;
    a.tap_x_idx = h->d_tap_x_idx;
    a.tap_x_a = h->d_tap_x_a;
    a.tap_y_i = h->d_tap_y_i;
    a.tap_y_b = h->d_tap_y_b;
    a.und_lut = h->d_und_lut;
    a.h = h->p.frame_h; a.w = h->p.frame_w;
    a.new_h = h->g.new_h; a.new_w = h->g.new_w;
    a.top = h->g.top; a.left = h->g.left;
    a.LH = h->g.LH; a.LW = h->g.LW;
    a.mode = h->resize_mode;
    a.flip = h->p.channel_flip;
    dim3 grid((a.LW + TX - 1) / TX, (a.LH + TY - 1) / TY, B);
    if (h->p.undistort)
        k1_letterbox_kernel<true><<<grid, K1_THREADS, 0, s>>>(a);
    else
        k1_letterbox_kernel<false><<<grid, K1_THREADS, 0, s>>>(a);
    h->launches++;
    VTI_CUDA(cudaGetLastError());
    return VTI_OK;
}
