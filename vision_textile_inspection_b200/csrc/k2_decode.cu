// K2 -- YOLOv8 head tail: class sigmoid + confidence filter (warp ballot, warp-aggregated compaction) and, for the
// surviving anchors only, DFL softmax-expectation + dist2bbox + xywh->xyxy.
//
// Replaces (SURVEY.md 8a U3 + the prelude of U4): ultralytics Detect._inference / DFL.forward / dist2bbox and the
// `xc = amax(cls) > conf` selection of ops.non_max_suppression, reached from /root/reference/measurement.py:208-210.
// Spec: oracle/post_spec.py decode_spec / candidates_spec -- every float op below is the same single-rounded
// IEEE op in the same order (no FMA contraction), so boxes and scores are bit-identical to the spec.
//
// One thread per anchor; a warp covers 32 consecutive anchors of one level so the class-logit reads are coalesced
// 128-byte lines.  Only candidate lanes touch their 64 box logits (strided by the level plane).  Candidates are
// appended with one atomicAdd per warp; the order inside the list is irrelevant because K3 sorts by the composite
// key (score bits, ~anchor) which reproduces torchvision's stable descending sort over ascending anchor order.
#include "vti_internal.h"

int vti_k3_cap_pad(int cap);

namespace {

constexpr int K2_THREADS = 256;

struct K2Args {
    const float* lvl[3];
    int lvl_h[3], lvl_w[3];
    int a_begin[4];            // anchor offset of each level, a_begin[3] = A
    int nc, A, cap, cap_pad;
    float conf;
    int32_t* cand_count;       // [B]
    unsigned long long* cand_key;   // [B][cap]
    float4* cand_box;          // [B][A]
};

__global__ void __launch_bounds__(K2_THREADS) k2_decode_kernel(const K2Args a) {
    const int b = blockIdx.y;
    // blocks are laid out per level so that a warp never straddles two levels
    int blk = blockIdx.x, l = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int nblk = (a.a_begin[l + 1] - a.a_begin[l] + K2_THREADS - 1) / K2_THREADS;
        if (blk >= nblk) { blk -= nblk; ++l; }
    }
    const int plane = a.lvl_h[l] * a.lvl_w[l];
    const int i = blk * K2_THREADS + threadIdx.x;       // index inside the level
    const bool valid = i < plane;
    const float* __restrict__ base = a.lvl[l] + (size_t)b * (64 + a.nc) * plane;

    float best = 0.0f;
    int cls = 0;
    if (valid) {
        for (int c = 0; c < a.nc; ++c) {
            const float pr = vti_sigmoid_spec(__ldg(base + (size_t)(64 + c) * plane + i));
            if (c == 0 || pr > best) { best = pr; cls = c; }     // first maximum wins, as torch.max
        }
    }
    const bool cand = valid && (best > a.conf);
    const unsigned m = __ballot_sync(0xffffffffu, cand);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(a.cand_count + b, __popc(m));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (!cand) return;

    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __ldg(base + (size_t)(side * 16 + k) * plane + i);
        float mx = v[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) mx = fmaxf(mx, v[k]);
        float S = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            v[k] = vti_exp_spec(__fsub_rn(v[k], mx));
            S = __fadd_rn(S, v[k]);
        }
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = __fadd_rn(acc, __fmul_rn((float)k, __fdiv_rn(v[k], S)));
        d[side] = acc;
    }
    const int gy = i / a.lvl_w[l], gx = i - gy * a.lvl_w[l];
    const float ax = __fadd_rn((float)gx, 0.5f), ay = __fadd_rn((float)gy, 0.5f);
    const float x1 = __fsub_rn(ax, d[0]), y1 = __fsub_rn(ay, d[1]);
    const float x2 = __fadd_rn(ax, d[2]), y2 = __fadd_rn(ay, d[3]);
    const float st = (float)(8 << l);
    const float cx = __fmul_rn(__fdiv_rn(__fadd_rn(x1, x2), 2.0f), st);
    const float cy = __fmul_rn(__fdiv_rn(__fadd_rn(y1, y2), 2.0f), st);
    const float w = __fmul_rn(__fsub_rn(x2, x1), st);
    const float hh = __fmul_rn(__fsub_rn(y2, y1), st);
    const float hw2 = __fdiv_rn(w, 2.0f), hh2 = __fdiv_rn(hh, 2.0f);
    const float4 box = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));

    const int anchor = a.a_begin[l] + i;
    a.cand_box[(size_t)b * a.A + anchor] = box;
    const int slot = slot0 + __popc(m & ((1u << lane) - 1u));
    if (slot < a.cap) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(best) << 32) |
                                       ((unsigned long long)(0xFFFFFFu - (unsigned)anchor) << 8) | (unsigned)cls;
        a.cand_key[(size_t)b * a.cap_pad + slot] = key;
    }
}

}  // namespace

int vti_launch_k2(vti_handle* h, const float* p3, const float* p4, const float* p5, int B, cudaStream_t s) {
    K2Args a;
    a.lvl[0] = p3; a.lvl[1] = p4; a.lvl[2] = p5;
    int off = 0, nblk = 0;
    for (int l = 0; l < 3; ++l) {
        a.lvl_h[l] = h->g.lvl_h[l];
        a.lvl_w[l] = h->g.lvl_w[l];
        a.a_begin[l] = off;
        off += a.lvl_h[l] * a.lvl_w[l];
        nblk += (a.lvl_h[l] * a.lvl_w[l] + K2_THREADS - 1) / K2_THREADS;
    }
    a.a_begin[3] = off;
    a.nc = h->p.nc; a.A = h->g.A; a.cap = h->g.max_candidates; a.cap_pad = vti_k3_cap_pad(a.cap);
    a.conf = h->p.conf;
    a.cand_count = h->d_cand_count;
    a.cand_key = h->d_cand_key;
    a.cand_box = h->d_cand_box;
    VTI_CUDA(cudaMemsetAsync(h->d_cand_count, 0, sizeof(int32_t) * (h->p.max_batch + 1), s));   // + K4 unit counter
    k2_decode_kernel<<<dim3(nblk, B), K2_THREADS, 0, s>>>(a);
    h->launches++;
    VTI_CUDA(cudaGetLastError());
    return VTI_OK;
}
