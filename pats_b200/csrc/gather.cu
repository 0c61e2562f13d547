// Feature gathers between the CNN stems and the attention layers -- integer index + copy work, bit-exact.
// Replaces, from zju3dv/pats:
//   models/second_layer.py:71-80   12x12 grid sampling of the three stem maps (2x2 average fused, never materialising the
//                                  pooled maps)
//   models/third_layer.py:119-146  8x8 window unfold around each level-2 point for both images, fused with the positional
//                                  encoding add and the rubbish token (one write of the [K,C,65] GNN input instead of
//                                  three gathers with [K*64,C] expanded int64 indices, a permute, an add and a cat)
// HBM-bound: every output element is written once with coalesced stores; inputs are read through L2.
#include "common.cuh"

namespace pats {

// out [N, C0+C1+C2, R*R]; one thread per output element (consecutive threads -> consecutive grid points)
__global__ void __launch_bounds__(256) grid_sample12_kernel(const float *__restrict__ f0, const float *__restrict__ f1,
                                                            const float *__restrict__ f2, int N, int C0, int C1, int C2, int R,
                                                            float *__restrict__ out) {
    const int RR = R * R, Ct = C0 + C1 + C2;
    const long long total = (long long)N * Ct * RR;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int pos = (int)(e % RR);
        const long long nc = e / RR;
        const int c = (int)(nc % Ct), n = (int)(nc / Ct);
        const int i = pos / R, j = pos - i * R;
        float v;
        if (c < C0 + C1) {
            const int lvl = c < C0 ? 0 : 1, s = lvl == 0 ? 4 : 2, S = s * R;
            const float *f = lvl == 0 ? f0 + ((size_t)n * C0 + c) * S * S : f1 + ((size_t)n * C1 + (c - C0)) * S * S;
            const int y = (int)(((float)i + 0.5f) * (float)s), x = (int)(((float)j + 0.5f) * (float)s);  // pooled coordinates (:76)
            // AvgPool2d(2, stride=1, padding=1), count_include_pad: window rows y-1..y, cols x-1..x, ATen's summation order, /4
            float acc = 0.f;
#pragma unroll
            for (int dh = -1; dh <= 0; ++dh)
#pragma unroll
                for (int dw = -1; dw <= 0; ++dw) {
                    const int h = y + dh, w = x + dw;
                    if (h >= 0 && h < S && w >= 0 && w < S) acc = __fadd_rn(acc, __ldg(f + h * S + w));
                }
            v = __fdiv_rn(acc, 4.0f);
        } else {
            v = __ldg(f2 + ((size_t)n * C2 + (c - C0 - C1)) * RR + pos);
        }
        out[e] = v;
    }
}

// one CTA per point k: out [K, C, 65]
__global__ void __launch_bounds__(256) third_unfold_kernel(const float *__restrict__ feat, int P, int C, int M,
                                                           const float *__restrict__ mkpts, const float *__restrict__ b_ids,
                                                           int clamp96, const float *__restrict__ kenc,
                                                           const float *__restrict__ rubbish, const float *__restrict__ mkpts0,
                                                           float *__restrict__ out, int *bad) {
    const int k = blockIdx.x;
    float px = mkpts[2 * k], py = mkpts[2 * k + 1];
    if (clamp96) {  // third_layer.py:127-128
        px = px >= 96.f ? 96.f : px, py = py >= 96.f ? 96.f : py;
        px = px <= 0.f ? 0.f : px, py = py <= 0.f ? 0.f : py;
    }
    const long long sx = (long long)rintf(__fdiv_rn(px, 4.0f)) * 4, sy = (long long)rintf(__fdiv_rn(py, 4.0f)) * 4;  // round-half-even (:122/:129)
    const float b = b_ids[k];
    const long long lx = (long long)rintf(__fdiv_rn(mkpts0[2 * k], 4.0f)) * 4, ly = (long long)rintf(__fdiv_rn(mkpts0[2 * k + 1], 4.0f)) * 4;
    const long long x2 = (long long)rintf(__fdiv_rn((float)lx, 8.0f)), y2 = (long long)rintf(__fdiv_rn((float)ly, 8.0f));
    const long long i2 = (long long)__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(b, 12.f), 12.f), (float)(y2 * 12)), (float)x2);  // :143
    const long long total = (long long)P * M * M, MM = (long long)M * M;
    const float fM = (float)M;
    for (int e = threadIdx.x; e < C * 65; e += blockDim.x) {
        const int c = e / 65, pos = e - c * 65;
        float v = 0.f;
        if (pos < 64) {
            const int wy = pos >> 3, wx = pos & 7;
            const float x0 = __fadd_rn(__fsub_rn(__fadd_rn((float)(sx / 2), (float)wx), 4.0f), 2.0f);   // :123
            const float y0 = __fadd_rn(__fsub_rn(__fadd_rn((float)(sy / 2), (float)wy), 4.0f), 2.0f);
            const long long idx = (long long)__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(b, fM), fM), __fmul_rn(y0, fM)), x0);  // :125
            if (idx < 0 || idx >= total) {
                if (c == 0) atomicAdd(bad, 1);
            } else {
                const long long bb = idx / MM, rem = idx - bb * MM;
                v = __ldg(feat + ((size_t)bb * C + c) * MM + rem);
            }
            v = __fadd_rn(v, __ldg(kenc + c * 64 + pos));
        } else {
            if (i2 < 0 || i2 >= (long long)P * 144) {
                if (c == 0) atomicAdd(bad, 1);
            } else {
                v = __ldg(rubbish + ((size_t)(i2 / 144) * C + c) * 144 + i2 % 144);
            }
        }
        out[(size_t)k * C * 65 + e] = v;
    }
}

}  // namespace pats

using namespace pats;

PATS_API int pats_grid_sample12_f32(const float *f0, const float *f1, const float *f2, int N, int C0, int C1, int C2, int row_num,
                                    float *out, void *stream) {
    if (N < 0 || C0 < 0 || C1 < 0 || C2 < 0 || row_num <= 0) return invalid("grid_sample12: bad sizes");
    const long long total = (long long)N * (C0 + C1 + C2) * row_num * row_num;
    if (total == 0) return PATS_OK;
    if ((C0 && !f0) || (C1 && !f1) || (C2 && !f2) || !out) return invalid("grid_sample12: null pointer");
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    grid_sample12_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(f0, f1, f2, N, C0, C1, C2, row_num, out);
    PATS_LAUNCH_CHECK("grid_sample12_kernel");
    return PATS_OK;
}

PATS_API int pats_third_unfold_f32(const float *feat, int P, int C, int M, const float *mkpts_c, const float *b_ids, int K,
                                   int clamp96, const float *kenc, const float *rubbish, const float *mkpts0_c, float *out,
                                   int *bad_index, void *stream) {
    if (P <= 0 || C <= 0 || M <= 0 || K < 0) return invalid("third_unfold: bad sizes");
    if (K == 0) return PATS_OK;
    if (!feat || !mkpts_c || !b_ids || !kenc || !rubbish || !mkpts0_c || !out || !bad_index) return invalid("third_unfold: null pointer");
    third_unfold_kernel<<<K, 256, 0, as_stream(stream)>>>(feat, P, C, M, mkpts_c, b_ids, clamp96, kenc, rubbish, mkpts0_c, out, bad_index);
    PATS_LAUNCH_CHECK("third_unfold_kernel");
    return PATS_OK;
}
