// Patch subdivision for sm_100a -- replaces setup/library.cpp:47-66 (tensor_resize),
// utils/utils.py:1300-1318 (origin_extract) and utils/utils.py:1343-1393 (Compute_imgs) of zju3dv/pats.
//
// These are HBM-bound byte / f32 movers: one launch per call, bounds read on the device (the
// reference does 5 .item() syncs and ~5 launches PER PATCH), coalesced 128-bit accesses where the
// layout allows, grids of K*C (or P) CTAs >> 148 SMs.
#include "common.cuh"

namespace pats {

// ---- bilinear lerp with a selectable rounding recipe ------------------------------------------------
// ATen's CUDA kernel (UpSampleBilinear2d.cu) evaluates
//     h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d)
// and lets nvcc contract it.  Variants 1..9 spell out every contraction (inner, outer in {none, fma on the
// first product, fma on the second product}); measured on B200 against torch 2.11 (tests/test_gpu_subdivide.py,
// profiles/r01_resize_variants.json) variant 5 -- fma(h0, fma(w0, a, w1*b), h1*t2) -- is bit-identical to ATen
// on 8.3 M outputs, so it is the shipping recipe (kShipVariant).  Variant 0 is the bare expression as compiled.
template <int VARIANT>
__device__ __forceinline__ float lerp2(float h0, float h1, float w0, float w1, float a, float b, float c, float d) {
    if (VARIANT == 0) {
        return h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d);
    } else {
        constexpr int inner = (VARIANT - 1) / 3, outer = (VARIANT - 1) % 3;
        float t1, t2;
        if (inner == 0) {
            t1 = __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b));
            t2 = __fadd_rn(__fmul_rn(w0, c), __fmul_rn(w1, d));
        } else if (inner == 1) {
            t1 = __fmaf_rn(w0, a, __fmul_rn(w1, b));
            t2 = __fmaf_rn(w0, c, __fmul_rn(w1, d));
        } else {
            t1 = __fmaf_rn(w1, b, __fmul_rn(w0, a));
            t2 = __fmaf_rn(w1, d, __fmul_rn(w0, c));
        }
        if (outer == 0) return __fadd_rn(__fmul_rn(h0, t1), __fmul_rn(h1, t2));
        if (outer == 1) return __fmaf_rn(h0, t1, __fmul_rn(h1, t2));
        return __fmaf_rn(h1, t2, __fmul_rn(h0, t1));
    }
}

constexpr int kShipVariant = 5;

struct Crop {
    long long y0, x0, h, w, img;
    bool ok;
};

__device__ __forceinline__ Crop read_crop(const int64_t *bound, int k, int B, int Hp, int Wp) {
    Crop c;
    const long long y1 = bound[k * 5 + 1], x1 = bound[k * 5 + 3];
    c.y0 = bound[k * 5 + 0];
    c.x0 = bound[k * 5 + 2];
    c.img = bound[k * 5 + 4] / 10000;  // library.cpp:54-55
    c.h = y1 - c.y0;                   // rows  [y0, y1)      library.cpp:56-57
    c.w = x1 - c.x0 + 1;               // cols  [x0, x1]      library.cpp:58-59
    c.ok = c.img >= 0 && c.img < B && c.h > 0 && c.w > 0 && c.y0 >= 0 && c.x0 >= 0 && c.y0 + c.h <= Hp && c.x0 + c.w <= Wp;
    return c;
}

// align_corners=True source coordinates (ATen UpSample.h area_pixel_compute_scale / _source_index)
struct Axis {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Axis axis_coord(float scale, int o, long long size) {
    Axis a;
    const float r = __fmul_rn(scale, (float)o);
    a.i0 = (int)r;
    a.i1 = a.i0 + ((a.i0 < size - 1) ? 1 : 0);
    a.l1 = __fsub_rn(r, (float)a.i0);
    a.l0 = __fsub_rn(1.0f, a.l1);
    return a;
}
__device__ __forceinline__ float axis_scale(long long in, int out) {
    return out > 1 ? __fdiv_rn((float)(in - 1), (float)(out - 1)) : 0.f;
}

// One CTA per (patch, channel): input [B,C,Hp,Wp] f32 -> out [K,C,oh,ow] f32.
template <int VARIANT>
__global__ void __launch_bounds__(256) tensor_resize_kernel(const float *__restrict__ input, int B, int C, int Hp, int Wp,
                                                            const int64_t *__restrict__ bound, int oh, int ow,
                                                            float *__restrict__ out, int *bad_rows) {
    const int k = blockIdx.x / C, ch = blockIdx.x % C;
    const Crop c = read_crop(bound, k, B, Hp, Wp);
    float *dst = out + (size_t)blockIdx.x * oh * ow;
    const int total = oh * ow;
    if (!c.ok) {
        for (int e = threadIdx.x; e < total; e += blockDim.x) dst[e] = 0.f;
        if (ch == 0 && threadIdx.x == 0 && bad_rows) atomicAdd(bad_rows, 1);
        return;
    }
    const float *src = input + (((size_t)c.img * C + ch) * Hp + c.y0) * Wp + c.x0;
    const float sh = axis_scale(c.h, oh), sw = axis_scale(c.w, ow);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int oy = e / ow, ox = e - oy * ow;
        const Axis ay = axis_coord(sh, oy, c.h), ax = axis_coord(sw, ox, c.w);
        const float *r0 = src + (size_t)ay.i0 * Wp, *r1 = src + (size_t)ay.i1 * Wp;
        dst[e] = lerp2<VARIANT>(ay.l0, ay.l1, ax.l0, ax.l1, __ldg(r0 + ax.i0), __ldg(r0 + ax.i1), __ldg(r1 + ax.i0),
                                __ldg(r1 + ax.i1));
    }
}

// ---- origin_extract: strided window copy with the widest vector the layout allows ---------------------
template <class V>
__global__ void __launch_bounds__(256) origin_extract_kernel(const V *__restrict__ src, V *__restrict__ dst, int height,
                                                             int width, int ps_v, int win_v, int win_rows,
                                                             size_t src_row_v, size_t src_plane_v) {
    // blockIdx.x = (b*C + c) * P + p ; all *_v quantities are in units of V
    const int P = height * width;
    const int p = blockIdx.x % P;
    const size_t bc = blockIdx.x / P;
    const int i = p / width, j = p % width;
    const V *s = src + bc * src_plane_v + (size_t)i * (win_rows / 3) * src_row_v + (size_t)j * ps_v;
    V *d = dst + (size_t)blockIdx.x * win_rows * win_v;
    const int total = win_rows * win_v;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int r = e / win_v, x = e - r * win_v;
        d[e] = __ldg(s + (size_t)r * src_row_v + x);
    }
}

// ---- Compute_imgs bound / scale arithmetic (utils/utils.py:1357-1372) ----------------------------------
struct BoundOut {
    long long b0, b1, b2, b3;
    float xs, ys, av0, av1;
};
__device__ __forceinline__ BoundOut bounds_one(float x_scale, float y_scale, float ay, float ax, int height, int width,
                                               int ps, int margin) {
    const float fps = (float)ps, fm = (float)margin;
    const float hy = __fdiv_rn(__fmul_rn(y_scale, 3.0f), 2.0f), hx = __fdiv_rn(__fmul_rn(x_scale, 3.0f), 2.0f);
    float b0 = __fadd_rn(__fmul_rn(__fsub_rn(ay, hy), fps), fm);
    float b1 = __fadd_rn(__fmul_rn(__fadd_rn(ay, hy), fps), fm);
    float b2 = __fadd_rn(__fmul_rn(__fsub_rn(ax, hx), fps), fm);
    float b3 = __fadd_rn(__fmul_rn(__fadd_rn(ax, hx), fps), fm);
    b0 = (b0 >= 0.f) ? b0 : 0.f;  // :1361 (NaN -> 0 like torch.where)
    b1 = (b1 >= 0.f) ? b1 : 0.f;
    b2 = (b2 >= 0.f) ? b2 : 0.f;
    b3 = (b3 >= 0.f) ? b3 : 0.f;
    b1 = (b1 < (float)(ps * height + 2 * margin)) ? b1 : (float)(ps * height - 1);  // :1362, board[1] :1351
    b3 = (b3 < (float)(ps * width + 2 * margin)) ? b3 : (float)(ps * width);        // :1363, board[3]
    BoundOut o;
    // `tensor / float(3 * patch_scale)`: on CUDA tensors ATen divides by a python scalar as a MULTIPLICATION by its reciprocal
    // (BinaryDivTrueKernel.cu: inv_b = 1.0 / b in double, rounded to f32), which differs from the true quotient in the last
    // bit for about a third of the values when the divisor is 96.  The reference runs on CUDA tensors (evaluate.py:26-28), so
    // that is the arithmetic to reproduce (seen live: tests/test_gpu_live_forward.py); the CPU oracle keeps the true division.
    const float inv = (float)(1.0 / (double)(3 * ps));
    o.xs = __fmul_rn(__fadd_rn(__fsub_rn(b1, b0), 1.0f), inv);  // :1364 ("x" from the y extent)
    o.ys = __fmul_rn(__fadd_rn(__fsub_rn(b3, b2), 1.0f), inv);  // :1365
    o.b0 = (long long)b0, o.b1 = (long long)b1, o.b2 = (long long)b2, o.b3 = (long long)b3;  // :1366 trunc
    o.av1 = __fadd_rn(__fsub_rn(__fdiv_rn((float)(o.b1 + o.b0), 2.0f), fm), 0.5f);  // :1368
    o.av0 = __fadd_rn(__fsub_rn(__fdiv_rn((float)(o.b2 + o.b3), 2.0f), fm), 0.5f);  // :1369
    return o;
}

__global__ void compute_bounds_kernel(const float *__restrict__ x_scale, const float *__restrict__ y_scale,
                                      const float *__restrict__ avg, int total, int height, int width, int ps, int margin,
                                      int64_t *__restrict__ bound, float *__restrict__ xs_new, float *__restrict__ ys_new,
                                      float *__restrict__ avg_new) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const BoundOut o = bounds_one(x_scale[k], y_scale[k], avg[2 * k], avg[2 * k + 1], height, width, ps, margin);
    bound[4 * k + 0] = o.b0, bound[4 * k + 1] = o.b1, bound[4 * k + 2] = o.b2, bound[4 * k + 3] = o.b3;
    xs_new[2 * k] = o.xs, xs_new[2 * k + 1] = 1.0f;  // :1380
    ys_new[2 * k] = o.ys, ys_new[2 * k + 1] = 1.0f;  // :1381
    avg_new[2 * k] = o.av0, avg_new[2 * k + 1] = o.av1;
}

// ---- fused Compute_imgs ----------------------------------------------------------------------------------
// Kernel A: bounds for every patch + ordered compaction of the matched ones (single CTA: B*n is a few
// hundred).  bound5 rows follow the row-major order of the boolean mask (utils.py:1376,1382).
__global__ void __launch_bounds__(1024) compute_imgs_plan_kernel(const float *__restrict__ x_scale, const float *__restrict__ y_scale,
                                                                 const float *__restrict__ avg, const uint8_t *__restrict__ nomatch,
                                                                 int B, int n, int height, int width, int ps, int margin,
                                                                 int64_t *__restrict__ bound5, float *__restrict__ xs_new,
                                                                 float *__restrict__ ys_new, float *__restrict__ avg_new,
                                                                 int capacity, int *count) {
    pdl_prologue();
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int total = B * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < total; start += blockDim.x) {
        const int k = start + threadIdx.x;
        bool matched = false;
        BoundOut o;
        if (k < total) {
            o = bounds_one(x_scale[k], y_scale[k], avg[2 * k], avg[2 * k + 1], height, width, ps, margin);
            xs_new[2 * k] = o.xs, xs_new[2 * k + 1] = 1.0f;
            ys_new[2 * k] = o.ys, ys_new[2 * k + 1] = 1.0f;
            avg_new[2 * k] = o.av0, avg_new[2 * k + 1] = o.av1;
            matched = nomatch[k] == 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, matched);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        const int pos = before + __popc(m & ((1u << lane) - 1u));
        if (matched && pos < capacity) {
            const int img = k / n, patch = k % n;
            bound5[(size_t)pos * 5 + 0] = o.b0, bound5[(size_t)pos * 5 + 1] = o.b1;
            bound5[(size_t)pos * 5 + 2] = o.b2, bound5[(size_t)pos * 5 + 3] = o.b3;
            bound5[(size_t)pos * 5 + 4] = (long long)img * 10000 + patch;  // :1373-1376
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// Kernel B: left windows straight from the un-padded NHWC image, zero outside (F.pad + origin_extract +
// permute + mask of utils.py:1353,1383-1384 in one pass).  V is the 16-byte vector; every 32-pixel
// segment of a window row is entirely inside or outside the image, so validity is per segment.
template <class V>
__global__ void __launch_bounds__(256) left_windows_kernel(const V *__restrict__ left, V *__restrict__ out,
                                                           const int64_t *__restrict__ bound5, const int *__restrict__ count,
                                                           int H, int W, int ps, int width, int seg_v /* V per ps pixels */) {
    pdl_prologue();
    const int p = blockIdx.x;
    if (p >= *count) return;
    const long long seq = bound5[(size_t)p * 5 + 4];
    const int img = (int)(seq / 10000), patch = (int)(seq % 10000);
    const int i = patch / width, j = patch % width;
    const int win = 3 * ps, row_v = 3 * seg_v;
    const size_t img_row_v = (size_t)(W / ps) * seg_v;
    V zero;
    memset(&zero, 0, sizeof(V));
    for (int e = threadIdx.x; e < win * row_v; e += blockDim.x) {
        const int r = e / row_v, xv = e - r * row_v;
        const int y = ps * (i - 1) + r;
        const int seg = xv / seg_v, jj = j - 1 + seg;  // patch column this segment lies in
        V v = zero;
        if (y >= 0 && y < H && jj >= 0 && jj < W / ps)
            v = __ldg(left + ((size_t)img * H + y) * img_row_v + (size_t)jj * seg_v + (xv - seg * seg_v));
        out[(size_t)p * win * row_v + e] = v;
    }
}

// Kernel C: right patches, crop + bilinear from the un-padded NHWC image (zero outside = the margin
// padding of utils.py:1352), written as [P,3,oh,ow] f32.
template <class T>
__global__ void __launch_bounds__(256) right_patches_kernel(const T *__restrict__ right, float *__restrict__ out,
                                                            const int64_t *__restrict__ bound5, const int *__restrict__ count,
                                                            int B, int H, int W, int margin, int oh, int ow, int *bad_rows) {
    // launched behind left_windows_kernel, which it does not depend on (both read the plan kernel's bounds): runs
    // concurrently with it and waits for it at the END, see count_rows_kernel in regroup.cu
    pdl_launch_dependents();
    // grid (P, bands): each CTA produces a band of output rows of one patch (P alone would leave most SMs idle)
    const int p = blockIdx.x;
    if (p >= *count) {
        pdl_wait();
        return;
    }
    const int Hp = H + 2 * margin, Wp = W + 2 * margin;
    const Crop c = read_crop(bound5, p, B, Hp, Wp);
    float *dst = out + (size_t)p * 3 * oh * ow;
    const int total = oh * ow;
    const int rows_per = (oh + gridDim.y - 1) / gridDim.y;
    const int e0 = blockIdx.y * rows_per * ow, e1 = min(total, e0 + rows_per * ow);  // the band as a flat range (zero fill below)
    if (!c.ok) {
        for (int ch = 0; ch < 3; ++ch)
            for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) dst[(size_t)ch * total + e] = 0.f;
        if (blockIdx.y == 0 && threadIdx.x == 0 && bad_rows) atomicAdd(bad_rows, 1);
        pdl_wait();
        return;
    }
    const T *img = right + (size_t)c.img * H * W * 3;
    const float sh = axis_scale(c.h, oh), sw = axis_scale(c.w, ow);
    // A thread owns output columns (its x coordinates, weights and validity are computed once) and walks down the band's
    // rows: no division per element, two row pointers per row, 32-bit offsets.
    const int row0 = blockIdx.y * rows_per, row1 = min(oh, row0 + rows_per);
    const int lanes_x = min(ow, (int)blockDim.x);          // threads along x
    const int ty = threadIdx.x / lanes_x, ny = max(1, (int)blockDim.x / lanes_x);
    if (ty < ny) {
        for (int ox = threadIdx.x - ty * lanes_x; ox < ow; ox += lanes_x) {
            const Axis ax = axis_coord(sw, ox, c.w);
            const int xa = (int)c.x0 + ax.i0 - margin, xb = (int)c.x0 + ax.i1 - margin;
            const bool vxa = xa >= 0 && xa < W, vxb = xb >= 0 && xb < W;
            const int xa3 = vxa ? xa * 3 : 0, xb3 = vxb ? xb * 3 : 0;
            for (int oy = row0 + ty; oy < row1; oy += ny) {
                const Axis ay = axis_coord(sh, oy, c.h);
                const int ya = (int)c.y0 + ay.i0 - margin, yb = (int)c.y0 + ay.i1 - margin;
                const bool vya = ya >= 0 && ya < H, vyb = yb >= 0 && yb < H;
                const T *ra = img + (size_t)(vya ? ya : 0) * W * 3, *rb = img + (size_t)(vyb ? yb : 0) * W * 3;
                float *o = dst + oy * ow + ox;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float a = (vya && vxa) ? (float)__ldg(ra + xa3 + ch) : 0.f;
                    const float b = (vya && vxb) ? (float)__ldg(ra + xb3 + ch) : 0.f;
                    const float cc = (vyb && vxa) ? (float)__ldg(rb + xa3 + ch) : 0.f;
                    const float d = (vyb && vxb) ? (float)__ldg(rb + xb3 + ch) : 0.f;
                    o[(size_t)ch * total] = lerp2<kShipVariant>(ay.l0, ay.l1, ax.l0, ax.l1, a, b, cc, d);
                }
            }
        }
    }
    pdl_wait();
}

template <int V>
static void launch_resize(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K, int oh, int ow,
                          float *out, int *bad, cudaStream_t st) {
    tensor_resize_kernel<V><<<K * C, 256, 0, st>>>(input, B, C, Hp, Wp, bound, oh, ow, out, bad);
}

}  // namespace pats

using namespace pats;

PATS_API int pats_tensor_resize_f32_variant(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K,
                                            int out_h, int out_w, float *out, int *bad_rows, int variant, void *stream) {
    if (B <= 0 || C <= 0 || Hp <= 0 || Wp <= 0 || K < 0 || out_h <= 0 || out_w <= 0) return invalid("tensor_resize: bad sizes");
    if (K == 0) return PATS_OK;
    if (!input || !bound || !out) return invalid("tensor_resize: null pointer");
    if ((long long)K * C > 0x7fffffffLL) return invalid("tensor_resize: too many patches");
    cudaStream_t st = as_stream(stream);
    switch (variant) {
        case 0: launch_resize<0>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 1: launch_resize<1>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 2: launch_resize<2>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 3: launch_resize<3>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 4: launch_resize<4>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 5: launch_resize<5>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 6: launch_resize<6>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 7: launch_resize<7>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 8: launch_resize<8>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        case 9: launch_resize<9>(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, st); break;
        default: return invalid("tensor_resize: variant %d out of range", variant);
    }
    PATS_LAUNCH_CHECK("tensor_resize_kernel");
    return PATS_OK;
}

PATS_API int pats_tensor_resize_f32(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K, int out_h,
                                    int out_w, float *out, int *bad_rows, void *stream) {
    return pats_tensor_resize_f32_variant(input, B, C, Hp, Wp, bound, K, out_h, out_w, out, bad_rows, kShipVariant, stream);
}

PATS_API int pats_tensor_resize_f32_host(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K,
                                         int out_h, int out_w, float *out) {
    if (B <= 0 || C <= 0 || Hp <= 0 || Wp <= 0 || K < 0 || out_h <= 0 || out_w <= 0) return invalid("tensor_resize (host): bad sizes");
    if (K == 0) return PATS_OK;
    if (!input || !bound || !out) return invalid("tensor_resize (host): null pointer");
    const size_t nin = (size_t)B * C * Hp * Wp * 4, nb = (size_t)K * 5 * 8, nout = (size_t)K * C * out_h * out_w * 4;
    void *d_in = nullptr, *d_b = nullptr, *d_out = nullptr, *d_bad = nullptr;
    int rc = PATS_OK, bad = 0;
    cudaError_t e;
    if ((e = cudaMalloc(&d_in, nin)) != cudaSuccess || (e = cudaMalloc(&d_b, nb)) != cudaSuccess ||
        (e = cudaMalloc(&d_out, nout)) != cudaSuccess || (e = cudaMalloc(&d_bad, 4)) != cudaSuccess) {
        rc = cuda_fail(e, "cudaMalloc (tensor_resize host)");
    } else {
        cudaMemcpyAsync(d_in, input, nin, cudaMemcpyHostToDevice, 0);
        cudaMemcpyAsync(d_b, bound, nb, cudaMemcpyHostToDevice, 0);
        cudaMemsetAsync(d_bad, 0, 4, 0);
        rc = pats_tensor_resize_f32((const float *)d_in, B, C, Hp, Wp, (const int64_t *)d_b, K, out_h, out_w, (float *)d_out,
                                    (int *)d_bad, nullptr);
        if (rc == PATS_OK) {
            cudaMemcpyAsync(out, d_out, nout, cudaMemcpyDeviceToHost, 0);
            cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, 0);
            e = cudaStreamSynchronize(0);
            if (e != cudaSuccess) rc = cuda_fail(e, "tensor_resize (host)");
            else if (bad) {
                set_error("tensor_resize: %d bound row(s) describe an empty or out-of-range crop", bad);
                rc = PATS_E_BAD_CROP;
            }
        }
    }
    cudaFree(d_in), cudaFree(d_b), cudaFree(d_out), cudaFree(d_bad);
    return rc;
}

PATS_API int pats_origin_extract(const void *left, int elem, int B, int C, int height, int width, int ps, void *out,
                                 void *stream) {
    if (elem <= 0 || B < 0 || C <= 0 || height <= 0 || width <= 0 || ps <= 0) return invalid("origin_extract: bad sizes");
    if (B == 0) return PATS_OK;
    if (!left || !out) return invalid("origin_extract: null pointer");
    const long long blocks = (long long)B * C * height * width;
    if (blocks > 0x7fffffffLL) return invalid("origin_extract: too many windows");
    cudaStream_t st = as_stream(stream);
    const size_t row_b = (size_t)ps * (width + 2) * elem, plane_b = row_b * ps * (height + 2);
    const size_t ps_b = (size_t)ps * elem, win_b = 3 * ps_b;
    const int win_rows = 3 * ps;
    const uintptr_t al = (uintptr_t)left | (uintptr_t)out | row_b | ps_b | plane_b;
    if (al % 16 == 0) {
        origin_extract_kernel<uint4><<<(int)blocks, 256, 0, st>>>((const uint4 *)left, (uint4 *)out, height, width, (int)(ps_b / 16),
                                                                 (int)(win_b / 16), win_rows, row_b / 16, plane_b / 16);
    } else if (al % 4 == 0) {
        origin_extract_kernel<uint32_t><<<(int)blocks, 256, 0, st>>>((const uint32_t *)left, (uint32_t *)out, height, width,
                                                                    (int)(ps_b / 4), (int)(win_b / 4), win_rows, row_b / 4, plane_b / 4);
    } else {
        origin_extract_kernel<uint8_t><<<(int)blocks, 256, 0, st>>>((const uint8_t *)left, (uint8_t *)out, height, width, (int)ps_b,
                                                                   (int)win_b, win_rows, row_b, plane_b);
    }
    PATS_LAUNCH_CHECK("origin_extract_kernel");
    return PATS_OK;
}

PATS_API int pats_compute_bounds_f32(const float *x_scale, const float *y_scale, const float *average_point, int B, int height,
                                     int width, int ps, int margin, int64_t *bound, float *x_scale_new, float *y_scale_new,
                                     float *average_new, void *stream) {
    if (B < 0 || height <= 0 || width <= 0 || ps <= 0 || margin < 0) return invalid("compute_bounds: bad sizes");
    const int total = B * height * width;
    if (total == 0) return PATS_OK;
    if (!x_scale || !y_scale || !average_point || !bound || !x_scale_new || !y_scale_new || !average_new)
        return invalid("compute_bounds: null pointer");
    compute_bounds_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(x_scale, y_scale, average_point, total, height, width, ps,
                                                                            margin, bound, x_scale_new, y_scale_new, average_new);
    PATS_LAUNCH_CHECK("compute_bounds_kernel");
    return PATS_OK;
}

PATS_API int pats_compute_imgs(const float *x_scale, const float *y_scale, const float *average_point,
                               const uint8_t *if_nomatching, const void *left, const void *right, int elem, int B, int height,
                               int width, int ps, int margin, void *new_left, float *new_right, int64_t *bound5,
                               float *x_scale_new, float *y_scale_new, float *average_new, int capacity, int *count,
                               int *bad_rows, void *stream) {
    if (B <= 0 || height <= 0 || width <= 0 || ps <= 0 || margin < 0 || capacity < 0) return invalid("compute_imgs: bad sizes");
    if (elem != 1 && elem != 4) return invalid("compute_imgs: images must be uint8 (elem=1) or float32 (elem=4)");
    if (width * height >= 10000) return invalid("compute_imgs: more than 9999 patches per image cannot be encoded (img*10000+patch)");
    if (!x_scale || !y_scale || !average_point || !if_nomatching || !left || !right || !x_scale_new || !y_scale_new ||
        !average_new || !count)
        return invalid("compute_imgs: null pointer");
    if (capacity > 0 && (!new_left || !new_right || !bound5)) return invalid("compute_imgs: null output pointer");
    if (((size_t)ps * 3 * elem) % 16 != 0 || ((uintptr_t)left | (uintptr_t)new_left) % 16 != 0)
        return invalid("compute_imgs: patch rows must be 16-byte multiples and images 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int n = height * width, H = ps * height, W = ps * width;
    PATS_CUDA_TRY(launch_chained(compute_imgs_plan_kernel, dim3(1), dim3(1024), 0, st, x_scale, y_scale, average_point, if_nomatching, B, n, height,
                                 width, ps, margin, bound5, x_scale_new, y_scale_new, average_new, capacity, count));
    const int grid = capacity < B * n ? capacity : B * n;
    if (grid == 0) return PATS_OK;
    const int seg_v = (int)((size_t)ps * 3 * elem / 16);
    PATS_CUDA_TRY(launch_chained(left_windows_kernel<uint4>, dim3(grid), dim3(256), 0, st, (const uint4 *)left, (uint4 *)new_left, bound5, count, H, W,
                                 ps, width, seg_v));
#ifndef RP_BANDS
#define RP_BANDS 8
#endif
    const dim3 rgrid(grid, RP_BANDS);
    const int ow_ = 3 * ps, rthreads = ow_ >= 256 ? 256 : (256 / ow_) * ow_;  // whole output rows per CTA pass (96 -> 192 threads)
    if (elem == 1)
        PATS_CUDA_TRY(launch_chained(right_patches_kernel<uint8_t>, rgrid, dim3(rthreads), 0, st, (const uint8_t *)right, new_right, bound5, count, B, H, W,
                                     margin, 3 * ps, 3 * ps, bad_rows));
    else
        PATS_CUDA_TRY(launch_chained(right_patches_kernel<float>, rgrid, dim3(rthreads), 0, st, (const float *)right, new_right, bound5, count, B, H, W,
                                     margin, 3 * ps, 3 * ps, bad_rows));
    return PATS_OK;
}
