// Crop geometry and OpenCV-compatible fixed-point bilinear taps, shared by the crop and head kernels.
// Follows lib/train/data/processing_utils.py:30-38,67 (float64 scalar geometry, Python round() =
// half-to-even) and OpenCV's cv::resize INTER_LINEAR 8U path (11-bit coefficients) - SURVEY 8a-R3.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vt {

struct CropGeom {
    int crop_sz;        // side of the square crop in frame pixels
    int x1, y1;         // top-left of the crop in frame coordinates (may be negative)
    int status;         // VT_TRACK_*
    double resize_factor;
};

// All arithmetic is explicit round-to-nearest double so that no FMA contraction can change it.
__device__ __forceinline__ CropGeom crop_geometry(double x, double y, double w, double h, double factor,
                                                  int S, int H, int W) {
    CropGeom g;
    g.status = 0;
    double area = __dmul_rn(w, h);
    double side = ceil(__dmul_rn(sqrt(area), factor));          // math.ceil(math.sqrt(w*h)*factor)
    if (!(side >= 1.0) || !(side < 1.0e9)) {                    // crop_sz < 1 (or NaN / absurd): 'Too small bounding box.'
        g.crop_sz = 1; g.x1 = 0; g.y1 = 0; g.status = 1; g.resize_factor = 0.0;
        return g;
    }
    g.crop_sz = (int)side;
    double half = __dmul_rn(side, 0.5);
    // round(x + 0.5*w - crop_sz*0.5): left-to-right evaluation, half-to-even
    double fx = __dsub_rn(__dadd_rn(x, __dmul_rn(0.5, w)), half);
    double fy = __dsub_rn(__dadd_rn(y, __dmul_rn(0.5, h)), half);
    fx = fmin(fmax(fx, -1.0e9), 1.0e9);
    fy = fmin(fmax(fy, -1.0e9), 1.0e9);
    g.x1 = (int)rint(fx);
    g.y1 = (int)rint(fy);
    g.resize_factor = __ddiv_rn((double)S, side);
    // the reference's slice im[y1+y1_pad : y2-y2_pad, x1+x1_pad : x2-x2_pad] must be non-empty
    long long xa = g.x1 > 0 ? g.x1 : 0, xb = (long long)g.x1 + g.crop_sz;
    long long ya = g.y1 > 0 ? g.y1 : 0, yb = (long long)g.y1 + g.crop_sz;
    if (xb > W - 1) xb = W - 1;
    if (yb > H - 1) yb = H - 1;
    if (!(xa < xb && ya < yb)) g.status = 2;
    return g;
}

// scale = 1.0 / (S / crop_sz) exactly as cv::resize computes it.
__device__ __forceinline__ double resize_scale(int S, int crop_sz) {
    return __ddiv_rn(1.0, __ddiv_rn((double)S, (double)crop_sz));
}

// Source index and fractional part for destination index d: fx = float((d + 0.5) * scale - 0.5).
__device__ __forceinline__ void resize_src(int d, double scale, int& s, float& f) {
    double fd = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    float ff = __double2float_rn(fd);
    float fl = floorf(ff);
    s = (int)fl;
    f = __fsub_rn(ff, fl);
}

// Horizontal taps (clamped, cv::resize xofs/ialpha).
__device__ __forceinline__ void tap_x(int d, double scale, int src, int& s0, int& s1, int& a0, int& a1,
                                      bool& w0nz, bool& w1nz) {
    int s; float f;
    resize_src(d, scale, s, f);
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
    float g = __fsub_rn(1.f, f);
    a0 = __float2int_rn(__fmul_rn(g, 2048.f));
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    s0 = s;
    s1 = min(s + 1, src - 1);
    w0nz = g != 0.f;
    w1nz = f != 0.f;
}

// Vertical taps: coefficients from the UNCLAMPED fraction, rows clamped (cv::resize yofs/ibeta +
// resizeGeneric_Invoker's row clipping).  The mask uses the clamped fraction (see oracle att_mask_spec).
__device__ __forceinline__ void tap_y(int d, double scale, int src, int& r0, int& r1, int& b0, int& b1,
                                      bool& w0nz, bool& w1nz) {
    int s; float f;
    resize_src(d, scale, s, f);
    b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    b1 = __float2int_rn(__fmul_rn(f, 2048.f));
    r0 = min(max(s, 0), src - 1);
    r1 = min(max(s + 1, 0), src - 1);
    float fc = f;
    if (s < 0) fc = 0.f;
    if (s >= src - 1) fc = 0.f;
    w0nz = __fsub_rn(1.f, fc) != 0.f;
    w1nz = fc != 0.f;
}

}  // namespace vt
