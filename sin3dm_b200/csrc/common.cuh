// Shared device-side types and helpers for the triplane denoising kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace s3d {

// Planes: 0 = xy [H,W], 1 = xz [H,D], 2 = yz [W,D]  (reference src/utils/triplane_util.py:20-25).
struct TriDims {
    int rows[3];
    int cols[3];
};
struct TriF {
    float* p[3];
};
struct TriCF {
    const float* p[3];
};
struct TriH {
    __half* p[3];
};
struct TriCH {
    const __half* p[3];
};

constexpr int kGroups = 32;        // GroupNorm32(32, C), reference src/diffusion/nn.py:93-100
constexpr float kGnEps = 1e-5f;    // torch.nn.GroupNorm default
constexpr float kLoScale = 2048.f; // lo = (v - fp16(v)) * 2^11 keeps the residual in fp16 normal range

// Programmatic dependent launch (PDL): every kernel of the step is launched with programmaticStreamSerialization, so
// its CTAs may become resident while the previous kernel is still draining.  pdl_wait() blocks until the previous grid
// has completed and its writes are visible; nothing that reads or writes activations may precede it.  pdl_trigger()
// lets the NEXT kernel start launching (it will block in its own pdl_wait()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float silu_f(float v) {
    // x * sigmoid(x), reference src/diffusion/nn.py:12-14.  ex2.approx + rcp: ~1e-6 relative.
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + __expf(-v)));     // 1 + e^-v >= 1: no range handling needed
    return v * r;
}

// fp32 -> (hi, lo) fp16 pair with v ~= hi + lo / 2048, |err| <= 2^-22 |v|.  Saturating: the conversions are
// cvt.rn.satfinite (F2FP.SATFINITE.F16.F32.PACK_AB converts, saturates and packs two values in one instruction — the explicit
// min/max clamp in front of a plain conversion was a tenth of the element-wise kernels' instructions, and they are issue bound).
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo_elem, float hi_elem) {     // -> {hi_elem : lo_elem} as one 32-bit word
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// two values at once: hi2 / lo2 hold (a, b) in (low, high) halves
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
    hi2 = cvt_f16x2_sat(a, b);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
    // packed fp32x2 arithmetic (sm_100): one FADD2 + one FMUL2 for the pair; both steps are exact in fp32
    const float2 d = __fmul2_rn(__fadd2_rn(make_float2(a, b), make_float2(-hf.x, -hf.y)), make_float2(kLoScale, kLoScale));
    lo2 = cvt_f16x2_sat(d.x, d.y);
}
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
    uint32_t h2, l2;
    split_f16x2(v, 0.f, h2, l2);
    hi = __ushort_as_half(static_cast<unsigned short>(h2 & 0xffffu));
    lo = __ushort_as_half(static_cast<unsigned short>(l2 & 0xffffu));
}

__device__ __forceinline__ void store_split4(__half* hi_ptr, __half* lo_ptr, float4 v) {
    uint2 h, l;
    split_f16x2(v.x, v.y, h.x, l.x);
    split_f16x2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi_ptr) = h;
    *reinterpret_cast<uint2*>(lo_ptr) = l;
}

// Opt-in device trace (s3d_unet_trace_enable): one thread per CTA stamps %globaltimer (ns, chip-wide) and clock64 (SM
// cycles) at named points of a kernel; tools/trace_step.py turns the stamps into a timeline of one step.
constexpr int kTraceSlots = 32;
struct Trace {
    unsigned long long* buf;   // [max_ctas][2][kTraceSlots]; nullptr: tracing off
    int max_ctas;
    int lt0;                   // k_conv_tc: the three local tiles lt0 .. lt0+2 of each CTA stamp their phases (S3D_TRACE_LT0)
};
__device__ __forceinline__ void trace_mark(const Trace& T, int slot) {
    if (T.buf) {
        const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (cta < static_cast<unsigned>(T.max_ctas)) {
            unsigned long long g;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
            T.buf[static_cast<size_t>(cta) * 2 * kTraceSlots + slot] = g;
            T.buf[static_cast<size_t>(cta) * 2 * kTraceSlots + kTraceSlots + slot] = static_cast<unsigned long long>(clock64());
        }
    }
}

// (hi, lo) pair -> fp32, exactly what a reader of the pair sees: hi + lo / 2048
__device__ __forceinline__ float4 join_halves4(uint2 hi, uint2 lo) {
    const __half2 h0 = *reinterpret_cast<const __half2*>(&hi.x), h1 = *reinterpret_cast<const __half2*>(&hi.y);
    const __half2 l0 = *reinterpret_cast<const __half2*>(&lo.x), l1 = *reinterpret_cast<const __half2*>(&lo.y);
    const float2 a = __half22float2(h0), b = __half22float2(h1), c = __half22float2(l0), d = __half22float2(l1);
    return make_float4(fmaf(c.x, 1.f / kLoScale, a.x), fmaf(c.y, 1.f / kLoScale, a.y), fmaf(d.x, 1.f / kLoScale, b.x),
                       fmaf(d.y, 1.f / kLoScale, b.y));
}
// value of v after a round trip through store_split4 (no memory access: recomputed from the same roundings)
__device__ __forceinline__ float4 roundtrip_split4(float4 v) {
    uint2 h, l;
    split_f16x2(v.x, v.y, h.x, l.x);
    split_f16x2(v.z, v.w, h.y, l.y);
    return join_halves4(h, l);
}

// Composed-layout address of plane pixel (r, c): [H+D, W+D] with yz stored transposed.
__device__ __forceinline__ int composed_offset(int plane, int r, int c, int H, int W, int Wc) {
    if (plane == 0) return r * Wc + c;
    if (plane == 1) return r * Wc + W + c;
    return (H + c) * Wc + r;
}

// ---------------------------------------------------------------- Philox4x32-10 (oracle/philox_ref.py)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// 4 standard normals for counter (idx4, step, sample, stream)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t sample, uint32_t step, uint32_t idx4) {
    uint4 r = philox4x32_10(make_uint4(idx4, step, sample, 0u),
                            make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const float s = 2.3283064365386963e-10f;   // 2^-32
    float u0 = fminf(__fmul_rn(__fadd_rn(__uint2float_rn(r.x), 1.f), s), 1.f);
    float u1 = fminf(__fmul_rn(__fadd_rn(__uint2float_rn(r.y), 1.f), s), 1.f);
    float u2 = fminf(__fmul_rn(__fadd_rn(__uint2float_rn(r.z), 1.f), s), 1.f);
    float u3 = fminf(__fmul_rn(__fadd_rn(__uint2float_rn(r.w), 1.f), s), 1.f);
    float rad0 = sqrtf(-2.f * logf(u0)), rad1 = sqrtf(-2.f * logf(u2));
    const float two_pi = 6.283185307179586f;
    float s0, c0, s1, c1;
    sincosf(two_pi * u1, &s0, &c0);
    sincosf(two_pi * u3, &s1, &c1);
    return make_float4(rad0 * c0, rad0 * s0, rad1 * c1, rad1 * s1);
}

}  // namespace s3d
